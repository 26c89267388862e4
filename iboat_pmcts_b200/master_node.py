"""Drop-in mirror of ``code/solver/master_node.py`` (reference): ``MasterNode`` and ``deepcopy_dict``.

A master node stores ``rewards[scenario][action]`` histograms -- one slice per scenario of the replicated
table the multi-GPU search all-gathers into.
"""
import numpy as np

from .simulatorTLKT import A_DICT, ACTIONS
from .utils import Hist


class MasterNode:
    """Node of the master tree (master_node.py:10-63): ``hash`` (key in the dictionary), ``arm`` (action
    from its parent), ``parentNode``, ``rewards`` (array of ``Hist``, shape scenarios x actions),
    ``children``, ``depth``."""

    def __init__(self, numscenarios, nodehash=None, parentNode=None, action=None, rewards=[]):
        self.hash = nodehash
        self.arm = action
        self.parentNode = parentNode
        self.children = []
        self.depth = None
        self.guessed_rewards = dict()
        if len(rewards) == 0:
            self._table = np.zeros((numscenarios, len(ACTIONS), Hist.N_BINS), dtype=int)
        else:
            self._table = np.array([[np.asarray(h.h) for h in row] for row in rewards], dtype=int)
        self._rewards = None

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_rewards"] = None  # views of _table: rebuilt on first use after unpickling
        return d

    @property
    def rewards(self):
        """Array of ``Hist`` (scenarios x actions), views of this node's table (built on first use)."""
        if self._rewards is None:
            self._rewards = np.array([[Hist(store=self._table[s, a]) for a in range(self._table.shape[1])]
                                      for s in range(self._table.shape[0])])
        return self._rewards

    def add_reward(self, idscenario, reward):
        """Adds a reward to a uniformly random action slot of one scenario (master_node.py:34-44)."""
        action = np.random.randint(len(ACTIONS))
        self.rewards[idscenario, action].add(reward)

    def add_reward_action(self, idscenario, action, reward):
        self.rewards[idscenario, A_DICT[action]].add(reward)

    def add_bins(self, idscenario, slot, bins):
        """Batched form of the two methods above: adds a whole per-leaf histogram of rewards to action
        slot ``slot`` (what ``pmcts_backup`` does on the device table)."""
        self._table[idscenario, slot] += np.asarray(bins, dtype=self._table.dtype)

    def is_expanded(self, idscenario):
        return bool(self._table[idscenario].any())


def deepcopy_dict(nodes):
    """Deep copy of the dictionary the search returns, rewired with ``parentNode``, ``children`` and
    ``depth`` (master_node.py:66-109).  Empties ``nodes`` as the reference does."""
    root_key = hash(tuple([]))
    n = nodes[root_key]._table.shape[0]
    new_dict = dict()
    for k, node in list(nodes.items()):
        # (the copy takes the node's counter table as one array; going through the Hist views the reference
        # copies -- 80 objects per node -- made this 25 times slower than the search it follows)
        new_node = MasterNode.__new__(MasterNode)
        new_node.hash = node.hash
        new_node.arm = node.arm
        new_node.parentNode = None
        new_node.children = []
        new_node.depth = None
        new_node.guessed_rewards = dict()
        new_node._table = np.array(node._table, dtype=int)
        new_node._rewards = None
        new_node.parentHash = node.parentNode.hash if node.parentNode else None
        new_dict[k] = new_node
        del nodes[k]
    for node in new_dict.values():
        if not node.parentHash:
            node.parentNode = None
        else:
            node.parentNode = new_dict[node.parentHash]
            node.parentNode.children.append(node)
    root = new_dict[root_key]
    root.depth = 0
    queue = [root]
    while queue:
        node = queue.pop(0)
        for child in node.children:
            child.depth = node.depth + 1
            queue.append(child)
    return new_dict
