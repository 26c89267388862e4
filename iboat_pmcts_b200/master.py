"""Drop-in mirror of ``code/solver/master.py`` (reference): ``MasterTree`` post-processing of a search --
best policy per scenario and globally, utilities, reward guesses for unexplored (node, scenario) pairs,
pickle save / load.  Plotting methods are out of scope.

The reference walks ``MasterNode`` objects and their ``Hist`` arrays in Python loops (master.py:57-185).
Here the whole tree is reduced at once on the GPU (``pmcts_master_policy``, csrc/pmcts_master.cu): for every
node, every scenario and the probability-weighted global tree -- the two utility terms, the guessed rewards,
the best child, and the greedy policies.  The methods of the reference's API read those arrays; node objects
are only built for what the caller looks at (the nodes along the best policies).
"""
import math
import pickle

import numpy as np

from .master_node import MasterDict, MasterNode
from .simulatorTLKT import ACTIONS
from .utils import out_of_scope
from .worker import UCT_COEFF  # bound at import, like the reference (master.py:11; quirk Q4)

ROOT_HASH = hash(tuple([]))


class MasterTree:
    """Final result of a parallel search: ``nodes`` (dictionary of ``MasterNode`` from
    ``master_node.deepcopy_dict``), ``probability`` of each scenario, ``best_policy`` / ``best_nodes_policy``
    per scenario id (``-1`` = global)."""

    def __init__(self, sims, destination, nodes=None, proba=[]):
        num_scenarios = len(sims)
        self.nodes = dict() if nodes is None else nodes
        if len(self.nodes) == 0:
            self.nodes[ROOT_HASH] = MasterNode(num_scenarios, nodehash=ROOT_HASH)
        self.Simulators = sims
        self.numScenarios = num_scenarios
        self.destination = destination
        self.best_policy = dict()
        self.best_nodes_policy = dict()
        if len(proba) != num_scenarios:
            self.probability = np.array([1 / num_scenarios for _ in range(num_scenarios)])
        else:
            self.probability = np.array(proba)
        self._solved = None

    # ---- the tree as arrays, reduced on the device -----------------------------------------------------
    def _arrays(self):
        """(parent, arm, row_index, rows, node_of_index): the layout ``pmcts_master_policy`` reads.  Straight from
        the native search when ``nodes`` is its array-backed dictionary; otherwise gathered from the node objects
        (children ranked in dictionary order, as ``deepcopy_dict`` links them)."""
        nodes = self.nodes
        if isinstance(nodes, MasterDict):
            sparse = nodes.sparse()
            if sparse is not None:
                return sparse + (nodes.node_at,)
        order, index = [], {}
        for node in nodes.values():  # parents first: a child is created after its parent in every search
            order.append(node)
        order.sort(key=lambda nd: _depth_of(nd))  # stable: keeps dictionary order within a depth
        for i, node in enumerate(order):
            index[id(node)] = i
            node._index = i
        n, S = len(order), self.numScenarios
        parent = np.array([index[id(nd.parentNode)] if nd.parentNode is not None else -1 for nd in order], np.int32)
        arm = np.array([list(ACTIONS).index(nd.arm) if nd.parentNode is not None else 0 for nd in order], np.int8)
        table = np.stack([np.asarray(nd._table) for nd in order]).astype(np.uint32).reshape(n * S, 96)
        used = table.any(axis=1)
        row_index = np.where(used, np.cumsum(used) - 1, -1).astype(np.int32).reshape(n, S)
        return parent, arm, row_index, np.ascontiguousarray(table[used]), order.__getitem__

    def _solve(self, want=("value", "explore", "guess", "best_child", "policy", "policy_nodes")):
        from .engine import default_engine
        eng = default_engine()
        max_depth = max(1, len(self.Simulators[0].times) - 1) if len(self.Simulators) else 64
        nodes = self.nodes
        native = getattr(nodes, "native", None) if isinstance(nodes, MasterDict) and not nodes._dirty else None
        if native is not None and native.handle:
            # the result of a native search, still in the host runtime: reduced from there, no array leaves it
            out = eng.forest_master_policy(native, self.probability, UCT_COEFF, max_depth, want=want)
            out["node_of"] = nodes.node_at
            return out
        parent, arm, row_index, rows, node_of = self._arrays()
        out = eng.master_policy(parent, arm, row_index, rows, self.probability, UCT_COEFF, max_depth, want=want)
        out["node_of"] = node_of
        return out

    def _full(self):
        """All per-node arrays (the API calls that look at single nodes read them); computed once."""
        if self._solved is None or "value" not in self._solved:
            self._solved = self._solve()
        return self._solved

    # ---- reference API ---------------------------------------------------------------------------------
    def get_best_policy(self, verbose=True):
        """Greedy descent from the root for every scenario and for the global tree (master.py:57-84): the
        device follows ``get_best_child`` for all of them at once; only the policies come back."""
        sol = self._solve(want=("policy", "policy_nodes", "value") if verbose else ("policy", "policy_nodes"))
        node_of = sol["node_of"]
        for id_scenario in range(-1, len(self.Simulators)):
            row = id_scenario + 1
            idx = [int(i) for i in sol["policy_nodes"][row] if i >= 0]
            policy = [ACTIONS[a] for a in sol["policy"][row] if a >= 0]
            if verbose:
                print("Global policy" if id_scenario == -1 else "Policy for scenario {}".format(id_scenario))
                col = self.numScenarios if id_scenario == -1 else id_scenario
                for depth, (i, action) in enumerate(zip(idx[1:], policy), 1):
                    print("Depth {}: best reward = {} for action = {}".format(depth, sol["value"][i, col], action))
            self.best_policy[id_scenario] = policy
            if isinstance(self.nodes, MasterDict) and not self.nodes._dirty and not self.nodes._filled:
                self.best_nodes_policy[id_scenario] = self.nodes.nodes_along(idx, policy)
            else:
                self.best_nodes_policy[id_scenario] = [node_of(i) for i in idx]

    def _col(self, idscenario):
        return self.numScenarios if idscenario == -1 else int(idscenario)

    def _index(self, node):
        self._full()  # (gathering the arrays numbers the nodes of a plain dictionary)
        return node._index

    def get_best_child(self, node, idscenario=-1, verbose=False):
        """Child with the largest exploitation value (first maximum wins; values must be > 0); ``(None, None)``
        when the scenario expanded none of the children (master.py:86-115)."""
        sol = self._full()
        c = int(sol["best_child"][self._index(node), self._col(idscenario)])
        if c < 0:
            return None, None
        child = sol["node_of"](c)
        if verbose:
            print("Depth {}: best reward = {} for action = {}".format(child.depth, sol["value"][c, self._col(idscenario)],
                                                                      child.arm))
        return child, child.arm

    def guess_reward(self, node, idscenario):
        """Reward of a node a scenario never explored: the lowest reward for which its nearest explored ancestor
        would have been expanded next (master.py:117-138)."""
        g = float(self._full()["guess"][self._index(node), int(idscenario)])
        if g != g:
            raise AttributeError("'NoneType' object has no attribute 'children'")  # what the reference runs into
        return g

    def get_utility(self, node, idscenario):
        """(exploitation, exploration) of a node for one scenario, or probability-weighted over the scenarios
        for ``idscenario == -1`` with guessed rewards where a scenario did not expand the node
        (master.py:140-185)."""
        sol = self._full()
        i, col = self._index(node), self._col(idscenario)
        if idscenario == -1:  # the reference leaves its guesses on the node
            for s in range(self.numScenarios):
                g = sol["guess"][i, s]
                if g == g:
                    node.guessed_rewards[s] = float(g)
        return float(sol["value"][i, col]), float(sol["explore"][i, col])

    def get_points(self, node, points, probability, coordinate=(0, 0), idscenario=-1, objective="depth"):
        """Segments of the tree below ``node`` for the reference's tree plots (master.py:187-220), parents before
        children, children in their order: ``(x0, y0, x, y, depth)`` or, for ``objective="uct"``,
        ``(x0, y0, x, y, utility, exploitation, exploration)``; a step of unit length in the direction of each arm.
        Subtrees a scenario never expanded are left out.  (The plots themselves are out of scope.)"""
        stack = [(child, coordinate[0], coordinate[1]) for child in reversed(node.children)]
        while stack:  # depth first, children in their order: a child's segment, its whole subtree, then its sibling
            child, x0, y0 = stack.pop()
            if idscenario != -1 and not child.is_expanded(idscenario):
                continue
            x = x0 + math.sin(child.arm * math.pi / 180)
            y = y0 + math.cos(child.arm * math.pi / 180)
            if objective == "uct":
                value, exploration = self.get_utility(child, idscenario)
                points.append((x0, y0, x, y, value + exploration, value, exploration))
            elif objective == "depth":
                points.append((x0, y0, x, y, child.depth))
            stack.extend((grandchild, x, y) for grandchild in reversed(child.children))
        return points

    def get_depth(self):
        if isinstance(self.nodes, MasterDict) and self.nodes.arrays() is not None:
            return int(self.nodes.depths().max())
        return max(node.depth for node in self.nodes.values())

    def save_tree(self, name, directory="../results/", reference_format=False):
        """``reference_format=True`` writes the file the reference's own ``MasterTree.load_tree`` and plotting tools
        read (its class names and attribute layout: ``refpickle.dump_reference_tree``); the default is a pickle of
        this class."""
        with open(directory + name + '.pickle', 'wb') as f:
            if reference_format:
                from .refpickle import dump_reference_tree
                dump_reference_tree(self, f)
            else:
                pickle.dump(self, f)

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_solved"] = None
        return d

    @classmethod
    def load_tree(cls, name, directory="../results/"):
        """A tree saved by this class, or by the REFERENCE's ``MasterTree.save_tree`` (master.py:390-410): its nodes,
        histograms, simulators and weather grids come back as this package's classes (``refpickle.load_reference_tree``)."""
        from .refpickle import load_reference_tree
        with open(directory + name + '.pickle', 'rb') as f:
            return load_reference_tree(f)


def _depth_of(node):
    d = 0
    while node.parentNode is not None:
        node = node.parentNode
        d += 1
    return d


for _name in ("plot_tree", "plot_tree_uct", "plot_best_policy", "plot_hist_best_policy", "draw_points"):
    setattr(MasterTree, _name, staticmethod(out_of_scope("MasterTree." + _name)))
