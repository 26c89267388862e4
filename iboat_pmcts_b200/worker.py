"""Drop-in mirror of ``code/solver/worker.py`` (reference): ``Node`` and the per-scenario UCT ``Tree``.

Selection and expansion stay on the host, as in the reference; the default-policy rollout
(``Tree.default_policy``, worker.py:255-281) runs on the GPU.  Two ways to search:

* ``Tree.uct_search`` -- the reference's sequential loop, one rollout per iteration.  The rollout is one
  literal-fp64 launch whose noise is taken from Python's ``random.gauss`` stream in the reference's
  order, and ``back_up`` / ``integrate_buffer`` consume ``np.random.randint`` as the reference does: a
  seeded search reproduces the reference's tree, visit count by visit count (tests/test_host_tree.py, tests/test_api_gpu.py).
* ``Tree.select_leaves`` / ``Tree.back_up_leaves`` -- the leaf-batched form used by ``forest.Forest``:
  ``B`` leaves per wave (virtual visits keep them apart), ``K`` noise replicas per leaf, per-leaf reward
  histograms reduced in the rollout kernel and added along the ancestor chains.

Every node's 8 x 12 counters live in one table per tree (``Tree.table[node][action][bin]``);
``node.Values[i].h`` is a view of its row.
"""
import math
import random as rand
from math import asin, exp, log, sqrt

import numpy as np

from . import _native as N
from . import simulatorTLKT as SimC
from .master_node import MasterNode
from .simulatorTLKT import A_DICT, ACTIONS  # noqa: F401
from .utils import Hist

#: Exploration coefficient of the UCT (worker.py:21) and share of the master's value in a child's
#: score (worker.py:24).  Read at call time.
UCT_COEFF = 1 / 5 * 1 / 2 ** 0.5
RHO = 0.3


class Node:
    """Node of a worker ``Tree`` (worker.py:27-104)."""

    def __init__(self, state=None, parent=None, origins=[], children=[], depth=0, store=None, rng=None):
        self.state = tuple(state) if state is not None else None  # root only
        self.parent = parent
        self.origins = list(origins)
        self.children = list(children)
        self.actions = list(SimC.ACTIONS)
        (rng or rand).shuffle(self.actions)  # the reference shuffles with the global generator (worker.py:51)
        self._table = store if store is not None else np.zeros((len(SimC.ACTIONS), Hist.N_BINS), dtype=int)
        self._values = None
        self.depth = depth
        self.pending = 0  # virtual visits of the wave being assembled (leaf-batched search)

    @property
    def Values(self):
        """Array of 8 ``Hist``, views of this node's rows of the tree table (built on first use)."""
        if self._values is None:
            self._values = np.array([Hist(store=self._table[i]) for i in range(self._table.shape[0])])
        return self._values

    def _rebind(self, rows):
        self._table = rows
        if self._values is not None:
            for a, hist in enumerate(self._values):
                hist.h = rows[a]

    def back_up(self, reward):
        """Adds the reward to a random action slot of this node (worker.py:62-63) and, up to the root, to
        the slot of the action taken from each ancestor."""
        action = np.random.randint(len(ACTIONS))
        self.Values[action].add(reward)
        node = self
        while node.parent:
            node.parent.Values[SimC.A_DICT[node.origins[-1]]].add(reward)
            node = node.parent

    def back_up_bins(self, slot, bins):
        """Batched ``back_up``: a whole histogram of rewards (the K replicas of one leaf) at once."""
        bins = np.asarray(bins, dtype=self._table.dtype)
        self._table[slot] += bins
        node = self
        while node.parent:
            node.parent._table[SimC.A_DICT[node.origins[-1]]] += bins
            node = node.parent

    def is_fully_expanded(self):
        return len(self.actions) == 0

    def visits(self):
        num = 0
        for val in self.Values:
            num += sum(val.h)
        return num

    def get_uct(self, num_parent):
        """max over actions of the binned mean + UCT_COEFF * sqrt(2 ln(N_parent) / N_node) (worker.py:82-104)."""
        best = 0
        num_node = self.visits()
        exploration = UCT_COEFF * (2 * math.log(num_parent) / num_node) ** 0.5
        for hist in self.Values:
            value = hist.get_mean()
            if value > best:
                best = value
        return best + exploration


def memo_counts(memo, node):
    """(visit count, max over actions of the binned mean) of a node, once per wave."""
    got = memo.get(id(node))
    if got is None:
        t = node._table
        tot = t.sum(1)
        visits = int(tot.sum())
        acc16 = t @ (2 * np.arange(Hist.N_BINS) + 1)
        value = 0.0
        for a in range(len(tot)):
            if tot[a] > 0:
                value = max(value, (float(acc16[a]) * 0.0625) / float(tot[a]))
        got = memo[id(node)] = (visits, value)
    return got


class Tree:
    """MCTS-UCT on one weather scenario (worker.py:107-445)."""

    def __init__(self, workerid, nscenario, probability=[], ite=0, budget=1000,
                 simulator=None, destination=[], TimeMin=0, buffer=None):
        self.id = workerid
        self.ite = ite
        self.budget = budget
        self.Simulator = simulator
        self.destination = destination
        self.TimeMax = self.Simulator.times[-1]
        self.TimeMin = TimeMin
        self.depth = 0
        self.Nodes = []
        self.buffer = [] if buffer is None else buffer  # (the reference shares one default list: quirk Q9)
        self.numScenarios = nscenario
        if len(probability) == 0:
            self.probability = np.array([1 / nscenario for _ in range(nscenario)])
        else:
            self.probability = probability
        self.rng = None  # generator of the nodes' action order; None = Python's global one, as in the reference
        self.vectorized = False  # leaf-batched search: NumPy forms of the UCT sums (same values up to rounding)
        self.replay_full_prefix = False  # True fixes quirk Q1 (worker.py:297 never replays past the first action)
        self._cap = 0
        self.table = np.zeros((0, len(SimC.ACTIONS), Hist.N_BINS), dtype=int)

    def __getattr__(self, name):
        # Only reached when the attribute is missing: after a native batched search (forest._export_tree) the node
        # objects are built on first use -- a 10^4-node tree costs 50 ms of interpreter time nobody may ask for.
        if name in ("Nodes", "rootNode", "depth", "table", "_cap") and "_lazy_nodes" in self.__dict__:
            from .forest import materialize_tree
            materialize_tree(self)
            return self.__dict__[name]
        raise AttributeError(name)

    # ---- storage ---------------------------------------------------------------------------------
    def _new_row(self):
        """Row of the tree's table for the next node; the table grows by doubling (views of earlier rows are
        re-pointed, so ``node.Values[i].h`` stays live)."""
        idx = len(self.Nodes)
        if idx >= self._cap:
            cap = max(256, 2 * self._cap)
            table = np.zeros((cap,) + self.table.shape[1:], dtype=int)
            table[:self._cap] = self.table[:self._cap]
            self.table, self._cap = table, cap
            for i, node in enumerate(self.Nodes):
                node._rebind(table[i])
        return self.table[idx]

    def _make_root(self, rootState):
        rootNode = Node(state=rootState, store=self._new_row(), rng=self.rng)
        self.rootNode = rootNode
        self.Nodes.append(rootNode)
        return rootNode

    # ---- the reference's sequential search ---------------------------------------------------------
    def uct_search(self, rootState, frequency, Master_nodes):
        """worker.py:147-186: ``budget`` iterations of select / expand -> rollout -> back-up, pushing the
        buffered updates to ``Master_nodes`` every ``frequency`` iterations."""
        self._make_root(rootState)
        while self.ite < self.budget:
            leafNode = self.tree_policy(self.rootNode, Master_nodes)
            reward = self.default_policy(leafNode)
            leafNode.back_up(reward)
            self.buffer.append([self.id, hash(tuple(leafNode.origins)), hash(tuple(leafNode.origins[:-1])),
                                leafNode.origins[-1], reward])
            self.ite = self.ite + 1
            if self.ite % 50 == 0:
                print('Iteration ' + str(self.ite) + ' on ' + str(self.budget) + ' for workers ' + str(self.id) + '\n')
            if self.ite % frequency == 0:
                self.integrate_buffer(Master_nodes)
                self.buffer = []

    def tree_policy(self, node, master_nodes):
        while not self.is_node_terminal(node):
            if not node.is_fully_expanded():
                return self.expand(node)
            node = self.best_child(node, master_nodes)
        return node

    def expand(self, node):
        row = self._new_row()  # before Node(): growing the table re-points existing views
        action = node.actions.pop()
        newNode = Node(parent=node, origins=node.origins + [action], depth=node.depth + 1, store=row, rng=self.rng)
        self.depth = max(self.depth, newNode.depth)
        node.children.append(newNode)
        self.Nodes.append(newNode)
        return newNode

    def best_child(self, node, Master_nodes):
        """Child with the largest score: its own UCT, mixed with the master's when the master knows the
        node (worker.py:224-253); first maximum wins."""
        best_score, best_id = -1, -1
        if self.vectorized:
            return self._best_child_vectorized(node, Master_nodes)
        num_node = node.visits() + node.pending
        for i, child in enumerate(node.children):
            uct_master = self.get_master_uct(hash(tuple(child.origins)), Master_nodes)
            own = self._child_uct(child, num_node)
            score = own if uct_master == 0 else (1 - RHO) * own + RHO * uct_master
            if score > best_score:
                best_score, best_id = score, i
        return node.children[best_id]

    def _best_child_vectorized(self, node, Master_nodes):
        """best_child for the leaf-batched search.  Within a wave neither the tree's counts nor the master
        change (only the virtual visits do), so the exploitation value and visit count of a node, and the
        master's UCT of a node, are computed once per wave (NumPy on the tables) and reused."""
        memo = self._wave_memo
        num_parent = memo_counts(memo, node)[0] + node.pending
        log_parent = 2 * math.log(num_parent)
        best_score, best_id = -1, -1
        for i, child in enumerate(node.children):
            visits, value = memo_counts(memo, child)
            own = value + UCT_COEFF * (log_parent / float(visits + child.pending)) ** 0.5
            key = child.__dict__.get("_hash")
            if key is None:
                key = child._hash = hash(tuple(child.origins))
            uct_master = memo.get(key)
            if uct_master is None:
                uct_master = memo[key] = self._master_uct_vectorized(Master_nodes.get(key, 0))
            score = own if uct_master == 0 else (1 - RHO) * own + RHO * uct_master
            if score > best_score:
                best_score, best_id = score, i
        return node.children[best_id]

    def _master_uct_vectorized(self, m):
        """get_master_uct on the master node's table: counts by NumPy, then scalar libm arithmetic summed in
        scenario order (the same operations, in the same order, as csrc/pmcts_forest_host.cpp)."""
        if m == 0:
            return 0
        mt, pt = m._table, m.parentNode._table
        n_node, n_par = mt.sum((1, 2)), pt.sum((1, 2))
        tot = mt.sum(2)
        acc16 = mt @ (2 * np.arange(Hist.N_BINS) + 1)   # sum h[b] (2b + 1): the binned mean is acc16 / 16 / tot
        total, any_ok = 0.0, False
        for s in range(mt.shape[0]):
            if n_node[s] != 0 and n_par[s] != 0:
                best = 0.0
                for a in range(mt.shape[1]):
                    if tot[s, a] > 0:
                        mean = (float(acc16[s, a]) * 0.0625) / float(tot[s, a])
                        if mean > best:
                            best = mean
                uct = best + UCT_COEFF * (2.0 * math.log(float(n_par[s])) / float(n_node[s])) ** 0.5
                total += uct * float(self.probability[s])
                any_ok = True
        return total if any_ok else 0.0

    @staticmethod
    def _child_uct(child, num_parent):
        if child.pending == 0:
            return child.get_uct(num_parent)
        # leaf-batched search: rollouts in flight count as visits in the exploration term only
        best = 0
        for hist in child.Values:
            value = hist.get_mean()
            if value > best:
                best = value
        return best + UCT_COEFF * (2 * math.log(num_parent) / (child.visits() + child.pending)) ** 0.5

    def default_policy(self, node):
        """Reward of a default-policy rollout from the node's estimated state (worker.py:255-281): ONE
        launch of the literal fp64 rollout kernel.  The noise is what ``random.gauss`` would have produced
        for the reference: the draws are taken from the global generator, one per transition made."""
        sim = self.Simulator
        eng, scen = sim.scenario()
        nT = len(sim.times)
        state = rand.getstate()
        z = np.array([[rand.gauss(0.0, 1.0)] for _ in range(nT)], np.float64)
        prec = eng.precision
        eng.set_precision(N.PREC_F64)
        try:
            flags = N.ROLL_REPLAY_FULL_PREFIX if self.replay_full_prefix else 0
            r = eng.rollout_batch_host(scen, 1, 1, self.rootNode.state, self.destination, self.TimeMin,
                                       prefix=np.array([node.origins], np.float64), z=z, traj_steps=nT, flags=flags,
                                       want=("reward", "steps", "status", "final_lat", "final_lon"))
        finally:
            eng.set_precision(prec)
        steps = int(r["steps"][0])
        self.transitions = getattr(self, "transitions", 0) + steps  # (counted for the search report)
        rand.setstate(state)
        for _ in range(steps):  # leave the generator where the reference's doStep calls would
            rand.gauss(0.0, 1.0)
        status = int(r["status"][0])
        if status in (N.ST_OUT_OF_GRID, N.ST_PAST_HORIZON, N.ST_DEGENERATE):
            SimC.raise_for_status(status)
        # the reference leaves its Simulator at the last state of the rollout
        k0 = int(self.rootNode.state[0])
        sim.state = [k0 + steps, float(r["final_lat"][0]), float(r["final_lon"][0])]
        if steps >= 2:
            sim.prevState = [k0 + steps - 1, float(r["traj_lat"][steps - 2, 0]), float(r["traj_lon"][steps - 2, 0])]
        else:
            sim.prevState = list(self.rootNode.state)
        return float(r["reward"][0])

    def get_sim_to_estimate_state(self, node):
        """Brings ``self.Simulator`` to the state a node stands for, with scalar ``doStep`` calls
        (worker.py:283-300).  As in the reference only the first action is replayed: the loop guard negates
        the non-empty list ``is_state_at_dest`` returns (quirk Q1).  ``replay_full_prefix`` fixes that."""
        actions = list(node.origins)
        actions.reverse()
        self.Simulator.reset(self.rootNode.state)
        self.Simulator.doStep(actions.pop())
        while actions and not Tree.is_state_terminal(self.Simulator, self.Simulator.state) and self.replay_full_prefix \
                and not Tree.is_state_at_dest(self.destination, self.Simulator.prevState, self.Simulator.state)[0]:
            self.Simulator.doStep(actions.pop())

    def get_master_uct(self, node_hash, Master_nodes):
        """UCT value of a node as the master sees it: per scenario that visited both the node and its parent,
        weighted by ``probability`` without renormalising (worker.py:302-346)."""
        master_node = Master_nodes.get(node_hash, 0)
        if master_node == 0:
            return 0
        idx_scenarios, uct_per_scenario = [], []
        for s, reward_per_scenario in enumerate(master_node.rewards):
            num_parent = 0
            for hist in master_node.parentNode.rewards[s]:
                num_parent += sum(hist.h)
            num_node = 0
            for hist in reward_per_scenario:
                num_node += sum(hist.h)
            if num_parent != 0 and num_node != 0:
                exploration = UCT_COEFF * (2 * log(num_parent) / num_node) ** 0.5
                best = 0
                for hist in reward_per_scenario:
                    value = hist.get_mean()
                    if value > best:
                        best = value
                uct_per_scenario.append(best + exploration)
                idx_scenarios.append(s)
        return np.dot(uct_per_scenario, [self.probability[i] for i in idx_scenarios])

    def integrate_buffer(self, Master_nodes):
        """Pushes the buffered updates ``[scenario, node hash, parent hash, action, reward]`` to the master
        dictionary and adds each reward along the master node's ancestors (worker.py:348-378)."""
        for scenarioId, newNodeHash, parentHash, action, reward in self.buffer:
            node = Master_nodes.get(newNodeHash, 0)
            if node == 0:
                node = MasterNode(self.numScenarios, nodehash=newNodeHash, parentNode=Master_nodes[parentHash],
                                  action=action)
            node.add_reward(scenarioId, reward)
            Master_nodes[newNodeHash] = node
            while node.parentNode is not None:
                parent_hash = Master_nodes[node.hash].parentNode.hash
                parent_node = Master_nodes[parent_hash]
                parent_node.add_reward_action(scenarioId, node.arm, reward)
                Master_nodes[parent_hash] = parent_node
                node = parent_node

    # ---- leaf-batched search (used by forest.Forest) ---------------------------------------------------
    def select_leaves(self, n_leaves, master_nodes):
        """``n_leaves`` runs of the tree policy without back-up in between.  Each selected leaf leaves a
        virtual visit on its path (exploration term only), so the runs spread over the tree instead of
        piling on one child.  Returns the list of leaf nodes."""
        leaves = []
        self.vectorized = True
        self._wave_memo = {}
        for _ in range(n_leaves):
            leaf = self.tree_policy(self.rootNode, master_nodes)
            node = leaf
            while node is not None:
                node.pending += 1
                node = node.parent
            leaves.append(leaf)
        self.vectorized = False
        return leaves

    def back_up_leaves(self, leaves, leaf_bins, slots):
        """Adds the per-leaf reward histograms of one wave (``leaf_bins[i][12]``, reduced on the GPU) along
        the ancestor chains and clears the virtual visits.  ``slots[i]`` is the random action slot of
        worker.py:62-63, one per leaf."""
        for leaf, bins, slot in zip(leaves, leaf_bins, slots):
            leaf.back_up_bins(int(slot), bins)
            node = leaf
            while node is not None:
                node.pending -= 1
                node = node.parent
        self.ite += len(leaves)

    # ---- statics -----------------------------------------------------------------------------------
    @staticmethod
    def is_state_at_dest(destination, stateA, stateB):
        """Has the boat passed the destination between ``stateA`` and ``stateB``?  (worker.py:380-415):
        the destination must lie within ``DESTINATION_ANGLE`` of the plane through the earth's centre and
        the two positions, and between them.  Returns ``[True, frac]`` or ``[False, None]``."""
        to_cart = SimC.Simulator.fromGeoToCartesian
        xa, ya, za = to_cart(stateA[1:])
        xb, yb, zb = to_cart(stateB[1:])
        xd, yd, zd = to_cart(destination)
        # plane x + b y + c z = 0 through A and B
        c = (yb / ya * xa - xb) / (zb - yb / ya * za)
        b = -(xa + c * za) / ya
        d = abs(xd + b * yd + c * zd) / sqrt(1 + b ** 2 + c ** 2)
        if asin(d) > SimC.DESTINATION_ANGLE:
            return [False, None]
        vad = np.array([xd, yd, zd]) - np.array([xa, ya, za])
        vdb = np.array([xb, yb, zb]) - np.array([xd, yd, zd])
        if np.dot(vad, vdb) < 0:
            return [False, None]
        vab = np.array([xb, yb, zb]) - np.array([xa, ya, za])
        return [True, np.dot(vad, vab) / np.dot(vab, vab)]

    @staticmethod
    def is_state_terminal(simulator, state):
        """Time horizon reached, or boat outside the zone ``lats`` x ``lons`` (worker.py:417-436)."""
        if simulator.times[state[0]] == simulator.times[-1]:
            return True
        if state[1] < simulator.lats[0] or state[1] > simulator.lats[-1]:
            return True
        if state[2] < simulator.lons[0] or state[2] > simulator.lons[-1]:
            return True
        return False

    def is_node_terminal(self, node):
        return self.Simulator.times[node.depth] == self.TimeMax
