"""Drop-in mirror of ``code/solver/forest.py`` (reference): ``Forest``, ``create_simulators`` and
``load_scenarios``, ``initialize_simulators``.  (Scenario download and plotting are out of scope.)

The reference runs one ``multiprocessing.Process`` per scenario around a ``Manager().dict()``
(forest.py:45-81).  Here the scenarios of this process are rolled out together on its GPU
(``pmcts_rollout_forest``: all local scenarios in one launch), scenarios are sharded over the ranks of a
``torch.distributed`` job, and the per-scenario update blocks of a wave are all-gathered so that every
rank applies the same updates, in the same order, to its replica of the master dictionary.
"""
import random as rand
from copy import deepcopy

import os

import numpy as np

from . import _native as N
from . import worker as mt
from .master_node import MasterNode
from .simulatorTLKT import A_DICT, ACTIONS, HOURS_TO_DAY, Simulator
from .utils import out_of_scope

ROOT_HASH = hash(tuple([]))


class Forest:
    """Coordinates the worker trees and the master dictionary of a parallel search (forest.py:14-81)."""

    def __init__(self, listsimulators=[], destination=[], timemin=0, budget=100):
        self.workers = dict()
        nscenario = len(listsimulators)
        self.nscenario = nscenario
        self.destination = list(destination)
        self.timemin = timemin
        for i, sim in enumerate(listsimulators):
            self.workers[i] = mt.Tree(workerid=i, nscenario=nscenario, ite=0, budget=budget,
                                      simulator=deepcopy(sim), destination=deepcopy(destination), TimeMin=timemin)

    # ------------------------------------------------------------------------------------------------
    def launch_search(self, root_state, frequency, mode="batched", leaves=64, replicas=32, seed=0,
                      replay_full_prefix=False, group=None, host="native", pipeline=1):
        """Runs the search and returns the master dictionary ``{hash(tuple(origins)): MasterNode}`` (deep
        copy it with ``master_node.deepcopy_dict`` before use, as with the reference).

        ``mode="sequential"``: every worker runs the reference's ``Tree.uct_search`` (one rollout per
        iteration, ``frequency`` iterations between pushes to the master), one worker after the other --
        what the reference's processes do when they do not interleave; reproducible from
        ``random.seed`` / ``np.random.seed``.  With ``host="native"`` (default) the loop runs in the C++ host
        runtime on the same random streams (``pmcts_forest_search_sequential``); ``host="python"`` runs the Python
        loop of ``worker.Tree.uct_search`` -- same tree, same master, same generator states afterwards
        (tests/test_api_gpu.py).

        ``mode="batched"`` (default): waves of ``leaves`` leaves per tree, ``replicas`` noise replicas per
        leaf (Philox noise from ``seed``), all local scenarios in one launch; the master is updated after
        every wave (the role ``frequency`` plays in the reference).  With ``torch.distributed`` initialised
        (``group``; ``False`` = ignore it), rank ``r`` owns workers ``r * S/world .. (r+1) * S/world - 1`` and the waves' update blocks are
        all-gathered; seeded alike, the ranks return the same dictionary as a single process would.
        ``host="native"`` runs the whole search as one native call (``pmcts_forest_search``): trees and master in
        the C++ host runtime, one launch per wave, update blocks all-gathered on the device; ``Node`` /
        ``MasterNode`` objects are built afterwards, when they are first read.  ``host="python"`` runs the same
        search on the Python classes (same leaves, same counts: tests/test_host_forest.py).
        ``pipeline=2`` (native host) keeps two waves in flight: wave ``w + 1`` is selected on the host while the
        GPU rolls out wave ``w``, whose leaves hold virtual visits until they are backed up -- other trees than
        ``pipeline=1`` (which reproduces the synchronous search), equally reproducible.
        """
        Master_nodes = dict()
        Master_nodes[ROOT_HASH] = MasterNode(len(self.workers), nodehash=ROOT_HASH)
        if mode == "sequential" and host == "native" and self._native_ok():
            return self._search_sequential_native(list(root_state), int(frequency), bool(replay_full_prefix))
        if mode == "sequential":
            for worker in self.workers.values():
                worker.replay_full_prefix = replay_full_prefix
                worker.uct_search(deepcopy(root_state), frequency, Master_nodes)
            self.last_search_report = dict(rollouts=sum(w.ite for w in self.workers.values()),
                                           transitions=sum(getattr(w, "transitions", 0) for w in self.workers.values()))
            return Master_nodes
        if mode != "batched":
            raise ValueError("mode must be 'sequential' or 'batched'")
        if host == "native":
            return self._search_native(list(root_state), int(leaves), int(replicas), int(seed),
                                       bool(replay_full_prefix), group, int(pipeline))
        return self._search_batched(list(root_state), Master_nodes, int(leaves), int(replicas), int(seed),
                                    bool(replay_full_prefix), group)

    # ------------------------------------------------------------------------------------------------
    def _native_ok(self):
        try:
            eng, _ = next(iter(self.workers.values())).Simulator.scenario()
        except Exception:
            return False
        return getattr(eng, "_h", None) is not None and np.random.get_state()[0] == "MT19937"

    def _search_sequential_native(self, root_state, frequency, replay_full):
        """``Tree.uct_search`` of every worker, one after the other (the reference's processes, serialised), as native
        calls (``pmcts_forest_search_sequential``): the same iterations, in the same order, on the same random streams
        as the Python loop of ``worker.Tree.uct_search`` -- ``random`` and ``np.random`` are continued from their
        current states and left where that loop would have left them -- with one small launch per rollout."""
        import ctypes as C
        from . import worker as W
        lib = N.lib()
        S = len(self.workers)
        trees = [self.workers[i] for i in range(S)]
        budget = trees[0].budget
        max_depth = len(trees[0].Simulator.times) - 1
        ids = np.arange(S, dtype=np.int32)
        prob = np.ascontiguousarray(trees[0].probability, np.float64)
        handle = C.c_void_p()
        N.check_forest(lib.pmcts_forest_create(S, S, ids.ctypes.data, max_depth, prob.ctypes.data, float(W.UCT_COEFF),
                                               float(W.RHO), None, budget + 2, C.byref(handle)))
        np_like = np.random.get_state()
        py, npst = N.MtState.from_python(rand.getstate()), N.MtState.from_numpy(np_like)
        root = np.ascontiguousarray(root_state, np.float64)
        dest = np.ascontiguousarray(self.destination, np.float64)
        flags = N.ROLL_REPLAY_FULL_PREFIX if replay_full else 0
        total = dict(rollouts=0, transitions=0, waves=0, launches=0, ms_total=0.0)
        try:
            for i, t in enumerate(trees):
                eng, slot = t.Simulator.scenario()
                report = N.SearchReport()
                rc = lib.pmcts_forest_search_sequential(eng._h, handle, i, int(slot), root.ctypes.data, dest.ctypes.data,
                                                        float(self.timemin), int(t.budget), int(frequency), flags,
                                                        C.byref(py), C.byref(npst), C.byref(report))
                # the generators stand where the reference's would, whatever happened
                rand.setstate(py.to_python())
                np.random.set_state(npst.to_numpy(np_like))
                if rc > 0:
                    from .simulatorTLKT import raise_for_status
                    raise_for_status(rc)
                N.check(rc)
                for k in total:
                    total[k] += getattr(report, k)
        except BaseException:
            lib.pmcts_forest_destroy(handle)
            raise
        self.last_search_report = total
        native = _NativeForest(lib, handle, S)
        for ti, t in enumerate(trees):
            _export_tree(native, ti, t, root_state, t.budget)
        return native.master_dict()

    # ------------------------------------------------------------------------------------------------
    def _search_batched(self, root_state, Master_nodes, B, K, seed, replay_full, group):
        dist, rank, world = _dist_env(group)
        S = len(self.workers)
        if S % world:
            raise ValueError("the number of scenarios (%d) must be a multiple of the world size (%d)" % (S, world))
        S_loc = S // world
        local_ids = list(range(rank * S_loc, (rank + 1) * S_loc))  # contiguous block: Philox stream = scenario id
        trees = [self.workers[i] for i in local_ids]
        for t in trees:
            t.rng = rand.Random(1000003 * seed + t.id)  # per-tree action orders: independent of the sharding
            t._make_root(root_state)
        budget = trees[0].budget
        eng, slots = None, []
        for t in trees:
            e, s = t.Simulator.scenario()
            eng = e
            slots.append(s)
        eng.set_precision(N.PREC_FAST)
        flags = N.ROLL_REPLAY_FULL_PREFIX if replay_full else 0
        wave = 0
        while trees[0].ite < budget:
            b = min(B, budget - trees[0].ite)
            leaves = [t.select_leaves(b, Master_nodes) for t in trees]
            depth = max(1, max(len(l.origins) for ls in leaves for l in ls))
            prefix = np.zeros((S_loc, b, depth), np.float64)
            plen = np.zeros((S_loc, b), np.int32)
            for si, ls in enumerate(leaves):
                for li, leaf in enumerate(ls):
                    plen[si, li] = len(leaf.origins)
                    prefix[si, li, :plen[si, li]] = leaf.origins
            bins = np.zeros((S_loc, b, 12), np.uint32)
            eng.rollout_forest(slots, b, K, root_state, self.destination, self.timemin, prefix.reshape(-1),
                               plen.reshape(-1), depth, prefix_shared=False, seed=seed,
                               stream=wave * S + rank * S_loc, id0=0,
                               out=dict(leaf_bins=bins.reshape(-1)), flags=flags)
            # random action slots of Node.back_up (worker.py:62) and MasterNode.add_reward (master_node.py:42):
            # drawn for ALL scenarios on every rank (ranks seeded alike draw alike), so the search does not
            # depend on the number of ranks
            slot_w = np.random.randint(len(ACTIONS), size=(S, b))[local_ids[0]:local_ids[0] + S_loc]
            slot_m = np.random.randint(len(ACTIONS), size=(S, b))[local_ids[0]:local_ids[0] + S_loc]
            for t, ls, bb, sw in zip(trees, leaves, bins, slot_w):
                t.back_up_leaves(ls, bb, sw)
            # update block of this rank: action paths, slots and histograms of its leaves
            width = len(trees[0].Simulator.times) - 1  # deepest possible leaf (worker.py:438-445)
            paths = np.full((S_loc, b, width), -1, np.int32)
            paths[:, :, :depth] = np.where(np.arange(depth)[None, None, :] < plen[:, :, None], prefix, -1).astype(np.int32)
            block = dict(ids=np.asarray(local_ids, np.int32), paths=paths, slots=slot_m.astype(np.int32), bins=bins)
            blocks = _all_gather_blocks(block, dist, group, world)
            for blk in blocks:  # rank order, then scenario order, then leaf order: identical on every rank
                for si, scen_id in enumerate(blk["ids"]):
                    for li in range(b):
                        path = [a for a in blk["paths"][si, li] if a >= 0]
                        _master_add(Master_nodes, S, int(scen_id), path, int(blk["slots"][si, li]), blk["bins"][si, li])
            wave += 1
        return Master_nodes


    # ------------------------------------------------------------------------------------------------
    def _search_native(self, root_state, B, K, seed, replay_full, group, pipeline=1):
        """The batched search as ONE native call (``pmcts_forest_search``): the trees and the master live in the
        C++ host runtime, every wave's rollouts are one launch over this rank's scenarios, and the update blocks
        (int8 paths + packed per-leaf histograms) are all-gathered on the device when several ranks search
        together."""
        import ctypes as C
        from . import worker as W
        lib = N.lib()
        dist, rank, world = _dist_env(group)
        S = len(self.workers)
        if S % world:
            raise ValueError("the number of scenarios (%d) must be a multiple of the world size (%d)" % (S, world))
        S_loc = S // world
        local_ids = list(range(rank * S_loc, (rank + 1) * S_loc))
        trees = [self.workers[i] for i in local_ids]
        budget = trees[0].budget
        nT = len(trees[0].Simulator.times)
        max_depth = nT - 1  # a node at depth len(times) - 1 is terminal (worker.py:438-445): no fixed path width
        # every node's shuffled action list, drawn tree by tree in creation order with the tree's own generator
        # (what Node.__init__ does, worker.py:50-51)
        max_nodes = budget + 2
        perms = np.empty((S_loc, max_nodes, len(ACTIONS)), np.uint8)
        # = rng = random.Random(1000003 * seed + t.id); rng.shuffle(list(range(8))) per node, generated natively, the
        # trees' tables side by side (tests/test_host_forest.py holds the two together)
        tree_seeds = [1000003 * seed + t.id for t in trees]
        if all(0 <= ts < 2 ** 64 for ts in tree_seeds):
            seeds64 = np.asarray(tree_seeds, np.uint64)
            N.check_forest(lib.pmcts_python_shuffles_batch(seeds64.ctypes.data, S_loc, max_nodes, len(ACTIONS),
                                                           perms.ctypes.data))
        else:  # (seeds beyond 64 bits: the interpreter's own generator)
            for ti, tree_seed in enumerate(tree_seeds):
                rng = rand.Random(tree_seed)
                for j in range(max_nodes):
                    order = list(range(len(ACTIONS)))
                    rng.shuffle(order)
                    perms[ti, j] = order
        ids = np.asarray(local_ids, np.int32)
        prob = np.ascontiguousarray(trees[0].probability, np.float64)
        # random action slots of Node.back_up (worker.py:62) and MasterNode.add_reward (master_node.py:42): drawn
        # wave by wave for ALL scenarios on every rank (ranks seeded alike draw alike), so the search does not
        # depend on the number of ranks
        slot_w, slot_m, done = [], [], 0
        while done < budget:
            b = min(B, budget - done)
            slot_w.append(np.random.randint(len(ACTIONS), size=(S, b)).astype(np.int8).ravel())
            slot_m.append(np.random.randint(len(ACTIONS), size=(S, b)).astype(np.int8).ravel())
            done += b
        slot_w, slot_m = np.concatenate(slot_w), np.concatenate(slot_m)
        handle = C.c_void_p()
        N.check_forest(lib.pmcts_forest_create(S, S_loc, ids.ctypes.data, max_depth, prob.ctypes.data,
                                               float(W.UCT_COEFF), float(W.RHO), perms.ctypes.data, max_nodes,
                                               C.byref(handle)))
        try:
            eng, slots = None, []
            for t in trees:
                eng, s = t.Simulator.scenario()
                slots.append(s)
            slots = np.asarray(slots, np.int32)
            params = N.SearchParams()
            params.n_leaves, params.replicas, params.budget, params.seed = B, K, budget, seed
            params.roll_flags = N.ROLL_REPLAY_FULL_PREFIX if replay_full else 0
            params.pipeline, params.world, params.rank = int(pipeline), world, rank
            params.slots_w, params.slots_m = slot_w.ctypes.data, slot_m.ctypes.data
            keep = []  # ctypes callbacks must outlive the call
            native_engine = getattr(eng, "_h", None) is not None
            if not native_engine:
                # tests: an engine object without a device context stands in for the launch (tests/test_host_forest.py)
                hook = _rollout_hook(eng, slots, K, root_state, self.destination, self.timemin, seed, S, rank * S_loc,
                                     params.roll_flags)
                keep.append(hook)
                params.rollout_hook = hook
            if world > 1:
                # between GPUs: the ranks' windows (NVLink stores, pmcts_peer) when they can be set up, else the
                # library all-gather (pmcts_comm); over a CPU process group the blocks go through the host
                block = int(lib.pmcts_search_block_bytes(S_loc, B, max_depth, K))
                peer = _device_peer(eng, dist, group, rank, world, block) if native_engine else None
                comm = _device_comm(eng, dist, group, rank, world) if native_engine and peer is None else None
                if peer is not None:
                    params.peer = peer._h
                elif comm is not None:
                    params.comm = comm
                else:
                    cb = _host_exchange(dist, group, world)
                    keep.append(cb)
                    params.exchange = cb
            report = N.SearchReport()
            root = np.ascontiguousarray(root_state, np.float64)
            dest = np.ascontiguousarray(self.destination, np.float64)
            rc = lib.pmcts_forest_search(eng._h if native_engine else None, handle, slots.ctypes.data, root.ctypes.data,
                                         dest.ctypes.data, float(self.timemin), C.byref(params), C.byref(report))
            N.check(rc)
            self.last_search_report = report.as_dict()
        except BaseException:
            lib.pmcts_forest_destroy(handle)
            raise
        # The native forest lives on behind the results: the arrays of a tree / of the master are copied out of it
        # when they are first read, and it is destroyed with the last object that refers to it.
        native = _NativeForest(lib, handle, S)
        for ti, t in enumerate(trees):
            _export_tree(native, ti, t, root_state, budget)
        from .master_node import MasterDict
        return native.master_dict()


class _NativeForest:
    """Owner of a ``pmcts_forest`` handle after a search."""

    def __init__(self, lib, handle, nscen):
        self.lib, self.handle, self.nscen = lib, handle, nscen

    def __del__(self):
        try:
            if self.handle:
                self.lib.pmcts_forest_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def master_size(self):
        return int(self.lib.pmcts_forest_master_size(self.handle))

    def master_dict(self):
        from .master_node import MasterDict
        d = MasterDict(loader=self.master_arrays, size=self.master_size(), sparse_loader=self.master_rows)
        d.native = self
        return d

    def master_arrays(self):
        n = self.master_size()
        parent = np.empty(n, np.int32)
        arm = np.empty(n, np.int8)
        table = np.empty((n, self.nscen, len(ACTIONS), 12), np.uint32)
        N.check_forest(self.lib.pmcts_forest_export_master(self.handle, parent.ctypes.data, arm.ctypes.data,
                                                           table.ctypes.data))
        return parent, arm, table

    def master_rows(self):
        n = self.master_size()
        parent = np.empty(n, np.int32)
        arm = np.empty(n, np.int8)
        row_index = np.empty((n, self.nscen), np.int32)
        rows = np.empty((int(self.lib.pmcts_forest_master_rows(self.handle)), 96), np.uint32)
        N.check_forest(self.lib.pmcts_forest_export_master_rows(self.handle, parent.ctypes.data, arm.ctypes.data,
                                                                row_index.ctypes.data, rows.ctypes.data))
        return parent, arm, row_index, rows

    def tree_arrays(self, ti):
        n = int(self.lib.pmcts_forest_tree_size(self.handle, ti))
        parent = np.empty(n, np.int32)
        arm = np.empty(n, np.int8)
        table = np.empty((n, len(ACTIONS), 12), np.int64)
        n_untried = np.empty(n, np.uint8)
        untried = np.empty((n, len(ACTIONS)), np.uint8)
        N.check_forest(self.lib.pmcts_forest_export_tree(self.handle, ti, parent.ctypes.data, arm.ctypes.data,
                                                         table.ctypes.data, n_untried.ctypes.data, untried.ctypes.data))
        return parent, arm, table.astype(int, copy=False), n_untried, untried


def _rollout_hook(eng, slots, K, root_state, destination, timemin, seed, S, first_scenario, flags):
    """ctypes callback that hands a wave to ``eng.rollout_forest`` (an object without a device context: the stand-in
    engines of the CPU tests)."""
    import ctypes as C

    def hook(_user, wave, n_local, _scen_ids, b, stride, plen_p, path_p, bins_p):
        try:
            plen = np.ctypeslib.as_array(C.cast(plen_p, C.POINTER(C.c_int32)), shape=(n_local * b,))
            path = np.ctypeslib.as_array(C.cast(path_p, C.POINTER(C.c_int8)), shape=(n_local, b, stride))
            bins = np.ctypeslib.as_array(C.cast(bins_p, C.POINTER(C.c_uint32)), shape=(n_local * b * 12,))
            depth = max(1, int(plen.max()))
            prefix = np.where(path[:, :, :depth] >= 0, path[:, :, :depth].astype(np.float64) * 45.0, 0.0)
            eng.rollout_forest(slots, b, K, root_state, destination, timemin, np.ascontiguousarray(prefix).reshape(-1),
                               plen, depth, prefix_shared=False, seed=seed, stream=wave * S + first_scenario, id0=0,
                               out=dict(leaf_bins=bins), flags=flags)
            return 0
        except Exception:  # an exception must not cross the C boundary
            import traceback
            traceback.print_exc()
            return 1
    return N.ROLLOUT_HOOK_FN(hook)


def _host_exchange(dist, group, world):
    """All-gather of the update blocks over host buffers (gloo in the CPU tests; any backend without a device
    communicator)."""
    import ctypes as C
    import torch

    def exchange(_user, send_p, nbytes, recv_p):
        try:
            send = torch.from_numpy(np.ctypeslib.as_array(C.cast(send_p, C.POINTER(C.c_uint8)), shape=(nbytes,)))
            recv = torch.from_numpy(np.ctypeslib.as_array(C.cast(recv_p, C.POINTER(C.c_uint8)), shape=(nbytes * world,)))
            if dist.get_backend(group) == "nccl":
                dev = torch.device("cuda", torch.cuda.current_device())
                out = torch.empty(nbytes * world, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(out, send.to(dev), group=group)
                recv.copy_(out.cpu())
            else:
                dist.all_gather_into_tensor(recv, send, group=group)
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return 1
    return N.EXCHANGE_FN(exchange)


def _device_peer(eng, dist, group, rank, world, block_bytes):
    """The engine's peer windows for this process group (``pmcts_peer``), opened on first use and reopened when a
    search needs larger blocks; ``None`` (on every rank) when they cannot be set up."""
    if dist.get_backend(group) != "nccl" or os.environ.get("PMCTS_EXCHANGE") == "nccl":
        return None
    from .peer import open_exchange
    cache = eng.__dict__.setdefault("_peers", {})
    key = (id(group), world, rank)
    have = cache.get(key)
    if have is not None and have[0] is not None and have[0].block_bytes < block_bytes:
        eng.synchronize()
        dist.barrier(group=group)
        have[0].close()
        have = None
    if have is None:
        px, _why = open_exchange(eng, dist, world, rank, (block_bytes * 2 + 15) // 16 * 16, group=group)
        have = cache[key] = (px, block_bytes)
    return have[0]


def _device_comm(eng, dist, group, rank, world):
    """The engine's NCCL communicator for this process group (``pmcts_comm``), created on first use: rank 0's
    unique id reaches the other ranks through one ``torch.distributed`` broadcast."""
    import ctypes as C
    import torch
    if dist.get_backend(group) != "nccl":
        return None
    cache = eng.__dict__.setdefault("_comms", {})
    key = (id(group), world, rank)
    if key not in cache:
        lib = N.lib()
        uid = np.zeros(128, np.uint8)
        if rank == 0:
            N.check(lib.pmcts_comm_unique_id(uid.ctypes.data))
        t = torch.from_numpy(uid).to(torch.device("cuda", eng.device))
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(t, src=src, group=group)
        uid = t.cpu().numpy()
        h = C.c_void_p()
        N.check(lib.pmcts_comm_create(eng._h, world, rank, uid.ctypes.data, C.byref(h)))
        cache[key] = h
    return cache[key]


def _dist_env(group):
    try:
        import torch.distributed as tdist
        if group is not False and tdist.is_available() and tdist.is_initialized():
            return tdist, tdist.get_rank(group), tdist.get_world_size(group)
    except ImportError:
        pass
    return None, 0, 1


def _export_tree(native, ti, tree, root_state, iterations):
    """Points a worker ``Tree`` at its native tree: ``tree.table`` and ``tree.Nodes`` / ``rootNode`` / ``depth``
    (``worker.Node`` objects over that table) are built the first time they are read (``Tree.__getattr__``)."""
    tree.ite = iterations
    for name in ("Nodes", "rootNode", "depth", "table", "_cap"):
        tree.__dict__.pop(name, None)
    tree._lazy_nodes = (native, ti, tuple(root_state))


def materialize_tree(tree):
    """Builds ``tree.table`` and ``tree.Nodes`` from the native tree (called by ``Tree.__getattr__``)."""
    native, ti, root_state = tree.__dict__.pop("_lazy_nodes")
    parent, arm, table, n_untried, untried = native.tree_arrays(ti)
    tree.table, tree._cap = table, parent.size
    nodes = []
    for i in range(parent.size):
        node = mt.Node.__new__(mt.Node)
        p = nodes[parent[i]] if parent[i] >= 0 else None
        node.state = root_state if p is None else None
        node.parent = p
        node.origins = [] if p is None else p.origins + [ACTIONS[arm[i]]]
        node.children = []
        node.actions = [ACTIONS[a] for a in untried[i, :n_untried[i]]]
        node._table = tree.table[i]
        node._values = None
        node.depth = 0 if p is None else p.depth + 1
        node.pending = 0
        if p is not None:
            p.children.append(node)
        nodes.append(node)
    tree.Nodes = nodes
    tree.rootNode = nodes[0]
    tree.depth = max(nd.depth for nd in nodes)


def _master_add(Master_nodes, nscen, scen_id, path, slot, bins):
    """``Tree.integrate_buffer`` (worker.py:348-378) for one leaf and a whole histogram of rewards."""
    actions = [int(a) for a in path]
    key = hash(tuple(actions))
    node = Master_nodes.get(key, 0)
    if node == 0:
        # ancestors first (a leaf of a later wave may sit several levels below the last known node)
        parent = Master_nodes[ROOT_HASH]
        for d in range(1, len(actions) + 1):
            k = hash(tuple(actions[:d]))
            nxt = Master_nodes.get(k, 0)
            if nxt == 0:
                nxt = MasterNode(nscen, nodehash=k, parentNode=parent, action=actions[d - 1])
                Master_nodes[k] = nxt
            parent = nxt
        node = Master_nodes[key]
    node.add_bins(scen_id, slot, bins)
    while node.parentNode is not None:
        node.parentNode.add_bins(scen_id, A_DICT[node.arm], bins)
        node = node.parentNode


def _all_gather_blocks(block, dist, group, world):
    """One all-gather of the fixed-size update block per wave (NCCL on the GPUs; gloo in the CPU tests)."""
    if dist is None or world == 1:
        return [block]
    import torch
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    flat = np.concatenate([block["ids"].ravel(), block["paths"].ravel(), block["slots"].ravel(),
                           block["bins"].astype(np.int64).astype(np.int32).ravel()]).astype(np.int32)
    send = torch.from_numpy(flat).to(dev)
    recv = torch.empty(world * flat.size, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, -1)
    out = []
    s_loc, b = block["paths"].shape[:2]
    for r in range(world):
        o = 0
        blk = {}
        for name, shape in (("ids", (s_loc,)), ("paths", (s_loc, b, block["paths"].shape[2])), ("slots", (s_loc, b)),
                            ("bins", (s_loc, b, 12))):
            size = int(np.prod(shape))
            blk[name] = recv[r, o:o + size].reshape(shape)
            o += size
        blk["bins"] = blk["bins"].astype(np.uint32)
        out.append(blk)
    return out


# ----------------------------------------------------------------------------------------------------
def load_scenarios(mydate, scenario_ids=range(0, 21), latBound=[-90, 90], lonBound=[0, 360], timeSteps=[0, 64],
                   directory='../data/'):
    """The pickled ``Weather`` of every scenario id, cropped (forest.py:118-151): files
    ``<directory><mydate>_<id, two digits>00z.obj`` as the reference's ``download_scenarios`` names them
    (id 0 = the ensemble's control run).  ``directory`` is an addition; the default is the reference's fixed path.
    Files written by the reference unpickle into the mirror ``Weather`` once the module alias of INTEGRATION.md is
    in place (``Weather.load``)."""
    from .weatherTLKT import Weather
    weather_scen = []
    for ii in scenario_ids:
        path = '%s%s_%02d00z.obj' % (directory, mydate, ii)
        weather_scen.append(Weather.load(path, latBound, lonBound, timeSteps))
        print("Loaded : " + path)
    return weather_scen


def create_simulators(weathers, numberofsim, simtimestep=6, stateinit=(0, 47.5, -3.5 + 360),
                      ndaysim=8, deltalatlon=0.5):
    """One ``Simulator`` per weather scenario (forest.py:154-187).  Shifts each ``weather.time`` to start
    at 0 in place, as the reference does."""
    sims = []
    for jj in range(numberofsim):
        weathers[jj].time = weathers[jj].time - weathers[jj].time[0]
        times = np.arange(0, ndaysim + simtimestep * HOURS_TO_DAY, simtimestep * HOURS_TO_DAY)
        lats = np.arange(weathers[jj].lat[0], weathers[jj].lat[-1], deltalatlon)
        lons = np.arange(weathers[jj].lon[0], weathers[jj].lon[-1], deltalatlon)
        sims.append(Simulator(times, lats, lons, weathers[jj], list(stateinit)))
    return sims


def _reference_noise(n_steps):
    """n_steps standard normals from Python's global generator, and a function that leaves the generator
    where ``used`` calls of ``random.gauss`` would have."""
    state = rand.getstate()
    z = np.array([rand.gauss(0.0, 1.0) for _ in range(n_steps)], np.float64)

    def consume(used):
        rand.setstate(state)
        for _ in range(used):
            rand.gauss(0.0, 1.0)
    return z, consume


def initialize_simulators(sims, ntra, stateinit, missionheading, plot=False, seed=None):
    """Destination point and reference time of a search (forest.py:210-331).

    Phase 1: per scenario, ``ntra`` boats hold ``missionheading`` until their next state would be terminal;
    the destination lies at the smallest of the scenarios' mean distances.  Phase 2: per scenario, ``ntra``
    default-policy runs to that destination; ``timemin`` is the smallest arrival time.

    ``seed=None``: the noise comes from ``random.gauss`` in the reference's order (one launch per
    trajectory), so a seeded script gets the reference's numbers.  ``seed=int``: Philox noise, one launch
    per scenario and phase.
    """
    if plot:
        raise NotImplementedError("plotting is outside the scope of pmcts_b200")
    k0, lat0, lon0 = int(stateinit[0]), float(stateinit[1]), float(stateinit[2])
    meanarrivaldistances = []
    for si, sim in enumerate(sims):
        eng, scen = sim.scenario()
        nT = len(sim.times)
        prec = eng.precision
        eng.set_precision(N.PREC_F64 if seed is None else N.PREC_FAST)
        try:
            if seed is None:
                ends = []
                for _ in range(ntra):
                    z, consume = _reference_noise(nT)
                    k, la, lo = np.array([k0], np.int32), np.array([lat0]), np.array([lon0])
                    st = np.zeros(1, np.uint8)
                    eng.step_batch(scen, k, la, lo, np.array([[float(missionheading)]]), nsteps=nT, seg_len=nT,
                                   status=st, z=z.reshape(nT, 1), flags=N.STEP_STOP_BEFORE_TERMINAL)
                    consume(int(k[0]) - k0)
                    _raise_unless(st, (N.ST_HORIZON, N.ST_OUT_OF_BOX))
                    ends.append((float(la[0]), float(lo[0])))
            else:
                k, la, lo = np.full(ntra, k0, np.int32), np.full(ntra, lat0), np.full(ntra, lon0)
                st = np.zeros(ntra, np.uint8)
                eng.step_batch(scen, k, la, lo, np.full((1, ntra), float(missionheading)), nsteps=nT, seg_len=nT,
                               status=st, seed=seed, stream=2 * si, flags=N.STEP_STOP_BEFORE_TERMINAL)
                _raise_unless(st, (N.ST_HORIZON, N.ST_OUT_OF_BOX))
                ends = list(zip(la.tolist(), lo.tolist()))
        finally:
            eng.set_precision(prec)
        arrivaldistances = [sim.getDistAndBearing([lat0, lon0], list(p))[0] for p in ends]
        meanarrivaldistances.append(np.mean(arrivaldistances))
    mindist = np.min(meanarrivaldistances)
    destination = sims[-1].getDestination(mindist, missionheading, [lat0, lon0])

    min_arrival_times = []
    for si, sim in enumerate(sims):
        eng, scen = sim.scenario()
        nT = len(sim.times)
        prec = eng.precision
        eng.set_precision(N.PREC_F64 if seed is None else N.PREC_FAST)
        try:
            if seed is None:
                arrivaltimes = []
                for _ in range(ntra):
                    z, consume = _reference_noise(nT)
                    r = eng.rollout_batch_host(scen, 1, 1, [k0, lat0, lon0], destination, 0.0, z=z.reshape(nT, 1),
                                               flags=N.ROLL_PLAN_EVAL, want=("t_arr", "steps", "status"))
                    consume(int(r["steps"][0]))
                    _raise_unless(r["status"], (N.ST_ARRIVED, N.ST_HORIZON, N.ST_OUT_OF_BOX))
                    if r["status"][0] == N.ST_ARRIVED:
                        arrivaltimes.append(float(r["t_arr"][0]))
            else:
                r = eng.rollout_batch_host(scen, ntra, 1, [k0, lat0, lon0], destination, 0.0, seed=seed,
                                           stream=2 * si + 1, flags=N.ROLL_PLAN_EVAL, want=("t_arr", "status"))
                _raise_unless(r["status"], (N.ST_ARRIVED, N.ST_HORIZON, N.ST_OUT_OF_BOX))
                arrivaltimes = r["t_arr"][r["status"] == N.ST_ARRIVED].tolist()
        finally:
            eng.set_precision(prec)
        if arrivaltimes:
            min_arrival_times.append(min(arrivaltimes))
        else:
            print("Scenario num : " + str(si) + " did not reach destination")
    timemin = min(min_arrival_times)
    return [destination, timemin]


def _raise_unless(status, allowed):
    from .simulatorTLKT import raise_for_status
    for s in np.asarray(status).ravel():
        if int(s) not in allowed and int(s) != N.ST_OK:
            raise_for_status(int(s))


download_scenarios = out_of_scope("forest.download_scenarios")
play_multiple_scenarios = out_of_scope("forest.play_multiple_scenarios")
