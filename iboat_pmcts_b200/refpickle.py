"""Pickles the reference's own tools can read.

``MasterTree.save_tree`` of the reference is ``pickle.dump(self)`` (code/solver/master.py:390-398): the file names
the classes ``master.MasterTree``, ``master_node.MasterNode``, ``utils.Hist`` and ``simulatorTLKT.Simulator`` and
holds their ``__dict__``s -- ``MasterNode.rewards`` as an object array of ``Hist`` whose ``h`` are int arrays,
``Simulator.uWindAvg`` / ``vWindAvg`` as SciPy ``RegularGridInterpolator`` objects.  ``dump_reference_tree`` writes
exactly that from a ``MasterTree`` of this package -- straight from the array-backed master dictionary of a native
search (``master_node.MasterDict``), node objects of the mirror classes never being built -- so that the reference's
``MasterTree.load_tree`` (master.py:400-410) and its plotting methods work on a GPU search's result.

The other way round, ``load_reference_tree`` reads a file the REFERENCE's ``save_tree`` wrote -- the result of a CPU
search of the reference -- into this package's classes (``MasterTree`` over ``MasterNode`` tables, ``Simulator`` /
``Weather`` rebuilt from the grids the pickled SciPy interpolators hold), without importing any reference code: its
decision is then one call of the device reductions.
"""
import pickle
import sys
import types

import numpy as np

from .master_node import MasterDict
from .simulatorTLKT import ACTIONS

_NAMES = {"master": ("MasterTree",), "master_node": ("MasterNode",), "utils": ("Hist",),
          "simulatorTLKT": ("Simulator",)}


class _Stubs:
    """Puts classes named like the reference's under the reference's module names for the duration of a dump: pickle
    stores a class as (module, name) and checks that the name resolves to the object being stored."""

    def __enter__(self):
        self.saved, self.cls = {}, {}
        for mod, names in _NAMES.items():
            self.saved[mod] = sys.modules.get(mod)
            m = types.ModuleType(mod)
            for n in names:
                c = type(n, (), {})
                c.__module__ = mod
                setattr(m, n, c)
                self.cls[n] = c
            sys.modules[mod] = m
        return self.cls

    def __exit__(self, *exc):
        for mod, old in self.saved.items():
            if old is None:
                sys.modules.pop(mod, None)
            else:
                sys.modules[mod] = old


def _make(cls, **attrs):
    obj = cls.__new__(cls)
    obj.__dict__.update(attrs)
    return obj


def _arrays_of(nodes):
    """(parent, arm, table[node][scenario][8][12], hashes) of a master dictionary, parents before children."""
    if isinstance(nodes, MasterDict) and nodes.arrays() is not None:
        parent, arm, table = nodes.arrays()
        paths, hashes = [], []
        for i in range(parent.size):
            path = () if parent[i] < 0 else paths[parent[i]] + (int(ACTIONS[arm[i]]),)
            paths.append(path)
            hashes.append(hash(path))
        return parent, arm, np.asarray(table), hashes
    order = sorted(nodes.values(), key=_depth)
    index = {id(nd): i for i, nd in enumerate(order)}
    parent = np.array([index[id(nd.parentNode)] if nd.parentNode is not None else -1 for nd in order], np.int32)
    arm = np.array([list(ACTIONS).index(nd.arm) if nd.parentNode is not None else 0 for nd in order], np.int8)
    table = np.stack([np.asarray(nd._table) for nd in order])
    return parent, arm, table, [nd.hash for nd in order]


def _depth(node):
    d = 0
    while node.parentNode is not None:
        node, d = node.parentNode, d + 1
    return d


def dump_reference_tree(tree, file, with_interpolators=True):
    """Writes ``tree`` (a ``master.MasterTree`` of this package) to the binary file object ``file`` as the reference's
    ``MasterTree.save_tree`` would have."""
    parent, arm, table, hashes = _arrays_of(tree.nodes)
    n = int(parent.size)
    depth = np.zeros(n, int)
    for i in range(1, n):
        depth[i] = depth[parent[i]] + 1
    with _Stubs() as C:
        objs = []
        for i in range(n):
            rewards = np.empty(table.shape[1:3], dtype=object)
            for s in range(table.shape[1]):
                for a in range(table.shape[2]):
                    rewards[s, a] = _make(C["Hist"], h=np.array(table[i, s, a], dtype=int))
            node = _make(C["MasterNode"], hash=hashes[i], arm=None if parent[i] < 0 else ACTIONS[arm[i]],
                         parentNode=objs[parent[i]] if parent[i] >= 0 else None, children=[], depth=int(depth[i]),
                         guessed_rewards=dict(), rewards=rewards,
                         parentHash=hashes[parent[i]] if parent[i] >= 0 else None)
            if parent[i] >= 0:
                objs[parent[i]].children.append(node)
            objs.append(node)
        by_hash = dict(zip(hashes, objs))
        sims = []
        for sim in tree.Simulators:
            attrs = dict(times=np.asarray(sim.times), lats=np.asarray(sim.lats), lons=np.asarray(sim.lons),
                         state=list(sim.state), prevState=list(sim.prevState))
            w = getattr(sim, "_weather", None)
            if with_interpolators and w is not None:
                from scipy.interpolate import RegularGridInterpolator as rgi
                grid = (np.asarray(w.time, float), np.asarray(w.lat, float), np.asarray(w.lon, float))
                attrs["uWindAvg"] = rgi(grid, np.asarray(w.u, float))
                attrs["vWindAvg"] = rgi(grid, np.asarray(w.v, float))
            sims.append(_make(C["Simulator"], **attrs))
        idx = {h: i for i, h in enumerate(hashes)}
        out = _make(C["MasterTree"], nodes=by_hash, Simulators=sims, numScenarios=tree.numScenarios,
                    destination=list(tree.destination), probability=np.asarray(tree.probability),
                    best_policy={k: list(v) for k, v in tree.best_policy.items()},
                    best_nodes_policy={k: [objs[idx[nd.hash]] for nd in v] for k, v in tree.best_nodes_policy.items()})
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))  # (the reference's node graph is recursive)
        pickle.dump(out, file, protocol=4)


class _Shell:
    """What a class of the reference unpickles into here: an object that only carries the pickled ``__dict__``."""


class _ShellUnpickler(pickle.Unpickler):
    def __init__(self, file):
        super().__init__(file)
        self._shells = {}

    def find_class(self, module, name):
        if module in _NAMES and name in _NAMES[module] or (module, name) == ("weatherTLKT", "Weather"):
            if (module, name) not in self._shells:
                self._shells[module, name] = type(name, (_Shell,), {"__module__": module})
            return self._shells[module, name]
        return super().find_class(module, name)


def is_reference_tree(obj):
    return isinstance(obj, _Shell) and type(obj).__name__ == "MasterTree"


def load_reference_tree(file):
    """A ``master.MasterTree`` of this package from the binary file object ``file`` written by the reference's
    ``MasterTree.save_tree`` (master.py:390-398); an object pickled by this package's own ``save_tree`` comes back as it
    is."""
    from .master import MasterTree
    from .master_node import MasterNode
    from .simulatorTLKT import Simulator
    from .weatherTLKT import Weather
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))
    src = _ShellUnpickler(file).load()
    if not is_reference_tree(src):
        return src
    sims = []
    for sim in src.Simulators:
        u, v = getattr(sim, "uWindAvg", None), getattr(sim, "vWindAvg", None)
        if u is None or v is None:
            raise ValueError("the pickled Simulator carries no wind interpolators")
        t, la, lo = (np.asarray(g, np.float64) for g in u.grid)
        new = Simulator(np.asarray(sim.times), np.asarray(sim.lats), np.asarray(sim.lons),
                        Weather(la, lo, t, np.asarray(u.values), np.asarray(v.values)), list(sim.state))
        new.prevState = list(sim.prevState)
        sims.append(new)
    # nodes in the file's dictionary order (the order deepcopy_dict and a reader of tree.nodes see); children in theirs
    built = {}
    for key, node in src.nodes.items():
        new = MasterNode(src.numScenarios, nodehash=node.hash, action=node.arm, rewards=node.rewards)
        new.depth = node.depth
        new.guessed_rewards = dict(node.guessed_rewards)
        built[id(node)] = new
    nodes = {}
    for key, node in src.nodes.items():
        new = built[id(node)]
        new.parentNode = built[id(node.parentNode)] if node.parentNode is not None else None
        new.children = [built[id(c)] for c in node.children]
        nodes[key] = new
    tree = MasterTree(sims, list(src.destination), nodes=nodes, proba=list(np.asarray(src.probability)))
    tree.best_policy = {k: list(v) for k, v in src.best_policy.items()}
    tree.best_nodes_policy = {k: [built[id(nd)] for nd in v] for k, v in src.best_nodes_policy.items()}
    return tree
