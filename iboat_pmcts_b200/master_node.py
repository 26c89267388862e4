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
        if d.get("_owner") is not None:  # a node of an array-backed dictionary pickles as a plain node
            d["children"] = self.children
            d["_owner"] = None
        if d.get("_source") is not None:
            d["_table"] = self._table
            d["_source"] = None
        return d

    def __getattr__(self, name):
        # Only reached when the attribute is missing: the children and the counters of a node of an array-backed
        # dictionary (MasterDict) are fetched the first time they are read.
        if name == "children" and self.__dict__.get("_owner") is not None:
            kids = self.__dict__["children"] = self._owner._children_of(self._index)
            return kids
        if name == "_table" and self.__dict__.get("_source") is not None:
            table = self.__dict__["_table"] = self._source._table_of(self._index)
            return table
        raise AttributeError(name)

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


class MasterDict(dict):
    """The master dictionary ``{hash(tuple(origins)): MasterNode}`` over the arrays of the native master tree
    (``parent[node]``, ``arm[node]`` = index in ACTIONS, ``table[node][scenario][8][12]``; a node's parent always
    has a smaller index).  A search of 10^5 nodes would spend longer building 10^5 ``MasterNode`` objects than
    searching, so they are built when the dictionary is first read -- all of them by the ``dict`` methods, one by
    one (with its ancestors) by ``node_at``.  ``wired`` dictionaries (what ``deepcopy_dict`` returns) also carry
    ``children`` and ``depth``.  ``arrays()`` hands the arrays to array consumers (``MasterTree``'s device
    reductions) as long as the dictionary has not been edited through the ``dict`` interface."""

    def __init__(self, parent=None, arm=None, table=None, wired=False, loader=None, size=None, sparse_loader=None):
        super().__init__()
        self._loader = loader  # () -> (parent, arm, table): the arrays are fetched on first need
        self._sparse_loader = sparse_loader  # () -> (parent, arm, row_index, rows): the native runtime's own layout
        self._sparse = None
        self.native = None  # the native forest behind a search result (forest._NativeForest), for consumers of its arrays
        self._parent, self._arm, self._table = parent, arm, table
        self._wired = bool(wired)
        n = int(size if loader is not None else parent.size)
        self._nodes = [None] * n
        self._filled = n == 0
        self._dirty = False
        self._depth = None
        self._kids = None

    # ---- array side ------------------------------------------------------------------------------------
    def _load(self):
        if self._loader is not None:
            self._parent, self._arm, self._table = self._loader()
            self._loader = None
            for node in self._nodes:  # nodes built from the sparse rows before now: re-point them at the table
                if node is not None and node.__dict__.get("_source") is not None:
                    if "_table" in node.__dict__:
                        self._table[node._index] = node.__dict__["_table"]
                    node.__dict__["_table"] = self._table[node._index]
                    node._source = None
                    node._rewards = None

    def arrays(self):
        """(parent, arm, table) or None once nodes have been added / removed through the dict interface."""
        if self._dirty:
            return None
        self._load()
        return self._parent, self._arm, self._table

    def sparse(self):
        """(parent, arm, row_index[node][scenario], rows[n_rows][96]) -- a (node, scenario) pair has a row only if
        the scenario visited the node -- or None once the dictionary has been edited."""
        if self._dirty:
            return None
        if self._sparse is not None:
            return self._sparse
        if self._sparse_loader is not None:
            self._sparse = self._sparse_loader()
            self._sparse_loader = None
            if self._parent is None:
                self._parent, self._arm = self._sparse[0], self._sparse[1]
            return self._sparse
        parent, arm, table = self.arrays()
        n, S = table.shape[0], table.shape[1]
        flat = np.ascontiguousarray(table, dtype=np.uint32).reshape(n * S, 96)
        used = flat.any(axis=1)
        row_index = np.where(used, np.cumsum(used) - 1, -1).astype(np.int32).reshape(n, S)
        return parent, arm, row_index, np.ascontiguousarray(flat[used])

    def depths(self):
        if self._parent is None:
            self._load()
        if self._depth is None:
            d = np.zeros(self._parent.size, np.int32)
            for i in range(1, d.size):  # parents come first
                d[i] = d[self._parent[i]] + 1
            self._depth = d
        return self._depth

    def _table_of(self, i):
        """[scenario][8][12] counters of node ``i``: a view of the dense table once that is loaded, else gathered from
        the sparse rows."""
        if self._table is not None:
            return self._table[i]
        sparse = self.sparse()
        rows = sparse[2][i]
        tab = np.zeros((rows.size, len(ACTIONS) * Hist.N_BINS), dtype=int)
        has = rows >= 0
        tab[has] = sparse[3][rows[has]]
        return tab.reshape(rows.size, len(ACTIONS), Hist.N_BINS)

    def _children_of(self, i):
        if self._parent is None:
            self._load()
        if self._kids is None:
            kids = [[] for _ in range(self._parent.size)]
            for j in range(1, self._parent.size):
                kids[self._parent[j]].append(j)
            self._kids = kids
        return [self.node_at(j) for j in self._kids[i]]

    def node_at(self, i):
        """The ``MasterNode`` of array index ``i`` (built with its ancestors on first use)."""
        node = self._nodes[i]
        if node is not None:
            return node
        # the dense table is only fetched when somebody asks for all of it (arrays / the dict methods): a single node
        # takes its counters from the sparse rows
        sparse = self.sparse() if (self._table is None and (self._sparse is not None or self._sparse_loader is not None)) else None
        if sparse is None:
            self._load()
        elif self._parent is None:
            self._parent, self._arm = sparse[0], sparse[1]
        p = int(self._parent[i])
        parent = self.node_at(p) if p >= 0 else None
        node = MasterNode.__new__(MasterNode)
        node._path = () if parent is None else parent._path + (int(ACTIONS[self._arm[i]]),)
        node.hash = hash(node._path)
        node.arm = None if parent is None else ACTIONS[self._arm[i]]
        node.parentNode = parent
        node.guessed_rewards = dict()
        if sparse is None:
            node._table = self._table[i]  # a view: edits through the node show in arrays()
            node._source = None
        else:
            node._source = self  # counters gathered from the sparse rows when first read (MasterNode.__getattr__)
        node._rewards = None
        node._index = i
        if self._wired:
            node._owner = self  # children are listed on first read (MasterNode.__getattr__)
            node.depth = 0 if parent is None else parent.depth + 1
        else:  # as the search leaves them: not linked downwards (master_node.py:21-22)
            node._owner = None
            node.children = []
            node.depth = None
        self._nodes[i] = node
        return node

    def nodes_along(self, indices, actions):
        """The nodes of a root-to-node chain given by array indices (root first) and the actions between them (degrees)
        -- what a greedy policy is -- built without fetching any array: their counters and children come later, if
        somebody reads them."""
        out, parent = [], None
        for d, i in enumerate(indices):
            node = self._nodes[i]
            if node is None:
                node = MasterNode.__new__(MasterNode)
                node._path = () if parent is None else parent._path + (int(actions[d - 1]),)
                node.hash = hash(node._path)
                node.arm = None if parent is None else actions[d - 1]
                node.parentNode = parent
                node.guessed_rewards = dict()
                node._rewards = None
                node._index = i
                if self._table is not None:
                    node._table, node._source = self._table[i], None
                else:
                    node._source = self
                if self._wired:
                    node._owner = self
                    node.depth = d
                else:
                    node._owner = None
                    node.children = []
                    node.depth = None
                self._nodes[i] = node
            out.append(node)
            parent = node
        return out

    def _fill(self):
        if not self._filled:
            self._filled = True
            self._load()  # every node is wanted: they view the dense table
            for i in range(len(self._nodes)):
                dict.__setitem__(self, self.node_at(i).hash, self._nodes[i])

    # ---- dict side: everything reads through _fill -----------------------------------------------------
    def __len__(self):
        return dict.__len__(self) if self._filled else len(self._nodes)

    def __reduce__(self):  # pickles as a plain dict of plain nodes (master.py:390-410 tools read it)
        self._fill()
        return (dict, (), None, None, iter(dict.items(self)))


def _reads(name, edits=False):
    base = getattr(dict, name)

    def method(self, *a, **k):
        self._fill()
        if edits:
            self._dirty = True
        return base(self, *a, **k)
    method.__name__ = name
    return method


for _name in ("__getitem__", "__contains__", "__iter__", "get", "keys", "values", "items", "__repr__", "__eq__",
              "__ne__", "copy", "__reversed__"):
    setattr(MasterDict, _name, _reads(_name))
for _name in ("__setitem__", "__delitem__", "pop", "popitem", "setdefault", "update", "clear", "__ior__"):
    setattr(MasterDict, _name, _reads(_name, edits=True))
MasterDict.__hash__ = None


def deepcopy_dict(nodes):
    """Deep copy of the dictionary the search returns, rewired with ``parentNode``, ``children`` and
    ``depth`` (master_node.py:66-109).  Empties ``nodes`` as the reference does."""
    if isinstance(nodes, MasterDict) and not nodes._dirty:
        # array-backed result of the native search: the copy is three array copies
        if nodes._loader is not None:
            # not read yet: the copy fetches its own arrays from the (immutable) native forest when it needs them
            loader = nodes._loader
            new = MasterDict(wired=True, size=len(nodes), sparse_loader=nodes._sparse_loader,
                             loader=lambda: tuple(np.array(a, dtype=int if a.ndim == 4 else a.dtype) for a in loader()))
            new._sparse = nodes._sparse
            new.native = nodes.native
        else:
            parent, arm, table = nodes.arrays()
            new = MasterDict(parent.copy(), arm.copy(), np.array(table, dtype=int), wired=True)
        dict.clear(nodes)
        empty = np.zeros(0, np.int32)
        nodes._loader = nodes._sparse_loader = nodes._sparse = None
        nodes._parent, nodes._arm, nodes._table = empty, empty.astype(np.int8), np.zeros((0, 1, 8, 12), np.uint32)
        nodes._nodes, nodes._filled, nodes._kids, nodes._depth = [], True, None, None
        return new
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
