"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* Python reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/gen_golden.py`` to produce the committed fixtures under ``tests/golden/`` and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference is not mounted).

Two shims are needed to import the reference with this image's libraries:

1. plotting / netCDF modules are absent -> register empty stub modules
   (imports at ``code/model/simulatorTLKT.py:17-24``, ``code/model/weatherTLKT.py:11-18``,
   ``code/solver/utils.py:4-6``, ``code/solver/forest.py:4``, ``code/solver/master.py:6-8``).
2. NumPy >= 2 rejects ``polyval2d(scalar, shape-(1,) array)`` at ``simulatorTLKT.py:502`` because
   SciPy returns shape-(1,) arrays from ``getWind`` (``simulatorTLKT.py:176-177``); ``getWind`` is
   replaced by a version that returns the same two numbers as Python floats.  Arithmetic unchanged.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PMCTS_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "code", "model"))


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _install_stubs():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = _stub("matplotlib", rcParams={}, rc=lambda *a, **k: None)
            mpl.pyplot = _stub("matplotlib.pyplot", figure=_Dummy, show=lambda *a, **k: None,
                               get_cmap=_Dummy, title=_Dummy, legend=_Dummy, annotate=_Dummy,
                               cm=_Dummy())
            mpl.animation = _stub("matplotlib.animation", FuncAnimation=_Dummy)
            mpl.widgets = _stub("matplotlib.widgets", Button=_Dummy)
    if "mpl_toolkits.basemap" not in sys.modules:
        tk = sys.modules.get("mpl_toolkits") or _stub("mpl_toolkits")
        tk.basemap = _stub("mpl_toolkits.basemap", Basemap=_Dummy)
        tk.axes_grid1 = _stub("mpl_toolkits.axes_grid1")
    if "netCDF4" not in sys.modules:
        _stub("netCDF4", Dataset=_Dummy)


_loaded = None


def load_reference():
    """Returns a namespace with the reference modules (simulatorTLKT, weatherTLKT, worker, ...)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    _install_stubs()
    for sub in ("model", "solver", "isochrones"):
        p = os.path.join(REFERENCE_ROOT, "code", sub)
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    import simulatorTLKT as simc
    import weatherTLKT as wth
    import utils as rutils
    import master_node as mnode
    import worker as wrk
    import forest as frst
    import master as mstr
    import isochrones as iso

    def _get_wind_scalar(self):
        p = [self.times[self.state[0]], self.state[1], self.state[2]]
        return float(self.uWindAvg(p)[0]), float(self.vWindAvg(p)[0])

    simc.Simulator.getWind = _get_wind_scalar

    _loaded = types.SimpleNamespace(simulatorTLKT=simc, weatherTLKT=wth, utils=rutils,
                                    master_node=mnode, worker=wrk, forest=frst, master=mstr,
                                    isochrones=iso)
    return _loaded
