"""The C-ABI library loads without a GPU and exports every symbol include/pmcts_b200.h declares; the product
has no CPU path (creating a context without a CUDA device fails loudly)."""
import ctypes as C
import os
import re

import pytest

from iboat_pmcts_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "pmcts_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmcts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 24
    lib = N.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(N.EXPORTS) == names  # the ctypes binding lists exactly the header's functions
    assert b"sm_100a" in lib.pmcts_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = N.lib().pmcts_create(0, C.byref(h))
    assert rc == N.ERR_NO_DEVICE and not h.value
    assert b"no CPU path" in N.lib().pmcts_last_error()
    from iboat_pmcts_b200.engine import Engine
    with pytest.raises(N.PmctsError):
        Engine(0)
    # the reference-API mirror goes through the same engine: no silent host implementation of doStep
    from iboat_pmcts_b200 import synth
    from iboat_pmcts_b200.forest import create_simulators
    from iboat_pmcts_b200.weatherTLKT import Weather
    t, la, lo, u, v = synth.wind_fields(0)
    with pytest.raises(N.PmctsError):
        sims = create_simulators([Weather(la, lo, t, u, v)], 1, stateinit=(0, 44.0, 355.0), ndaysim=3)
        sims[0].doStep(0)


def test_product_does_not_import_test_infrastructure():
    pkg = os.path.join(ROOT, "iboat_pmcts_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "from oracle" not in text, fn
                assert "fastmodel" not in text.replace("tests/fastmodel", "") or fn.endswith((".cuh", ".h")), fn
    # nor do the measurement scripts (what loads the oracle by hand lives under tests/tools/); bench.py does, and only in
    # its CPU-baseline / reference-arm leg
    for fn in os.listdir(os.path.join(ROOT, "scripts")):
        if fn.endswith((".py", ".sh")):
            text = open(os.path.join(ROOT, "scripts", fn)).read()
            for word in ("import oracle", "from oracle", "pmcts_helpers", "import fastmodel", "oracle_engine", "liboracle"):
                assert word not in text, (fn, word)
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for node in tree.body:   # every import of the oracle sits inside one of the cpu_*_sample functions
        inside = isinstance(node, ast.FunctionDef) and node.name.startswith("cpu_") and node.name.endswith("_sample")
        for sub in ast.walk(node):
            uses = (isinstance(sub, ast.ImportFrom) and (sub.module or "").split(".")[0] == "oracle") or \
                   (isinstance(sub, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in sub.names))
            assert not uses or inside, getattr(node, "name", type(node).__name__)


def test_weather_pickle_ingest(tmp_path):
    """Weather.load: a pickle written under the reference's module name unpickles into the mirror class and
    is cropped like the reference's (strict bounds)."""
    import pickle
    import sys
    import numpy as np
    from iboat_pmcts_b200 import synth, weatherTLKT
    t, la, lo, u, v = synth.wind_fields(1)
    sys.modules["weatherTLKT"] = weatherTLKT
    try:
        w = weatherTLKT.Weather(la, lo, t, u.astype(np.float32), v.astype(np.float32))
        weatherTLKT.Weather.__module__ = "weatherTLKT"      # what a pickle written by the reference says
        try:
            blob = pickle.dumps(w)
        finally:
            weatherTLKT.Weather.__module__ = "iboat_pmcts_b200.weatherTLKT"
        assert b"iboat_pmcts_b200" not in blob
        path = tmp_path / "20180108_gep_0100z.obj"
        path.write_bytes(blob)
        got = weatherTLKT.Weather.load(str(path), latBound=[40, 50], lonBound=[345, 360], timeSteps=[0, 24])
    finally:
        del sys.modules["weatherTLKT"]
    assert isinstance(got, weatherTLKT.Weather)
    assert got.lat[0] == 40.5 and got.lat[-1] == 49.5 and got.lon[0] == 345.5 and got.time.size == 24
    assert got.u.dtype == np.float64 and got.u.shape == (24, 19, 29)
    assert np.array_equal(got.u, u[:24, 1:-1, 1:-1])


def test_the_boundary_from_plain_c(tmp_path):
    """tests/c_abi/abi_from_c.c: include/pmcts_b200.h compiled as strict C99 (no warnings), linked against the
    library, run without Python: no device -> PMCTS_ERR_NO_DEVICE ("no CPU path"); the host runtime of the search driven
    through the ABI from C."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    from iboat_pmcts_b200 import _native
    lib_dir = os.path.dirname(_native.LIB_PATH)
    exe = str(tmp_path / "abi_from_c")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-o", exe,
                          os.path.join(ROOT, "tests", "c_abi", "abi_from_c.c"), "-L", lib_dir,
                          "-l:" + os.path.basename(_native.LIB_PATH), "-Wl,-rpath," + lib_dir],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    run = subprocess.run([exe] + (["gpu"] if has_gpu else []), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=120)
    assert run.returncode == 0 and "abi_from_c ok" in run.stdout, run.stdout


def test_the_ctypes_stub_of_integration_md_binds():
    """Route B of INTEGRATION.md -- the stub a maintainer of the reference would add -- is executed as written (library
    path filled in): every symbol it names resolves, its argument lists have the header's lengths, and on a machine
    without a GPU it stops where it must, at pmcts_create ("no CPU path")."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    start = text.index("```python\nimport ctypes as C, numpy as np")
    code = text[start + len("```python\n"):text.index("```", start + 10)]
    code = code.replace('"/path/to/libpmcts_b200.so"', repr(N.LIB_PATH))
    header = open(os.path.join(ROOT, "include", "pmcts_b200.h")).read()

    def n_params(name):
        m = re.search(r"\b%s\(([^;]*?)\);" % name, header, re.S)
        return len([p for p in m.group(1).split(",") if p.strip()])

    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    scope = {}
    if has_gpu:
        exec(code, scope)
    else:
        with pytest.raises(RuntimeError, match="no CPU path"):
            exec(code, scope)
    L = scope["_L"]
    for name in ("pmcts_create", "pmcts_scenario_upload", "pmcts_rollout_batch"):
        assert len(getattr(L, name).argtypes) == n_params(name), name
    assert scope["PMCTS_MEM_HOST"] == N.MEM_HOST and C.sizeof(scope["Noise"]) == C.sizeof(N.Noise)


def test_the_module_swap_of_integration_md_works_as_written():
    """Route A of INTEGRATION.md, executed as written in a fresh interpreter: after the swap the reference's own import
    lines resolve to the mirror modules, which carry the names a script of the reference uses (no device is touched)."""
    import subprocess
    import sys
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    start = text.index("```python\nimport sys, iboat_pmcts_b200")
    code = text[start + len("```python\n"):text.index("```", start + 10)]
    check = code + "\n" + "\n".join([
        "import master, master_node, isochrones, utils",
        "assert ft.__name__ == 'iboat_pmcts_b200.forest' and mt.__name__ == 'iboat_pmcts_b200.worker'",
        "assert Weather.__module__ == 'iboat_pmcts_b200.weatherTLKT' and Simulator.__module__ == 'iboat_pmcts_b200.simulatorTLKT'",
        "assert Boat.UNCERTAINTY_COEFF == 0.2 and mt.RHO == 0.3 and abs(mt.UCT_COEFF - 0.2 * 2 ** -0.5) < 1e-15",
        "for name in ('Forest', 'create_simulators', 'initialize_simulators', 'load_scenarios'): assert hasattr(ft, name), name",
        "assert hasattr(master.MasterTree, 'get_best_policy') and hasattr(master_node, 'deepcopy_dict')",
        "assert hasattr(isochrones, 'Isochrone') and hasattr(isochrones, 'estimate_perfomance_plan') and hasattr(utils, 'Hist')",
        "print('swap ok')"])
    res = subprocess.run([sys.executable, "-c", check], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300,
                         cwd=ROOT)
    assert res.returncode == 0 and "swap ok" in res.stdout, res.stderr[-2000:]
