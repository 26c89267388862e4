"""N3 / N4 around the path, against the UNMODIFIED reference (live: runs where /root/reference is mounted, no GPU):

* ``forest.load_scenarios`` (forest.py:118-151) + ``Weather.load / crop`` (weatherTLKT.py:51-78, 165-218) on scenario
  files the REFERENCE's own ``Weather`` class pickled -- GEFS-shaped: fp32 masked fields, masked axes, time in days
  since year 1 -- return the grids the reference's ``load_scenarios`` returns, value for value;
* ``Weather.getPolarVel`` (weatherTLKT.py:135-146) fills the arrays the reference's triple loop fills;
* ``MasterTree.get_points`` (master.py:187-220) lists the reference's segments for the tree plots (objective "depth",
  global tree and one scenario) on a tree the reference loaded from ``save_tree(reference_format=True)``.
"""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="reference not mounted")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, cwd):
    res = subprocess.run([sys.executable, "-W", "ignore", "-c", script], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True, timeout=300, cwd=cwd)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout


def test_load_scenarios_and_polar_velocity_like_the_reference(tmp_path):
    (tmp_path / "data").mkdir()
    (tmp_path / "code").mkdir()
    # the reference writes three scenario files with its own class and reads them back with its own load_scenarios
    script = textwrap.dedent("""
        import pickle, sys
        import numpy as np
        sys.path.insert(0, %r)
        from oracle import refshim
        from iboat_pmcts_b200 import synth
        R = refshim.load_reference()
        for s in range(3):
            t, la, lo, u, v = synth.wind_fields(s)
            w = R.weatherTLKT.Weather(np.ma.masked_array(la), np.ma.masked_array(lo), np.ma.masked_array(t + 736702.0),
                                      np.ma.masked_array(u.astype(np.float32)), np.ma.masked_array(v.astype(np.float32)))
            with open("../data/20180108_%%02d00z.obj" %% s, "wb") as f:
                pickle.dump(w, f)
        out = {}
        for name, kw in (("crop", dict(latBound=[41, 49], lonBound=[346, 359], timeSteps=[2, 30])),
                         ("time", dict(timeSteps=[0, 16])), ("all", dict())):
            ws = R.forest.load_scenarios("20180108", range(0, 3), **kw)
            for s, w in enumerate(ws):
                for a in ("lat", "lon", "time", "u", "v"):
                    out["%%s_%%d_%%s" %% (name, s, a)] = np.ma.getdata(getattr(w, a))
        w = ws[1].crop(latBound=[44, 47], lonBound=[350, 354], timeSteps=[0, 5])
        w.getPolarVel()
        out["wMag"], out["wAng"] = w.wMag, w.wAng
        np.savez("../want.npz", **out)
        """) % ROOT
    _run(script, str(tmp_path / "code"))
    want = np.load(tmp_path / "want.npz")

    from iboat_pmcts_b200 import forest as ft, weatherTLKT
    sys.modules["weatherTLKT"] = weatherTLKT  # (the module alias under which the reference's files unpickle)
    try:
        got = {}
        for name, kw in (("crop", dict(latBound=[41, 49], lonBound=[346, 359], timeSteps=[2, 30])),
                         ("time", dict(timeSteps=[0, 16])), ("all", dict())):
            ws = ft.load_scenarios("20180108", range(0, 3), directory=str(tmp_path / "data") + "/", **kw)
            assert all(type(w) is weatherTLKT.Weather for w in ws)
            for s, w in enumerate(ws):
                for a in ("lat", "lon", "time", "u", "v"):
                    got["%s_%d_%s" % (name, s, a)] = np.ma.getdata(getattr(w, a))
        w = ws[1].crop(latBound=[44, 47], lonBound=[350, 354], timeSteps=[0, 5])
        w.getPolarVel()
        got["wMag"], got["wAng"] = w.wMag, w.wAng
    finally:
        del sys.modules["weatherTLKT"]
    # without the alias too: the reference's class name is read as the mirror class
    assert "weatherTLKT" not in sys.modules or sys.modules["weatherTLKT"] is not weatherTLKT
    plain = ft.load_scenarios("20180108", range(1, 2), directory=str(tmp_path / "data") + "/", timeSteps=[0, 16])[0]
    assert type(plain) is weatherTLKT.Weather and np.array_equal(np.ma.getdata(plain.u), want["time_1_u"])
    assert set(got) == set(want.files)
    for k in want.files:
        assert got[k].shape == want[k].shape and got[k].dtype == want[k].dtype, k
        assert np.array_equal(got[k], want[k]), k
    assert got["crop_0_u"].shape == (28, 15, 25) and got["crop_0_u"].dtype == np.float64
    assert got["all_2_u"].dtype == np.float32 and got["wAng"].shape == (5, 5, 7)


def test_tree_plot_segments_like_the_reference(tmp_path):
    from fake_device import run
    from iboat_pmcts_b200 import master_node as mn, synth
    from iboat_pmcts_b200.forest import ROOT_HASH, create_simulators
    from iboat_pmcts_b200.master import MasterTree
    from iboat_pmcts_b200.weatherTLKT import Weather
    forest, master, _ = run("native", n_scen=3, budget=70)
    nodes = mn.deepcopy_dict(master)
    weathers = [Weather(*[synth.wind_fields(s)[i] for i in (1, 2, 0, 3, 4)]) for s in range(3)]
    sims = create_simulators(weathers, 3, simtimestep=6, stateinit=(0, 44.0, 355.0), ndaysim=3)
    mt = MasterTree(sims, [46.7, 355.0], nodes=nodes)
    mt.save_tree("tree", directory=str(tmp_path) + "/", reference_format=True)
    script = textwrap.dedent("""
        import pickle, sys
        import numpy as np
        sys.path.insert(0, %r)
        from oracle import refshim
        R = refshim.load_reference()
        with open(%r, "rb") as f:
            tree = pickle.load(f)
        root = tree.nodes[hash(())]
        out = {}
        for ids in (-1, 0, 2):
            out["p%%d" %% ids] = np.array(tree.get_points(root, [], tree.probability, idscenario=ids, objective="depth"), float)
        np.savez(%r, **out)
        """) % (ROOT, str(tmp_path / "tree.pickle"), str(tmp_path / "want.npz"))
    _run(script, str(tmp_path))
    want = np.load(tmp_path / "want.npz")
    root = mt.nodes[ROOT_HASH]
    for ids in (-1, 0, 2):
        got = np.array(mt.get_points(root, [], mt.probability, idscenario=ids, objective="depth"), float)
        assert got.shape == want["p%d" % ids].shape and got.shape[1] == 5
        assert np.array_equal(got, want["p%d" % ids]), ids
    assert want["p-1"].shape[0] == len(mt.nodes) - 1 and want["p0"].shape[0] < want["p-1"].shape[0]
