"""The reference's walk-through (code/Tutorial.py) run against this package through the module aliases of
INTEGRATION.md section A: the script's own import lines and calls, synthetic weather instead of the NOMADS
download, plotting calls left out."""
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAMES = ("weatherTLKT", "simulatorTLKT", "utils", "worker", "master_node", "forest", "master", "isochrones")


@pytest.fixture()
def reference_names():
    import importlib
    saved = {n: sys.modules.get(n) for n in NAMES}
    for n in NAMES:
        sys.modules[n] = importlib.import_module("iboat_pmcts_b200." + n)
    yield
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m


def test_tutorial_script_runs_unchanged(reference_names, capsys):
    # ---- the import lines of Tutorial.py:9-24 -------------------------------------------------------
    import forest as ft
    import worker as mt
    from simulatorTLKT import Boat
    from weatherTLKT import Weather
    from master_node import deepcopy_dict
    from master import MasterTree
    import isochrones as IC
    from iboat_pmcts_b200 import synth

    # ---- "Load the weather scenarios" (Tutorial.py:26-37): synthetic GEFS-shaped members ----------------
    Weathers = []
    for s in range(3):
        t, la, lo, u, v = synth.wind_fields(s)
        Weathers.append(Weather(la, lo, t + 10.0, u, v))   # forecast times do not start at 0: create_simulators shifts them
    # ---- "Create the simulators" (Tutorial.py:39-49) --------------------------------------------------------
    Boat.UNCERTAINTY_COEFF = 0.2
    NUMBER_OF_SIM, SIM_TIME_STEP, STATE_INIT, N_DAYS_SIM = 3, 6, [0, 44., 355.], 3
    Sims = ft.create_simulators(Weathers, numberofsim=NUMBER_OF_SIM, simtimestep=SIM_TIME_STEP,
                                stateinit=STATE_INIT, ndaysim=N_DAYS_SIM)
    assert Weathers[0].time[0] == 0.0 and len(Sims[0].times) == 13
    # ---- "Initialize the simulators" (Tutorial.py:51-56) ------------------------------------------------------
    random.seed(0)
    np.random.seed(0)
    missionheading, ntra = 0, 50
    destination, timemin = ft.initialize_simulators(Sims, ntra, STATE_INIT, missionheading, plot=False)
    assert 46.0 < destination[0] < 48.5 and abs(destination[1] - 355.0) < 1e-9 and 1.5 < timemin < 3.0
    # ---- "Create and launch a forest" (Tutorial.py:58-68) --------------------------------------------------------
    mt.RHO = 0.5
    mt.UCT_COEFF = 1 / 2 ** 0.5
    budget, frequency = 400, 10
    forest = ft.Forest(listsimulators=Sims, destination=destination, timemin=timemin, budget=budget)
    master_nodes = forest.launch_search(STATE_INIT, frequency)
    # ---- "Process the results" (Tutorial.py:70-75) ------------------------------------------------------------------
    new_dict = deepcopy_dict(master_nodes)
    forest.master = MasterTree(Sims, destination, nodes=new_dict)
    forest.master.get_best_policy()
    plan_PMCTS = forest.master.best_policy[-1]
    assert "Global policy" in capsys.readouterr().out
    assert len(plan_PMCTS) >= 2 and int(plan_PMCTS[0]) in (0, 45, 315)
    # ---- "Isochrones" (Tutorial.py:77-92) ---------------------------------------------------------------------------------
    Boat.UNCERTAINTY_COEFF = 0
    sim = ft.create_simulators(Weathers, numberofsim=NUMBER_OF_SIM, simtimestep=SIM_TIME_STEP, stateinit=STATE_INIT,
                               ndaysim=N_DAYS_SIM)[0]
    solver_iso = IC.Isochrone(sim, STATE_INIT, destination, delta_cap=5, increment_cap=9, nb_secteur=300,
                              resolution=100)
    temps_estime, plan_iso, plan_iso_ligne_droite, trajectoire = solver_iso.isochrone_methode()
    assert 1.5 < temps_estime < 3.0 and len(plan_iso) >= 6 and trajectoire[-1] == destination
    # ---- "Comparision" (Tutorial.py:94-104) --------------------------------------------------------------------------------
    Boat.UNCERTAINTY_COEFF = 0.2
    mean_PMCTS, var_PMCTS = IC.estimate_perfomance_plan(Sims, ntra, STATE_INIT, destination, plan_PMCTS, plot=False,
                                                        verbose=False)
    mean_iso, var_iso = IC.estimate_perfomance_plan(Sims, ntra, STATE_INIT, destination, plan_iso, plot=False,
                                                    verbose=False)
    assert len(mean_PMCTS) == 4 and len(var_iso) == 4
    assert all(1.5 < m <= 3.0 for m in mean_PMCTS + mean_iso)
    # the deterministic optimum is a lower bound for its own noisy replays, up to Monte-Carlo noise
    assert mean_iso[1] > temps_estime - 0.1
    # the batched evaluation of the same plans agrees with the reference-order one
    mean_fast, _ = IC.estimate_perfomance_plan(Sims, 20000, STATE_INIT, destination, plan_iso, verbose=False, seed=1)
    assert abs(mean_fast[0] - mean_iso[0]) < 0.1
