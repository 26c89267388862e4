"""TEST INFRASTRUCTURE: an engine whose launches are answered by the fp64 oracle, for CPU tests of the HOST logic of the
mirror classes (noise order, bookkeeping, exceptions, reductions) against the fixtures recorded from the reference.  It
implements the two host-array calls those classes make -- ``step_batch`` and ``rollout_batch_host`` -- with the oracle's
batches (bit-exact with the reference's ``doStep`` / ``default_policy``); the same tests run on the GPU with the kernels
in its place (tests/test_api_gpu.py).  The product never imports this module."""
import numpy as np

from iboat_pmcts_b200 import _native as N
from iboat_pmcts_b200 import simulatorTLKT as SimC


class OracleEngine:
    precision = N.PREC_F64

    def __init__(self, osims):
        """``osims``: {scenario slot: oracle.cport.OracleSim}."""
        self.osims = dict(osims)
        self.launches = 0

    def set_precision(self, mode):
        self.precision = mode

    def _sim(self, scen):
        o = self.osims[int(scen)]
        o.s.sigma_coeff = float(SimC.Boat.UNCERTAINTY_COEFF)   # read at call time, like the reference (and sync_params)
        return o

    def step_batch(self, scen, k, lat, lon, heading, nsteps=1, seg_len=1, status=None, z=None, flags=0, **_):
        assert self.precision == N.PREC_F64, "the reference-order paths run the literal kernel"
        r = self._sim(scen).step_batch(k, lat, lon, heading, nsteps=nsteps, seg_len=seg_len, z=z,
                                       flags=1 if flags & N.STEP_STOP_BEFORE_TERMINAL else 0)
        k[:], lat[:], lon[:] = r["k"], r["lat"], r["lon"]
        if status is not None:
            status[:] = r["status"]
        self.launches += 1

    def rollout_batch_host(self, scen, n_leaf, replicas, root, dest, tmin, prefix=None, prefix_len=None, z=None,
                           traj_steps=0, flags=0, want=None, **_):
        assert self.precision == N.PREC_F64 and replicas == 1
        r = self._sim(scen).rollout_batch(int(n_leaf), root, dest, tmin, prefix=prefix, prefix_len=prefix_len,
                                          replay_full=bool(flags & N.ROLL_REPLAY_FULL_PREFIX),
                                          mode=1 if flags & N.ROLL_PLAN_EVAL else 0, z=z, traj_steps=traj_steps)
        self.launches += 1
        keep = set(want or r) | {"traj_lat", "traj_lon"}
        return {name: val for name, val in r.items() if name in keep}


def on_oracle(sims, oracle_sims, monkeypatch):
    """Points the mirror simulators at one OracleEngine (slot i = oracle_sims[i]); returns the engine."""
    eng = OracleEngine({i: o for i, o in enumerate(oracle_sims)})
    for i, sim in enumerate(sims):
        monkeypatch.setattr(sim, "scenario", lambda i=i: (eng, i))
    return eng
