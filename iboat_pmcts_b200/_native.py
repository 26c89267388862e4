"""ctypes binding of libpmcts_b200.so (include/pmcts_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
Python / CPU fallback: if the library is missing, or no CUDA device is visible when an engine is
created, this module raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMCTS_B200_LIB") or os.path.join(HERE, "libpmcts_b200.so")  # (the override: A/B builds)

# mirrors of the header constants
ST_OK, ST_ARRIVED, ST_HORIZON, ST_OUT_OF_BOX, ST_OUT_OF_GRID, ST_DEGENERATE, ST_PAST_HORIZON = range(7)
STEP_STOP_BEFORE_TERMINAL = 0x1
STEP_HEADING_ACTIONS = 0x2
STEP_DEFER = 0x4
ROLL_REPLAY_FULL_PREFIX = 0x2
ROLL_PLAN_EVAL = 0x4
ROLL_ACCUMULATE = 0x8
ROLL_DEFER_REDO = 0x10
ROLL_PREFIX_ACTIONS = 0x20
OPT_STAGE_SLABS = 1
OPT_PACKED_FP32 = 2
(VARIANT_LITERAL, VARIANT_FULL_OUTPUT, VARIANT_GENERIC, VARIANT_DEFAULT_BOAT, VARIANT_TURBO, VARIANT_STAGED,
 VARIANT_TURBO_PACKED) = range(7)
MEM_HOST = 0x100
BACKUP_BINS_U8, BACKUP_BINS_U16 = 0x1, 0x2
PREC_F64, PREC_FAST = 0, 1
NOISE_INJECTED, NOISE_PHILOX, NOISE_NONE = 0, 1, 2
KERNEL_STEP, KERNEL_ROLLOUT, KERNEL_BACKUP, KERNEL_OTHER, KERNEL_FINISHER = 0, 1, 2, 3, 4
ERR_NO_DEVICE = -3
PEER_HANDLE_BYTES = 64

EXPORTS = [
    "pmcts_create", "pmcts_destroy", "pmcts_last_error", "pmcts_version", "pmcts_set_stream", "pmcts_reset_stream",
    "pmcts_get_stream", "pmcts_get_device", "pmcts_comm_unique_id", "pmcts_comm_create", "pmcts_comm_destroy",
    "pmcts_allgather_updates", "pmcts_peer_create", "pmcts_peer_connect", "pmcts_peer_put_bins", "pmcts_peer_put",
    "pmcts_peer_wait", "pmcts_peer_poll", "pmcts_peer_destroy", "pmcts_search_block_bytes", "pmcts_forest_search", "pmcts_master_policy", "pmcts_pack_bins",
    "pmcts_forest_search_sequential", "pmcts_forest_search_sequential_hook", "pmcts_python_gauss", "pmcts_numpy_randint8", "pmcts_forest_master_policy",
    "pmcts_synchronize", "pmcts_join", "pmcts_wait", "pmcts_set_option", "pmcts_set_precision", "pmcts_set_fast_margins", "pmcts_last_redo_count", "pmcts_last_kernel_variant", "pmcts_launch_count", "pmcts_probe_fp32_peak", "pmcts_enable_timing",
    "pmcts_kernel_ms_total", "pmcts_kernel_ms", "pmcts_set_params", "pmcts_scenario_upload", "pmcts_scenario_free",
    "pmcts_get_wind", "pmcts_interp_wind", "pmcts_step_batch", "pmcts_rollout_batch", "pmcts_rollout_forest", "pmcts_backup", "pmcts_hist_stats",
    "pmcts_forest_create", "pmcts_forest_destroy", "pmcts_forest_last_error", "pmcts_forest_set_coefficients",
    "pmcts_forest_select", "pmcts_forest_backup", "pmcts_forest_master_apply", "pmcts_forest_tree_size",
    "pmcts_forest_master_size", "pmcts_forest_export_tree", "pmcts_forest_export_master", "pmcts_python_shuffles", "pmcts_python_shuffles_batch",
    "pmcts_forest_master_rows", "pmcts_forest_export_master_rows",
]


class Noise(C.Structure):
    _fields_ = [("mode", C.c_int32), ("reserved", C.c_int32), ("z", C.c_void_p),
                ("seed", C.c_uint64), ("stream", C.c_uint64), ("id0", C.c_uint64)]


#: int exchange(void *user, const void *send, size_t bytes, void *recv) -- all-gather over host buffers
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
#: int hook(void *user, int wave, int n_local, const int32_t *scen_ids, int n_leaves, int stride,
#:          const int32_t *prefix_len, const int8_t *prefix_act, uint32_t *leaf_bins) -- test stand-in for the GPU launch
ROLLOUT_HOOK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                              C.c_void_p, C.c_void_p)


#: int hook(void *user, int64 iteration, int scen_id, const int8_t *path, int len, const double *z, int n_instants,
#:          int32_t *bin, int32_t *transitions) -- test stand-in for the rollout of the sequential search
SEQ_ROLLOUT_HOOK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int8), C.c_int,
                                  C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32))


class SearchParams(C.Structure):
    """pmcts_search_params (include/pmcts_b200.h)."""
    _fields_ = [("n_leaves", C.c_int32), ("replicas", C.c_int32), ("budget", C.c_int64), ("seed", C.c_uint64),
                ("roll_flags", C.c_int32), ("pipeline", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32),
                ("slots_w", C.c_void_p), ("slots_m", C.c_void_p), ("comm", C.c_void_p), ("peer", C.c_void_p),
                ("exchange", EXCHANGE_FN), ("exchange_user", C.c_void_p),
                ("rollout_hook", ROLLOUT_HOOK_FN), ("hook_user", C.c_void_p)]


class MtState(C.Structure):
    """pmcts_mt_state: a Mersenne Twister as random.getstate() / np.random.get_state() describe it."""
    _fields_ = [("key", C.c_uint32 * 624), ("pos", C.c_int32), ("has_gauss", C.c_int32), ("gauss_next", C.c_double)]

    @classmethod
    def from_python(cls, state):
        version, internal, gauss_next = state
        st = cls()
        st.key[:] = internal[:624]
        st.pos = internal[624]
        st.has_gauss = 0 if gauss_next is None else 1
        st.gauss_next = 0.0 if gauss_next is None else float(gauss_next)
        return st

    def to_python(self):
        return (3, tuple(int(x) for x in self.key) + (int(self.pos),), float(self.gauss_next) if self.has_gauss else None)

    @classmethod
    def from_numpy(cls, state):
        import numpy as np
        name, key, pos = state[0], state[1], state[2]
        if name != "MT19937":
            raise ValueError("np.random must be the legacy MT19937 generator")
        st = cls()
        st.key[:] = [int(x) for x in np.asarray(key, dtype=np.uint32)]
        st.pos = int(pos)
        return st

    def to_numpy(self, like):
        import numpy as np
        return (like[0], np.array(list(self.key), dtype=np.uint32), int(self.pos)) + tuple(like[3:])


class SearchReport(C.Structure):
    """pmcts_search_report."""
    _fields_ = [(k, C.c_int64) for k in ("waves", "leaves", "rollouts", "transitions", "arrived", "launches")] + \
               [(k, C.c_double) for k in ("ms_total", "ms_select", "ms_wait", "ms_backup", "ms_master", "ms_exchange")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PmctsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pmcts_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Loads the CUDA extension; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pmcts_b200: CUDA extension %s is missing -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    L.pmcts_create.argtypes = [i32, C.POINTER(vp)]
    L.pmcts_destroy.argtypes = [vp]
    L.pmcts_destroy.restype = None
    L.pmcts_last_error.restype = C.c_char_p
    L.pmcts_version.restype = C.c_char_p
    L.pmcts_set_stream.argtypes = [vp, vp]
    L.pmcts_reset_stream.argtypes = [vp]
    L.pmcts_get_stream.argtypes = [vp]
    L.pmcts_get_stream.restype = vp
    L.pmcts_get_device.argtypes = [vp]
    L.pmcts_comm_unique_id.argtypes = [vp]
    L.pmcts_comm_create.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
    L.pmcts_comm_destroy.argtypes = [vp]
    L.pmcts_comm_destroy.restype = None
    L.pmcts_allgather_updates.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.pmcts_peer_create.argtypes = [vp, i32, i32, C.c_size_t, C.POINTER(vp), vp]
    L.pmcts_peer_connect.argtypes = [vp, vp]
    L.pmcts_peer_put_bins.argtypes = [vp, vp, vp, i64, i32]
    L.pmcts_peer_put.argtypes = [vp, vp, vp, C.c_size_t]
    L.pmcts_peer_wait.argtypes = [vp, vp, C.POINTER(vp)]
    L.pmcts_peer_poll.argtypes = [vp, vp, C.POINTER(i32)]
    L.pmcts_peer_destroy.argtypes = [vp]
    L.pmcts_peer_destroy.restype = None
    L.pmcts_search_block_bytes.argtypes = [i32, i32, i32, i32]
    L.pmcts_search_block_bytes.restype = C.c_size_t
    L.pmcts_master_policy.argtypes = [vp, i64, i32, vp, vp, vp, i64, vp, vp, dbl, i32, vp, vp, vp, vp, vp, vp, i32]
    L.pmcts_forest_search_sequential.argtypes = [vp, vp, i32, i32, vp, vp, dbl, i64, i32, i32, C.POINTER(MtState),
                                                 C.POINTER(MtState), C.POINTER(SearchReport)]
    L.pmcts_forest_search_sequential_hook.argtypes = [vp, i32, i64, i32, C.POINTER(MtState), C.POINTER(MtState),
                                                      SEQ_ROLLOUT_HOOK_FN, vp, C.POINTER(SearchReport)]
    L.pmcts_python_gauss.argtypes = [C.POINTER(MtState), i64, vp]
    L.pmcts_numpy_randint8.argtypes = [C.POINTER(MtState), i64, vp]
    L.pmcts_forest_master_policy.argtypes = [vp, vp, vp, dbl, i32, vp, vp, vp, vp, vp, vp]
    L.pmcts_forest_search.argtypes = [vp, vp, vp, vp, vp, dbl, C.POINTER(SearchParams), C.POINTER(SearchReport)]
    L.pmcts_synchronize.argtypes = [vp]
    L.pmcts_join.argtypes = [vp, i32]
    L.pmcts_wait.argtypes = [vp, i32]
    L.pmcts_set_option.argtypes = [vp, i32, i32]
    L.pmcts_set_precision.argtypes = [vp, i32]
    L.pmcts_set_fast_margins.argtypes = [vp, dbl, dbl, dbl, dbl, dbl]
    L.pmcts_last_redo_count.argtypes = [vp]
    L.pmcts_last_kernel_variant.argtypes = [vp]
    L.pmcts_probe_fp32_peak.argtypes = [vp, vp]
    L.pmcts_last_redo_count.restype = i64
    L.pmcts_launch_count.argtypes = [vp]
    L.pmcts_launch_count.restype = i64
    L.pmcts_enable_timing.argtypes = [vp, i32]
    L.pmcts_kernel_ms_total.argtypes = [vp, i32]
    L.pmcts_kernel_ms_total.restype = dbl
    L.pmcts_kernel_ms.argtypes = [vp, i32, i32]
    L.pmcts_kernel_ms.restype = dbl
    L.pmcts_set_params.argtypes = [vp, vp, dbl, dbl, dbl, dbl, dbl, dbl]
    L.pmcts_scenario_upload.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, i32, vp, i32, vp]
    L.pmcts_scenario_free.argtypes = [vp, i32]
    L.pmcts_get_wind.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp, i32]
    L.pmcts_interp_wind.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp, i32]
    L.pmcts_step_batch.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, i32, C.POINTER(Noise),
                                   vp, vp, vp, i32]
    L.pmcts_rollout_batch.argtypes = [vp, i32, i64, i32, vp, vp, dbl, vp, vp, i32, C.POINTER(Noise),
                                      vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32]
    L.pmcts_rollout_forest.argtypes = [vp, i32, vp, i64, i32, vp, vp, dbl, vp, vp, i32, i32, C.POINTER(Noise),
                                       vp, vp, vp, vp, vp, vp, vp, i32]
    L.pmcts_backup.argtypes = [vp, vp, i64, i32, vp, vp, i64, vp, vp, vp, i32]
    L.pmcts_hist_stats.argtypes = [vp, vp, i64, vp, vp, i32]
    L.pmcts_pack_bins.argtypes = [vp, vp, i64, vp, i32]
    L.pmcts_forest_create.argtypes = [i32, i32, vp, i32, vp, dbl, dbl, vp, i64, C.POINTER(vp)]
    L.pmcts_forest_destroy.argtypes = [vp]
    L.pmcts_forest_destroy.restype = None
    L.pmcts_forest_last_error.restype = C.c_char_p
    L.pmcts_forest_set_coefficients.argtypes = [vp, dbl, dbl]
    L.pmcts_forest_select.argtypes = [vp, i32, i32, vp, vp, vp]
    L.pmcts_forest_backup.argtypes = [vp, i32, vp, vp]
    L.pmcts_forest_master_apply.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp]
    L.pmcts_forest_tree_size.argtypes = [vp, i32]
    L.pmcts_forest_tree_size.restype = i64
    L.pmcts_forest_master_size.argtypes = [vp]
    L.pmcts_forest_master_size.restype = i64
    L.pmcts_python_shuffles.argtypes = [C.c_uint64, i64, i32, vp]
    L.pmcts_python_shuffles_batch.argtypes = [vp, i32, i64, i32, vp]
    L.pmcts_forest_export_tree.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.pmcts_forest_export_master.argtypes = [vp, vp, vp, vp]
    L.pmcts_forest_master_rows.argtypes = [vp]
    L.pmcts_forest_master_rows.restype = i64
    L.pmcts_forest_export_master_rows.argtypes = [vp, vp, vp, vp, vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise PmctsError(rc, lib().pmcts_last_error().decode("utf-8", "replace"))


def check_forest(rc):
    if rc != 0:
        raise PmctsError(rc, lib().pmcts_forest_last_error().decode("utf-8", "replace"))
