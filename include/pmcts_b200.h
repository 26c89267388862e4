/* pmcts_b200 -- C ABI of the B200-native rollout engine for IBoat-PMCTS.
 *
 * The reference (PBarde/IBoat-PMCTS) is pure Python and has no FFI; the boundary it exposes for
 * this path is its Python object API.  Each entry point below names the reference interface it
 * replaces (file:line in the reference checkout).  The Python mirror of that API
 * (the modules of iboat_pmcts_b200/) binds these symbols with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every call returns 0 on success or a negative PMCTS_ERR_* code; pmcts_last_error() returns the
 *    message of the last failure on the calling thread;
 *  - one pmcts_ctx per GPU; calls on a ctx are ordered on its CUDA stream (pmcts_set_stream);
 *  - array arguments are DEVICE pointers and the call returns without synchronising, unless
 *    PMCTS_MEM_HOST is set in `flags`: then every array argument is HOST memory (pinned or
 *    pageable), the library stages it through device scratch and the call returns after the
 *    results are back in the caller's arrays (or, with PMCTS_ROLL_DEFER_REDO / PMCTS_STEP_DEFER, as
 *    soon as its work is enqueued: see pmcts_wait);
 *  - optional outputs may be NULL;
 *  - there is no CPU fallback: without a CUDA device pmcts_create fails.
 */
#ifndef PMCTS_B200_H
#define PMCTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmcts_ctx pmcts_ctx;

/* error codes */
#define PMCTS_OK 0
#define PMCTS_ERR_CUDA (-1)       /* a CUDA runtime call failed (message has the CUDA error string) */
#define PMCTS_ERR_ARG (-2)        /* invalid argument */
#define PMCTS_ERR_NO_DEVICE (-3)  /* no usable CUDA device: the engine has no CPU path */
#define PMCTS_ERR_AXIS (-4)       /* weather axes not ascending / not uniformly spaced */
#define PMCTS_ERR_SCENARIO (-5)   /* unknown scenario slot */

/* per-boat / per-rollout status byte (what the Python layer turns into the reference's exceptions) */
#define PMCTS_ST_OK 0            /* still sailing */
#define PMCTS_ST_ARRIVED 1       /* Tree.is_state_at_dest true (worker.py:380-415) */
#define PMCTS_ST_HORIZON 2       /* times[k] == times[-1] (worker.py:427) */
#define PMCTS_ST_OUT_OF_BOX 3    /* outside [lats[0],lats[-1]] x [lons[0],lons[-1]] (worker.py:430-434) */
#define PMCTS_ST_OUT_OF_GRID 4   /* SciPy bounds_error -> ValueError (weatherTLKT.py:377-378) */
#define PMCTS_ST_DEGENERATE 5    /* ZeroDivisionError inside is_state_at_dest (worker.py:396-397) */
#define PMCTS_ST_PAST_HORIZON 6  /* IndexError: times[k+1] past the axis (simulatorTLKT.py:208) */

/* flags */
#define PMCTS_STEP_STOP_BEFORE_TERMINAL 0x1 /* step_batch: fixed-heading sweep of initialize_simulators (forest.py:239) */
#define PMCTS_STEP_HEADING_ACTIONS 0x2      /* step_batch: `heading` points to uint8_t[nseg][n] action indices instead of
                                               fp64 degrees: heading = 45 a degrees, the reference's ACTIONS
                                               (simulatorTLKT.py:29) -- an eighth of the bytes for a host caller */
#define PMCTS_STEP_DEFER 0x4                /* step_batch with PMCTS_MEM_HOST: the call returns once its work is enqueued;
                                               copies in, kernel and copies out run on three streams, so the stages of
                                               consecutive calls overlap.  Results are in the caller's arrays after
                                               pmcts_wait (or pmcts_synchronize); the arrays must stay alive until
                                               then.  Two calls may be in flight. */
#define PMCTS_ROLL_REPLAY_FULL_PREFIX 0x2   /* rollout_batch: replay the whole prefix (fixes quirk Q1, worker.py:297) */
#define PMCTS_ROLL_PLAN_EVAL 0x4            /* rollout_batch: estimate_perfomance_plan semantics (isochrones.py:498-530) */
#define PMCTS_ROLL_ACCUMULATE 0x8           /* rollout_batch, device arrays only: add to leaf_bins / stats instead of zeroing them first */
#define PMCTS_ROLL_DEFER_REDO 0x10          /* rollout_batch / rollout_forest, PMCTS_PREC_FAST: the fp64 finisher, the
                                               statistics fold and -- with PMCTS_MEM_HOST -- the copies back into the
                                               caller's arrays run on a side stream, so they overlap the NEXT launch on
                                               this ctx.  The outputs of this call are complete only after pmcts_join
                                               (device arrays: orders the stream), pmcts_wait (blocks the host: what a
                                               caller with host arrays needs) or pmcts_synchronize; host arrays must
                                               stay alive until then.  Two launches may be in flight. */
#define PMCTS_ROLL_PREFIX_ACTIONS 0x20       /* rollout_batch / rollout_forest: `prefix` points to int8_t[..][prefix_stride] indices into
                                               the reference's ACTIONS (simulatorTLKT.py:29; heading = 45 a degrees, -1 pads a path)
                                               instead of fp64 degrees: what the host tree hands over, an eighth of the bytes */
#define PMCTS_MEM_HOST 0x100                /* array arguments are host memory */

/* arithmetic mode */
#define PMCTS_PREC_F64 0   /* literal fp64 transcription of the reference formulas */
#define PMCTS_PREC_FAST 1  /* fp32 arithmetic on re-derived, difference-form formulas; positions within 1e-6
                              relative of the fp64 path; every decision (arrival, terminal box, histogram bin)
                              closer than its error margin is re-taken by the fp64 kernel (pmcts_fast.cuh) */

/* noise source for Boat.addUncertainty (simulatorTLKT.py:504-517): speed + z * (speed * sigma) */
#define PMCTS_NOISE_INJECTED 0 /* z[step * n + i], fp64 (parity tests) */
#define PMCTS_NOISE_PHILOX 1   /* Philox4x32-10, counter (id0 + i, step / 4, stream), key seed */
#define PMCTS_NOISE_NONE 2     /* z = 0 (UNCERTAINTY_COEFF = 0, Tutorial.py:81) */
typedef struct {
    int32_t mode;
    int32_t reserved;
    const double *z;
    uint64_t seed, stream, id0;
} pmcts_noise;

/* ---- context ------------------------------------------------------------------------------- */
int pmcts_create(int device, pmcts_ctx **out);
void pmcts_destroy(pmcts_ctx *ctx);
const char *pmcts_last_error(void);
const char *pmcts_version(void);
/* cudaStream_t to run on (NULL = the CUDA default stream); a new ctx runs on its own non-blocking
 * stream, which pmcts_reset_stream restores. */
int pmcts_set_stream(pmcts_ctx *ctx, void *cuda_stream);
int pmcts_reset_stream(pmcts_ctx *ctx);
/* The cudaStream_t the ctx currently runs on, and its device ordinal (for callers that enqueue their own work
 * -- copies, collectives -- in order with the library's). */
void *pmcts_get_stream(pmcts_ctx *ctx);
int pmcts_get_device(pmcts_ctx *ctx);
int pmcts_synchronize(pmcts_ctx *ctx);
/* Makes the ctx stream wait for the deferred finishers (PMCTS_ROLL_DEFER_REDO) of all but the `keep_last`
 * most recent launches: work enqueued on the stream afterwards sees their complete outputs. */
int pmcts_join(pmcts_ctx *ctx, int keep_last);
/* Blocks the calling thread until the deferred launches (PMCTS_ROLL_DEFER_REDO) of all but the `keep_last` most
 * recent calls are complete, copies into host arrays included (rollout and PMCTS_STEP_DEFER step calls are counted
 * separately). */
int pmcts_wait(pmcts_ctx *ctx, int keep_last);
int pmcts_set_precision(pmcts_ctx *ctx, int mode);
/* Decision margins of PMCTS_PREC_FAST (a negative value keeps the current one; all negative restores
 * the defaults): relative margin on sin^2 of the arrival angle vs sin^2(DESTINATION_ANGLE); margin on the
 * position of the destination along the last step (units of the step); (distance to the destination /
 * step)^2 under which a missed arrival is re-run; terminal-box margin in grid cells; margin on reward * 8
 * against the histogram bin edges.  Huge margins send every rollout through the fp64 kernel. */
int pmcts_set_fast_margins(pmcts_ctx *ctx, double angle, double along, double near_miss2, double box,
                           double bin);
/* Rollouts (or K1 hand-over flag) the last FAST launch left to the fp64 kernel; synchronises the stream. */
int64_t pmcts_last_redo_count(pmcts_ctx *ctx);
/* Engine options.  PMCTS_OPT_STAGE_SLABS (default 0): 1 makes K1 in PMCTS_PREC_FAST stage the wind slab of the
 * current step in shared memory (TMA bulk copy into a two-slot mbarrier ring) for blocks whose boats share
 * their time index; 0 reads every gather through L1 / L2.  Same results bit for bit; measured slower on a
 * B200 (the per-step block barrier costs more than the L1 hits it replaces), kept for grids that miss L1. */
#define PMCTS_OPT_STAGE_SLABS 1
/* PMCTS_OPT_PACKED_FP32 (default 0): 1 makes K2's turbo variant carry TWO rollouts per thread, all their FP32
 * arithmetic issued as packed FFMA2 / FMUL2 / FADD2 (csrc/pmcts_fast2.cuh).  Same histograms and statistics; 18 % fewer
 * instructions, but at 128 registers only 4 warps per scheduler stay resident and the kernel measured 5 % slower on a
 * B200 than the one-rollout-per-thread form (which packs the pairs a single rollout offers), so it is off. */
#define PMCTS_OPT_PACKED_FP32 2
int pmcts_set_option(pmcts_ctx *ctx, int option, int value);
/* Which compile-time variant of the K1 / K2 kernel the last pmcts_step_batch / pmcts_rollout_* call launched.
 * All variants give the same results (tests/test_fastmodel.py, tests/test_gpu_parity.py); the specialised ones
 * only drop work the host proved redundant:
 *   LITERAL       PMCTS_PREC_F64, or a scenario / launch the mixed-precision path does not cover
 *   FULL_OUTPUT   mixed precision with per-rollout / per-step outputs requested
 *   GENERIC       mixed precision, histogram / statistics outputs only, any pmcts_set_params values
 *   DEFAULT_BOAT  + the reference's boat, geometry constants and default margins as instruction immediates
 *   TURBO         + (K2, tree mode) equal steps, at most 16 of them, terminal box inside the grid, root able to
 *                 start a step: range checks, bounds test, step-length load, sin / cos refresh compiled out
 *   STAGED        K1 with PMCTS_OPT_STAGE_SLABS
 *   TURBO_PACKED  TURBO with PMCTS_OPT_PACKED_FP32 */
#define PMCTS_VARIANT_LITERAL 0
#define PMCTS_VARIANT_FULL_OUTPUT 1
#define PMCTS_VARIANT_GENERIC 2
#define PMCTS_VARIANT_DEFAULT_BOAT 3
#define PMCTS_VARIANT_TURBO 4
#define PMCTS_VARIANT_STAGED 5
#define PMCTS_VARIANT_TURBO_PACKED 6 /* TURBO, two rollouts per thread on the packed FP32 pipe */
int pmcts_last_kernel_variant(pmcts_ctx *ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches). */
int64_t pmcts_launch_count(pmcts_ctx *ctx);
/* Measures what the FP32 FMA pipes of the ctx's device sustain (16 independent FFMA chains per thread, best of
 * four timed launches, CUDA events): the denominator of bench.py's roofline next to the nominal figure. */
int pmcts_probe_fp32_peak(pmcts_ctx *ctx, double *tflops);
/* While timing is enabled every kernel launch is bracketed by a cudaEvent pair on the ctx stream;
 * pmcts_kernel_ms* synchronise the stream and return the accumulated device milliseconds (all
 * kernels, or one PMCTS_KERNEL_* kind). */
int pmcts_enable_timing(pmcts_ctx *ctx, int on);
double pmcts_kernel_ms_total(pmcts_ctx *ctx, int reset);
#define PMCTS_KERNEL_STEP 0
#define PMCTS_KERNEL_ROLLOUT 1
#define PMCTS_KERNEL_BACKUP 2
#define PMCTS_KERNEL_OTHER 3
#define PMCTS_KERNEL_FINISHER 4 /* the fp64 kernel on the work list of a PMCTS_PREC_FAST rollout launch */
double pmcts_kernel_ms(pmcts_ctx *ctx, int kind, int reset);

/* Module globals the reference reads at call time -- Boat.FIT_VELOCITY / POLAR_* /
 * UNCERTAINTY_COEFF (simulatorTLKT.py:454-471), DESTINATION_ANGLE (:43), EARTH_RADIUS (:40). */
int pmcts_set_params(pmcts_ctx *ctx, const double polar[36], double wind_max, double pofsail_min,
                     double pofsail_max, double sigma_coeff, double dest_angle, double earth_radius);

/* ---- scenarios ------------------------------------------------------------------------------
 * Replaces Weather.__init__ + Weather.Interpolators (weatherTLKT.py:40-49, 369-378) and
 * Simulator.__init__ (simulatorTLKT.py:75-88) for scenario slot `scen`.
 * Weather axes ascending and uniformly spaced; u, v are [nt][nlat][nlon] host arrays, fp64
 * (dtype 0) or fp32 (dtype 1).  sim_times[nT] is the simulator axis in days; box = {lats[0],
 * lats[-1], lons[0], lons[-1]} of the Simulator (the terminal box of worker.py:430-434).
 * On the device the scenario is stored as one time-blended (u, v) slab per simulator instant, plus the
 * raw grid (for pmcts_interp_wind).  nT = 0 (sim_times / box may be NULL) uploads the weather alone:
 * such a slot serves pmcts_interp_wind only. */
int pmcts_scenario_upload(pmcts_ctx *ctx, int scen, const double *wtime, int nt, const double *wlat,
                          int nlat, const double *wlon, int nlon, const void *u, const void *v,
                          int dtype, const double *sim_times, int nT, const double box[4]);
int pmcts_scenario_free(pmcts_ctx *ctx, int scen);
/* Simulator.getWind (simulatorTLKT.py:168-179) for n query points at time index k[i]. */
int pmcts_get_wind(pmcts_ctx *ctx, int scen, int64_t n, const int32_t *k, const double *lat,
                   const double *lon, double *u, double *v, uint8_t *status, int flags);

/* Weather.uInterpolator / vInterpolator (weatherTLKT.py:369-378) at n arbitrary points (t days, lat, lon):
 * SciPy's linear RegularGridInterpolator, same evaluation order, bit-for-bit; out-of-bounds points get
 * NaN and PMCTS_ST_OUT_OF_GRID (SciPy raises ValueError). */
int pmcts_interp_wind(pmcts_ctx *ctx, int scen, int64_t n, const double *t, const double *lat,
                      const double *lon, double *u, double *v, uint8_t *status, int flags);

/* ---- K1: Simulator.doStep (simulatorTLKT.py:181-216) for n boats, nsteps fused ---------------
 * State is structure-of-arrays: k[n] time index, lat[n], lon[n] degrees, prev_lat/prev_lon[n]
 * (Simulator.prevState), updated in place.  heading[(j / seg_len) * n + i] is the heading in degrees
 * of boat i at step j.  A boat whose status is or becomes non-zero is frozen.  traj_lat / traj_lon
 * (optional) receive the position after every step, [nsteps][n] (NaN where a boat was frozen). */
int pmcts_step_batch(pmcts_ctx *ctx, int scen, int64_t n, int nsteps, int32_t *k, double *lat,
                     double *lon, double *prev_lat, double *prev_lon, const double *heading,
                     int seg_len, const pmcts_noise *noise, uint8_t *status, double *traj_lat,
                     double *traj_lon, int flags);

/* ---- K2 / K4 / K5: Tree.get_sim_to_estimate_state + Tree.default_policy (worker.py:255-300),
 * estimate_perfomance_plan (isochrones.py:491-530), initialize_simulators phase 2 (forest.py:261-299)
 * n_leaf * replicas rollouts, rollout r = leaf * replicas + rep, all starting at root = {k, lat, lon}.
 * prefix[leaf * prefix_stride + a] (degrees), prefix_len[leaf] actions (NULL: 0 everywhere).
 * Outputs per rollout (all optional): reward, t_arr (days; NaN if not arrived, times[-1] in plan
 * mode), steps, status, bin (Hist bin of the reward, utils.py:34-47), frac, final_lat, final_lon;
 * traj_lat/traj_lon [traj_stride][n] positions after each transition (NaN where not reached).
 * leaf_bins (optional) [n_leaf][12] uint32: per-leaf histogram of the replicas' reward bins.
 * stats (optional) [8] doubles over all rollouts: {rollouts with an arrival time, sum t_arr,
 * sum t_arr^2, arrived, transitions, rollouts, 0, 0}. */
int pmcts_rollout_batch(pmcts_ctx *ctx, int scen, int64_t n_leaf, int replicas, const double root[3],
                        const double dest[2], double tmin, const double *prefix,
                        const int32_t *prefix_len, int prefix_stride, const pmcts_noise *noise,
                        double *reward, double *t_arr, int32_t *steps, uint8_t *status, int32_t *bin,
                        double *frac, double *final_lat, double *final_lon, double *traj_lat,
                        double *traj_lon, int traj_stride, uint32_t *leaf_bins, double *stats,
                        int flags);

/* K2 for several scenarios in ONE launch -- what Forest.launch_search (forest.py:45-81) spreads over one
 * process per scenario: rollouts of the scenario slots scen[0 .. n_scen) (a HOST array, n_scen <= 16; the
 * slots must share their axes, terminal box and simulator instants), n_leaf * replicas rollouts each,
 * same root / dest / tmin.  Arrays carry a leading scenario dimension: prefix [n_scen][n_leaf][stride]
 * and prefix_len [n_scen][n_leaf] (or one copy for all when prefix_shared != 0), outputs [n_scen][n],
 * leaf_bins [n_scen][n_leaf][12], stats [n_scen][8], injected noise z [n_scen][nT][n].  Scenario i draws
 * its Philox noise from stream noise->stream + i. */
int pmcts_rollout_forest(pmcts_ctx *ctx, int n_scen, const int32_t *scen, int64_t n_leaf, int replicas,
                         const double root[3], const double dest[2], double tmin, const double *prefix,
                         const int32_t *prefix_len, int prefix_stride, int prefix_shared,
                         const pmcts_noise *noise, double *reward, double *t_arr, int32_t *steps,
                         uint8_t *status, int32_t *bin, uint32_t *leaf_bins, double *stats, int flags);

/* ---- K3: Node.back_up (worker.py:55-70) / MasterNode.add_reward* + the ancestor walk of
 * Tree.integrate_buffer (master_node.py:34-54, worker.py:358-378) -------------------------------
 * table[scen][node][8][12] uint32 visit counts for n_scen scenarios of n_nodes nodes each (a worker
 * tree is the n_scen = 1 case; the replicated master table has one slice per scenario).  The tree
 * shape is shared: parent[node] (-1 for the root) and arm[node] = index in ACTIONS of the action
 * that expanded `node`.  For scenario s and leaf l, with c = leaf_bins[s][l][0..12):
 *   table[s][leaf_node[s][l]][leaf_slot[s][l]] += c        (the random slot of worker.py:62-63)
 *   table[s][parent[x]][arm[x]] += c  for x = leaf, parent(leaf), ... up to the root's child.
 * One launch covers all scenarios; additions are integer atomics, so the result is exact and
 * independent of the order.
 * PMCTS_BACKUP_BINS_U8 / _U16: leaf_bins points to uint8 / uint16 counts (the form histograms travel in between
 * GPUs: pmcts_pack_bins). */
#define PMCTS_BACKUP_BINS_U8 0x1
#define PMCTS_BACKUP_BINS_U16 0x2
int pmcts_backup(pmcts_ctx *ctx, uint32_t *table, int64_t n_nodes, int n_scen, const int32_t *parent,
                 const int8_t *arm, int64_t n_leaf, const int32_t *leaf_node, const int8_t *leaf_slot,
                 const uint32_t *leaf_bins, int flags);
/* Narrows `count` uint32 histogram counters (device) to `width` = 1 or 2 bytes each (device): what a wave's
 * histograms are all-gathered as -- a count never exceeds the number of replicas per leaf. */
int pmcts_pack_bins(pmcts_ctx *ctx, const uint32_t *bins, int64_t count, void *out, int width);

/* Per-(node, action) binned means and visit totals of a histogram table (Hist.get_mean,
 * utils.py:49-59; Node.get_uct's exploitation term, worker.py:91-104):
 * mean[node][8] = sum(h * centre) / sum(h) (0 if empty), visits[node] = sum over 8x12. */
int pmcts_hist_stats(pmcts_ctx *ctx, const uint32_t *table, int64_t n_nodes, double *mean,
                     uint32_t *visits, int flags);

/* ---- host runtime of the leaf-batched forest search -------------------------------------------------
 * Array-based worker trees (one per local scenario) and the replicated master tree, with the selection /
 * expansion / back-up rules of the reference's Python classes: Node, Tree.tree_policy / expand / best_child /
 * get_master_uct (worker.py:27-253, 302-346), Node.back_up (worker.py:55-70), Tree.integrate_buffer and
 * MasterNode.add_reward* (worker.py:348-378, master_node.py:34-54).  Host code: it picks the leaves the GPU
 * rolls out (pmcts_rollout_forest) and files the histograms that come back.  All arrays are HOST memory.
 * Actions are indices into ACTIONS (0..7).
 *
 * create: n_scen scenarios in total, n_local trees on this rank for the scenarios local_ids[]; a node at depth
 *   max_depth is terminal (Tree.is_node_terminal, worker.py:438-445: len(times) - 1); perms[n_local][max_nodes][8]
 *   gives every node, in creation order, its shuffled list of untried actions (worker.py:50-51; consumed from
 *   the end like list.pop()); the caller keeps perms alive.  perms = NULL: a forest for
 *   pmcts_forest_search_sequential, which shuffles with the caller's generator as nodes are created.
 * select: n_leaves runs of the tree policy per local tree, a virtual visit left on each selected path;
 *   outputs [n_local][n_leaves]: the leaf's node index, its depth, its action path (prefix_stride >= max_depth
 *   entries, -1 padded).
 * backup: the wave's histograms leaf_bins[n_local][n_leaves][12] and random action slots (worker.py:62) along
 *   the ancestor chains; clears the virtual visits.
 * master_apply: update blocks of any rank (scenario ids, paths, slots of master_node.py:42, histograms) into
 *   the master tree, creating missing nodes ancestors first; call it with the same blocks in the same order
 *   on every rank. */
typedef struct pmcts_forest pmcts_forest;
int pmcts_forest_create(int n_scen, int n_local, const int32_t *local_ids, int max_depth,
                        const double *probability, double uct_coeff, double rho, const uint8_t *perms,
                        int64_t max_nodes, pmcts_forest **out);
void pmcts_forest_destroy(pmcts_forest *f);
const char *pmcts_forest_last_error(void);
int pmcts_forest_set_coefficients(pmcts_forest *f, double uct_coeff, double rho);
int pmcts_forest_select(pmcts_forest *f, int n_leaves, int prefix_stride, int32_t *leaf_node,
                        int32_t *prefix_len, int8_t *prefix_act);
int pmcts_forest_backup(pmcts_forest *f, int n_leaves, const uint32_t *leaf_bins, const int8_t *slots);
int pmcts_forest_master_apply(pmcts_forest *f, int n_blocks_scen, const int32_t *scen_ids, int n_leaves,
                              int prefix_stride, const int32_t *prefix_len, const int8_t *prefix_act,
                              const int8_t *slots, const uint32_t *leaf_bins);
int64_t pmcts_forest_tree_size(pmcts_forest *f, int local_index);
int64_t pmcts_forest_master_size(pmcts_forest *f);
/* out[count][n_items] = `count` consecutive random.shuffle(list(range(n_items))) of CPython's
 * random.Random(seed): the action order Node.__init__ draws for every new node (worker.py:50-51), bit for bit. */
int pmcts_python_shuffles(uint64_t seed, int64_t count, int n_items, uint8_t *out);
/* The same for n_seeds generators (the worker trees of one search, worker.py:50-51 per tree), filled side by side on
 * the host runtime's threads: out[n_seeds][count][n_items]. */
int pmcts_python_shuffles_batch(const uint64_t *seeds, int n_seeds, int64_t count, int n_items, uint8_t *out);
/* Copies of a worker tree (parent, arm, table[node][8][12] int64, untried actions) and of the master tree
 * (parent, arm, table[node][n_scen][8][12] uint32); any pointer may be NULL. */
int pmcts_forest_export_tree(pmcts_forest *f, int local_index, int32_t *parent, int8_t *arm, int64_t *table,
                             uint8_t *n_untried, uint8_t *untried);
int pmcts_forest_export_master(pmcts_forest *f, int32_t *parent, int8_t *arm, uint32_t *table);
/* The master as the runtime keeps it -- a node near the root is visited by every scenario, a deep one by the tree
 * that grew it, so a (node, scenario) pair owns an 8 x 12 row only once that scenario has filed a histogram there:
 * row_index[node][n_scen] (-1: no row) into rows[pmcts_forest_master_rows()][8][12] uint32.  This is the layout
 * pmcts_master_policy reads. */
int64_t pmcts_forest_master_rows(pmcts_forest *f);
int pmcts_forest_export_master_rows(pmcts_forest *f, int32_t *parent, int8_t *arm, int32_t *row_index, uint32_t *rows);

/* ---- MasterTree.get_utility / guess_reward / get_best_child / get_best_policy (master.py:57-185) ------------
 * over the replicated master table, for every node, every scenario and the probability-weighted global tree
 * (idscenario = -1) at once.  Inputs in the layout of pmcts_forest_export_master_rows: parent[n_nodes] (-1 for the
 * root; a parent has a smaller index than its children; children are ranked by index, i.e. creation order),
 * arm[n_nodes], row_index[n_nodes][n_scen], rows[n_rows][8][12], probability[n_scen]; uct_coeff is master.py's
 * UCT_COEFF (bound at import, master.py:11).  Outputs, any of which may be NULL (column n_scen = the global tree):
 *   value, explore [n_nodes][n_scen + 1]   the two terms of get_utility (0, 0 where a scenario never expanded the node)
 *   guess          [n_nodes][n_scen]       guess_reward where the scenario never expanded the node, NaN elsewhere
 *   best_child     [n_nodes][n_scen + 1]   node index get_best_child returns, -1 for None
 *   policy         [n_scen + 1][max_depth] int8 action indices of best_policy, -1 padded; row 0 is the global
 *                                          policy (best_policy[-1]), row 1 + s that of scenario s
 *   policy_nodes   [n_scen + 1][max_depth + 1] int32 node indices of best_nodes_policy (root first), -1 padded
 * PMCTS_MEM_HOST: all arrays are host memory and the call returns with the outputs filled. */
int pmcts_master_policy(pmcts_ctx *ctx, int64_t n_nodes, int n_scen, const int32_t *parent, const int8_t *arm,
                        const int32_t *row_index, int64_t n_rows, const uint32_t *rows, const double *probability,
                        double uct_coeff, int max_depth, double *value, double *explore, double *guess,
                        int32_t *best_child, int8_t *policy, int32_t *policy_nodes, int flags);

/* The same, straight from the master tree of a host forest (after a search): the runtime's rows go to the device
 * through page-locked staging kept with the ctx; the outputs are HOST arrays (any may be NULL). */
int pmcts_forest_master_policy(pmcts_ctx *ctx, pmcts_forest *f, const double *probability, double uct_coeff,
                               int max_depth, double *value, double *explore, double *guess, int32_t *best_child,
                               int8_t *policy, int32_t *policy_nodes);

/* ---- multi-GPU exchange: the all-gather of the per-wave update blocks ------------------------------------
 * Replaces the Manager().dict() proxy through which the reference's worker processes reach the master tree
 * (forest.py:60-66, worker.py:348-378): every rank contributes the block of its own scenarios and receives
 * everybody's, in rank order.  One communicator per ctx (= per GPU, one process per GPU); the collective is
 * NCCL's all-gather over NVLink / NVSwitch (libnccl.so.2, loaded on first use -- the library has no link-time
 * dependency on it), enqueued on the ctx stream.
 *   pmcts_comm_unique_id: on rank 0; hand the 128 bytes to every rank by any channel (torch.distributed
 *     broadcast, MPI, a file) before pmcts_comm_create.
 *   pmcts_allgather_updates: send [bytes] and recv [world * bytes] are DEVICE buffers; returns without
 *     synchronising. */
typedef struct pmcts_comm pmcts_comm;
int pmcts_comm_unique_id(uint8_t id[128]);
int pmcts_comm_create(pmcts_ctx *ctx, int world, int rank, const uint8_t id[128], pmcts_comm **out);
void pmcts_comm_destroy(pmcts_comm *comm);
int pmcts_allgather_updates(pmcts_ctx *ctx, pmcts_comm *comm, const void *send, size_t bytes, void *recv);

/* ---- the same exchange over peer memory: NVLink stores into windows the ranks map from each other ----------
 * (code/solver/worker.py:348-378 again -- every worker hands its buffer of results to all the others.)  One process per
 * GPU on one node.  Every rank owns a window [2][world][block_bytes] in device memory; a put narrows / copies the rank's
 * block and stores it into ALL the windows in one small kernel (the packing pass is the all-gather; nothing of a
 * library runs, no large block has to find room under a resident rollout kernel), then raises the rank's sequence
 * number in every window; pmcts_peer_wait makes ctx's stream wait for the sequence numbers of all ranks in its own
 * window (stream memory operations -- no kernel waits on another GPU) and hands out the gathered blocks in place,
 * laid out like pmcts_allgather_updates' recv.  Two parities alternate: a rank can be one put ahead of the slowest
 * reader, never two, provided every rank orders put k + 1 after its own reading of put k on the same stream.
 *   pmcts_peer_create: allocates the window, returns the 64-byte handle the OTHER processes open it with; hand the
 *     handles of all ranks (rank order, [world][64]) to pmcts_peer_connect by any channel.  world == 1 needs no connect.
 *   pmcts_peer_put_bins: count uint32 counters, narrowed to width 1 or 2 bytes (as pmcts_pack_bins);
 *   pmcts_peer_put: bytes as they are (an update block of pmcts_forest_search).  Sources are 16-byte aligned.
 *   pmcts_peer_poll: host-side check (synchronous, its own stream): how many ranks' blocks of the last put have
 *     arrived in this window -- for bring-up checks with a deadline before anything is made to wait on a stream.
 *   All ranks must make the same sequence of puts.  Destroy only after every rank is done with the windows. */
#define PMCTS_PEER_HANDLE_BYTES 64
typedef struct pmcts_peer pmcts_peer;
int pmcts_peer_create(pmcts_ctx *ctx, int world, int rank, size_t block_bytes, pmcts_peer **out,
                      uint8_t handle[PMCTS_PEER_HANDLE_BYTES]);
int pmcts_peer_connect(pmcts_peer *peer, const uint8_t *handles);
int pmcts_peer_put_bins(pmcts_ctx *ctx, pmcts_peer *peer, const uint32_t *bins, int64_t count, int width);
int pmcts_peer_put(pmcts_ctx *ctx, pmcts_peer *peer, const void *send, size_t bytes);
int pmcts_peer_wait(pmcts_ctx *ctx, pmcts_peer *peer, const void **gathered);
int pmcts_peer_poll(pmcts_ctx *ctx, pmcts_peer *peer, int *arrived);
void pmcts_peer_destroy(pmcts_peer *peer);

/* ---- the leaf-batched search of Forest.launch_search as ONE native call -------------------------------
 * (forest.py:45-81 + Tree.uct_search, worker.py:147-186, restated for waves of leaves.)  Per wave:
 *   host   pmcts_forest_select: n_leaves runs of the tree policy per local tree (virtual visits);
 *   GPU    the leaves' int8 action paths go up (one copy), K2 rolls out replicas x n_leaves x local scenarios in
 *          one launch (PMCTS_PREC_FAST + its fp64 finisher on the side stream), the per-leaf histograms are
 *          packed to uint8 / uint16 counts behind the paths, all-gathered when world > 1 (pmcts_comm), and come
 *          back in one copy;
 *   host   pmcts_forest_backup on the local trees, pmcts_forest_master_apply of every rank's block in rank order.
 * `pipeline` waves are in flight: 1 reproduces the synchronous search (wave w + 1 is selected after wave w has
 * been backed up); 2 selects wave w + 1 on the host while the GPU rolls out wave w (its leaves hold virtual
 * visits until they are backed up).  Rank r owns the scenarios r * n_local .. (r + 1) * n_local - 1 (the forest's
 * local ids) in the ctx slots scen_slots[0 .. n_local).  Scenario s of wave w draws Philox stream w * n_scen + s.
 * slots_w / slots_m: the random action slots of Node.back_up (worker.py:62) and MasterNode.add_reward
 * (master_node.py:42), wave after wave, [n_scen][leaves of that wave] each -- drawn by the caller (the
 * reference uses np.random.randint), for ALL scenarios, identically on every rank.
 * exchange (optional): an all-gather over HOST buffers used instead of pmcts_comm / pmcts_peer (gloo, MPI ...).
 * rollout_hook (test instrumentation; the product passes NULL): stands in for the GPU launch -- fills
 * leaf_bins[n_local][n_leaves][12] for the given paths -- so that the host logic runs without a device;
 * then ctx may be NULL. */
typedef int (*pmcts_exchange_fn)(void *user, const void *send, size_t bytes, void *recv);
typedef int (*pmcts_rollout_hook_fn)(void *user, int wave, int n_local, const int32_t *scen_ids, int n_leaves,
                                     int prefix_stride, const int32_t *prefix_len, const int8_t *prefix_act,
                                     uint32_t *leaf_bins);
/* bytes of one rank's update block (lengths + int8 paths + narrowed counts), the block_bytes of a pmcts_peer that
 * serves such a search */
size_t pmcts_search_block_bytes(int n_local, int n_leaves, int max_depth, int replicas);
typedef struct {
    int32_t n_leaves;   /* leaves per tree per wave */
    int32_t replicas;   /* noise replicas per leaf */
    int64_t budget;     /* leaves per tree in total */
    uint64_t seed;      /* Philox key */
    int32_t roll_flags; /* PMCTS_ROLL_REPLAY_FULL_PREFIX or 0 */
    int32_t pipeline;   /* waves in flight, 1 or 2 */
    int32_t world, rank;
    const int8_t *slots_w, *slots_m;
    pmcts_comm *comm;   /* world > 1 with device buffers: library all-gather ... */
    pmcts_peer *peer;   /* ... or, preferred when set, peer windows of at least pmcts_search_block_bytes() per block */
    pmcts_exchange_fn exchange;
    void *exchange_user;
    pmcts_rollout_hook_fn rollout_hook;
    void *hook_user;
} pmcts_search_params;
/* State of a Mersenne Twister as Python and NumPy hand it out: random.getstate() = (3, key[624] + (pos,), gauss_next)
 * and np.random.get_state() = ('MT19937', key[624], pos, has_gauss, cached_gaussian) -- the latter's Gaussian cache is
 * not used here. */
typedef struct {
    uint32_t key[624];
    int32_t pos;
    int32_t has_gauss; /* random.getstate()[2] is not None */
    double gauss_next;
} pmcts_mt_state;
typedef struct {
    int64_t waves, leaves, rollouts, transitions, arrived, launches;
    double ms_total, ms_select, ms_wait, ms_backup, ms_master, ms_exchange;
} pmcts_search_report;
int pmcts_forest_search(pmcts_ctx *ctx, pmcts_forest *f, const int32_t *scen_slots, const double root[3],
                        const double dest[2], double tmin, const pmcts_search_params *params,
                        pmcts_search_report *report);

/* Tree.uct_search (worker.py:147-186) for the worker tree `local_index` of a forest created WITHOUT perms: the
 * reference's sequential search -- one rollout per iteration, back-up, updates pushed to the master every `frequency`
 * iterations -- in the reference's order and on the reference's random streams: node creation shuffles and the
 * transitions' normals continue Python's generator from *py_state, the random action slots of Node.back_up /
 * MasterNode.add_reward continue NumPy's from *np_state, and both states come back as the reference would have left
 * them, so a seeded script grows the reference's tree.  A rollout is one launch of K2 (PMCTS_PREC_FAST with its fp64
 * re-run: the histogram bin and the number of transitions -- all the tree and the generators depend on -- are exact).
 * Returns 0, a negative error, or the positive PMCTS_ST_* of a rollout the reference would have raised on.
 * Call it for the trees one after the other (forest.py's processes, serialised); report->waves counts master updates.
 * From the second iteration on the device work of an iteration is replayed from a CUDA graph recorded on the ctx
 * stream (not while pmcts_enable_timing is on, nor on a stream that cannot be captured, e.g. the legacy default
 * stream; PMCTS_SEQ_GRAPH=0 in the environment turns it off): the same launches, counted as such in report->launches. */
int pmcts_forest_search_sequential(pmcts_ctx *ctx, pmcts_forest *f, int local_index, int scen_slot,
                                   const double root[3], const double dest[2], double tmin, int64_t budget,
                                   int frequency, int roll_flags, pmcts_mt_state *py_state, pmcts_mt_state *np_state,
                                   pmcts_search_report *report);
/* The same search with the rollouts handed to a callback instead of the GPU -- TEST INSTRUMENTATION, like
 * pmcts_search_params.rollout_hook: it lets the host logic of Tree.uct_search (worker.py:147-253, 348-378: tree policy,
 * node creation shuffles, back-up, master updates, both random streams) be replayed against a recorded run of the
 * reference without a device.  hook(user, iteration (0-based), scenario id, the leaf's action path, its length, the
 * normals z[n_instants] its transitions would draw, &bin, &transitions) returns 0 or an error that ends the search. */
typedef int (*pmcts_seq_rollout_hook_fn)(void *user, int64_t iteration, int scen_id, const int8_t *path, int len,
                                         const double *z, int n_instants, int32_t *bin, int32_t *transitions);
int pmcts_forest_search_sequential_hook(pmcts_forest *f, int local_index, int64_t budget, int frequency,
                                        pmcts_mt_state *py_state, pmcts_mt_state *np_state,
                                        pmcts_seq_rollout_hook_fn hook, void *user, pmcts_search_report *report);
/* The two generators above, for checking them against the interpreter: count x random.gauss(0, 1) and count x
 * np.random.randint(8), continuing the given states. */
int pmcts_python_gauss(pmcts_mt_state *state, int64_t count, double *out);
int pmcts_numpy_randint8(pmcts_mt_state *state, int64_t count, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
