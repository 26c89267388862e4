/* TEST INFRASTRUCTURE ONLY -- CPU fp64 restatement of the IBoat-PMCTS rollout hot path.
 *
 * Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (libpmcts_b200.so) never links it.
 *
 * Parity status: the reference ships no tests or golden vectors ("parity unpinned" by its own
 * suite); this restatement is pinned instead against outputs of the reference
 * itself, imported in the build container (oracle/gen_golden.py -> tests/golden/, and
 * tests/test_oracle_vs_reference.py).
 *
 * All citations are relative to the reference checkout (PBarde/IBoat-PMCTS).
 */
#ifndef PMCTS_ORACLE_H
#define PMCTS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Weather scenario on its native (time, lat, lon) grid: code/model/weatherTLKT.py:40-49. */
typedef struct {
    int nt, nlat, nlon;
    const double *time, *lat, *lon; /* ascending axes */
    const double *u, *v;            /* [nt][nlat][nlon], C order */
} orc_weather;

/* Simulator description: code/model/simulatorTLKT.py:75-88 (times / terminal box) plus the module
 * globals read at call time (Boat.*, DESTINATION_ANGLE, EARTH_RADIUS; simulatorTLKT.py:29-46,454-471). */
typedef struct {
    orc_weather w;
    int nT;
    const double *times;   /* simulator time axis in days */
    double lat_lo, lat_hi; /* lats[0], lats[-1] of the Simulator (worker.py:430) */
    double lon_lo, lon_hi; /* lons[0], lons[-1] (worker.py:433) */
    double polar[36];      /* Boat.FIT_VELOCITY, row i = power of point-of-sail, col j = power of wind */
    double wind_max, pofsail_min, pofsail_max, sigma_coeff;
    double dest_angle, earth_radius;
} orc_sim;

/* Status codes shared with the product ABI (include/pmcts_b200.h). */
enum { ORC_OK = 0, ORC_ARRIVED = 1, ORC_HORIZON = 2, ORC_OUT_OF_BOX = 3, ORC_OUT_OF_GRID = 4,
       ORC_DEGENERATE = 5, ORC_PAST_HORIZON = 6 };

void orc_default_params(orc_sim *s);

/* scalar building blocks */
int orc_interp(const orc_weather *w, double t, double lat, double lon, double *u, double *v);
void orc_polar_vel(double u, double v, double *mag, double *ang);
double orc_deter_dyn(const orc_sim *s, double pofsail, double wmag);
void orc_get_destination(const orc_sim *s, double dist, double bearing, double lat, double lon,
                         double *olat, double *olon);
void orc_dist_bearing(const orc_sim *s, double lat, double lon, double dlat, double dlon,
                      double *dist, double *bearing);
void orc_geo_to_cart(double lat, double lon, double xyz[3]);
int orc_is_at_dest(const orc_sim *s, const double dest[2], const double a[2], const double b[2],
                   double *frac);
int orc_is_terminal(const orc_sim *s, int k, double lat, double lon);
int orc_do_step(const orc_sim *s, int *k, double *lat, double *lon, double action, double z);
int orc_hist_bin(double reward);
double orc_reward(double tmax, double tmin, double t_arr);

/* Philox4x32-10 counter scheme shared with the device kernels. */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_philox_normal(uint64_t seed, uint64_t stream, uint64_t boat, uint32_t step);

/* Batched drivers (OpenMP over boats / rollouts).
 * noise: z != NULL -> injected z[step*n + i]; else Philox(seed, stream, boat=id0+i, step). */
typedef struct {
    const double *z;
    uint64_t seed, stream, id0;
} orc_noise;

/* heading[(j / seg_len) * n + i] degrees; traj_lat/traj_lon optional [nsteps][n].
 * flags: ORC_STEP_STOP_BEFORE_TERMINAL = the fixed-heading sweep of initialize_simulators
 * (forest.py:239-241): a boat stops (status HORIZON / OUT_OF_BOX) as soon as its next state
 * [k+1, lat, lon] would be terminal. */
#define ORC_STEP_STOP_BEFORE_TERMINAL 1
void orc_step_batch(const orc_sim *s, int64_t n, int nsteps, int32_t *k, double *lat, double *lon,
                    double *prev_lat, double *prev_lon, const double *heading, int seg_len,
                    const orc_noise *noise, uint8_t *status, double *traj_lat, double *traj_lon,
                    int flags, int nthreads);

/* One default-policy rollout per i (worker.py:255-300): start at root, apply the prefix actions
 * prefix[i*prefix_stride .. +prefix_len[i]) -- only the first one unless replay_full (quirk Q1) --
 * then head for dest until arrival / terminal.  mode 1 = plan evaluation (isochrones.py:491-530):
 * every prefix action is applied without terminal checks and a missed arrival reports t_arr = times[-1]. */
void orc_rollout_batch(const orc_sim *s, int64_t n, const double root[3], const double dest[2],
                       double tmin, const double *prefix, const int32_t *prefix_len,
                       int prefix_stride, int replay_full, int mode, const orc_noise *noise,
                       double *reward, double *t_arr, int32_t *steps, uint8_t *status,
                       int32_t *bin, double *frac_out, double *final_lat, double *final_lon,
                       double *traj_lat, double *traj_lon, int traj_stride, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
