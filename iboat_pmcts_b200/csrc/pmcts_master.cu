// pmcts_b200 -- the reductions of MasterTree over the replicated master table (pmcts_master_policy):
// get_utility / guess_reward / get_best_child / get_best_policy of the reference (code/solver/master.py:57-185),
// for every node, every scenario and the probability-weighted global tree at once.
//
// Layout: the master as the host runtime keeps it (pmcts_forest_export_master_rows) -- parent[node] (a parent has a
// smaller index than its children), arm[node] = index in ACTIONS of the action that led to the node,
// row_index[node][n_scen] (-1: that scenario never visited the node) into rows[n_rows][8 actions][12 bins] uint32.
//
// What the reference computes, restated per (node, scenario s):
//   expanded       the scenario filed at least one reward in the node              (master_node.py: is_expanded)
//   value          max over the 8 actions of the binned mean, 0 if not expanded     (master.py:140-185)
//   explore        UCT_COEFF * sqrt(2 ln(N_parent) / N_node), N = visit totals (the root: N_parent = N_node)
//   best_child     first child (in creation order) with the largest value > 0      (master.py:86-115)
//   guess          for a node s never expanded: with f its nearest expanded ancestor,
//                  value(u) + explore(u) - explore(f), u = best_child(parent(f))   (master.py:117-138; the
//                  reference hands the number down the greedy path through node.guessed_rewards)
// and for the global tree (idscenario = -1):
//   value          max over actions j of sum_s probability[s] * (expanded ? mean(s, j) : guess(s))
//   explore        the same formula on visit totals summed over the scenarios that expanded the node
//   best_child     first child with the largest global value > 0
// best_policy follows best_child from the root for every scenario and for the global tree.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "pmcts_internal.h"

namespace {

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    pmcts::set_last_error(buf);
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(PMCTS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                    \
    } while (0)

constexpr int kA = 8, kB = 12, kRow = kA * kB;
constexpr int kThreads = 128;

// Hist.get_mean of the 8 histograms of one row (utils.py:49-59): sum h (b + 0.5) / 8 / sum h, exact in fp64 whatever
// the summation order (integers below 2^53 scaled by a power of two).
__device__ __forceinline__ void row_means(const uint32_t *__restrict__ row, double mean[kA], long long &visits) {
    visits = 0;
#pragma unroll
    for (int a = 0; a < kA; ++a) {
        const uint4 *q = reinterpret_cast<const uint4 *>(row + a * kB);
        const uint4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        const uint32_t h[kB] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
        long long tot = 0, acc = 0;
#pragma unroll
        for (int b = 0; b < kB; ++b) {
            tot += h[b];
            acc += (long long)h[b] * (2 * b + 1);
        }
        visits += tot;
        mean[a] = tot > 0 ? ((double)acc * 0.0625) / (double)tot : 0.0;
    }
}

__global__ void row_stats_kernel(const uint32_t *__restrict__ rows, int64_t n_rows, long long *__restrict__ visits,
                                 double *__restrict__ best) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    double mean[kA];
    long long v;
    row_means(rows + r * kRow, mean, v);
    double b = 0.0;
#pragma unroll
    for (int a = 0; a < kA; ++a) b = mean[a] > b ? mean[a] : b;
    visits[r] = v;
    best[r] = b;
}

// child[node][action] from parent / arm
__global__ void children_kernel(const int32_t *__restrict__ parent, const int8_t *__restrict__ arm, int64_t n,
                                int32_t *__restrict__ child) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || parent[i] < 0) return;
    child[(int64_t)parent[i] * kA + arm[i]] = (int32_t)i;
}

// value / explore of every (node, scenario)
__global__ void scenario_utility_kernel(const int32_t *__restrict__ parent, const int32_t *__restrict__ row_index,
                                        const long long *__restrict__ visits, const double *__restrict__ best,
                                        int64_t n, int S, double coeff, double *__restrict__ value,
                                        double *__restrict__ explore) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * S) return;
    const int64_t node = t / S;
    const int s = (int)(t % S);
    const int32_t r = row_index[t];
    double v = 0.0, e = 0.0;
    if (r >= 0 && visits[r] > 0) {
        v = best[r];
        const long long n_node = visits[r];
        long long n_par = n_node;
        const int32_t p = parent[node];
        if (p >= 0) {
            const int32_t rp = row_index[(int64_t)p * S + s];
            n_par = rp >= 0 ? visits[rp] : 0;
        }
        e = coeff * sqrt(2.0 * log((double)n_par) / (double)n_node);
    }
    value[node * (S + 1) + s] = v;
    explore[node * (S + 1) + s] = e;
}

// get_best_child for column `col0 .. col0 + cols` of value[node][S + 1]: first child in creation order (= ascending
// index) with the largest value > 0, -1 if none.
__global__ void best_child_kernel(const int32_t *__restrict__ child, const double *__restrict__ value, int64_t n, int S,
                                  int col0, int cols, int32_t *__restrict__ best_child) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * cols) return;
    const int64_t node = t / cols;
    const int col = col0 + (int)(t % cols);
    double bv = 0.0;
    int32_t bc = -1;
#pragma unroll
    for (int a = 0; a < kA; ++a) {
        const int32_t c = child[node * kA + a];
        if (c < 0) continue;
        const double v = value[(int64_t)c * (S + 1) + col];
        if (v > bv || (v == bv && bc >= 0 && c < bc)) {
            bv = v;
            bc = c;
        }
    }
    best_child[node * (S + 1) + col] = bc;
}

// guess_reward of every (node, scenario) the scenario did not expand (NaN elsewhere)
__global__ void guess_kernel(const int32_t *__restrict__ parent, const int32_t *__restrict__ row_index,
                             const long long *__restrict__ visits, const double *__restrict__ value,
                             const double *__restrict__ explore, const int32_t *__restrict__ best_child, int64_t n, int S,
                             double *__restrict__ guess) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * S) return;
    const int64_t node = t / S;
    const int s = (int)(t % S);
    const int32_t r = row_index[t];
    double g = NAN;
    if (!(r >= 0 && visits[r] > 0)) {
        // nearest ancestor the scenario expanded
        int32_t f = parent[node];
        while (f >= 0) {
            const int32_t rf = row_index[(int64_t)f * S + s];
            if (rf >= 0 && visits[rf] > 0) break;
            f = parent[f];
        }
        if (f >= 0 && parent[f] >= 0) {
            const int32_t u = best_child[(int64_t)parent[f] * (S + 1) + s];
            if (u >= 0) g = value[(int64_t)u * (S + 1) + s] + explore[(int64_t)u * (S + 1) + s] - explore[(int64_t)f * (S + 1) + s];
        }
        // (f = root: the reference asks for the best child of the root's parent and fails; NaN here)
    }
    guess[t] = g;
}

// get_utility(node, -1): one thread per node
__global__ void global_utility_kernel(const int32_t *__restrict__ parent, const int32_t *__restrict__ row_index,
                                      const uint32_t *__restrict__ rows, const long long *__restrict__ visits,
                                      const double *__restrict__ guess, const double *__restrict__ prob, int64_t n, int S,
                                      double coeff, double *__restrict__ value, double *__restrict__ explore) {
    const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n) return;
    double acc[kA];
#pragma unroll
    for (int a = 0; a < kA; ++a) acc[a] = 0.0;
    long long n_node = 0, n_par = 0;
    const int32_t p = parent[node];
    for (int s = 0; s < S; ++s) {
        const int32_t r = row_index[node * S + s];
        const double w = prob[s];
        if (r >= 0 && visits[r] > 0) {
            double mean[kA];
            long long v;
            row_means(rows + (int64_t)r * kRow, mean, v);
#pragma unroll
            for (int a = 0; a < kA; ++a) acc[a] += mean[a] * w;
            n_node += v;
            if (p >= 0) {
                const int32_t rp = row_index[(int64_t)p * S + s];
                n_par += rp >= 0 ? visits[rp] : 0;
            } else {
                n_par += v;
            }
        } else {
            const double g = guess[node * S + s] * w;
#pragma unroll
            for (int a = 0; a < kA; ++a) acc[a] += g;
        }
    }
    double v = acc[0];
#pragma unroll
    for (int a = 1; a < kA; ++a) v = acc[a] > v ? acc[a] : v;  // (NaN guesses propagate as in np.max)
    value[node * (S + 1) + S] = v;
    explore[node * (S + 1) + S] = n_node > 0 ? coeff * sqrt(2.0 * log((double)n_par) / (double)n_node) : NAN;
}

// get_best_policy: thread c follows best_child[.][col(c)] from the root; row 0 = global tree, row 1 + s = scenario s
__global__ void policy_kernel(const int32_t *__restrict__ best_child, const int8_t *__restrict__ arm, int S, int max_depth,
                              int8_t *__restrict__ policy, int32_t *__restrict__ policy_nodes) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > S) return;
    const int col = c == 0 ? S : c - 1;
    int32_t node = 0;
    policy_nodes[(int64_t)c * (max_depth + 1)] = 0;
    int d = 0;
    for (; d < max_depth; ++d) {
        const int32_t nxt = best_child[(int64_t)node * (S + 1) + col];
        if (nxt < 0) break;
        policy[(int64_t)c * max_depth + d] = arm[nxt];
        policy_nodes[(int64_t)c * (max_depth + 1) + d + 1] = nxt;
        node = nxt;
    }
    for (; d < max_depth; ++d) {
        policy[(int64_t)c * max_depth + d] = -1;
        policy_nodes[(int64_t)c * (max_depth + 1) + d + 1] = -1;
    }
}

inline unsigned grid(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

extern "C" int pmcts_master_policy(pmcts_ctx *ctx, int64_t n_nodes, int n_scen, const int32_t *parent, const int8_t *arm,
                                   const int32_t *row_index, int64_t n_rows, const uint32_t *rows,
                                   const double *probability, double uct_coeff, int max_depth, double *value,
                                   double *explore, double *guess, int32_t *best_child, int8_t *policy,
                                   int32_t *policy_nodes, int flags) {
    if (!ctx || !parent || !arm || !row_index || !rows || !probability) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (n_nodes < 1 || n_scen < 1 || n_rows < 0 || max_depth < 1) return fail(PMCTS_ERR_ARG, "bad sizes");
    const bool host = (flags & PMCTS_MEM_HOST) != 0;
    int prev = -1;
    cudaGetDevice(&prev);
    const int dev = pmcts_get_device(ctx);
    if (prev != dev) CUDA_TRY(cudaSetDevice(dev));
    struct Restore {
        int prev, dev;
        ~Restore() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
    } restore{prev, dev};
    cudaStream_t st = (cudaStream_t)pmcts_get_stream(ctx);
    const int64_t n = n_nodes;
    const int S = n_scen, C = n_scen + 1;
    // one stream-ordered allocation for the inputs' device copies and all the scratch (sizes are known up front): a
    // dozen separate allocations and frees were a third of the call's host time on a 10-scenario master
    auto a256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t arena_bytes = a256((size_t)n_rows * sizeof(long long)) + a256((size_t)n_rows * sizeof(double)) +
                         a256((size_t)n * kA * sizeof(int32_t)) + 2 * a256((size_t)n * C * sizeof(double)) +
                         a256((size_t)n * S * sizeof(double)) + a256((size_t)n * C * sizeof(int32_t)) +
                         a256((size_t)C * max_depth) + a256((size_t)C * (max_depth + 1) * sizeof(int32_t)) + 4096;  // (+ room for empty arrays: they still get a slot)
    if (host)
        arena_bytes += a256((size_t)n * sizeof(int32_t)) + a256((size_t)n) + a256((size_t)n * S * sizeof(int32_t)) +
                       a256((size_t)n_rows * kRow * sizeof(uint32_t)) + a256((size_t)S * sizeof(double));
    char *arena = nullptr;
    size_t arena_used = 0;
    if (cudaMallocAsync((void **)&arena, arena_bytes, st) != cudaSuccess)
        return fail(PMCTS_ERR_CUDA, "device scratch for the master reductions: %s", cudaGetErrorString(cudaGetLastError()));
    auto temp = [&](size_t bytes) -> void * {
        const size_t need = a256(bytes ? bytes : 16);
        if (arena_used + need > arena_bytes) return nullptr;
        void *p = arena + arena_used;
        arena_used += need;
        return p;
    };
    auto release = [&]() {
        if (arena) cudaFreeAsync(arena, st);
        arena = nullptr;
    };
#define M_TRY(expr)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            release();                                                                           \
            return fail(PMCTS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));         \
        }                                                                                        \
    } while (0)
    // inputs on the device
    const int32_t *d_parent = parent, *d_row_index = row_index;
    const int8_t *d_arm = arm;
    const uint32_t *d_rows = rows;
    const double *d_prob = probability;
    if (host) {
        auto up = [&](const void *src, size_t bytes) -> void * {
            void *d = temp(bytes);
            if (d && bytes && cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
            return d;
        };
        d_parent = (const int32_t *)up(parent, n * sizeof(int32_t));
        d_arm = (const int8_t *)up(arm, n);
        d_row_index = (const int32_t *)up(row_index, n * S * sizeof(int32_t));
        d_rows = (const uint32_t *)up(rows, (size_t)n_rows * kRow * sizeof(uint32_t));
        d_prob = (const double *)up(probability, S * sizeof(double));
        if (!d_parent || !d_arm || !d_row_index || !d_rows || !d_prob) {
            release();
            return fail(PMCTS_ERR_CUDA, "staging the master table on the device failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    long long *d_visits = (long long *)temp((size_t)n_rows * sizeof(long long));
    double *d_best = (double *)temp((size_t)n_rows * sizeof(double));
    int32_t *d_child = (int32_t *)temp((size_t)n * kA * sizeof(int32_t));
    double *d_value = (double *)temp((size_t)n * C * sizeof(double));
    double *d_explore = (double *)temp((size_t)n * C * sizeof(double));
    double *d_guess = (double *)temp((size_t)n * S * sizeof(double));
    int32_t *d_bc = (int32_t *)temp((size_t)n * C * sizeof(int32_t));
    int8_t *d_policy = (int8_t *)temp((size_t)C * max_depth);
    int32_t *d_pnodes = (int32_t *)temp((size_t)C * (max_depth + 1) * sizeof(int32_t));
    if (!d_visits || !d_best || !d_child || !d_value || !d_explore || !d_guess || !d_bc || !d_policy || !d_pnodes) {
        release();
        return fail(PMCTS_ERR_CUDA, "device scratch for the master reductions: %s", cudaGetErrorString(cudaGetLastError()));
    }
    M_TRY(cudaMemsetAsync(d_child, 0xFF, (size_t)n * kA * sizeof(int32_t), st));
    if (n_rows > 0) row_stats_kernel<<<grid(n_rows), kThreads, 0, st>>>(d_rows, n_rows, d_visits, d_best);
    children_kernel<<<grid(n), kThreads, 0, st>>>(d_parent, d_arm, n, d_child);
    scenario_utility_kernel<<<grid(n * S), kThreads, 0, st>>>(d_parent, d_row_index, d_visits, d_best, n, S, uct_coeff, d_value,
                                                             d_explore);
    best_child_kernel<<<grid(n * S), kThreads, 0, st>>>(d_child, d_value, n, S, 0, S, d_bc);
    guess_kernel<<<grid(n * S), kThreads, 0, st>>>(d_parent, d_row_index, d_visits, d_value, d_explore, d_bc, n, S, d_guess);
    global_utility_kernel<<<grid(n), kThreads, 0, st>>>(d_parent, d_row_index, d_rows, d_visits, d_guess, d_prob, n, S,
                                                       uct_coeff, d_value, d_explore);
    best_child_kernel<<<grid(n), kThreads, 0, st>>>(d_child, d_value, n, S, S, 1, d_bc);
    policy_kernel<<<grid(C), kThreads, 0, st>>>(d_bc, d_arm, S, max_depth, d_policy, d_pnodes);
    M_TRY(cudaGetLastError());
    pmcts::count_launches(ctx, n_rows > 0 ? 8 : 7);
    const cudaMemcpyKind kind = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (value) M_TRY(cudaMemcpyAsync(value, d_value, (size_t)n * C * sizeof(double), kind, st));
    if (explore) M_TRY(cudaMemcpyAsync(explore, d_explore, (size_t)n * C * sizeof(double), kind, st));
    if (guess) M_TRY(cudaMemcpyAsync(guess, d_guess, (size_t)n * S * sizeof(double), kind, st));
    if (best_child) M_TRY(cudaMemcpyAsync(best_child, d_bc, (size_t)n * C * sizeof(int32_t), kind, st));
    if (policy) M_TRY(cudaMemcpyAsync(policy, d_policy, (size_t)C * max_depth, kind, st));
    if (policy_nodes) M_TRY(cudaMemcpyAsync(policy_nodes, d_pnodes, (size_t)C * (max_depth + 1) * sizeof(int32_t), kind, st));
    release();
    if (host) CUDA_TRY(cudaStreamSynchronize(st));
#undef M_TRY
    return PMCTS_OK;
}

// ---- the same reductions straight from the host runtime's master (no round trip through the caller's arrays) ------
namespace {

struct MasterStage {  // page-locked staging of the exported master, kept with the ctx (grow-only)
    char *p = nullptr;
    size_t cap = 0;
};
void free_master_stage(void *q) {
    MasterStage *m = (MasterStage *)q;
    if (m->p) cudaFreeHost(m->p);
    delete m;
}
size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" int pmcts_forest_master_policy(pmcts_ctx *ctx, pmcts_forest *f, const double *probability, double uct_coeff,
                                          int max_depth, double *value, double *explore, double *guess,
                                          int32_t *best_child, int8_t *policy, int32_t *policy_nodes) {
    if (!ctx || !f || !probability) return fail(PMCTS_ERR_ARG, "NULL argument");
    const int64_t n = pmcts_forest_master_size(f), n_rows = pmcts_forest_master_rows(f);
    const int S = pmcts::forest_info(f).n_scen;
    if (n < 1) return fail(PMCTS_ERR_ARG, "empty master");
    int prev = -1;
    cudaGetDevice(&prev);
    const int dev = pmcts_get_device(ctx);
    if (prev != dev) CUDA_TRY(cudaSetDevice(dev));
    void **slot = pmcts::ctx_attachment(ctx, free_master_stage, 1);
    if (!*slot) *slot = new MasterStage();
    MasterStage *st = (MasterStage *)*slot;
    const size_t o_arm = al256((size_t)n * sizeof(int32_t)), o_idx = o_arm + al256((size_t)n);
    const size_t o_rows = o_idx + al256((size_t)n * S * sizeof(int32_t));
    const size_t total = o_rows + al256((size_t)n_rows * kRow * sizeof(uint32_t));
    int rc = PMCTS_OK;
    if (total > st->cap) {
        if (st->p) cudaFreeHost(st->p);
        st->p = nullptr;
        st->cap = 0;
        if (cudaMallocHost((void **)&st->p, total + total / 2) != cudaSuccess)
            rc = fail(PMCTS_ERR_CUDA, "page-locked staging for the master: %s", cudaGetErrorString(cudaGetLastError()));
        else
            st->cap = total + total / 2;
    }
    if (!rc) {
        char *h = st->p;
        rc = pmcts_forest_export_master_rows(f, (int32_t *)h, (int8_t *)(h + o_arm), (int32_t *)(h + o_idx),
                                             (uint32_t *)(h + o_rows));
        if (rc) fail(rc, "%s", pmcts_forest_last_error());
        else
            rc = pmcts_master_policy(ctx, n, S, (const int32_t *)h, (const int8_t *)(h + o_arm), (const int32_t *)(h + o_idx),
                                     n_rows, (const uint32_t *)(h + o_rows), probability, uct_coeff, max_depth, value, explore,
                                     guess, best_child, policy, policy_nodes, PMCTS_MEM_HOST);
    }
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    return rc;
}
