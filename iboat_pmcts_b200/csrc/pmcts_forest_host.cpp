// pmcts_b200 host runtime of the leaf-batched forest search: array-based worker trees and the replicated
// master tree, with the selection / expansion / back-up logic of the reference's Python classes
// (code/solver/worker.py:27-253, 302-378; code/solver/master_node.py:10-63) restated over
// structure-of-arrays storage.  Pure host code (no CUDA): it decides *which* rollouts the GPU runs and
// files their histograms; it never computes a rollout.
//
// Semantics are those of iboat_pmcts_b200/worker.py's batched path (Tree.select_leaves /
// back_up_leaves, forest._master_add), which tests/test_host_forest.py compares it with leaf by leaf and
// counter by counter:
//   * a node's untried actions are a permutation handed in by the caller (the reference shuffles them with
//     Python's generator, worker.py:51) and are consumed from the end (list.pop());
//   * children are kept in expansion order and the best child is the FIRST maximum of
//       own            = max_a mean_a + C sqrt(2 ln(N_parent + pending) / (N_child + pending))
//       score          = own                      if the master has no value for the child
//                      = (1 - rho) own + rho m    otherwise            (worker.py:224-253)
//     with m = sum over scenarios that visited the node and its parent of probability[s] * (max_a mean_a +
//     C sqrt(2 ln N_parent / N_node)) (worker.py:302-346), evaluated once per wave;
//   * mean_a is the binned mean of utils.Hist (utils.py:49-59): sum h[b] (b + 0.5) / 8 / sum h[b] -- exact in
//     fp64 whatever the summation order;
//   * a leaf's wave histogram goes to a random action slot of the leaf (worker.py:62-63,
//     master_node.py:42-44) and to the slot of the action taken at every ancestor.
#include <sched.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <thread>
#include <algorithm>
#include <vector>

#include "../../include/pmcts_b200.h"
#include "pmcts_internal.h"

namespace {

constexpr int kA = 8;    // actions
constexpr int kB = 12;   // histogram bins
constexpr int kRow = kA * kB;

thread_local std::string g_err;

int fail(int code, const char *msg) {
    g_err = msg;
    return code;
}

// The reference runs one process per worker tree (forest.py:69-75).  Here the trees of a wave are walked by a few
// host threads: selection and back-up touch only their own tree (the master is read-only while leaves are
// selected, apart from its per-wave memo, which is written with release / acquire ordering and holds the same
// value whoever computes it), and the master update of a wave is split by scenario slice.  The threads are a
// process-wide pool that sleeps between waves: a wave of a small search is a few hundred tree walks, less than the
// cost of starting threads for it.  Tiny waves stay on the calling thread.
constexpr int64_t kMinParallelWork = 128;

class Pool {
  public:
    static Pool &get() {
        static Pool *pool = nullptr;
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        if (!pool || pool->pid_ != getpid()) pool = new Pool();  // (a forked child starts its own threads)
        return *pool;
    }
    int size() const { return (int)workers_.size() + 1; }

    // fn(i) for i in [0, n), on the pool's threads and the caller; returns when all are done.
    // A search hands the pool three small jobs per wave: the back-up and the master update back to back, the next
    // selection right behind them, then nothing while the GPU rolls the wave out.  Two things keep such jobs cheap.
    // (1) A job is finished when its TASKS are, not when every thread has reported: indices are claimed from one
    // counter that also carries the job's number, the caller claims like everybody else, and a thread that wakes
    // late finds nothing to claim -- a sleeping or descheduled worker never holds a job up.  (2) A worker that has
    // run out of tasks polls for the next job for 50 us before it sleeps, which covers the back-to-back jobs of a
    // wave without keeping the cores busy through the GPU's part (longer polling was measured: faster on a quiet box,
    // several times slower on a box whose cores were contended).
    void run(int n, const std::function<void(int)> &fn) {
        std::lock_guard<std::mutex> serial(run_mu_);  // one job at a time
        const uint64_t gen = ((ticket_.load() >> 32) + 1) & 0xffffffffu;
        Job &j = jobs_[gen & 1];
        j.fn.store(&fn);
        j.n.store(n);
        done_.store(0);
        ticket_.store(gen << 32);  // publishes the job
        if (sleepers_.load() > 0) {
            // (the lock orders this wake-up after a sleeper's look at the counter: it either saw the new job or is
            //  waiting by now)
            { std::lock_guard<std::mutex> lk(mu_); }
            cv_.notify_all();
        }
        claim(gen);
        for (int spins = 0; done_.load() != n; ++spins) {  // tasks other threads are still running
            relax();
            if ((spins & 255) == 255) std::this_thread::yield();
        }
    }

  private:
    int spin_us_ = 50;  // (PMCTS_POOL_SPIN_US overrides; 0: sleep at once)
    struct Job {
        std::atomic<const std::function<void(int)> *> fn{nullptr};
        std::atomic<int> n{0};
    };

    // runs unclaimed tasks of job `gen` until there are none, or the pool has moved on to another job.  (Two job
    // slots alternate: a thread that still holds the number of a finished job may read the slot of the job after
    // next, but it cannot claim from it -- the counter no longer carries its number.)
    void claim(uint64_t gen) {
        Job &j = jobs_[gen & 1];
        for (;;) {
            uint64_t t = ticket_.load();
            if ((t >> 32) != gen) return;
            if ((int64_t)(t & 0xffffffffu) >= (int64_t)j.n.load()) return;
            if (!ticket_.compare_exchange_weak(t, t + 1)) continue;
            (*j.fn.load())((int)(t & 0xffffffffu));
            done_.fetch_add(1);
        }
    }

    Pool() : pid_(getpid()) {
        // up to 16 threads, of this process's share of the cores when a launcher runs several ranks on the box
        // (LOCAL_WORLD_SIZE as torchrun sets it): polling threads must not outnumber the cores.  PMCTS_HOST_THREADS
        // overrides.
        unsigned hw = std::thread::hardware_concurrency();
        {   // (the cores this process may run on: a cpuset / taskset narrower than the machine counts)
            cpu_set_t allowed;
            CPU_ZERO(&allowed);
            if (sched_getaffinity(0, sizeof allowed, &allowed) == 0 && CPU_COUNT(&allowed) > 0)
                hw = hw ? std::min<unsigned>(hw, (unsigned)CPU_COUNT(&allowed)) : (unsigned)CPU_COUNT(&allowed);
        }
        if (const char *lw = std::getenv("LOCAL_WORLD_SIZE")) {
            const int ranks = std::atoi(lw);
            if (ranks > 1 && hw) hw = std::max(1u, hw / (unsigned)ranks);
        }
        int nt = (int)std::min<unsigned>(hw ? hw : 1u, 16u);
        if (const char *ht = std::getenv("PMCTS_HOST_THREADS")) {
            const int want = std::atoi(ht);
            if (want >= 1) nt = std::min(want, 64);
        }
        if (const char *sp = std::getenv("PMCTS_POOL_SPIN_US")) spin_us_ = std::max(0, std::min(std::atoi(sp), 10000));
        for (int t = 1; t < nt; ++t) {
            workers_.emplace_back([this] { loop(); });
            workers_.back().detach();
        }
    }
    static void relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    void loop() {
        using clock = std::chrono::steady_clock;
        uint64_t seen = 0;
        for (;;) {
            // poll for the next job, then sleep
            const clock::time_point until = clock::now() + std::chrono::microseconds(spin_us_);
            bool have = false;
            for (int spins = 0;; ++spins) {
                if ((ticket_.load() >> 32) != seen) {
                    have = true;
                    break;
                }
                relax();
                if ((spins & 63) == 63 && clock::now() >= until) break;
            }
            if (!have) {
                std::unique_lock<std::mutex> lk(mu_);
                sleepers_.fetch_add(1);
                cv_.wait(lk, [&] { return (ticket_.load() >> 32) != seen; });
                sleepers_.fetch_sub(1);
            }
            seen = ticket_.load() >> 32;
            claim(seen);
        }
    }
    pid_t pid_;
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_;
    Job jobs_[2];
    std::atomic<uint64_t> ticket_{0};  // job number << 32 | next unclaimed index
    std::atomic<int> done_{0}, sleepers_{0};
};

template <typename F>
void parallel_for(int n, int64_t work, F &&fn) {
    if (n <= 1 || work < kMinParallelWork || Pool::get().size() <= 1) {
        for (int i = 0; i < n; ++i) fn(i);
        return;
    }
    std::atomic<bool> failed{false};
    const std::function<void(int)> body = [&](int i) {
        try {
            fn(i);
        } catch (...) {  // (allocation failure while a tree grows): reported by the caller, never thrown across a thread
            failed.store(true);
        }
    };
    Pool::get().run(n, body);
    if (failed.load()) throw std::bad_alloc();
}

// (visit count, max over actions of the binned mean) of one node's 8 x 12 counters
template <typename T>
inline void node_value(const T *row, int64_t &visits, double &best) {
    visits = 0;
    best = 0.0;
    for (int a = 0; a < kA; ++a) {
        int64_t tot = 0, acc = 0;  // acc = sum h[b] * (2b + 1): the mean is acc / 16 / tot
        for (int b = 0; b < kB; ++b) {
            tot += (int64_t)row[a * kB + b];
            acc += (int64_t)row[a * kB + b] * (2 * b + 1);
        }
        visits += tot;
        if (tot > 0) {
            const double mean = ((double)acc * 0.0625) / (double)tot;
            if (mean > best) best = mean;
        }
    }
}

// ---- the reference's random streams ---------------------------------------------------------------------------
// The sequential search (Tree.uct_search, worker.py:147-186) interleaves three consumers of two generators:
// random.shuffle at every node creation (worker.py:51) and random.gauss at every transition (simulatorTLKT.py:515)
// draw from CPython's global Mersenne Twister; np.random.randint picks the action slots of Node.back_up and
// MasterNode.add_reward (worker.py:62, master_node.py:42) from NumPy's legacy one.  A seeded script must see the same
// tree here, so both are continued from the caller's state (random.getstate() / np.random.get_state()) and handed
// back where the reference would have left them.
struct Mt19937 {
    uint32_t mt[624];
    int pos = 624;
    uint32_t next32() {
        if (pos >= 624) {
            for (int kk = 0; kk < 624; ++kk) {
                const uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
                mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            pos = 0;
        }
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

// CPython's random.Random (Lib/random.py, Modules/_randommodule.c)
struct PyRandom {
    Mt19937 g;
    bool has_gauss = false;
    double gauss_next = 0.0;
    double random() {  // genrand_res53
        const uint32_t a = g.next32() >> 5, b = g.next32() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    double gauss() {  // gauss(0.0, 1.0): 0.0 + z * 1.0 == z
        if (has_gauss) {
            has_gauss = false;
            return gauss_next;
        }
        const double x2pi = random() * 6.283185307179586476925286766559;  // TWOPI = 2.0 * _pi
        const double g2rad = std::sqrt(-2.0 * std::log(1.0 - random()));
        gauss_next = std::sin(x2pi) * g2rad;
        has_gauss = true;
        return std::cos(x2pi) * g2rad;
    }
    void shuffle(uint8_t *x, int n) {  // x[i], x[_randbelow(i + 1)] swapped for i = n - 1 .. 1
        for (int i = n - 1; i >= 1; --i) {
            const uint32_t m = (uint32_t)i + 1u;
            int bits = 0;
            for (uint32_t v = m; v; v >>= 1) ++bits;
            uint32_t r;
            do r = g.next32() >> (32 - bits); while (r >= m);
            const uint8_t tmp = x[i];
            x[i] = x[r];
            x[r] = tmp;
        }
    }
};

struct Tree {
    int scen = 0;  // global scenario id
    PyRandom *pyrng = nullptr;  // sequential search: a node's untried actions are shuffled when it is created
    std::vector<int32_t> parent, mnode, pending;
    std::vector<int8_t> arm;
    std::vector<int16_t> depth;
    std::vector<int32_t> child;       // [node][kA] in expansion order, -1 beyond n_child
    std::vector<uint8_t> n_child, n_untried;
    std::vector<uint8_t> untried;     // [node][kA] action indices, consumed from the end
    std::vector<int64_t> table;       // [node][kA][kB]
    std::vector<int64_t> visits;      // cached
    std::vector<double> best;         // cached
    const uint8_t *perms = nullptr;   // caller-owned [max_nodes][kA]
    int64_t max_nodes = 0;
    // leaves of the waves selected and not yet backed up, oldest first (the pipelined search keeps two in flight)
    std::deque<std::vector<int32_t>> waves;
    // A wave files hundreds of histograms through the same ancestors: the cached (visits, best) of a node are
    // recomputed once per wave, after the last add, from the nodes listed here
    std::vector<int64_t> touch_stamp;
    std::vector<int32_t> touched;
    int64_t backups = 0;

    void reserve(size_t n) {
        parent.reserve(n); mnode.reserve(n); pending.reserve(n); arm.reserve(n); depth.reserve(n);
        child.reserve(n * kA); n_child.reserve(n); n_untried.reserve(n); untried.reserve(n * kA);
        table.reserve(n * kRow); visits.reserve(n); best.reserve(n); touch_stamp.reserve(n);
    }

    int32_t size() const { return (int32_t)parent.size(); }

    int32_t add_node(int32_t par, int a) {
        const int32_t idx = size();
        parent.push_back(par);
        arm.push_back((int8_t)a);
        depth.push_back(par < 0 ? 0 : (int16_t)(depth[par] + 1));
        mnode.push_back(-1);
        pending.push_back(0);
        child.insert(child.end(), kA, -1);
        n_child.push_back(0);
        n_untried.push_back(kA);
        if (pyrng) {
            uint8_t order[kA];
            for (int i = 0; i < kA; ++i) order[i] = (uint8_t)i;
            pyrng->shuffle(order, kA);
            untried.insert(untried.end(), order, order + kA);
        } else {
            untried.insert(untried.end(), perms + (size_t)idx * kA, perms + (size_t)idx * kA + kA);
        }
        table.insert(table.end(), kRow, 0);
        visits.push_back(0);
        best.push_back(0.0);
        touch_stamp.push_back(-1);
        return idx;
    }
};

struct Master {
    int n_scen = 0;
    std::vector<int32_t> parent, child;  // child [node][kA] by action index
    std::vector<int8_t> arm;
    // Counters.  A node near the root is visited by every scenario, a deep one by the one tree that grew it: a dense
    // table[node][scenario][8][12] is nine tenths zeros at 10 scenarios (and was most of the cost of a search:
    // 384 MB to allocate, clear and page in for 10^5 nodes).  So a (node, scenario) pair gets its 8 x 12 row when
    // the scenario first files a histogram there: row[node][scenario] indexes the scenario's own pool (-1: none
    // yet).  One pool per scenario, because the update of a wave runs one thread per scenario slice.
    static constexpr int kChunk = 256;  // rows per allocation: a pool never moves its rows
    struct Slice {
        std::vector<std::vector<uint32_t>> chunks;
        std::vector<int64_t> visits, touch;  // cached visit total; the update that last added to the row
        std::vector<double> best;            // cached max over actions of the binned mean
        std::vector<int32_t> node;           // the row's node
        int32_t rows = 0;
        uint32_t *counters(int32_t r) { return chunks[r / kChunk].data() + (size_t)(r % kChunk) * kRow; }
        const uint32_t *counters(int32_t r) const { return chunks[r / kChunk].data() + (size_t)(r % kChunk) * kRow; }
        int32_t add_row(int32_t nd) {
            if (rows % kChunk == 0) chunks.emplace_back((size_t)kChunk * kRow, 0u);
            visits.push_back(0);
            touch.push_back(-1);
            best.push_back(0.0);
            node.push_back(nd);
            return rows++;
        }
    };
    std::vector<Slice> slices;           // [n_scen]
    std::vector<int32_t> row;            // [node][n_scen]
    // the scenarios that own a row of a node, as a list in ascending scenario order (first[node] -> ent_next ...):
    // the master's value of a node sums over them (worker.py:302-346), and with 80 scenarios on 8 GPUs a walk over all
    // n_scen slots of every node was most of a wave's selection
    std::vector<int32_t> first, ent_scen, ent_row, ent_next;
    std::vector<double> memo;            // master UCT per node, valid when stamp == wave
    std::vector<int64_t> stamp;
    int64_t applies = 0;

    int32_t size() const { return (int32_t)parent.size(); }
    int64_t visits_of(int32_t node, int s) const {
        const int32_t r = row[(size_t)node * n_scen + s];
        return r < 0 ? 0 : slices[s].visits[r];
    }

    int32_t add_node(int32_t par, int a) {
        const int32_t idx = size();
        parent.push_back(par);
        arm.push_back((int8_t)a);
        child.insert(child.end(), kA, -1);
        row.insert(row.end(), n_scen, -1);
        first.push_back(-1);
        memo.push_back(0.0);
        stamp.push_back(-1);
        if (par >= 0) child[(size_t)par * kA + a] = idx;
        return idx;
    }

    // node reached by a path of action indices, created on the way (ancestors first)
    int32_t descend(const int8_t *path, int len) {
        int32_t node = 0;
        for (int d = 0; d < len; ++d) {
            int32_t nxt = child[(size_t)node * kA + path[d]];
            if (nxt < 0) nxt = add_node(node, path[d]);
            node = nxt;
        }
        return node;
    }

    // links (node, s) into the node's scenario list; serial (lists of different scenarios share the node's head)
    void link(int32_t node, int s) {
        const int32_t e = (int32_t)ent_scen.size();
        ent_scen.push_back(s);
        ent_row.push_back(row[(size_t)node * n_scen + s]);
        ent_next.push_back(-1);
        int32_t prev = -1, cur = first[node];
        while (cur >= 0 && ent_scen[cur] < s) {
            prev = cur;
            cur = ent_next[cur];
        }
        ent_next[e] = cur;
        if (prev < 0) first[node] = e;
        else ent_next[prev] = e;
    }

    // adds a histogram; the cached (visits, best) of the row are refreshed by the caller once per update, from
    // `touched` (rows of scenario s; scenario slices are disjoint, so concurrent callers never share an entry); a
    // node the scenario had no row for yet goes to `created`, for the caller to link() once the slices are done
    void add(int32_t node, int s, int slot, const uint32_t *bins, std::vector<int32_t> &touched,
             std::vector<int32_t> &created) {
        Slice &sl = slices[s];
        int32_t &r = row[(size_t)node * n_scen + s];
        if (r < 0) {
            r = sl.add_row(node);
            created.push_back(node);
        }
        uint32_t *c = sl.counters(r);
        for (int b = 0; b < kB; ++b) c[slot * kB + b] += bins[b];
        if (sl.touch[r] != applies) {
            sl.touch[r] = applies;
            touched.push_back(r);
        }
    }
    void refresh(int s, const std::vector<int32_t> &touched) {
        Slice &sl = slices[s];
        for (int32_t r : touched) node_value(sl.counters(r), sl.visits[r], sl.best[r]);
    }
};

}  // namespace

struct pmcts_forest {
    int n_scen = 0, n_local = 0, max_depth = 0;
    double uct_coeff = 0.0, rho = 0.0;
    std::vector<double> probability;
    std::vector<Tree> trees;
    Master master;
    int64_t wave = 0;

    double master_uct(int32_t m) {
        if (m < 0) return 0.0;
        if (__atomic_load_n(&master.stamp[m], __ATOMIC_ACQUIRE) == wave) {
            double memo;
            __atomic_load(&master.memo[m], &memo, __ATOMIC_RELAXED);
            return memo;
        }
        const int32_t p = master.parent[m];
        double acc = 0.0;
        bool any = false;
        for (int32_t e = master.first[m]; e >= 0; e = master.ent_next[e]) {  // ascending scenario order
            const int s = master.ent_scen[e];
            const int32_t r = master.ent_row[e];
            const int64_t n_node = master.slices[s].visits[r];
            const int64_t n_par = p >= 0 ? master.visits_of(p, s) : 0;
            if (n_node != 0 && n_par != 0) {
                const double uct = master.slices[s].best[r] +
                                   uct_coeff * std::pow(2.0 * std::log((double)n_par) / (double)n_node, 0.5);
                acc += uct * probability[s];
                any = true;
            }
        }
        const double value = any ? acc : 0.0;
        // (trees that meet at a node nobody has valued this wave race here: they store the same value, as atomic
        //  accesses, and whoever reads the stamp afterwards finds it)
        __atomic_store(&master.memo[m], &value, __ATOMIC_RELAXED);
        __atomic_store_n(&master.stamp[m], wave, __ATOMIC_RELEASE);
        return value;
    }

    // master node of a worker node (by its action path), or -1 when the master does not know it yet
    int32_t master_of(Tree &t, int32_t node) {
        if (t.mnode[node] >= 0) return t.mnode[node];
        if (t.parent[node] < 0) return t.mnode[node] = 0;
        const int32_t pm = master_of(t, t.parent[node]);
        if (pm < 0) return -1;
        const int32_t m = master.child[(size_t)pm * kA + t.arm[node]];
        if (m >= 0) t.mnode[node] = m;
        return m;
    }

    int32_t best_child(Tree &t, int32_t node) {
        const int64_t num_parent = t.visits[node] + t.pending[node];
        const double log_parent = 2.0 * std::log((double)num_parent);
        double best_score = -1.0;
        int32_t best = -1;
        for (int j = 0; j < t.n_child[node]; ++j) {
            const int32_t c = t.child[(size_t)node * kA + j];
            const double own = t.best[c] + uct_coeff * std::pow(log_parent / (double)(t.visits[c] + t.pending[c]), 0.5);
            const double um = master_uct(master_of(t, c));
            const double score = um == 0.0 ? own : (1.0 - rho) * own + rho * um;
            if (score > best_score) {
                best_score = score;
                best = c;
            }
        }
        return best;
    }

    int32_t tree_policy(Tree &t) {
        int32_t node = 0;
        while (t.depth[node] != max_depth) {
            if (t.n_untried[node] > 0) {
                if (t.size() >= t.max_nodes) return -1;
                const int a = t.untried[(size_t)node * kA + --t.n_untried[node]];
                const int32_t nn = t.add_node(node, a);
                t.child[(size_t)node * kA + t.n_child[node]++] = nn;
                return nn;
            }
            node = best_child(t, node);
        }
        return node;
    }
};

namespace {

void load_state(Mt19937 &g, const pmcts_mt_state &st) {
    std::memcpy(g.mt, st.key, sizeof g.mt);
    g.pos = st.pos;
}
void store_state(const Mt19937 &g, pmcts_mt_state &st) {
    std::memcpy(st.key, g.mt, sizeof g.mt);
    st.pos = g.pos;
}

}  // namespace

namespace pmcts {

// Tree.uct_search (worker.py:147-186) for one worker tree of the forest, in the reference's order and on the
// reference's random streams; `rollout` runs Tree.default_policy (worker.py:255-281) for a leaf given the normals its
// transitions would draw, and reports the reward's histogram bin and the transitions made.
int forest_sequential(pmcts_forest *f, int local_index, int64_t budget, int frequency, int n_instants,
                      pmcts_mt_state *py_state, pmcts_mt_state *np_state, const SeqRollout &rollout,
                      pmcts_search_report *rep) {
    if (!f || local_index < 0 || local_index >= f->n_local || !py_state || !np_state)
        return fail(PMCTS_ERR_ARG, "bad arguments");
    if (budget < 1 || frequency < 1 || n_instants < 2) return fail(PMCTS_ERR_ARG, "bad sizes");
    Tree &t = f->trees[local_index];
    if (t.perms) return fail(PMCTS_ERR_ARG, "the forest was created for the batched search (perms given)");
    PyRandom py;
    Mt19937 npr;
    load_state(py.g, *py_state);
    py.has_gauss = py_state->has_gauss != 0;
    py.gauss_next = py_state->gauss_next;
    load_state(npr, *np_state);
    t.pyrng = &py;
    t.max_nodes = std::max<int64_t>(t.max_nodes, t.size() + budget + 2);
    if (t.size() == 0) t.add_node(-1, 0);  // Tree._make_root: the root's actions are shuffled first
    Master &m = f->master;
    const int s = t.scen;
    struct Update { int32_t leaf; int bin; };
    std::vector<Update> buffer;
    std::vector<double> z(n_instants);
    std::vector<int8_t> path(f->max_depth);
    std::vector<int32_t> touched, created;
    uint32_t one_hot[kB];
    int rc = PMCTS_OK;
    for (int64_t ite = 1; ite <= budget && !rc; ++ite) {
        const int32_t leaf = f->tree_policy(t);
        if (leaf < 0) { rc = fail(PMCTS_ERR_ARG, "tree capacity exhausted"); break; }
        const int len = t.depth[leaf];
        {
            int d = len;
            for (int32_t x = leaf; t.parent[x] >= 0; x = t.parent[x]) path[--d] = t.arm[x];
        }
        // the normals the rollout's transitions draw: taken ahead, then the generator is put back and advanced by
        // the draws actually made (one per transition, simulatorTLKT.py:515)
        const PyRandom saved = py;
        for (int j = 0; j < n_instants; ++j) z[j] = py.gauss();
        int bin = 0, steps = 0;
        rc = rollout(s, path.data(), len, z.data(), &bin, &steps);
        if (rc) break;
        py = saved;
        for (int j = 0; j < steps; ++j) py.gauss();
        if (rep) { rep->rollouts += 1; rep->transitions += steps; rep->leaves += 1; }
        // Node.back_up (worker.py:55-70)
        int slot = (int)(npr.next32() & 7u);
        for (int32_t x = leaf; x >= 0; x = t.parent[x]) {
            int64_t *row = t.table.data() + (size_t)x * kRow;
            row[slot * kB + bin] += 1;
            node_value(row, t.visits[x], t.best[x]);
            slot = t.arm[x];
        }
        buffer.push_back({leaf, bin});
        // Tree.integrate_buffer (worker.py:348-378) every `frequency` iterations
        if (ite % frequency == 0) {
            m.applies += 1;
            touched.clear();
            for (const Update &u : buffer) {
                const int ulen = t.depth[u.leaf];
                int d = ulen;
                for (int32_t x = u.leaf; t.parent[x] >= 0; x = t.parent[x]) path[--d] = t.arm[x];
                int32_t node = m.descend(path.data(), ulen);
                std::memset(one_hot, 0, sizeof one_hot);
                one_hot[u.bin] = 1;
                m.add(node, s, (int)(npr.next32() & 7u), one_hot, touched, created);  // MasterNode.add_reward: a random slot
                while (m.parent[node] >= 0) {
                    m.add(m.parent[node], s, m.arm[node], one_hot, touched, created);
                    node = m.parent[node];
                }
            }
            m.refresh(s, touched);
            for (int32_t node : created) m.link(node, s);
            created.clear();
            f->wave += 1;
            buffer.clear();
            if (rep) rep->waves += 1;
        }
    }
    t.pyrng = nullptr;
    store_state(py.g, *py_state);
    py_state->has_gauss = py.has_gauss ? 1 : 0;
    py_state->gauss_next = py.gauss_next;
    store_state(npr, *np_state);
    return rc;
}

ForestInfo forest_info(const pmcts_forest *f) { return ForestInfo{f->n_scen, f->n_local, f->max_depth}; }
int forest_local_id(const pmcts_forest *f, int local_index) { return f->trees[local_index].scen; }
void forest_abandon_pending(pmcts_forest *f) {
    for (Tree &t : f->trees) {
        for (const std::vector<int32_t> &wave_leaves : t.waves)
            for (int32_t leaf : wave_leaves)
                for (int32_t x = leaf; x >= 0; x = t.parent[x]) t.pending[x] -= 1;
        t.waves.clear();
    }
}
}  // namespace pmcts

extern "C" {

const char *pmcts_forest_last_error(void) { return g_err.c_str(); }

int pmcts_forest_create(int n_scen, int n_local, const int32_t *local_ids, int max_depth,
                        const double *probability, double uct_coeff, double rho, const uint8_t *perms,
                        int64_t max_nodes, pmcts_forest **out) {
    if (!out || !local_ids || !probability) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (n_scen < 1 || n_local < 1 || n_local > n_scen || max_depth < 1 || max_nodes < 1)
        return fail(PMCTS_ERR_ARG, "bad sizes");
    pmcts_forest *f = new pmcts_forest();
    f->n_scen = n_scen;
    f->n_local = n_local;
    f->max_depth = max_depth;
    f->uct_coeff = uct_coeff;
    f->rho = rho;
    f->probability.assign(probability, probability + n_scen);
    f->trees.resize(n_local);
    for (int i = 0; i < n_local; ++i) {
        Tree &t = f->trees[i];
        t.scen = local_ids[i];
        t.perms = perms ? perms + (size_t)i * max_nodes * kA : nullptr;
        t.max_nodes = max_nodes;
        t.reserve((size_t)max_nodes);
        if (perms) t.add_node(-1, 0);  // (without perms: pmcts_forest_search_sequential creates the root, shuffled by the caller's generator)
    }
    f->master.n_scen = n_scen;
    f->master.slices.resize(n_scen);
    f->master.add_node(-1, 0);
    *out = f;
    return PMCTS_OK;
}

void pmcts_forest_destroy(pmcts_forest *f) { delete f; }

int pmcts_forest_set_coefficients(pmcts_forest *f, double uct_coeff, double rho) {
    if (!f) return fail(PMCTS_ERR_ARG, "forest is NULL");
    f->uct_coeff = uct_coeff;
    f->rho = rho;
    return PMCTS_OK;
}

int pmcts_forest_select(pmcts_forest *f, int n_leaves, int prefix_stride, int32_t *leaf_node,
                        int32_t *prefix_len, int8_t *prefix_act) try {
    if (!f || !leaf_node || !prefix_len || !prefix_act) return fail(PMCTS_ERR_ARG, "NULL argument");
    if (n_leaves < 1 || prefix_stride < f->max_depth) return fail(PMCTS_ERR_ARG, "prefix_stride must hold max_depth actions");
    std::atomic<int> exhausted{0};
    parallel_for(f->n_local, (int64_t)f->n_local * n_leaves, [&](int i) {
        Tree &t = f->trees[i];
        t.waves.emplace_back();
        std::vector<int32_t> &wave_leaves = t.waves.back();
        wave_leaves.reserve(n_leaves);
        for (int l = 0; l < n_leaves; ++l) {
            const int32_t leaf = f->tree_policy(t);
            if (leaf < 0) {
                exhausted.store(1);
                return;
            }
            for (int32_t x = leaf; x >= 0; x = t.parent[x]) t.pending[x] += 1;
            wave_leaves.push_back(leaf);
            const size_t o = (size_t)i * n_leaves + l;
            leaf_node[o] = leaf;
            const int len = t.depth[leaf];
            prefix_len[o] = len;
            int8_t *p = prefix_act + o * prefix_stride;
            std::memset(p, -1, prefix_stride);
            int d = len;
            for (int32_t x = leaf; t.parent[x] >= 0; x = t.parent[x]) p[--d] = t.arm[x];
        }
    });
    if (exhausted.load()) return fail(PMCTS_ERR_ARG, "tree capacity (max_nodes) exhausted");
    return PMCTS_OK;
} catch (const std::bad_alloc &) {
    return fail(PMCTS_ERR_ARG, "out of host memory");
}

int pmcts_forest_backup(pmcts_forest *f, int n_leaves, const uint32_t *leaf_bins, const int8_t *slots) try {
    if (!f || !leaf_bins || !slots) return fail(PMCTS_ERR_ARG, "NULL argument");
    for (int i = 0; i < f->n_local; ++i)
        if (f->trees[i].waves.empty() || (int)f->trees[i].waves.front().size() != n_leaves)
            return fail(PMCTS_ERR_ARG, "back-up does not match the oldest selection not yet backed up");
    parallel_for(f->n_local, (int64_t)f->n_local * n_leaves, [&](int i) {
        Tree &t = f->trees[i];
        const std::vector<int32_t> &wave_leaves = t.waves.front();
        const int64_t stamp = ++t.backups;
        t.touched.clear();
        for (int l = 0; l < n_leaves; ++l) {
            const int32_t leaf = wave_leaves[l];
            const uint32_t *bins = leaf_bins + ((size_t)i * n_leaves + l) * kB;
            int slot = slots[(size_t)i * n_leaves + l];
            for (int32_t x = leaf; x >= 0; x = t.parent[x]) {
                int64_t *row = t.table.data() + (size_t)x * kRow;
                for (int b = 0; b < kB; ++b) row[slot * kB + b] += bins[b];
                if (t.touch_stamp[x] != stamp) {
                    t.touch_stamp[x] = stamp;
                    t.touched.push_back(x);
                }
                t.pending[x] -= 1;
                slot = t.arm[x];  // the ancestors take the histogram in the slot of the action that led here
            }
        }
        for (int32_t x : t.touched) node_value(t.table.data() + (size_t)x * kRow, t.visits[x], t.best[x]);
        t.waves.pop_front();
    });
    return PMCTS_OK;
} catch (const std::bad_alloc &) {
    return fail(PMCTS_ERR_ARG, "out of host memory");
}

int pmcts_forest_master_apply(pmcts_forest *f, int n_blocks_scen, const int32_t *scen_ids, int n_leaves,
                              int prefix_stride, const int32_t *prefix_len, const int8_t *prefix_act,
                              const int8_t *slots, const uint32_t *leaf_bins) try {
    if (!f || !scen_ids || !prefix_len || !prefix_act || !slots || !leaf_bins) return fail(PMCTS_ERR_ARG, "NULL argument");
    Master &m = f->master;
    std::vector<uint8_t> seen(f->n_scen, 0);
    bool distinct = true;
    for (int i = 0; i < n_blocks_scen; ++i) {
        const int s = scen_ids[i];
        if (s < 0 || s >= f->n_scen) return fail(PMCTS_ERR_ARG, "scenario id out of range");
        if (seen[s]) distinct = false;
        seen[s] = 1;
    }
    // 1. the nodes of every path, created in block / leaf order (the order the serial walk creates them in)
    std::vector<int32_t> leaf_master((size_t)n_blocks_scen * n_leaves);
    for (size_t o = 0; o < leaf_master.size(); ++o)
        leaf_master[o] = m.descend(prefix_act + o * prefix_stride, prefix_len[o]);
    // 2. the counters: block i only touches the slice of its own scenario, so distinct scenarios run side by side
    m.applies += 1;
    std::vector<std::vector<int32_t>> created(n_blocks_scen);
    auto apply_block = [&](int i) {
        const int s = scen_ids[i];
        std::vector<int32_t> touched;
        touched.reserve((size_t)n_leaves * 2);
        for (int l = 0; l < n_leaves; ++l) {
            const size_t o = (size_t)i * n_leaves + l;
            const uint32_t *bins = leaf_bins + o * kB;
            int32_t node = leaf_master[o];
            m.add(node, s, slots[o], bins, touched, created[i]);
            while (m.parent[node] >= 0) {
                m.add(m.parent[node], s, m.arm[node], bins, touched, created[i]);
                node = m.parent[node];
            }
        }
        m.refresh(s, touched);
        if (!distinct) m.applies += 1;  // (a scenario listed twice: its second block must refresh again)
    };
    if (distinct) parallel_for(n_blocks_scen, (int64_t)n_blocks_scen * n_leaves, apply_block);
    else
        for (int i = 0; i < n_blocks_scen; ++i) apply_block(i);
    for (int i = 0; i < n_blocks_scen; ++i)
        for (int32_t node : created[i]) m.link(node, scen_ids[i]);
    f->wave += 1;  // the master changed: its UCT values are recomputed on next use
    return PMCTS_OK;
} catch (const std::bad_alloc &) {
    return fail(PMCTS_ERR_ARG, "out of host memory");
}

int64_t pmcts_forest_tree_size(pmcts_forest *f, int local_index) {
    if (!f || local_index < 0 || local_index >= f->n_local) return -1;
    return f->trees[local_index].size();
}

int64_t pmcts_forest_master_size(pmcts_forest *f) { return f ? f->master.size() : -1; }

// `count` consecutive random.shuffle(list(range(n_items))) of CPython's random.Random(seed) -- what Node.__init__
// does with the tree's generator (worker.py:50-51) -- so that the native search pops the same untried actions as
// the Python classes without a hundred thousand interpreter-level shuffles.  CPython: Random(int) seeds MT19937
// with init_by_array over the 32-bit limbs of |seed| (Modules/_randommodule.c), shuffle swaps x[i] with
// x[_randbelow(i + 1)] for i = n - 1 .. 1, and _randbelow(n) draws getrandbits(n.bit_length()) until < n
// (Lib/random.py); getrandbits(k <= 32) is the top k bits of one 32-bit output.
int pmcts_python_shuffles(uint64_t seed, int64_t count, int n_items, uint8_t *out) {
    if (!out || count < 0 || n_items < 1 || n_items > 255) return fail(PMCTS_ERR_ARG, "bad arguments");
    uint32_t mt[624];
    int mti = 624;
    auto init_genrand = [&](uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    };
    {
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        const int klen = key[1] ? 2 : 1;
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = 624 > klen ? 624 : klen; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
            if (++j >= klen) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    auto next32 = [&]() -> uint32_t {
        if (mti >= 624) {
            for (int kk = 0; kk < 624; ++kk) {
                const uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
                mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            mti = 0;
        }
        uint32_t y = mt[mti++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    };
    for (int64_t c = 0; c < count; ++c) {
        uint8_t *x = out + (size_t)c * n_items;
        for (int i = 0; i < n_items; ++i) x[i] = (uint8_t)i;
        for (int i = n_items - 1; i >= 1; --i) {
            const uint32_t n = (uint32_t)i + 1u;
            int bits = 0;
            for (uint32_t v = n; v; v >>= 1) ++bits;
            uint32_t r;
            do r = next32() >> (32 - bits); while (r >= n);
            const uint8_t tmp = x[i];
            x[i] = x[r];
            x[r] = tmp;
        }
    }
    return PMCTS_OK;
}

// The same for several generators at once -- one per worker tree of a search -- on the pool's threads (a tree's table
// is ~3 000 draws of its own Mersenne Twister: 40 us each, which was a tenth of a 10-scenario search when the trees'
// tables were filled one after the other).  out[n_seeds][count][n_items].
int pmcts_python_shuffles_batch(const uint64_t *seeds, int n_seeds, int64_t count, int n_items, uint8_t *out) try {
    if (!seeds || !out || n_seeds < 0 || count < 0 || n_items < 1 || n_items > 255) return fail(PMCTS_ERR_ARG, "bad arguments");
    std::atomic<int> bad{0};
    parallel_for(n_seeds, (int64_t)n_seeds * count, [&](int i) {
        if (pmcts_python_shuffles(seeds[i], count, n_items, out + (size_t)i * count * n_items)) bad.store(1);
    });
    return bad.load() ? fail(PMCTS_ERR_ARG, "bad arguments") : PMCTS_OK;
} catch (const std::bad_alloc &) {
    return fail(PMCTS_ERR_ARG, "out of host memory");
}

int pmcts_forest_search_sequential_hook(pmcts_forest *f, int local_index, int64_t budget, int frequency,
                                        pmcts_mt_state *py_state, pmcts_mt_state *np_state,
                                        pmcts_seq_rollout_hook_fn hook, void *user, pmcts_search_report *report) try {
    if (!f || !hook) return fail(PMCTS_ERR_ARG, "NULL argument");
    pmcts_search_report rep{};
    const int n_instants = f->max_depth + 1;
    int64_t iteration = 0;
    const pmcts::SeqRollout rollout = [&](int scen, const int8_t *path, int len, const double *z, int *bin, int *steps) -> int {
        int32_t b = 0, t = 0;
        const int rc = hook(user, iteration++, scen, path, len, z, n_instants, &b, &t);
        if (rc) return fail(rc < 0 ? rc : PMCTS_ERR_ARG, "the rollout hook refused the iteration");
        if (b < 0 || b >= kB || t < 0 || t > n_instants) return fail(PMCTS_ERR_ARG, "the rollout hook returned a bin / transition count out of range");
        *bin = b;
        *steps = t;
        return 0;
    };
    const int rc = pmcts::forest_sequential(f, local_index, budget, frequency, n_instants, py_state, np_state, rollout, &rep);
    if (report) *report = rep;
    return rc;
} catch (const std::bad_alloc &) {
    return fail(PMCTS_ERR_ARG, "out of host memory");
}

int pmcts_python_gauss(pmcts_mt_state *state, int64_t count, double *out) {
    if (!state || !out || count < 0) return fail(PMCTS_ERR_ARG, "bad arguments");
    PyRandom py;
    load_state(py.g, *state);
    py.has_gauss = state->has_gauss != 0;
    py.gauss_next = state->gauss_next;
    for (int64_t i = 0; i < count; ++i) out[i] = py.gauss();
    store_state(py.g, *state);
    state->has_gauss = py.has_gauss ? 1 : 0;
    state->gauss_next = py.gauss_next;
    return PMCTS_OK;
}

int pmcts_numpy_randint8(pmcts_mt_state *state, int64_t count, uint8_t *out) {
    if (!state || !out || count < 0) return fail(PMCTS_ERR_ARG, "bad arguments");
    Mt19937 g;
    load_state(g, *state);
    for (int64_t i = 0; i < count; ++i) out[i] = (uint8_t)(g.next32() & 7u);
    store_state(g, *state);
    return PMCTS_OK;
}

int pmcts_forest_export_tree(pmcts_forest *f, int local_index, int32_t *parent, int8_t *arm, int64_t *table,
                             uint8_t *n_untried, uint8_t *untried) {
    if (!f || local_index < 0 || local_index >= f->n_local) return fail(PMCTS_ERR_ARG, "bad tree index");
    const Tree &t = f->trees[local_index];
    const size_t n = t.parent.size();
    if (parent) std::memcpy(parent, t.parent.data(), n * sizeof(int32_t));
    if (arm) std::memcpy(arm, t.arm.data(), n);
    if (table) std::memcpy(table, t.table.data(), n * kRow * sizeof(int64_t));
    if (n_untried) std::memcpy(n_untried, t.n_untried.data(), n);
    if (untried) std::memcpy(untried, t.untried.data(), n * kA);
    return PMCTS_OK;
}

int pmcts_forest_export_master(pmcts_forest *f, int32_t *parent, int8_t *arm, uint32_t *table) {
    if (!f) return fail(PMCTS_ERR_ARG, "forest is NULL");
    const Master &m = f->master;
    const size_t n = m.parent.size();
    if (parent) std::memcpy(parent, m.parent.data(), n * sizeof(int32_t));
    if (arm) std::memcpy(arm, m.arm.data(), n);
    if (table) {
        std::memset(table, 0, n * (size_t)m.n_scen * kRow * sizeof(uint32_t));
        for (int s = 0; s < m.n_scen; ++s) {
            const Master::Slice &sl = m.slices[s];
            for (int32_t r = 0; r < sl.rows; ++r)
                std::memcpy(table + ((size_t)sl.node[r] * m.n_scen + s) * kRow, sl.counters(r), kRow * sizeof(uint32_t));
        }
    }
    return PMCTS_OK;
}

int64_t pmcts_forest_master_rows(pmcts_forest *f) {
    if (!f) return -1;
    int64_t total = 0;
    for (const Master::Slice &sl : f->master.slices) total += sl.rows;
    return total;
}

int pmcts_forest_export_master_rows(pmcts_forest *f, int32_t *parent, int8_t *arm, int32_t *row_index, uint32_t *rows) {
    if (!f) return fail(PMCTS_ERR_ARG, "forest is NULL");
    const Master &m = f->master;
    const size_t n = m.parent.size();
    if (parent) std::memcpy(parent, m.parent.data(), n * sizeof(int32_t));
    if (arm) std::memcpy(arm, m.arm.data(), n);
    std::vector<int64_t> first(m.n_scen + 1, 0);  // rows are numbered scenario by scenario
    for (int s = 0; s < m.n_scen; ++s) first[s + 1] = first[s] + m.slices[s].rows;
    if (row_index)
        for (size_t i = 0; i < n; ++i)
            for (int s = 0; s < m.n_scen; ++s) {
                const int32_t r = m.row[i * m.n_scen + s];
                row_index[i * m.n_scen + s] = r < 0 ? -1 : (int32_t)(first[s] + r);
            }
    if (rows)
        for (int s = 0; s < m.n_scen; ++s) {
            const Master::Slice &sl = m.slices[s];
            for (int32_t c = 0; c * Master::kChunk < sl.rows; ++c) {
                const size_t cnt = std::min<size_t>(Master::kChunk, sl.rows - (size_t)c * Master::kChunk);
                std::memcpy(rows + ((size_t)first[s] + (size_t)c * Master::kChunk) * kRow, sl.chunks[c].data(),
                            cnt * kRow * sizeof(uint32_t));
            }
        }
    return PMCTS_OK;
}

}  // extern "C"
