// pmcts_b200 -- what the translation units of the library share beyond the public C ABI.
#pragma once
#include <functional>
#include <mutex>

#include "../../include/pmcts_b200.h"

namespace pmcts {

// message returned by pmcts_last_error() on the calling thread (pmcts_b200.cu)
void set_last_error(const char *msg);

// Objects other translation units keep with a ctx for the ctx's lifetime (staging buffers of the search driver: key 0,
// of the master reductions: key 1): returns the slot; pmcts_destroy calls free_fn on what it holds.
void **ctx_attachment(pmcts_ctx *ctx, void (*free_fn)(void *), int key = 0);

// kernels launched on behalf of the ctx by another translation unit (pmcts_launch_count)
void count_launches(pmcts_ctx *ctx, int n);

// is per-launch timing on (pmcts_enable_timing)?  Its event pairs bracket single launches, so nothing is replayed
// from a recorded graph while it is.
bool ctx_timing(const pmcts_ctx *ctx);

// the lock every entry point of a ctx takes (recursive): held by a translation unit across a SEQUENCE of calls that
// must not interleave with another host thread's -- a stream capture, whose graph would otherwise record them too
std::unique_lock<std::recursive_mutex> ctx_lock(pmcts_ctx *ctx);

// size of the blocks a peer exchange was created for (pmcts_peer.cu)
size_t peer_block_bytes(const pmcts_peer *peer);

// shape of a host forest (pmcts_forest_host.cpp)
struct ForestInfo {
    int n_scen, n_local, max_depth;
};
ForestInfo forest_info(const pmcts_forest *f);
int forest_local_id(const pmcts_forest *f, int local_index);
// A search that fails between a selection and its back-up leaves waves selected and not filed: their leaves' virtual
// visits are taken back and the waves dropped, so that the forest can be searched again (the nodes they expanded stay:
// unvisited leaves, which the next selection treats like any other child without statistics of its own).
void forest_abandon_pending(pmcts_forest *f);

// Tree.uct_search for one worker tree (pmcts_forest_host.cpp); `rollout(scenario, path, len, z, &bin, &steps)` runs the
// default policy of a leaf with the normals z[0 .. n_instants) and returns 0 or a negative PMCTS_ERR_* / the positive
// status byte of a rollout the reference would have raised on.
using SeqRollout = std::function<int(int, const int8_t *, int, const double *, int *, int *)>;
int forest_sequential(pmcts_forest *f, int local_index, int64_t budget, int frequency, int n_instants,
                      pmcts_mt_state *py_state, pmcts_mt_state *np_state, const SeqRollout &rollout,
                      pmcts_search_report *rep);

}  // namespace pmcts
