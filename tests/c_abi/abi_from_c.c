/* The boundary from plain C (C99, no Python, no CUDA headers): what a cgo / JNI / N-API binding of
 * include/pmcts_b200.h sees.  Built and run by tests/test_abi.py on the CPU:
 *   - the header compiles as C and the library links;
 *   - without a CUDA device pmcts_create reports PMCTS_ERR_NO_DEVICE (there is no CPU path) and says why;
 *   - the host runtime of the search (pmcts_forest_*: selection, back-up, master update) is driven from C: two waves
 *     of a two-scenario forest, every histogram filed once.
 * On a machine with a GPU the first check is skipped (argv[1] = "gpu"). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pmcts_b200.h"

#define CHECK(cond, msg)                                              \
    do {                                                              \
        if (!(cond)) {                                                \
            fprintf(stderr, "FAILED: %s (%s:%d)\n", msg, __FILE__, __LINE__); \
            return 1;                                                 \
        }                                                             \
    } while (0)

/* The multi-GPU snippets of INTEGRATION.md (route B), kept compiling: never called here (they need GPUs and peers). */
int integration_md_route_b(pmcts_ctx *ctx, pmcts_forest *forest, const int32_t *scen_slots, const double root[3],
                           const double dest[2], double tmin, int world, int rank, const int8_t *slots_w,
                           const int8_t *slots_m, int n_local, int max_depth, uint8_t (*all)[PMCTS_PEER_HANDLE_BYTES],
                           const uint32_t *d_leaf_bins, int64_t count, void *d_send, void *d_recv, size_t bytes) {
    uint8_t id[128];
    if (rank == 0 && pmcts_comm_unique_id(id)) return 1;
    pmcts_comm *comm;
    if (pmcts_comm_create(ctx, world, rank, id, &comm)) return 1;
    pmcts_search_params p;
    memset(&p, 0, sizeof p);
    p.n_leaves = 64; p.replicas = 32; p.budget = 320; p.seed = 5; p.pipeline = 1;
    p.world = world; p.rank = rank; p.comm = comm;
    p.slots_w = slots_w; p.slots_m = slots_m;
    pmcts_search_report rep;
    if (pmcts_forest_search(ctx, forest, scen_slots, root, dest, tmin, &p, &rep)) return 1;
    if (pmcts_allgather_updates(ctx, comm, d_send, bytes, d_recv)) return 1;
    uint8_t mine[PMCTS_PEER_HANDLE_BYTES];
    pmcts_peer *peer;
    if (pmcts_peer_create(ctx, world, rank, pmcts_search_block_bytes(n_local, p.n_leaves, max_depth, p.replicas), &peer, mine))
        return 1;
    if (pmcts_peer_connect(peer, &all[0][0])) return 1;
    p.comm = NULL;
    p.peer = peer;
    const void *gathered;
    if (pmcts_peer_put_bins(ctx, peer, d_leaf_bins, count, 1) || pmcts_peer_wait(ctx, peer, &gathered)) return 1;
    int arrived = 0;
    if (pmcts_peer_poll(ctx, peer, &arrived)) return 1;
    pmcts_peer_destroy(peer);
    pmcts_comm_destroy(comm);
    return 0;
}

int main(int argc, char **argv) {
    enum { S = 2, B = 8, K = 4, D = 12, MAX_NODES = 64 };
    printf("%s\n", pmcts_version());
    if (!(argc > 1 && strcmp(argv[1], "gpu") == 0)) {
        pmcts_ctx *ctx = NULL;
        const int rc = pmcts_create(0, &ctx);
        CHECK(rc == PMCTS_ERR_NO_DEVICE && ctx == NULL, "pmcts_create must fail without a device");
        CHECK(strstr(pmcts_last_error(), "no CPU path") != NULL, "the error must say that there is no CPU path");
    }
    CHECK(pmcts_search_block_bytes(S, B, D, K) == (size_t)(S * B * 4 + S * B * D + S * B * 12), "update block layout");

    static uint8_t perms[S][MAX_NODES][8];
    uint64_t seeds[S] = {11u, 12u};
    CHECK(pmcts_python_shuffles_batch(seeds, S, MAX_NODES, 8, &perms[0][0][0]) == PMCTS_OK, "shuffles");
    int32_t ids[S] = {0, 1};
    double prob[S] = {0.5, 0.5};
    pmcts_forest *f = NULL;
    CHECK(pmcts_forest_create(S, S, ids, D, prob, 0.1, 0.3, &perms[0][0][0], MAX_NODES, &f) == PMCTS_OK, "forest");

    int32_t leaf[S * B], plen[S * B];
    int8_t path[S * B * D], slots[S * B];
    uint32_t bins[S * B * 12];
    memset(slots, 0, sizeof slots);
    for (int wave = 0; wave < 2; ++wave) {
        CHECK(pmcts_forest_select(f, B, D, leaf, plen, path) == PMCTS_OK, pmcts_forest_last_error());
        for (int i = 0; i < S * B; ++i) {
            CHECK(plen[i] >= 1 && plen[i] <= D && path[i * D] >= 0 && path[i * D] < 8, "a leaf's path");
            memset(&bins[i * 12], 0, 12 * sizeof(uint32_t));
            bins[i * 12 + 2 + path[i * D] % 8] = K; /* (the stand-in for K rollouts of the leaf) */
        }
        CHECK(pmcts_forest_backup(f, B, bins, slots) == PMCTS_OK, pmcts_forest_last_error());
        CHECK(pmcts_forest_master_apply(f, S, ids, B, D, plen, path, slots, bins) == PMCTS_OK, pmcts_forest_last_error());
    }
    CHECK(pmcts_forest_tree_size(f, 0) == 1 + 2 * B && pmcts_forest_tree_size(f, 1) == 1 + 2 * B, "one node per leaf");
    const int64_t n = pmcts_forest_master_size(f);
    CHECK(n > 8 && n <= 1 + 2 * S * B, "master size");
    uint32_t *table = (uint32_t *)malloc((size_t)n * S * 96 * sizeof(uint32_t));
    CHECK(table != NULL && pmcts_forest_export_master(f, NULL, NULL, table) == PMCTS_OK, "export");
    for (int s = 0; s < S; ++s) { /* the root's row of a scenario holds every rollout of its tree */
        uint64_t total = 0;
        for (int j = 0; j < 96; ++j) total += table[(size_t)s * 96 + j];
        CHECK(total == (uint64_t)(2 * B * K), "root visits");
    }
    free(table);
    CHECK(pmcts_forest_backup(f, B, bins, slots) != PMCTS_OK, "a back-up without a selection must be refused");
    pmcts_forest_destroy(f);
    printf("abi_from_c ok\n");
    return 0;
}
