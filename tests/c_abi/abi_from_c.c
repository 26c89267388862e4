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
