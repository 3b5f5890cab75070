/* Plain-C client of the branching-table entry points (no Python, no torch): what a `ccall` host does for
 * branching_table(p, TensorNetworkSolver(), region) (/root/reference/src/branch.jl:79).
 * Region: the path u - x - v - y with u (vertex 0) and v (vertex 2) open.  Known table (bit 0 = u, bit 1 = v):
 *   00 -> size 2, {x, y};  01 -> size 2, {u, y};  10 -> size 1, {v};  11 -> size 2, {u, v}
 * mis_compactify keeps only 00 (every other row chooses a superset of its boundary vertices and is no larger).  Also runs
 * the two regions {this one, a triangle with one open vertex} through the batched call.  Prints "ok" and exits 0 when
 * everything matches (the expected values are the oracle's, tests/test_table_configs.py checks the same calls at size). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tbcuda.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "c_table: check failed at line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(void) {
    /* vertex tensors 0..3, edge tensors (0,1) (1,2) (2,3) */
    static const int32_t off1[] = {0, 1, 2, 3, 4, 6, 8, 10};
    static const int32_t lab1[] = {0, 1, 2, 3, 0, 1, 1, 2, 2, 3};
    static const int32_t bnd1[] = {0, 2};
    /* triangle 0 1 2, vertex 1 open */
    static const int32_t off2[] = {0, 1, 2, 3, 5, 7, 9};
    static const int32_t lab2[] = {0, 1, 2, 0, 1, 1, 2, 0, 2};
    tb_network nets[2];
    memset(nets, 0, sizeof nets);
    nets[0].n_labels = 4; nets[0].n_leaves = 7; nets[0].leaf_off = off1; nets[0].leaf_labels = lab1;
    nets[1].n_labels = 3; nets[1].n_leaves = 6; nets[1].leaf_off = off2; nets[1].leaf_labels = lab2;
    tb_options opts;
    memset(&opts, 0, sizeof opts);
    tb_ctx* ctx = NULL;
    int rc = tb_init(&opts, &ctx);
    if (rc) { fprintf(stderr, "tb_init: %d %s\n", rc, tb_last_error(NULL)); return 1; }

    uint8_t keep[8];
    double sizes[8];
    int64_t row_off[9], total = -1;
    uint32_t cfg[64];
    rc = tb_branching_table(ctx, &nets[0], bnd1, 2, keep, sizes, row_off, cfg, 64, &total);
    if (rc) { fprintf(stderr, "tb_branching_table: %d %s\n", rc, tb_last_error(ctx)); return 1; }
    CHECK(sizes[0] == 2 && sizes[1] == 2 && sizes[2] == 1 && sizes[3] == 2);
    CHECK(keep[0] == 1 && keep[1] == 0 && keep[2] == 0 && keep[3] == 0);
    CHECK(total == 1 && row_off[0] == 0 && row_off[1] == 1 && row_off[2] == 1 && row_off[3] == 1 && row_off[4] == 1);
    CHECK(cfg[0] == 0xa);

    /* every row (no reduction) through tb_table_configs */
    rc = tb_table_configs(ctx, &nets[0], bnd1, 2, NULL, sizes, row_off, cfg, 64, &total);
    CHECK(rc == TB_OK && total == 4 && row_off[1] == 1 && row_off[2] == 2 && row_off[3] == 3 && row_off[4] == 4);
    CHECK(cfg[0] == 0xa && cfg[1] == 0x9 && cfg[2] == 0x4 && cfg[3] == 0x5);

    /* both regions in one call: rows 0..3 = the path, rows 4..5 = the triangle (vertex 1 out: {0} or {2}; in: {1}) */
    static const int32_t boff[] = {0, 2, 3};
    static const int32_t blab[] = {0, 2, 1};
    rc = tb_branching_tables(ctx, nets, boff, blab, 2, keep, sizes, row_off, cfg, 64, &total);
    if (rc) { fprintf(stderr, "tb_branching_tables: %d %s\n", rc, tb_last_error(ctx)); return 1; }
    CHECK(sizes[4] == 1 && sizes[5] == 1 && keep[4] == 1 && keep[5] == 0);
    CHECK(total == 3 && row_off[1] == 1 && row_off[4] == 1 && row_off[5] == 3 && row_off[6] == 3);
    CHECK(cfg[0] == 0xa && cfg[1] == 0x1 && cfg[2] == 0x4);

    /* a short buffer is an error that still reports the total */
    total = -1;
    rc = tb_branching_tables(ctx, nets, boff, blab, 2, keep, sizes, row_off, cfg, 2, &total);
    CHECK(rc == TB_ERR_BAD_ARGUMENT && total == 3);
    tb_shutdown(ctx);
    printf("ok\n");
    return 0;
}
