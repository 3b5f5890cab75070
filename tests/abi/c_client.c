/* Plain-C client of libtbcuda.so: no Python, no torch, no CUDA headers -- exactly what a Julia `ccall` host sees.
 * Reads a branch list written by tests/test_abi_layout.py (a flat little-endian binary file), contracts it with ONE
 * tb_contract_networks call on the given devices (tb_init_multi: LPT sharding + one ncclAllReduce(max) inside the
 * library when there is more than one device) and prints one value per branch.
 *
 *   c_client <file> <dev0,dev1,...>
 * file: int32 n_branches, then per branch: int32 n_labels, n_leaves, n_lab_total, double r,
 *       int32 leaf_off[n_leaves+1], leaf_labels[n_lab_total], node_left[n_leaves-1], node_right[n_leaves-1]
 *       (n_leaves == 0: an empty graph, nothing follows).  Unit weights. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tbcuda.h"

static int32_t* read_i32(FILE* f, size_t n) {
    int32_t* p = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    if (n && fread(p, sizeof(int32_t), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s file dev0,dev1,...\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    int32_t devices[64], n_dev = 0;
    for (char* tok = strtok(argv[2], ","); tok && n_dev < 64; tok = strtok(NULL, ",")) devices[n_dev++] = atoi(tok);
    int32_t n;
    if (fread(&n, 4, 1, f) != 1) return 2;
    tb_network* nets = (tb_network*)calloc((size_t)(n ? n : 1), sizeof(tb_network));
    double* r = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
    for (int i = 0; i < n; ++i) {
        int32_t hdr[3];
        if (fread(hdr, 4, 3, f) != 3 || fread(&r[i], 8, 1, f) != 1) return 2;
        nets[i].n_labels = hdr[0];
        nets[i].n_leaves = hdr[1];
        if (hdr[1] == 0) continue;
        nets[i].leaf_off = read_i32(f, (size_t)hdr[1] + 1);
        nets[i].leaf_labels = read_i32(f, (size_t)hdr[2]);
        nets[i].node_left = read_i32(f, (size_t)hdr[1] - 1);
        nets[i].node_right = read_i32(f, (size_t)hdr[1] - 1);
        nets[i].weight_dtype = TB_WEIGHT_UNIT;
        nets[i].value_type = TB_VALUE_AUTO;
    }
    fclose(f);
    tb_options opts;
    memset(&opts, 0, sizeof opts);
    tb_ctx* ctx = NULL;
    int rc = tb_init_multi(devices, n_dev, &opts, &ctx);
    if (rc) { fprintf(stderr, "tb_init_multi: %d %s\n", rc, tb_last_error(NULL)); return 1; }
    double* vals = (double*)calloc((size_t)(n ? n : 1), sizeof(double));
    int32_t* status = (int32_t*)calloc((size_t)(n ? n : 1), sizeof(int32_t));
    double mx = 0;
    rc = tb_contract_networks(ctx, nets, r, n, vals, status, &mx);
    if (rc) { fprintf(stderr, "tb_contract_networks: %d %s\n", rc, tb_last_error(ctx)); return 1; }
    double ms = 0;
    int64_t launches = 0;
    tb_last_timing(ctx, &ms, &launches);
    printf("devices %d max %.17g device_ms %.3f launches %lld\n", tb_device_count(ctx), mx, ms, (long long)launches);
    for (int i = 0; i < n; ++i) printf("%.17g %d\n", vals[i], status[i]);
    tb_shutdown(ctx);
    return 0;
}
