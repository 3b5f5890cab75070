// desc.h -- step descriptors shared by the host plan compiler and the device kernels.
//
// A plan is the device-resident form of one branch's contraction tree (the thing
// uncompress(branch.code) + OMEinsum's recursive executor are in the reference,
// /root/reference/src/types.jl:75-79, src/dynamic_ob.jl:31-32): a flat list of binary contractions
// whose operand layouts were fixed at compile time so that no permutedims pass is needed.
//
// Layout rule (all label sizes are 2, so a layout is a list of labels in address-bit order):
//   for a node  C[M,N,Bt] = max_{K,KA,KB}  A[M,K,KA,Bt] + B[N,K,KB,Bt]
//   A is stored as  [ M_lo (tm bits) | K | KA | M_hi | Bt ]      (bit 0 first)
//   B is stored as  [ N_lo (tn bits) | K | KB | N_hi | Bt ]
//   (packed int16 GEMM nodes: [ K0 | M_lo | K_rest | M_hi | Bt ] / [ K0 | N_lo | K_rest | N_hi | Bt ], "kfirst")
//   i.e. each operand is written by its producer in the order its (single) consumer wants to read it:
//   reads are regular (a GEMM panel is one contiguous block), writes are bit-scattered.
#pragma once
#include <stdint.h>

namespace tb {

enum : uint8_t { LOC_ARENA = 0, LOC_POOL = 1, LOC_SMEM = 2 };
enum : uint8_t { KIND_FUSED = 0, KIND_GENERIC = 1, KIND_GEMM = 2 };
enum : uint8_t { STORE_SCALAR = 0, STORE_VEC_M = 1, STORE_VEC_N = 2 };
enum : uint8_t { OPV_GATHER = 0, OPV_VEC = 1, OPV_BCAST = 2 };  // how the 4 outputs of a vec4 thread map into an operand

constexpr uint8_t NO_BIT = 0xFF;
constexpr int MAX_RANK = 31;        // tensors up to 2^31 elements; shift tables hold 32 entries
constexpr int FUSED_MAX_RANK = 10;  // every tensor inside a fused subtree has rank <= this
constexpr int FUSED_MAX_TC = 16;    // log2 ops of one fused step
constexpr int FUSED_SMEM_ELEMS = 8192;  // 32 KB of 4-byte values per subtree
constexpr int GEMM_TILE_MAX = 7;    // 128 x 128 output tile (4-byte values)
constexpr int GEMM_TILE_MAX_M16 = 8;  // packed int16: 256 x 128 tile, every thread owns 16 (m) x 8 (n) outputs
constexpr int GEMM_STAGE_ELEMS = 4096;  // elements per pipeline stage (A + B panels of all sub-tiles)

// pool layout (elements): [0..3] edge tensor (0,0,0,-inf)  [4] unit scalar (0)  [8+2i, 8+2i+1] vertex i
constexpr int POOL_EDGE = 0;
constexpr int POOL_UNIT = 4;
constexpr int POOL_VERTEX0 = 8;

// One step of a fused subtree (every tensor <= 2^FUSED_MAX_RANK elements, operands in shared memory
// or the leaf pool).  For output address bit i: a_shift[i] / b_shift[i] = bit position of that label
// in A / B (NO_BIT if absent).  The reduction index R has nk shared bits (low), then nka A-only bits,
// then nkb B-only bits; in A the K|KA bits start at bit sa, in B the K|KB bits start at bit sb.
struct SubStep {
    uint16_t a_off, b_off, c_off;  // element offsets (smem or pool); c_off unused when c_loc == LOC_ARENA
    uint8_t a_loc, b_loc, c_loc;
    uint8_t rc, nk, nka, nkb, sa, sb;
    uint8_t pad;  // kfirst: 1 = reduction bit 0 is address bit 0 of A and B, remaining K bits start at sa+1 / sb+1
    uint8_t a_shift[16];
    uint8_t b_shift[16];
};
static_assert(sizeof(SubStep) == 48, "SubStep layout");

struct SubTree {
    int64_t out_off;      // arena element offset of the subtree's result
    uint32_t first_step;  // index into the plan's SubStep array
    uint32_t n_steps;
    uint32_t smem_elems;
    uint32_t pad;
};
static_assert(sizeof(SubTree) == 24, "SubTree layout");

// One non-fused step.
//  KIND_GENERIC: tables indexed by OUTPUT ADDRESS bit i (as SubStep).
//  KIND_GEMM:    tables indexed by canonical output coordinate bit q:
//                q in [0,tm) = M_lo, [tm,tm+tn) = N_lo, then ng grid bits = M_hi | N_hi | Bt.
//                c_shift[q] = bit position in C.  Panels: A panel of grid index g starts at
//                ((g_mhi | g_bt << n_mhi) << (tm+nk)), 2^(tm+nk) contiguous elements laid out [k][m_lo];
//                B likewise with tn / n_nhi.
struct BigStep {
    int64_t a_off, b_off, c_off;  // element offsets in the arena (or pool for leaf operands)
    uint8_t a_loc, b_loc, kind, rc;
    uint8_t nk, nka, nkb, sa, sb, tm, tn, kc;  // kc = log2(k-chunk) for gemm
    uint8_t ng, n_mhi, n_nhi, store_mode;
    uint32_t n_tiles;  // CTAs this step needs
    uint8_t a_shift[32];
    uint8_t b_shift[32];
    uint8_t c_shift[32];
    uint8_t ks;  // generic: log2 threads of a CTA cooperating on one output (block-level split-k)
    uint8_t po;  // generic: log2 outputs per CTA (po + ks <= 8); n_tiles = 2^(rc - po)
    uint8_t lane_n_first;  // gemm: the n tile index varies fastest across the lanes of a warp (else m)
    uint8_t vec4;          // generic: every thread owns 4 consecutive outputs (128-bit stores); po = 10, ks = 0
};
static_assert(sizeof(BigStep) == 144, "BigStep layout");

// per-launch work lists built by the wave executor
struct SubInst {           // one fused subtree of one branch
    const SubStep* steps;
    const void* pool;
    void* out;             // arena address of the result
    uint32_t n_steps;
    uint32_t n_pool;       // pool slots to stage in shared memory (0: read the pool from global memory)
};
struct BigInst {           // one non-fused step of one branch
    const BigStep* step;
    const void* pool;
    void* arena;
    uint32_t tile_start;   // exclusive prefix sum of n_tiles over the launch
    // dataflow launches (all levels of a wave in one persistent kernel): index, in this launch's instance array, of
    // the instance that produces operand A / B, or -1 when the operand is ready before the launch starts (leaf pool,
    // fused subtree, earlier launch).  Instance i is complete when the launch's done[i] counter has reached 0.
    int32_t dep_a, dep_b;
    uint32_t pad;
};
static_assert(sizeof(BigInst) == 40, "BigInst layout");

template <typename T> struct Tropical;
template <> struct Tropical<int32_t> {
    static constexpr int32_t kNegInf = -(1 << 30);
};
template <> struct Tropical<int16_t> {
    static constexpr int32_t kNegInf = -(1 << 14);  // a + b >= -2^15 never wraps; needs sum |w| < 2^13
};

}  // namespace tb
