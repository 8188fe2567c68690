// a3 -- set attention as a three-kernel pipeline (DSVT_ATTN_FP32_TC and DSVT_ATTN_FP16_GEMM):
//
//   1. proj_tile_kernel (roles Q, K, V): per VOXEL (not per set slot) projections on tcgen05,
//        Q = ((x+pos) Wq^T + bq) / sqrt(24),  K = (x+pos) Wk^T + bk,  V = x Wv^T + bv        -> qkv [V, 576] f32
//      Every voxel belongs to exactly one set per axis (getSet's rank formula never lets two sets share a voxel),
//      so projecting voxels instead of the 36 slots of every set removes the padding repeats (2/3 of the slots on the
//      reference frame, ~1/2 at Waymo density) before any arithmetic happens.
//   2. attn_core_kernel: one CTA per set gathers the K/V rows of the set's distinct tokens into shared memory and
//      evaluates scores + key mask -> softmax -> PV in FP32 on the CUDA cores (8.6 % of the FLOPs, 36x36x24 per head:
//      too small for a 128-row UMMA tile)                                                            -> o [V, 192] f32
//   3. proj_tile_kernel (role O): out = o Wout^T + bout, tail rows zero-filled                   -> out [V, 192] f32
//
// Restates multHeadAttention() (reference src/dsvt-ai-trt.cpp:288-458) with GetValueByIndex
// (getValueByIndex.cu:282-303) and MapSetFeature2Voxel (mapSetFeature2voxel.cu:258-275) folded in.
//
// FP32 accuracy on FP16 tensor cores (DSVT_ATTN_FP32_TC): every operand a is split as a = hi + lo with
// hi = fp16(a), lo = fp16(a - hi) (22 significand bits), and a product is evaluated as hi*hi + hi*lo + lo*hi -- three
// tcgen05.mma.kind::f16 per K step accumulating in FP32 in TMEM; the dropped lo*lo term is 2^-22 relative.  Weights
// are pre-scaled by a power of two per role so hi/lo stay in FP16's normal range (undone exactly in the epilogue);
// activations must satisfy |x + pos| < 65504 (they are LayerNorm outputs).  DSVT_ATTN_FP16_GEMM uses the hi terms only.
//
// GEMM kernel (proj_tile_kernel): one CTA = one 128-row x 192-column tile of one role, K streamed in chunks of 32 through a
// 2-stage ring (A: 8 producer warps read FP32 rows with 256-bit loads, add pos, split, and write the UMMA K-major
// interleaved layout of tc_common.cuh conflict-free; B: one cp.async.bulk per chunk of the pre-arranged weight image); one
// thread issues the MMAs; the producer warps then drain the 192 accumulator columns from TMEM (optionally through a chain
// of LayerNorms).  80 KB shared memory and 256 TMEM columns per CTA -> two CTAs per SM overlap each other's phases.
//
// Tile-wide successors (one CTA per 128-row tile and SM, 16 worker warps, a converged issuer warp, a copier; same operand
// images, MMA order and epilogue arithmetic, bit-identical results) are included below and are what the fused frame launches:
//   qkv_fused.cuh  step 1 for the fused FP32_TC entry (x + pos converted once for Q and K, roles in sequence)
//   ffn_fused.cuh  the FFN + the LayerNorm chain behind it; with step 3 + norm1 in front it is the layer tail
//   pos_fused.cuh  the position-embedding MLPs of a frame as roles in sequence
//   vfe_fused.cuh  the pillar feature net (PFN 0, per-pillar max, concat, PFN 1, per-pillar max)
// proj_tile_kernel stays the kernel of the plugin-shaped entry points (q/k/v form, LinearPlugin, dsvt_linear_rows_*).
#include "attention_common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace dsvt {
namespace {

using namespace tc;

constexpr int kC = 192, kH = 8, kD = 24;
constexpr int kBM = 128, kBN = 192, kBK = 32;
constexpr int kNumK = kC / kBK;                      // 6 K chunks
constexpr int kEpiWarps = 8;                         // epilogue warps (TMEM lane quarter = warp % 4, column half = warp / 4)
constexpr int kATerm = kBM * kBK * 2;                // 8192 B: one precision term of an A chunk
constexpr int kBTerm = kBN * kBK * 2;                // 12288 B
constexpr int kWChunkBytes = 2 * kBTerm;             // weight image per (role, K chunk): hi | lo
constexpr int kWRoleBytes = kNumK * kWChunkBytes;    // 147456
constexpr int kRoles = 4;                            // Q, K, V, O
constexpr int kEpiScratch = 32 * 32 * 4;             // 4096 B per epilogue warp: 32 rows x 32 columns, XOR-swizzled float4s

// ---- attention plan: the set partition of one (frame, window partition, axis) in token order ----------------------
// Built once by dsvt_set_attention_plan_launch and shared by every attention layer that uses the partition.
// Per batch item, in ints (each array padded to 64): hdr[64] (hdr[0] = T, number of distinct tokens) |
// set_off[max_sets + 1] (exclusive prefix of tokens per set) | nu[max_sets] | order[max_sets] as int4 (set, token offset,
// token count, 0), sets by descending token count: the core kernel starts the longest sets first | vox_su[max_pillars] (voxel -> set * 64 + u,
// -1 if the voxel is in no set) | tok[max_sets * S] as int2 (voxel row, slot) of the u-th distinct token of each set.
__host__ __device__ inline size_t pad64(size_t n) { return (n + 63) & ~(size_t) 63; }
struct PlanView { int* hdr; int* set_off; int* nu; int4* order; int* vox_su; int2* tok; };
__host__ __device__ inline size_t plan_words(int max_sets, int S, int max_pillars) {
    return 64 + pad64((size_t) max_sets + 1) + 5 * pad64(max_sets) + pad64(max_pillars) + pad64((size_t) 2 * max_sets * S);
}
__host__ __device__ inline PlanView plan_view(int* base, int max_sets, int max_pillars) {
    PlanView v;
    v.hdr = base;
    v.set_off = base + 64;
    v.nu = v.set_off + pad64((size_t) max_sets + 1);
    v.order = reinterpret_cast<int4*>(v.nu + pad64(max_sets));      // (set, token offset, token count, 0), longest sets first
    v.vox_su = v.nu + 5 * pad64(max_sets);
    v.tok = reinterpret_cast<int2*>(v.vox_su + pad64(max_pillars));
    return v;
}

struct GemmRole {
    const float* a0;        // [rows, 192] f32
    const float* a1;        // optional addend (pos), or nullptr
    const uint8_t* wimg;    // this role's weight image
    const float* bias;      // [192]
    float* out;             // row-major, ld_out floats per row
    int ld_out, col0;
    float out_mul;          // 2^-s: undoes the weight pre-scaling
    float post_mul;         // 1/sqrt(C/heads) for q, applied AFTER the biased projection like the reference's division
                            // (:386-405; multiplying by the rounded reciprocal differs from dividing by <= 1 ulp)
    const int* plan;        // attention plan (per-batch stride plan_stride ints) or nullptr: when set, output row of voxel v
    size_t plan_stride;     //   is its TOKEN position set_off[set(v)] + u(v) (set-major order), not v
    int pad_hi;             // floats inserted in front of columns 96..191 (K / V rows: bank-conflict-free head layout)
    int lda;                // floats per row of a0 / a1 (192 for the attention tensors; K for a wider linear layer)
    int accumulate;         // 1: add to the rows already in `out` (second 192-wide K block of a linear layer)
    int act;                // 0: none, 1: GELU (tanh form, gelu.cu:201-211), 2: ReLU on the finished value
    const float* add_src;   // optional rows [rows, ld_out-compatible N] added to the finished value (a residual folded into the
    int ld_add;             //   epilogue: the addElementWise(kSUM) behind the FFN, src/dsvt-ai-trt.cpp:685); nullptr: off
    const float* a0b;       // optional second source of the A operand: columns >= ksplit of the K block come from a0b
    int ksplit, ldb;        //   (the PFN's concatenation [point features | per-pillar max], src/dsvt-ai-trt.cpp:583-587,
                            //   read in place instead of being materialised); ksplit is a multiple of 32, a0b == nullptr: off
    // generated A operand (tile kernel only): row = relu((x2 W1^T) * scale + shift), the narrow first layer of a position-
    // embedding MLP (Linear(2 -> 192) + BatchNorm1d + ReLU, src/dsvt-ai-trt.cpp:461-492) evaluated by the producers instead of
    // being read back from memory.  gen_x [rows, 2]; gen_blob = W1 [192][2] | scale [192] | shift [192] (dsvt_small_linear)
    const float* gen_x;
    const float* gen_blob;
    int kchunks;            // K / 32 (tile kernel only): 6, or 12 for a K = 384 layer whose two weight blocks are consecutive
    // LayerNorm-chain epilogue (tile kernel only, N == 192, no plan): the finished row y0 = acc * out_mul + bias goes through
    //   y = LN_s(y + ln_res[s]) for s < n_ln  (the addElementWise(kSUM) + LayerNormPlugin pairs that follow the attention's
    //   out-projection and the FFN's second linear in the reference graph, src/dsvt-ai-trt.cpp:669-697, :750-756) before it
    //   is stored -- the intermediate tensors never reach memory.  Arithmetic per stage = layer_norm192_kernel (rowwise.cu)
    //   except that the quotient by the standard deviation is a reciprocal-multiply (<= 1 ulp per element).
    int n_ln;
    const float* ln_res[3];     // [rows, 192] each, or nullptr
    const float* ln_gamma[3];
    const float* ln_beta[3];
    float ln_eps;
    int ln_res0_written_here;   // 1: ln_res[0] was written by THIS kernel (the fused out-projection + FFN form): read it through L2
                                //    (ld.global.cg), not through the non-coherent read-only path
    const int* cover;       // optional voxel -> (set, token) map of the attention plan (per-batch stride cover_stride ints):
    size_t cover_stride;    //   rows whose entry is negative belong to no set and are written as exact zeros -- the
                            //   reference scatters the set features into a zero-filled tensor (mapSetFeature2voxel.cu:312),
                            //   so a voxel dropped by a capacity guard must not receive bias + W * (stale workspace row)
};
constexpr int kMaxRoles = 8;                         // roles (blockIdx.y) of one launch: Q / K / V, the column blocks of a wide layer,
struct GemmRoles { GemmRole r[kMaxRoles]; };         // or the eight position-embedding MLPs of a frame

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
}
// 256-bit read-only global load (sm_100: LDG.E.256): p must be 32-byte aligned
__device__ __forceinline__ void ldg256(const float* p, float* d) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7]) : "l"(p));
}
// asynchronous L2 prefetch of a contiguous range (bytes: multiple of 16)
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// GELU, tanh form with the reference's constants (gelu.cu:201-211, params.h:75-77) as the logistic identity
// 0.5 + 0.5 tanh(u) = 1 / (1 + e^(-2u)) -- the f32 formulation of rowwise.cu's GeluPlugin kernel, with the quotient as
// reciprocal-multiply (MUFU.RCP, <= 2 ulp; e^(-2u) = inf gives x * 0 = -0 like the exact division): the IEEE division's
// Newton step + slow-path call per element made the FFN epilogue 2.3x as long as the plain one (21 k vs 9 k cycles per CTA,
// profiles/r2_tile_profile.txt; half of its stall samples were instruction-fetch misses in the unrolled division code)
__device__ __forceinline__ float gelu_tanh(float x) {
    const float u = x * (0.035677408136300125f * x * x + 0.7978845608028654f);
    return __fdividef(x, 1.0f + __expf(-2.0f * u));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// phase stamps of one CTA (tools/split_profile.py): compiled in with -DDSVT_PROFILE only, never in the product build
#ifdef DSVT_PROFILE
__device__ long long g_split_prof[64];
#define CP(i) do { if (blockIdx.x == 5 && blockIdx.y == 0 && threadIdx.x == 0) g_split_prof[32 + i] = clock64(); } while (0)
#define TP(i) do { if (blockIdx.x == 20 && blockIdx.y == 0) g_split_prof[(i)] = clock64(); } while (0)
#else
#define CP(i) do { } while (0)
#define TP(i) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------
// The projection / linear-layer GEMM: one CTA = one 128-row tile of one role, NOT persistent, two CTAs per SM.
//
// Round 1 ran this GEMM as a persistent kernel with a 147 KB resident weight image per CTA (one 229 KB CTA per SM).  At
// these sizes (30 k rows = 241 tiles, 148 SMs) that kernel never reached a steady state -- a CTA computed 1-5 tiles, so the
// arrival of its weight image (4-5 us when every SM pulls the same lines), the first row loads and the final drain were most
// of its life and nothing overlapped them: its ncu source view showed 46 % of the warp samples waiting on global loads and
// 23 % parked at the final barrier (profiles/r2_gemm_persistent_stalls.txt).  Here the weights are NOT resident: each K
// chunk's [hi | lo] image (24 KB, an L2 hit for every CTA but the first) streams through a 2-stage ring next to the
// converted A chunk, the CTA needs 80 KB of shared memory and 192 TMEM columns, and TWO CTAs share an SM so that one's loads
// overlap the other's epilogue.  L2->SM traffic per tile rises from 96-192 KB (rows) to 243-339 KB (rows + weights).
// Operand images, MMA sequence and epilogue arithmetic are unchanged from round 1: results are bit-identical.
constexpr int kLnStride = 196;                        // floats per row of the LayerNorm tile: 49 x 16 B, conflict-free row writes
#ifndef DSVT_TSTAGES
#define DSVT_TSTAGES 2
#endif
constexpr int kTStages = DSVT_TSTAGES;                // converted A chunks
constexpr int kTWStages = 2;                          // weight chunks (a third stage and per-row L2 prefetches of every operand
                                                      // were measured: slower -- the memory system, not latency, is the limit)
constexpr int kTWorkers = 256;                        // warps 0-7: A producers, then the epilogue
constexpr int kTIssuerWarp = 8;                       // warp 8: MMA issue;  warp 9: weight-chunk copies + L2 prefetches
constexpr int kTThreads = 10 * 32;
template <bool SPLIT> struct TLay {
    static constexpr int terms = SPLIT ? 2 : 1;
    static constexpr int a_stage = terms * kATerm;            // 16384 / 8192
    static constexpr int w_stage = terms * kBTerm;            // 24576 / 12288
    static constexpr int w = kTStages * a_stage;              // weight ring behind the A ring
    static constexpr int ring = w + kTWStages * w_stage;      // 106496 / 53248
    static constexpr int ln_tile = kBM * kLnStride * 4;       // 100352: the finished FP32 tile of the LayerNorm epilogue
    static constexpr int plain = ring > kEpiWarps * kEpiScratch ? ring : kEpiWarps * kEpiScratch;   // epilogue scratch aliases the ring
    static constexpr int total = plain > ln_tile ? plain : ln_tile;                                  // ... and so does the LayerNorm tile
};

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm-chain epilogue of a 128 x 192 accumulator tile (shared by the tile GEMM and the fused FFN kernel): the eight
// epilogue warps (TMEM lane quarter = warp & 3, column half = warp >> 2) turn the accumulators at `tmem_acc` into
//   y = LN_s(y + ln_res[s]) for s < n_ln,  y0 = acc * out_mul + bias
// and store the rows once.  `tile` = 128 x kLnStride floats of shared memory that no asynchronous operation still touches.
template <int NW>
__device__ __forceinline__ void ln_chain_epilogue(const GemmRole& g, float* tile, uint32_t tmem_acc, int warp, int lane, int tid,
                                                  int row_base, int V, int b, int max_pillars, float* out, int zero_tails)
{
    static_assert(NW == 8 || NW == 16, "epilogue warps");
    constexpr int kCols = kBN / (NW / 4);          // accumulator columns per warp in pass A: 96 / 48
    constexpr int kRowsW = kBM / NW;               // rows per warp in pass B: 16 / 8
    constexpr int kIters = kRowsW / 2;             // two rows (half-warps) per iteration
    const int q4 = warp & 3, hf = warp >> 2;
    const uint32_t tlane = tmem_acc + ((uint32_t) (q4 * 32) << 16) + hf * kCols;
    const int row0 = row_base + q4 * 32;
    {
        // ---- pass A, TMEM -> finished FP32 rows in shared memory (lane = row) ---------------------------------------
        {
            float* trow = tile + (size_t) (q4 * 32 + lane) * kLnStride + hf * kCols;
            const int grow = row0 + lane;
            const bool is_dead = g.cover && grow < V && __ldg(g.cover + (size_t) b * g.cover_stride + grow) < 0;
#pragma unroll 1
            for (int j0 = 0; j0 < kCols; j0 += 16) {
                uint32_t r[16];
                tmem_ld16(tlane + j0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + hf * kCols + j0 + 4 * j));
                    float4 y = make_float4(__uint_as_float(r[4 * j]) * g.out_mul + bb.x, __uint_as_float(r[4 * j + 1]) * g.out_mul + bb.y,
                                           __uint_as_float(r[4 * j + 2]) * g.out_mul + bb.z, __uint_as_float(r[4 * j + 3]) * g.out_mul + bb.w);
                    if (is_dead) y = make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(trow + j0 + 4 * j) = y;
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");          // the epilogue warps
        if (tid == 0) TP(10);
        // ---- pass B: half a warp per row, the chain of layer_norm192_chain_kernel (rowwise.cu) on the staged rows ---
        const int sub = lane & 15;
        const unsigned hmask = 0xFFFFu << (lane & 16);
        // The work of a half-warp is the list of entries n = (row it, stage st), st fastest.  The residual pieces of entries
        // n + 1 .. n + 3 are in flight while entry n is computed (a rotating queue of three register slots, refilled as soon
        // as a slot is consumed): one row of lead for a 3-stage chain, three rows for norm1 alone.  Issued at their place in
        // program order, the loads put 8 x n_ln dependent memory round trips on the epilogue (27 k of a CTA's 47 k cycles).
        float4 rq[3][3];
        const int n_ln = g.n_ln, total = kIters * n_ln;
        auto res_load = [&](int n, float4 (&d)[3]) {
            const int it = n / n_ln, st = n - it * n_ln;
            const int grow = row_base + warp * kRowsW + it * 2 + (lane >> 4);
            const float* base = n < total ? g.ln_res[st] : nullptr;
            if (base != nullptr && grow < V) {
                const float4* rp = reinterpret_cast<const float4*>(base + ((size_t) b * max_pillars + grow) * kC);
                if (st == 0 && g.ln_res0_written_here) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d[k].x), "=f"(d[k].y), "=f"(d[k].z), "=f"(d[k].w) : "l"(rp + k * 16 + sub));
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) d[k] = ldg_stream4(rp + k * 16 + sub);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) d[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
#pragma unroll
        for (int q = 0; q < 3; ++q) res_load(q, rq[q]);
        float4 v[3];
#pragma unroll 1
        for (int n0 = 0; n0 < total; n0 += 3) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int n = n0 + q;
                if (n >= total) break;
                const int it = n / n_ln, st = n - it * n_ln;
                const int rloc = warp * kRowsW + it * 2 + (lane >> 4), grow = row_base + rloc;
                if (st == 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) v[k] = *reinterpret_cast<const float4*>(tile + (size_t) rloc * kLnStride + (k * 16 + sub) * 4);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) { v[k].x += rq[q][k].x; v[k].y += rq[q][k].y; v[k].z += rq[q][k].z; v[k].w += rq[q][k].w; }
                res_load(n + 3, rq[q]);
                float sum = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(hmask, sum, o);
                const float mean = sum / 192.f;
                float qs = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float a = v[k].x - mean, c = v[k].y - mean, d = v[k].z - mean, e = v[k].w - mean;
                    qs += (a * a + c * c) + (d * d + e * e);
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) qs += __shfl_xor_sync(hmask, qs, o);
                // (v - mean) / sd as (v - mean) * (1 / sd): <= 1 ulp per element from layer_norm192_kernel's quotient, 11 IEEE
                // divisions per lane and stage less (the pass is bound by instruction issue, not by memory)
                const float inv_sd = 1.0f / sqrtf(qs / 192.f + g.ln_eps);
                const float4* gp = reinterpret_cast<const float4*>(g.ln_gamma[st]);
                const float4* bp = reinterpret_cast<const float4*>(g.ln_beta[st]);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 ga = __ldg(gp + k * 16 + sub), be = __ldg(bp + k * 16 + sub);
                    v[k].x = (v[k].x - mean) * inv_sd * ga.x + be.x;
                    v[k].y = (v[k].y - mean) * inv_sd * ga.y + be.y;
                    v[k].z = (v[k].z - mean) * inv_sd * ga.z + be.z;
                    v[k].w = (v[k].w - mean) * inv_sd * ga.w + be.w;
                }
                if (st == n_ln - 1) {
                    float4* orow4 = reinterpret_cast<float4*>(out + (size_t) grow * g.ld_out);
                    if (grow < V) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) stg_stream4(orow4 + k * 16 + sub, v[k]);
                    } else if (zero_tails && grow < max_pillars) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) stg_zero4(orow4 + k * 16 + sub);
                    }
                    if (tid == 0 && it == kIters / 2 - 1) TP(11);
                }
            }
        }
    }
}

template <bool SPLIT>
__global__ void __launch_bounds__(kTThreads, 2)
proj_tile_kernel(const __grid_constant__ GemmRoles roles, const int* __restrict__ voxel_num, int rows_host, int max_pillars,
                 int max_sets, int zero_tails)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t w_full[kTWStages], w_empty[kTWStages], a_full[kTStages], s_empty[kTStages], acc_full;
    __shared__ uint32_t tmem_slot;
    __shared__ float4 s_gen[kC];                     // per column (w0, w1, scale, shift) of a generated A operand
    using L = TLay<SPLIT>;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, b = blockIdx.z;
    const GemmRole& g = roles.r[blockIdx.y];       // read in place from the parameter space (no local copy of the struct)
    int V = voxel_num ? voxel_num[b] : rows_host;
    V = V < max_pillars ? V : max_pillars;
    const int row_base = tile * kBM;
    float* out = g.out + (size_t) b * max_pillars * g.ld_out + g.col0;
    if (row_base >= V) {                                  // a tile of tail rows: zero-filled (plugin outputs), or nothing to do
        if (zero_tails)
            for (int i = tid; i < kBM * (kBN / 4); i += kTThreads) {
                const int rloc = i / (kBN / 4), cc4 = i - rloc * (kBN / 4);
                if (row_base + rloc < max_pillars)
                    stg_zero4(reinterpret_cast<float4*>(out + (size_t) (row_base + rloc) * g.ld_out + cc4 * 4));
            }
        return;
    }
    const float* a0 = g.a0 + (size_t) b * max_pillars * g.lda;
    const float* a1 = g.a1 ? g.a1 + (size_t) b * max_pillars * g.lda : nullptr;
    const float* a0b = g.a0b ? g.a0b + (size_t) b * max_pillars * g.ldb : nullptr;
    if (tid == 0) TP(0);
#ifdef DSVT_EXP_IMG        // experiment: the A operand arrives as ready-made 16 KB chunk images (bulk copies of the same bytes)
    const bool a_img = !g.gen_x && !g.a0b;
#else
    const bool a_img = false;
#endif

    if (tid == 0) {
        for (int s = 0; s < kTStages; ++s) { mbar_init(&a_full[s], a_img ? 1 : kTWorkers); mbar_init(&s_empty[s], 1); }
        for (int s = 0; s < kTWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(&acc_full, 1);
        fence_barrier_init();
    }
    if (warp == kTIssuerWarp) tmem_alloc<256>(&tmem_slot);
    if (g.gen_x)
        for (int t = tid; t < kC; t += kTThreads)
            s_gen[t] = make_float4(__ldg(g.gen_blob + 2 * t), __ldg(g.gen_blob + 2 * t + 1), __ldg(g.gen_blob + 2 * kC + t),
                                   __ldg(g.gen_blob + 3 * kC + t));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) TP(1);

    if (warp < 8) {
        // =========================== A PRODUCERS =========================================================
        // step s = (K chunk kc = s >> 1, half = s & 1): rows half*64 + warp*8 + (lane & 7), 16-byte K piece c16 = lane >> 3
        constexpr int kDepth = 3, kUnroll = 6;                  // lcm(2 halves, kDepth): static buffer / half indices
        const int n_steps = g.kchunks * 2;                      // a multiple of 12
        const int rl = warp * 8 + (lane & 7), c16 = lane >> 3;
        float buf[kDepth][16];                                  // [0..7] = a0 row piece, [8..15] = a1 (pos) row piece
        const float* gen_x = g.gen_x ? g.gen_x + (size_t) b * max_pillars * 2 : nullptr;
        float2 gxy[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};      // the two rows of this thread (half 0 / 1)
        if (gen_x) {
#pragma unroll
            for (int hf2 = 0; hf2 < 2; ++hf2)
                if (row_base + hf2 * 64 + rl < V) gxy[hf2] = __ldg(reinterpret_cast<const float2*>(gen_x) + row_base + hf2 * 64 + rl);
        }
        auto issue = [&](int s, float (&d)[16]) {
            const int kc = s >> 1, row = row_base + (s & 1) * 64 + rl;
            if (row < V && gen_x) {               // same arithmetic as small_linear_kernel<2>: FC, Scale (folded BatchNorm), ReLU
                const int col = kc * kBK + c16 * 8;
                const float2 xy = gxy[s & 1];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 wv = s_gen[col + e];
                    float acc = fmaf(xy.x, wv.x, 0.f);
                    acc = fmaf(xy.y, wv.y, acc);
                    acc = fmaf(acc, wv.z, wv.w);
                    d[e] = fmaxf(acc, 0.f);
                    d[8 + e] = 0.f;
                }
            }
#ifdef DSVT_EXP_NOA       // experiment: the A operand is not read (wrong results; what the row loads cost)
            else if (row < V) {
#pragma unroll
                for (int e = 0; e < 16; ++e) d[e] = 1.f;
            }
#endif
            else if (row < V) {
                const int col = kc * kBK + c16 * 8;
                if (a0b && col >= g.ksplit) ldg256(a0b + (size_t) row * g.ldb + (col - g.ksplit), &d[0]);
                else ldg256(a0 + (size_t) row * g.lda + col, &d[0]);
                if (a1) ldg256(a1 + (size_t) row * g.lda + col, &d[8]);
                else {
#pragma unroll
                    for (int e = 8; e < 16; ++e) d[e] = 0.f;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) d[e] = 0.f;
            }
        };
        if (!a_img) {
#pragma unroll
        for (int s = 0; s < kDepth - 1; ++s) issue(s, buf[s]);
        }
#pragma unroll 1
        for (int s0 = 0; s0 < (a_img ? 0 : n_steps); s0 += kUnroll) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int s = s0 + u;
                if (s + kDepth - 1 < n_steps) issue(s + kDepth - 1, buf[(u + kDepth - 1) % kDepth]);
                const int cc = s >> 1, st = cc % kTStages, r = (u & 1) * 64 + rl;
                if ((u & 1) == 0 && cc >= kTStages) mbar_wait(&s_empty[st], ((cc / kTStages) - 1) & 1);
                float (&d)[16] = buf[u % kDepth];
                const float v[8] = {d[0] + d[8], d[1] + d[9], d[2] + d[10], d[3] + d[11],
                                    d[4] + d[12], d[5] + d[13], d[6] + d[14], d[7] + d[15]};
                const uint4 hi = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                uint8_t* stage = smem + st * L::a_stage;
                *reinterpret_cast<uint4*>(stage + c16 * (kBM * 16) + r * 16) = hi;
                if (SPLIT) {
                    const float2 h0 = unpack_h2(hi.x), h1 = unpack_h2(hi.y), h2 = unpack_h2(hi.z), h3 = unpack_h2(hi.w);
                    const uint4 lo = make_uint4(pack_h2(v[0] - h0.x, v[1] - h0.y), pack_h2(v[2] - h1.x, v[3] - h1.y),
                                                pack_h2(v[4] - h2.x, v[5] - h2.y), pack_h2(v[6] - h3.x, v[7] - h3.y));
                    *reinterpret_cast<uint4*>(stage + kATerm + c16 * (kBM * 16) + r * 16) = lo;
                }
                if (u & 1) {
                    fence_proxy_async_smem();
                    mbar_arrive(&a_full[st]);
                    if (tid == 0 && cc < 6) TP(2 + cc);
                }
            }
        }

        // =========================== EPILOGUE (same warps) ===============================================
        // warp = (TMEM lane quarter q4, column half hf): 3 slabs of 32 columns.  Slab: TMEM -> registers (lane = row)
        // -> swizzled scratch -> (lane = 4 columns of 8 rows) scale + bias -> full 128-byte row segments to global.
        const int q4 = warp & 3, hf = warp >> 2;
        const uint32_t tlane = tmem + ((uint32_t) (q4 * 32) << 16) + hf * 96;
        const int rg = lane >> 3, c4 = lane & 7;
        const float* bias = g.bias + hf * 96 + c4 * 4;
        float* outc = out + hf * (96 + g.pad_hi) + c4 * 4;
        const PlanView pv = plan_view(const_cast<int*>(g.plan) + (size_t) b * g.plan_stride, max_sets, max_pillars);
        const int row0 = row_base + q4 * 32;
        int orow[8];                                       // output row of this lane's 8 rows, -1: not written
        unsigned dead = 0;                                 // bit rr: a valid row that belongs to no set -> zeros
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            const int grow = row0 + rr * 4 + rg;
            if (g.cover && grow < V && __ldg(g.cover + (size_t) b * g.cover_stride + grow) < 0) dead |= 1u << rr;
            if (g.plan) {                                  // voxel row -> token position (set-major order)
                orow[rr] = -1;
                if (grow < V) {
                    const int su = __ldg(pv.vox_su + grow);
                    if (su >= 0) {
                        const int t = __ldg(pv.set_off + (su >> 6)) + (su & 63);
                        if (t < max_pillars) orow[rr] = t;
                    }
                }
            } else {
                orow[rr] = (grow < V || (zero_tails && grow < max_pillars)) ? grow : -1;
            }
        }
        float4 bias3[3];                                   // loaded BEFORE the accumulators are waited for: a load issued per slab
#pragma unroll                                             // would put one L2 round trip on every slab's critical path
        for (int j = 0; j < 3; ++j) bias3[j] = __ldg(reinterpret_cast<const float4*>(bias + j * 32));
        if (tid == 0) TP(8);
        mbar_wait(&acc_full, 0);                           // every MMA has completed: the ring is dead, the scratch may alias it
        tc_fence_after_sync();
        if (tid == 0) TP(9);
        if (g.n_ln > 0) {
            ln_chain_epilogue<8>(g, reinterpret_cast<float*>(smem), tmem, warp, lane, tid, row_base, V, b, max_pillars, out, zero_tails);
        } else {
        float4* scr = reinterpret_cast<float4*>(smem + warp * kEpiScratch);
#pragma unroll 1
        for (int j0 = 0; j0 < 96; j0 += 32) {
            // rows added in the epilogue (second K block / residual): all eight loads of the slab are in flight before the
            // accumulators are touched (they were the longest stall of the persistent kernel's epilogue)
            float4 extra[8];
            const bool has_acc = g.accumulate != 0, has_add = g.add_src != nullptr;
            if (has_acc || has_add) {
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    const int grow = row0 + rr * 4 + rg;
                    extra[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (orow[rr] >= 0 && grow < V) {
                        if (has_acc) extra[rr] = *reinterpret_cast<const float4*>(outc + (size_t) orow[rr] * g.ld_out + j0);
                        if (has_add) {
                            const float4 ad = __ldg(reinterpret_cast<const float4*>(
                                g.add_src + ((size_t) b * max_pillars + grow) * g.ld_add + g.col0 + hf * 96 + c4 * 4 + j0));
                            extra[rr].x += ad.x; extra[rr].y += ad.y; extra[rr].z += ad.z; extra[rr].w += ad.w;
                        }
                    }
                }
            }
            const float4 bb = j0 == 0 ? bias3[0] : (j0 == 32 ? bias3[1] : bias3[2]);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {               // two 16-column loads: keeps the live registers under the cap
                uint32_t r[16];
                tmem_ld16(tlane + j0 + hh * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j)                // float4 j of row `lane` lands in slot j ^ (lane & 7)
                    scr[lane * 8 + ((hh * 4 + j) ^ (lane & 7))] =
                        make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                    __uint_as_float(r[4 * j + 3]));
            }
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {               // 4 rows x 128 contiguous bytes per store instruction
                const int grow = row0 + rr * 4 + rg, rloc = rr * 4 + rg;
                const float4 v = scr[rloc * 8 + (c4 ^ (rloc & 7))];
                // out_mul is a power of two: the product is exact, so this is one rounding of (acc + bias)
                float4 ov = make_float4((v.x * g.out_mul + bb.x) * g.post_mul, (v.y * g.out_mul + bb.y) * g.post_mul,
                                        (v.z * g.out_mul + bb.z) * g.post_mul, (v.w * g.out_mul + bb.w) * g.post_mul);
                if (has_acc || has_add) { ov.x += extra[rr].x; ov.y += extra[rr].y; ov.z += extra[rr].z; ov.w += extra[rr].w; }
                if (g.act == 1) { ov.x = gelu_tanh(ov.x); ov.y = gelu_tanh(ov.y); ov.z = gelu_tanh(ov.z); ov.w = gelu_tanh(ov.w); }
                else if (g.act == 2) { ov.x = fmaxf(ov.x, 0.f); ov.y = fmaxf(ov.y, 0.f); ov.z = fmaxf(ov.z, 0.f); ov.w = fmaxf(ov.w, 0.f); }
                if (grow >= V || (dead >> rr & 1u)) ov = make_float4(0.f, 0.f, 0.f, 0.f);
                if (orow[rr] >= 0) *reinterpret_cast<float4*>(outc + (size_t) orow[rr] * g.ld_out + j0) = ov;
            }
            __syncwarp();
            if (tid == 0) TP(10 + (j0 >> 5));
        }
        }
    } else if (warp == kTIssuerWarp) {
        // =========================== MMA ISSUE ===========================================================
        if (lane == 0) {       // (the converged-warp issue of ffn_fused.cuh was measured here: no gain, six MMAs per wait)
            const uint32_t idesc = make_idesc(kFmtF16, kBM, kBN);
            const uint32_t sbase = smem_u32(smem);
#pragma unroll 1
            for (int kc = 0; kc < g.kchunks; ++kc) {
                const int st = kc % kTStages, ws = kc % kTWStages;
                mbar_wait(&w_full[ws], (kc / kTWStages) & 1);
                if (kc < 6) TP(14 + kc);
                mbar_wait(&a_full[st], (kc / kTStages) & 1);
                tc_fence_after_sync();
                if (kc < 6) TP(20 + kc);
                const uint32_t sa = sbase + st * L::a_stage, sw = sbase + L::w + ws * L::w_stage;
#pragma unroll
                for (int ks = 0; ks < kBK / 16; ++ks) {
                    const uint64_t a_hi = make_smem_desc(sa + ks * 2 * (kBM * 16), kBM * 16, 128);
                    const uint64_t b_hi = make_smem_desc(sw + ks * 2 * (kBN * 16), kBN * 16, 128);
                    if (SPLIT) {
                        const uint64_t a_lo = make_smem_desc(sa + kATerm + ks * 2 * (kBM * 16), kBM * 16, 128);
                        const uint64_t b_lo = make_smem_desc(sw + kBTerm + ks * 2 * (kBN * 16), kBN * 16, 128);
                        umma_f16(tmem, a_lo, b_hi, idesc, (kc | ks) != 0);
                        umma_f16(tmem, a_hi, b_lo, idesc, 1);
                        umma_f16(tmem, a_hi, b_hi, idesc, 1);
                    } else {
                        umma_f16(tmem, a_hi, b_hi, idesc, (kc | ks) != 0);
                    }
                }
                umma_commit(&s_empty[st]);
                umma_commit(&w_empty[ws]);
            }
            umma_commit(&acc_full);
        }
        __syncwarp();
    } else {
        // =========================== WEIGHT-CHUNK COPIES =================================================
        const int nrows = V - row_base < kBM ? V - row_base : kBM;
        if (lane == 0) {
            const uint64_t w_policy = l2_policy_evict_last();
            if (g.lda == kC && !g.gen_x) {   // the tile's rows are one contiguous block: pull them into L2 as large sequential requests
                const uint32_t bytes = (uint32_t) (nrows * kC * sizeof(float));
                l2_prefetch(a0 + (size_t) row_base * kC, bytes);
                if (a1) l2_prefetch(a1 + (size_t) row_base * kC, bytes);
            }
            // residual rows of the LayerNorm-chain epilogue: in L2 by the time the accumulators are complete
            for (int st = 0; st < g.n_ln; ++st)
                if (g.ln_res[st] != nullptr)
                    l2_prefetch(g.ln_res[st] + ((size_t) b * max_pillars + row_base) * kC, (uint32_t) (nrows * kC * sizeof(float)));
#pragma unroll 1
            for (int kc = 0; kc < g.kchunks; ++kc) {
                const int ws = kc % kTWStages;
                if (kc >= kTWStages) mbar_wait(&w_empty[ws], ((kc / kTWStages) - 1) & 1);
#ifdef DSVT_EXP_NOW       // experiment: weight chunks are not streamed (wrong results; what the W stream costs)
                if (kc >= kTWStages) { mbar_arrive(&w_full[ws]); continue; }
#endif
#ifdef DSVT_EXP_NOW2      // experiment: no weight chunk is copied at all (what the FIRST chunk's arrival costs)
                if (a_img) {
                    const int st = kc % kTStages;
                    if (kc >= kTStages) mbar_wait(&s_empty[st], ((kc / kTStages) - 1) & 1);
                    mbar_arrive_expect_tx(&a_full[st], L::a_stage);
                    bulk_g2s(smem + st * L::a_stage, reinterpret_cast<const uint8_t*>(a0) + ((size_t) tile * g.kchunks + kc) * L::a_stage,
                             L::a_stage, &a_full[st]);
                }
                mbar_arrive(&w_full[ws]); continue;
#endif
                if (a_img) {
                    const int st = kc % kTStages;
                    if (kc >= kTStages) mbar_wait(&s_empty[st], ((kc / kTStages) - 1) & 1);
                    mbar_arrive_expect_tx(&a_full[st], L::a_stage);
                    bulk_g2s(smem + st * L::a_stage, reinterpret_cast<const uint8_t*>(a0) + ((size_t) tile * g.kchunks + kc) * L::a_stage,
                             L::a_stage, &a_full[st]);
                }
                mbar_arrive_expect_tx(&w_full[ws], L::w_stage);
                bulk_g2s_hint(smem + L::w + ws * L::w_stage, g.wimg + (size_t) kc * kWChunkBytes, L::w_stage, &w_full[ws], w_policy);
            }
        }
        __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) TP(13);
    if (warp == kTIssuerWarp) tmem_dealloc<256>(tmem);
}

#include "ffn_fused.cuh"
#include "vfe_fused.cuh"

// ---------------------------------------------------------------------------------------------------------------
// Plan kernel 1: one warp per set.  Token compaction, same rule as attention_fp32.cu: a slot that repeats the previous
// voxel AND is masked as a key by every head (getSet.cu:546-563) is the same token as its twin.
template <int S>
__global__ void __launch_bounds__(256)
attn_plan_sets_kernel(const int* __restrict__ idx, const float* __restrict__ mask, const int* __restrict__ set_num,
                      int* __restrict__ plan, size_t plan_stride, int max_sets, int max_pillars, int axis)
{
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int set = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (set >= max_sets) return;
    const PlanView pv = plan_view(plan + (size_t) b * plan_stride, max_sets, max_pillars);
    int ns = set_num[b];
    ns = ns < max_sets ? ns : max_sets;
    if (set >= ns) { if (lane == 0) pv.nu[set] = 0; return; }
    const int* my_idx = idx + (((size_t) b * 2 + axis) * max_sets + set) * S;
    const float* my_mask = mask + ((size_t) b * max_sets + set) * kH * S;
    int base_u = 0;
#pragma unroll
    for (int k0 = 0; k0 < S; k0 += 32) {
        const int k = k0 + lane;
        bool keep = false;
        int gidx = 0;
        if (k < S) {
            gidx = my_idx[k];
            keep = (k == 0) || (gidx != my_idx[k - 1]);
            if (!keep)
#pragma unroll
                for (int hh = 0; hh < kH; ++hh) keep |= !(my_mask[hh * S + k] < -1e30f);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int u = base_u + __popc(bal & ((1u << lane) - 1u));
            pv.tok[(size_t) set * S + u] = make_int2(gidx, k);
            if (gidx >= 0 && gidx < max_pillars) pv.vox_su[gidx] = set * 64 + u;
        }
        base_u += __popc(bal);
    }
    if (lane == 0) pv.nu[set] = base_u;
}

// Plan kernel 2: exclusive prefix of the tokens per set (one CTA per batch item), clamped to the row capacity.
__global__ void __launch_bounds__(1024)
attn_plan_scan_kernel(const int* __restrict__ set_num, int* __restrict__ plan, size_t plan_stride, int max_sets,
                      int max_pillars)
{
    __shared__ int warp_sums[33];
    const int b = blockIdx.x;
    const PlanView pv = plan_view(plan + (size_t) b * plan_stride, max_sets, max_pillars);
    int ns = set_num[b];
    ns = ns < max_sets ? ns : max_sets;
    int carry = 0;
    for (int i0 = 0; i0 < ns; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int v = i < ns ? pv.nu[i] : 0;
        int total;
        const int excl = block_excl_scan(v, warp_sums, &total) + carry;
        if (i < ns) pv.set_off[i] = excl < max_pillars ? excl : max_pillars;
        carry += total;
    }
    if (threadIdx.x == 0) {
        const int T = carry < max_pillars ? carry : max_pillars;
        pv.set_off[ns] = T;
        pv.hdr[0] = T;
    }
    // order[]: sets by descending token count (counting sort over the <= 65 possible counts; any order within a count).
    // The core kernel's cost per set grows with the square of the count: longest-first removes its tail.
    __shared__ int bucket[65];
    if (threadIdx.x < 65) bucket[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < ns; i += 1024) atomicAdd(&bucket[min(pv.nu[i], 64)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = 64; c >= 0; --c) { const int n = bucket[c]; bucket[c] = run; run += n; }
    }
    __syncthreads();
    // one 16-byte record per set -- (set, token offset, token count): the core kernel's CTA finds everything it needs to issue
    // its K / V copy with ONE load instead of a chain of three dependent ones (order -> set_off[set] -> set_off[set + 1])
    for (int i = threadIdx.x; i < ns; i += 1024) {
        const int off = pv.set_off[i];
        pv.order[atomicAdd(&bucket[min(pv.nu[i], 64)], 1)] = make_int4(i, off, pv.set_off[i + 1] - off, 0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Per-set attention core.  One CTA per set (grid-stride), thread = (token pair, head h) with h fastest: a warp is 4 token
// pairs x 8 heads.  The QKV GEMM has written Q and [K | V] rows in TOKEN order (set-major), so a set's keys and values
// are ONE contiguous block that lands in shared memory with a single cp.async.bulk; the chain of dependent global
// round trips per CTA is  plan entry -> {bulk copy, q rows, token list, mask}  (the first version gathered rows by
// voxel id after compacting the set itself: four dependent trips, which made the kernel latency-bound).
// Each K/V piece read from shared memory feeds TWO queries (register tiling: the kernel is bound by instruction issue,
// not by FLOPs), the dot products and the PV update use the packed FFMA2 (two FP32 FMAs per instruction, same IEEE
// arithmetic).  A K/V read is 8 distinct 16-byte pieces (one per head) broadcast to the 4 pairs; heads 4..7 sit 16 bytes
// further (row = 96 | 4 pad | 96 floats), which puts the 8 pieces into 8 different bank groups.
// Online softmax over key chunks of 4 in the log2 domain (one rescale per chunk): K, V are read once per query pair.
#ifndef DSVT_CORE_MINB
#define DSVT_CORE_MINB 3   // 128 registers: three CTAs (15 warps) per SM measured faster than two at 168 registers
#endif
constexpr int kKvHalf = 96 + 4;          // floats: heads 0-3 | pad
constexpr int kKvRow = 2 * kKvHalf - 4;  // 196 floats per K (or V) row
constexpr int kKvTok = 2 * kKvRow;       // 392 floats per token: K row | V row

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
constexpr float kLog2e = 1.4426950408889634f;
// 2^x on the SFU (one MUFU.EX2; 2^-inf = +0): the softmax runs in the log2 domain, exp(s - m) = 2^((s - m) log2 e)
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int S>
__global__ void __launch_bounds__(S / 2 * kH, DSVT_CORE_MINB)
attn_core_kernel(const float* __restrict__ qbuf, const float* __restrict__ kvbuf, const int* __restrict__ plan,
                 size_t plan_stride, const float* __restrict__ mask, const int* __restrict__ set_num,
                 float* __restrict__ o, int max_sets, int max_pillars)
{
    static_assert(S % 4 == 0 && S <= 64, "set size");
    constexpr int NT = S / 2 * kH;      // 144 threads for S = 36
    extern __shared__ __align__(128) float sm[];        // [S][392]: K row | V row per token
    __shared__ __align__(8) uint64_t kv_bar;
    __shared__ int s_rows[S], s_slot[S];
    __shared__ float s_mask[kH * S];    // the set's additive key mask as given: [head][slot]
    __shared__ float s_cmask[kH][S];    // ... compacted, log2 domain: [head][token]

    const int b = blockIdx.y, tid = threadIdx.x;
    CP(0);
    const PlanView pv = plan_view(const_cast<int*>(plan) + (size_t) b * plan_stride, max_sets, max_pillars);
    qbuf += (size_t) b * max_pillars * kC;
    kvbuf += (size_t) b * max_pillars * kKvTok;
    o += (size_t) b * max_pillars * kC;
    const int pr = tid >> 3, h = tid & 7;
    const int i0 = 2 * pr, i1 = i0 + 1;
    const int hoff = h * kD + (h >= 4 ? 4 : 0);
    if (tid == 0) { mbar_init(&kv_bar, 1); fence_barrier_init(); }
    __syncthreads();
    int ns = set_num[b];
    ns = ns < max_sets ? ns : max_sets;
    uint32_t phase = 0;

    for (int si = blockIdx.x; si < ns; si += gridDim.x, phase ^= 1) {
        const int4 rec = __ldg(pv.order + si);                         // longest sets first: (set, token offset, token count)
        const int set = rec.x;
#ifdef DSVT_CORE_PHASE_PROFILE
        const long long t_begin = clock64();
#endif
        const int off = rec.y;
        const int nu = rec.z;                                       // distinct tokens of the set (clamped to the row capacity)
        if (nu <= 0) { phase ^= 1; continue; }
        const int nu4 = (nu + 3) & ~3;                              // keys are processed in chunks of 4
        if (tid == 0) {
            fence_proxy_async_smem();                               // earlier generic accesses to the tile -> async-proxy write
            mbar_arrive_expect_tx(&kv_bar, (uint32_t) nu * kKvTok * 4);
            bulk_g2s(sm, kvbuf + (size_t) off * kKvTok, (uint32_t) nu * kKvTok * 4, &kv_bar);
        }
        for (int t = tid; t < kH * S; t += NT) s_mask[t] = __ldg(mask + ((size_t) b * max_sets + set) * kH * S + t);
        if (tid < nu) {
            const int2 tk = __ldg(pv.tok + (size_t) set * S + tid);
            s_rows[tid] = tk.x;
            s_slot[tid] = tk.y;
        }
        float2 q0[kD / 2], q1[kD / 2];
        if (i0 < nu) {
            const float4* qp = reinterpret_cast<const float4*>(qbuf + (size_t) (off + i0) * kC + h * kD);
            const float4* qp1 = reinterpret_cast<const float4*>(qbuf + (size_t) (off + (i1 < nu ? i1 : i0)) * kC + h * kD);
#pragma unroll
            for (int d4 = 0; d4 < kD / 4; ++d4) {
                const float4 t = __ldg(qp + d4), u = __ldg(qp1 + d4);
                q0[2 * d4] = f2(t.x, t.y); q0[2 * d4 + 1] = f2(t.z, t.w);
                q1[2 * d4] = f2(u.x, u.y); q1[2 * d4 + 1] = f2(u.z, u.w);
            }
        }
        for (int t = tid; t < (nu4 - nu) * (kKvTok / 4); t += NT)   // pad keys: zero rows (masked with -inf below)
            reinterpret_cast<float4*>(sm + (size_t) nu * kKvTok)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        for (int t = tid; t < kH * S; t += NT) {                    // key mask in the log2 domain of the softmax
            const int hh = t / S, j = t % S;
            s_cmask[hh][j] = j < nu ? s_mask[hh * S + s_slot[j]] * kLog2e : -INFINITY;
        }
        mbar_wait(&kv_bar, phase);
        __syncthreads();
        if (si == (int) blockIdx.x) CP(2);
#ifdef DSVT_CORE_PHASE_PROFILE
        long long t_staged = 0;
        if (tid == 0) { t_staged = clock64(); atomicAdd((unsigned long long*) &g_split_prof[36], (unsigned long long) (t_staged - t_begin)); }
#endif

        if (i0 < nu) {
            float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
            float2 acc0[kD / 2], acc1[kD / 2];
#pragma unroll
            for (int d2 = 0; d2 < kD / 2; ++d2) { acc0[d2] = f2(0.f, 0.f); acc1[d2] = f2(0.f, 0.f); }
#pragma unroll 1
            for (int j0 = 0; j0 < nu; j0 += 4) {
                // branch-free chunk of 4 keys (pad keys carry zero rows and a -inf mask): scores in the log2 domain
                float sc0[4], sc1[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = j0 + jj;
                    const float4* kp = reinterpret_cast<const float4*>(sm + j * kKvTok + hoff);
                    float2 a0 = f2(0.f, 0.f), a1 = f2(0.f, 0.f), b0 = f2(0.f, 0.f), b1 = f2(0.f, 0.f);
#pragma unroll
                    for (int d4 = 0; d4 < kD / 4; ++d4) {
                        const float4 kv = kp[d4];
                        const float2 klo = f2(kv.x, kv.y), khi = f2(kv.z, kv.w);
                        a0 = __ffma2_rn(q0[2 * d4], klo, a0); a1 = __ffma2_rn(q0[2 * d4 + 1], khi, a1);
                        b0 = __ffma2_rn(q1[2 * d4], klo, b0); b1 = __ffma2_rn(q1[2 * d4 + 1], khi, b1);
                    }
                    const float mk = s_cmask[h][j];
                    sc0[jj] = fmaf((a0.x + a0.y) + (a1.x + a1.y), kLog2e, mk);
                    sc1[jj] = fmaf((b0.x + b0.y) + (b1.x + b1.y), kLog2e, mk);
                }
                const float m0n = fmaxf(fmaxf(m0, fmaxf(sc0[0], sc0[1])), fmaxf(sc0[2], sc0[3]));
                const float m1n = fmaxf(fmaxf(m1, fmaxf(sc1[0], sc1[1])), fmaxf(sc1[2], sc1[3]));
                const float al0 = ex2(m0 - m0n), al1 = ex2(m1 - m1n);            // first chunk: 2^-inf = 0
                l0 *= al0; l1 *= al1;
                const float2 al02 = f2(al0, al0), al12 = f2(al1, al1);
#pragma unroll
                for (int d2 = 0; d2 < kD / 2; ++d2) { acc0[d2] = __fmul2_rn(acc0[d2], al02); acc1[d2] = __fmul2_rn(acc1[d2], al12); }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = j0 + jj;
                    const float p0 = ex2(sc0[jj] - m0n), p1 = ex2(sc1[jj] - m1n);
                    l0 += p0; l1 += p1;
                    const float2 p02 = f2(p0, p0), p12 = f2(p1, p1);
                    const float4* vp = reinterpret_cast<const float4*>(sm + j * kKvTok + kKvRow + hoff);
#pragma unroll
                    for (int d4 = 0; d4 < kD / 4; ++d4) {
                        const float4 vv = vp[d4];
                        const float2 vlo = f2(vv.x, vv.y), vhi = f2(vv.z, vv.w);
                        acc0[2 * d4] = __ffma2_rn(p02, vlo, acc0[2 * d4]); acc0[2 * d4 + 1] = __ffma2_rn(p02, vhi, acc0[2 * d4 + 1]);
                        acc1[2 * d4] = __ffma2_rn(p12, vlo, acc1[2 * d4]); acc1[2 * d4 + 1] = __ffma2_rn(p12, vhi, acc1[2 * d4 + 1]);
                    }
                }
                m0 = m0n; m1 = m1n;
            }
            const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
            const int r0 = s_rows[i0];
            if (r0 >= 0 && r0 < max_pillars) {
                float4* op = reinterpret_cast<float4*>(o + (size_t) r0 * kC + h * kD);
#pragma unroll
                for (int d4 = 0; d4 < kD / 4; ++d4)
                    op[d4] = make_float4(acc0[2 * d4].x * inv0, acc0[2 * d4].y * inv0, acc0[2 * d4 + 1].x * inv0, acc0[2 * d4 + 1].y * inv0);
            }
            const int r1 = i1 < nu ? s_rows[i1] : -1;
            if (r1 >= 0 && r1 < max_pillars) {
                float4* op1 = reinterpret_cast<float4*>(o + (size_t) r1 * kC + h * kD);
#pragma unroll
                for (int d4 = 0; d4 < kD / 4; ++d4)
                    op1[d4] = make_float4(acc1[2 * d4].x * inv1, acc1[2 * d4].y * inv1, acc1[2 * d4 + 1].x * inv1, acc1[2 * d4 + 1].y * inv1);
            }
        }
        if (si == (int) blockIdx.x) CP(3);
        __syncthreads();        // shared memory is recycled by the next set
#ifdef DSVT_CORE_PHASE_PROFILE
        if (tid == 0) {
            atomicAdd((unsigned long long*) &g_split_prof[37], (unsigned long long) (clock64() - t_staged));
            atomicAdd((unsigned long long*) &g_split_prof[38], 1ull);
            atomicAdd((unsigned long long*) &g_split_prof[39], (unsigned long long) nu);
        }
#endif
    }
}

#include "qkv_fused.cuh"
#include "pos_fused.cuh"

template <int S>
int launch_plan(const dsvt_set_attention_params* p, const int* idx, const float* mask, const int* set_num, int* plan,
                cudaStream_t st)
{
    const size_t stride = plan_words(p->max_set_num, S, p->max_pillars_num);
    const PlanView pv = plan_view(plan, p->max_set_num, p->max_pillars_num);
    // voxel -> (set, u) map: -1 for voxels that are in no set
    DSVT_CUDA(cudaMemset2DAsync(pv.vox_su, stride * sizeof(int), 0xFF, (size_t) p->max_pillars_num * sizeof(int), p->batch, st));
    count_launch();
    attn_plan_sets_kernel<S><<<dim3((p->max_set_num + 7) / 8, p->batch), 256, 0, st>>>(
        idx, mask, set_num, plan, stride, p->max_set_num, p->max_pillars_num, p->axis_id);
    DSVT_LAUNCH_CHECK();
    attn_plan_scan_kernel<<<p->batch, 1024, 0, st>>>(set_num, plan, stride, p->max_set_num, p->max_pillars_num);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

template <int S>
int launch_core(const dsvt_set_attention_params* p, const float* qbuf, const float* kvbuf, const int* plan,
                const float* mask, const int* set_num, float* o, cudaStream_t st)
{
    const size_t smem = (size_t) S * kKvTok * sizeof(float);
    {
        static PerDeviceOnce once;
        const int rc = once.raise_smem(attn_core_kernel<S>, (int) smem, true);
        if (rc != DSVT_OK) return rc;
    }
    const int cap = 12 * sm_count();        // idle CTAs (sets >= set_num) cost a launch slot each: do not start thousands
    const int grid = p->max_set_num < cap ? p->max_set_num : cap;
    attn_core_kernel<S><<<dim3(grid, p->batch), S / 2 * kH, smem, st>>>(
        qbuf, kvbuf, plan, plan_words(p->max_set_num, S, p->max_pillars_num), mask, set_num, o, p->max_set_num,
        p->max_pillars_num);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

// split-image header kept in front of the device blob
struct SplitBlobHeader {
    float out_mul[kRoles];
};

}  // namespace

// One GEMM launch over `n_roles` roles (<= 3): every role is [rows, 192 k] x [192 k, 192]^T with its own operands / epilogue.
static int launch_gemm(const GemmRoles& roles, int n_roles, const int* rows_dev, int rows_host, int max_rows, int max_sets,
                       int zero_tails, int batch, bool split, cudaStream_t st)
{
    DSVT_RAISE_SMEM(proj_tile_kernel<true>, TLay<true>::total);
    DSVT_RAISE_SMEM(proj_tile_kernel<false>, TLay<false>::total);
    const dim3 grid((max_rows + kBM - 1) / kBM, n_roles, batch);
    const bool ln = roles.r[0].n_ln > 0;       // the LayerNorm epilogue stages the finished tile: 98 KB instead of 80 KB
    if (split) proj_tile_kernel<true><<<grid, kTThreads, ln ? TLay<true>::total : TLay<true>::plain, st>>>(roles, rows_dev, rows_host, max_rows, max_sets, zero_tails);
    else proj_tile_kernel<false><<<grid, kTThreads, ln ? TLay<false>::total : TLay<false>::plain, st>>>(roles, rows_dev, rows_host, max_rows, max_sets, zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

// One 192 x 192 weight block W[n0 .. n0+192][k0 .. k0+192] (row stride ldw) as the kernel's image: six K chunks of
// [hi 192x32 | lo 192x32] FP16, pre-scaled by 2^sh.
static void build_block_image(const float* W, int ldw, int n0, int k0, int sh, uint8_t* img) {
    const float ws = ldexpf(1.0f, sh);
    for (int kc = 0; kc < kNumK; ++kc) {
        uint8_t* hi_img = img + (size_t) kc * kWChunkBytes;
        uint8_t* lo_img = hi_img + kBTerm;
        for (int n = 0; n < kBN; ++n)
            for (int c16 = 0; c16 < kBK / 8; ++c16)
                for (int e = 0; e < 8; ++e) {
                    const float w = W[(size_t) (n0 + n) * ldw + k0 + kc * kBK + c16 * 8 + e] * ws;
                    const __half h = __float2half_rn(w);
                    const __half l = __float2half_rn(w - __half2float(h));
                    const uint16_t hb = __half_as_ushort(h), lb = __half_as_ushort(l);
                    const size_t off = (size_t) c16 * (kBN * 16) + (size_t) n * 16 + e * 2;
                    memcpy(hi_img + off, &hb, 2);
                    memcpy(lo_img + off, &lb, 2);
                }
    }
}
static int scale_shift(const float* W, size_t n) {
    float maxabs = 0.f;
    for (size_t t = 0; t < n; ++t) maxabs = fmaxf(maxabs, fabsf(W[t]));
    if (!(maxabs > 0.f) || !std::isfinite(maxabs)) return 0;
    int e;
    frexpf(maxabs, &e);              // maxabs = m * 2^e, m in [0.5, 1)
    int sh = 14 - e;                 // maxabs * 2^sh in [2^13, 2^14)
    return sh > 60 ? 60 : (sh < -60 ? -60 : sh);
}

// ---- dense linear layer on the same kernel: y[M,N] = act(x[M,K] W[N,K]^T + b), K and N multiples of 192 -------------
// Device blob: [N/192][K/192] block images (147456 B each), then bias [N] f32, then 192 zero floats (the bias of K blocks > 0).  One launch per 192-wide K block (the
// CTA's resident weight image is one block); the second and later K blocks add to the rows written by the first.
void* linear_split_prepare(int N, int K, const float* W, const float* b, float* out_mul) {
    const int nb = N / kBN, kb = K / kC;
    const size_t img_bytes = (size_t) nb * kb * kWRoleBytes;
    std::vector<uint8_t> host(img_bytes + ((size_t) N + kBN) * sizeof(float), 0);     // images | bias [N] | 192 zeros
    const int sh = scale_shift(W, (size_t) N * K);
    *out_mul = ldexpf(1.0f, -sh);
    for (int i = 0; i < nb; ++i)
        for (int j = 0; j < kb; ++j)
            build_block_image(W, K, i * kBN, j * kC, sh, host.data() + ((size_t) i * kb + j) * kWRoleBytes);
    float* bias = reinterpret_cast<float*>(host.data() + img_bytes);
    for (int n = 0; n < N; ++n) bias[n] = b ? b[n] : 0.f;
    void* dev = nullptr;
    if (cudaMalloc(&dev, host.size()) != cudaSuccess) return nullptr;
    if (cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dev); return nullptr; }
    return dev;
}

int linear_split_launch(const void* blob, int N, int K, float out_mul, bool split, int act, const float* x,
                        const float* x_hi, int k_split, const int* rows_dev, int rows_host, int max_rows, float* y,
                        int zero_tails, cudaStream_t st)
{
    // x_hi != nullptr: the input row is the concatenation [x (k_split columns, dense) | x_hi (K - k_split columns, dense)];
    // only single-K-block layers (K == 192) take it
    const int nb = N / kBN, kb = K / kC;
    const uint8_t* img = static_cast<const uint8_t*>(blob);
    const float* bias = reinterpret_cast<const float*>(img + (size_t) nb * kb * kWRoleBytes);
    const float* zero_bias = bias + N;           // K blocks after the first add no bias: 192 zeros kept behind the bias
    for (int j = 0; j < kb; ++j)
        for (int i0 = 0; i0 < nb; i0 += 3) {            // up to three 192-column output blocks (roles) per launch
            const int n_roles = nb - i0 < 3 ? nb - i0 : 3;
            GemmRoles roles;
            for (int r = 0; r < 3; ++r) {
                const int i = i0 + (r < n_roles ? r : 0);
                GemmRole& g = roles.r[r];
                g.kchunks = kNumK; g.n_ln = 0; g.ln_eps = 0.f; g.ln_res0_written_here = 0; g.gen_x = nullptr; g.gen_blob = nullptr;
                g.a0 = x + (size_t) j * kC; g.a1 = nullptr; g.lda = x_hi ? k_split : K;
                g.a0b = x_hi; g.ksplit = x_hi ? k_split : 0; g.ldb = x_hi ? K - k_split : 0;
                g.add_src = nullptr; g.ld_add = 0; g.cover = nullptr; g.cover_stride = 0;
                g.wimg = img + ((size_t) i * kb + j) * kWRoleBytes;
                g.bias = j == 0 ? bias + i * kBN : zero_bias;
                g.out = y; g.ld_out = N; g.col0 = i * kBN;
                g.out_mul = out_mul; g.post_mul = 1.0f;
                g.plan = nullptr; g.plan_stride = 0; g.pad_hi = 0;
                g.accumulate = j > 0;
                g.act = j == kb - 1 ? act : 0;
            }
            const int zt = (zero_tails && j == 0) ? 1 : 0;
            const int rc = launch_gemm(roles, n_roles, rows_dev, rows_host, max_rows, 1, zt, 1, split, st);
            if (rc != DSVT_OK) return rc;
        }
    return DSVT_OK;
}

// Split-K form of a [*, K] -> [*, 192] layer (K = 192 * kb, kb <= 3): ONE launch whose roles are the K blocks; role j writes
// the partial product of K block j to y_parts[j] ([max_rows, 192] each; bias and the optional residual rows `add` go into
// part 0).  The consumer sums the parts (the LayerNorm behind the FFN takes part 1 as its residual input), so the
// read-modify-write of the accumulating two-launch form and one pipeline fill / drain disappear.
int linear_split_k_launch(const void* blob, int N, int K, float out_mul, bool split, const float* x, const float* add,
                          const int* rows_dev, int max_rows, float* y_parts, cudaStream_t st)
{
    const int kb = K / kC;
    const uint8_t* img = static_cast<const uint8_t*>(blob);
    const float* bias = reinterpret_cast<const float*>(img + (size_t) kb * kWRoleBytes);
    const float* zero_bias = bias + N;           // 192 zeros kept behind the bias in the weight blob
    GemmRoles roles;
    for (int r = 0; r < 3; ++r) {
        const int j = r < kb ? r : 0;
        GemmRole& g = roles.r[r];
        g.kchunks = kNumK; g.n_ln = 0; g.ln_eps = 0.f; g.ln_res0_written_here = 0; g.gen_x = nullptr; g.gen_blob = nullptr;
        g.a0 = x + (size_t) j * kC; g.a1 = nullptr; g.lda = K;
        g.a0b = nullptr; g.ksplit = 0; g.ldb = 0; g.cover = nullptr; g.cover_stride = 0;
        g.wimg = img + (size_t) j * kWRoleBytes;
        g.bias = j == 0 ? bias : zero_bias;
        g.out = y_parts + (size_t) j * max_rows * N; g.ld_out = N; g.col0 = 0;
        g.out_mul = out_mul; g.post_mul = 1.0f;
        g.plan = nullptr; g.plan_stride = 0; g.pad_hi = 0;
        g.accumulate = 0; g.act = 0;
        g.add_src = j == 0 ? add : nullptr; g.ld_add = N;
    }
    return launch_gemm(roles, kb, rows_dev, 0, max_rows, 1, 0, 1, split, st);
}

// Position-embedding MLP in one kernel: y = relu((x2 W1^T) * scale + shift) W2^T + b2 with x2 [rows, 2]; the first layer is
// evaluated by the GEMM's producers (GemmRole::gen_x), its [rows, 192] output never reaches memory (tile kernel only).
int linear_gen_launch(const void* blob, float out_mul, bool split, const float* x2, const float* small_blob, const int* rows_dev,
                      int max_rows, float* y, int zero_tails, cudaStream_t st)
{
    const uint8_t* img = static_cast<const uint8_t*>(blob);
    GemmRoles roles;
    GemmRole& g = roles.r[0];
    g.a0 = x2; g.a1 = nullptr; g.lda = kC;          // a0 is not read in this mode
    g.a0b = nullptr; g.ksplit = 0; g.ldb = 0; g.cover = nullptr; g.cover_stride = 0;
    g.wimg = img;
    g.bias = reinterpret_cast<const float*>(img + (size_t) kWRoleBytes);
    g.out = y; g.ld_out = kC; g.col0 = 0;
    g.out_mul = out_mul; g.post_mul = 1.0f;
    g.plan = nullptr; g.plan_stride = 0; g.pad_hi = 0;
    g.accumulate = 0; g.act = 0; g.add_src = nullptr; g.ld_add = 0;
    g.kchunks = kNumK; g.n_ln = 0; g.ln_eps = 0.f; g.ln_res0_written_here = 0;
    for (int s = 0; s < 3; ++s) { g.ln_res[s] = nullptr; g.ln_gamma[s] = nullptr; g.ln_beta[s] = nullptr; }
    g.gen_x = x2; g.gen_blob = small_blob;
    roles.r[1] = roles.r[2] = g;
    return launch_gemm(roles, 1, rows_dev, 0, max_rows, 1, zero_tails, 1, split, st);
}

// n (<= 8) position-embedding MLPs in ONE launch (role = MLP): the eight MLPs of a frame depend on the window coordinates only,
// and eight launches of 241 CTAs each leave a fifth of the 296 CTA slots of every wave empty.
int linear_gen_batch_launch(int n, const void* const* blobs, const float* out_muls, bool split, const float* const* x2s,
                            const float* const* small_blobs, const int* rows_dev, int max_rows, float* const* ys, int zero_tails,
                            cudaStream_t st)
{
#ifndef DSVT_POS_TILE_KERNEL
    if (split) {                                   // FP32 configuration: one tile-wide CTA, the MLPs as roles in sequence (pos_fused.cuh)
        PosArgs a;
        a.n = n;
        for (int r = 0; r < n; ++r) {
            const uint8_t* img = static_cast<const uint8_t*>(blobs[r]);
            a.r[r] = PosRole{x2s[r], small_blobs[r], img, reinterpret_cast<const float*>(img + (size_t) kWRoleBytes), ys[r], out_muls[r]};
        }
        DSVT_RAISE_SMEM(pos_fused_kernel, kPSmem);
        pos_fused_kernel<<<(max_rows + kBM - 1) / kBM, kQThreads, kPSmem, st>>>(a, rows_dev, max_rows, zero_tails);
        DSVT_LAUNCH_CHECK();
        return DSVT_OK;
    }
#endif
    GemmRoles roles;
    for (int r = 0; r < n; ++r) {
        const uint8_t* img = static_cast<const uint8_t*>(blobs[r]);
        GemmRole& g = roles.r[r];
        g.a0 = x2s[r]; g.a1 = nullptr; g.lda = kC;          // a0 is not read in this mode
        g.a0b = nullptr; g.ksplit = 0; g.ldb = 0; g.cover = nullptr; g.cover_stride = 0;
        g.wimg = img;
        g.bias = reinterpret_cast<const float*>(img + (size_t) kWRoleBytes);
        g.out = ys[r]; g.ld_out = kC; g.col0 = 0;
        g.out_mul = out_muls[r]; g.post_mul = 1.0f;
        g.plan = nullptr; g.plan_stride = 0; g.pad_hi = 0;
        g.accumulate = 0; g.act = 0; g.add_src = nullptr; g.ld_add = 0;
        g.kchunks = kNumK; g.n_ln = 0; g.ln_eps = 0.f; g.ln_res0_written_here = 0;
        for (int s = 0; s < 3; ++s) { g.ln_res[s] = nullptr; g.ln_gamma[s] = nullptr; g.ln_beta[s] = nullptr; }
        g.gen_x = x2s[r]; g.gen_blob = small_blobs[r];
    }
    return launch_gemm(roles, n, rows_dev, 0, max_rows, 1, zero_tails, 1, split, st);
}

// [*, K] -> [*, 192] layer (K = 192 or 384) followed by a chain of up to three (residual add + LayerNorm) stages, in ONE
// kernel: the linear's rows go through the LayerNorms in the epilogue and are stored once (tile kernel only).
int linear_ln_launch(const void* blob, int K, float out_mul, bool split, const float* x, const int* rows_dev, int max_rows,
                     int n_ln, const float* const* res, const float* const* gamma, const float* const* beta, float eps,
                     float* y, int zero_tails, cudaStream_t st)
{
    const int kb = K / kC;
    const uint8_t* img = static_cast<const uint8_t*>(blob);
    GemmRoles roles;
    GemmRole& g = roles.r[0];
    g.a0 = x; g.a1 = nullptr; g.lda = K;
    g.a0b = nullptr; g.ksplit = 0; g.ldb = 0; g.cover = nullptr; g.cover_stride = 0;
    g.wimg = img;                                    // K block j's image follows block j-1's: chunks 0 .. 6 kb - 1 are consecutive
    g.bias = reinterpret_cast<const float*>(img + (size_t) kb * kWRoleBytes);
    g.out = y; g.ld_out = kC; g.col0 = 0;
    g.out_mul = out_mul; g.post_mul = 1.0f;
    g.plan = nullptr; g.plan_stride = 0; g.pad_hi = 0;
    g.accumulate = 0; g.act = 0; g.add_src = nullptr; g.ld_add = 0;
    g.kchunks = kb * kNumK; g.n_ln = n_ln; g.ln_eps = eps; g.ln_res0_written_here = 0; g.gen_x = nullptr; g.gen_blob = nullptr;
    for (int s = 0; s < 3; ++s) {
        g.ln_res[s] = s < n_ln ? res[s] : nullptr;
        g.ln_gamma[s] = s < n_ln ? gamma[s] : nullptr;
        g.ln_beta[s] = s < n_ln ? beta[s] : nullptr;
    }
    roles.r[1] = roles.r[2] = g;
    return launch_gemm(roles, 1, rows_dev, 0, max_rows, 1, zero_tails, 1, split, st);
}

// ---- fused FFN (ffn_fused.cuh) ------------------------------------------------------------------------------------
// W1 [384, 192] as six pieces of 64 output columns: piece p = six K chunks of [hi 64x32 | lo 64x32] FP16 (8 KB each),
// pre-scaled like the block images of linear_split_prepare (same shift, so out_mul is the layer's).
void* ffn_w1_pieces_prepare(const float* W) {
    const int sh = scale_shift(W, (size_t) 2 * kC * kC);
    const float ws = ldexpf(1.0f, sh);
    std::vector<uint8_t> host((size_t) kFPieces * kFW1Piece, 0);
    for (int p = 0; p < kFPieces; ++p)
        for (int kc = 0; kc < kNumK; ++kc) {
            uint8_t* hi_img = host.data() + (size_t) p * kFW1Piece + (size_t) kc * kFW1Chunk;
            uint8_t* lo_img = hi_img + kFW1Chunk / 2;
            for (int n = 0; n < kFP; ++n)
                for (int c16 = 0; c16 < kBK / 8; ++c16)
                    for (int e = 0; e < 8; ++e) {
                        const float w = W[(size_t) (p * kFP + n) * kC + kc * kBK + c16 * 8 + e] * ws;
                        const __half h = __float2half_rn(w);
                        const __half l = __float2half_rn(w - __half2float(h));
                        const uint16_t hb = __half_as_ushort(h), lb = __half_as_ushort(l);
                        const size_t off = (size_t) c16 * (kFP * 16) + (size_t) n * 16 + e * 2;
                        memcpy(hi_img + off, &hb, 2);
                        memcpy(lo_img + off, &lb, 2);
                    }
        }
    void* dev = nullptr;
    if (cudaMalloc(&dev, host.size()) != cudaSuccess) return nullptr;
    if (cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dev); return nullptr; }
    return dev;
}

// blob1 / pieces1: the 192 -> 384 layer (its bias sits behind the two block images), blob2: the 384 -> 192 layer
struct FfnOutProj {            // the attention tail folded in front of the FFN (attn_ffn_fused_launch); all null / 0: plain FFN
    const float* o; const uint8_t* wo_img; const float* bias_o; float out_mul_o; const int* cover;
    const float* res1; const float* gamma1; const float* beta1; float eps1; float* src_out;
};
static int ffn_fused_launch_impl(const void* blob1, const void* pieces1, float out_mul1, const void* blob2, float out_mul2,
                                 const float* x, const FfnOutProj* op, const int* rows_dev, int max_rows, int n_ln,
                                 const float* const* res, const float* const* gamma, const float* const* beta, float eps, float* y,
                                 int zero_tails, cudaStream_t st)
{
    const uint8_t* img2 = static_cast<const uint8_t*>(blob2);
    FfnArgs a;
    GemmRole& g = a.g;
    g.a0 = x; g.a1 = nullptr; g.lda = kC;
    g.a0b = nullptr; g.ksplit = 0; g.ldb = 0; g.cover = nullptr; g.cover_stride = 0;
    g.wimg = img2;
    g.bias = reinterpret_cast<const float*>(img2 + (size_t) 2 * kWRoleBytes);
    g.out = y; g.ld_out = kC; g.col0 = 0;
    g.out_mul = out_mul2; g.post_mul = 1.0f;
    g.plan = nullptr; g.plan_stride = 0; g.pad_hi = 0;
    g.accumulate = 0; g.act = 0; g.add_src = nullptr; g.ld_add = 0;
    g.kchunks = 2 * kNumK; g.n_ln = n_ln; g.ln_eps = eps; g.ln_res0_written_here = op ? 1 : 0; g.gen_x = nullptr; g.gen_blob = nullptr;
    for (int s = 0; s < 3; ++s) {
        g.ln_res[s] = s < n_ln ? res[s] : nullptr;
        g.ln_gamma[s] = s < n_ln ? gamma[s] : nullptr;
        g.ln_beta[s] = s < n_ln ? beta[s] : nullptr;
    }
    a.w1_img = static_cast<const uint8_t*>(pieces1);
    a.bias1 = reinterpret_cast<const float*>(static_cast<const uint8_t*>(blob1) + (size_t) 2 * kWRoleBytes);
    a.out_mul1 = out_mul1;
    a.o = nullptr; a.wo_img = nullptr; a.bias_o = nullptr; a.out_mul_o = 1.0f; a.cover = nullptr;
    a.res1 = nullptr; a.gamma1 = nullptr; a.beta1 = nullptr; a.eps1 = 0.f; a.src_out = nullptr;
    const dim3 grid((max_rows + kBM - 1) / kBM, 1, 1);
    if (op) {
        a.o = op->o; a.wo_img = op->wo_img; a.bias_o = op->bias_o; a.out_mul_o = op->out_mul_o; a.cover = op->cover;
        a.res1 = op->res1; a.gamma1 = op->gamma1; a.beta1 = op->beta1; a.eps1 = op->eps1; a.src_out = op->src_out;
        g.ln_res[0] = op->src_out;                  // norm2's residual = the src rows this kernel writes
        DSVT_RAISE_SMEM(ffn_fused_kernel<true>, kFSmemTotal);
        ffn_fused_kernel<true><<<grid, kFThreads, kFSmemTotal, st>>>(a, rows_dev, max_rows, zero_tails);
    } else {
        DSVT_RAISE_SMEM(ffn_fused_kernel<false>, kFSmemTotal);
        ffn_fused_kernel<false><<<grid, kFThreads, kFSmemTotal, st>>>(a, rows_dev, max_rows, zero_tails);
    }
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}
int ffn_fused_launch(const void* blob1, const void* pieces1, float out_mul1, const void* blob2, float out_mul2, const float* x,
                     const int* rows_dev, int max_rows, int n_ln, const float* const* res, const float* const* gamma,
                     const float* const* beta, float eps, float* y, int zero_tails, cudaStream_t st)
{
    return ffn_fused_launch_impl(blob1, pieces1, out_mul1, blob2, out_mul2, x, nullptr, rows_dev, max_rows, n_ln, res, gamma, beta,
                                 eps, y, zero_tails, st);
}


// ---- fused VFE (vfe_fused.cuh) ------------------------------------------------------------------------------------
size_t vfe_fused_workspace(int max_points, int npv) {
    const int win = ((kBM - (npv - 1)) / 16) * 16;
    return ((size_t) (max_points + win - 1) / win + 4) * sizeof(int);
}
int vfe_fused_launch(const float* pfn0_blob, const void* blob1, float out_mul1, const float* point_features, const int* piv,
                     const int* voxel_num, const int* point_num, int max_points, int max_pillars, int npv, float* max_voxel,
                     void* workspace, int zero_tails, cudaStream_t st)
{
    int win = ((kBM - (npv - 1)) / 16) * 16;            // rows of a tile <= win + npv - 1 <= 128
    win = win < kVfeMaxPillars ? win : kVfeMaxPillars;   // pillars of a tile <= win
    const uint8_t* img1 = static_cast<const uint8_t*>(blob1);
    VfeArgs a;
    a.x = point_features; a.piv = piv; a.pfn0 = pfn0_blob;
    a.w1_img = img1; a.bias1 = reinterpret_cast<const float*>(img1 + (size_t) kWRoleBytes); a.out_mul1 = out_mul1;
    a.out = max_voxel; a.tiles = static_cast<const int*>(workspace); a.npv = npv; a.win = win;
    const int pgrid = (max_pillars + 255) / 256 < 4 * sm_count() ? (max_pillars + 255) / 256 : 4 * sm_count();
    vfe_plan_kernel<<<pgrid, 256, 0, st>>>(piv, voxel_num, static_cast<int*>(workspace), reinterpret_cast<float4*>(max_voxel),
                                           max_pillars, npv, win, zero_tails);
    DSVT_LAUNCH_CHECK();
    DSVT_RAISE_SMEM(vfe_fused_kernel, kVfeSmem);
    vfe_fused_kernel<<<(max_points + win - 1) / win, kVfeThreads, kVfeSmem, st>>>(a, voxel_num, point_num, max_points, max_pillars);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

// Device blob: [kRoles][kNumK][hi 12288 | lo 12288] weight images, then bias [kRoles][192] f32.
// out_mul[] (host) receives the per-role power-of-two that undoes the weight pre-scaling.
void* attention_split_prepare(const float* w_in, const float* b_in, const float* w_out, const float* b_out,
                              float* out_mul)
{
    const size_t img_bytes = (size_t) kRoles * kWRoleBytes;
    std::vector<uint8_t> host(img_bytes + (size_t) kRoles * kC * sizeof(float));
    for (int role = 0; role < kRoles; ++role) {
        const float* W = role < 3 ? w_in + (size_t) role * kC * kC : w_out;       // [192 out][192 in]
        float maxabs = 0.f;
        for (int t = 0; t < kC * kC; ++t) maxabs = fmaxf(maxabs, fabsf(W[t]));
        int sh = 0;
        if (maxabs > 0.f && std::isfinite(maxabs)) {
            int e;
            frexpf(maxabs, &e);              // maxabs = m * 2^e, m in [0.5, 1)
            sh = 14 - e;                     // maxabs * 2^sh in [2^13, 2^14)
            if (sh > 60) sh = 60;
            if (sh < -60) sh = -60;
        }
        const float ws = ldexpf(1.0f, sh);
        out_mul[role] = ldexpf(1.0f, -sh);
        for (int kc = 0; kc < kNumK; ++kc) {
            uint8_t* hi_img = host.data() + (size_t) role * kWRoleBytes + (size_t) kc * kWChunkBytes;
            uint8_t* lo_img = hi_img + kBTerm;
            for (int n = 0; n < kBN; ++n)
                for (int c16 = 0; c16 < kBK / 8; ++c16)
                    for (int e = 0; e < 8; ++e) {
                        const float w = W[(size_t) n * kC + kc * kBK + c16 * 8 + e] * ws;
                        const __half h = __float2half_rn(w);
                        const __half l = __float2half_rn(w - __half2float(h));
                        const uint16_t hb = __half_as_ushort(h), lb = __half_as_ushort(l);
                        const size_t off = (size_t) c16 * (kBN * 16) + (size_t) n * 16 + e * 2;
                        memcpy(hi_img + off, &hb, 2);
                        memcpy(lo_img + off, &lb, 2);
                    }
        }
        float* bias = reinterpret_cast<float*>(host.data() + img_bytes) + role * kC;
        const float* bsrc = role < 3 ? b_in + role * kC : b_out;
        for (int n = 0; n < kC; ++n) bias[n] = bsrc[n];
    }
    void* dev = nullptr;
    if (cudaMalloc(&dev, host.size()) != cudaSuccess) return nullptr;
    if (cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(dev);
        return nullptr;
    }
    return dev;
}

static size_t plan_bytes_of(const dsvt_set_attention_params* p) {
    return align_up((size_t) p->batch * plan_words(p->max_set_num, p->voxel_num_set, p->max_pillars_num) * sizeof(int), kWsAlign);
}
size_t attention_split_plan_bytes(const dsvt_set_attention_params* p) { return plan_bytes_of(p); }

size_t attention_split_workspace(const dsvt_set_attention_params* p) {
    // q [B, max_pillars, 192] | kv [B, max_pillars, 392] | o [B, max_pillars, 192] (f32) | a plan (used when the caller
    // does not pass one)
    const size_t rows = (size_t) p->batch * p->max_pillars_num;
    return align_up(rows * kC * sizeof(float), kWsAlign) + align_up(rows * kKvTok * sizeof(float), kWsAlign) +
           align_up(rows * kC * sizeof(float), kWsAlign) + plan_bytes_of(p);
}

static int split_check(const dsvt_set_attention_params* p) {
    if (p->channel_num != kC || p->num_heads != kH ||
        (p->voxel_num_set != 24 && p->voxel_num_set != 36 && p->voxel_num_set != 48)) {
        set_last_error("set attention (GEMM pipeline): only C=192, heads=8, set in {24,36,48} is built");
        return DSVT_ERR_UNSUPPORTED;
    }
    return DSVT_OK;
}

int attention_split_plan(const dsvt_set_attention_params* p, const int* idx, const float* mask, const int* set_num,
                         void* plan, size_t plan_bytes, cudaStream_t st)
{
    int rc = split_check(p);
    if (rc != DSVT_OK) return rc;
    if (!plan || ((uintptr_t) plan & 255) || plan_bytes < plan_bytes_of(p)) {
        set_last_error("set attention plan: a 256-byte aligned buffer of %zu bytes is required (dsvt_set_attention_plan_size)",
                       plan_bytes_of(p));
        return DSVT_ERR_INVALID_ARGUMENT;
    }
    switch (p->voxel_num_set) {
        case 24: return launch_plan<24>(p, idx, mask, set_num, static_cast<int*>(plan), st);
        case 36: return launch_plan<36>(p, idx, mask, set_num, static_cast<int*>(plan), st);
        default: return launch_plan<48>(p, idx, mask, set_num, static_cast<int*>(plan), st);
    }
}

int set_attention_split_fused(const dsvt_set_attention_params* p, const void* split_blob, const float* out_mul,
                              bool split, const float* x, const float* pos, const int* idx, const float* mask,
                              const int* set_num, const int* voxel_num, float* out, const void* plan_in,
                              void* workspace, size_t workspace_bytes, cudaStream_t st, const AttnNorm* norm, int stages,
                              const AttnPosTable* pos_table)
{
    int rc = split_check(p);
    if (pos_table && (stages & 1) && (!split || !pos_table->cell || pos_table->win_x < 1)) {
        set_last_error("set attention: the position-embedding table form is built for DSVT_ATTN_FP32_TC (cell map and win_x required)");
        return DSVT_ERR_UNSUPPORTED;
    }
    if (!split_blob) {
        set_last_error("set attention (GEMM pipeline): weights were not prepared");
        return DSVT_ERR_INVALID_ARGUMENT;
    }
    if (((uintptr_t) x | (uintptr_t) pos | (uintptr_t) workspace) & 31) {
        set_last_error("set attention (GEMM pipeline): x, pos and workspace must be 32-byte aligned (256-bit loads)");
        return DSVT_ERR_INVALID_ARGUMENT;
    }
    if (!workspace || workspace_bytes < attention_split_workspace(p)) {
        set_last_error("set attention (GEMM pipeline): workspace of %zu bytes required (dsvt_set_attention_workspace_size)",
                       attention_split_workspace(p));
        return DSVT_ERR_INVALID_ARGUMENT;
    }
    const size_t rows = (size_t) p->batch * p->max_pillars_num;
    WsCarver ws(workspace);
    float* qbuf = ws.take<float>(rows * kC);
    float* kvbuf = ws.take<float>(rows * kKvTok);
    float* o = ws.take<float>(rows * kC);
    int* own_plan = ws.take<int>(plan_bytes_of(p) / sizeof(int));
    const uint8_t* img = static_cast<const uint8_t*>(split_blob);
    const float* bias = reinterpret_cast<const float*>(img + (size_t) kRoles * kWRoleBytes);

    const int* plan = static_cast<const int*>(plan_in);
    if (!plan) {                                   // stateless call: build the partition's plan first
        if ((rc = attention_split_plan(p, idx, mask, set_num, own_plan, plan_bytes_of(p), st)) != DSVT_OK) return rc;
        plan = own_plan;
    }
    const size_t plan_stride = plan_words(p->max_set_num, p->voxel_num_set, p->max_pillars_num);
    GemmRoles in_roles, out_roles;
    for (int r = 0; r < 3; ++r) {
        GemmRole& g = in_roles.r[r];
        g.kchunks = kNumK; g.n_ln = 0; g.ln_eps = 0.f; g.ln_res0_written_here = 0; g.gen_x = nullptr; g.gen_blob = nullptr;
        g.a0 = x;
        g.a1 = r < 2 ? pos : nullptr;
        g.wimg = img + (size_t) r * kWRoleBytes;
        g.bias = bias + r * kC;
        g.out = r == 0 ? qbuf : kvbuf;             // rows in token order (set-major): see attn_core_kernel
        g.ld_out = r == 0 ? kC : kKvTok;
        g.col0 = r == 2 ? kKvRow : 0;
        g.out_mul = out_mul[r];
        g.post_mul = r == 0 ? 1.0f / sqrtf((float) (kC / kH)) : 1.0f;
        g.plan = plan;
        g.plan_stride = plan_stride;
        g.pad_hi = r == 0 ? 0 : 4;
        g.lda = kC; g.accumulate = 0; g.act = 0;
        g.a0b = nullptr; g.ksplit = 0; g.ldb = 0; g.add_src = nullptr; g.ld_add = 0; g.cover = nullptr; g.cover_stride = 0;
    }
    {
        GemmRole& g = out_roles.r[0];
        g.kchunks = kNumK; g.n_ln = 0; g.ln_eps = 0.f; g.ln_res0_written_here = 0; g.gen_x = nullptr; g.gen_blob = nullptr;
        g.a0 = o; g.a1 = nullptr;
        g.wimg = img + (size_t) 3 * kWRoleBytes;
        g.bias = bias + 3 * kC;
        g.out = out; g.ld_out = kC; g.col0 = 0;
        g.out_mul = out_mul[3];
        g.post_mul = 1.0f;
        g.plan = nullptr; g.plan_stride = 0; g.pad_hi = 0;
        g.lda = kC; g.accumulate = 0; g.act = 0;
        g.a0b = nullptr; g.ksplit = 0; g.ldb = 0; g.add_src = nullptr; g.ld_add = 0;
        g.cover = plan_view(const_cast<int*>(plan), p->max_set_num, p->max_pillars_num).vox_su;   // voxels in no set -> 0
        g.cover_stride = plan_stride;
        if (norm) {                                // out = LayerNorm(attention + residual): norm1(y + x), src/dsvt-ai-trt.cpp:669-676
            g.n_ln = 1; g.ln_eps = norm->eps; g.ln_res0_written_here = 0;
            g.ln_res[0] = norm->residual; g.ln_gamma[0] = norm->gamma; g.ln_beta[0] = norm->beta;
            g.ln_res[1] = g.ln_res[2] = nullptr; g.ln_gamma[1] = g.ln_gamma[2] = nullptr; g.ln_beta[1] = g.ln_beta[2] = nullptr;
        }
        out_roles.r[1] = out_roles.r[2] = g;
    }
    // `stages`: bit 0 QKV projection GEMM, bit 1 per-set core, bit 2 out-projection GEMM (all three in normal operation; the
    // instrumented entry point launches them one at a time so that each can be bracketed by CUDA events)
    if (stages & 1) {
#ifndef DSVT_QKV_TILE_KERNEL
        if (split) {                               // FP32 configuration: one tile-wide CTA for the three roles (qkv_fused.cuh)
            QkvArgs qa;
            qa.x = x; qa.pos = pos; qa.wimg = img; qa.bias = bias;
            qa.pos_cell = pos_table ? pos_table->cell : nullptr; qa.win_x = pos_table ? pos_table->win_x : 0;
            for (int r = 0; r < 3; ++r) qa.out_mul[r] = out_mul[r];
            qa.q_post_mul = 1.0f / sqrtf((float) (kC / kH));
            qa.qbuf = qbuf; qa.kvbuf = kvbuf; qa.plan = plan; qa.plan_stride = plan_stride;
            DSVT_RAISE_SMEM(qkv_fused_kernel, kQSmem);
            qkv_fused_kernel<<<dim3((p->max_pillars_num + kBM - 1) / kBM, 1, p->batch), kQThreads, kQSmem, st>>>(
                qa, voxel_num, p->max_pillars_num, p->max_set_num);
            DSVT_LAUNCH_CHECK();
        } else
#endif
        if ((rc = launch_gemm(in_roles, 3, voxel_num, 0, p->max_pillars_num, p->max_set_num, 0, p->batch, split, st)) != DSVT_OK)
            return rc;
    }
    if (stages & 2) switch (p->voxel_num_set) {
        case 24: rc = launch_core<24>(p, qbuf, kvbuf, plan, mask, set_num, o, st); break;
        case 36: rc = launch_core<36>(p, qbuf, kvbuf, plan, mask, set_num, o, st); break;
        default: rc = launch_core<48>(p, qbuf, kvbuf, plan, mask, set_num, o, st); break;
    }
    if (rc != DSVT_OK) return rc;
    if ((stages & 4) &&
        (rc = launch_gemm(out_roles, 1, voxel_num, 0, p->max_pillars_num, p->max_set_num, p->zero_tails, p->batch, split, st)) != DSVT_OK)
        return rc;
    return DSVT_OK;
}

// The tail of an encoder layer in ONE kernel: out-projection of the attention (rows `o` of the per-set core, in the caller's
// set-attention workspace) + norm1(. + x) + the FFN + the LayerNorm chain behind it.  Batch 1.
int attn_ffn_fused_launch(const dsvt_set_attention_params* p, const void* attn_blob, const float* attn_out_mul, const void* plan,
                          void* workspace, size_t workspace_bytes, const float* x_res, const float* gamma1, const float* beta1,
                          float eps1, const void* blob1, const void* pieces1, float out_mul1, const void* blob2, float out_mul2,
                          const int* rows_dev, int n_ln, const float* const* res, const float* const* gamma, const float* const* beta,
                          float eps, float* src_out, float* y, int zero_tails, cudaStream_t st)
{
    int rc = split_check(p);
    if (rc != DSVT_OK) return rc;
    if (p->batch != 1 || !plan || !workspace || workspace_bytes < attention_split_workspace(p)) {
        set_last_error("attention tail + FFN: batch 1, a prebuilt plan and the set-attention workspace are required");
        return DSVT_ERR_INVALID_ARGUMENT;
    }
    const size_t rows = (size_t) p->max_pillars_num;
    WsCarver ws(workspace);
    ws.take<float>(rows * kC);                      // q
    ws.take<float>(rows * kKvTok);                  // k | v
    const float* o = ws.take<float>(rows * kC);
    const uint8_t* img = static_cast<const uint8_t*>(attn_blob);
    const float* bias = reinterpret_cast<const float*>(img + (size_t) kRoles * kWRoleBytes);
    FfnOutProj op;
    op.o = o; op.wo_img = img + (size_t) 3 * kWRoleBytes; op.bias_o = bias + 3 * kC; op.out_mul_o = attn_out_mul[3];
    op.cover = plan_view(const_cast<int*>(static_cast<const int*>(plan)), p->max_set_num, p->max_pillars_num).vox_su;
    op.res1 = x_res; op.gamma1 = gamma1; op.beta1 = beta1; op.eps1 = eps1; op.src_out = src_out;
    return ffn_fused_launch_impl(blob1, pieces1, out_mul1, blob2, out_mul2, nullptr, &op, rows_dev, p->max_pillars_num, n_ln, res,
                                 gamma, beta, eps, y, zero_tails, st);
}


// ---------------------------------------------------------------------------------------------------------------
// Plugin-shaped form (the drop-in for multHeadAttention() itself, src/dsvt-ai-trt.cpp:288-458: q, k, v arrive pre-gathered as
// [B, max_sets, S, 192]) on the same tensor-core pipeline.  Every slot of a valid set is a token here (the gather has already
// duplicated the padding slots; the key mask removes them), so the "plan" is the identity: token = row = set * S + slot.
namespace {
__global__ void __launch_bounds__(256)
attn_identity_plan_kernel(const int* __restrict__ set_num, int* __restrict__ plan, size_t plan_stride, int* __restrict__ rows_out,
                          int max_sets, int S)
{
    const int b = blockIdx.y, rows_cap = max_sets * S;
    const PlanView pv = plan_view(plan + (size_t) b * plan_stride, max_sets, rows_cap);
    int ns = set_num ? set_num[b] : max_sets;
    ns = ns < max_sets ? (ns < 0 ? 0 : ns) : max_sets;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { pv.hdr[0] = ns * S; rows_out[b] = ns * S; rows_out[gridDim.y + b] = ns; }     // rows | sets, per batch item
    if (i <= max_sets) pv.set_off[i] = (i < ns ? i : ns) * S;
    if (i < max_sets) { pv.nu[i] = i < ns ? S : 0; pv.order[i] = make_int4(i, (i < ns ? i : ns) * S, i < ns ? S : 0, 0); }
    for (int t = i; t < rows_cap; t += gridDim.x * blockDim.x) {
        const int set = t / S, u = t - set * S;
        pv.vox_su[t] = set < ns ? set * 64 + u : -1;
        pv.tok[t] = make_int2(t, u);
    }
}
}  // namespace

size_t attention_split_plugin_workspace(const dsvt_set_attention_params* p) {
    dsvt_set_attention_params q = *p;
    q.max_pillars_num = p->max_set_num * p->voxel_num_set;        // one row per set slot
    return attention_split_workspace(&q) + align_up(2 * (size_t) p->batch * sizeof(int), kWsAlign);   // + the device-side counts
}

int set_attention_split_plugin(const dsvt_set_attention_params* p, const void* split_blob, const float* out_mul, bool split,
                               const float* q, const float* k, const float* v, const float* mask, const int* set_num,
                               float* out, void* workspace, size_t workspace_bytes, cudaStream_t st)
{
    int rc = split_check(p);
    if (rc != DSVT_OK) return rc;
    if (!split_blob) { set_last_error("set attention (GEMM pipeline): weights were not prepared"); return DSVT_ERR_INVALID_ARGUMENT; }
    if (((uintptr_t) q | (uintptr_t) k | (uintptr_t) v | (uintptr_t) workspace) & 31) {
        set_last_error("set attention (GEMM pipeline): q, k, v and workspace must be 32-byte aligned (256-bit loads)");
        return DSVT_ERR_INVALID_ARGUMENT;
    }
    if (!workspace || workspace_bytes < attention_split_plugin_workspace(p)) {
        set_last_error("set attention (GEMM pipeline): workspace of %zu bytes required (dsvt_set_attention_workspace_size)",
                       attention_split_plugin_workspace(p));
        return DSVT_ERR_INVALID_ARGUMENT;
    }
    dsvt_set_attention_params pp = *p;
    const int rows_cap = p->max_set_num * p->voxel_num_set;
    pp.max_pillars_num = rows_cap;
    const size_t rows = (size_t) p->batch * rows_cap;
    WsCarver ws(workspace);
    float* qbuf = ws.take<float>(rows * kC);
    float* kvbuf = ws.take<float>(rows * kKvTok);
    float* o = ws.take<float>(rows * kC);
    int* plan = ws.take<int>(plan_bytes_of(&pp) / sizeof(int));
    int* rows_dev = ws.take<int>(2 * (size_t) p->batch);     // [B] valid rows (= sets * S) | [B] valid sets
    const size_t plan_stride = plan_words(p->max_set_num, p->voxel_num_set, rows_cap);
    const uint8_t* img = static_cast<const uint8_t*>(split_blob);
    const float* bias = reinterpret_cast<const float*>(img + (size_t) kRoles * kWRoleBytes);

    const int pgrid = (rows_cap + 255) / 256 < 1024 ? (rows_cap + 255) / 256 : 1024;
    attn_identity_plan_kernel<<<dim3(pgrid > (p->max_set_num + 256) / 256 ? pgrid : (p->max_set_num + 256) / 256, p->batch), 256, 0, st>>>(
        set_num, plan, plan_stride, rows_dev, p->max_set_num, p->voxel_num_set);
    DSVT_LAUNCH_CHECK();

    GemmRoles in_roles, out_roles;
    const float* srcs[3] = {q, k, v};
    for (int r = 0; r < 3; ++r) {
        GemmRole& g = in_roles.r[r];
        g.kchunks = kNumK; g.n_ln = 0; g.ln_eps = 0.f; g.ln_res0_written_here = 0; g.gen_x = nullptr; g.gen_blob = nullptr;
        for (int s2 = 0; s2 < 3; ++s2) { g.ln_res[s2] = nullptr; g.ln_gamma[s2] = nullptr; g.ln_beta[s2] = nullptr; }
        g.a0 = srcs[r]; g.a1 = nullptr;                      // the gather plugin has already formed q = k = x + pos, v = x
        g.wimg = img + (size_t) r * kWRoleBytes;
        g.bias = bias + r * kC;
        g.out = r == 0 ? qbuf : kvbuf;
        g.ld_out = r == 0 ? kC : kKvTok;
        g.col0 = r == 2 ? kKvRow : 0;
        g.out_mul = out_mul[r];
        g.post_mul = r == 0 ? 1.0f / sqrtf((float) (kC / kH)) : 1.0f;
        g.plan = plan; g.plan_stride = plan_stride;
        g.pad_hi = r == 0 ? 0 : 4;
        g.lda = kC; g.accumulate = 0; g.act = 0;
        g.a0b = nullptr; g.ksplit = 0; g.ldb = 0; g.add_src = nullptr; g.ld_add = 0; g.cover = nullptr; g.cover_stride = 0;
    }
    {
        GemmRole& g = out_roles.r[0];
        g.kchunks = kNumK; g.n_ln = 0; g.ln_eps = 0.f; g.ln_res0_written_here = 0; g.gen_x = nullptr; g.gen_blob = nullptr;
        for (int s2 = 0; s2 < 3; ++s2) { g.ln_res[s2] = nullptr; g.ln_gamma[s2] = nullptr; g.ln_beta[s2] = nullptr; }
        g.a0 = o; g.a1 = nullptr;
        g.wimg = img + (size_t) 3 * kWRoleBytes;
        g.bias = bias + 3 * kC;
        g.out = out; g.ld_out = kC; g.col0 = 0;
        g.out_mul = out_mul[3]; g.post_mul = 1.0f;
        g.plan = nullptr; g.plan_stride = 0; g.pad_hi = 0;
        g.lda = kC; g.accumulate = 0; g.act = 0;
        g.a0b = nullptr; g.ksplit = 0; g.ldb = 0; g.add_src = nullptr; g.ld_add = 0; g.cover = nullptr; g.cover_stride = 0;
        out_roles.r[1] = out_roles.r[2] = g;
    }
    const int* ns_dev = rows_dev + p->batch;                 // set_num == NULL: all max_set_num sets, as the reference graph does
    if ((rc = launch_gemm(in_roles, 3, rows_dev, 0, rows_cap, p->max_set_num, 0, p->batch, split, st)) != DSVT_OK) return rc;
    switch (p->voxel_num_set) {
        case 24: rc = launch_core<24>(&pp, qbuf, kvbuf, plan, mask, ns_dev, o, st); break;
        case 36: rc = launch_core<36>(&pp, qbuf, kvbuf, plan, mask, ns_dev, o, st); break;
        default: rc = launch_core<48>(&pp, qbuf, kvbuf, plan, mask, ns_dev, o, st); break;
    }
    if (rc != DSVT_OK) return rc;
    return launch_gemm(out_roles, 1, rows_dev, 0, rows_cap, p->max_set_num, p->zero_tails, p->batch, split, st);
}

}  // namespace dsvt

#ifdef DSVT_PROFILE
extern "C" int dsvt_debug_split_profile(long long* out64) {
    return cudaMemcpyFromSymbol(out64, dsvt::g_split_prof, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif
