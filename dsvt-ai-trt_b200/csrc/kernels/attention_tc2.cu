// a3 -- set attention on tcgen05, WARP-SPECIALISED pipeline (second generation of attention_tc.cu; same math,
// same operand layouts, same weight images).
//
// attention_tc.cu runs the per-head chain  proj-MMA -> epilogue -> S-MMA -> softmax -> PV-MMA  strictly in sequence:
// its phase profile (tools/tc_profile.py) shows ~5.2k cycles per head of which only ~2k are tensor-pipe issue time.
// Here three engines run concurrently on different heads, synchronised with mbarriers only:
//   warp 8 (one lane)  ISSUER : all tcgen05.mma + the cp.async.bulk weight stream
//   warps 0-3          E group: projection epilogue (TMEM -> +bias -> FP16 Q/K/V operand tiles) and the per-head
//                               O epilogue (TMEM -> FP16 O operand tile)
//   warps 4-7          S group: softmax (one thread per token row), P written back to tensor memory
// Double-buffered per head parity b = h & 1: PROJ accumulators, S/P, O accumulators (TMEM) and the Q/K/V tiles (smem).
// TMEM map: PROJ[b] = 80 b .. | SP[b] = 160 + 112 b .. (P aliases the first 56 columns of its S) | O[b] = 384 + 32 b;
// the out-projection accumulator reuses [0,192).
#include "attention_common.cuh"
#include "tc_common.cuh"
#include <cstring>
#include <vector>
#include <cuda_fp16.h>

namespace dsvt {
namespace {

using namespace tc;

constexpr int kC = 192, kH = 8, kD = 24, kS = 36;
constexpr int kSetsPerTile = 3;
constexpr int kRows = 128;
constexpr int kValidRows = kSetsPerTile * kS;   // 108
constexpr int kKeys = 112;
constexpr int kWorkers = 256;                   // E group + S group
constexpr int kThreads = kWorkers + 32;         // + issuer warp

constexpr uint32_t C_PROJ = 0, C_SP = 160, C_O = 384, C_OUT = 0;
constexpr uint32_t kProjCols = 80, kSpCols = 112, kOCols = 32;

constexpr int kChunkStride = kRows * 16;
constexpr int SM_AQK = 0;                                // 49152 (later: Wout rows 0..95, then FP32 output tile)
constexpr int SM_AV = SM_AQK + 24 * kChunkStride;        // 49152 (later: Wout rows 96..191, output tile cont.)
constexpr int SM_O = SM_AV + 24 * kChunkStride;          // 49152  O operand tile of the out-projection
constexpr int SM_W = SM_O + 24 * kChunkStride;           // 30720  per-head weights
constexpr int SM_QKV = SM_W + 30720;                     // 2 x (Q 8192 | K 8192 | V 8192)
constexpr int kQkvBytes = 3 * 8192;
constexpr int SM_TOTAL = SM_QKV + 2 * kQkvBytes;         // 227328

constexpr int kWqkBytes = 48 * kC * 2, kWvBytes = 32 * kC * 2, kWHeadBytes = kWqkBytes + kWvBytes;
constexpr int kWoutHalfBytes = 96 * kC * 2;
constexpr int kWImgBytes = kH * kWHeadBytes + 2 * kWoutHalfBytes;

struct TcBlobView2 {
    const uint8_t* w_img;
    const float* bias;     // b_q (pre-scaled) | b_k | b_v | b_out, 4 x 192
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 cvt8(const uint32_t* r, const float* b) {
    return make_uint4(pack_h2(__uint_as_float(r[0]) + b[0], __uint_as_float(r[1]) + b[1]),
                      pack_h2(__uint_as_float(r[2]) + b[2], __uint_as_float(r[3]) + b[3]),
                      pack_h2(__uint_as_float(r[4]) + b[4], __uint_as_float(r[5]) + b[5]),
                      pack_h2(__uint_as_float(r[6]) + b[6], __uint_as_float(r[7]) + b[7]));
}
__device__ __forceinline__ uint4 cvt8_nobias(const uint32_t* r) {
    return make_uint4(pack_h2(__uint_as_float(r[0]), __uint_as_float(r[1])), pack_h2(__uint_as_float(r[2]), __uint_as_float(r[3])),
                      pack_h2(__uint_as_float(r[4]), __uint_as_float(r[5])), pack_h2(__uint_as_float(r[6]), __uint_as_float(r[7])));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

#ifdef DSVT_PROFILE      // phase stamps of one CTA (tools/tc_profile.py): never in the product build
__device__ long long g_tc2_prof[64];
#define P2(i) do { if (blockIdx.x == 0 && tile == (int) blockIdx.x) g_tc2_prof[i] = clock64(); } while (0)
#else
#define P2(i) do { } while (0)
#endif

struct Barriers {
    uint64_t w_full, wout_full, a_full, so_full, out_full;
    uint64_t proj_full[2], proj_free[2], qkv_full[2], s_full[2], p_full[2], o_full[2], o_free[2];
};

__global__ void __launch_bounds__(kThreads, 1)
set_attention_tc2_kernel(const float* __restrict__ x, const float* __restrict__ pos, const int* __restrict__ idx,
                         const float* __restrict__ mask, const int* __restrict__ set_num,
                         const int* __restrict__ voxel_num, float* __restrict__ out, TcBlobView2 wb,
                         int max_sets, int max_pillars, int axis, int zero_tails)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) Barriers bars;
    __shared__ uint32_t tmem_slot;
    __shared__ int s_rows[kRows];
    __shared__ __align__(16) float s_bias[4 * kC];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    int ns = set_num[b];
    ns = ns < max_sets ? ns : max_sets;
    const int n_tiles = (ns + kSetsPerTile - 1) / kSetsPerTile;
    x += (size_t) b * max_pillars * kC;
    pos += (size_t) b * max_pillars * kC;
    out += (size_t) b * max_pillars * kC;
    idx += ((size_t) b * 2 + axis) * max_sets * kS;
    mask += (size_t) b * max_sets * kH * kS;

    if (zero_tails) {
        int V = voxel_num[b];
        V = V < max_pillars ? V : max_pillars;
        float4* o4 = reinterpret_cast<float4*>(out + (size_t) V * kC);
        const long long n4 = (long long) (max_pillars - V) * (kC / 4);
        for (long long t = (long long) blockIdx.x * kThreads + tid; t < n4; t += (long long) gridDim.x * kThreads)
            o4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if ((int) blockIdx.x >= n_tiles) return;

    if (tid == 0) {
        mbar_init(&bars.w_full, 1); mbar_init(&bars.wout_full, 1);
        mbar_init(&bars.a_full, kWorkers); mbar_init(&bars.so_full, 128); mbar_init(&bars.out_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars.proj_full[i], 1); mbar_init(&bars.proj_free[i], 128); mbar_init(&bars.qkv_full[i], 128);
            mbar_init(&bars.s_full[i], 1); mbar_init(&bars.p_full[i], 128);
            mbar_init(&bars.o_full[i], 1); mbar_init(&bars.o_free[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc<512>(&tmem_slot);
    // padding that stays zero: chunk 3 of Q/K, d-group 3 of V, in both buffers
    for (int t = tid; t < 2 * kRows; t += kThreads) {
        uint8_t* base = smem + SM_QKV + (t >> 7) * kQkvBytes;
        const int r = t & 127;
        *reinterpret_cast<uint4*>(base + 3 * kChunkStride + r * 16) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(base + 8192 + 3 * kChunkStride + r * 16) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(base + 16384 + (r >> 3) * 512 + 3 * 128 + (r & 7) * 16) = make_uint4(0, 0, 0, 0);
    }
    for (int t = tid; t < 4 * kC; t += kThreads) s_bias[t] = __ldg(wb.bias + t);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t sbase = smem_u32(smem);

    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t) (q4 * 32) << 16);
    // number of completed uses of each double-buffered barrier pair, per role (parity = count & 1)
    uint32_t tile_iter = 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
        const int set0 = tile * kSetsPerTile;
        const uint32_t tpar = tile_iter & 1;             // parity of per-tile barriers (one use per tile)
        // the [2] barriers are used 4 times per tile each (heads b, b+2, b+4, b+6): parity of use n is (4*tile_iter + n) & 1 = n & 1

        if (warp == 8) {
            // =========================== ISSUER ==========================================================
            if (lane == 0) {
                const uint32_t idesc_qk = make_idesc(kFmtF16, kRows, 48), idesc_v = make_idesc(kFmtF16, kRows, 32);
                const uint32_t idesc_s = make_idesc(kFmtF16, kRows, kKeys), idesc_pv = make_idesc(kFmtF16, kRows, 32, 0, 1);
                const uint32_t idesc_out = make_idesc(kFmtF16, kRows, 96);
                const uint64_t d_aqk = make_smem_desc(sbase + SM_AQK, kChunkStride, 128);
                const uint64_t d_av = make_smem_desc(sbase + SM_AV, kChunkStride, 128);
                const uint64_t d_o = make_smem_desc(sbase + SM_O, kChunkStride, 128);
                const uint64_t d_wqk = make_smem_desc(sbase + SM_W, 48 * 16, 128);
                const uint64_t d_wv = make_smem_desc(sbase + SM_W + kWqkBytes, 32 * 16, 128);
                const uint64_t d_wo0 = make_smem_desc(sbase + SM_AQK, 96 * 16, 128);
                const uint64_t d_wo1 = make_smem_desc(sbase + SM_AV, 96 * 16, 128);
                uint32_t ph_w = (tile_iter * 8) & 1;     // 8 weight loads per tile -> parity restarts even each tile
                auto load_w = [&](int h) {
                    mbar_arrive_expect_tx(&bars.w_full, kWHeadBytes);
                    bulk_g2s(smem + SM_W, wb.w_img + (size_t) h * kWHeadBytes, kWHeadBytes, &bars.w_full);
                };
                P2(0);
                load_w(0);
                mbar_wait(&bars.a_full, tpar);           // A tiles staged by the workers
                P2(1);
                // Dynamic schedule: never block on one dependency while another MMA group is ready.  Per head the
                // issuer owes S(h), PV(h) and proj(h) (+ the weight copy that precedes each proj); each is issued as
                // soon as its barriers have completed, S first (it feeds the softmax group, the slowest engine).
                int n_proj = 0, n_s = 0, n_pv = 0, n_w = 1;      // issued so far; weights: n_w copies issued
                bool w_inflight = true;                           // copy n_w-1 not yet consumed by a proj
                while (n_pv < kH) {
                    bool progressed = false;
                    if (n_s < kH && n_s < n_proj && mbar_try_wait(&bars.qkv_full[n_s & 1], (n_s >> 1) & 1)) {
                        const int h = n_s, bb = h & 1;
                        tc_fence_after_sync();
                        const uint64_t d_q = make_smem_desc(sbase + SM_QKV + bb * kQkvBytes, kChunkStride, 128);
                        const uint64_t d_k = make_smem_desc(sbase + SM_QKV + bb * kQkvBytes + 8192, kChunkStride, 128);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            umma_f16(tmem + C_SP + bb * kSpCols, d_q + ks * (2 * kChunkStride / 16),
                                     d_k + ks * (2 * kChunkStride / 16), idesc_s, ks > 0);
                        umma_commit(&bars.s_full[bb]);
                        ++n_s; progressed = true;
                    }
                    if (n_pv < n_s && mbar_try_wait(&bars.p_full[n_pv & 1], (n_pv >> 1) & 1) &&
                        (n_pv < 2 || mbar_try_wait(&bars.o_free[n_pv & 1], ((n_pv >> 1) - 1) & 1))) {
                        const int h = n_pv, bb = h & 1;
                        tc_fence_after_sync();
                        const uint64_t d_v = make_smem_desc(sbase + SM_QKV + bb * kQkvBytes + 16384, 512, 128);
#pragma unroll
                        for (int ks = 0; ks < kKeys / 16; ++ks)
                            umma_f16_ts(tmem + C_O + bb * kOCols, tmem + C_SP + bb * kSpCols + 8 * ks, d_v + ks * (1024 / 16),
                                        idesc_pv, ks > 0);
                        umma_commit(&bars.o_full[bb]);
                        ++n_pv; progressed = true;
                        if (h < kH) P2(2 + h);
                    }
                    // weight copy for the next projection: sW is free once the previous projection has completed
                    if (!w_inflight && n_w < kH && mbar_try_wait(&bars.proj_full[(n_w - 1) & 1], ((n_w - 1) >> 1) & 1)) {
                        load_w(n_w);
                        ++n_w; w_inflight = true; progressed = true;
                    }
                    if (n_proj < kH && w_inflight && n_proj == n_w - 1 && mbar_try_wait(&bars.w_full, ph_w) &&
                        (n_proj < 2 || mbar_try_wait(&bars.proj_free[n_proj & 1], ((n_proj >> 1) - 1) & 1))) {
                        const int h = n_proj;
                        ph_w ^= 1;
                        tc_fence_after_sync();
                        const uint32_t dq = tmem + C_PROJ + (h & 1) * kProjCols, dv = dq + 48;
                        uint64_t aq = d_aqk, av = d_av, bq = d_wqk, bv = d_wv;
#pragma unroll
                        for (int ks = 0; ks < 12; ++ks) {
                            umma_f16(dq, aq, bq, idesc_qk, ks > 0);
                            umma_f16(dv, av, bv, idesc_v, ks > 0);
                            aq += 2 * kChunkStride / 16; av += 2 * kChunkStride / 16; bq += 2 * 48; bv += 2 * 32;
                        }
                        umma_commit(&bars.proj_full[h & 1]);
                        ++n_proj; w_inflight = false; progressed = true;
                    }
                    if (!progressed) __nanosleep(20);
                }
                P2(10);
                // out-projection: Wout halves into the (dead) A tiles, O tile from the E group
                mbar_wait(&bars.proj_full[1], 1);        // proj(7) complete -> sAqk / sAv free   (4th use of parity-1 barrier)
                mbar_arrive_expect_tx(&bars.wout_full, 2 * kWoutHalfBytes);
                bulk_g2s(smem + SM_AQK, wb.w_img + (size_t) kH * kWHeadBytes, kWoutHalfBytes, &bars.wout_full);
                bulk_g2s(smem + SM_AV, wb.w_img + (size_t) kH * kWHeadBytes + kWoutHalfBytes, kWoutHalfBytes, &bars.wout_full);
                mbar_wait(&bars.wout_full, tpar);
                mbar_wait(&bars.so_full, tpar);
                P2(11);
                tc_fence_after_sync();
#pragma unroll
                for (int ks = 0; ks < 12; ++ks) {
                    umma_f16(tmem + C_OUT, d_o + ks * (2 * kChunkStride / 16), d_wo0 + ks * (2 * 96), idesc_out, ks > 0);
                    umma_f16(tmem + C_OUT + 96, d_o + ks * (2 * kChunkStride / 16), d_wo1 + ks * (2 * 96), idesc_out, ks > 0);
                }
                umma_commit(&bars.out_full);
                P2(12);
            }
            __syncwarp();        // reconverge the issuer warp before the CTA-wide barrier below
        } else {
            // =========================== WORKERS: stage the token tile ====================================
            const int wt = tid;                          // 0..255
            if (wt < kRows) {
                const int r = wt, st = set0 + r / kS;
                s_rows[r] = (r < kValidRows && st < ns) ? idx[(size_t) st * kS + (r % kS)] : -1;
            }
            workers_sync();
            {
                const int r = wt & (kRows - 1);
                const int g = s_rows[r];
                const float4* xr = reinterpret_cast<const float4*>(x + (size_t) (g < 0 ? 0 : g) * kC);
                const float4* pr = reinterpret_cast<const float4*>(pos + (size_t) (g < 0 ? 0 : g) * kC);
                constexpr int NJ = 4;
                for (int c0 = wt >> 7; c0 < 24; c0 += 2 * NJ) {
                    float4 a[NJ][2], p[NJ][2];
                    if (g >= 0) {
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const int c = c0 + 2 * j;
                            a[j][0] = __ldg(xr + 2 * c); a[j][1] = __ldg(xr + 2 * c + 1);
                            p[j][0] = __ldg(pr + 2 * c); p[j][1] = __ldg(pr + 2 * c + 1);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int c = c0 + 2 * j;
                        uint4 vq = make_uint4(0, 0, 0, 0), vv = vq;
                        if (g >= 0) {
                            const float4 a0 = a[j][0], a1 = a[j][1], p0 = p[j][0], p1 = p[j][1];
                            vv = make_uint4(pack_h2(a0.x, a0.y), pack_h2(a0.z, a0.w), pack_h2(a1.x, a1.y), pack_h2(a1.z, a1.w));
                            vq = make_uint4(pack_h2(a0.x + p0.x, a0.y + p0.y), pack_h2(a0.z + p0.z, a0.w + p0.w),
                                            pack_h2(a1.x + p1.x, a1.y + p1.y), pack_h2(a1.z + p1.z, a1.w + p1.w));
                        }
                        *reinterpret_cast<uint4*>(smem + SM_AQK + c * kChunkStride + r * 16) = vq;
                        *reinterpret_cast<uint4*>(smem + SM_AV + c * kChunkStride + r * 16) = vv;
                    }
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars.a_full);
            if (tid == 0) P2(16);

            if (warp < 4) {
                // =========================== E GROUP ======================================================
                auto o_epilogue = [&](int h) {           // O[h&1] (TMEM) -> FP16 columns 24h..24h+23 of the O operand tile
                    const int bb = h & 1, n = h >> 1;
                    mbar_wait(&bars.o_full[bb], n & 1);
                    tc_fence_after_sync();
                    uint32_t a[16], c8[8];
                    tmem_ld16(tlane + C_O + bb * kOCols, a);
                    tmem_ld8(tlane + C_O + bb * kOCols + 16, c8);
                    tmem_ld_wait();
                    tc_fence_before_sync();
                    mbar_arrive(&bars.o_free[bb]);
                    *reinterpret_cast<uint4*>(smem + SM_O + (3 * h + 0) * kChunkStride + row * 16) = cvt8_nobias(a);
                    *reinterpret_cast<uint4*>(smem + SM_O + (3 * h + 1) * kChunkStride + row * 16) = cvt8_nobias(a + 8);
                    *reinterpret_cast<uint4*>(smem + SM_O + (3 * h + 2) * kChunkStride + row * 16) = cvt8_nobias(c8);
                };
                for (int h = 0; h < kH; ++h) {
                    const int bb = h & 1, n = h >> 1;
                    // the Q/K/V tiles of parity bb were last read by S(h-2) / PV(h-2): PV(h-2) complete == o_full
                    if (h >= 2) o_epilogue(h - 2);
                    mbar_wait(&bars.proj_full[bb], n & 1);
                    tc_fence_after_sync();
                    uint32_t qa[16], qb[8], ka[16], kb[8], va[16], vb[8];
                    const uint32_t pb = tlane + C_PROJ + bb * kProjCols;
                    tmem_ld16(pb, qa);      tmem_ld8(pb + 16, qb);
                    tmem_ld16(pb + 24, ka); tmem_ld8(pb + 40, kb);
                    tmem_ld16(pb + 48, va); tmem_ld8(pb + 64, vb);
                    tmem_ld_wait();
                    tc_fence_before_sync();
                    mbar_arrive(&bars.proj_free[bb]);
                    uint8_t* qkv = smem + SM_QKV + bb * kQkvBytes;
                    const float* bq = s_bias + h * kD; const float* bk = s_bias + kC + h * kD; const float* bv = s_bias + 2 * kC + h * kD;
                    *reinterpret_cast<uint4*>(qkv + 0 * kChunkStride + row * 16) = cvt8(qa, bq);
                    *reinterpret_cast<uint4*>(qkv + 1 * kChunkStride + row * 16) = cvt8(qa + 8, bq + 8);
                    *reinterpret_cast<uint4*>(qkv + 2 * kChunkStride + row * 16) = cvt8(qb, bq + 16);
                    *reinterpret_cast<uint4*>(qkv + 8192 + 0 * kChunkStride + row * 16) = cvt8(ka, bk);
                    *reinterpret_cast<uint4*>(qkv + 8192 + 1 * kChunkStride + row * 16) = cvt8(ka + 8, bk + 8);
                    *reinterpret_cast<uint4*>(qkv + 8192 + 2 * kChunkStride + row * 16) = cvt8(kb, bk + 16);
                    *reinterpret_cast<uint4*>(qkv + 16384 + (row >> 3) * 512 + 0 * 128 + (row & 7) * 16) = cvt8(va, bv);
                    *reinterpret_cast<uint4*>(qkv + 16384 + (row >> 3) * 512 + 1 * 128 + (row & 7) * 16) = cvt8(va + 8, bv + 8);
                    *reinterpret_cast<uint4*>(qkv + 16384 + (row >> 3) * 512 + 2 * 128 + (row & 7) * 16) = cvt8(vb, bv + 16);
                    fence_proxy_async_smem();
                    mbar_arrive(&bars.qkv_full[bb]);
                    if (tid == 0) P2(17 + h);
                }
                o_epilogue(kH - 2);
                o_epilogue(kH - 1);
                fence_proxy_async_smem();
                mbar_arrive(&bars.so_full);
            } else {
                // =========================== S GROUP ======================================================
                const int sw = warp - 4;                 // TMEM lane quarter
                const int a_set = sw == 0 ? 0 : sw - 1;
                const bool two = (sw == 1) || (sw == 2);
                const int sl = row / kS;
                const bool use_b = two && (sl == a_set + 1);
                const int st = set0 + sl;
                const bool live = (row < kValidRows) && (st < ns);
                const int my_off = (use_b ? a_set + 1 : a_set) * 18;     // packed column of this row's first key
                for (int h = 0; h < kH; ++h) {
                    const int bb = h & 1, n = h >> 1;
                    const float* mrow = mask + ((size_t) (live ? st : 0) * kH + h) * kS;
                    if (live) asm volatile("prefetch.global.L1 [%0];" ::"l"(mrow));   // 144 B: lands while we wait for S
                    if (live) asm volatile("prefetch.global.L1 [%0];" ::"l"(mrow + 32));
                    mbar_wait(&bars.s_full[bb], n & 1);
                    tc_fence_after_sync();
                    const uint32_t sp = tlane + C_SP + bb * kSpCols;
                    uint32_t ra[32], ra4[4], rb[32], rb4[4];
                    tmem_ld32(sp + a_set * kS, ra);
                    tmem_ld4(sp + a_set * kS + 32, ra4);
                    if (two) {
                        tmem_ld32(sp + a_set * kS + kS, rb);
                        tmem_ld4(sp + a_set * kS + kS + 32, rb4);
                    }
                    tmem_ld_wait();
                    float s[kS];
#pragma unroll
                    for (int k = 0; k < 32; ++k) s[k] = __uint_as_float(use_b ? rb[k] : ra[k]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) s[32 + k] = __uint_as_float(use_b ? rb4[k] : ra4[k]);
                    if (live) {
                        const float4* mp = reinterpret_cast<const float4*>(mrow);
#pragma unroll
                        for (int k4 = 0; k4 < kS / 4; ++k4) {
                            const float4 m = __ldg(mp + k4);
                            s[4 * k4] += m.x; s[4 * k4 + 1] += m.y; s[4 * k4 + 2] += m.z; s[4 * k4 + 3] += m.w;
                        }
                    }
                    float m4[4] = {s[0], s[1], s[2], s[3]};
#pragma unroll
                    for (int k = 4; k < kS; ++k) m4[k & 3] = fmaxf(m4[k & 3], s[k]);
                    const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                    float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int k = 0; k < kS; ++k) { s[k] = exp2f((s[k] - mx) * 1.4426950408889634f); a4[k & 3] += s[k]; }
                    const float inv = live ? 1.0f / ((a4[0] + a4[1]) + (a4[2] + a4[3])) : 0.f;
                    uint32_t pk[18];
#pragma unroll
                    for (int k = 0; k < 18; ++k) pk[k] = pack_h2(s[2 * k] * inv, s[2 * k + 1] * inv);
                    // P aliases the first 56 columns of this S buffer: the whole packed row is rewritten (zeros outside
                    // the row's own 18 columns).  Column c of the row holds keys 2c, 2c+1.
                    uint32_t p0[32], p1[16], p2[8];
                    const bool s0 = my_off == 0, s1 = my_off == 18, s2 = my_off == 36;
#pragma unroll
                    for (int c = 0; c < 32; ++c)             // packed columns 0..31: set 0 owns 0..17, set 1 owns 18..35
                        p0[c] = c < 18 ? (s0 ? pk[c] : 0u) : (s1 ? pk[c - 18] : 0u);
#pragma unroll
                    for (int c = 0; c < 16; ++c)             // packed columns 32..47: set 1 owns ..35, set 2 owns 36..53
                        p1[c] = c < 4 ? (s1 ? pk[c + 14] : 0u) : (s2 ? pk[c - 4] : 0u);
#pragma unroll
                    for (int c = 0; c < 8; ++c)              // packed columns 48..55
                        p2[c] = c < 6 ? (s2 ? pk[c + 12] : 0u) : 0u;
                    tmem_st32(sp, p0);
                    tmem_st16(sp + 32, p1);
                    tmem_st8(sp + 48, p2);
                    tmem_st_wait();
                    tc_fence_before_sync();
                    mbar_arrive(&bars.p_full[bb]);
                    if (tid == 128) P2(32 + h);
                }
            }
            // =========================== WORKERS: final epilogue ==========================================
            mbar_wait(&bars.out_full, tpar);
            tc_fence_after_sync();
            if (tid == 0) P2(26);
            {
                constexpr int kOutStride = 196;
                float* s_out = reinterpret_cast<float*>(smem + SM_AQK);   // Wout tiles are dead: out-proj MMAs completed
                const int hf = warp >> 2;
                const float* bo = s_bias + 3 * kC + hf * 96;
#pragma unroll 1
                for (int j0 = 0; j0 < 96; j0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tlane + C_OUT + hf * 96 + j0, r);
                    tmem_ld_wait();
                    if (row < kValidRows) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 o;
                            o.x = __uint_as_float(r[j + 0]) + bo[j0 + j + 0];
                            o.y = __uint_as_float(r[j + 1]) + bo[j0 + j + 1];
                            o.z = __uint_as_float(r[j + 2]) + bo[j0 + j + 2];
                            o.w = __uint_as_float(r[j + 3]) + bo[j0 + j + 3];
                            *reinterpret_cast<float4*>(s_out + row * kOutStride + hf * 96 + j0 + j) = o;
                        }
                    }
                }
                workers_sync();
                for (int i = tid; i < kValidRows * (kC / 4); i += kWorkers) {
                    const int rr = i / (kC / 4), c4 = i - rr * (kC / 4);
                    const int g = s_rows[rr];
                    if (g >= 0)
                        *reinterpret_cast<float4*>(out + (size_t) g * kC + c4 * 4) =
                            *reinterpret_cast<const float4*>(s_out + rr * kOutStride + c4 * 4);
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();          // everything (smem tiles, TMEM, s_rows) is recycled by the next tile
        tc_fence_after_sync();
        if (tid == 0) P2(27);
    }
    if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace

// ---- host: FP16 operand images of one attention layer ----------------------------------------------------------
void* attention_tc_prepare(int C, int H, const float* w_in, const float* b_in, const float* w_out, const float* b_out)
{
    if (C != kC || H != kH) return nullptr;
    const float inv_scale = 1.0f / sqrtf((float) kD);
    std::vector<uint8_t> blob(kWImgBytes + 4 * kC * sizeof(float), 0);
    auto put = [&](uint8_t* tile, int rows, int n, int k, float v) {   // [chunk][row][8 halves]
        const __half hv = __float2half_rn(v);
        memcpy(tile + ((size_t) (k / 8) * rows + n) * 16 + (k % 8) * 2, &hv, 2);
    };
    for (int h = 0; h < kH; ++h) {
        uint8_t* qk = blob.data() + (size_t) h * kWHeadBytes;
        uint8_t* vv = qk + kWqkBytes;
        for (int d = 0; d < kD; ++d)
            for (int k = 0; k < kC; ++k) {
                put(qk, 48, d, k, w_in[(size_t) (0 * kC + h * kD + d) * kC + k] * inv_scale);      // query rows, pre-scaled
                put(qk, 48, 24 + d, k, w_in[(size_t) (1 * kC + h * kD + d) * kC + k]);             // key rows
                put(vv, 32, d, k, w_in[(size_t) (2 * kC + h * kD + d) * kC + k]);                  // value rows (+8 zero rows)
            }
    }
    for (int half = 0; half < 2; ++half) {
        uint8_t* t = blob.data() + (size_t) kH * kWHeadBytes + (size_t) half * kWoutHalfBytes;
        for (int n = 0; n < 96; ++n)
            for (int k = 0; k < kC; ++k) put(t, 96, n, k, w_out[(size_t) (half * 96 + n) * kC + k]);
    }
    float* fb = reinterpret_cast<float*>(blob.data() + kWImgBytes);
    for (int i = 0; i < kC; ++i) {
        fb[i] = b_in[i] * inv_scale;
        fb[kC + i] = b_in[kC + i];
        fb[2 * kC + i] = b_in[2 * kC + i];
        fb[3 * kC + i] = b_out[i];
    }
    void* dev = nullptr;
    if (cudaMalloc(&dev, blob.size()) != cudaSuccess) return nullptr;
    if (cudaMemcpy(dev, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dev); return nullptr; }
    return dev;
}

int set_attention_tc2_fused(const dsvt_set_attention_params* p, const void* tc_blob,
                            const float* x, const float* pos, const int* idx, const float* mask,
                            const int* set_num, const int* voxel_num, float* out, cudaStream_t st)
{
    if (p->voxel_num_set != kS || p->channel_num != kC || p->num_heads != kH) {
        set_last_error("set attention (FP16 tensor-core path): only set=36, C=192, heads=8 is built");
        return DSVT_ERR_UNSUPPORTED;
    }
    if (!tc_blob) {
        set_last_error("set attention (FP16 tensor-core path): weights were not prepared");
        return DSVT_ERR_INVALID_ARGUMENT;
    }
    TcBlobView2 wb;
    wb.w_img = static_cast<const uint8_t*>(tc_blob);
    wb.bias = reinterpret_cast<const float*>(wb.w_img + kWImgBytes);
    DSVT_RAISE_SMEM(set_attention_tc2_kernel, SM_TOTAL);
    const int max_tiles = (p->max_set_num + kSetsPerTile - 1) / kSetsPerTile;
    const int grid = max_tiles < sm_count() ? max_tiles : sm_count();
    set_attention_tc2_kernel<<<dim3(grid, p->batch), kThreads, SM_TOTAL, st>>>(
        x, pos, idx, mask, set_num, voxel_num, out, wb, p->max_set_num, p->max_pillars_num, p->axis_id, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

}  // namespace dsvt

#ifdef DSVT_PROFILE
extern "C" int dsvt_debug_tc2_profile(long long* out64) {
    return cudaMemcpyFromSymbol(out64, dsvt::g_tc2_prof, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif
