// a3 -- set attention, FP32 CUDA-core path (DSVT_ATTN_FP32).
// Restates the TensorRT layer sub-graph built by multHeadAttention() (reference
// src/dsvt-ai-trt.cpp:288-458): per set of S tokens, nn.MultiheadAttention(C=192, heads=8) with an
// additive key mask, q scaled by a DIVISION by sqrt(C/heads) after the biased projection (:386-405).
//
// One CTA per set keeps everything on chip: token tile -> Q,K,V -> scores -> softmax -> PV ->
// out-projection, no intermediate ever touches HBM (the reference materialises ~25 tensors of
// 22 MB each per call).  The fused entry point also folds GetValueByIndex (gather x+pos / x,
// getValueByIndex.cu:282-303) into the tile load and MapSetFeature2Voxel
// (mapSetFeature2voxel.cu:258-275) into the epilogue.
// This is the exact-arithmetic mode and the on-GPU yard-stick for the tcgen05 path
// (attention_tc.cu); weights stream from L2 (0.59 MB per layer, resident).
#include "attention_common.cuh"

namespace dsvt {
namespace {

constexpr int kC = 192;
constexpr int kThreadsA = 288;            // 48 column groups (4 columns each) x 6 row groups
constexpr int kColGroups = kC / 4;        // 48
constexpr int kRowGroups = kThreadsA / kColGroups;   // 6
constexpr int kXs = 196;                  // row stride (floats) of every tile: 49 x 16 B, odd in 16-byte units ->
                                          // conflict-free 128-bit accesses whether lanes walk rows or columns
constexpr int kHeadGroup = 4;             // heads whose score matrices are resident at a time

// Per set: X (token tile) and the score buffer of one head group share a region (X is dead once V is projected);
// the attention output O is written in place over Q (head h's Q columns are dead after head h's scores).
// S = 36: 4 x 28.2 KB = 112.9 KB per CTA -> two CTAs (18 warps) per SM.  (A variant with two sets per CTA and one CTA
// per SM was measured 20 % slower: halving the warps hurts the score / softmax / PV phases more than sharing the
// weight stream helps.)
template <int S> struct AttnSmem {
    static constexpr int NS = 1;     // sets per CTA (2 was measured slower: 9 warps/SM starve the non-GEMM phases)
    static constexpr int tile = S * kXs;
    static constexpr int sc_floats = kHeadGroup * S * (S + 1);
    static constexpr int xreg = tile > sc_floats ? tile : sc_floats;
    static constexpr int x = 0;                       // [NS][xreg]
    static constexpr int q = x + NS * xreg;           // [NS][tile]
    static constexpr int k = q + NS * tile;
    static constexpr int v = k + NS * tile;
    static constexpr int total = v + NS * tile;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t) __cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kStageK = 8;                          // k-values per weight stage
constexpr int kStageFloats = kStageK * kC;          // 1536 floats = 6 KB
// ring depth: 4 stages (24 KB) fit the 28 KB destination tile of S >= 36, 3 stages (18 KB) that of S = 24
template <int S> struct RingDepth { static constexpr int value = (S * kXs >= 4 * kStageFloats) ? 4 : 3; };

// out[r][4cg..4cg+3] = (sum_k X[r][k] * Wt[k][4cg..] + bias) / div for this thread's R rows.
// The 147 KB weight slice streams from L2 ONCE per CTA through a 4-stage cp.async ring in shared memory (the ring
// lives in the projection's own destination tile, which is only written after the k loop), so neither the L2
// latency (~1000 cycles under load: the previous register-prefetch version was bound by it) nor the six-fold
// re-read by the six row groups is on the critical path.  Register tile R x 4; row groups whose rows are all
// >= n_rows (token compaction) skip the FMAs but still help with the copies.
template <int S, bool TO_GLOBAL>
__device__ __forceinline__ void project(const float* __restrict__ Xs, const float* __restrict__ Wt, int ldw,
                                        const float* __restrict__ bias, float div, float* ring, float* out_s,
                                        float* out_g, const int* row_map, int n_rows)
{
    constexpr int R = S / kRowGroups;
    constexpr int kStages = RingDepth<S>::value;
    constexpr int kNumStages = kC / kStageK;        // 24
    const int tid = threadIdx.x;
    const int cg = tid % kColGroups;
    const int r0 = (tid / kColGroups) * R;
    const bool active = r0 < n_rows;
    // (walking k from a per-CTA offset, to de-correlate the CTAs' L2 requests, was measured: no gain; not kept so
    // that a set's result does not depend on which CTA computes it)
    constexpr int rot = 0;
    float acc[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }

    auto issue = [&](int st) {                      // copy 8 weight rows x 192 columns into ring slot st % kStages
        if (st < kNumStages) {
            float* dst = ring + (st % kStages) * kStageFloats;
            const int slab = (st + rot) % kNumStages;
            for (int c = tid; c < kStageFloats / 4; c += kThreadsA) {
                const int kk = c / (kC / 4), j = c - kk * (kC / 4);
                cp_async16(dst + kk * kC + j * 4, Wt + (size_t) (slab * kStageK + kk) * ldw + j * 4);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int st = 0; st < kStages - 1; ++st) issue(st);
#pragma unroll 1
    for (int st = 0; st < kNumStages; ++st) {
        cp_async_wait<kStages - 2>();               // this thread's copies of stage st have landed
        __syncthreads();                            // ... everyone's have, and slot (st-1)%kStages is no longer being read
        issue(st + kStages - 1);
        if (active) {
            const float* wst = ring + (st % kStages) * kStageFloats + cg * 4;
#pragma unroll
            for (int k4 = 0; k4 < kStageK / 4; ++k4) {
                float4 wc[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) wc[kk] = *reinterpret_cast<const float4*>(wst + (k4 * 4 + kk) * kC);
                const int k = ((st + rot) % kNumStages) * kStageK + k4 * 4;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r0 + r >= n_rows) break;
                    const float4 xv = *reinterpret_cast<const float4*>(Xs + (r0 + r) * kXs + k);
                    const float xk[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        acc[r][0] = fmaf(xk[kk], wc[kk].x, acc[r][0]);
                        acc[r][1] = fmaf(xk[kk], wc[kk].y, acc[r][1]);
                        acc[r][2] = fmaf(xk[kk], wc[kk].z, acc[r][2]);
                        acc[r][3] = fmaf(xk[kk], wc[kk].w, acc[r][3]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();                                // the ring (== the destination tile) is free to be overwritten
    if (!active) return;
    const float4 bn = __ldg(reinterpret_cast<const float4*>(bias) + cg);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int lr = r0 + r;
        if (lr >= n_rows) break;
        float4 o = make_float4(acc[r][0] + bn.x, acc[r][1] + bn.y, acc[r][2] + bn.z, acc[r][3] + bn.w);
        if (div != 1.0f) { o.x = o.x / div; o.y = o.y / div; o.z = o.z / div; o.w = o.w / div; }
        if (TO_GLOBAL) {
            const int row = row_map ? row_map[lr] : lr;
            *reinterpret_cast<float4*>(out_g + (size_t) row * kC + cg * 4) = o;
        } else {
            *reinterpret_cast<float4*>(out_s + lr * kXs + cg * 4) = o;
        }
    }
}

#ifdef DSVT_PROFILE      // phase stamps of one CTA (tools/fp32_profile.py): never in the product build
__device__ long long g_fp32_prof[32];
#define F32_PROF(i) do { if (blockIdx.x == 5 && blockIdx.y == 0 && threadIdx.x == 0) g_fp32_prof[i] = clock64(); } while (0)
#else
#define F32_PROF(i) do { } while (0)
#endif

template <int S, bool FUSED>
__global__ void __launch_bounds__(kThreadsA, (S <= 36) ? 2 : 1)
set_attention_fp32_kernel(const float* __restrict__ qin, const float* __restrict__ kin, const float* __restrict__ vin,
                          const float* __restrict__ pos, const int* __restrict__ idx,
                          const float* __restrict__ mask, const int* __restrict__ set_num,
                          const int* __restrict__ voxel_num, float* __restrict__ out,
                          AttnWeightsDev w, int max_sets, int max_pillars, int axis, int zero_tails)
{
    extern __shared__ __align__(16) float sm[];
    using L = AttnSmem<S>;
    constexpr int NS = L::NS;
    constexpr int TILE = L::tile;
    float* Xs = sm + L::x;      // [NS][xreg]
    float* Qs = sm + L::q;      // [NS][tile]; becomes the attention output O in place
    float* Ks = sm + L::k;
    float* Vs = sm + L::v;
    float* Ss = sm + L::x;      // scores of one head group per set, aliases the (dead) token tile
    __shared__ int s_rows[NS * S];      // voxel row of compact token u
    __shared__ int s_slot[NS * S];      // original slot of compact token u (for the key mask)
    __shared__ int s_nu;

    constexpr int H = 8, D = kC / H;
    const int b = blockIdx.y;
    const int set0 = blockIdx.x * NS;
    const int tid = threadIdx.x;
    F32_PROF(0);
    int ns = set_num ? set_num[b] : max_sets;
    ns = ns < max_sets ? ns : max_sets;

    if (set0 >= ns) {
        if (!zero_tails) return;
        if (!FUSED) {
            for (int s = 0; s < NS; ++s) {
                if (set0 + s >= max_sets) break;
                float4* o = reinterpret_cast<float4*>(out + ((size_t) b * max_sets + set0 + s) * S * kC);
                for (int t = tid; t < S * kC / 4; t += kThreadsA) o[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            // tail rows [voxel_num, max_pillars) are split over the idle CTAs
            int V = voxel_num[b];
            V = V < max_pillars ? V : max_pillars;
            const int first_idle = (ns + NS - 1) / NS;
            const int idle = (int) gridDim.x - first_idle;
            const long long tail = (long long) (max_pillars - V) * (kC / 4);
            const long long per = (tail + idle - 1) / idle;
            long long lo = (long long) ((int) blockIdx.x - first_idle) * per, hi = lo + per;
            if (hi > tail) hi = tail;
            float4* o = reinterpret_cast<float4*>(out + ((size_t) b * max_pillars + V) * kC);
            for (long long t = lo + tid; t < hi; t += kThreadsA) o[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    bool live[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) live[s] = set0 + s < ns;
    if (!FUSED && zero_tails) {                       // an odd trailing set of the pair: zero its output rows
        for (int s = 0; s < NS; ++s) {
            if (live[s] || set0 + s >= max_sets) continue;
            float4* o = reinterpret_cast<float4*>(out + ((size_t) b * max_sets + set0 + s) * S * kC);
            for (int t = tid; t < S * kC / 4; t += kThreadsA) o[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    // Token compaction (fused form).  getSet pads a set by repeating the previous voxel and masks the repeats
    // (getSet.cu:546-563).  A repeated slot is the SAME token: as a query it yields the same row as its twin (and
    // scatters to the same voxel row), as a key it is masked out exactly (exp underflows to 0).  So only slots that
    // are a new voxel -- or that are not fully masked -- need projecting: work drops from 36 rows to the n_u
    // distinct tokens of the set (~1/2 at Waymo densities, ~1/3 on the reference frame) with identical results.
    if (FUSED) {
        static_assert(NS == 1, "token compaction assumes one set per CTA");
        const int* my_idx = idx + (((size_t) b * 2 + axis) * max_sets + set0) * S;
        if (tid < 32) {
            int base_u = 0;
            for (int k0 = 0; k0 < S; k0 += 32) {
                const int k = k0 + tid;
                bool keep = false;
                int g = 0;
                if (k < S) {
                    g = my_idx[k];
                    keep = (k == 0) || (g != my_idx[k - 1]);
                    if (!keep) {            // a repeat: droppable only if every head masks it as a key
                        const float* mrow = mask + ((size_t) b * max_sets + set0) * H * S + k;
                        for (int h = 0; h < H; ++h) keep |= !(mrow[h * S] < -1e30f);
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const int u = base_u + __popc(bal & ((1u << tid) - 1u));
                    s_rows[u] = g;
                    s_slot[u] = k;
                }
                base_u += __popc(bal);
            }
            if (tid == 0) s_nu = base_u;
        }
    } else {
        if (tid < NS * S) s_slot[tid] = tid % S;
        if (tid == 0) s_nu = S;
    }
    __syncthreads();
    const int nu = s_nu;
    F32_PROF(1);
#ifdef DSVT_PROFILE
    if (blockIdx.x == 5 && blockIdx.y == 0 && threadIdx.x == 0) g_fp32_prof[20] = nu;
#endif

    auto load_tile = [&](const float* src_plain, bool add_pos) {
        for (int t = tid; t < nu * (kC / 4); t += kThreadsA) {
            const int sr = t / (kC / 4), c4 = t - sr * (kC / 4), s = 0, r = sr;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live[s]) {
                if (FUSED) {
                    const size_t row = (size_t) b * max_pillars + s_rows[sr];
                    val = __ldg(reinterpret_cast<const float4*>(qin + row * kC) + c4);
                    if (add_pos) {
                        const float4 pp = __ldg(reinterpret_cast<const float4*>(pos + row * kC) + c4);
                        val.x += pp.x; val.y += pp.y; val.z += pp.z; val.w += pp.w;
                    }
                } else {
                    val = __ldg(reinterpret_cast<const float4*>(src_plain + (((size_t) b * max_sets + set0 + s) * S + r) * kC) + c4);
                }
            }
            *reinterpret_cast<float4*>(Xs + s * L::xreg + r * kXs + c4 * 4) = val;
        }
    };

    // ---- Q, K, V projections -------------------------------------------------
    const float scale = sqrtf((float) D);     // sqrt(dim_3 / num_heads), integer division (:386)
    load_tile(qin, true);
    __syncthreads();
    F32_PROF(2);
    project<S, false>(Xs, w.w_in_t + 0 * kC, 3 * kC, w.b_in + 0 * kC, scale, Qs, Qs, nullptr, nullptr, nu);
    if (!FUSED) {
        __syncthreads();
        load_tile(kin, true);
        __syncthreads();
    }
    F32_PROF(3);
    project<S, false>(Xs, w.w_in_t + 1 * kC, 3 * kC, w.b_in + 1 * kC, 1.0f, Ks, Ks, nullptr, nullptr, nu);
    __syncthreads();
    F32_PROF(4);
    load_tile(vin, false);
    __syncthreads();
    F32_PROF(5);
    project<S, false>(Xs, w.w_in_t + 2 * kC, 3 * kC, w.b_in + 2 * kC, 1.0f, Vs, Vs, nullptr, nullptr, nu);
    __syncthreads();

    F32_PROF(6);
    for (int hg = 0; hg < H; hg += kHeadGroup) {
        // ---- scores + mask (:410-412): thread (set, head, key j, half of the queries) keeps K[j] in registers -----
        for (int t = tid; t < NS * kHeadGroup * S * 2; t += kThreadsA) {
            const int s = t / (kHeadGroup * S * 2), t1 = t - s * (kHeadGroup * S * 2);
            const int half = t1 / (kHeadGroup * S), rem = t1 - half * (kHeadGroup * S);
            const int hl = rem / S, j = rem - hl * S, h = hg + hl;
            if (j >= nu) continue;
            const float* Kt = Ks + s * TILE;
            const float* Qt = Qs + s * TILE;
            float* St = Ss + s * L::xreg;
            float4 kr[D / 4];
#pragma unroll
            for (int d4 = 0; d4 < D / 4; ++d4) kr[d4] = *reinterpret_cast<const float4*>(Kt + j * kXs + h * D + d4 * 4);
            const float mj = live[s] ? __ldg(mask + (((size_t) b * max_sets + set0 + s) * H + h) * S + s_slot[j]) : 0.f;
            const int ih = (nu + 1) / 2, i0 = half * ih, i1 = half ? nu : ih;
#pragma unroll 2
            for (int i = i0; i < i1; ++i) {
                const float4* qp = reinterpret_cast<const float4*>(Qt + i * kXs + h * D);
                float a = 0.f;
#pragma unroll
                for (int d4 = 0; d4 < D / 4; ++d4) {
                    const float4 qv = qp[d4];
                    a = fmaf(qv.x, kr[d4].x, a); a = fmaf(qv.y, kr[d4].y, a);
                    a = fmaf(qv.z, kr[d4].z, a); a = fmaf(qv.w, kr[d4].w, a);
                }
                St[(hl * S + i) * (S + 1) + j] = a + mj;
            }
        }
        __syncthreads();
        if (hg == 0) F32_PROF(7);
        // ---- softmax over keys (:414-415) -----------------------------------------------------------------------
        for (int t = tid; t < NS * kHeadGroup * S; t += kThreadsA) {
            const int s = t / (kHeadGroup * S);
            if ((t - s * (kHeadGroup * S)) % S >= nu) continue;
            float* row = Ss + s * L::xreg + (t - s * (kHeadGroup * S)) * (S + 1);
            float mx = row[0];
            for (int j = 1; j < nu; ++j) mx = fmaxf(mx, row[j]);
            float sum = 0.f;
            for (int j = 0; j < nu; ++j) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
            const float inv = 1.0f / sum;
            for (int j = 0; j < nu; ++j) row[j] *= inv;
        }
        __syncthreads();
        if (hg == 0) F32_PROF(8);
        // ---- P.V (:417): head-major channel concat, written in place over this head group's Q columns ---------
        {
            constexpr int GC = kHeadGroup * D;               // 96 channels in this head group
            constexpr int IG = kThreadsA / GC;               // 3 query groups
            constexpr int RI = S / IG;                       // 12 queries per thread
            const int cl = tid % GC, i0 = (tid / GC) * RI, hl = cl / D, c = hg * D + cl;
#pragma unroll 1
            for (int s = 0; s < NS; ++s) {
                float acc[RI];
#pragma unroll
                for (int r = 0; r < RI; ++r) acc[r] = 0.f;
                if (i0 >= nu) continue;
                const float* pr = Ss + s * L::xreg + (hl * S + i0) * (S + 1);
                const float* Vt = Vs + s * TILE;
#pragma unroll 4
                for (int j = 0; j < nu; ++j) {
                    const float vj = Vt[j * kXs + c];
#pragma unroll
                    for (int r = 0; r < RI; ++r) acc[r] = fmaf(pr[r * (S + 1) + j], vj, acc[r]);   // rows >= nu: unused
                }
#pragma unroll
                for (int r = 0; r < RI; ++r)
                    if (i0 + r < nu) Qs[s * TILE + (i0 + r) * kXs + c] = acc[r];
            }
        }
        __syncthreads();
        if (hg == 0) F32_PROF(9);
    }
    F32_PROF(10);
    // ---- out-projection (:448) + (fused) scatter to voxel rows -----------------
    float* dst = FUSED ? out + (size_t) b * max_pillars * kC : out + ((size_t) b * max_sets + set0) * S * kC;
    project<S, true>(Qs, w.w_out_t, kC, w.b_out, 1.0f, Ks /* K is dead: weight ring */, nullptr, dst,
                     FUSED ? s_rows : nullptr, nu);
    F32_PROF(11);
}

template <int S, bool FUSED>
int launch_fp32(const dsvt_set_attention_params* p, const AttnWeightsDev& w,
                const float* q, const float* k, const float* v, const float* pos, const int* idx,
                const float* mask, const int* set_num, const int* voxel_num, float* out, cudaStream_t st)
{
    // (one CTA per SM with half of the SM left to L1 for the weight stream was measured 40 % slower)
    const size_t smem = (size_t) AttnSmem<S>::total * sizeof(float);
    DSVT_RAISE_SMEM((set_attention_fp32_kernel<S, FUSED>), smem);
    constexpr int NS = AttnSmem<S>::NS;
    set_attention_fp32_kernel<S, FUSED><<<dim3((p->max_set_num + NS - 1) / NS, p->batch), kThreadsA, smem, st>>>(
        q, k, v, pos, idx, mask, set_num, voxel_num, out, w, p->max_set_num, p->max_pillars_num, p->axis_id,
        p->zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

}  // namespace

int set_attention_fp32(const dsvt_set_attention_params* p, const AttnWeightsDev& w, bool fused,
                       const float* q, const float* k, const float* v, const float* pos, const int* idx,
                       const float* mask, const int* set_num, const int* voxel_num, float* out, cudaStream_t st)
{
#define DSVT_DISPATCH_S(SV)                                                                                     \
    case SV:                                                                                                    \
        return fused ? launch_fp32<SV, true>(p, w, q, k, v, pos, idx, mask, set_num, voxel_num, out, st)         \
                     : launch_fp32<SV, false>(p, w, q, k, v, pos, idx, mask, set_num, voxel_num, out, st);
    switch (p->voxel_num_set) {
        DSVT_DISPATCH_S(24)
        DSVT_DISPATCH_S(36)
        DSVT_DISPATCH_S(48)
        default:
            set_last_error("set attention: voxel_num_set must be 24, 36 or 48 (got %d)", p->voxel_num_set);
            return DSVT_ERR_UNSUPPORTED;
    }
#undef DSVT_DISPATCH_S
}

}  // namespace dsvt

#ifdef DSVT_PROFILE
extern "C" int dsvt_debug_fp32_profile(long long* out32) {
    return cudaMemcpyFromSymbol(out32, dsvt::g_fp32_prof, sizeof(long long) * 32) == cudaSuccess ? 0 : 1;
}
#endif
