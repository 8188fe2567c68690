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
constexpr int kThreadsA = 384;
constexpr int kXs = 196;   // token-tile row stride (float4 aligned, broadcast reads)
constexpr int kQs = 193;   // Q/K/V row stride (odd: conflict-free column walks)

template <int S> struct AttnSmem {
    static constexpr int x = 0;
    static constexpr int q = x + S * kXs;
    static constexpr int k = q + S * kQs;
    static constexpr int v = k + S * kQs;
    static constexpr int s = v + S * kQs;
    static constexpr int total = s + 8 * S * (S + 1);
};

// out[r][n] = (sum_k X[r][k] * Wt[k][n] + bias[n]) / div   for r in this thread's half of the rows
template <int S, bool TO_GLOBAL>
__device__ __forceinline__ void project(const float* __restrict__ Xs, const float* __restrict__ Wt, int ldw,
                                        const float* __restrict__ bias, float div, float* out_s,
                                        float* out_g, const int* row_map, int C)
{
    constexpr int R = S / 2;
    const int n = threadIdx.x % kC;
    const int r0 = (threadIdx.x / kC) * R;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    for (int k = 0; k < kC; k += 4) {
        const float w0 = __ldg(Wt + (size_t) (k + 0) * ldw + n);
        const float w1 = __ldg(Wt + (size_t) (k + 1) * ldw + n);
        const float w2 = __ldg(Wt + (size_t) (k + 2) * ldw + n);
        const float w3 = __ldg(Wt + (size_t) (k + 3) * ldw + n);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(Xs + (r0 + r) * kXs + k);
            acc[r] = fmaf(xv.x, w0, acc[r]);
            acc[r] = fmaf(xv.y, w1, acc[r]);
            acc[r] = fmaf(xv.z, w2, acc[r]);
            acc[r] = fmaf(xv.w, w3, acc[r]);
        }
    }
    const float bn = __ldg(bias + n);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float o = acc[r] + bn;
        if (div != 1.0f) o = o / div;
        if (TO_GLOBAL) {
            const int row = row_map ? row_map[r0 + r] : (r0 + r);
            out_g[(size_t) row * C + n] = o;
        } else {
            out_s[(r0 + r) * kQs + n] = o;
        }
    }
}

template <int S, bool FUSED>
__global__ void __launch_bounds__(kThreadsA, 1)
set_attention_fp32_kernel(const float* __restrict__ qin, const float* __restrict__ kin, const float* __restrict__ vin,
                          const float* __restrict__ pos, const int* __restrict__ idx,
                          const float* __restrict__ mask, const int* __restrict__ set_num,
                          const int* __restrict__ voxel_num, float* __restrict__ out,
                          AttnWeightsDev w, int max_sets, int max_pillars, int axis, int zero_tails)
{
    extern __shared__ __align__(16) float sm[];
    using L = AttnSmem<S>;
    float* Xs = sm + L::x;
    float* Qs = sm + L::q;
    float* Ks = sm + L::k;
    float* Vs = sm + L::v;
    float* Ss = sm + L::s;
    __shared__ int s_rows[S];

    constexpr int H = 8, D = kC / H;
    const int b = blockIdx.y;
    const int set = blockIdx.x;
    const int tid = threadIdx.x;
    int ns = set_num ? set_num[b] : max_sets;
    ns = ns < max_sets ? ns : max_sets;

    if (set >= ns) {
        if (!zero_tails) return;
        if (!FUSED) {
            float4* o = reinterpret_cast<float4*>(out + ((size_t) b * max_sets + set) * S * kC);
            for (int t = tid; t < S * kC / 4; t += kThreadsA) o[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            // tail rows [voxel_num, max_pillars) are split over the idle CTAs
            int V = voxel_num[b];
            V = V < max_pillars ? V : max_pillars;
            const int idle = max_sets - ns;
            const long long tail = (long long) (max_pillars - V) * (kC / 4);
            const long long per = (tail + idle - 1) / idle;
            long long lo = (long long) (set - ns) * per, hi = lo + per;
            if (hi > tail) hi = tail;
            float4* o = reinterpret_cast<float4*>(out + ((size_t) b * max_pillars + V) * kC);
            for (long long t = lo + tid; t < hi; t += kThreadsA) o[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }

    const int* my_idx = FUSED ? idx + (((size_t) b * 2 + axis) * max_sets + set) * S : nullptr;
    if (FUSED && tid < S) s_rows[tid] = my_idx[tid];
    __syncthreads();

    auto load_tile = [&](const float* src_plain, bool add_pos) {
        for (int t = tid; t < S * (kC / 4); t += kThreadsA) {
            const int r = t / (kC / 4), c4 = t - r * (kC / 4);
            float4 val;
            if (FUSED) {
                const size_t row = (size_t) b * max_pillars + s_rows[r];
                val = __ldg(reinterpret_cast<const float4*>(qin + row * kC) + c4);
                if (add_pos) {
                    const float4 pp = __ldg(reinterpret_cast<const float4*>(pos + row * kC) + c4);
                    val.x += pp.x; val.y += pp.y; val.z += pp.z; val.w += pp.w;
                }
            } else {
                val = __ldg(reinterpret_cast<const float4*>(src_plain + (((size_t) b * max_sets + set) * S + r) * kC) + c4);
            }
            *reinterpret_cast<float4*>(Xs + r * kXs + c4 * 4) = val;
        }
    };

    // ---- Q, K, V projections -------------------------------------------------
    const float scale = sqrtf((float) D);     // sqrt(dim_3 / num_heads), integer division (:386)
    load_tile(qin, true);
    __syncthreads();
    project<S, false>(Xs, w.w_in_t + 0 * kC, 3 * kC, w.b_in + 0 * kC, scale, Qs, nullptr, nullptr, kC);
    if (!FUSED) {
        __syncthreads();
        load_tile(kin, true);
        __syncthreads();
    }
    project<S, false>(Xs, w.w_in_t + 1 * kC, 3 * kC, w.b_in + 1 * kC, 1.0f, Ks, nullptr, nullptr, kC);
    __syncthreads();
    load_tile(vin, false);
    __syncthreads();
    project<S, false>(Xs, w.w_in_t + 2 * kC, 3 * kC, w.b_in + 2 * kC, 1.0f, Vs, nullptr, nullptr, kC);
    __syncthreads();

    // ---- scores + mask (:410-412) ---------------------------------------------
    const float* mk = mask + ((size_t) b * max_sets + set) * H * S;
    for (int t = tid; t < H * S * S; t += kThreadsA) {
        const int h = t / (S * S), rem = t - h * S * S, i = rem / S, j = rem - i * S;
        const float* qp = Qs + i * kQs + h * D;
        const float* kp = Ks + j * kQs + h * D;
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) a = fmaf(qp[d], kp[d], a);
        Ss[(h * S + i) * (S + 1) + j] = a + __ldg(mk + h * S + j);
    }
    __syncthreads();
    // ---- softmax over keys (:414-415) -----------------------------------------
    for (int t = tid; t < H * S; t += kThreadsA) {
        float* row = Ss + t * (S + 1);
        float mx = row[0];
        for (int j = 1; j < S; ++j) mx = fmaxf(mx, row[j]);
        float sum = 0.f;
        for (int j = 0; j < S; ++j) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
        const float inv = 1.0f / sum;
        for (int j = 0; j < S; ++j) row[j] *= inv;
    }
    __syncthreads();
    // ---- P.V (:417), head-major channel concat -> token tile -------------------
    for (int t = tid; t < S * kC; t += kThreadsA) {
        const int i = t / kC, c = t - i * kC, h = c / D;
        const float* pr = Ss + (h * S + i) * (S + 1);
        float a = 0.f;
#pragma unroll 4
        for (int j = 0; j < S; ++j) a = fmaf(pr[j], Vs[j * kQs + c], a);
        Xs[i * kXs + c] = a;
    }
    __syncthreads();
    // ---- out-projection (:448) + (fused) scatter to voxel rows -----------------
    if (FUSED) {
        project<S, true>(Xs, w.w_out_t, kC, w.b_out, 1.0f, nullptr, out + (size_t) b * max_pillars * kC, s_rows, kC);
    } else {
        project<S, true>(Xs, w.w_out_t, kC, w.b_out, 1.0f, nullptr,
                         out + ((size_t) b * max_sets + set) * S * kC, nullptr, kC);
    }
}

template <int S, bool FUSED>
int launch_fp32(const dsvt_set_attention_params* p, const AttnWeightsDev& w,
                const float* q, const float* k, const float* v, const float* pos, const int* idx,
                const float* mask, const int* set_num, const int* voxel_num, float* out, cudaStream_t st)
{
    const size_t smem = (size_t) AttnSmem<S>::total * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        DSVT_CUDA(cudaFuncSetAttribute(set_attention_fp32_kernel<S, FUSED>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    set_attention_fp32_kernel<S, FUSED><<<dim3(p->max_set_num, p->batch), kThreadsA, smem, st>>>(
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
