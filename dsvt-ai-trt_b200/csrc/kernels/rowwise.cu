// a4 GELU, a5 LayerNorm, a6 filterBoxByScore -- the HBM-bound row-wise plugins.
//   GeluPlugin::enqueue             reference plugins/src/gelu.cu:227-251
//   LayerNormPlugin::enqueue        reference plugins/src/layerNorm.cu:357-402
//   FilterBoxByScorePlugin::enqueue reference plugins/src/filterBoxByScore.cu:328-379
// All three are pure streaming kernels: 128-bit coalesced loads/stores, no shared-memory staging
// (no reuse), grid sized as a multiple of the SM count, valid counts read from device memory.
#include "common.cuh"

namespace dsvt {
namespace {

// ---------------------------------------------------------------------------
// GELU (tanh form, params.h:75-77).  The reference evaluates it in double (gelu.cu:210); we use
// the algebraically identical logistic form 0.5+0.5*tanh(u) = 1/(1+exp(-2u)) in f32, which stays
// within ~2e-7 relative of the double result.
__device__ __forceinline__ float gelu_f(float x) {
    const float kB = 0.7978845608028654f, kC = 0.035677408136300125f;
    const float u = x * fmaf(kC * x, x, kB);
    const float e = __expf(-2.0f * u);
    return __fdividef(x, 1.0f + e);
}

__global__ void __launch_bounds__(256)
gelu_kernel(const float4* __restrict__ in, const int* __restrict__ voxel_num, float4* __restrict__ out,
            int max_pillars, int C4 /*channels / 4*/, int zero_tails)
{
    const int b = blockIdx.y;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    const size_t frame = (size_t) max_pillars * C4;
    const size_t valid = (size_t) V * C4;
    const size_t end = zero_tails ? frame : valid;
    const float4* src = in + (size_t) b * frame;
    float4* dst = out + (size_t) b * frame;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    // two independent 16-B loads in flight per thread
    for (; i + stride < valid; i += 2 * stride) {
        const float4 a = ldg_stream4(src + i);
        const float4 c = ldg_stream4(src + i + stride);
        stg_stream4(dst + i, make_float4(gelu_f(a.x), gelu_f(a.y), gelu_f(a.z), gelu_f(a.w)));
        stg_stream4(dst + i + stride, make_float4(gelu_f(c.x), gelu_f(c.y), gelu_f(c.z), gelu_f(c.w)));
    }
    for (; i < end; i += stride) {
        if (i < valid) {
            const float4 a = ldg_stream4(src + i);
            stg_stream4(dst + i, make_float4(gelu_f(a.x), gelu_f(a.y), gelu_f(a.z), gelu_f(a.w)));
        } else {
            stg_zero4(dst + i);
        }
    }
}

__global__ void __launch_bounds__(256)
gelu_scalar_kernel(const float* __restrict__ in, const int* __restrict__ voxel_num, float* __restrict__ out,
                   int max_pillars, int C, int zero_tails)
{
    const int b = blockIdx.y;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    const size_t frame = (size_t) max_pillars * C, valid = (size_t) V * C;
    const size_t end = zero_tails ? frame : valid;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (size_t) gridDim.x * blockDim.x)
        out[(size_t) b * frame + i] = i < valid ? gelu_f(in[(size_t) b * frame + i]) : 0.f;
}

// ---------------------------------------------------------------------------
// LayerNorm, C = 192 fast path: HALF a warp per row, three float4 per lane (48 float4 = 192 ch),
// two-pass in registers: mean, then sum((x-mean)^2) -- the reference's formulation
// (layerNorm.cu:297-338) -- then (x-mean)/sqrtf(var+eps)*gamma+beta (:272-276).
template <bool HAS_RES>
__global__ void __launch_bounds__(256)
layer_norm192_kernel(const float4* __restrict__ x, const float4* __restrict__ res,
                     const int* __restrict__ voxel_num, const float4* __restrict__ gamma,
                     const float4* __restrict__ beta, float4* __restrict__ out,
                     int max_pillars, float eps, int zero_tails)
{
    const int b = blockIdx.y;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    const int sub = threadIdx.x & 15;                 // lane within the half-warp
    // the two half-warps of a warp own different rows and diverge at the valid-count / end-row boundary: the reductions
    // name only the 16 lanes of the own half
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);
    const int rows_per_block = blockDim.x >> 4;
    const int end_row = zero_tails ? max_pillars : V;
    float4 g[3], be[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { g[k] = __ldg(gamma + k * 16 + sub); be[k] = __ldg(beta + k * 16 + sub); }

    for (int row = blockIdx.x * rows_per_block + (threadIdx.x >> 4); row < end_row;
         row += gridDim.x * rows_per_block) {
        const size_t off = ((size_t) b * max_pillars + row) * 48;
        float4 v[3];
        if (row < V) {
#pragma unroll
            for (int k = 0; k < 3; ++k) v[k] = ldg_stream4(x + off + k * 16 + sub);
            if (HAS_RES) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 r = ldg_stream4(res + off + k * 16 + sub);
                    v[k].x += r.x; v[k].y += r.y; v[k].z += r.z; v[k].w += r.w;
                }
            }
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(hmask, s, o);
            const float mean = s / 192.f;
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float a = v[k].x - mean, c = v[k].y - mean, d = v[k].z - mean, e = v[k].w - mean;
                q += (a * a + c * c) + (d * d + e * e);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(hmask, q, o);
            const float var = q / 192.f;
            const float sd = sqrtf(var + eps);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float4 r;
                r.x = (v[k].x - mean) / sd * g[k].x + be[k].x;
                r.y = (v[k].y - mean) / sd * g[k].y + be[k].y;
                r.z = (v[k].z - mean) / sd * g[k].z + be[k].z;
                r.w = (v[k].w - mean) / sd * g[k].w + be[k].w;
                stg_stream4(out + off + k * 16 + sub, r);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) stg_zero4(out + off + k * 16 + sub);
        }
    }
}

// Chain of N LayerNorms (C = 192) with residual adds in between; same arithmetic per stage as layer_norm192_kernel.
struct LnChain {
    const float4* res[3];
    const float4* gamma[3];
    const float4* beta[3];
};
template <int N>
__global__ void __launch_bounds__(256)
layer_norm192_chain_kernel(const float4* __restrict__ x, LnChain ch, const int* __restrict__ voxel_num,
                           float4* __restrict__ out, int max_pillars, float eps, int zero_tails)
{
    const int b = blockIdx.y;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    const int sub = threadIdx.x & 15;
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);     // reductions stay inside the own half-warp (see above)
    const int rows_per_block = blockDim.x >> 4;
    const int end_row = zero_tails ? max_pillars : V;
    for (int row = blockIdx.x * rows_per_block + (threadIdx.x >> 4); row < end_row;
         row += gridDim.x * rows_per_block) {
        const size_t off = ((size_t) b * max_pillars + row) * 48;
        if (row >= V) {
#pragma unroll
            for (int k = 0; k < 3; ++k) stg_zero4(out + off + k * 16 + sub);
            continue;
        }
        float4 v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = ldg_stream4(x + off + k * 16 + sub);
#pragma unroll
        for (int st = 0; st < N; ++st) {
            if (ch.res[st] != nullptr) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 r = ldg_stream4(ch.res[st] + off + k * 16 + sub);
                    v[k].x += r.x; v[k].y += r.y; v[k].z += r.z; v[k].w += r.w;
                }
            }
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(hmask, s, o);
            const float mean = s / 192.f;
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float a = v[k].x - mean, c = v[k].y - mean, d = v[k].z - mean, e = v[k].w - mean;
                q += (a * a + c * c) + (d * d + e * e);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(hmask, q, o);
            const float sd = sqrtf(q / 192.f + eps);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 g = __ldg(ch.gamma[st] + k * 16 + sub), be = __ldg(ch.beta[st] + k * 16 + sub);
                v[k].x = (v[k].x - mean) / sd * g.x + be.x;
                v[k].y = (v[k].y - mean) / sd * g.y + be.y;
                v[k].z = (v[k].z - mean) / sd * g.z + be.z;
                v[k].w = (v[k].w - mean) / sd * g.w + be.w;
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) stg_stream4(out + off + k * 16 + sub, v[k]);
    }
}

// generic channel count: one warp per row, scalar accesses
__global__ void __launch_bounds__(256)
layer_norm_generic_kernel(const float* __restrict__ x, const float* __restrict__ res,
                          const int* __restrict__ voxel_num, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ out,
                          int max_pillars, int C, float eps, int zero_tails)
{
    const int b = blockIdx.y;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    const int lane = threadIdx.x & 31;
    const int rows_per_block = blockDim.x >> 5;
    const int end_row = zero_tails ? max_pillars : V;
    for (int row = blockIdx.x * rows_per_block + (threadIdx.x >> 5); row < end_row;
         row += gridDim.x * rows_per_block) {
        const size_t off = ((size_t) b * max_pillars + row) * C;
        if (row >= V) {
            for (int c = lane; c < C; c += 32) out[off + c] = 0.f;
            continue;
        }
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += x[off + c] + (res ? res[off + c] : 0.f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / C;
        float q = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float d = x[off + c] + (res ? res[off + c] : 0.f) - mean;
            q += d * d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float sd = sqrtf(q / C + eps);
        for (int c = lane; c < C; c += 32) {
            const float d = x[off + c] + (res ? res[off + c] : 0.f) - mean;
            out[off + c] = d / sd * gamma[c] + beta[c];
        }
    }
}

// ---------------------------------------------------------------------------
// filterBoxByScore: one CTA per frame, ordered (stable) compaction via ballot + warp prefix.
// Unlike the reference (filterBoxByScore.cu:319 launches 512 threads over 500 candidates with no
// guard) every access is bounded by max_top_k.
__global__ void __launch_bounds__(1024)
filter_box_kernel(const float* __restrict__ scores, const int* __restrict__ classes, const int* __restrict__ xs,
                  const int* __restrict__ ys, const float* __restrict__ center, const float* __restrict__ center_z,
                  const float* __restrict__ angle, const float* __restrict__ dim,
                  float* __restrict__ boxes, int* __restrict__ valid, int K,
                  float x_min, float x_max, float y_min, float y_max, float z_min, float z_max,
                  float vx, float vy, float score_thr, int zero_tails)
{
    __shared__ int warp_cnt[32];
    __shared__ int s_base;
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    scores += (size_t) b * K; classes += (size_t) b * K; xs += (size_t) b * K; ys += (size_t) b * K;
    center += (size_t) b * K * 2; center_z += (size_t) b * K; angle += (size_t) b * K; dim += (size_t) b * K * 3;
    boxes += (size_t) b * K * 9;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < K; c0 += blockDim.x) {
        const int i = c0 + threadIdx.x;
        bool keep = false;
        float nx = 0.f, ny = 0.f, cz = 0.f, sc = 0.f;
        if (i < K) {
            sc = scores[i];
            // (float)unsigned + offset, then *voxel + min (contracted to an FMA by nvcc in the reference too)
            nx = (float) (unsigned) xs[i] + center[i * 2 + 0];
            ny = (float) (unsigned) ys[i] + center[i * 2 + 1];
            nx = nx * vx + x_min;
            ny = ny * vy + y_min;
            cz = center_z[i];
            keep = (nx >= x_min && nx < x_max && ny >= y_min && ny < y_max && cz >= z_min && cz < z_max) &&
                   (sc >= score_thr);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int before = s_base, total = 0;
        for (int w = 0; w < nw; ++w) { const int c = warp_cnt[w]; if (w < wid) before += c; total += c; }
        if (keep) {
            float* o = boxes + (size_t) (before + __popc(bal & ((1u << lane) - 1))) * 9;
            o[0] = nx; o[1] = ny; o[2] = cz;
            o[3] = dim[i * 3 + 0]; o[4] = dim[i * 3 + 1]; o[5] = dim[i * 3 + 2];
            o[6] = angle[i];
            o[7] = (float) (unsigned) classes[i];
            o[8] = sc;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
    const int nvalid = s_base;
    if (threadIdx.x == 0) valid[b] = nvalid;
    if (zero_tails)
        for (int t = nvalid * 9 + threadIdx.x; t < K * 9; t += blockDim.x) boxes[t] = 0.f;
}

// ---------------------------------------------------------------------------
// standalone gather / scatter plugins around the set attention (next #2)
__global__ void __launch_bounds__(256)
get_value_by_index_kernel(const float4* __restrict__ x, const float4* __restrict__ pos, const int* __restrict__ idx,
                          const int* __restrict__ set_num, float4* __restrict__ q, float4* __restrict__ k,
                          float4* __restrict__ v, int max_sets, int S, int C4, int max_pillars, int axis,
                          int zero_tails)
{
    const int b = blockIdx.y;
    int ns = set_num[b];
    ns = ns < max_sets ? ns : max_sets;
    const size_t valid = (size_t) ns * S * C4, frame = (size_t) max_sets * S * C4;
    const size_t end = zero_tails ? frame : valid;
    const int* id = idx + ((size_t) b * 2 + axis) * max_sets * S;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (size_t) gridDim.x * blockDim.x) {
        float4 qq = make_float4(0.f, 0.f, 0.f, 0.f), vv = qq;
        if (i < valid) {
            const int tok = (int) (i / C4), c = (int) (i - (size_t) tok * C4);
            const int g = id[tok];
            const size_t src = ((size_t) b * max_pillars + g) * C4 + c;
            vv = __ldg(x + src);
            const float4 pp = __ldg(pos + src);
            qq = make_float4(vv.x + pp.x, vv.y + pp.y, vv.z + pp.z, vv.w + pp.w);   // getValueByIndex.cu:298-300
        }
        q[(size_t) b * frame + i] = qq; k[(size_t) b * frame + i] = qq; v[(size_t) b * frame + i] = vv;
    }
}

__global__ void __launch_bounds__(256)
map_set_feature2voxel_kernel(const float4* __restrict__ feat, const int* __restrict__ idx,
                             const int* __restrict__ set_num, float4* __restrict__ out,
                             int max_sets, int S, int C4, int max_pillars, int axis)
{
    const int b = blockIdx.y;
    int ns = set_num[b];
    ns = ns < max_sets ? ns : max_sets;
    const size_t valid = (size_t) ns * S * C4, frame = (size_t) max_sets * S * C4;
    const int* id = idx + ((size_t) b * 2 + axis) * max_sets * S;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < valid; i += (size_t) gridDim.x * blockDim.x) {
        const int tok = (int) (i / C4), c = (int) (i - (size_t) tok * C4);
        const int g = id[tok];
        out[((size_t) b * max_pillars + g) * C4 + c] = feat[(size_t) b * frame + i];   // mapSetFeature2voxel.cu:273
    }
}

inline int grid_for(size_t work_items, int threads, int ctas_per_sm) {
    const int sms = sm_count();
    size_t need = (work_items + threads - 1) / threads;
    size_t cap = (size_t) sms * ctas_per_sm;
    if (need < 1) need = 1;
    if (need >= cap) return (int) cap;
    // round up to a multiple of the SM count when that does not exceed the need by much
    return (int) need;
}

}  // namespace
}  // namespace dsvt

using namespace dsvt;

extern "C" int dsvt_gelu_launch(const dsvt_gelu_params* p, const float* x, const int32_t* voxel_num, float* out,
                                dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && p->batch >= 1 && p->max_pillars_num >= 1 && p->channel_num >= 1, "params");
    DSVT_CHECK_ARG(x && voxel_num && out, "NULL tensor pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t frame = (size_t) p->max_pillars_num * p->channel_num;
    if (p->channel_num % 4 == 0 && ((uintptr_t) x & 15) == 0 && ((uintptr_t) out & 15) == 0) {
        const int grid = grid_for(frame / 4 / 2, 256, 8);
        gelu_kernel<<<dim3(grid, p->batch), 256, 0, st>>>(reinterpret_cast<const float4*>(x), voxel_num,
                                                          reinterpret_cast<float4*>(out), p->max_pillars_num,
                                                          p->channel_num / 4, p->zero_tails);
    } else {
        const int grid = grid_for(frame, 256, 8);
        gelu_scalar_kernel<<<dim3(grid, p->batch), 256, 0, st>>>(x, voxel_num, out, p->max_pillars_num,
                                                                 p->channel_num, p->zero_tails);
    }
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

extern "C" int dsvt_layer_norm_launch(const dsvt_layer_norm_params* p, const float* x, const float* residual,
                                      const int32_t* voxel_num, const float* gamma, const float* beta, float* out,
                                      dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && p->batch >= 1 && p->max_pillars_num >= 1 && p->channel_num >= 1, "params");
    DSVT_CHECK_ARG(x && voxel_num && gamma && beta && out, "NULL tensor pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool aligned = !(((uintptr_t) x | (uintptr_t) out | (uintptr_t) gamma | (uintptr_t) beta |
                            (uintptr_t) residual) & 15);
    if (p->channel_num == 192 && aligned) {
        const int rows_per_block = 16;
        const int grid = grid_for((size_t) p->max_pillars_num * 16, 256, 8);
        (void) rows_per_block;
        if (residual) {
            layer_norm192_kernel<true><<<dim3(grid, p->batch), 256, 0, st>>>(
                reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(residual), voxel_num,
                reinterpret_cast<const float4*>(gamma), reinterpret_cast<const float4*>(beta),
                reinterpret_cast<float4*>(out), p->max_pillars_num, p->eps, p->zero_tails);
        } else {
            layer_norm192_kernel<false><<<dim3(grid, p->batch), 256, 0, st>>>(
                reinterpret_cast<const float4*>(x), nullptr, voxel_num,
                reinterpret_cast<const float4*>(gamma), reinterpret_cast<const float4*>(beta),
                reinterpret_cast<float4*>(out), p->max_pillars_num, p->eps, p->zero_tails);
        }
    } else {
        const int grid = grid_for((size_t) p->max_pillars_num * 32, 256, 8);
        layer_norm_generic_kernel<<<dim3(grid, p->batch), 256, 0, st>>>(x, residual, voxel_num, gamma, beta, out,
                                                                        p->max_pillars_num, p->channel_num,
                                                                        p->eps, p->zero_tails);
    }
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

extern "C" int dsvt_layer_norm_chain_launch(const dsvt_layer_norm_params* p, const float* x, const int32_t* voxel_num,
                                            const dsvt_ln_stage* stages, int32_t n_stages, float* out,
                                            dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && p->batch >= 1 && p->max_pillars_num >= 1, "params");
    DSVT_CHECK_ARG(p->channel_num == 192, "the fused chain is built for channel_num = 192");
    DSVT_CHECK_ARG(x && voxel_num && stages && out && n_stages >= 1 && n_stages <= 3, "arguments");
    LnChain ch{};
    uintptr_t align = (uintptr_t) x | (uintptr_t) out;
    for (int i = 0; i < n_stages; ++i) {
        DSVT_CHECK_ARG(stages[i].gamma && stages[i].beta, "stage gamma/beta is NULL");
        ch.res[i] = reinterpret_cast<const float4*>(stages[i].residual);
        ch.gamma[i] = reinterpret_cast<const float4*>(stages[i].gamma);
        ch.beta[i] = reinterpret_cast<const float4*>(stages[i].beta);
        align |= (uintptr_t) stages[i].residual | (uintptr_t) stages[i].gamma | (uintptr_t) stages[i].beta;
    }
    DSVT_CHECK_ARG(!(align & 15), "16-B alignment");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for((size_t) p->max_pillars_num * 16, 256, 8);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* o4 = reinterpret_cast<float4*>(out);
    switch (n_stages) {
        case 1: layer_norm192_chain_kernel<1><<<dim3(grid, p->batch), 256, 0, st>>>(x4, ch, voxel_num, o4, p->max_pillars_num, p->eps, p->zero_tails); break;
        case 2: layer_norm192_chain_kernel<2><<<dim3(grid, p->batch), 256, 0, st>>>(x4, ch, voxel_num, o4, p->max_pillars_num, p->eps, p->zero_tails); break;
        default: layer_norm192_chain_kernel<3><<<dim3(grid, p->batch), 256, 0, st>>>(x4, ch, voxel_num, o4, p->max_pillars_num, p->eps, p->zero_tails); break;
    }
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

extern "C" int dsvt_filter_box_launch(const dsvt_filter_box_params* p,
                                      const float* scores, const int32_t* classes, const int32_t* xs,
                                      const int32_t* ys, const float* center, const float* center_z,
                                      const float* angle, const float* dim, float* boxes, int32_t* valid,
                                      dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && p->batch >= 1 && p->max_top_k >= 1, "params");
    DSVT_CHECK_ARG(scores && classes && xs && ys && center && center_z && angle && dim && boxes && valid,
                   "NULL tensor pointer");
    int threads = ((p->max_top_k + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    filter_box_kernel<<<p->batch, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        scores, classes, xs, ys, center, center_z, angle, dim, boxes, valid, p->max_top_k,
        p->x_min, p->x_max, p->y_min, p->y_max, p->z_min, p->z_max, p->voxel_x, p->voxel_y,
        p->score_threshold, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

extern "C" int dsvt_get_value_by_index_launch(const dsvt_set_attention_params* p, const float* x, const float* pos,
                                              const int32_t* global_index_in_set, const int32_t* set_num,
                                              float* q, float* k, float* v, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && p->batch >= 1 && p->max_set_num >= 1 && p->voxel_num_set >= 1 && p->max_pillars_num >= 1,
                   "params");
    DSVT_CHECK_ARG(p->channel_num % 4 == 0, "channel_num must be a multiple of 4");
    DSVT_CHECK_ARG(p->axis_id == 0 || p->axis_id == 1, "axis_id");
    DSVT_CHECK_ARG(x && pos && global_index_in_set && set_num && q && k && v, "NULL tensor pointer");
    const size_t frame = (size_t) p->max_set_num * p->voxel_num_set * (p->channel_num / 4);
    const int grid = grid_for(frame, 256, 8);
    get_value_by_index_kernel<<<dim3(grid, p->batch), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(pos), global_index_in_set, set_num,
        reinterpret_cast<float4*>(q), reinterpret_cast<float4*>(k), reinterpret_cast<float4*>(v),
        p->max_set_num, p->voxel_num_set, p->channel_num / 4, p->max_pillars_num, p->axis_id, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

extern "C" int dsvt_map_set_feature2voxel_launch(const dsvt_set_attention_params* p, const float* set_features,
                                                 const int32_t* global_index_in_set, const int32_t* set_num,
                                                 float* voxel_features, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && p->batch >= 1 && p->max_set_num >= 1 && p->voxel_num_set >= 1 && p->max_pillars_num >= 1,
                   "params");
    DSVT_CHECK_ARG(p->channel_num % 4 == 0, "channel_num must be a multiple of 4");
    DSVT_CHECK_ARG(p->axis_id == 0 || p->axis_id == 1, "axis_id");
    DSVT_CHECK_ARG(set_features && global_index_in_set && set_num && voxel_features, "NULL tensor pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p->zero_tails) {   // the reference clears the whole voxel tensor first (mapSetFeature2voxel.cu:312)
        DSVT_CUDA(cudaMemsetAsync(voxel_features, 0,
                                  (size_t) p->batch * p->max_pillars_num * p->channel_num * sizeof(float), st));
        count_launch();
    }
    const size_t frame = (size_t) p->max_set_num * p->voxel_num_set * (p->channel_num / 4);
    const int grid = grid_for(frame, 256, 8);
    map_set_feature2voxel_kernel<<<dim3(grid, p->batch), 256, 0, st>>>(
        reinterpret_cast<const float4*>(set_features), global_index_in_set, set_num,
        reinterpret_cast<float4*>(voxel_features), p->max_set_num, p->voxel_num_set, p->channel_num / 4,
        p->max_pillars_num, p->axis_id);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}
