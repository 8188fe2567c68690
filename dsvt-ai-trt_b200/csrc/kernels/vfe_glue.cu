// next #3 -- the two remaining reference plugins either side of the DSVT blocks:
//
//   TorchScatterMaxPlugin::enqueue  (reference plugins/src/torchScatterMax.cu:282-309, kernel :201-262)
//       per pillar: channel-wise max over its points' feature rows (start value -1000000.0, :214-217); the result is
//       written to the pillar row of max_voxel_features and broadcast to every point row of max_point_features.
//       The reference runs one THREAD per pillar with a 200-float local array and strided 4-byte accesses, after
//       memsetting both outputs in full (:300-301).  Here: one warp per pillar, lanes over float4 channel groups, rows
//       read and written as full 128-byte lines; only the tails are zero-filled.
//   Map2BevPlugin::enqueue          (reference plugins/src/map2bev.cu:283-312, kernel :250-265)
//       dense BEV map [gy, gx, C] = 0, then map[y, x, :] = voxel_features[v, :] for the valid pillars.  The 168 MB
//       clear is the contract (a dense map) and is done by one memset node at write bandwidth; the scatter moves
//       whole 768-byte rows with 128-bit accesses (reference: one thread per float).
#include "common.cuh"

namespace dsvt {
namespace {

constexpr float kMaxInit = -1000000.0f;      // torchScatterMax.cu:216

// warp per pillar; F4 = feature_num / 4 float4 per row handled by lanes (lane, lane + 32, ...)
template <int NF4>      // float4s per lane: ceil(F4 / 32), 1 (F <= 128) or 2 (F <= 256)
__global__ void __launch_bounds__(256)
scatter_max_kernel(const float4* __restrict__ feat, const int* __restrict__ piv, const int* __restrict__ pnv,
                   const int* __restrict__ voxel_num, const int* __restrict__ point_num, float4* __restrict__ max_point,
                   float4* __restrict__ max_voxel, int max_points, int max_pillars, int npv, int F4, int zero_tails)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    const float4* fb = feat + (size_t) b * max_points * F4;
    float4* mp = max_point + (size_t) b * max_points * F4;
    float4* mv = max_voxel + (size_t) b * max_pillars * F4;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); v < max_pillars; v += warps) {
        if (v >= V) {
            if (!zero_tails) break;
            for (int c = lane; c < F4; c += 32) stg_zero4(mv + (size_t) v * F4 + c);
            continue;
        }
        int n = pnv[(size_t) b * max_pillars + v];
        n = n < npv ? n : npv;
        // the pillar's row ids: lanes hold up to 64 of them (npv <= 64)
        const int* ids = piv + ((size_t) b * max_pillars + v) * npv;
        const int id0 = lane < n ? ids[lane] : 0, id1 = lane + 32 < n ? ids[lane + 32] : 0;
        float4 m[NF4];
#pragma unroll
        for (int k = 0; k < NF4; ++k) m[k] = make_float4(kMaxInit, kMaxInit, kMaxInit, kMaxInit);
        // kRows rows in flight per step (the pillar's rows are independent loads; only the max is a chain)
        constexpr int kRows = 8;
        for (int i0 = 0; i0 < n; i0 += kRows) {
            float4 a[kRows][NF4];
#pragma unroll
            for (int u = 0; u < kRows; ++u) {
                const int i = i0 + u < n ? i0 + u : n - 1;           // clamped: a repeated row does not change the max
                const int row = i < 32 ? __shfl_sync(0xffffffffu, id0, i) : __shfl_sync(0xffffffffu, id1, i - 32);
#pragma unroll
                for (int k = 0; k < NF4; ++k) {
                    const int c = lane + 32 * k;
                    a[u][k] = c < F4 ? ldg_stream4(fb + (size_t) row * F4 + c) : make_float4(kMaxInit, kMaxInit, kMaxInit, kMaxInit);
                }
            }
#pragma unroll
            for (int u = 0; u < kRows; ++u)
#pragma unroll
                for (int k = 0; k < NF4; ++k) {
                    // strict '>' like the reference (:230): a NaN never replaces the running maximum
                    m[k].x = a[u][k].x > m[k].x ? a[u][k].x : m[k].x; m[k].y = a[u][k].y > m[k].y ? a[u][k].y : m[k].y;
                    m[k].z = a[u][k].z > m[k].z ? a[u][k].z : m[k].z; m[k].w = a[u][k].w > m[k].w ? a[u][k].w : m[k].w;
                }
        }
#pragma unroll
        for (int k = 0; k < NF4; ++k) {
            const int c = lane + 32 * k;
            if (c < F4) mv[(size_t) v * F4 + c] = m[k];
        }
        for (int i = 0; i < n; ++i) {
            const int row = i < 32 ? __shfl_sync(0xffffffffu, id0, i) : __shfl_sync(0xffffffffu, id1, i - 32);
#pragma unroll
            for (int k = 0; k < NF4; ++k) {
                const int c = lane + 32 * k;
                if (c < F4) stg_stream4(mp + (size_t) row * F4 + c, m[k]);
            }
        }
    }
    // point rows that belong to no pillar: with point_num given they are exactly [point_num, max_points)
    if (zero_tails && point_num) {
        int Pc = point_num[b];
        Pc = Pc < max_points ? Pc : max_points;
        const size_t n4 = (size_t) (max_points - Pc) * F4;
        float4* tail = mp + (size_t) Pc * F4;
        for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (size_t) gridDim.x * blockDim.x)
            stg_zero4(tail + t);
    }
}

__global__ void __launch_bounds__(256)
map2bev_kernel(const float4* __restrict__ feat, const int4* __restrict__ coords, const int* __restrict__ voxel_num,
               float4* __restrict__ map, int max_pillars, int C4, int gx, int gy)
{
    const int b = blockIdx.y;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    const size_t n = (size_t) V * C4;
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t) gridDim.x * blockDim.x) {
        const int v = (int) (t / C4), c = (int) (t - (size_t) v * C4);
        const int4 co = __ldg(coords + (size_t) b * max_pillars + v);          // (0, 0, y, x), map2bev.cu:258-260
        const unsigned y = (unsigned) co.z, x = (unsigned) co.w;
        if (y < (unsigned) gy && x < (unsigned) gx)                            // the reference does not guard
            map[((size_t) b * gy * gx + (size_t) y * gx + x) * C4 + c] = ldg_stream4(feat + ((size_t) b * max_pillars + v) * C4 + c);
    }
}

// Alternative clear of the dense BEV map with evict-first stores (a 168 MB cudaMemset node sweeps the whole 126 MB L2);
// tried, slower than the memset node, kept behind DSVT_BEV_ZERO_KERNEL.
__global__ void __launch_bounds__(256)
zero_fill_kernel(float4* __restrict__ dst, size_t n4)
{
    for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (size_t) gridDim.x * blockDim.x)
        stg_zero4(dst + t);
}

inline int grid_cap(size_t work_items, int threads, int ctas_per_sm) {
    const size_t need = (work_items + threads - 1) / threads, cap = (size_t) sm_count() * ctas_per_sm;
    return (int) (need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace
}  // namespace dsvt

using namespace dsvt;

extern "C" int dsvt_torch_scatter_max_launch(const dsvt_torch_scatter_max_params* p, const float* point_features,
                                             const int32_t* point_index_in_voxel, const int32_t* point_num_in_voxel,
                                             const int32_t* voxel_num, const int32_t* point_num,
                                             float* max_point_features, float* max_voxel_features, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && p->batch >= 1 && p->max_points_num >= 1 && p->max_pillars_num >= 1, "params");
    DSVT_CHECK_ARG(p->feature_num >= 4 && p->feature_num % 4 == 0 && p->feature_num <= 256,
                   "feature_num must be a multiple of 4 in [4, 256]");
    DSVT_CHECK_ARG(p->max_num_points_per_voxel >= 1 && p->max_num_points_per_voxel <= 64, "max_num_points_per_voxel");
    DSVT_CHECK_ARG(point_features && point_index_in_voxel && point_num_in_voxel && voxel_num && max_point_features &&
                   max_voxel_features, "NULL tensor pointer");
    DSVT_CHECK_ARG(!(((uintptr_t) point_features | (uintptr_t) max_point_features | (uintptr_t) max_voxel_features) & 15),
                   "16-B alignment");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int F4 = p->feature_num / 4;
    if (p->zero_tails && !point_num) {
        // without the voxeliser's point count the rows that belong to no pillar are unknown: clear everything first,
        // like the reference (torchScatterMax.cu:300)
        DSVT_CUDA(cudaMemsetAsync(max_point_features, 0, (size_t) p->batch * p->max_points_num * p->feature_num * sizeof(float), st));
        count_launch();
    }
    const int grid = grid_cap((size_t) p->max_pillars_num * 32, 256, 8);
    const float4* f4 = reinterpret_cast<const float4*>(point_features);
    float4* mp4 = reinterpret_cast<float4*>(max_point_features);
    float4* mv4 = reinterpret_cast<float4*>(max_voxel_features);
    if (F4 <= 32)
        scatter_max_kernel<1><<<dim3(grid, p->batch), 256, 0, st>>>(f4, point_index_in_voxel, point_num_in_voxel, voxel_num,
                                                                   point_num, mp4, mv4, p->max_points_num, p->max_pillars_num,
                                                                   p->max_num_points_per_voxel, F4, p->zero_tails);
    else
        scatter_max_kernel<2><<<dim3(grid, p->batch), 256, 0, st>>>(f4, point_index_in_voxel, point_num_in_voxel, voxel_num,
                                                                   point_num, mp4, mv4, p->max_points_num, p->max_pillars_num,
                                                                   p->max_num_points_per_voxel, F4, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

extern "C" int dsvt_map2bev_launch(const dsvt_map2bev_params* p, const float* voxel_features, const int32_t* coords,
                                   const int32_t* voxel_num, float* map_features, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && p->batch >= 1 && p->max_pillars_num >= 1 && p->grid_size_x >= 1 && p->grid_size_y >= 1, "params");
    DSVT_CHECK_ARG(p->channel_num >= 4 && p->channel_num % 4 == 0, "channel_num must be a multiple of 4");
    DSVT_CHECK_ARG(voxel_features && coords && voxel_num && map_features, "NULL tensor pointer");
    DSVT_CHECK_ARG(!(((uintptr_t) voxel_features | (uintptr_t) coords | (uintptr_t) map_features) & 15), "16-B alignment");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // the dense map is the contract (map2bev.cu:305): cleared at write bandwidth
    const int C4 = p->channel_num / 4;
    const size_t map4 = (size_t) p->batch * p->grid_size_x * p->grid_size_y * C4;
#ifdef DSVT_BEV_ZERO_KERNEL      // measured: 43.7 us vs 39.6 us for the memset node, frames/s -0.4 %: not the default
    zero_fill_kernel<<<grid_cap(map4, 256, 8), 256, 0, st>>>(reinterpret_cast<float4*>(map_features), map4);
    DSVT_LAUNCH_CHECK();
#else
    DSVT_CUDA(cudaMemsetAsync(map_features, 0, map4 * sizeof(float4), st));
    count_launch();
#endif
    const int grid = grid_cap((size_t) p->max_pillars_num * C4, 256, 8);
    map2bev_kernel<<<dim3(grid, p->batch), 256, 0, st>>>(reinterpret_cast<const float4*>(voxel_features),
                                                         reinterpret_cast<const int4*>(coords), voxel_num,
                                                         reinterpret_cast<float4*>(map_features), p->max_pillars_num, C4,
                                                         p->grid_size_x, p->grid_size_y);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}
