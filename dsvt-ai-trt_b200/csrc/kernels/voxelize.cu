// a1 -- pillar voxelisation + 10-channel point decoration.
// Replaces Points2FeaturesPlugin::enqueue (reference plugins/src/points2Features.cu:896-990).
//
// B200-first design (not the reference's dense 168 MB scatter grid):
//   count    : one thread per point, float4 streaming load, f32 cell id, RED.ADD into a
//              876 KB (468x468) L2-resident counter grid; the cell id is parked in `pcell`.
//   scan     : two-kernel (reduce, then scan) exclusive scan over the cell grid -> canonical
//              pillar ids (ascending y*gx+x), list bases and output-row bases.  No spin-waits.
//   scatter  : counting-sort step: each in-range point claims a slot in its cell's segment.
//   rank     : one THREAD per point: its slot is the number of smaller input indices in its pillar's segment
//              (segments are ~10 points, the scan stops once `npv` smaller ones are seen); the lowest `npv` get an
//              output row.  Pillars with more than 64 points go through a warp-per-pillar radix-select instead.
//   mean     : one thread per pillar: sequential-order f32 sum of the kept points, exactly like the reference.
//   feat     : one thread per output row: the 10-channel decoration, written through shared memory as full lines.
//   (The first version ran one WARP per pillar for all of this: ~1160 warp instructions per pillar at 31 % lane
//    utilisation made the voxeliser issue-bound at 6 % of the HBM roofline on batched frames.)
// Results are deterministic: the reference's three atomicAdd races (:697, :751, :829) are
// replaced by their canonical serial-order outcome.
#include "common.cuh"

namespace dsvt {
namespace {

constexpr int kThreads = 256;
constexpr int kCellsPerThread = 4;
constexpr int kCellsPerBlock = kThreads * kCellsPerThread;  // 1024
constexpr int kMaxNpv = 64;
constexpr int kBigN = kMaxNpv;   // pillars with more points are ranked by vox_rank_big_kernel (radix select)

struct VoxGeom {
    float x_min, x_max, y_min, y_max, z_min, z_max;
    float vx, vy, vz;
    int gx, gy;
};

struct VoxWs {  // per-frame workspace pointers (frame stride applied by the kernels)
    unsigned int* cell_count;  // [G]
    unsigned int* cursor;      // [G]
    int* pcell;                // [maxP]
    int* list;                 // [maxP]
    int* pillar_base;          // [maxV]
    int* pillar_n;             // [maxV]
    int* pillar_row;           // [maxV]
    int* block_sums;           // [nblk*3]
    size_t frame_stride_u32;   // all sub-buffers are carved per frame with this many 4-byte words
};

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
vox_count_kernel(const float4* __restrict__ points, const int* __restrict__ points_size, int max_points,
                 VoxGeom g, unsigned int* __restrict__ cell_count, int* __restrict__ pcell,
                 size_t ws_stride, int G)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = points_size[b];
    n = n < max_points ? n : max_points;
    if (i >= n) return;
    const float4 p = ldg_stream4(points + (size_t) b * max_points + i);
    int cell = -1;
    // same predicate and f32 arithmetic as points2Features.cu:683-690
    if (!(p.x < g.x_min || p.x >= g.x_max || p.y < g.y_min || p.y >= g.y_max || p.z < g.z_min || p.z >= g.z_max)) {
        const int ix = (int) floorf((p.x - g.x_min) / g.vx);
        const int iy = (int) floorf((p.y - g.y_min) / g.vy);
        // the reference does not clamp; an index == grid size can only arise from f32 rounding at the
        // upper edge and would write out of bounds there.  We drop such a point.
        if (ix >= 0 && ix < g.gx && iy >= 0 && iy < g.gy) {
            cell = iy * g.gx + ix;
            atomicAdd(cell_count + (size_t) b * ws_stride + cell, 1u);
        }
    }
    pcell[(size_t) b * ws_stride + i] = cell;
    (void) G;
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
vox_block_reduce_kernel(const unsigned int* __restrict__ cell_count, int* __restrict__ block_sums,
                        size_t ws_stride, int G, int npv)
{
    __shared__ int red[3][kThreads / 32];
    const int b = blockIdx.y;
    const unsigned int* cc = cell_count + (size_t) b * ws_stride;
    const int c0 = blockIdx.x * kCellsPerBlock + threadIdx.x * kCellsPerThread;
    int ne = 0, sa = 0, sk = 0;
    if (c0 + kCellsPerThread <= G) {
        const uint4 v = *reinterpret_cast<const uint4*>(cc + c0);
        const unsigned int a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { ne += a[j] > 0; sa += (a[j] + 3u) & ~3u; sk += min((int) a[j], npv); }
    } else {
        for (int j = 0; j < kCellsPerThread; ++j) {
            if (c0 + j < G) { unsigned int a = cc[c0 + j]; ne += a > 0; sa += (a + 3u) & ~3u; sk += min((int) a, npv); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ne += __shfl_xor_sync(0xffffffffu, ne, o);
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sk += __shfl_xor_sync(0xffffffffu, sk, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = ne; red[1][wid] = sa; red[2][wid] = sk; }
    __syncthreads();
    if (threadIdx.x < 3) {
        int s = 0;
        for (int w = 0; w < kThreads / 32; ++w) s += red[threadIdx.x][w];
        block_sums[((size_t) b * ws_stride) + blockIdx.x * 3 + threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
vox_scan_kernel(const unsigned int* __restrict__ cell_count, const int* __restrict__ block_sums,
                unsigned int* __restrict__ cursor, int* __restrict__ pillar_base /* int4 records */,
                int* __restrict__ coords, int* __restrict__ point_num_in_voxel,
                int* __restrict__ pillar_num, int* __restrict__ point_num,
                int* __restrict__ cell_pid, int* __restrict__ big_list, unsigned int* __restrict__ big_count_base,
                int* __restrict__ list, size_t ws_stride, int G, int gx, int npv, int max_pillars, int max_rows)
{
    __shared__ int warp_sums[33];
    __shared__ int prefix[3];
    const int b = blockIdx.y;
    const unsigned int* cc = cell_count + (size_t) b * ws_stride;
    const int* bs = block_sums + (size_t) b * ws_stride;

    // exclusive prefix over the preceding blocks' aggregates (<= ~1k values, warp 0)
    if (threadIdx.x < 32) {
        int p0 = 0, p1 = 0, p2 = 0;
        for (int k = threadIdx.x; k < (int) blockIdx.x; k += 32) {
            p0 += bs[k * 3 + 0]; p1 += bs[k * 3 + 1]; p2 += bs[k * 3 + 2];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            p0 += __shfl_xor_sync(0xffffffffu, p0, o);
            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
            p2 += __shfl_xor_sync(0xffffffffu, p2, o);
        }
        if (threadIdx.x == 0) { prefix[0] = p0; prefix[1] = p1; prefix[2] = p2; }
    }
    __syncthreads();

    const int c0 = blockIdx.x * kCellsPerBlock + threadIdx.x * kCellsPerThread;
    unsigned int a[kCellsPerThread];
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j) a[j] = (c0 + j < G) ? cc[c0 + j] : 0u;
    int ne = 0, sa = 0, sk = 0;
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j) { ne += a[j] > 0; sa += (a[j] + 3u) & ~3u; sk += min((int) a[j], npv); }
    int tot0, tot1, tot2;
    int e0 = block_excl_scan(ne, warp_sums, &tot0) + prefix[0];
    int e1 = block_excl_scan(sa, warp_sums, &tot1) + prefix[1];
    int e2 = block_excl_scan(sk, warp_sums, &tot2) + prefix[2];

    int* coords_b = coords + (size_t) b * max_pillars * 4;
    int* pnv_b = point_num_in_voxel + (size_t) b * max_pillars;
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j) {
        const int cell = c0 + j;
        if (cell < G && a[j] > 0) {
            const int pid = e0;
            cursor[(size_t) b * ws_stride + cell] = (unsigned int) e1;
            cell_pid[(size_t) b * ws_stride + cell] = pid < max_pillars ? pid : -1;
            if (pid < max_pillars && (int) a[j] > kBigN)        // ranked by a warp (vox_rank_big_kernel)
                big_list[(size_t) b * ws_stride + atomicAdd(big_count_base + (size_t) b * ws_stride + G, 1u)] = pid;
            if (pid < max_pillars) {
                int keep = min((int) a[j], npv);
                keep = max(0, min(keep, max_rows - e2));  // row-capacity guard (reference has none, SURVEY A-5)
                // one 16-byte record per pillar: list base, point count, first output row, rows kept
                reinterpret_cast<int4*>(pillar_base)[((size_t) b * ws_stride >> 2) + pid] = make_int4(e1, (int) a[j], e2, keep);
                const int y = cell / gx, x = cell - y * gx;
                *reinterpret_cast<int4*>(coords_b + (size_t) pid * 4) = make_int4(0, 0, y, x);  // :755
                pnv_b[pid] = keep;
            } else if (pid == max_pillars) {
                // pillars beyond the capacity are the highest cells, so their rows all lie behind this base:
                // the emitted rows are exactly [0, e2)
                point_num[b] = min(e2, max_rows);
            }
            // list segments start on 16-byte boundaries (vox_rank_kernel reads them as int4); the <= 3 pad entries
            // compare greater than every input index
            for (int q = (int) a[j]; q < (((int) a[j] + 3) & ~3); ++q) list[(size_t) b * ws_stride + e1 + q] = 0x7fffffff;
            e0 += 1; e1 += ((int) a[j] + 3) & ~3; e2 += min((int) a[j], npv);
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        pillar_num[b] = min(prefix[0] + tot0, max_pillars);
        if (prefix[0] + tot0 <= max_pillars) point_num[b] = min(prefix[2] + tot2, max_rows);
    }
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ void
vox_scatter_body(const int* __restrict__ points_size, int max_points, int* __restrict__ pcell,
                 const int* __restrict__ cell_pid, unsigned int* __restrict__ cursor, int* __restrict__ list,
                 size_t ws_stride)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = points_size[b];
    n = n < max_points ? n : max_points;
    if (i >= n) return;
    const int cell = pcell[(size_t) b * ws_stride + i];
    if (cell < 0) return;
    const unsigned int pos = atomicAdd(cursor + (size_t) b * ws_stride + cell, 1u);
    list[(size_t) b * ws_stride + pos] = i;
    pcell[(size_t) b * ws_stride + i] = cell_pid[(size_t) b * ws_stride + cell];     // cell -> pillar id (-1: beyond capacity)
}

// ---------------------------------------------------------------------------
// One thread per input point: slot = number of smaller input indices among the pillar's points.
__device__ __forceinline__ void
vox_rank_body(const int* __restrict__ points_size, int max_points, const int* __restrict__ ppid,
              const int* __restrict__ list, const int4* __restrict__ pinfo, int2* __restrict__ row_info,
              size_t ws_stride)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n_pts = points_size[b];
    n_pts = n_pts < max_points ? n_pts : max_points;
    if (i >= n_pts) return;
    const size_t wb = (size_t) b * ws_stride;
    const int pid = ppid[wb + i];                         // written by vox_scatter_kernel; < 0: dropped point
    if (pid < 0) return;
    const int4 pi = __ldg(pinfo + (wb >> 2) + pid);       // list base, n, first row, keep = min(n, npv) after the capacity guard
    const int n = pi.y, keep = pi.w;
    if (n > kBigN) return;                                // ranked by vox_rank_big_kernel
    const int4* lst = reinterpret_cast<const int4*>(list + wb + pi.x);    // padded to a multiple of 4
    int rank = 0;
    for (int j = 0; j < n && rank < keep; j += 4) {       // a point with `keep` smaller indices before it is dropped
        const int4 v = __ldg(lst + (j >> 2));
        rank += (v.x < i) + (v.y < i) + (v.z < i) + (v.w < i);
    }
    if (rank < keep) row_info[(wb >> 1) + pi.z + rank] = make_int2(i, pid);
}

// One warp per pillar with more than kBigN points (tens to hundreds per frame; thousands of points each near the
// sensor): radix-select the `npv` lowest input indices (8-bit digits, per-warp shared-memory histogram), rank-sort them.
template <int WARPS>
__device__ __forceinline__ void
vox_rank_big_body(int cta, int n_ctas, int max_points, const int* __restrict__ list, const int4* __restrict__ pinfo,
                  const int* __restrict__ big_list, const unsigned int* __restrict__ big_count_base,
                  int2* __restrict__ row_info, size_t ws_stride, int G, int npv)
{
    __shared__ int s_idx[WARPS][kMaxNpv];
    __shared__ int s_hist[WARPS][256];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const size_t wb = (size_t) b * ws_stride;
    const int n_big = (int) big_count_base[wb + G];
    for (int bi = cta * WARPS + w; bi < n_big; bi += n_ctas * WARPS) {
        const int pid = big_list[wb + bi];
        const int4 pi = pinfo[(wb >> 2) + pid];
        const int n = pi.y, row0 = pi.z, keep = pi.w;
        const int* lst = list + wb + pi.x;
        const int kk = min(n, npv);
        int prefix_v = 0, k = kk;
        const int nbits = 32 - __clz(max_points | 1);
        for (int shift = ((nbits - 1) / 8) * 8; shift >= 0; shift -= 8) {
            for (int bn = lane; bn < 256; bn += 32) s_hist[w][bn] = 0;
            __syncwarp();
            const unsigned himask = shift + 8 >= 32 ? 0u : ~((1u << (shift + 8)) - 1u);
            for (int t = lane; t < n; t += 32) {
                const unsigned e = (unsigned) lst[t];
                if ((e & himask) == (unsigned) prefix_v) atomicAdd(&s_hist[w][(e >> shift) & 255u], 1);
            }
            __syncwarp();
            int c[8], lsum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i] = s_hist[w][lane * 8 + i]; lsum += c[i]; }
            const int incl = warp_incl_scan(lsum, lane);
            const int excl = incl - lsum;
            const unsigned hit = __ballot_sync(0xffffffffu, incl >= k);     // the lane whose bins contain the k-th element
            const int src = __ffs(hit) - 1;
            int digit = 0, below = 0;
            if (lane == src) {
                int run = excl;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (run + c[i] >= k) { digit = lane * 8 + i; below = run; break; }
                    run += c[i];
                }
            }
            digit = __shfl_sync(0xffffffffu, digit, src);
            below = __shfl_sync(0xffffffffu, below, src);
            k -= below;
            prefix_v |= digit << shift;
            __syncwarp();
        }
        const int T = prefix_v;                               // the kk-th smallest input index (indices are unique)
        int filled = 0;
        for (int t0 = 0; t0 < n; t0 += 32) {
            const int t = t0 + lane;
            const int e = t < n ? lst[t] : 0x7fffffff;
            const bool take = e <= T;
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            if (take) s_idx[w][filled + __popc(bal & ((1u << lane) - 1))] = e;
            filled += __popc(bal);
        }
        __syncwarp();
        const int e0 = lane < kk ? s_idx[w][lane] : 0x7fffffff;
        const int e1 = lane + 32 < kk ? s_idx[w][lane + 32] : 0x7fffffff;
        int r0 = 0, r1 = 0;                                   // rank-sort kk (<= 64) unique indices
        for (int j = 0; j < kk; ++j) {
            const int v = j < 32 ? __shfl_sync(0xffffffffu, e0, j) : __shfl_sync(0xffffffffu, e1, j - 32);
            r0 += v < e0;
            r1 += v < e1;
        }
        if (lane < kk && r0 < keep) row_info[(wb >> 1) + row0 + r0] = make_int2(e0, pid);
        if (lane + 32 < kk && r1 < keep) row_info[(wb >> 1) + row0 + r1] = make_int2(e1, pid);
        __syncwarp();
    }
}

// One thread per pillar: sequential-order f32 sums of the kept points (points2Features.cu:809-824).
__global__ void __launch_bounds__(kThreads)
vox_mean_kernel(const float4* __restrict__ points, int max_points, const int* __restrict__ pillar_num,
                const int4* __restrict__ pinfo, const int2* __restrict__ row_info, float4* __restrict__ mean,
                size_t ws_stride)
{
    const int b = blockIdx.y;
    const int pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= pillar_num[b]) return;
    const size_t wb = (size_t) b * ws_stride;
    const int4 pi = __ldg(pinfo + (wb >> 2) + pid);
    const int keep = pi.w;
    const int2* src = row_info + (wb >> 1) + pi.z;
    const float4* pts = points + (size_t) b * max_points;
    float ax = 0.f, ay = 0.f, az = 0.f;
    constexpr int kChunk = 4;                                 // loads of 4 slots in flight, sums strictly in slot order
    for (int s0 = 0; s0 < keep; s0 += kChunk) {
        float4 pbuf[kChunk];
#pragma unroll
        for (int u = 0; u < kChunk; ++u)                      // clamped, not predicated: a repeat of the last slot is harmless
            pbuf[u] = __ldg(pts + __ldg(&src[min(s0 + u, keep - 1)].x));
#pragma unroll
        for (int u = 0; u < kChunk; ++u)
            if (s0 + u < keep) { ax += pbuf[u].x; ay += pbuf[u].y; az += pbuf[u].z; }
    }
    mean[(wb >> 2) + pid] = make_float4(ax / keep, ay / keep, az / keep, 0.f);
}

// One thread per output row: the 10-channel decoration; a CTA's 256 rows leave through shared memory as 10 KB of
// contiguous, 16-byte vectorised stores.
__global__ void __launch_bounds__(kThreads)
vox_feat_kernel(const float4* __restrict__ points, int max_points, VoxGeom g, const int* __restrict__ point_num,
                const int4* __restrict__ pinfo, const int2* __restrict__ row_info,
                const float4* __restrict__ mean, float* __restrict__ point_features,
                int* __restrict__ point_index_in_voxel, size_t ws_stride, int npv, int max_pillars, int max_rows)
{
    __shared__ __align__(16) float s_feat[kThreads * 10];
    const int b = blockIdx.y;
    const int Pc = point_num[b];
    const int r0 = blockIdx.x * kThreads;
    if (r0 >= Pc) return;
    const size_t wb = (size_t) b * ws_stride;
    const int r = r0 + threadIdx.x;
    if (r < Pc) {
        const int2 info = row_info[(wb >> 1) + r];
        const int pid = info.y;
        const float4 p = points[(size_t) b * max_points + info.x];
        const float4 m = mean[(wb >> 2) + pid];
        const float x = p.x, y = p.y, z = p.z;
        const int ix = (int) floorf((x - g.x_min) / g.vx);
        const int iy = (int) floorf((y - g.y_min) / g.vy);
        const int iz = (int) floorf((z - g.z_min) / g.vz);
        // pillar-centre offsets are evaluated in double by the reference (:849-851)
        const float fx = x - ((ix + 0.5) * g.vx + g.x_min);
        const float fy = y - ((iy + 0.5) * g.vy + g.y_min);
        const float fz = z - ((iz + 0.5) * g.vz + g.z_min);
        float* f = &s_feat[threadIdx.x * 10];
        f[0] = x; f[1] = y; f[2] = z; f[3] = p.w;
        f[4] = x - m.x; f[5] = y - m.y; f[6] = z - m.z;
        f[7] = fx; f[8] = fy; f[9] = fz;
        point_index_in_voxel[((size_t) b * max_pillars + pid) * npv + (r - __ldg(pinfo + (wb >> 2) + pid).z)] = r;
    }
    __syncthreads();
    const int rows = Pc - r0 < kThreads ? Pc - r0 : kThreads;
    float* dst = point_features + ((size_t) b * max_rows + r0) * 10;      // 256 rows x 40 B: 16-byte aligned
    const int n4 = (reinterpret_cast<uintptr_t>(dst) & 15) ? 0 : rows * 10 / 4;    // odd capacities: scalar path
    for (int t = threadIdx.x; t < n4; t += kThreads)
        reinterpret_cast<float4*>(dst)[t] = reinterpret_cast<const float4*>(s_feat)[t];
    for (int t = n4 * 4 + threadIdx.x; t < rows * 10; t += kThreads) dst[t] = s_feat[t];
}

// Zero-fill of everything beyond the valid counts (the reference memsets all outputs, points2Features.cu:944-952).
// A plain grid-stride streaming kernel: keeping this out of vox_emit removes ~half of that kernel's instructions.
__device__ __forceinline__ void
vox_zero_tails_body(int cta, int n_ctas, const int* __restrict__ pillar_num, const int* __restrict__ point_num,
                    float* __restrict__ point_features, int* __restrict__ point_index_in_voxel,
                    int* __restrict__ coords, int* __restrict__ point_num_in_voxel,
                    int npv, int max_pillars, int max_rows)
{
    const int b = blockIdx.y;
    const int V = pillar_num[b], Pc = point_num[b];
    const unsigned stride = n_ctas * blockDim.x, t0 = cta * blockDim.x + threadIdx.x;
    // feature rows [Pc, max_rows): 10 floats per row = 5 float2 (rows are 8-byte aligned)
    {
        float2* f = reinterpret_cast<float2*>(point_features + ((size_t) b * max_rows + Pc) * 10);
        const unsigned n = (unsigned) (max_rows - Pc) * 5u;
        for (unsigned t = t0; t < n; t += stride) f[t] = make_float2(0.f, 0.f);
    }
    {   // the whole index tensor: tail pillars and the unused slots of valid ones (vox_feat_kernel fills the used slots)
        int4* q = reinterpret_cast<int4*>(point_index_in_voxel + (size_t) b * max_pillars * npv);
        const unsigned n = (reinterpret_cast<uintptr_t>(q) & 15) ? 0u : (unsigned) (((size_t) max_pillars * npv) / 4);
        for (unsigned t = t0; t < n; t += stride) q[t] = make_int4(0, 0, 0, 0);
        for (unsigned t = n * 4 + t0; t < (unsigned) max_pillars * (unsigned) npv; t += stride)
            point_index_in_voxel[(size_t) b * max_pillars * npv + t] = 0;
    }
    {
        int* c = coords + ((size_t) b * max_pillars + V) * 4;
        const unsigned n = (unsigned) (max_pillars - V) * 4u;
        for (unsigned t = t0; t < n; t += stride) c[t] = 0;
        int* m = point_num_in_voxel + (size_t) b * max_pillars + V;
        for (unsigned t = t0; t < (unsigned) (max_pillars - V); t += stride) m[t] = 0;
    }
}

// Two independent stages share one launch each (fewer launches on the single-frame latency path): the first
// `pts_ctas` CTAs of a row run the per-point stage, the rest the other one.
__global__ void __launch_bounds__(kThreads)
vox_scatter_zero_kernel(int pts_ctas, const int* __restrict__ points_size, int max_points, int* __restrict__ pcell,
                        const int* __restrict__ cell_pid, unsigned int* __restrict__ cursor, int* __restrict__ list,
                        size_t ws_stride, const int* __restrict__ pillar_num, const int* __restrict__ point_num,
                        float* __restrict__ point_features, int* __restrict__ point_index_in_voxel,
                        int* __restrict__ coords, int* __restrict__ point_num_in_voxel, int npv, int max_pillars,
                        int max_rows)
{
    if ((int) blockIdx.x < pts_ctas)
        vox_scatter_body(points_size, max_points, pcell, cell_pid, cursor, list, ws_stride);
    else
        vox_zero_tails_body((int) blockIdx.x - pts_ctas, (int) gridDim.x - pts_ctas, pillar_num, point_num, point_features,
                            point_index_in_voxel, coords, point_num_in_voxel, npv, max_pillars, max_rows);
}

__global__ void __launch_bounds__(kThreads)
vox_rank_kernel(int pts_ctas, const int* __restrict__ points_size, int max_points, const int* __restrict__ ppid,
                const int* __restrict__ list, const int4* __restrict__ pinfo, int2* __restrict__ row_info,
                size_t ws_stride, const int* __restrict__ big_list, const unsigned int* __restrict__ big_count_base,
                int G, int npv)
{
    if ((int) blockIdx.x < pts_ctas)
        vox_rank_body(points_size, max_points, ppid, list, pinfo, row_info, ws_stride);
    else
        vox_rank_big_body<kThreads / 32>((int) blockIdx.x - pts_ctas, (int) gridDim.x - pts_ctas, max_points, list, pinfo,
                                         big_list, big_count_base, row_info, ws_stride, G, npv);
}

}  // namespace
}  // namespace dsvt

using namespace dsvt;

static int vox_check(const dsvt_points2features_params* p) {
    DSVT_CHECK_ARG(p != nullptr, "params is NULL");
    DSVT_CHECK_ARG(p->batch >= 1, "batch < 1");
    DSVT_CHECK_ARG(p->point_feature_num == 4, "point_feature_num must be 4 (x,y,z,intensity)");
    DSVT_CHECK_ARG(p->feature_num == 10, "feature_num must be 10");
    DSVT_CHECK_ARG(p->max_num_points_per_voxel >= 1 && p->max_num_points_per_voxel <= kMaxNpv,
                   "max_num_points_per_voxel must be in [1,64]");
    DSVT_CHECK_ARG(p->grid_z == 1, "grid_size z must be 1 (pillars)");
    DSVT_CHECK_ARG(p->grid_x >= 1 && p->grid_y >= 1, "grid_size");
    DSVT_CHECK_ARG((long long) p->grid_x * p->grid_y <= (1ll << 30), "grid too large");
    DSVT_CHECK_ARG(p->max_points_num >= 1 && p->max_points_num_voxel_filter >= 1 && p->max_pillars_num >= 1,
                   "capacities must be >= 1");
    DSVT_CHECK_ARG(p->voxel_x > 0.f && p->voxel_y > 0.f && p->voxel_z > 0.f, "voxel_size must be > 0");
    return DSVT_OK;
}

constexpr int kWsArrays = 13;
static size_t vox_frame_stride_words(const dsvt_points2features_params* p, size_t* offs /*kWsArrays*/) {
    const size_t G = (size_t) p->grid_x * p->grid_y;
    const size_t nblk = (G + kCellsPerBlock - 1) / kCellsPerBlock;
    // cell_count (+ the big-pillar counter right behind it, cleared by the same memset) | cursor | pcell | list |
    // pillar records (int4: list base, n, first row, rows kept) | (unused) | (unused) | block_sums | cell_pid | row_info (int2: point index, pillar) | (unused) | mean (float4) | big_list
    const size_t list_cap = (size_t) p->max_points_num + 3 * (G < (size_t) p->max_points_num ? G : (size_t) p->max_points_num);
    const size_t sizes[kWsArrays] = {G + 8, G, (size_t) p->max_points_num, list_cap,
                                     (size_t) p->max_pillars_num * 4, 0, 0, nblk * 3, G,
                                     (size_t) p->max_points_num_voxel_filter * 2, 0,
                                     (size_t) p->max_pillars_num * 4, (size_t) p->max_pillars_num};
    size_t off = 0;
    for (int i = 0; i < kWsArrays; ++i) {
        if (offs) offs[i] = off;
        off += align_up(sizes[i] * 4, kWsAlign) / 4;
    }
    return off;
}

extern "C" size_t dsvt_points2features_workspace_size(const dsvt_points2features_params* p) {
    if (vox_check(p) != DSVT_OK) return 0;
    return vox_frame_stride_words(p, nullptr) * 4 * (size_t) p->batch;
}

extern "C" int dsvt_points2features_launch(const dsvt_points2features_params* p,
                                           const float* points, const int32_t* points_size,
                                           float* point_features, int32_t* point_index_in_voxel, int32_t* coords,
                                           int32_t* point_num_in_voxel, int32_t* pillar_num, int32_t* point_num,
                                           void* workspace, size_t workspace_bytes, dsvt_stream_t stream)
{
    int rc = vox_check(p);
    if (rc != DSVT_OK) return rc;
    DSVT_CHECK_ARG(points && points_size && point_features && point_index_in_voxel && coords &&
                   point_num_in_voxel && pillar_num && point_num && workspace, "NULL tensor pointer");
    DSVT_CHECK_ARG(((uintptr_t) points & 15) == 0 && ((uintptr_t) coords & 15) == 0 &&
                   ((uintptr_t) workspace & 255) == 0, "points/coords must be 16-B, workspace 256-B aligned");
    size_t offs[kWsArrays];
    const size_t stride = vox_frame_stride_words(p, offs);
    if (workspace_bytes < stride * 4 * (size_t) p->batch) {
        set_last_error("dsvt_points2features_launch: workspace too small (%zu < %zu)", workspace_bytes,
                       stride * 4 * (size_t) p->batch);
        return DSVT_ERR_WORKSPACE_TOO_SMALL;
    }
    unsigned int* w32 = static_cast<unsigned int*>(workspace);
    unsigned int* cell_count = w32 + offs[0];
    unsigned int* cursor = w32 + offs[1];
    int* pcell = reinterpret_cast<int*>(w32 + offs[2]);
    int* list = reinterpret_cast<int*>(w32 + offs[3]);
    int* pillar_base = reinterpret_cast<int*>(w32 + offs[4]);
    int* block_sums = reinterpret_cast<int*>(w32 + offs[7]);
    int* cell_pid = reinterpret_cast<int*>(w32 + offs[8]);
    int2* row_info = reinterpret_cast<int2*>(w32 + offs[9]);
    float4* mean = reinterpret_cast<float4*>(w32 + offs[11]);
    int* big_list = reinterpret_cast<int*>(w32 + offs[12]);

    const int B = p->batch;
    const int G = p->grid_x * p->grid_y;
    const int nblk = (G + kCellsPerBlock - 1) / kCellsPerBlock;
    const VoxGeom g{p->x_min, p->x_max, p->y_min, p->y_max, p->z_min, p->z_max,
                    p->voxel_x, p->voxel_y, p->voxel_z, p->grid_x, p->grid_y};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    // only the counter grid needs clearing (876 KB/frame at 468^2); everything else is fully overwritten
    if (B == 1) {
        DSVT_CUDA(cudaMemsetAsync(cell_count, 0, (size_t) (G + 8) * 4, st));
    } else {
        DSVT_CUDA(cudaMemset2DAsync(cell_count, stride * 4, 0, (size_t) (G + 8) * 4, B, st));
    }
    count_launch();

    const dim3 grid_pts((p->max_points_num + kThreads - 1) / kThreads, B);
    vox_count_kernel<<<grid_pts, kThreads, 0, st>>>(reinterpret_cast<const float4*>(points), points_size,
                                                    p->max_points_num, g, cell_count, pcell, stride, G);
    DSVT_LAUNCH_CHECK();
    vox_block_reduce_kernel<<<dim3(nblk, B), kThreads, 0, st>>>(cell_count, block_sums, stride, G,
                                                                p->max_num_points_per_voxel);
    DSVT_LAUNCH_CHECK();
    vox_scan_kernel<<<dim3(nblk, B), kThreads, 0, st>>>(cell_count, block_sums, cursor, pillar_base, coords, point_num_in_voxel, pillar_num,
                                                        point_num, cell_pid, big_list, cell_count, list, stride, G, p->grid_x,
                                                        p->max_num_points_per_voxel, p->max_pillars_num,
                                                        p->max_points_num_voxel_filter);
    DSVT_LAUNCH_CHECK();
    // scatter (per point) + zero-fill of the outputs beyond the valid counts (before vox_feat_kernel, which fills the
    // used slots of point_index_in_voxel)
    const int zero_ctas = p->zero_tails ? sm_count() * 4 : 0;
    vox_scatter_zero_kernel<<<dim3(grid_pts.x + zero_ctas, B), kThreads, 0, st>>>(
        (int) grid_pts.x, points_size, p->max_points_num, pcell, cell_pid, cursor, list, stride, pillar_num, point_num,
        point_features, point_index_in_voxel, coords, point_num_in_voxel, p->max_num_points_per_voxel,
        p->max_pillars_num, p->max_points_num_voxel_filter);
    DSVT_LAUNCH_CHECK();
    const int4* pinfo = reinterpret_cast<const int4*>(pillar_base);
    vox_rank_kernel<<<dim3(grid_pts.x + 64, B), kThreads, 0, st>>>(
        (int) grid_pts.x, points_size, p->max_points_num, pcell, list, pinfo, row_info, stride, big_list, cell_count, G,
        p->max_num_points_per_voxel);
    DSVT_LAUNCH_CHECK();
    vox_mean_kernel<<<dim3((p->max_pillars_num + kThreads - 1) / kThreads, B), kThreads, 0, st>>>(
        reinterpret_cast<const float4*>(points), p->max_points_num, pillar_num, pinfo, row_info, mean, stride);
    DSVT_LAUNCH_CHECK();
    vox_feat_kernel<<<dim3((p->max_points_num_voxel_filter + kThreads - 1) / kThreads, B), kThreads, 0, st>>>(
        reinterpret_cast<const float4*>(points), p->max_points_num, g, point_num, pinfo, row_info, mean,
        point_features, point_index_in_voxel, stride, p->max_num_points_per_voxel, p->max_pillars_num,
        p->max_points_num_voxel_filter);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}
