// a1 -- pillar voxelisation + 10-channel point decoration.
// Replaces Points2FeaturesPlugin::enqueue (reference plugins/src/points2Features.cu:896-990).
//
// B200-first design (not the reference's dense 168 MB scatter grid):
//   count    : one thread per point, float4 streaming load, f32 cell id, RED.ADD into a
//              876 KB (468x468) L2-resident counter grid; the cell id is parked in `pcell`.
//   scan     : two-kernel (reduce, then scan) exclusive scan over the cell grid -> canonical
//              pillar ids (ascending y*gx+x), list bases and output-row bases.  No spin-waits.
//   scatter  : counting-sort step: each in-range point claims a slot in its cell's segment.
//   emit     : one warp per pillar sorts its (tiny) segment by input index, keeps the lowest
//              `npv`, computes the sequential-order f32 mean exactly like the reference and
//              writes rows with coalesced stores.  Tail rows are zeroed by the same grid.
// Results are deterministic: the reference's three atomicAdd races (:697, :751, :829) are
// replaced by their canonical serial-order outcome.
#include "common.cuh"

namespace dsvt {
namespace {

constexpr int kThreads = 256;
constexpr int kCellsPerThread = 4;
constexpr int kCellsPerBlock = kThreads * kCellsPerThread;  // 1024
constexpr int kMaxNpv = 64;

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
        for (int j = 0; j < 4; ++j) { ne += a[j] > 0; sa += a[j]; sk += min((int) a[j], npv); }
    } else {
        for (int j = 0; j < kCellsPerThread; ++j) {
            if (c0 + j < G) { unsigned int a = cc[c0 + j]; ne += a > 0; sa += a; sk += min((int) a, npv); }
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
                unsigned int* __restrict__ cursor, int* __restrict__ pillar_base, int* __restrict__ pillar_n,
                int* __restrict__ pillar_row,
                int* __restrict__ coords, int* __restrict__ point_num_in_voxel,
                int* __restrict__ pillar_num, int* __restrict__ point_num,
                size_t ws_stride, int G, int gx, int npv, int max_pillars, int max_rows)
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
    for (int j = 0; j < kCellsPerThread; ++j) { ne += a[j] > 0; sa += a[j]; sk += min((int) a[j], npv); }
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
            if (pid < max_pillars) {
                int keep = min((int) a[j], npv);
                keep = max(0, min(keep, max_rows - e2));  // row-capacity guard (reference has none, SURVEY A-5)
                pillar_base[(size_t) b * ws_stride + pid] = e1;
                pillar_n[(size_t) b * ws_stride + pid] = (int) a[j];
                pillar_row[(size_t) b * ws_stride + pid] = e2;
                const int y = cell / gx, x = cell - y * gx;
                *reinterpret_cast<int4*>(coords_b + (size_t) pid * 4) = make_int4(0, 0, y, x);  // :755
                pnv_b[pid] = keep;
            } else if (pid == max_pillars) {
                // pillars beyond the capacity are the highest cells, so their rows all lie behind this base:
                // the emitted rows are exactly [0, e2)
                point_num[b] = min(e2, max_rows);
            }
            e0 += 1; e1 += (int) a[j]; e2 += min((int) a[j], npv);
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        pillar_num[b] = min(prefix[0] + tot0, max_pillars);
        if (prefix[0] + tot0 <= max_pillars) point_num[b] = min(prefix[2] + tot2, max_rows);
    }
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
vox_scatter_kernel(const int* __restrict__ points_size, int max_points, const int* __restrict__ pcell,
                   unsigned int* __restrict__ cursor, int* __restrict__ list, size_t ws_stride)
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
}

// ---------------------------------------------------------------------------
// One warp per pillar slot (valid pillars emit, the rest zero-fill their rows).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
vox_emit_kernel(const float4* __restrict__ points, int max_points, VoxGeom g,
                const int* __restrict__ list, const int* __restrict__ pillar_base,
                const int* __restrict__ pillar_n, const int* __restrict__ pillar_row,
                const int* __restrict__ pillar_num, const int* __restrict__ point_num,
                const int* point_num_in_voxel,  // aliases point_num_in_voxel_out (disjoint rows)
                float* __restrict__ point_features, int* __restrict__ point_index_in_voxel,
                int* __restrict__ coords, int* point_num_in_voxel_out,
                size_t ws_stride, int npv, int max_pillars, int max_rows, int zero_tails)
{
    __shared__ int s_idx[WARPS][kMaxNpv];
    __shared__ float s_pts[WARPS][kMaxNpv][4];
    __shared__ float s_feat[WARPS][kMaxNpv * 10];
    __shared__ int s_hist[WARPS][256];

    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pid = blockIdx.x * WARPS + w;
    const int V = pillar_num[b];
    const int Pc = point_num[b];
    float* feat_b = point_features + (size_t) b * max_rows * 10;
    int* piv_b = point_index_in_voxel + (size_t) b * max_pillars * npv;

    if (pid >= V) return;          // tails are zero-filled by vox_zero_tails_kernel
    const int base = pillar_base[(size_t) b * ws_stride + pid];
    const int n = pillar_n[(size_t) b * ws_stride + pid];
    const int row0 = pillar_row[(size_t) b * ws_stride + pid];
    const int keep = point_num_in_voxel[(size_t) b * max_pillars + pid];  // min(n, npv) after capacity guard
    const int* lst = list + (size_t) b * ws_stride + base;
    const int kk = min(n, npv);  // how many lowest indices we need sorted

    int e0, e1, m;  // up to 64 candidate indices in registers, m = how many
    if (n <= 64) {
        m = n;
        e0 = lane < n ? lst[lane] : 0x7fffffff;
        e1 = lane + 32 < n ? lst[lane + 32] : 0x7fffffff;
    } else {
        // radix-select (8-bit digits, per-warp shared-memory histogram) the kk-th smallest input index T,
        // then gather the kk elements <= T
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
            // the lane whose bin range contains the k-th element
            const unsigned hit = __ballot_sync(0xffffffffu, incl >= k);
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
        const int T = prefix_v;
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
        m = kk;
        e0 = lane < m ? s_idx[w][lane] : 0x7fffffff;
        e1 = lane + 32 < m ? s_idx[w][lane + 32] : 0x7fffffff;
        __syncwarp();
    }
    // rank-sort m (<= 64) unique indices
    int r0 = 0, r1 = 0;
    for (int j = 0; j < m; ++j) {
        const int v = j < 32 ? __shfl_sync(0xffffffffu, e0, j) : __shfl_sync(0xffffffffu, e1, j - 32);
        r0 += v < e0;
        r1 += v < e1;
    }
    if (lane < m && r0 < kMaxNpv) s_idx[w][r0] = e0;
    if (lane + 32 < m && r1 < kMaxNpv) s_idx[w][r1] = e1;
    __syncwarp();

    // load the kept points in slot order
    for (int s = lane; s < keep; s += 32) {
        const float4 p = points[(size_t) b * max_points + s_idx[w][s]];
        s_pts[w][s][0] = p.x; s_pts[w][s][1] = p.y; s_pts[w][s][2] = p.z; s_pts[w][s][3] = p.w;
    }
    __syncwarp();
    // sequential-order f32 sums, one lane per axis (points2Features.cu:809-824)
    float mean = 0.f;
    if (lane < 3) {
        float acc = 0.f;
        for (int s = 0; s < keep; ++s) acc += s_pts[w][s][lane];
        mean = acc / keep;
    }
    const float mx = __shfl_sync(0xffffffffu, mean, 0);
    const float my = __shfl_sync(0xffffffffu, mean, 1);
    const float mz = __shfl_sync(0xffffffffu, mean, 2);

    for (int s = lane; s < keep; s += 32) {
        const float x = s_pts[w][s][0], y = s_pts[w][s][1], z = s_pts[w][s][2], it = s_pts[w][s][3];
        const int ix = (int) floorf((x - g.x_min) / g.vx);
        const int iy = (int) floorf((y - g.y_min) / g.vy);
        const int iz = (int) floorf((z - g.z_min) / g.vz);
        // pillar-centre offsets are evaluated in double by the reference (:849-851)
        const float fx = x - ((ix + 0.5) * g.vx + g.x_min);
        const float fy = y - ((iy + 0.5) * g.vy + g.y_min);
        const float fz = z - ((iz + 0.5) * g.vz + g.z_min);
        float* f = &s_feat[w][s * 10];
        f[0] = x; f[1] = y; f[2] = z; f[3] = it;
        f[4] = x - mx; f[5] = y - my; f[6] = z - mz;
        f[7] = fx; f[8] = fy; f[9] = fz;
    }
    __syncwarp();
    float* dst = feat_b + (size_t) row0 * 10;
    for (int t = lane; t < keep * 10; t += 32) dst[t] = s_feat[w][t];
    for (int s = lane; s < npv; s += 32) {
        if (s < keep) piv_b[(size_t) pid * npv + s] = row0 + s;
        else if (zero_tails) piv_b[(size_t) pid * npv + s] = 0;
    }
}

// Zero-fill of everything beyond the valid counts (the reference memsets all outputs, points2Features.cu:944-952).
// A plain grid-stride streaming kernel: keeping this out of vox_emit removes ~half of that kernel's instructions.
__global__ void __launch_bounds__(256)
vox_zero_tails_kernel(const int* __restrict__ pillar_num, const int* __restrict__ point_num,
                      float* __restrict__ point_features, int* __restrict__ point_index_in_voxel,
                      int* __restrict__ coords, int* __restrict__ point_num_in_voxel,
                      int npv, int max_pillars, int max_rows)
{
    const int b = blockIdx.y;
    const int V = pillar_num[b], Pc = point_num[b];
    const unsigned stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    // feature rows [Pc, max_rows): 10 floats per row = 5 float2 (rows are 8-byte aligned)
    {
        float2* f = reinterpret_cast<float2*>(point_features + ((size_t) b * max_rows + Pc) * 10);
        const unsigned n = (unsigned) (max_rows - Pc) * 5u;
        for (unsigned t = t0; t < n; t += stride) f[t] = make_float2(0.f, 0.f);
    }
    {
        int* q = point_index_in_voxel + ((size_t) b * max_pillars + V) * npv;
        const unsigned n = (unsigned) (max_pillars - V) * (unsigned) npv;
        for (unsigned t = t0; t < n; t += stride) q[t] = 0;
    }
    {
        int* c = coords + ((size_t) b * max_pillars + V) * 4;
        const unsigned n = (unsigned) (max_pillars - V) * 4u;
        for (unsigned t = t0; t < n; t += stride) c[t] = 0;
        int* m = point_num_in_voxel + (size_t) b * max_pillars + V;
        for (unsigned t = t0; t < (unsigned) (max_pillars - V); t += stride) m[t] = 0;
    }
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

static size_t vox_frame_stride_words(const dsvt_points2features_params* p, size_t* offs /*8*/) {
    const size_t G = (size_t) p->grid_x * p->grid_y;
    const size_t nblk = (G + kCellsPerBlock - 1) / kCellsPerBlock;
    const size_t sizes[8] = {G, G, (size_t) p->max_points_num, (size_t) p->max_points_num,
                             (size_t) p->max_pillars_num, (size_t) p->max_pillars_num,
                             (size_t) p->max_pillars_num, nblk * 3};
    size_t off = 0;
    for (int i = 0; i < 8; ++i) {
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
    size_t offs[8];
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
    int* pillar_n = reinterpret_cast<int*>(w32 + offs[5]);
    int* pillar_row = reinterpret_cast<int*>(w32 + offs[6]);
    int* block_sums = reinterpret_cast<int*>(w32 + offs[7]);

    const int B = p->batch;
    const int G = p->grid_x * p->grid_y;
    const int nblk = (G + kCellsPerBlock - 1) / kCellsPerBlock;
    const VoxGeom g{p->x_min, p->x_max, p->y_min, p->y_max, p->z_min, p->z_max,
                    p->voxel_x, p->voxel_y, p->voxel_z, p->grid_x, p->grid_y};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    // only the counter grid needs clearing (876 KB/frame at 468^2); everything else is fully overwritten
    if (B == 1) {
        DSVT_CUDA(cudaMemsetAsync(cell_count, 0, (size_t) G * 4, st));
    } else {
        DSVT_CUDA(cudaMemset2DAsync(cell_count, stride * 4, 0, (size_t) G * 4, B, st));
    }
    count_launch();

    const dim3 grid_pts((p->max_points_num + kThreads - 1) / kThreads, B);
    vox_count_kernel<<<grid_pts, kThreads, 0, st>>>(reinterpret_cast<const float4*>(points), points_size,
                                                    p->max_points_num, g, cell_count, pcell, stride, G);
    DSVT_LAUNCH_CHECK();
    vox_block_reduce_kernel<<<dim3(nblk, B), kThreads, 0, st>>>(cell_count, block_sums, stride, G,
                                                                p->max_num_points_per_voxel);
    DSVT_LAUNCH_CHECK();
    vox_scan_kernel<<<dim3(nblk, B), kThreads, 0, st>>>(cell_count, block_sums, cursor, pillar_base, pillar_n,
                                                        pillar_row, coords, point_num_in_voxel, pillar_num,
                                                        point_num, stride, G, p->grid_x,
                                                        p->max_num_points_per_voxel, p->max_pillars_num,
                                                        p->max_points_num_voxel_filter);
    DSVT_LAUNCH_CHECK();
    vox_scatter_kernel<<<grid_pts, kThreads, 0, st>>>(points_size, p->max_points_num, pcell, cursor, list, stride);
    DSVT_LAUNCH_CHECK();
    constexpr int kWarps = 8;
    const dim3 grid_emit((p->max_pillars_num + kWarps - 1) / kWarps, B);
    vox_emit_kernel<kWarps><<<grid_emit, kWarps * 32, 0, st>>>(
        reinterpret_cast<const float4*>(points), p->max_points_num, g, list, pillar_base, pillar_n, pillar_row,
        pillar_num, point_num, point_num_in_voxel, point_features, point_index_in_voxel, coords,
        point_num_in_voxel, stride, p->max_num_points_per_voxel, p->max_pillars_num,
        p->max_points_num_voxel_filter, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    if (p->zero_tails) {
        vox_zero_tails_kernel<<<dim3(sm_count() * 4, B), 256, 0, st>>>(
            pillar_num, point_num, point_features, point_index_in_voxel, coords, point_num_in_voxel,
            p->max_num_points_per_voxel, p->max_pillars_num, p->max_points_num_voxel_filter);
        DSVT_LAUNCH_CHECK();
    }
    return DSVT_OK;
}
