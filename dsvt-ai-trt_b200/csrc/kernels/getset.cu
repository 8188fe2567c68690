// a2 -- rotated-set partition.  Replaces GetSetPlugin::enqueue (reference plugins/src/getSet.cu:629-704)
// and, below, the producer of its inputs, WindowPartitionPlugin::enqueue (windowPartition.cu:397-470).
//
// getSet: ONE kernel, one CTA per window slot (the reference: five kernels, one THREAD per window
// running a local-memory quicksort).  Per CTA:
//   * set base  = sum over preceding windows of ceil(N/S)        (replaces the atomicAdd race, :337)
//   * Y-major / X-major orderings by stable rank-by-counting in shared memory (keys are unique, so
//     the result equals the reference quicksort's, SURVEY.md A-6 ii)
//   * slot k of set j takes in-window rank ((j*S+k)*N/S)/n_sets   (integer arithmetic, :346)
//   * mask = -FLT_MAX where a slot repeats the previous voxel     (:546-563)
//   * the 8-head broadcast of both masks                           (:589-609)
// CTAs whose index is >= set_num zero-fill their set row (the reference memsets everything first).
#include "common.cuh"
#include <cfloat>

namespace dsvt {
namespace {

constexpr int kGsThreads = 128;

__global__ void __launch_bounds__(kGsThreads)
get_set_kernel(const int* __restrict__ global_index, const int* __restrict__ coors_in_win,
               const int* __restrict__ voxel_num_in_win, const int* __restrict__ win_num,
               int* __restrict__ global_index_in_set, float* __restrict__ set_voxel_mask,
               int* __restrict__ set_num, float* __restrict__ mask_expand_0, float* __restrict__ mask_expand_1,
               int S, int max_win, int max_vpw, int wx, int wy, int wz, int heads, int zero_tails)
{
    extern __shared__ int smem[];
    int* key_y = smem;                 // [max_vpw]
    int* key_x = key_y + max_vpw;      // [max_vpw]
    int* sorted_y = key_x + max_vpw;   // [max_vpw]
    int* sorted_x = sorted_y + max_vpw;
    __shared__ int red[kGsThreads / 32][2];
    __shared__ int s_base, s_total;

    const int b = blockIdx.y;
    const int w = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int W = win_num[b];
    W = W < max_win ? W : max_win;
    const int* vn = voxel_num_in_win + (size_t) b * max_win;

    // prefix / total of per-window set counts
    int before = 0, total = 0;
    for (int i = tid; i < W; i += kGsThreads) {
        int n = vn[i];
        n = n < max_vpw ? n : max_vpw;
        const int ns = (n + S - 1) / S;   // == int(ceilf(float(n)/S)) for n <= 2^24 (getSet.cu:335)
        total += ns;
        if (i < w) before += ns;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (lane == 0) { red[wid][0] = before; red[wid][1] = total; }
    __syncthreads();
    if (tid == 0) {
        int bsum = 0, tsum = 0;
        for (int i = 0; i < kGsThreads / 32; ++i) { bsum += red[i][0]; tsum += red[i][1]; }
        s_base = bsum;
        s_total = tsum < max_win ? tsum : max_win;   // capacity guard: sets beyond max_win are dropped
        if (w == 0) set_num[b] = s_total;
    }
    __syncthreads();
    const int set_base = s_base;
    const int n_total = s_total;

    int* gis0 = global_index_in_set + ((size_t) b * 2 + 0) * max_win * S;
    int* gis1 = global_index_in_set + ((size_t) b * 2 + 1) * max_win * S;
    float* svm0 = set_voxel_mask + ((size_t) b * 2 + 0) * max_win * S;
    float* svm1 = set_voxel_mask + ((size_t) b * 2 + 1) * max_win * S;
    float* me0 = mask_expand_0 + (size_t) b * max_win * heads * S;
    float* me1 = mask_expand_1 + (size_t) b * max_win * heads * S;

    // zero-fill duty for set row `w` if it is beyond the valid sets
    if (zero_tails && w >= n_total) {
        for (int k = tid; k < S; k += kGsThreads) {
            gis0[(size_t) w * S + k] = 0; gis1[(size_t) w * S + k] = 0;
            svm0[(size_t) w * S + k] = 0.f; svm1[(size_t) w * S + k] = 0.f;
        }
        for (int k = tid; k < heads * S; k += kGsThreads) {
            me0[(size_t) w * heads * S + k] = 0.f; me1[(size_t) w * heads * S + k] = 0.f;
        }
    }
    if (w >= W) return;

    int N = vn[w];
    N = N < max_vpw ? N : max_vpw;
    if (N <= 0) return;
    const int n_sets = (N + S - 1) / S;
    const int* gi = global_index + ((size_t) b * max_win + w) * max_vpw;
    const int* cw = coors_in_win + ((size_t) b * max_win + w) * max_vpw * 3;

    // In-window keys are unique (one voxel per cell), so a voxel's rank in either ordering is the number of OCCUPIED
    // cells with a smaller key: a bitmap over the wx*wy*wz cells + word prefix sums gives it in O(1) per voxel
    // (the reference quicksorts per window on one thread, getSet.cu:274-324; our first version counted pairs, O(N^2)).
    // Out-of-range coordinates or duplicate keys (malformed input) fall back to the pair count below.
    const int cells = wx * wy * wz, words = (cells + 31) >> 5;
    unsigned* occ_y = reinterpret_cast<unsigned*>(sorted_x + max_vpw);   // [words]
    unsigned* occ_x = occ_y + words;                                     // [words]
    int* pre_y = reinterpret_cast<int*>(occ_x + words);                  // [words] exclusive popcount prefix
    int* pre_x = pre_y + words;
    __shared__ int s_bad;
    for (int i = tid; i < words; i += kGsThreads) { occ_y[i] = 0u; occ_x[i] = 0u; }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int m = tid; m < N; m += kGsThreads) {
        const int z = cw[m * 3 + 0], y = cw[m * 3 + 1], x = cw[m * 3 + 2];
        const int ky = y * wx * wz + x * wz + z;    // getSet.cu:388-389
        const int kx = x * wy * wz + y * wz + z;    // getSet.cu:463-464
        key_y[m] = ky;
        key_x[m] = kx;
        if ((unsigned) x >= (unsigned) wx || (unsigned) y >= (unsigned) wy || (unsigned) z >= (unsigned) wz) {
            s_bad = 1;
        } else {
            const unsigned oy = atomicOr(&occ_y[ky >> 5], 1u << (ky & 31));
            atomicOr(&occ_x[kx >> 5], 1u << (kx & 31));
            if (oy & (1u << (ky & 31))) s_bad = 1;  // two voxels in one cell
        }
    }
    __syncthreads();
    if (!s_bad) {
        if (tid < 32) {                             // exclusive prefix of the word popcounts (<= 18 words at 24x24)
            int cy = 0, cx = 0;
            for (int i0 = 0; i0 < words; i0 += 32) {
                const int i = i0 + tid;
                const int py = i < words ? __popc(occ_y[i]) : 0, px = i < words ? __popc(occ_x[i]) : 0;
                const int iy = warp_incl_scan(py, lane), ix = warp_incl_scan(px, lane);
                if (i < words) { pre_y[i] = cy + iy - py; pre_x[i] = cx + ix - px; }
                cy += __shfl_sync(0xffffffffu, iy, 31);
                cx += __shfl_sync(0xffffffffu, ix, 31);
            }
        }
        __syncthreads();
        for (int m = tid; m < N; m += kGsThreads) {
            const int ky = key_y[m], kx = key_x[m];
            const int ry = pre_y[ky >> 5] + __popc(occ_y[ky >> 5] & ((1u << (ky & 31)) - 1u));
            const int rx = pre_x[kx >> 5] + __popc(occ_x[kx >> 5] & ((1u << (kx & 31)) - 1u));
            const int g = gi[m];
            sorted_y[ry] = g;
            sorted_x[rx] = g;
        }
    } else {
        for (int m = tid; m < N; m += kGsThreads) {
            const int ky = key_y[m], kx = key_x[m];
            int ry = 0, rx = 0;
            for (int q = 0; q < N; ++q) {
                const int qy = key_y[q], qx = key_x[q];
                ry += (qy < ky) || (qy == ky && q < m);
                rx += (qx < kx) || (qx == kx && q < m);
            }
            const int g = gi[m];
            sorted_y[ry] = g;
            sorted_x[rx] = g;
        }
    }
    __syncthreads();

    const float kMasked = -FLT_MAX;   // the double literal -3.4028235e+38 rounds to exactly this (SURVEY A-6 iv)
    for (int t = tid; t < n_sets * S; t += kGsThreads) {
        const int j = t / S, k = t - j * S;
        const int set = set_base + j;
        if (set >= max_win) continue;
        const int r = ((j * S + k) * N / S) / n_sets;
        const int gy = sorted_y[r], gx = sorted_x[r];
        float my = 0.f, mx = 0.f;
        if (k > 0) {
            const int rp = ((j * S + k - 1) * N / S) / n_sets;
            my = (sorted_y[rp] == gy) ? kMasked : 0.f;
            mx = (sorted_x[rp] == gx) ? kMasked : 0.f;
        }
        gis0[(size_t) set * S + k] = gy;
        gis1[(size_t) set * S + k] = gx;
        svm0[(size_t) set * S + k] = my;
        svm1[(size_t) set * S + k] = mx;
        for (int h = 0; h < heads; ++h) {
            me0[((size_t) set * heads + h) * S + k] = my;
            me1[((size_t) set * heads + h) * S + k] = mx;
        }
    }
}

// ---------------------------------------------------------------------------
// window partition (deterministic; the reference spins on a racy window-id publication,
// windowPartition.cu:323-331)
constexpr int kWpThreads = 256;

__global__ void __launch_bounds__(kWpThreads)
wp_count_kernel(const int* __restrict__ coords, const int* __restrict__ voxel_num, int max_pillars,
                int* __restrict__ dense_count, int* __restrict__ vox_win, size_t ws_stride,
                int sx_, int sy_, int sz_, int wx, int wy, int wz, int nwx, int nwy, int nwz,
                int* __restrict__ coors_2d, float* __restrict__ coors_xy, int zero_tails)
{
    const int b = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= max_pillars) return;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    int* c2d = coors_2d + ((size_t) b * max_pillars + v) * 3;
    float* cxy = coors_xy + ((size_t) b * max_pillars + v) * 2;
    if (v >= V) {
        if (zero_tails) { c2d[0] = 0; c2d[1] = 0; c2d[2] = 0; cxy[0] = 0.f; cxy[1] = 0.f; }
        return;
    }
    const int4 c = *reinterpret_cast<const int4*>(coords + ((size_t) b * max_pillars + v) * 4);
    const unsigned sx = (unsigned) c.w + sx_, sy = (unsigned) c.z + sy_, sz = (unsigned) c.y + sz_;  // :290-292
    const unsigned wcx = sx / wx, wcy = sy / wy, wcz = sz / wz;                                       // :294-296
    int dense = -1;
    if (wcx < (unsigned) nwx && wcy < (unsigned) nwy && wcz < (unsigned) nwz) {
        dense = (int) (wcz * (nwy * nwx) + wcy * nwx + wcx);                                            // :299
        atomicAdd(dense_count + (size_t) b * ws_stride + dense, 1);
    }
    vox_win[(size_t) b * ws_stride + v] = dense;
    const int cx = sx % wx, cy = sy % wy, cz = sz % wz;                                                 // :343-345
    c2d[0] = cz; c2d[1] = cy; c2d[2] = cx;                                                              // :353-355
    cxy[0] = float(cx) - float(wx) / 2;                                                                 // :358-359
    cxy[1] = float(cy) - float(wy) / 2;
}

// single CTA per frame: dense window grid -> compact window slots (ascending dense index)
__global__ void __launch_bounds__(1024)
wp_scan_kernel(const int* __restrict__ dense_count, int* __restrict__ dense_slot, int* __restrict__ dense_cursor,
               int* __restrict__ voxel_num_in_win, int* __restrict__ win_num, size_t ws_stride,
               int dense_n, int max_win, int max_vpw, int zero_tails)
{
    __shared__ int warp_sums[33];
    __shared__ int s_carry;
    const int b = blockIdx.x;
    const int* dc = dense_count + (size_t) b * ws_stride;
    int* ds = dense_slot + (size_t) b * ws_stride;
    int* dcur = dense_cursor + (size_t) b * ws_stride;
    int* vnw = voxel_num_in_win + (size_t) b * max_win;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < dense_n; c0 += blockDim.x) {
        const int i = c0 + threadIdx.x;
        const int cnt = i < dense_n ? dc[i] : 0;
        int tot;
        const int carry = s_carry;
        const int ex = block_excl_scan(cnt > 0, warp_sums, &tot) + carry;
        if (i < dense_n) {
            int slot = -1;
            if (cnt > 0 && ex < max_win) {
                slot = ex;
                vnw[ex] = cnt < max_vpw ? cnt : max_vpw;    // :336-339
            }
            ds[i] = slot;
            dcur[i] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    const int W = s_carry < max_win ? s_carry : max_win;
    if (threadIdx.x == 0) win_num[b] = W;
    if (zero_tails) for (int i = W + threadIdx.x; i < max_win; i += blockDim.x) vnw[i] = 0;
}

// every voxel claims a (race-ordered) slot inside its window's staging list; the finalize kernel sorts the list
__global__ void __launch_bounds__(kWpThreads)
wp_scatter_kernel(const int* __restrict__ voxel_num, int max_pillars, const int* __restrict__ vox_win,
                  const int* __restrict__ dense_count, const int* __restrict__ dense_slot, int* __restrict__ dense_cursor,
                  int* __restrict__ global_index, size_t ws_stride, int max_win, int max_vpw)
{
    const int b = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    int V = voxel_num[b];
    V = V < max_pillars ? V : max_pillars;
    if (v >= V) return;
    const int dense = vox_win[(size_t) b * ws_stride + v];
    if (dense < 0) return;
    const int slot = dense_slot[(size_t) b * ws_stride + dense];
    if (slot < 0) return;
    if (dense_count[(size_t) b * ws_stride + dense] > max_vpw) {
        // capacity overflow (only possible when max_voxel_num_per_win < cells per window): the serial-execution outcome
        // keeps the max_vpw LOWEST voxel ids (the reference drops by race, :303).  Rank by counting -- a correctness
        // path, O(V) per voxel of an overfull window.
        const int* vw = vox_win + (size_t) b * ws_stride;
        int rank = 0;
        for (int u = 0; u < v && rank < max_vpw; ++u) rank += vw[u] == dense;
        if (rank < max_vpw) global_index[((size_t) b * max_win + slot) * max_vpw + rank] = v;
        return;
    }
    const int pos = atomicAdd(dense_cursor + (size_t) b * ws_stride + dense, 1);   // staged unsorted
    global_index[((size_t) b * max_win + slot) * max_vpw + pos] = v;
}

// one CTA per window slot: sort the staged voxel ids ascending, emit in-window coordinates
__global__ void __launch_bounds__(kGsThreads)
wp_finalize_kernel(const int* __restrict__ coords, int max_pillars, const int* __restrict__ voxel_num_in_win,
                   const int* __restrict__ win_num, int* __restrict__ global_index, int* __restrict__ coors_in_win,
                   int max_win, int max_vpw, int sx_, int sy_, int sz_, int wx, int wy, int wz, int zero_tails)
{
    extern __shared__ int smem[];
    int* ids = smem;               // [max_vpw]
    int* sorted = smem + max_vpw;  // [max_vpw]
    const int b = blockIdx.y, w = blockIdx.x, tid = threadIdx.x;
    int W = win_num[b];
    W = W < max_win ? W : max_win;
    int* gi = global_index + ((size_t) b * max_win + w) * max_vpw;
    int* cw = coors_in_win + ((size_t) b * max_win + w) * max_vpw * 3;
    int N = 0;
    if (w < W) {
        N = voxel_num_in_win[(size_t) b * max_win + w];
        // Ascending voxel id.  With canonical pillar ids (ascending y*gx+x, what our voxeliser emits) that is also the
        // ascending order of the in-window cell key, and the key's rank is the number of occupied cells below it: a
        // bitmap over the window's cells gives the order in O(N).  The result is verified (ids ascending in that order);
        // any other input -- e.g. the reference voxeliser's race-ordered ids -- falls back to the O(N^2) pair count.
        const int cells = wx * wy * wz, words = (cells + 31) >> 5;
        unsigned* occ = reinterpret_cast<unsigned*>(sorted + max_vpw);   // [words]
        int* pre = reinterpret_cast<int*>(occ + words);                  // [words]
        int* keys = pre + words;                                         // [max_vpw]
        __shared__ int s_bad;
        for (int i = tid; i < words; i += kGsThreads) occ[i] = 0u;
        if (tid == 0) s_bad = 0;
        for (int m = tid; m < N; m += kGsThreads) ids[m] = gi[m];
        __syncthreads();
        for (int m = tid; m < N; m += kGsThreads) {
            const int4 c = *reinterpret_cast<const int4*>(coords + ((size_t) b * max_pillars + ids[m]) * 4);
            const unsigned sx = (unsigned) c.w + sx_, sy = (unsigned) c.z + sy_, sz = (unsigned) c.y + sz_;
            const int k = (int) (sy % wy) * wx * wz + (int) (sx % wx) * wz + (int) (sz % wz);
            keys[m] = k;
            if (atomicOr(&occ[k >> 5], 1u << (k & 31)) & (1u << (k & 31))) s_bad = 1;
        }
        __syncthreads();
        if (tid < 32) {
            int carry = 0;
            const int lane = tid;
            for (int i0 = 0; i0 < words; i0 += 32) {
                const int i = i0 + lane;
                const int pc = i < words ? __popc(occ[i]) : 0;
                const int inc = warp_incl_scan(pc, lane);
                if (i < words) pre[i] = carry + inc - pc;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        __syncthreads();
        if (!s_bad)
            for (int m = tid; m < N; m += kGsThreads) {
                const int k = keys[m];
                sorted[pre[k >> 5] + __popc(occ[k >> 5] & ((1u << (k & 31)) - 1u))] = ids[m];
            }
        __syncthreads();
        if (!s_bad)
            for (int m = tid; m + 1 < N; m += kGsThreads)
                if (sorted[m] >= sorted[m + 1]) s_bad = 1;              // not the ascending-id order: redo below
        __syncthreads();
        if (s_bad) {
            for (int m = tid; m < N; m += kGsThreads) {
                const int e = ids[m];
                int r = 0;
                for (int q = 0; q < N; ++q) r += ids[q] < e;
                sorted[r] = e;
            }
        }
        __syncthreads();
        for (int m = tid; m < N; m += kGsThreads) {
            const int v = sorted[m];
            gi[m] = v;
            const int4 c = *reinterpret_cast<const int4*>(coords + ((size_t) b * max_pillars + v) * 4);
            const unsigned sx = (unsigned) c.w + sx_, sy = (unsigned) c.z + sy_, sz = (unsigned) c.y + sz_;
            cw[m * 3 + 0] = (int) (sz % wz);   // :348-350 (z,y,x)
            cw[m * 3 + 1] = (int) (sy % wy);
            cw[m * 3 + 2] = (int) (sx % wx);
        }
    }
    if (zero_tails) {
        for (int m = N + tid; m < max_vpw; m += kGsThreads) {
            gi[m] = 0;
            cw[m * 3 + 0] = 0; cw[m * 3 + 1] = 0; cw[m * 3 + 2] = 0;
        }
    }
}

}  // namespace
}  // namespace dsvt

using namespace dsvt;

// ---- getSet host side ------------------------------------------------------
static int gs_check(const dsvt_get_set_params* p) {
    DSVT_CHECK_ARG(p != nullptr, "params is NULL");
    DSVT_CHECK_ARG(p->batch >= 1, "batch < 1");
    DSVT_CHECK_ARG(p->voxel_num_set >= 1 && p->voxel_num_set <= 1024, "voxel_num_set out of range");
    DSVT_CHECK_ARG(p->max_win_num >= 1 && p->max_voxel_num_per_win >= 1, "capacities must be >= 1");
    DSVT_CHECK_ARG((size_t) p->max_voxel_num_per_win * 16 <= 200 * 1024, "max_voxel_num_per_win too large for shared memory");
    DSVT_CHECK_ARG(p->win_shape_x >= 1 && p->win_shape_y >= 1 && p->win_shape_z >= 1, "win_shape");
    DSVT_CHECK_ARG(p->num_heads >= 1, "num_heads");
    return DSVT_OK;
}

extern "C" size_t dsvt_get_set_workspace_size(const dsvt_get_set_params* p) {
    (void) p;
    return 0;   // the reference needs 2 x 1.84 MB + 115 KB + 3.2 KB (getSet.cu:237-255); we sort in shared memory
}

extern "C" int dsvt_get_set_launch(const dsvt_get_set_params* p,
                                   const int32_t* global_index, const int32_t* coors_in_win,
                                   const int32_t* voxel_num_in_win, const int32_t* win_num,
                                   int32_t* global_index_in_set, float* set_voxel_mask, int32_t* set_num,
                                   float* mask_expand_0, float* mask_expand_1,
                                   void* workspace, size_t workspace_bytes, dsvt_stream_t stream)
{
    (void) workspace; (void) workspace_bytes;
    int rc = gs_check(p);
    if (rc != DSVT_OK) return rc;
    DSVT_CHECK_ARG(global_index && coors_in_win && voxel_num_in_win && win_num && global_index_in_set &&
                   set_voxel_mask && set_num && mask_expand_0 && mask_expand_1, "NULL tensor pointer");
    // keys + sorted ids (4 x max_vpw ints) + two occupancy bitmaps and their popcount prefixes over the window's cells
    const size_t cell_words = ((size_t) p->win_shape_x * p->win_shape_y * p->win_shape_z + 31) / 32;
    const size_t smem = ((size_t) p->max_voxel_num_per_win * 4 + 4 * cell_words) * sizeof(int);
    DSVT_CHECK_ARG(smem <= 200 * 1024, "window too large for shared memory");
    if (smem > 48 * 1024) DSVT_RAISE_SMEM(get_set_kernel, 200 * 1024);
    get_set_kernel<<<dim3(p->max_win_num, p->batch), kGsThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        global_index, coors_in_win, voxel_num_in_win, win_num, global_index_in_set, set_voxel_mask, set_num,
        mask_expand_0, mask_expand_1, p->voxel_num_set, p->max_win_num, p->max_voxel_num_per_win,
        p->win_shape_x, p->win_shape_y, p->win_shape_z, p->num_heads, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

// ---- window partition host side -------------------------------------------
static int wp_check(const dsvt_window_partition_params* p) {
    DSVT_CHECK_ARG(p != nullptr, "params is NULL");
    DSVT_CHECK_ARG(p->batch >= 1, "batch < 1");
    DSVT_CHECK_ARG(p->max_pillars_num >= 1 && p->max_win_num >= 1 && p->max_voxel_num_per_win >= 1, "capacities");
    DSVT_CHECK_ARG(p->win_shape_x >= 1 && p->win_shape_y >= 1 && p->win_shape_z >= 1, "win_shape");
    DSVT_CHECK_ARG(p->sparse_shape_x >= 1 && p->sparse_shape_y >= 1 && p->sparse_shape_z >= 1, "sparse_shape");
    DSVT_CHECK_ARG(p->shift_x >= 0 && p->shift_y >= 0 && p->shift_z >= 0, "shift must be >= 0");
    DSVT_CHECK_ARG((size_t) p->max_voxel_num_per_win * 8 <= 200 * 1024, "max_voxel_num_per_win too large");
    return DSVT_OK;
}

static void wp_dense(const dsvt_window_partition_params* p, int* nwx, int* nwy, int* nwz) {
    // int(ceilf(sparse/win) + 1) with INTEGER division inside (windowPartition.cu:425-427)
    *nwx = p->sparse_shape_x / p->win_shape_x + 1;
    *nwy = p->sparse_shape_y / p->win_shape_y + 1;
    *nwz = p->sparse_shape_z / p->win_shape_z + 1;
}

static size_t wp_stride_words(const dsvt_window_partition_params* p, size_t* offs /*4*/) {
    int nwx, nwy, nwz;
    wp_dense(p, &nwx, &nwy, &nwz);
    const size_t dn = (size_t) nwx * nwy * nwz;
    const size_t sizes[4] = {dn, dn, dn, (size_t) p->max_pillars_num};
    size_t off = 0;
    for (int i = 0; i < 4; ++i) {
        if (offs) offs[i] = off;
        off += align_up(sizes[i] * 4, kWsAlign) / 4;
    }
    return off;
}

extern "C" size_t dsvt_window_partition_workspace_size(const dsvt_window_partition_params* p) {
    if (wp_check(p) != DSVT_OK) return 0;
    return wp_stride_words(p, nullptr) * 4 * (size_t) p->batch;
}

extern "C" int dsvt_window_partition_launch(const dsvt_window_partition_params* p,
                                            const int32_t* coords, const int32_t* voxel_num,
                                            int32_t* global_index, int32_t* coors_in_win, int32_t* voxel_num_in_win,
                                            int32_t* win_num, int32_t* coors_in_win_2d, float* coors_in_win_x_y,
                                            void* workspace, size_t workspace_bytes, dsvt_stream_t stream)
{
    int rc = wp_check(p);
    if (rc != DSVT_OK) return rc;
    DSVT_CHECK_ARG(coords && voxel_num && global_index && coors_in_win && voxel_num_in_win && win_num &&
                   coors_in_win_2d && coors_in_win_x_y && workspace, "NULL tensor pointer");
    DSVT_CHECK_ARG(((uintptr_t) coords & 15) == 0 && ((uintptr_t) workspace & 255) == 0, "alignment");
    size_t offs[4];
    const size_t stride = wp_stride_words(p, offs);
    if (workspace_bytes < stride * 4 * (size_t) p->batch) {
        set_last_error("dsvt_window_partition_launch: workspace too small");
        return DSVT_ERR_WORKSPACE_TOO_SMALL;
    }
    int nwx, nwy, nwz;
    wp_dense(p, &nwx, &nwy, &nwz);
    const int dn = nwx * nwy * nwz;
    int* w32 = static_cast<int*>(workspace);
    int* dense_count = w32 + offs[0];
    int* dense_slot = w32 + offs[1];
    int* dense_cursor = w32 + offs[2];
    int* vox_win = w32 + offs[3];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int B = p->batch;
    if (B == 1) DSVT_CUDA(cudaMemsetAsync(dense_count, 0, (size_t) dn * 4, st));
    else DSVT_CUDA(cudaMemset2DAsync(dense_count, stride * 4, 0, (size_t) dn * 4, B, st));
    count_launch();
    const dim3 grid_v((p->max_pillars_num + kWpThreads - 1) / kWpThreads, B);
    wp_count_kernel<<<grid_v, kWpThreads, 0, st>>>(coords, voxel_num, p->max_pillars_num, dense_count, vox_win,
                                                   stride, p->shift_x, p->shift_y, p->shift_z, p->win_shape_x,
                                                   p->win_shape_y, p->win_shape_z, nwx, nwy, nwz,
                                                   coors_in_win_2d, coors_in_win_x_y, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    wp_scan_kernel<<<B, 1024, 0, st>>>(dense_count, dense_slot, dense_cursor, voxel_num_in_win, win_num, stride,
                                       dn, p->max_win_num, p->max_voxel_num_per_win, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    wp_scatter_kernel<<<grid_v, kWpThreads, 0, st>>>(voxel_num, p->max_pillars_num, vox_win, dense_count, dense_slot,
                                                     dense_cursor, global_index, stride, p->max_win_num,
                                                     p->max_voxel_num_per_win);
    DSVT_LAUNCH_CHECK();
    // ids + sorted + keys (3 x max_vpw ints) + the occupancy bitmap and its popcount prefix over the window's cells
    const size_t wp_cell_words = ((size_t) p->win_shape_x * p->win_shape_y * p->win_shape_z + 31) / 32;
    const size_t smem = ((size_t) p->max_voxel_num_per_win * 3 + 2 * wp_cell_words) * sizeof(int);
    if (smem > 48 * 1024) DSVT_RAISE_SMEM(wp_finalize_kernel, 200 * 1024);
    wp_finalize_kernel<<<dim3(p->max_win_num, B), kGsThreads, smem, st>>>(
        coords, p->max_pillars_num, voxel_num_in_win, win_num, global_index, coors_in_win, p->max_win_num,
        p->max_voxel_num_per_win, p->shift_x, p->shift_y, p->shift_z, p->win_shape_x, p->win_shape_y,
        p->win_shape_z, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}
