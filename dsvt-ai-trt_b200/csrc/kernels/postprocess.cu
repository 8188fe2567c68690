// (next #4 tail) the CenterHead post-process graph and the rotated NMS behind FilterBoxByScorePlugin.
//
// 1. dsvt_center_head_topk_launch -- the TensorRT-native graph in front of FilterBoxByScorePlugin (reference
//    src/dsvt-ai-trt.cpp:1471-1691): sigmoid(heat map) -> TopK(500) per class over H*W -> TopK(500) over the 10 x 500
//    survivors -> index arithmetic (class = ind / 500, y = cell / W, x = cell % W) -> gather of center / center_z / dim /
//    rot at the winning cells -> exp(dim), atan(sin / cos).  Its eight outputs are the eight inputs of the plugin.
//    The two-stage TopK equals ONE global top-500 over the 10 * H * W scores (a class can contribute at most 500 anyway),
//    and the sigmoid is monotonic, so we select on the LOGITS: a two-level radix histogram (12 + 12 key bits) finds the
//    24-bit key prefix of the 500th largest value, one pass collects the few hundred candidates at or above it, and one
//    CTA sorts them (score descending, flat index ascending = the order a stable serial top-k produces) and gathers.
//    The 8.8 MB heat map is read three times (L2-resident after the first pass); the reference materialises ~20 tensors.
// 2. dsvt_rotated_nms_launch -- nms_cpu (reference include/helper.h:257-283, box_overlap :166-255), which the reference
//    runs on the HOST after a D2H copy (src/dsvt-ai-trt.cpp:1939-1954): sort by score, pairwise rotated IoU as a
//    suppression bit matrix (one thread per pair), then the sequential greedy scan by one warp.  The polygon clipping
//    code is restated operation by operation (this file is compiled with -fmad=false) so that every IoU decision matches
//    the host arithmetic up to the last bits of cosf / sinf / atan2f.
#include "common.cuh"

namespace dsvt {
namespace {

constexpr int kBins = 4096;
constexpr int kCandCap = 8192;
constexpr int kPpThreads = 256;

__device__ __forceinline__ unsigned ord_key(float f) {          // ascending unsigned order == ascending float order
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct PpWs { unsigned* hist1; unsigned* hist2; unsigned* cand_count; uint2* cand; };
__host__ __device__ inline size_t pp_ws_words() { return 2 * kBins + 64 + 2 * (size_t) kCandCap; }
__device__ __forceinline__ PpWs pp_ws(unsigned* base, int b) {
    unsigned* p = base + (size_t) b * pp_ws_words();
    return PpWs{p, p + kBins, p + 2 * kBins, reinterpret_cast<uint2*>(p + 2 * kBins + 64)};
}

// Largest bin b with count(bins > b) < K <= count(bins >= b); *above = count(bins > b).  All threads of a 256-thread CTA.
__device__ int pick_bin(const unsigned* __restrict__ hist, int K, int* above, int* s_tmp /*[kPpThreads + 2]*/)
{
    const int t = threadIdx.x;
    constexpr int per = kBins / kPpThreads;                      // 16 consecutive bins per thread, thread 0 = the top bins
    int local = 0;
    for (int i = 0; i < per; ++i) local += (int) hist[kBins - 1 - (t * per + i)];
    s_tmp[t] = local;
    __syncthreads();
    if (t == 0) {
        int run = 0, who = kPpThreads - 1;
        for (int i = 0; i < kPpThreads; ++i) {
            if (run + s_tmp[i] >= K) { who = i; break; }
            run += s_tmp[i];
        }
        int bin = 0, ab = run;
        for (int i = 0; i < per; ++i) {
            const int bidx = kBins - 1 - (who * per + i);
            const int c = (int) hist[bidx];
            if (ab + c >= K || i == per - 1) { bin = bidx; break; }
            ab += c;
        }
        s_tmp[kPpThreads] = bin;
        s_tmp[kPpThreads + 1] = ab;
    }
    __syncthreads();
    const int bin = s_tmp[kPpThreads];
    *above = s_tmp[kPpThreads + 1];
    __syncthreads();
    return bin;
}

__global__ void __launch_bounds__(kPpThreads)
ph_hist1_kernel(const float* __restrict__ hm, unsigned* __restrict__ ws, int N)
{
    __shared__ unsigned s_h[kBins];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < kBins; i += kPpThreads) s_h[i] = 0;
    __syncthreads();
    const float* src = hm + (size_t) b * N;
    for (int i = blockIdx.x * kPpThreads + threadIdx.x; i < N; i += gridDim.x * kPpThreads)
        atomicAdd(&s_h[ord_key(__ldg(src + i)) >> 20], 1u);
    __syncthreads();
    unsigned* h = pp_ws(ws, b).hist1;
    for (int i = threadIdx.x; i < kBins; i += kPpThreads)
        if (s_h[i]) atomicAdd(h + i, s_h[i]);
}

__global__ void __launch_bounds__(kPpThreads)
ph_hist2_kernel(const float* __restrict__ hm, unsigned* __restrict__ ws, int N, int K)
{
    __shared__ unsigned s_h[kBins];
    __shared__ int s_tmp[kPpThreads + 2];
    const int b = blockIdx.y;
    const PpWs w = pp_ws(ws, b);
    int above;
    const unsigned b1 = (unsigned) pick_bin(w.hist1, K, &above, s_tmp);
    for (int i = threadIdx.x; i < kBins; i += kPpThreads) s_h[i] = 0;
    __syncthreads();
    const float* src = hm + (size_t) b * N;
    for (int i = blockIdx.x * kPpThreads + threadIdx.x; i < N; i += gridDim.x * kPpThreads) {
        const unsigned k = ord_key(__ldg(src + i));
        if ((k >> 20) == b1) atomicAdd(&s_h[(k >> 8) & 0xFFFu], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += kPpThreads)
        if (s_h[i]) atomicAdd(w.hist2 + i, s_h[i]);
}

__global__ void __launch_bounds__(kPpThreads)
ph_collect_kernel(const float* __restrict__ hm, unsigned* __restrict__ ws, int N, int K)
{
    __shared__ int s_tmp[kPpThreads + 2];
    const int b = blockIdx.y;
    const PpWs w = pp_ws(ws, b);
    int above1, above2;
    const unsigned b1 = (unsigned) pick_bin(w.hist1, K, &above1, s_tmp);
    const unsigned b2 = (unsigned) pick_bin(w.hist2, K - above1, &above2, s_tmp);
    const unsigned prefix = (b1 << 12) | b2;                      // 24-bit key prefix of the K-th largest value
    const float* src = hm + (size_t) b * N;
    for (int i = blockIdx.x * kPpThreads + threadIdx.x; i < N; i += gridDim.x * kPpThreads) {
        const unsigned k = ord_key(__ldg(src + i));
        if ((k >> 8) >= prefix) {
            const unsigned pos = atomicAdd(w.cand_count, 1u);
            if (pos < (unsigned) kCandCap) w.cand[pos] = make_uint2(k, (unsigned) i);
        }
    }
}

// One CTA per frame: sort the candidates (key descending, flat index ascending), emit the top K with everything
// FilterBoxByScorePlugin wants.
__global__ void __launch_bounds__(1024)
ph_emit_kernel(const unsigned* __restrict__ ws_c, const float* __restrict__ center, const float* __restrict__ center_z,
               const float* __restrict__ dim, const float* __restrict__ rot, float* __restrict__ scores,
               int* __restrict__ classes, int* __restrict__ xs, int* __restrict__ ys, float* __restrict__ center_g,
               float* __restrict__ center_z_g, float* __restrict__ angle, float* __restrict__ dim_g, int HW, int W, int K)
{
    extern __shared__ unsigned long long s_key[];                 // [n_pow2]
    const int b = blockIdx.x;
    const PpWs w = pp_ws(const_cast<unsigned*>(ws_c), b);
    const int n = min((int) *w.cand_count, kCandCap);
    int np2 = 1024;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        unsigned long long v = 0ull;                              // padding sorts last
        if (i < n) { const uint2 c = w.cand[i]; v = ((unsigned long long) c.x << 32) | (unsigned long long) (0xFFFFFFFFu - c.y); }
        s_key[i] = v;
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)                            // bitonic sort, descending
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = s_key[i], c = s_key[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? a < c : a > c) { s_key[i] = c; s_key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const size_t o = (size_t) b * K + k;
        if (k >= n) {                                             // fewer scores than K (cannot happen for N >= K)
            scores[o] = 0.f; classes[o] = 0; xs[o] = 0; ys[o] = 0;
            center_g[o * 2] = 0.f; center_g[o * 2 + 1] = 0.f; center_z_g[o] = 0.f; angle[o] = 0.f;
            dim_g[o * 3] = 0.f; dim_g[o * 3 + 1] = 0.f; dim_g[o * 3 + 2] = 0.f;
            continue;
        }
        const unsigned long long v = s_key[k];
        const unsigned idx = 0xFFFFFFFFu - (unsigned) (v & 0xFFFFFFFFull);
        const float logit = key_float((unsigned) (v >> 32));
        const int cls = (int) (idx / (unsigned) HW), cell = (int) (idx - (unsigned) cls * (unsigned) HW);
        scores[o] = 1.0f / (1.0f + expf(-logit));                 // kSIGMOID (:1479)
        classes[o] = cls;                                         // topk_ind / K (:1570)
        ys[o] = cell / W;                                         // :1543
        xs[o] = cell - (cell / W) * W;                            // :1545-1547
        const float* c = center + (size_t) b * 2 * HW;
        center_g[o * 2] = __ldg(c + cell);
        center_g[o * 2 + 1] = __ldg(c + HW + cell);
        center_z_g[o] = __ldg(center_z + (size_t) b * HW + cell);
        const float* d = dim + (size_t) b * 3 * HW;
        dim_g[o * 3] = expf(__ldg(d + cell));                     // kEXP (:1489)
        dim_g[o * 3 + 1] = expf(__ldg(d + HW + cell));
        dim_g[o * 3 + 2] = expf(__ldg(d + 2 * HW + cell));
        const float* r = rot + (size_t) b * 2 * HW;               // channel 0 = cos, 1 = sin (:1492-1503)
        angle[o] = atanf(__ldg(r + HW + cell) / __ldg(r + cell)); // kDIV then kATAN (:1665-1666): atan, not atan2
    }
}

// ---------------------------------------------------------------------------------------------------------------
// rotated NMS
struct F2 { float x, y; };
constexpr float kNmsEps = 1e-8f;                                   // helper.h:26

__device__ __forceinline__ float nms_cross(F2 p1, F2 p2, F2 p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }
struct NBox { float x, y, w, l, rt; };
__device__ __forceinline__ bool nms_check_box2d(const NBox& box, F2 p) {               // helper.h:111-121
    const float MARGIN = 1e-2f;
    const float ac = cosf(-box.rt), as = sinf(-box.rt);
    const float rot_x = (p.x - box.x) * ac + (p.y - box.y) * (-as);
    const float rot_y = (p.x - box.x) * as + (p.y - box.y) * ac;
    return fabsf(rot_x) < box.w / 2 + MARGIN && fabsf(rot_y) < box.l / 2 + MARGIN;
}
__device__ __forceinline__ bool nms_intersection(F2 p1, F2 p0, F2 q1, F2 q0, F2* ans) {  // helper.h:123-157
    if (!(fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
          fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y)))
        return false;
    const float s1 = nms_cross(q0, p1, p0), s2 = nms_cross(p1, q1, p0), s3 = nms_cross(p0, q1, q0), s4 = nms_cross(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
    const float s5 = nms_cross(q1, p1, p0);
    if (fabsf(s5 - s1) > kNmsEps) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}
__device__ __forceinline__ void nms_rotate(F2 c, float ac, float as, F2* p) {           // helper.h:159-164
    const float nx = (p->x - c.x) * ac + (p->y - c.y) * (-as) + c.x;
    const float ny = (p->x - c.x) * as + (p->y - c.y) * ac + c.y;
    p->x = nx; p->y = ny;
}
__device__ float nms_box_overlap(const NBox& a, const NBox& b) {                        // helper.h:166-255
    const float a_dx = a.w / 2, b_dx = b.w / 2, a_dy = a.l / 2, b_dy = b.l / 2;
    F2 ca[5] = {{a.x - a_dx, a.y - a_dy}, {a.x + a_dx, a.y - a_dy}, {a.x + a_dx, a.y + a_dy}, {a.x - a_dx, a.y + a_dy}, {0.f, 0.f}};
    F2 cb[5] = {{b.x - b_dx, b.y - b_dy}, {b.x + b_dx, b.y - b_dy}, {b.x + b_dx, b.y + b_dy}, {b.x - b_dx, b.y + b_dy}, {0.f, 0.f}};
    const F2 center_a = {a.x, a.y}, center_b = {b.x, b.y};
    F2 pts[16], pc = {0.f, 0.f};
    int cnt = 0;
    const float a_cos = cosf(a.rt), a_sin = sinf(a.rt), b_cos = cosf(b.rt), b_sin = sinf(b.rt);
    for (int k = 0; k < 4; ++k) { nms_rotate(center_a, a_cos, a_sin, &ca[k]); nms_rotate(center_b, b_cos, b_sin, &cb[k]); }
    ca[4] = ca[0]; cb[4] = cb[0];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (nms_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) { pc.x += pts[cnt].x; pc.y += pts[cnt].y; ++cnt; }
    for (int k = 0; k < 4; ++k) {
        if (nms_check_box2d(a, cb[k])) { pc.x += cb[k].x; pc.y += cb[k].y; pts[cnt++] = cb[k]; }
        if (nms_check_box2d(b, ca[k])) { pc.x += ca[k].x; pc.y += ca[k].y; pts[cnt++] = ca[k]; }
    }
    pc.x /= cnt; pc.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(pts[i].y - pc.y, pts[i].x - pc.x) > atan2f(pts[i + 1].y - pc.y, pts[i + 1].x - pc.x)) {
                const F2 t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const F2 u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
        area += (u.x * v.y - u.y * v.x);
    }
    return (float) (fabs((double) area) / 2.0);
}

// workspace per frame (ints): order[Kp] | mask[Kp * Kp / 32]
__host__ __device__ inline size_t nms_ws_words(int K) { const size_t Kp = (size_t) ((K + 31) / 32 * 32); return Kp + Kp * (Kp / 32); }

// 1 CTA per frame: order = box indices by (score descending, index ascending); clears the suppression matrix
__global__ void __launch_bounds__(1024)
nms_sort_kernel(const float* __restrict__ boxes, const int* __restrict__ valid, int* __restrict__ ws, int K)
{
    extern __shared__ unsigned long long s_key[];                 // [np2]
    const int b = blockIdx.x;
    const int Kp = (K + 31) / 32 * 32;
    int* order = ws + (size_t) b * nms_ws_words(K);
    unsigned* mask = reinterpret_cast<unsigned*>(order + Kp);
    int n = valid[b];
    n = n < 0 ? 0 : (n > K ? K : n);
    int np2 = 32;
    while (np2 < K) np2 <<= 1;
    const float* bx = boxes + (size_t) b * K * 9;
    for (int i = threadIdx.x; i < np2; i += blockDim.x)
        s_key[i] = i < n ? (((unsigned long long) ord_key(bx[i * 9 + 8]) << 32) | (unsigned long long) (0xFFFFFFFFu - (unsigned) i)) : 0ull;
    for (int i = threadIdx.x; i < Kp * (Kp / 32); i += blockDim.x) mask[i] = 0u;
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = s_key[i], c = s_key[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? a < c : a > c) { s_key[i] = c; s_key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < Kp; i += blockDim.x)
        order[i] = i < n ? (int) (0xFFFFFFFFu - (unsigned) (s_key[i] & 0xFFFFFFFFull)) : -1;
}

// grid (K, B): row i of the (sorted) suppression matrix, one thread per j > i
__global__ void __launch_bounds__(128)
nms_pairs_kernel(const float* __restrict__ boxes, const int* __restrict__ valid, int* __restrict__ ws, int K, float thresh)
{
    const int b = blockIdx.y, i = blockIdx.x;
    int n = valid[b];
    n = n < 0 ? 0 : (n > K ? K : n);
    if (i >= n) return;
    const int Kp = (K + 31) / 32 * 32;
    const int* order = ws + (size_t) b * nms_ws_words(K);
    unsigned* mask = reinterpret_cast<unsigned*>(const_cast<int*>(order) + Kp);
    const float* bx = boxes + (size_t) b * K * 9;
    const float* pa = bx + (size_t) order[i] * 9;
    const NBox A{pa[0], pa[1], pa[3], pa[4], pa[6]};
    const float sa = A.w * A.l;
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
        const float* pb = bx + (size_t) order[j] * 9;
        const NBox Bx{pb[0], pb[1], pb[3], pb[4], pb[6]};
        const float sb = Bx.w * Bx.l;
        const float so = nms_box_overlap(A, Bx);
        const float iou = so / fmaxf(sa + sb - so, kNmsEps);      // helper.h:272-275
        if (iou >= thresh) atomicOr(mask + (size_t) i * (Kp / 32) + (j >> 5), 1u << (j & 31));
    }
}

// 1 warp per frame: the sequential greedy pass of nms_cpu over the bit matrix, then the output rows
__global__ void __launch_bounds__(32)
nms_scan_kernel(const float* __restrict__ boxes, const int* __restrict__ valid, const int* __restrict__ ws, float* __restrict__ out_boxes,
                int* __restrict__ out_num, int* __restrict__ keep_index, int K, int zero_tails)
{
    __shared__ int s_keep[1024];
    const int b = blockIdx.x, lane = threadIdx.x;
    int n = valid[b];
    n = n < 0 ? 0 : (n > K ? K : n);
    const int Kp = (K + 31) / 32 * 32, words = Kp / 32;           // words <= 32
    const int* order = ws + (size_t) b * nms_ws_words(K);
    const unsigned* mask = reinterpret_cast<const unsigned*>(order + Kp);
    unsigned removed = 0u;                                        // lane w holds word w of the suppressed set
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        const unsigned wv = __shfl_sync(0xffffffffu, removed, i >> 5);
        if (wv >> (i & 31) & 1u) continue;
        if (lane == 0) s_keep[kept] = i;
        ++kept;
        if (lane < words) removed |= mask[(size_t) i * words + lane];
    }
    __syncwarp();
    const float* bx = boxes + (size_t) b * K * 9;
    float* ob = out_boxes + (size_t) b * K * 9;
    for (int t = lane; t < kept * 9; t += 32) {
        const int k = t / 9, c = t - k * 9;
        ob[t] = bx[(size_t) order[s_keep[k]] * 9 + c];
    }
    if (keep_index)
        for (int k = lane; k < K; k += 32) {
            if (k < kept) keep_index[(size_t) b * K + k] = order[s_keep[k]];
            else if (zero_tails) keep_index[(size_t) b * K + k] = 0;
        }
    if (zero_tails)
        for (int t = kept * 9 + lane; t < K * 9; t += 32) ob[t] = 0.f;
    if (lane == 0) out_num[b] = kept;
}

}  // namespace
}  // namespace dsvt

using namespace dsvt;

static int ch_check(const dsvt_center_head_params* p) {
    DSVT_CHECK_ARG(p != nullptr, "params is NULL");
    DSVT_CHECK_ARG(p->batch >= 1 && p->num_classes >= 1 && p->height >= 1 && p->width >= 1, "shape");
    DSVT_CHECK_ARG(p->max_top_k >= 1 && p->max_top_k <= 1024, "max_top_k must be in [1, 1024]");
    DSVT_CHECK_ARG((long long) p->num_classes * p->height * p->width < (1ll << 31), "heat map too large");
    DSVT_CHECK_ARG((long long) p->num_classes * p->height * p->width >= p->max_top_k, "fewer scores than max_top_k");
    return DSVT_OK;
}

extern "C" size_t dsvt_center_head_topk_workspace_size(const dsvt_center_head_params* p) {
    if (ch_check(p) != DSVT_OK) return 0;
    return align_up(pp_ws_words() * 4 * (size_t) p->batch, kWsAlign);
}

extern "C" int dsvt_center_head_topk_launch(const dsvt_center_head_params* p, const float* heatmap, const float* center,
                                            const float* center_z, const float* dim, const float* rot, float* scores,
                                            int32_t* classes, int32_t* xs, int32_t* ys, float* center_g, float* center_z_g,
                                            float* angle, float* dim_g, void* workspace, size_t workspace_bytes,
                                            dsvt_stream_t stream)
{
    int rc = ch_check(p);
    if (rc != DSVT_OK) return rc;
    DSVT_CHECK_ARG(heatmap && center && center_z && dim && rot && scores && classes && xs && ys && center_g && center_z_g && angle &&
                   dim_g && workspace, "NULL tensor pointer");
    DSVT_CHECK_ARG(workspace_bytes >= dsvt_center_head_topk_workspace_size(p) && !((uintptr_t) workspace & 15), "workspace");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int HW = p->height * p->width, N = p->num_classes * HW, B = p->batch, K = p->max_top_k;
    unsigned* ws = static_cast<unsigned*>(workspace);
    // histograms + the candidate counter of every frame (the candidate list itself is overwritten)
    DSVT_CUDA(cudaMemset2DAsync(ws, pp_ws_words() * 4, 0, (2 * kBins + 64) * 4, B, st));
    count_launch();
    const int ctas = sm_count() * 4;
    ph_hist1_kernel<<<dim3(ctas, B), kPpThreads, 0, st>>>(heatmap, ws, N);
    DSVT_LAUNCH_CHECK();
    ph_hist2_kernel<<<dim3(ctas, B), kPpThreads, 0, st>>>(heatmap, ws, N, K);
    DSVT_LAUNCH_CHECK();
    ph_collect_kernel<<<dim3(ctas, B), kPpThreads, 0, st>>>(heatmap, ws, N, K);
    DSVT_LAUNCH_CHECK();
    DSVT_RAISE_SMEM(ph_emit_kernel, kCandCap * 8);
    ph_emit_kernel<<<B, 1024, kCandCap * 8, st>>>(ws, center, center_z, dim, rot, scores, classes, xs, ys, center_g, center_z_g,
                                                  angle, dim_g, HW, p->width, K);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

static int nms_check(const dsvt_nms_params* p) {
    DSVT_CHECK_ARG(p != nullptr, "params is NULL");
    DSVT_CHECK_ARG(p->batch >= 1 && p->max_boxes >= 1 && p->max_boxes <= 1024, "max_boxes must be in [1, 1024]");
    return DSVT_OK;
}

extern "C" size_t dsvt_rotated_nms_workspace_size(const dsvt_nms_params* p) {
    if (nms_check(p) != DSVT_OK) return 0;
    return align_up(nms_ws_words(p->max_boxes) * 4 * (size_t) p->batch, kWsAlign);
}

extern "C" int dsvt_rotated_nms_launch(const dsvt_nms_params* p, const float* boxes, const int32_t* valid, float* out_boxes,
                                       int32_t* out_num, int32_t* keep_index, void* workspace, size_t workspace_bytes,
                                       dsvt_stream_t stream)
{
    int rc = nms_check(p);
    if (rc != DSVT_OK) return rc;
    DSVT_CHECK_ARG(boxes && valid && out_boxes && out_num && workspace, "NULL tensor pointer");
    DSVT_CHECK_ARG(workspace_bytes >= dsvt_rotated_nms_workspace_size(p) && !((uintptr_t) workspace & 15), "workspace");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int K = p->max_boxes, B = p->batch;
    int np2 = 32;
    while (np2 < K) np2 <<= 1;
    int* ws = static_cast<int*>(workspace);
    nms_sort_kernel<<<B, 1024, (size_t) np2 * 8, st>>>(boxes, valid, ws, K);
    DSVT_LAUNCH_CHECK();
    nms_pairs_kernel<<<dim3(K, B), 128, 0, st>>>(boxes, valid, ws, K, p->nms_thresh);
    DSVT_LAUNCH_CHECK();
    nms_scan_kernel<<<B, 32, 0, st>>>(boxes, valid, ws, out_boxes, out_num, keep_index, K, p->zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}
