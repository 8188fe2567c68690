// Dense linear layer on the 5th-generation tensor cores: Y[M,N] = X[M,K] * W[N,K]^T + b  (FP32 in/out).
// This is the building block the set-attention kernel's projections use, exposed on its own as the
// replacement of the TensorRT FullyConnected layers around the plugins (FFN linears).
//
// One CTA owns a 128-row tile of X.  Rows are converted (FP16 or TF32) and staged by the threads into the
// UMMA K-major interleaved layout (tc_common.cuh); W tiles are pre-arranged on the host in that same order
// and arrive with one cp.async.bulk per N tile; one elected thread issues tcgen05.mma over K with the
// accumulator in TMEM; all four warps read their TMEM lane quarter back with tcgen05.ld, add the bias and
// store.
#include "common.cuh"
#include "tc_common.cuh"
#include "attention_common.cuh"
#include <cuda_fp16.h>
#include <vector>
#include <cmath>
#include <cstring>
#include <new>

namespace dsvt {
namespace {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kTileN = 64;

template <int ESIZE> struct Elem;
template <> struct Elem<2> { static constexpr int per_chunk = 8; static constexpr uint32_t fmt = kFmtF16; };
template <> struct Elem<4> { static constexpr int per_chunk = 4; static constexpr uint32_t fmt = kFmtTF32; };

// smem: A tile [K/per_chunk chunks][128 rows][16 B] | B tile [chunks][64 rows][16 B] | barriers
template <int ESIZE>
__global__ void __launch_bounds__(128, 1)
tc_linear_kernel(const float* __restrict__ X, const uint8_t* __restrict__ Wimg, const float* __restrict__ bias,
                 float* __restrict__ Y, int M, int N, int K)
{
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int EPC = Elem<ESIZE>::per_chunk;
    const int chunks = K / EPC;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t) chunks * kTileM * 16;
    __shared__ __align__(8) uint64_t bar_b, bar_mma;
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.x * kTileM;

    if (tid == 0) {
        mbar_init(&bar_b, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<64>(&tmem_base_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_slot;

    // ---- stage A: thread t handles (row = t, all chunks): lanes of a warp write 32 consecutive rows of a chunk
    for (int c = 0; c < chunks; ++c) {
        const int r = tid;
        const int grow = row0 + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (grow < M) {
            const float* src = X + (size_t) grow * K + c * EPC;
            if (ESIZE == 4) {
                const float4 f = *reinterpret_cast<const float4*>(src);
                v.x = __float_as_uint(to_tf32(f.x)); v.y = __float_as_uint(to_tf32(f.y));
                v.z = __float_as_uint(to_tf32(f.z)); v.w = __float_as_uint(to_tf32(f.w));
            } else {
                const float4 f0 = *reinterpret_cast<const float4*>(src);
                const float4 f1 = *reinterpret_cast<const float4*>(src + 4);
                __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f0.z, f0.w);
                __half2 h2 = __floats2half2_rn(f1.x, f1.y), h3 = __floats2half2_rn(f1.z, f1.w);
                v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
                v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
            }
        }
        *reinterpret_cast<uint4*>(sA + ((size_t) c * kTileM + r) * 16) = v;
    }
    fence_proxy_async_smem();
    __syncthreads();

    const uint32_t idesc = make_idesc(Elem<ESIZE>::fmt, kTileM, kTileN);
    const uint32_t b_tile_bytes = (uint32_t) chunks * kTileN * 16;
    uint32_t phase = 0;
    for (int n0 = 0; n0 < N; n0 += kTileN) {
        if (tid == 0) {
            mbar_arrive_expect_tx(&bar_b, b_tile_bytes);
            bulk_g2s(sB, Wimg + (size_t) (n0 / kTileN) * b_tile_bytes, b_tile_bytes, &bar_b);
            mbar_wait(&bar_b, phase);
            tc_fence_after_sync();
            const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
            const int ksteps = chunks / 2;
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t ad = make_smem_desc(a_addr + ks * 2 * kTileM * 16, kTileM * 16, 128);
                const uint64_t bd = make_smem_desc(b_addr + ks * 2 * kTileN * 16, kTileN * 16, 128);
                if (ESIZE == 4) umma_tf32(tmem, ad, bd, idesc, ks > 0);
                else umma_f16(tmem, ad, bd, idesc, ks > 0);
            }
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, phase);
        tc_fence_after_sync();
        phase ^= 1;
        // ---- epilogue: warp w reads lanes [32w, 32w+32), 64 columns
        const int grow = row0 + warp * 32 + lane;
        uint32_t r[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            tmem_ld32(tmem + ((uint32_t) (warp * 32) << 16) + half * 32, r);
            tmem_ld_wait();
            if (grow < M) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int n = n0 + half * 32 + j;
                    float4 o;
                    o.x = __uint_as_float(r[j + 0]) + bias[n + 0];
                    o.y = __uint_as_float(r[j + 1]) + bias[n + 1];
                    o.z = __uint_as_float(r[j + 2]) + bias[n + 2];
                    o.w = __uint_as_float(r[j + 3]) + bias[n + 3];
                    *reinterpret_cast<float4*>(Y + (size_t) grow * N + n) = o;
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();      // TMEM accumulator and sB are reused by the next N tile
        tc_fence_after_sync();
    }
    if (warp == 0) tmem_dealloc<64>(tmem);
}

}  // namespace
}  // namespace dsvt

using namespace dsvt;

namespace dsvt {   // attention_split.cu: the persistent split-precision GEMM as a plain linear layer
void* linear_split_prepare(int N, int K, const float* W, const float* b, float* out_mul);
int linear_split_launch(const void* blob, int N, int K, float out_mul, bool split, int act, const float* x,
                        const float* x_hi, int k_split, const int* rows_dev, int rows_host, int max_rows, float* y,
                        int zero_tails, cudaStream_t st);
int linear_split_k_launch(const void* blob, int N, int K, float out_mul, bool split, const float* x, const float* add,
                          const int* rows_dev, int max_rows, float* y_parts, cudaStream_t st);
int linear_gen_launch(const void* blob, float out_mul, bool split, const float* x2, const float* small_blob, const int* rows_dev,
                      int max_rows, float* y, int zero_tails, cudaStream_t st);
int linear_ln_launch(const void* blob, int K, float out_mul, bool split, const float* x, const int* rows_dev, int max_rows,
                     int n_ln, const float* const* res, const float* const* gamma, const float* const* beta, float eps,
                     float* y, int zero_tails, cudaStream_t st);
int linear_gen_batch_launch(int n, const void* const* blobs, const float* out_muls, bool split, const float* const* x2s,
                            const float* const* small_blobs, const int* rows_dev, int max_rows, float* const* ys, int zero_tails,
                            cudaStream_t st);
void* ffn_w1_pieces_prepare(const float* W);
int attn_ffn_fused_launch(const dsvt_set_attention_params* p, const void* attn_blob, const float* attn_out_mul, const void* plan,
                          void* workspace, size_t workspace_bytes, const float* x_res, const float* gamma1, const float* beta1,
                          float eps1, const void* blob1, const void* pieces1, float out_mul1, const void* blob2, float out_mul2,
                          const int* rows_dev, int n_ln, const float* const* res, const float* const* gamma, const float* const* beta,
                          float eps, float* src_out, float* y, int zero_tails, cudaStream_t st);
size_t vfe_fused_workspace(int max_points, int npv);
int vfe_fused_launch(const float* pfn0_blob, const void* blob1, float out_mul1, const float* point_features, const int* piv,
                     const int* voxel_num, const int* point_num, int max_points, int max_pillars, int npv, float* max_voxel,
                     void* workspace, int zero_tails, cudaStream_t st);
int ffn_fused_launch(const void* blob1, const void* pieces1, float out_mul1, const void* blob2, float out_mul2, const float* x,
                     const int* rows_dev, int max_rows, int n_ln, const float* const* res, const float* const* gamma,
                     const float* const* beta, float eps, float* y, int zero_tails, cudaStream_t st);
}

struct dsvt_linear_weights {
    int N, K, precision;
    uint8_t* img;     // device: [N/64 tiles][K chunks][64 rows][16 B]           (TF32 / FP16 single-tile kernel)
    float* bias;      // device [N]
    void* split_blob; // device: 192 x 192 block images + bias                   (FP32_TC / FP16_GEMM persistent kernel)
    float out_mul;
    void* piece_blob; // device: a 192 -> 384 FP32_TC layer as 64-column pieces  (first layer of dsvt_ffn_fused_launch)
};

static uint16_t f32_to_f16_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
static uint32_t f32_to_tf32_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u += 0x1000u;   // round to nearest (ties away), like cvt.rna.tf32.f32
    return u & 0xFFFFE000u;
}

extern "C" dsvt_linear_weights* dsvt_linear_weights_create(int32_t N, int32_t K, const float* W, const float* b,
                                                           int32_t precision)
{
    if (precision == DSVT_ATTN_FP32_TC || precision == DSVT_ATTN_FP16_GEMM) {
        if (N <= 0 || K <= 0 || N % 192 || K % 192 || !W) {
            set_last_error("dsvt_linear_weights_create: the FP32_TC / FP16_GEMM linear layer needs N %% 192 == 0 and K %% 192 == 0");
            return nullptr;
        }
        auto* lw = new (std::nothrow) dsvt_linear_weights{N, K, precision, nullptr, nullptr, nullptr, 1.0f, nullptr};
        if (!lw) return nullptr;
        lw->split_blob = dsvt::linear_split_prepare(N, K, W, b, &lw->out_mul);
        if (!lw->split_blob) {
            set_last_error("dsvt_linear_weights_create: CUDA allocation/copy failed");
            delete lw;
            return nullptr;
        }
        if (precision == DSVT_ATTN_FP32_TC && N == 384 && K == 192) {
            lw->piece_blob = dsvt::ffn_w1_pieces_prepare(W);
            if (!lw->piece_blob) {
                set_last_error("dsvt_linear_weights_create: CUDA allocation/copy failed");
                cudaFree(lw->split_blob);
                delete lw;
                return nullptr;
            }
        }
        return lw;
    }
    const int esize = precision == DSVT_ATTN_FP16 ? 2 : 4;
    const int epc = 16 / esize;
    if (N <= 0 || K <= 0 || N % kTileN || K % (2 * epc) || !W || (precision != DSVT_ATTN_FP16 && precision != DSVT_LINEAR_TF32)) {
        set_last_error("dsvt_linear_weights_create: need N %% 64 == 0, K %% %d == 0, precision TF32 or FP16", 2 * epc);
        return nullptr;
    }
    const int chunks = K / epc;
    std::vector<uint8_t> img((size_t) N * K * esize);
    for (int n = 0; n < N; ++n) {
        const int tile = n / kTileN, nl = n % kTileN;
        for (int c = 0; c < chunks; ++c) {
            uint8_t* dst = img.data() + (((size_t) tile * chunks + c) * kTileN + nl) * 16;
            for (int e = 0; e < epc; ++e) {
                const float w = W[(size_t) n * K + c * epc + e];
                if (esize == 2) { uint16_t h = f32_to_f16_bits(w); memcpy(dst + e * 2, &h, 2); }
                else { uint32_t t = f32_to_tf32_bits(w); memcpy(dst + e * 4, &t, 4); }
            }
        }
    }
    auto* lw = new (std::nothrow) dsvt_linear_weights{N, K, precision, nullptr, nullptr, nullptr, 1.0f, nullptr};
    if (!lw) return nullptr;
    std::vector<float> zero(N, 0.f);
    if (cudaMalloc(&lw->img, img.size()) != cudaSuccess || cudaMalloc(&lw->bias, N * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(lw->img, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(lw->bias, b ? b : zero.data(), N * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_last_error("dsvt_linear_weights_create: CUDA allocation/copy failed");
        cudaFree(lw->img); cudaFree(lw->bias);
        delete lw;
        return nullptr;
    }
    return lw;
}

extern "C" void dsvt_linear_weights_destroy(dsvt_linear_weights* w) {
    if (!w) return;
    cudaFree(w->img);
    cudaFree(w->bias);
    cudaFree(w->split_blob);
    cudaFree(w->piece_blob);
    delete w;
}

extern "C" int dsvt_linear_launch(const dsvt_linear_weights* w, const float* x, int32_t M, float* y, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(w && x && y && M >= 0, "NULL argument");
    DSVT_CHECK_ARG(!(((uintptr_t) x | (uintptr_t) y) & 15), "16-B alignment");
    if (M == 0) return DSVT_OK;
    if (w->split_blob) {
        DSVT_CHECK_ARG(!((uintptr_t) x & 31), "32-B alignment of x (256-bit loads)");
        return dsvt::linear_split_launch(w->split_blob, w->N, w->K, w->out_mul, w->precision == DSVT_ATTN_FP32_TC, 0, x,
                                         nullptr, 0, nullptr, M, M, y, 0, reinterpret_cast<cudaStream_t>(stream));
    }
    const int esize = w->precision == DSVT_ATTN_FP16 ? 2 : 4;
    const size_t smem = (size_t) w->K * esize * (kTileM + kTileN);
    DSVT_CHECK_ARG(smem <= 220 * 1024, "K too large for a single-stage tile");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (M + kTileM - 1) / kTileM;
    if (esize == 4) {
        DSVT_CUDA(cudaFuncSetAttribute(tc_linear_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        tc_linear_kernel<4><<<grid, 128, smem, st>>>(x, w->img, w->bias, y, M, w->N, w->K);
    } else {
        DSVT_CUDA(cudaFuncSetAttribute(tc_linear_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        tc_linear_kernel<2><<<grid, 128, smem, st>>>(x, w->img, w->bias, y, M, w->N, w->K);
    }
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

extern "C" int dsvt_linear_rows_launch(const dsvt_linear_weights* w, const float* x, const int32_t* rows, int32_t max_rows,
                                       int32_t activation, float* y, int32_t zero_tails, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(w && x && y && rows && max_rows >= 1, "NULL argument");
    DSVT_CHECK_ARG(w->split_blob != nullptr, "built for precision DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM weights only");
    DSVT_CHECK_ARG(activation >= 0 && activation <= 2, "activation: 0 none, 1 GELU, 2 ReLU");
    DSVT_CHECK_ARG(!(((uintptr_t) x & 31) | ((uintptr_t) y & 15)), "alignment (x 32 B, y 16 B)");
    return dsvt::linear_split_launch(w->split_blob, w->N, w->K, w->out_mul, w->precision == DSVT_ATTN_FP32_TC, activation, x,
                                     nullptr, 0, rows, 0, max_rows, y, zero_tails, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dsvt_linear_rows_concat_launch(const dsvt_linear_weights* w, const float* x_lo, const float* x_hi,
                                              int32_t k_split, const int32_t* rows, int32_t max_rows, int32_t activation,
                                              float* y, int32_t zero_tails, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(w && x_lo && x_hi && y && rows && max_rows >= 1, "NULL argument");
    DSVT_CHECK_ARG(w->split_blob != nullptr, "built for precision DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM weights only");
    DSVT_CHECK_ARG(w->K == 192, "the in-place concatenation is built for K == 192 layers");
    DSVT_CHECK_ARG(k_split > 0 && k_split < w->K && k_split % 32 == 0, "k_split must be a multiple of 32 inside (0, K)");
    DSVT_CHECK_ARG(activation >= 0 && activation <= 2, "activation: 0 none, 1 GELU, 2 ReLU");
    DSVT_CHECK_ARG(!((((uintptr_t) x_lo | (uintptr_t) x_hi) & 31) | ((uintptr_t) y & 15)), "alignment (x 32 B, y 16 B)");
    return dsvt::linear_split_launch(w->split_blob, w->N, w->K, w->out_mul, w->precision == DSVT_ATTN_FP32_TC, activation, x_lo,
                                     x_hi, k_split, rows, 0, max_rows, y, zero_tails, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dsvt_linear_rows_splitk_launch(const dsvt_linear_weights* w, const float* x, const float* add,
                                              const int32_t* rows, int32_t max_rows, float* y_parts, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(w && x && y_parts && rows && max_rows >= 1, "NULL argument");
    DSVT_CHECK_ARG(w->split_blob != nullptr, "built for precision DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM weights only");
    DSVT_CHECK_ARG(w->N == 192 && w->K >= 192 && w->K <= 576, "split-K form: N == 192, K in {192, 384, 576}");
    DSVT_CHECK_ARG(!(((uintptr_t) x & 31) | ((uintptr_t) y_parts & 15) | ((uintptr_t) add & 15)), "alignment (x 32 B, y / add 16 B)");
    return dsvt::linear_split_k_launch(w->split_blob, w->N, w->K, w->out_mul, w->precision == DSVT_ATTN_FP32_TC, x, add, rows,
                                       max_rows, y_parts, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dsvt_linear_rows_norm_launch(const dsvt_linear_weights* w, const float* x, const int32_t* rows, int32_t max_rows,
                                            const dsvt_ln_stage* stages, int32_t n_stages, float eps, float* y,
                                            int32_t zero_tails, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(w && x && y && rows && stages && max_rows >= 1, "NULL argument");
    DSVT_CHECK_ARG(w->split_blob != nullptr, "built for precision DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM weights only");
    DSVT_CHECK_ARG(w->N == 192 && (w->K == 192 || w->K == 384), "linear + LayerNorm chain: N == 192, K in {192, 384}");
    DSVT_CHECK_ARG(n_stages >= 1 && n_stages <= 3, "1..3 LayerNorm stages");
    DSVT_CHECK_ARG(!(((uintptr_t) x & 31) | ((uintptr_t) y & 15)), "alignment (x 32 B, y 16 B)");
    const float* res[3] = {nullptr, nullptr, nullptr};
    const float* gamma[3] = {nullptr, nullptr, nullptr};
    const float* beta[3] = {nullptr, nullptr, nullptr};
    for (int s = 0; s < n_stages; ++s) {
        DSVT_CHECK_ARG(stages[s].gamma && stages[s].beta, "NULL gamma / beta");
        DSVT_CHECK_ARG(!(((uintptr_t) stages[s].residual | (uintptr_t) stages[s].gamma | (uintptr_t) stages[s].beta) & 15), "16-B alignment");
        res[s] = stages[s].residual; gamma[s] = stages[s].gamma; beta[s] = stages[s].beta;
    }
    return dsvt::linear_ln_launch(w->split_blob, w->K, w->out_mul, w->precision == DSVT_ATTN_FP32_TC, x, rows, max_rows,
                                  n_stages, res, gamma, beta, eps, y, zero_tails, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dsvt_ffn_fused_launch(const dsvt_linear_weights* fc1, const dsvt_linear_weights* fc2, const float* x,
                                     const int32_t* rows, int32_t max_rows, const dsvt_ln_stage* stages, int32_t n_stages,
                                     float eps, float* y, int32_t zero_tails, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(fc1 && fc2 && x && y && rows && stages && max_rows >= 1, "NULL argument");
    DSVT_CHECK_ARG(fc1->piece_blob != nullptr && fc1->N == 384 && fc1->K == 192,
                   "first layer: Linear(192 -> 384) created with DSVT_ATTN_FP32_TC");
    DSVT_CHECK_ARG(fc2->split_blob != nullptr && fc2->precision == DSVT_ATTN_FP32_TC && fc2->N == 192 && fc2->K == 384,
                   "second layer: Linear(384 -> 192) created with DSVT_ATTN_FP32_TC");
    DSVT_CHECK_ARG(n_stages >= 1 && n_stages <= 3, "1..3 LayerNorm stages");
    DSVT_CHECK_ARG(!(((uintptr_t) x & 31) | ((uintptr_t) y & 15)), "alignment (x 32 B, y 16 B)");
    const float* res[3] = {nullptr, nullptr, nullptr};
    const float* gamma[3] = {nullptr, nullptr, nullptr};
    const float* beta[3] = {nullptr, nullptr, nullptr};
    for (int s = 0; s < n_stages; ++s) {
        DSVT_CHECK_ARG(stages[s].gamma && stages[s].beta, "NULL gamma / beta");
        DSVT_CHECK_ARG(!(((uintptr_t) stages[s].residual | (uintptr_t) stages[s].gamma | (uintptr_t) stages[s].beta) & 15), "16-B alignment");
        res[s] = stages[s].residual; gamma[s] = stages[s].gamma; beta[s] = stages[s].beta;
    }
    return dsvt::ffn_fused_launch(fc1->split_blob, fc1->piece_blob, fc1->out_mul, fc2->split_blob, fc2->out_mul, x, rows, max_rows,
                                  n_stages, res, gamma, beta, eps, y, zero_tails, reinterpret_cast<cudaStream_t>(stream));
}

// small_linear.cu
struct dsvt_small_linear { int N, K; float* blob; };

extern "C" int dsvt_pos_embed_mlp_launch(const dsvt_small_linear* first, const dsvt_linear_weights* second, const float* x2,
                                         const int32_t* rows, int32_t max_rows, float* y, int32_t zero_tails,
                                         dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(first && second && x2 && rows && y && max_rows >= 1, "NULL argument");
    DSVT_CHECK_ARG(first->K == 2 && first->N == 192, "first layer: Linear(2 -> 192) (+ folded BatchNorm, ReLU)");
    DSVT_CHECK_ARG(second->split_blob != nullptr && second->N == 192 && second->K == 192,
                   "second layer: Linear(192 -> 192) created with DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM");
    DSVT_CHECK_ARG(!(((uintptr_t) x2 & 7) | ((uintptr_t) y & 15)), "alignment (x2 8 B, y 16 B)");
    return dsvt::linear_gen_launch(second->split_blob, second->out_mul, second->precision == DSVT_ATTN_FP32_TC, x2, first->blob,
                                   rows, max_rows, y, zero_tails, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dsvt_attention_tail_ffn_launch(const dsvt_set_attention_params* p, const dsvt_attention_weights* w, const void* plan,
                                              void* workspace, size_t workspace_bytes, const int32_t* voxel_num, const float* x,
                                              const float* norm1_gamma, const float* norm1_beta, float norm1_eps,
                                              const dsvt_linear_weights* fc1, const dsvt_linear_weights* fc2,
                                              const dsvt_ln_stage* stages, int32_t n_stages, float eps, float* src, float* y,
                                              dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(p && w && plan && workspace && voxel_num && x && norm1_gamma && norm1_beta && fc1 && fc2 && stages && src && y,
                   "NULL argument");
    DSVT_CHECK_ARG(p->precision == DSVT_ATTN_FP32_TC && w->split_blob != nullptr, "built for precision DSVT_ATTN_FP32_TC");
    DSVT_CHECK_ARG(p->channel_num == 192 && w->channel_num == 192, "channel_num 192");
    DSVT_CHECK_ARG(fc1->piece_blob != nullptr && fc1->N == 384 && fc1->K == 192,
                   "first FFN layer: Linear(192 -> 384) created with DSVT_ATTN_FP32_TC");
    DSVT_CHECK_ARG(fc2->split_blob != nullptr && fc2->precision == DSVT_ATTN_FP32_TC && fc2->N == 192 && fc2->K == 384,
                   "second FFN layer: Linear(384 -> 192) created with DSVT_ATTN_FP32_TC");
    DSVT_CHECK_ARG(n_stages >= 1 && n_stages <= 3, "1..3 LayerNorm stages (the first one's residual is `src`)");
    DSVT_CHECK_ARG(!(((uintptr_t) x | (uintptr_t) src | (uintptr_t) y | (uintptr_t) norm1_gamma | (uintptr_t) norm1_beta) & 15), "16-B alignment");
    const float* res[3] = {nullptr, nullptr, nullptr};
    const float* gamma[3] = {nullptr, nullptr, nullptr};
    const float* beta[3] = {nullptr, nullptr, nullptr};
    for (int s = 0; s < n_stages; ++s) {
        DSVT_CHECK_ARG(stages[s].gamma && stages[s].beta, "NULL gamma / beta");
        DSVT_CHECK_ARG(!(((uintptr_t) stages[s].residual | (uintptr_t) stages[s].gamma | (uintptr_t) stages[s].beta) & 15), "16-B alignment");
        res[s] = stages[s].residual; gamma[s] = stages[s].gamma; beta[s] = stages[s].beta;
    }
    return dsvt::attn_ffn_fused_launch(p, w->split_blob, w->split_out_mul, plan, workspace, workspace_bytes, x, norm1_gamma, norm1_beta,
                                       norm1_eps, fc1->split_blob, fc1->piece_blob, fc1->out_mul, fc2->split_blob, fc2->out_mul,
                                       voxel_num, n_stages, res, gamma, beta, eps, src, y, p->zero_tails,
                                       reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dsvt_pos_embed_mlp_batch_launch(const dsvt_small_linear* const* firsts, const dsvt_linear_weights* const* seconds,
                                               const float* const* x2s, int32_t n, const int32_t* rows, int32_t max_rows,
                                               float* const* ys, int32_t zero_tails, dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(firsts && seconds && x2s && ys && rows && max_rows >= 1, "NULL argument");
    DSVT_CHECK_ARG(n >= 1 && n <= 8, "1..8 MLPs per launch");
    const void* blobs[8];
    const float* small_blobs[8];
    float out_muls[8];
    for (int i = 0; i < n; ++i) {
        DSVT_CHECK_ARG(firsts[i] && seconds[i] && x2s[i] && ys[i], "NULL entry");
        DSVT_CHECK_ARG(firsts[i]->K == 2 && firsts[i]->N == 192, "first layer: Linear(2 -> 192) (+ folded BatchNorm, ReLU)");
        DSVT_CHECK_ARG(seconds[i]->split_blob != nullptr && seconds[i]->N == 192 && seconds[i]->K == 192 &&
                       seconds[i]->precision == seconds[0]->precision,
                       "second layer: Linear(192 -> 192) created with DSVT_ATTN_FP32_TC / DSVT_ATTN_FP16_GEMM (one precision per launch)");
        DSVT_CHECK_ARG(!(((uintptr_t) x2s[i] & 7) | ((uintptr_t) ys[i] & 15)), "alignment (x2 8 B, y 16 B)");
        blobs[i] = seconds[i]->split_blob; small_blobs[i] = firsts[i]->blob; out_muls[i] = seconds[i]->out_mul;
    }
    return dsvt::linear_gen_batch_launch(n, blobs, out_muls, seconds[0]->precision == DSVT_ATTN_FP32_TC, x2s, small_blobs, rows,
                                         max_rows, ys, zero_tails, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" size_t dsvt_vfe_fused_workspace_size(int32_t max_points_num, int32_t max_num_points_per_voxel) {
    if (max_points_num < 1 || max_num_points_per_voxel < 1 || max_num_points_per_voxel > 64) return 0;
    return dsvt::vfe_fused_workspace(max_points_num, max_num_points_per_voxel);
}

extern "C" int dsvt_vfe_fused_launch(const dsvt_small_linear* pfn0, const dsvt_linear_weights* pfn1, const float* point_features,
                                     const int32_t* point_index_in_voxel, const int32_t* voxel_num, const int32_t* point_num,
                                     int32_t max_points_num, int32_t max_pillars_num, int32_t max_num_points_per_voxel,
                                     float* voxel_features, void* workspace, size_t workspace_bytes, int32_t zero_tails,
                                     dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(pfn0 && pfn1 && point_features && point_index_in_voxel && voxel_num && point_num && voxel_features && workspace,
                   "NULL argument");
    DSVT_CHECK_ARG(pfn0->K == 10 && pfn0->N == 96, "PFN layer 0: Linear(10 -> 96) (+ folded BatchNorm, ReLU)");
    DSVT_CHECK_ARG(pfn1->split_blob != nullptr && pfn1->precision == DSVT_ATTN_FP32_TC && pfn1->N == 192 && pfn1->K == 192,
                   "PFN layer 1: Linear(192 -> 192) created with DSVT_ATTN_FP32_TC (BatchNorm folded)");
    DSVT_CHECK_ARG(max_points_num >= 1 && max_pillars_num >= 1 && max_num_points_per_voxel >= 1 && max_num_points_per_voxel <= 64,
                   "capacities");
    DSVT_CHECK_ARG(workspace_bytes >= dsvt::vfe_fused_workspace(max_points_num, max_num_points_per_voxel), "workspace too small");
    DSVT_CHECK_ARG(!(((uintptr_t) point_features & 7) | ((uintptr_t) voxel_features & 15) | ((uintptr_t) workspace & 3)), "alignment");
    return dsvt::vfe_fused_launch(pfn0->blob, pfn1->split_blob, pfn1->out_mul, point_features, point_index_in_voxel, voxel_num,
                                  point_num, max_points_num, max_pillars_num, max_num_points_per_voxel, voxel_features, workspace,
                                  zero_tails, reinterpret_cast<cudaStream_t>(stream));
}
