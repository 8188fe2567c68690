// (next #4) the narrow first layers of the VFE / position-embedding MLPs, which TensorRT runs as FullyConnected + Scale +
// ReLU (fullyConnectedBnLELU, reference src/dsvt-ai-trt.cpp:268-286):
//   PFN layer 0           Linear(10 -> 96, no bias) + BatchNorm1d + ReLU on the decorated points   (:577)
//   position embedding    Linear(2 -> 192) + BatchNorm1d + ReLU on the in-window coordinates       (:603-637, :461-492)
// K is far too small for a tensor-core tile: the layer is a streaming kernel -- one thread per (row, 4 output channels),
// the row's K inputs read through L1, weights and the folded BatchNorm (scale, shift) in shared memory, 16-byte coalesced
// stores; bound by the HBM write of the output rows.
#include "common.cuh"
#include <new>
#include <vector>

struct dsvt_small_linear {
    int N, K;
    float* blob;      // device: W [N][K] | scale [N] | shift [N]
};

namespace dsvt {
namespace {

constexpr int kMaxK = 16;

template <int K>
__global__ void __launch_bounds__(192)
small_linear_kernel(const float* __restrict__ x, const float* __restrict__ blob, const int* __restrict__ rows,
                    float* __restrict__ y, int max_rows, int N, int act, int zero_tails)
{
    extern __shared__ float s_w[];                  // [N][K] | scale [N] | shift [N]
    for (int t = threadIdx.x; t < N * (K + 2); t += blockDim.x) s_w[t] = blob[t];
    __syncthreads();
    const float* s_scale = s_w + N * K;
    const float* s_shift = s_scale + N;
    const int b = blockIdx.y;
    int R = rows[b];
    R = R < max_rows ? R : max_rows;
    const int n4 = N / 4, rows_per_block = blockDim.x / n4;
    const int c4 = threadIdx.x % n4, rl = threadIdx.x / n4;
    if (rl >= rows_per_block) return;
    const int end = zero_tails ? max_rows : R;
    x += (size_t) b * max_rows * K;
    y += (size_t) b * max_rows * N;
    for (int row = blockIdx.x * rows_per_block + rl; row < end; row += gridDim.x * rows_per_block) {
        float4* dst = reinterpret_cast<float4*>(y + (size_t) row * N) + c4;
        if (row >= R) { stg_zero4(dst); continue; }
        float xin[K];
#pragma unroll
        for (int k = 0; k < K; ++k) xin[k] = __ldg(x + (size_t) row * K + k);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = c4 * 4 + j;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) acc = fmaf(xin[k], s_w[n * K + k], acc);
            acc = fmaf(acc, s_scale[n], s_shift[n]);              // FullyConnected, then the Scale layer of the folded BatchNorm
            o[j] = act == 2 ? fmaxf(acc, 0.f) : acc;
        }
        stg_stream4(dst, make_float4(o[0], o[1], o[2], o[3]));
    }
}

template <int K>
int launch_k(const dsvt_small_linear* w, const float* x, const int* rows, int batch, int max_rows, int act, float* y,
             int zero_tails, cudaStream_t st)
{
    const int n4 = w->N / 4, rows_per_block = 192 / n4;
    const size_t need = ((size_t) max_rows + rows_per_block - 1) / rows_per_block, cap = (size_t) sm_count() * 8;
    const int grid = (int) (need < cap ? need : cap);
    small_linear_kernel<K><<<dim3(grid, batch), 192, (size_t) w->N * (K + 2) * sizeof(float), st>>>(
        x, w->blob, rows, y, max_rows, w->N, act, zero_tails);
    DSVT_LAUNCH_CHECK();
    return DSVT_OK;
}

}  // namespace
}  // namespace dsvt

using namespace dsvt;

extern "C" dsvt_small_linear* dsvt_small_linear_create(int32_t N, int32_t K, const float* W, const float* scale,
                                                       const float* shift)
{
    if (N <= 0 || N % 4 || 192 % (N / 4) || K <= 0 || K > kMaxK || !W) {
        set_last_error("dsvt_small_linear_create: need K in [1,16] and N / 4 dividing 192 (N = 96, 192, ...)");
        return nullptr;
    }
    std::vector<float> host((size_t) N * (K + 2));
    for (size_t i = 0; i < (size_t) N * K; ++i) host[i] = W[i];
    for (int n = 0; n < N; ++n) {
        host[(size_t) N * K + n] = scale ? scale[n] : 1.0f;
        host[(size_t) N * (K + 1) + n] = shift ? shift[n] : 0.0f;
    }
    auto* w = new (std::nothrow) dsvt_small_linear{N, K, nullptr};
    if (!w) return nullptr;
    if (cudaMalloc(&w->blob, host.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(w->blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_last_error("dsvt_small_linear_create: CUDA allocation/copy failed");
        cudaFree(w->blob);
        delete w;
        return nullptr;
    }
    return w;
}

extern "C" void dsvt_small_linear_destroy(dsvt_small_linear* w) {
    if (!w) return;
    cudaFree(w->blob);
    delete w;
}

extern "C" int dsvt_small_linear_launch(const dsvt_small_linear* w, const float* x, const int32_t* rows, int32_t batch,
                                        int32_t max_rows, int32_t activation, float* y, int32_t zero_tails,
                                        dsvt_stream_t stream)
{
    DSVT_CHECK_ARG(w && x && rows && y && batch >= 1 && max_rows >= 1, "NULL argument");
    DSVT_CHECK_ARG(activation == 0 || activation == 2, "activation: 0 none, 2 ReLU");
    DSVT_CHECK_ARG(!((uintptr_t) y & 15), "y must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (w->K) {
#define DSVT_SL_CASE(KK) case KK: return launch_k<KK>(w, x, rows, batch, max_rows, activation, y, zero_tails, st);
        DSVT_SL_CASE(1) DSVT_SL_CASE(2) DSVT_SL_CASE(3) DSVT_SL_CASE(4) DSVT_SL_CASE(5) DSVT_SL_CASE(6) DSVT_SL_CASE(7) DSVT_SL_CASE(8)
        DSVT_SL_CASE(9) DSVT_SL_CASE(10) DSVT_SL_CASE(11) DSVT_SL_CASE(12) DSVT_SL_CASE(13) DSVT_SL_CASE(14) DSVT_SL_CASE(15) DSVT_SL_CASE(16)
#undef DSVT_SL_CASE
        default:
            set_last_error("dsvt_small_linear_launch: K = %d out of range [1, 16]", w->K);
            return DSVT_ERR_UNSUPPORTED;
    }
}
