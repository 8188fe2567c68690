// Micro-probe: how fast does a 147 KB weight image arrive in shared memory via cp.async.bulk when all 148 CTAs want
// (a) the same image, (b) different images; in 6 / 24 / 1 pieces; cold vs warm L2.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(544, 1) probe(const uint8_t* src, size_t cta_stride, int pieces, long long* out, int pollers) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ volatile int done;
    const uint32_t total = 147456, piece = total / pieces;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;");
        done = 0;
    }
    __syncthreads();
    if (threadIdx.x >= 32 && (int) (threadIdx.x >> 5) <= pollers) {      // warps 1..pollers poll a barrier that completes at the end
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar2)) : "memory");
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(total) : "memory");
        for (int i = 0; i < pieces; ++i)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + (size_t) i * piece)), "l"(src + blockIdx.x * cta_stride + (size_t) i * piece), "r"(piece), "r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        out[blockIdx.x] = clock64() - t0;
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    }
}
int main() {
    const size_t img = 147456;
    uint8_t* d; cudaMalloc(&d, img * 148);
    cudaMemset(d, 1, img * 148);
    long long* o; cudaMalloc(&o, 148 * 8);
    uint8_t* flush; cudaMalloc(&flush, 256u << 20);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) img);
    std::vector<long long> h(148);
    for (int pollers : {0, 4, 16})
      for (int shared = 1; shared >= 1; --shared)
        for (int pieces : {6})
            for (int warm = 0; warm < 2; ++warm) {
                if (!warm) cudaMemset(flush, 0, 256u << 20);
                else probe<<<148, 544, img>>>(d, shared ? 0 : img, pieces, o, pollers);
                probe<<<148, 544, img>>>(d, shared ? 0 : img, pieces, o, pollers);
                cudaMemcpy(h.data(), o, 148 * 8, cudaMemcpyDeviceToHost);
                long long mx = 0, sum = 0;
                for (auto v : h) { mx = v > mx ? v : mx; sum += v; }
                printf("%2d polling warps, %s image, %2d pieces, %s L2: avg %6lld cycles, max %6lld cycles  (%.1f B/cycle/SM at avg)\n", pollers, shared ? "same     " : "per-CTA  ",
                       pieces, warm ? "warm" : "cold", sum / 148, mx, 147456.0 / (sum / 148.0));
            }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return 0;
}
