// Shared helpers for the sm_100a kernels of the DSVT hot path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include "dsvt_b200.h"

namespace dsvt {

// --- error plumbing ---------------------------------------------------------
void set_last_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

inline void count_launch(int n = 1) { g_launch_count.fetch_add((uint64_t) n, std::memory_order_relaxed); }

#define DSVT_CHECK_ARG(cond, msg)                                                  \
    do {                                                                           \
        if (!(cond)) {                                                             \
            ::dsvt::set_last_error("%s: invalid argument: %s", __func__, msg);     \
            return DSVT_ERR_INVALID_ARGUMENT;                                      \
        }                                                                          \
    } while (0)

#define DSVT_CUDA(call)                                                            \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            ::dsvt::set_last_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, \
                                   cudaGetErrorString(e__));                       \
            return DSVT_ERR_CUDA;                                                  \
        }                                                                          \
    } while (0)

#define DSVT_LAUNCH_CHECK()                                                        \
    do {                                                                           \
        ::dsvt::count_launch();                                                    \
        DSVT_CUDA(cudaGetLastError());                                             \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
constexpr size_t kWsAlign = 256;  // the reference aligns workspace sub-buffers to 256 B (points2Features.cu:70-102)

// carve consecutive 256-B aligned sub-buffers out of a workspace
struct WsCarver {
    uint8_t* base;
    size_t off = 0;
    explicit WsCarver(void* p) : base(static_cast<uint8_t*>(p)) {}
    template <typename T> T* take(size_t n) {
        T* r = reinterpret_cast<T*>(base + off);
        off += align_up(n * sizeof(T), kWsAlign);
        return r;
    }
};

int sm_count();

// One-time, PER-DEVICE raise of a kernel's dynamic shared-memory limit: cudaFuncSetAttribute acts on the current device's
// context, so the "done" state is a bit per device ordinal, not a process-wide flag.  Thread-safe without a lock: the
// attribute call is idempotent, two racing first calls on a device merely both make it.  Never allocates or synchronises,
// so it is legal inside a stream capture.
struct PerDeviceOnce {
    std::atomic<uint64_t> done{0};
    template <typename K> int raise_smem(K kernel, int bytes, bool max_carveout = false) {
        int dev = 0;
        DSVT_CUDA(cudaGetDevice(&dev));
        const uint64_t bit = 1ull << (dev & 63);
        if (done.load(std::memory_order_acquire) & bit) return DSVT_OK;
        DSVT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if (max_carveout)
            DSVT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared));
        done.fetch_or(bit, std::memory_order_release);
        return DSVT_OK;
    }
};
#define DSVT_RAISE_SMEM(kernel, bytes)                                             \
    do {                                                                           \
        static ::dsvt::PerDeviceOnce once__;                                       \
        const int rc__ = once__.raise_smem(kernel, (int) (bytes));                 \
        if (rc__ != DSVT_OK) return rc__;                                          \
    } while (0)

// --- device helpers ---------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Zero-fill of output tails (rows beyond the valid count: the reference memsets every output in full).  Nobody reads these
// lines again, so they are written with an L2 evict-first policy and do not push the rows the next kernel needs out of L2.
// DSVT_NO_EVICT_FIRST builds the plain streaming store (A/B runs).
__device__ __forceinline__ void stg_zero4(float4* p) {
#ifdef DSVT_NO_EVICT_FIRST
    stg_stream4(p, make_float4(0.f, 0.f, 0.f, 0.f));
#else
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%1,%1,%1}, %2;"
                 :: "l"(p), "f"(0.f), "l"(pol) : "memory");
#endif
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// `warp_sums` must hold 32 ints of shared memory.  Returns the exclusive prefix; *total = block sum.
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int incl = warp_incl_scan(v, lane);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int s = lane < nw ? warp_sums[lane] : 0;
        int si = warp_incl_scan(s, lane);
        warp_sums[lane] = si - s;           // exclusive warp offsets
        if (lane == 31) warp_sums[32] = si; // total (needs 33 ints!)
    }
    __syncthreads();
    int r = warp_sums[wid] + incl - v;
    *total = warp_sums[32];
    __syncthreads();
    return r;
}

}  // namespace dsvt
