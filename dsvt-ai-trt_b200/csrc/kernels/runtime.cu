// Library-wide plumbing: error text, launch counter, device queries.
#include "common.cuh"
#include <cstdarg>

namespace dsvt {

std::atomic<uint64_t> g_launch_count{0};
static thread_local char t_err[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace dsvt

extern "C" int dsvt_abi_version(void) { return DSVT_B200_ABI_VERSION; }
extern "C" const char* dsvt_last_error(void) { return dsvt::t_err; }
extern "C" int dsvt_device_sm_count(void) { return dsvt::sm_count(); }
extern "C" uint64_t dsvt_launch_count(void) { return dsvt::g_launch_count.load(); }
