// Library plumbing: thread-local error string, device query.
#include <stdarg.h>
#include <string.h>

#include "dd_common.cuh"

namespace dd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached_dev = dev;
        cached_sms = n;
    }
    return cached_sms;
}

}  // namespace dd

extern "C" {

int dd_abi_version(void) { return DD_ABI_VERSION; }

const char* dd_last_error(void) { return dd::g_err; }

int dd_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    DD_CUDA_OK(cudaGetDevice(&dev));
    int sms = 0, maj = 0, min = 0;
    DD_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DD_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    DD_CUDA_OK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    DD_REQUIRE(maj == 10, DD_EUNSUPPORTED, "libdistdiff_sm100 is built for sm_100a only; device is sm_%d%d", maj, min);
    return 0;
}

}  // extern "C"
