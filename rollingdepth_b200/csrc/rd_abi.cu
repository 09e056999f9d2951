// librd_b200: version / error plumbing of the C ABI (include/rd_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace rd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return RD_ERR_CUDA;
}

int current_device() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int num_sms() {
    static int cached[kMaxDevices] = {};
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDevices) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace rd

extern "C" {

int rd_version(void) { return RD_B200_ABI_VERSION; }

const char* rd_last_error(void) { return rd::g_err; }

int rd_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return rd::cuda_fail(e, "cudaGetDevice");
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return rd::cuda_fail(e, "cudaDeviceGetAttribute");
    e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) return rd::cuda_fail(e, "cudaDeviceGetAttribute");
    return major * 10 + minor;
}

}  // extern "C"
