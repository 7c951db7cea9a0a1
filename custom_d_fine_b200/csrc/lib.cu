// Library-level entry points: error string, version, device probe.
#include "common.cuh"
#include <cstdarg>
#include <cstring>

static thread_local char g_err[512] = "";

void dfine_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

DFINE_API const char* dfine_last_error(void) { return g_err; }
DFINE_API int dfine_abi_version(void) { return 1; }

// 0 if the current device can run this library (compute capability 10.x), else negative + error string.
DFINE_API int dfine_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { dfine_set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return -2; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) { dfine_set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return -2; }
    if (p.major != 10) { dfine_set_error("libdfine_sm100 needs an sm_100 device, found sm_%d%d", p.major, p.minor); return -3; }
    return 0;
}
