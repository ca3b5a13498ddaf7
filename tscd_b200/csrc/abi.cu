// Library identification entry points of libtscd_b200.so.
#include <stdio.h>
#include "common.cuh"

extern "C" const char* tscd_version(void) { return "tscd_b200 0.1 (sm_100a)"; }

extern "C" int tscd_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

static char g_last_err[256] = "";
extern "C" void tscd_set_last_cuda_error(int code, const char* where) {
    snprintf(g_last_err, sizeof(g_last_err), "%s (%d) at %s", cudaGetErrorString((cudaError_t)code), code, where);
}
extern "C" const char* tscd_last_cuda_error(void) { return g_last_err; }
