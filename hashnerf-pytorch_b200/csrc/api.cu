// api.cu -- library-level entry points and error bookkeeping.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace hn {

static thread_local char g_last_error[512] = "";

int fail(int code, const char* what) {
  if (code > 0) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", what,
             cudaGetErrorString((cudaError_t)code), cudaGetErrorName((cudaError_t)code));
  } else {
    snprintf(g_last_error, sizeof(g_last_error), "invalid argument: %s", what);
  }
  return code;
}

int check_launch(const char* kernel_name) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, kernel_name);
  return 0;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace hn

extern "C" {

int hn_abi_version(void) { return HN_ABI_VERSION; }

const char* hn_last_error_string(void) { return hn::g_last_error; }

int hn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return hn::fail((int)e, "cudaGetDevice");
  int v = 0;
  if (sm_count) {
    e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return hn::fail((int)e, "cudaDeviceGetAttribute");
    *sm_count = v;
  }
  if (cc_major) {
    e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return hn::fail((int)e, "cudaDeviceGetAttribute");
    *cc_major = v;
  }
  if (cc_minor) {
    e = cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) return hn::fail((int)e, "cudaDeviceGetAttribute");
    *cc_minor = v;
  }
  return 0;
}

}  // extern "C"
