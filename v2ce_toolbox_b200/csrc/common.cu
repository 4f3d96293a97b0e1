#include "common.cuh"

#include <string.h>

namespace v2ce {

char* last_error_buffer() {
  static thread_local char buf[1024] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return dev & 63;
}

bool pdl_enabled() {
  static const bool on = getenv("V2CE_PDL") && atoi(getenv("V2CE_PDL")) != 0;   // measured: no gain (9.4 ms either way), off by default
  return on;
}

int sm_count_cached() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

}  // namespace v2ce

extern "C" const char* v2ce_last_error(void) { return v2ce::last_error_buffer(); }

extern "C" int v2ce_version(void) { return 100; }

extern "C" int v2ce_device_check(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  V2CE_CUDA_CHECK(cudaGetDeviceCount(&n));
  V2CE_REQUIRE(device >= 0 && device < n, "device %d out of range (%d visible)", device, n);
  int major = 0, minor = 0, sms = 0;
  V2CE_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  V2CE_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  V2CE_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = major;
  if (cc_minor) *cc_minor = minor;
  V2CE_REQUIRE(major == 10, "libv2ce_b200 is built for sm_100a only; device %d is sm_%d%d", device, major, minor);
  return V2CE_OK;
}
