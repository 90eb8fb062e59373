// Library-level entry points: version, last-error string, device query.
#include "common.cuh"
#include <string.h>

static thread_local char g_last_error[512] = {0};
static thread_local int g_pdl_mode = 0;
namespace mb { int& pdl_mode() { return g_pdl_mode; } }

extern "C" {

void mb200_set_last_error(const char* msg) {
  if (!msg) { g_last_error[0] = 0; return; }
  strncpy(g_last_error, msg, sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}
const char* mb200_last_error(void) { return g_last_error; }
int mb200_version(void) { return 100; }   // 0.1.0

// 0 if the current device is an sm_100 part (the only target of this library), -ENODEV otherwise.
int mb200_check_device(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { mb200_set_last_error("no CUDA device"); return -ENODEV; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) { mb200_set_last_error("mantis_b200 requires an sm_100a (B200) device"); return -ENODEV; }
  return 0;
}
int mb200_num_sms(void) { return mb::num_sms(); }

}  // extern "C"
