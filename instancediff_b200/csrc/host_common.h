// Host-side helpers shared by the .cu files (error text, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/idiff.h"

namespace idiff {
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
int watchdog_conv(int clear);   // conv_gemm.cu
int watchdog_attn(int clear);   // attention.cu
int watchdog_stem(int clear);   // stem_tc.cu
int watchdog_lattn(int clear);  // linattn_fused.cu
int watchdog_rowpair(int clear);  // conv3_rowpair.cu
// One-time PER-DEVICE launch setup: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to the
// current device, so a process that drives several GPUs (ConditionalUNet.to(other), two nets on two devices) must
// repeat it on each.  Thread-safe: a racing second thread at worst repeats the idempotent setup.
struct DeviceOnce {
  std::atomic<unsigned long long> done{0};
  int sms[64] = {};
};
template <typename F>
inline cudaError_t per_device_setup(DeviceOnce& d, int* sms_out, F setup) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const int slot = dev & 63;
  if (!((d.done.load(std::memory_order_acquire) >> slot) & 1ull)) {
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = setup();
    if (e != cudaSuccess) return e;
    d.sms[slot] = sms;
    d.done.fetch_or(1ull << slot, std::memory_order_release);
  }
  if (sms_out) *sms_out = d.sms[slot];
  return cudaSuccess;
}
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace idiff

#define IDIFF_REQUIRE(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return idiff::fail(IDIFF_ERR_ARG, __VA_ARGS__);   \
  } while (0)
