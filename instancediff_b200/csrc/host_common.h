// Host-side helpers shared by the .cu files (error text, launch checks).
#pragma once
#include <cuda_runtime.h>
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
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
}  // namespace idiff

#define IDIFF_REQUIRE(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return idiff::fail(IDIFF_ERR_ARG, __VA_ARGS__);   \
  } while (0)
