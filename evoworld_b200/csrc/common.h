// Shared helpers for the evoworld_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/evoworld_b200.h"

namespace evw {

void set_error(const char* fmt, ...);

#define EVW_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      evw::set_error(__VA_ARGS__);               \
      return EVW_ERR_INVALID;                    \
    }                                            \
  } while (0)

#define EVW_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      evw::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return EVW_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define EVW_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      evw::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return EVW_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

int sm_count();

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace evw
