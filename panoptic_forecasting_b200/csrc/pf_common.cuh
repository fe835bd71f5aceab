// Shared helpers for libpf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/pf_b200.h"

namespace pf {

void set_error(const char* fmt, ...);

#define PF_CHECK_CUDA(expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      pf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                              \
    }                                                                              \
  } while (0)

#define PF_REQUIRE(cond, code, ...)        \
  do {                                     \
    if (!(cond)) {                         \
      pf::set_error(__VA_ARGS__);          \
      return (code);                       \
    }                                      \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

}  // namespace pf
