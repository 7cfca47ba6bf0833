// Test-only: run a barrier-free, shared-memory-free CUDA kernel body on the host, one (block, thread) at a time.
// The kernel source is #included verbatim after this header (tests/test_kernel_index_math_cpu.py), so the index
// arithmetic that runs on the GPU is the arithmetic tested here.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 v = {a, b, c, d}; return v; }
static dim3 blockIdx, threadIdx, gridDim, blockDim;
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
template <typename T> static inline T __ldg(const T* p) { return *p; }
#define CUDART_INF_F (__builtin_huge_valf())
// round-to-nearest single operations (the host compiler must not contract them either: built with -ffp-contract=off)
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }

// serial launch: EMU_LAUNCH(kernel<...>, grid, block, args...)
#define EMU_LAUNCH(kern, grid, block, ...)                                   \
  do {                                                                       \
    gridDim = dim3(grid); blockDim = dim3(block);                            \
    for (unsigned bx = 0; bx < gridDim.x; ++bx)                              \
      for (unsigned tx = 0; tx < blockDim.x; ++tx) {                         \
        blockIdx = dim3(bx); threadIdx = dim3(tx);                           \
        kern(__VA_ARGS__);                                                   \
      }                                                                      \
  } while (0)
