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

// Block-shared memory and one barrier: `__shared__` arrays become statics that persist between the serial thread calls,
// and a kernel whose threads meet at ONE __syncthreads() is run in two passes per block -- pass 0 stops every thread at
// the barrier (all shared-memory writes before it have then happened), pass 1 re-runs every thread from the top (the part
// before the barrier must be idempotent: it recomputes and rewrites the same shared values) and continues past it.
// Barrier-free kernels never set `emu_barrier_hit` and run once (in-place kernels must not run twice).
static int emu_pass = 0;
static bool emu_barrier_hit = false;
#define __shared__ static
#define __syncthreads() do { if (emu_pass == 0) { emu_barrier_hit = true; return; } } while (0)

// serial launch: EMU_LAUNCH(kernel<...>, grid, block, args...); grid may be 1-, 2- or 3-dimensional, blocks 1-dimensional
#define EMU_LAUNCH(kern, grid, block, ...)                                   \
  do {                                                                       \
    gridDim = dim3(grid); blockDim = dim3(block);                            \
    for (unsigned bz = 0; bz < gridDim.z; ++bz)                              \
      for (unsigned by = 0; by < gridDim.y; ++by)                            \
        for (unsigned bx = 0; bx < gridDim.x; ++bx) {                        \
          emu_barrier_hit = false;                                           \
          for (emu_pass = 0; emu_pass < (emu_barrier_hit ? 2 : 1); ++emu_pass) \
            for (unsigned tx = 0; tx < blockDim.x; ++tx) {                   \
              blockIdx = dim3(bx, by, bz); threadIdx = dim3(tx, 0, 0);       \
              kern(__VA_ARGS__);                                             \
            }                                                                \
          emu_pass = 0;                                                      \
        }                                                                    \
  } while (0)
