// read_from_pixel, second half: merge the 64-pixel tiles of one attention row (query i, head h) of object bk -- tile-local
// maxima m_t, sums l_t and unnormalised Z_t [256] written by qt_p2q_tc_kernel (csrc/qt_tc.cu) -- normalise, and apply the
// per-head value projection (transformer_layers.py:88-93, value half of in_proj).  One body for the stand-alone kernel
// (qt_tc.cu) and for the fused query chain (qt_chain_kernel, qt.cu).  256 threads; coef [tiles] and zn [E] in shared memory.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace cutie {

template <int E, int H, int Q>
__device__ __forceinline__ void qt_p2q_combine_tile(const float* __restrict__ ws, int tiles, const float* __restrict__ wv,
                                                    long long ldwv, const float* __restrict__ bv,
                                                    float* __restrict__ attn, const int i, const int h, const long long bk,
                                                    float* coef, float* zn) {
  constexpr int ROWS_ = Q * H;
  constexpr int WS_ = ROWS_ * (E + 2);              // per (object, tile): Z [128][256], m [128], l [128]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = i * H + h;
  const float* base = ws + bk * tiles * (long long)WS_;
  if (warp == 0) {
    float M = -CUDART_INF_F;
    for (int t = lane; t < tiles; t += 32) M = fmaxf(M, base[(long long)t * WS_ + ROWS_ * E + r]);
    M = warp_max(M);
    float L = 0.f;
    for (int t = lane; t < tiles; t += 32) {
      const float mt = base[(long long)t * WS_ + ROWS_ * E + r];
      const float f = (mt == -CUDART_INF_F) ? 0.f : expf(mt - M);
      coef[t] = f;
      L += f * base[(long long)t * WS_ + ROWS_ * E + ROWS_ + r];
    }
    L = warp_sum(L);
    __syncwarp();
    const float inv = 1.f / L;
    for (int t = lane; t < tiles; t += 32) coef[t] *= inv;
  }
  __syncthreads();
  float acc = 0.f;
  for (int t = 0; t < tiles; ++t) {
    const float c = coef[t];
    if (c != 0.f) acc = fmaf(c, base[(long long)t * WS_ + (long long)r * E + tid], acc);   // skip fully masked tiles
  }
  zn[tid] = acc;
  __syncthreads();
  for (int e = warp; e < 32; e += 8) {
    const float* wr = wv + (long long)(h * 32 + e) * ldwv;
    float d = 0.f;
#pragma unroll
    for (int c = lane; c < E; c += 32) d = fmaf(zn[c], wr[c], d);
    d = warp_sum(d);
    if (lane == 0) attn[(bk * Q + i) * E + h * 32 + e] = d + bv[h * 32 + e];
  }
}

}  // namespace cutie
