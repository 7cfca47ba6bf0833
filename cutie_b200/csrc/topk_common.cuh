// Shared device helpers of the affinity kernels (exact scan, tcgen05 filter, re-rank).
#pragma once
#include <limits.h>
#include <math_constants.h>

#include "common.cuh"

namespace cutie {

constexpr int CKD = 64;        // key channels
constexpr int KPAD_MAX = 64;   // largest padded top-k list

// Insert candidate (s, idx) into a descending (value, then ascending index) list of `top_k` live slots
// stored at lv/li[0..32*NS).  Executed by one full warp.  Returns the list's k-th value afterwards.
template <int NS>
__device__ __forceinline__ float list_insert(float* lv, int* li, int lane, int top_k, float s, int idx) {
  const unsigned full = 0xffffffffu;
  float v[NS];
  int ix[NS];
  int pos = 0;
#pragma unroll
  for (int u = 0; u < NS; ++u) {
    v[u] = lv[lane + 32 * u];
    ix[u] = li[lane + 32 * u];
    bool better = (v[u] > s) || (v[u] == s && ix[u] < idx);
    pos += __popc(__ballot_sync(full, better));
  }
  float tau = 0.f;
  if (pos < top_k) {
    float nv[NS];
    int ni[NS];
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      int slot = lane + 32 * u;
      float pv = __shfl_up_sync(full, v[u], 1);
      int pi = __shfl_up_sync(full, ix[u], 1);
      if (u > 0) {
        float cv = __shfl_sync(full, v[u > 0 ? u - 1 : 0], 31);
        int ci = __shfl_sync(full, ix[u > 0 ? u - 1 : 0], 31);
        if (lane == 0) { pv = cv; pi = ci; }
      }
      nv[u] = slot < pos ? v[u] : (slot == pos ? s : pv);
      ni[u] = slot < pos ? ix[u] : (slot == pos ? idx : pi);
      if (slot >= top_k) { nv[u] = -CUDART_INF_F; ni[u] = INT_MAX; }
      lv[slot] = nv[u];
      li[slot] = ni[u];
    }
#pragma unroll
    for (int u = 0; u < NS; ++u)
      if (u == ((top_k - 1) >> 5)) tau = __shfl_sync(full, nv[u], (top_k - 1) & 31);
  } else {
#pragma unroll
    for (int u = 0; u < NS; ++u)
      if (u == ((top_k - 1) >> 5)) tau = __shfl_sync(full, v[u], (top_k - 1) & 31);
  }
  __syncwarp();
  return tau;
}


// Exact fp32 direct-form similarity of one memory token against one query, bit-identical everywhere it is
// used: S = (sum_c (a_c k_c - b_c)^2) * (-shr / sqrt(CK)), channels accumulated in order, a = sqrt(qe), b = a qk.
__device__ __forceinline__ float exact_similarity(const float* __restrict__ krow, float shr,
                                                  const float* __restrict__ a, const float* __restrict__ b) {
  float acc = 0.f;
#pragma unroll
  for (int c4 = 0; c4 < CKD / 4; ++c4) {
    const float4 kf = __ldg(reinterpret_cast<const float4*>(krow) + c4);
    float d;
    d = fmaf(a[4 * c4 + 0], kf.x, -b[4 * c4 + 0]); acc = fmaf(d, d, acc);
    d = fmaf(a[4 * c4 + 1], kf.y, -b[4 * c4 + 1]); acc = fmaf(d, d, acc);
    d = fmaf(a[4 * c4 + 2], kf.z, -b[4 * c4 + 2]); acc = fmaf(d, d, acc);
    d = fmaf(a[4 * c4 + 3], kf.w, -b[4 * c4 + 3]); acc = fmaf(d, d, acc);
  }
  return acc * (-shr * rsqrtf((float)CKD));
}

// Final step shared by the merge and re-rank kernels: softmax over the sorted winners held in lv/li
// (max-subtracted; equals exp(S)/sum exp(S) of memory_utils.py:60-61 whenever that is finite), padded
// outputs, optional fixed-point usage accumulation.  One warp.
template <int NS>
__device__ __forceinline__ void finalize_topk(const float* lv, const int* li, int lane, int top_k, int kp,
                                              int* out_idx, float* out_w, float* out_sim,
                                              unsigned long long* usage_row) {
  const float smax = lv[0];
  float e[NS], sum = 0.f;
#pragma unroll
  for (int u = 0; u < NS; ++u) {
    const int slot = lane + 32 * u;
    e[u] = (slot < top_k && li[slot] != INT_MAX) ? expf(lv[slot] - smax) : 0.f;
    sum += e[u];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int u = 0; u < NS; ++u) {
    const int slot = lane + 32 * u;
    if (slot < kp) {
      const bool live = slot < top_k && li[slot] != INT_MAX;
      const float w = live ? e[u] * inv : 0.f;
      const int id = live ? li[slot] : -1;
      out_idx[slot] = id;
      out_w[slot] = w;
      if (out_sim) out_sim[slot] = live ? lv[slot] : 0.f;
      if (usage_row && live)
        atomicAdd(&usage_row[id], (unsigned long long)((double)w * (double)(1ull << CUTIE_B200_USAGE_FRAC_BITS)));
    }
  }
}

}  // namespace cutie
