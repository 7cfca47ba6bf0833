// Internal (non-ABI) interfaces between affinity.cu (exact scan, orchestration) and affinity_tc.cu
// (tcgen05 candidate filter, exact re-rank).
#pragma once
#include "common.cuh"

namespace cutie {

// A "sample" of the bank: virtual indices i in [0, samp_count) map to tokens g = samp_begin + i * samp_stride
// of the concatenated segments.  stride 1 = the whole bank.

struct TcFilterParams {
  KeySegments segs;
  const float* qk;
  const float* qe;
  long long Q;
  long long samp_begin, samp_stride, samp_count;
  int tiles_per_split, nsplit;
  const float* tau;        // per query: k-th best exact similarity of an earlier (nested) sample; may be null
  long long tau_stride;    // tau[(b*Q + q) * tau_stride]
  int* cand;               // [B][nsplit][Q][cap] token indices
  int* count;              // [B][nsplit][Q]  (-1 = overflow)
  int cap;
  float* dbg_energy;       // optional [B][Q][samp_count] tf32 energies (tests)
};

struct RerankParams {
  KeySegments segs;
  const float* qk;
  const float* qe;
  long long Q, n_total;
  long long samp_begin, samp_stride, samp_count;
  int tiles_per_split, nsplit;
  const int* cand;
  const int* count;
  int cap, top_k, kpad;
  int* out_idx;
  float* out_w;
  float* out_sim;
  unsigned long long* usage_acc;
};

size_t tc_filter_smem_bytes();
int tc_split_count(long long B, long long Q, long long samp_count);
int launch_tc_filter(const TcFilterParams& p, long long B, cudaStream_t st);
int launch_rerank(const RerankParams& p, long long B, cudaStream_t st);

}  // namespace cutie
