// Object-transformer cross attention on the 5th-gen tensor cores (sm_100a, tcgen05 + TMEM).
//
// Both cross attentions of a QueryTransformerBlock (transformer_layers.py:45-98 via object_transformer.py:35-73) are
// algebraically folded (model/object_transformer.py) so that per object they are two dense contractions around a
// softmax, with M = 128 = 16 object queries x 8 heads:
//
//   read_from_pixel  (qt_p2q_tc_kernel, one CTA per 64-pixel tile and object)
//       S[r, p]  = Qf[r, :] . (X + PE)[:, p]          r = i*8 + h       GEMM 1: M=128 (r), N=64 (p), K=256 (channels)
//       P        = masked exp(S - rowmax)             (foreground / background rule of _get_aux_mask, :179-205)
//       Z[r, c]  = sum_p P[r, p] X[c, p]                                GEMM 2: M=128 (r), N=256 (c), K=64 (p)
//     (tile-local max / sum / Z go to a workspace; qt_p2q_combine_kernel merges tiles and applies the value projection)
//
//   read_from_query  (qt_q2p_tc_kernel, one CTA per 128-pixel tile and object)
//       S[p, r]  = (X + PE)[:, p] . Kf[r, :] + kd[r]  r = j*8 + h       GEMM 1: M=128 (p), N=128 (r), K=256
//       P        = softmax over the 16 queries j inside each head h
//       O[p, c]  = X[c, p] + bo[c] + sum_r P[p, r] Vf[r, c]             GEMM 2: M=128 (p), N=256 (c), K=128 (r)
//
// Pixel tensors are channel-major [BK, 256, HW] (what the cuDNN convolutions around these kernels emit), i.e. the pixel
// axis is contiguous: as an MMA operand that is MN-major when pixels are the M/N index (GEMM 1 of both kernels) and
// K-major when pixels are contracted (GEMM 2 of read_from_pixel); Vf [r][c] is MN-major for GEMM 2 of read_from_query.
// The MN-major SWIZZLE_128B descriptors were checked on hardware with tests/cuda/umma_probe.cu.
//
// Precision: kind::tf32 with both operands split x = hi + lo (hi = tf32(x) RN, lo = tf32(x - hi)) and three MMAs per
// k-step (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM): relative error ~2^-21 per product, i.e. fp32-class --
// plain 1xTF32 (2^-11) is not enough for the 1e-3 bar on segmentation logits after the decoder's amplification.
//
// Warp roles (416 threads): warps 0-3 softmax / epilogue (thread == TMEM lane), warps 4-11 producers (global fp32 ->
// hi/lo split -> swizzled shared memory, next chunk's loads in flight while the current one is converted), warp 12
// TMEM allocator + single-thread MMA issuer.  mbarrier pipelines between the roles; no __syncthreads in the main loop.
#include <math_constants.h>

#include "common.cuh"
#include "qt_combine.cuh"
#include "tc_ptx.cuh"

namespace cutie {

namespace {

constexpr int E_ = 256, H_ = 8, NQ = 16;
constexpr int ROWS = NQ * H_;            // 128 folded (query, head) rows
constexpr int QT_THREADS = 416;
constexpr int N_PROD = 256;              // producer threads (warps 4-11)
constexpr int KC = 32;                   // channels per GEMM-1 chunk (one 128-byte swizzle row of K)

__device__ __forceinline__ void split_tf32(float4 v, float4& hi, float4& lo) {
  hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  lo = make_float4(to_tf32(v.x - hi.x), to_tf32(v.y - hi.y), to_tf32(v.z - hi.z), to_tf32(v.w - hi.w));
}
__device__ __forceinline__ void store_split(unsigned char* hi_base, unsigned char* lo_base, int off, float4 v) {
  float4 hi, lo;
  split_tf32(v, hi, lo);
  *reinterpret_cast<float4*>(hi_base + off) = hi;
  *reinterpret_cast<float4*>(lo_base + off) = lo;
}
// 4 consecutive pixels of one channel row; VEC: rows are 16-byte aligned (HW % 4 == 0), else element-wise
template <bool VEC>
__device__ __forceinline__ float4 load_px4(const float* row, long long p, long long HW) {
  if (VEC) {
    return p < HW ? __ldg(reinterpret_cast<const float4*>(row + p)) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    float4 v;
    v.x = p < HW ? __ldg(row + p) : 0.f;
    v.y = p + 1 < HW ? __ldg(row + p + 1) : 0.f;
    v.z = p + 2 < HW ? __ldg(row + p + 2) : 0.f;
    v.w = p + 3 < HW ? __ldg(row + p + 3) : 0.f;
    return v;
  }
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// three-product tf32 MMA: D (+)= A.B with A = a_hi + a_lo, B = b_hi + b_lo (lo*lo dropped)
__device__ __forceinline__ void mma3(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                     bool accumulate) {
  tc_mma_tf32(d, a_lo, b_hi, idesc, accumulate ? 1u : 0u);
  tc_mma_tf32(d, a_hi, b_lo, idesc, 1u);
  tc_mma_tf32(d, a_hi, b_hi, idesc, 1u);
}

}  // namespace

// =====================================================================================================
// read_from_pixel
// =====================================================================================================
constexpr int P2Q_TP = 64;                         // pixels per CTA
constexpr int P2Q_NST1 = 3;                        // GEMM-1 stages
constexpr int P2Q_ST1_BYTES = 2 * 16384 + 2 * 8192;   // A_hi | A_lo [128 x 128 B]  +  B_hi | B_lo [2 groups x 32 rows x 128 B]
constexpr int P2Q_P_BYTES = 2 * 128 * 128;         // P_hi (or P_lo): 2 K-blocks of 32 pixels x 128 rows x 128 B
constexpr int P2Q_NC2 = 64;                        // channels per GEMM-2 chunk
constexpr int P2Q_ST2_BYTES = 2 * 16384;           // X_hi | X_lo: 2 K-blocks x 64 rows x 128 B each
constexpr int P2Q_REGION = P2Q_NST1 * P2Q_ST1_BYTES;   // 147456; phase 2 (P 64 KB + 2 stages 64 KB) aliases it
constexpr int P2Q_WS = ROWS * (E_ + 2);            // per (object, tile): Z [128][256], m [128], l [128]

struct P2QTail {
  unsigned long long full1[P2Q_NST1], empty1[P2Q_NST1], full2[2], empty2[2], s_full, p_full, z_full;
  uint32_t tmem_base;
  int cnt;
  unsigned char fg[P2Q_TP];
};

struct P2QTcParams {
  const float* qfold;   // [BK*16, 8, 256]: row (bk, i, h) == folded row r = i*8 + h of object bk
  const float* pixel;   // [BK, 256, HW]
  const float* pe;      // [BK, 256, HW]
  const uint8_t* fg;    // [BK, HW]
  const int* fg_count;  // [BK]
  long long HW;
  int tiles;
  float* ws;            // [BK][tiles][P2Q_WS]
};

template <bool VEC>
__global__ void __launch_bounds__(QT_THREADS, 1) qt_p2q_tc_kernel(const P2QTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  P2QTail& T = *reinterpret_cast<P2QTail*>(smem + P2Q_REGION);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const long long bk = blockIdx.y;
  const long long p0 = (long long)tile * P2Q_TP;
  if (tid == 0) {
    for (int s = 0; s < P2Q_NST1; ++s) { mbar_init(smem_u32(&T.full1[s]), N_PROD); mbar_init(smem_u32(&T.empty1[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&T.full2[s]), N_PROD); mbar_init(smem_u32(&T.empty2[s]), 1); }
    mbar_init(smem_u32(&T.s_full), 1);
    mbar_init(smem_u32(&T.p_full), 128);
    mbar_init(smem_u32(&T.z_full), 1);
    mbar_init_fence();
    T.cnt = p.fg_count[bk];
  }
  if (tid < P2Q_TP) T.fg[tid] = (p0 + tid < p.HW) ? p.fg[bk * p.HW + p0 + tid] : 0;
  if (warp == 12) tmem_alloc<512>(smem_u32(&T.tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = T.tmem_base;
  const uint32_t S_COL = 0, Z_COL = 64;

  if (warp >= 4 && warp < 12) {
    // ================================== producers ==================================
    const int pt = tid - 128;
    const float* qf = p.qfold + bk * ROWS * E_;
    const float* xb = p.pixel + bk * E_ * p.HW;
    const float* eb = p.pe + bk * E_ * p.HW;
    float4 av[4], xv[2], ev[2];
    auto load1 = [&](int j) {
      const int c0 = j * KC;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = pt + N_PROD * u, row = f >> 3, c4 = f & 7;
        av[u] = __ldg(reinterpret_cast<const float4*>(qf + (long long)row * E_ + c0 + 4 * c4));
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int f = pt + N_PROD * u, ch = f >> 4, q4 = f & 15;
        const long long roff = (long long)(c0 + ch) * p.HW;
        xv[u] = load_px4<VEC>(xb + roff, p0 + 4 * q4, p.HW);
        ev[u] = load_px4<VEC>(eb + roff, p0 + 4 * q4, p.HW);
      }
    };
    load1(0);
    for (int j = 0; j < E_ / KC; ++j) {
      const int s = j % P2Q_NST1;
      mbar_wait(smem_u32(&T.empty1[s]), ((j / P2Q_NST1) & 1) ^ 1);
      unsigned char* st = smem + s * P2Q_ST1_BYTES;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = pt + N_PROD * u, row = f >> 3, c4 = f & 7;
        store_split(st, st + 16384, row * 128 + ((c4 ^ (row & 7)) << 4), av[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int f = pt + N_PROD * u, ch = f >> 4, q4 = f & 15, grp = q4 >> 3, c4 = q4 & 7;
        store_split(st + 32768, st + 40960, grp * 4096 + off_mn32(ch, c4), add4(xv[u], ev[u]));
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&T.full1[s]));
      if (j + 1 < E_ / KC) load1(j + 1);
    }
    // phase 2: X chunks of 64 channels as the K-major B operand (K = pixels); the first chunk's loads are issued
    // before GEMM 1 has drained
    float4 zv[4];
    auto load2 = [&](int c) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = pt + N_PROD * u, ch = f >> 4, q4 = f & 15;
        zv[u] = load_px4<VEC>(xb + (long long)(c * P2Q_NC2 + ch) * p.HW, p0 + 4 * q4, p.HW);
      }
    };
    load2(0);
    mbar_wait(smem_u32(&T.s_full), 0);         // every GEMM-1 read of the aliased region has completed
    for (int c = 0; c < E_ / P2Q_NC2; ++c) {
      const int s = c & 1;
      mbar_wait(smem_u32(&T.empty2[s]), ((c >> 1) & 1) ^ 1);
      unsigned char* st = smem + 2 * P2Q_P_BYTES + s * P2Q_ST2_BYTES;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = pt + N_PROD * u, ch = f >> 4, q4 = f & 15, blk = q4 >> 3, c4 = q4 & 7;
        store_split(st, st + 16384, blk * 8192 + ch * 128 + ((c4 ^ (ch & 7)) << 4), zv[u]);
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&T.full2[s]));
      if (c + 1 < E_ / P2Q_NC2) load2(c + 1);
    }
  } else if (warp == 12) {
    // ================================== MMA issuer ==================================
    if (lane == 0) {
      const uint32_t id1 = idesc_tf32(128, P2Q_TP, false, true);      // A = Qf (K-major), B = X+PE (MN-major)
      for (int j = 0; j < E_ / KC; ++j) {
        const int s = j % P2Q_NST1;
        mbar_wait(smem_u32(&T.full1[s]), (j / P2Q_NST1) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * P2Q_ST1_BYTES), a_lo = a_hi + 16384, b_hi = a_hi + 32768, b_lo = a_hi + 40960;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma3(tmem + S_COL, desc_sw128_kmajor(a_hi + ks * 32), desc_sw128_kmajor(a_lo + ks * 32),
               desc_sw128_mnmajor(b_hi + ks * 1024, 4096), desc_sw128_mnmajor(b_lo + ks * 1024, 4096), id1, (j | ks) != 0);
        tc_commit(smem_u32(&T.empty1[s]));
      }
      tc_commit(smem_u32(&T.s_full));
      const uint32_t id2 = idesc_tf32(128, P2Q_NC2, false, false);     // A = P (K-major), B = X (K-major, K = pixels)
      mbar_wait(smem_u32(&T.p_full), 0);
      const uint32_t p_hi = smem_u32(smem), p_lo = p_hi + P2Q_P_BYTES;
      for (int c = 0; c < E_ / P2Q_NC2; ++c) {
        const int s = c & 1;
        mbar_wait(smem_u32(&T.full2[s]), (c >> 1) & 1);
        tc_fence_after();
        const uint32_t x_hi = smem_u32(smem + 2 * P2Q_P_BYTES + s * P2Q_ST2_BYTES), x_lo = x_hi + 16384;
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma3(tmem + Z_COL + c * P2Q_NC2, desc_sw128_kmajor(p_hi + blk * 16384 + ks * 32),
                 desc_sw128_kmajor(p_lo + blk * 16384 + ks * 32), desc_sw128_kmajor(x_hi + blk * 8192 + ks * 32),
                 desc_sw128_kmajor(x_lo + blk * 8192 + ks * 32), id2, (blk | ks) != 0);
        tc_commit(smem_u32(&T.empty2[s]));
      }
      tc_commit(smem_u32(&T.z_full));
    }
  } else {
    // ================================== softmax / epilogue: thread == row r ==================================
    const int r = tid;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float* wsr = p.ws + (bk * p.tiles + tile) * (long long)P2Q_WS;
    const bool fgq = r < ROWS / 2;                  // rows of the foreground queries (i < 8)
    const int cnt = T.cnt;
    const bool open = fgq ? (cnt == 0) : (cnt == (int)p.HW);     // a fully blocked row is opened (:203)
    mbar_wait(smem_u32(&T.s_full), 0);
    tc_fence_after();
    uint32_t v0[32], v1[32];
    tmem_ld32(lane_base + S_COL, v0);
    tmem_ld32(lane_base + S_COL + 32, v1);
    float mx = -CUDART_INF_F;
    unsigned ok0 = 0u, ok1 = 0u;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const bool in0 = p0 + j < p.HW, in1 = p0 + 32 + j < p.HW;
      const bool f0 = T.fg[j] != 0, f1 = T.fg[32 + j] != 0;
      const bool a0 = in0 && (open || (fgq ? f0 : !f0)), a1 = in1 && (open || (fgq ? f1 : !f1));
      ok0 |= a0 ? (1u << j) : 0u;
      ok1 |= a1 ? (1u << j) : 0u;
      if (a0) mx = fmaxf(mx, __uint_as_float(v0[j]));
      if (a1) mx = fmaxf(mx, __uint_as_float(v1[j]));
    }
    float sum = 0.f;
    unsigned char* Ph = smem;
    unsigned char* Pl = smem + P2Q_P_BYTES;
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        float e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = 4 * c4 + i;
          const bool a = ((blk ? ok1 : ok0) >> j) & 1u;
          const float sv = __uint_as_float(blk ? v1[j] : v0[j]);
          e[i] = a ? expf(sv - mx) : 0.f;
          sum += e[i];
        }
        store_split(Ph, Pl, blk * 16384 + r * 128 + ((c4 ^ (r & 7)) << 4), make_float4(e[0], e[1], e[2], e[3]));
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(smem_u32(&T.p_full));
    wsr[ROWS * E_ + r] = mx;
    wsr[ROWS * E_ + ROWS + r] = sum;
    mbar_wait(smem_u32(&T.z_full), 0);
    tc_fence_after();
    float4* zrow = reinterpret_cast<float4*>(wsr + (long long)r * E_);
#pragma unroll 1
    for (int g = 0; g < E_ / 32; ++g) {
      uint32_t z[32];
      tmem_ld32(lane_base + Z_COL + g * 32, z);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        zrow[g * 8 + i] = make_float4(__uint_as_float(z[4 * i]), __uint_as_float(z[4 * i + 1]), __uint_as_float(z[4 * i + 2]),
                                      __uint_as_float(z[4 * i + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// merge the pixel tiles of one attention row (i, h), normalise, apply the per-head value projection:
// grid (16 query rows, heads, BK), 256 threads (thread == channel, then 8 warps x 4 outputs); body in qt_combine.cuh
// (shared with qt_chain_kernel, csrc/qt.cu)
__global__ void __launch_bounds__(256) qt_p2q_combine_kernel(const float* __restrict__ ws, int tiles,
                                                             const float* __restrict__ wv, long long ldwv,
                                                             const float* __restrict__ bv, float* __restrict__ attn) {
  extern __shared__ float coef[];   // [tiles]: exp(m_t - M) / L
  __shared__ float zn[E_];
  qt_p2q_combine_tile<E_, H_, NQ>(ws, tiles, wv, ldwv, bv, attn, (int)blockIdx.x, (int)blockIdx.y, (long long)blockIdx.z,
                                  coef, zn);
}

// =====================================================================================================
// read_from_query
// =====================================================================================================
constexpr int Q2P_TP = 128;                          // pixels per CTA (MMA M)
constexpr int Q2P_ST1_BYTES = 4 * 16384;             // A_hi | A_lo (X+PE, MN-major: 4 groups x 32 rows x 128 B) | B_hi | B_lo (Kf)
constexpr int Q2P_P_BYTES = 4 * 128 * 128;           // P_hi (or P_lo): 4 K-blocks of 32 r x 128 pixel rows x 128 B
constexpr int Q2P_NC2 = 32;                          // channels per GEMM-2 chunk
constexpr int Q2P_ST2_BYTES = 2 * 16384;             // V_hi | V_lo: 128 r rows x 128 B (32 channels)
constexpr int Q2P_REGION = 2 * Q2P_ST1_BYTES;        // 131072 = P_hi + P_lo (aliases the two GEMM-1 stages)

struct Q2PTail {
  unsigned long long full1[2], empty1[2], full2[2], empty2[2], s_full, p_full, z_full;
  uint32_t tmem_base;
  float kd[ROWS];
  float bo[E_];
};

struct Q2PTcParams {
  const float* kfold;   // [BK*16, 8, 256]  row r = j*8 + h
  const float* kdots;   // [BK*16, 8]
  const float* vfold;   // [BK*16, 8, 256]
  const float* out_bias;
  const float* pixel;
  const float* pe;
  long long HW;
  float* out;
};

template <bool VEC>
__global__ void __launch_bounds__(QT_THREADS, 1) qt_q2p_tc_kernel(const Q2PTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  Q2PTail& T = *reinterpret_cast<Q2PTail*>(smem + Q2P_REGION + 2 * Q2P_ST2_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long bk = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * Q2P_TP;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&T.full1[s]), N_PROD); mbar_init(smem_u32(&T.empty1[s]), 1);
      mbar_init(smem_u32(&T.full2[s]), N_PROD); mbar_init(smem_u32(&T.empty2[s]), 1);
    }
    mbar_init(smem_u32(&T.s_full), 1);
    mbar_init(smem_u32(&T.p_full), 128);
    mbar_init(smem_u32(&T.z_full), 1);
    mbar_init_fence();
  }
  if (tid < ROWS) T.kd[tid] = p.kdots[bk * ROWS + tid];
  if (tid >= 128 && tid < 128 + E_) T.bo[tid - 128] = p.out_bias[tid - 128];
  if (warp == 12) tmem_alloc<512>(smem_u32(&T.tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = T.tmem_base;
  const uint32_t S_COL = 0, O_COL = 128;
  const float* xb = p.pixel + bk * E_ * p.HW;

  if (warp >= 4 && warp < 12) {
    // ================================== producers ==================================
    const int pt = tid - 128;
    const float* eb = p.pe + bk * E_ * p.HW;
    const float* kf = p.kfold + bk * ROWS * E_;
    const float* vf = p.vfold + bk * ROWS * E_;
    float4 xv[4], ev[4], kv[4];
    auto load1 = [&](int j) {
      const int c0 = j * KC;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = pt + N_PROD * u;
        const int ch = f >> 5, q4 = f & 31;
        const long long roff = (long long)(c0 + ch) * p.HW;
        xv[u] = load_px4<VEC>(xb + roff, p0 + 4 * q4, p.HW);
        ev[u] = load_px4<VEC>(eb + roff, p0 + 4 * q4, p.HW);
        const int row = f >> 3, c4 = f & 7;
        kv[u] = __ldg(reinterpret_cast<const float4*>(kf + (long long)row * E_ + c0 + 4 * c4));
      }
    };
    load1(0);
    for (int j = 0; j < E_ / KC; ++j) {
      const int s = j & 1;
      mbar_wait(smem_u32(&T.empty1[s]), ((j >> 1) & 1) ^ 1);
      unsigned char* st = smem + s * Q2P_ST1_BYTES;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = pt + N_PROD * u;
        const int ch = f >> 5, q4 = f & 31, grp = q4 >> 3, c4 = q4 & 7;
        store_split(st, st + 16384, grp * 4096 + off_mn32(ch, c4), add4(xv[u], ev[u]));
        const int row = f >> 3, k4 = f & 7;
        store_split(st + 32768, st + 49152, row * 128 + ((k4 ^ (row & 7)) << 4), kv[u]);
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&T.full1[s]));
      if (j + 1 < E_ / KC) load1(j + 1);
    }
    // phase 2: Vf chunks of 32 channels (MN-major B: k = r rows, 32 channels = one 128-byte row); their buffers do not
    // alias GEMM 1, so the first chunks are converted while the softmax runs
    auto load2 = [&](int c) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = pt + N_PROD * u, row = f >> 3, c4 = f & 7;
        kv[u] = __ldg(reinterpret_cast<const float4*>(vf + (long long)row * E_ + c * Q2P_NC2 + 4 * c4));
      }
    };
    load2(0);
    for (int c = 0; c < E_ / Q2P_NC2; ++c) {
      const int s = c & 1;
      mbar_wait(smem_u32(&T.empty2[s]), ((c >> 1) & 1) ^ 1);
      unsigned char* st = smem + Q2P_REGION + s * Q2P_ST2_BYTES;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = pt + N_PROD * u, row = f >> 3, c4 = f & 7;
        store_split(st, st + 16384, off_mn32(row, c4), kv[u]);
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&T.full2[s]));
      if (c + 1 < E_ / Q2P_NC2) load2(c + 1);
    }
  } else if (warp == 12) {
    // ================================== MMA issuer ==================================
    if (lane == 0) {
      const uint32_t id1 = idesc_tf32(128, ROWS, true, false);        // A = X+PE (MN-major), B = Kf (K-major)
      for (int j = 0; j < E_ / KC; ++j) {
        const int s = j & 1;
        mbar_wait(smem_u32(&T.full1[s]), (j >> 1) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * Q2P_ST1_BYTES), a_lo = a_hi + 16384, b_hi = a_hi + 32768, b_lo = a_hi + 49152;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma3(tmem + S_COL, desc_sw128_mnmajor(a_hi + ks * 1024, 4096), desc_sw128_mnmajor(a_lo + ks * 1024, 4096),
               desc_sw128_kmajor(b_hi + ks * 32), desc_sw128_kmajor(b_lo + ks * 32), id1, (j | ks) != 0);
        tc_commit(smem_u32(&T.empty1[s]));
      }
      tc_commit(smem_u32(&T.s_full));
      const uint32_t id2 = idesc_tf32(128, Q2P_NC2, false, true);      // A = P (K-major), B = Vf (MN-major)
      mbar_wait(smem_u32(&T.p_full), 0);
      const uint32_t p_hi = smem_u32(smem), p_lo = p_hi + Q2P_P_BYTES;
      for (int c = 0; c < E_ / Q2P_NC2; ++c) {
        const int s = c & 1;
        mbar_wait(smem_u32(&T.full2[s]), (c >> 1) & 1);
        tc_fence_after();
        const uint32_t v_hi = smem_u32(smem + Q2P_REGION + s * Q2P_ST2_BYTES), v_lo = v_hi + 16384;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)
          mma3(tmem + O_COL + c * Q2P_NC2, desc_sw128_kmajor(p_hi + (ks >> 2) * 16384 + (ks & 3) * 32),
               desc_sw128_kmajor(p_lo + (ks >> 2) * 16384 + (ks & 3) * 32), desc_sw128_mnmajor(v_hi + ks * 1024, 4096),
               desc_sw128_mnmajor(v_lo + ks * 1024, 4096), id2, ks != 0);
        tc_commit(smem_u32(&T.empty2[s]));
      }
      tc_commit(smem_u32(&T.z_full));
    }
  } else {
    // ================================== softmax / epilogue: thread == pixel ==================================
    const int px = tid;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    mbar_wait(smem_u32(&T.s_full), 0);
    tc_fence_after();
    // column r = j*8 + h: three passes over the 128 columns (TMEM re-reads are cheap; 128 live registers are not)
    float mx[H_], inv[H_];
#pragma unroll
    for (int h = 0; h < H_; ++h) { mx[h] = -CUDART_INF_F; inv[h] = 0.f; }
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      uint32_t v[32];
      tmem_ld32(lane_base + S_COL + g * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) mx[j & 7] = fmaxf(mx[j & 7], __uint_as_float(v[j]) + T.kd[g * 32 + j]);
    }
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      uint32_t v[32];
      tmem_ld32(lane_base + S_COL + g * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) inv[j & 7] += expf(__uint_as_float(v[j]) + T.kd[g * 32 + j] - mx[j & 7]);
    }
#pragma unroll
    for (int h = 0; h < H_; ++h) inv[h] = 1.f / inv[h];
    unsigned char* Ph = smem;
    unsigned char* Pl = smem + Q2P_P_BYTES;
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
      uint32_t v[32];
      tmem_ld32(lane_base + S_COL + g * 32, v);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        float e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = 4 * c4 + i;
          e[i] = expf(__uint_as_float(v[j]) + T.kd[g * 32 + j] - mx[j & 7]) * inv[j & 7];
        }
        store_split(Ph, Pl, g * 16384 + px * 128 + ((c4 ^ (px & 7)) << 4), make_float4(e[0], e[1], e[2], e[3]));
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(smem_u32(&T.p_full));
    mbar_wait(smem_u32(&T.z_full), 0);
    tc_fence_after();
    const long long pg = p0 + px;
    float* ob = p.out + bk * E_ * p.HW;
#pragma unroll 1
    for (int g = 0; g < E_ / 32; ++g) {
      float res[32];
      if (pg < p.HW) {
#pragma unroll
        for (int j = 0; j < 32; ++j) res[j] = __ldg(xb + (long long)(g * 32 + j) * p.HW + pg);
      }
      uint32_t o[32];
      tmem_ld32(lane_base + O_COL + g * 32, o);
      if (pg < p.HW) {
#pragma unroll
        for (int j = 0; j < 32; ++j) ob[(long long)(g * 32 + j) * p.HW + pg] = res[j] + T.bo[g * 32 + j] + __uint_as_float(o[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace cutie

using namespace cutie;

extern "C" int cutie_qt_pixel_to_query_splits(int64_t BK, int64_t HW, int num_heads) {
  (void)BK; (void)num_heads;
  return (int)((HW + P2Q_TP - 1) / P2Q_TP);
}

extern "C" int64_t cutie_qt_pixel_to_query_workspace_floats(int64_t BK, int64_t HW) {
  return BK * ((HW + P2Q_TP - 1) / P2Q_TP) * (int64_t)P2Q_WS;
}

extern "C" int cutie_qt_pixel_to_query(const float* qfold, const float* pixel, const float* pixel_pe,
                                       const uint8_t* fg, const int32_t* fg_count, const float* wv, int64_t ldwv,
                                       const float* bv, int64_t BK, int64_t E, int64_t HW, int num_queries,
                                       int num_heads, int splits, float* workspace, float* attn_out, void* stream) {
  // attn_out == nullptr: tiles only -- the merge + value projection then runs as an op of cutie_qt_chain (qt.cu)
  CUTIE_REQUIRE(qfold && pixel && pixel_pe && fg && fg_count && workspace, "null argument");
  CUTIE_REQUIRE(attn_out == nullptr || (wv && bv), "the combine step needs wv and bv");
  CUTIE_REQUIRE(E == E_ && num_heads == H_ && num_queries == NQ, "embed_dim 256, 8 heads, 16 queries");
  CUTIE_REQUIRE(BK >= 1 && HW >= 1, "empty");
  const int tiles = (int)((HW + P2Q_TP - 1) / P2Q_TP);
  CUTIE_REQUIRE(splits == tiles, "splits must equal cutie_qt_pixel_to_query_splits() (one CTA per 64-pixel tile)");
  CUTIE_REQUIRE(tiles <= 4096, "at most 262144 pixels");
  P2QTcParams p;
  p.qfold = qfold; p.pixel = pixel; p.pe = pixel_pe; p.fg = fg; p.fg_count = fg_count; p.HW = HW; p.tiles = tiles;
  p.ws = workspace;
  const size_t smem = (size_t)P2Q_REGION + sizeof(P2QTail) + 64;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(pixel) | reinterpret_cast<uintptr_t>(pixel_pe)) % 16 == 0);
  static bool attr_done[64] = {};
  if (first_use_on_device(attr_done)) {
    cudaFuncSetAttribute(qt_p2q_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(qt_p2q_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (vec)
    qt_p2q_tc_kernel<true><<<dim3((unsigned)tiles, (unsigned)BK), QT_THREADS, smem, st>>>(p);
  else
    qt_p2q_tc_kernel<false><<<dim3((unsigned)tiles, (unsigned)BK), QT_THREADS, smem, st>>>(p);
  CUTIE_CHECK_LAUNCH();
  if (attn_out == nullptr) return 0;
  qt_p2q_combine_kernel<<<dim3(NQ, H_, (unsigned)BK), 256, (size_t)tiles * sizeof(float), st>>>(workspace, tiles, wv, ldwv, bv,
                                                                                                 attn_out);
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_qt_query_to_pixel(const float* kfold, const float* kdots, const float* vfold,
                                       const float* out_bias, const float* pixel, const float* pixel_pe, int64_t BK,
                                       int64_t E, int64_t HW, int num_queries, int num_heads, float* out,
                                       void* stream) {
  CUTIE_REQUIRE(kfold && kdots && vfold && out_bias && pixel && pixel_pe && out, "null argument");
  CUTIE_REQUIRE(E == E_ && num_heads == H_ && num_queries == NQ, "embed_dim 256, 8 heads, 16 queries");
  CUTIE_REQUIRE(BK >= 1 && HW >= 1, "empty");
  Q2PTcParams p;
  p.kfold = kfold; p.kdots = kdots; p.vfold = vfold; p.out_bias = out_bias; p.pixel = pixel; p.pe = pixel_pe;
  p.HW = HW; p.out = out;
  const size_t smem = (size_t)Q2P_REGION + 2 * Q2P_ST2_BYTES + sizeof(Q2PTail) + 64;
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(pixel) | reinterpret_cast<uintptr_t>(pixel_pe)) % 16 == 0);
  static bool attr_done[64] = {};
  if (first_use_on_device(attr_done)) {
    cudaFuncSetAttribute(qt_q2p_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(qt_q2p_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  dim3 grid((unsigned)((HW + Q2P_TP - 1) / Q2P_TP), (unsigned)BK);
  if (vec)
    qt_q2p_tc_kernel<true><<<grid, QT_THREADS, smem, (cudaStream_t)stream>>>(p);
  else
    qt_q2p_tc_kernel<false><<<grid, QT_THREADS, smem, (cudaStream_t)stream>>>(p);
  CUTIE_CHECK_LAUNCH();
  return 0;
}
