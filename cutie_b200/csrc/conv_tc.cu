// 3x3 / stride 1 / zero-pad 1 convolution as a tcgen05 implicit GEMM with fp32-class accuracy (3xTF32), sm_100a.
//
// the PixelFFN / CAResBlock (transformer_layers.py:121-136, channel_attn.py:7-39) and
// PixelFeatureFuser (big_modules.py:192-235) convolutions.  With the numerics the parity tests validate -- fp32, cuDNN
// TF32 off -- cuDNN runs them on the FP32 pipe; they are the largest cost inside the replaced subsystem.
//
//   Y[n, co, y, x] = act( bias[co] + sum_{ci, dy, dx} W[co, ci, dy, dx] * pre(X)[n, ci, y + dy - 1, x + dx - 1]  (+ Z[n, co, y, x]) )
//
// GEMM view per CTA: D[co (M = 128), position (N <= 128)] += A_tap[co, ci] . B_tap[ci, position], K = 32 input channels
// per (chunk, tap) step, 9 taps x Cin / 32 chunks.
//
//   * A (weights) comes from a precomputed OPERAND IMAGE (cutie_conv3x3_weight_image, built once per layer): for every
//     (128-channel output tile, 32-channel input chunk, tap) the [128 x 32] tf32 hi and lo planes in K-major
//     SWIZZLE_128B order, 32 KB, fetched by ONE cp.async.bulk per step through a 3-stage mbarrier ring.
//   * B (activations): the CTA's spatial tile is TH x TW output pixels; its input window with the 1-pixel halo is laid
//     out in shared memory ONCE per 32-channel chunk as rows of a local zero-padded grid -- row r = ly * (TW + 2) + lx + 1
//     holds the 32 channels of input pixel (ty0 + ly - 1, tx0 + lx - 1) as 128 bytes, K-major SWIZZLE_128B, hi and lo
//     planes.  Output position j = ty * (TW + 2) + lx reads, for tap (dy, dx), row j + dy * (TW + 2) + dx: EVERY TAP IS
//     THE SAME TILE READ THROUGH A DESCRIPTOR WHOSE START ADDRESS IS SHIFTED BY WHOLE ROWS -- no im2col copies, no
//     per-tap producer work.  (The hardware applies the 128-byte swizzle to absolute shared-memory address bits, so a
//     start shifted by r rows reads rows r.. of a tile that was written with the address-based pattern, with the
//     descriptor's base-offset field left 0; tests/cuda/umma_probe.cu checks this on the device:
//     profiles/r02_umma_probe.txt.)  The two padding columns of every local row are computed and
//     discarded (TW / (TW + 2) efficiency); image borders are zero rows.
//   * 3xTF32: x = hi + lo with hi = tf32(x) RN, lo = tf32(x - hi); three MMAs per k-step (lo.hi + hi.lo + hi.hi), fp32
//     accumulation in TMEM: relative error ~2^-21 per product.  The tensor core adds each MMA's result to the accumulator
//     with TRUNCATION, a bias that grows with the number of accumulations (measured: 864 of them into one accumulator at
//     Cin = 256 left 1.6e-5 of max|y|, cuDNN's fp32 FMA chain 1.1e-5).  So the K loop is spread over FOUR accumulators
//     (4 x 128 TMEM columns): the small cross terms lo.hi + hi.lo (2^-11 of the result: their truncation is invisible) in
//     one, the hi.hi products of chunk c in accumulator c % 3; the epilogue adds the four in fp32 (round to nearest).
//     Same MMA count, ~5x smaller error.
//   * Epilogue: thread == output channel (TMEM lane); bias, optional residual, optional ReLU; ReLU on the INPUT (the
//     pre-activation blocks' conv(relu(x))) is applied by the producers for free.
//
// Warp roles (448 threads): warps 0-3 epilogue, warps 4-11 activation producers (global fp32 -> hi/lo -> swizzled smem,
// next chunk's loads in flight during the current chunk's MMAs), warp 12 MMA issuer (one thread), warp 13 weight loader
// (one thread).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cutie {

namespace {

constexpr int CV_M = 128;                         // output channels per CTA
constexpr int CV_KC = 32;                         // input channels per chunk
constexpr int CV_A_BYTES = 2 * CV_M * 128;        // hi | lo planes of one (chunk, tap) weight block: 32768
constexpr int CV_A_STAGES = 3;
constexpr int CV_ROWS = 248;                      // activation tile rows (31 x 8: planes stay 1024-byte aligned)
constexpr int CV_X_PLANE = CV_ROWS * 128;         // 31744
constexpr int CV_X_BYTES = 2 * CV_X_PLANE;        // hi | lo
constexpr int CV_THREADS = 448;
constexpr int CV_PROD = 256;

struct ConvTail {
  unsigned long long a_full[CV_A_STAGES], a_empty[CV_A_STAGES], x_full[2], x_empty[2], acc_full;
  uint32_t tmem_base;
};
constexpr int CV_SMEM = 2 * CV_X_BYTES + CV_A_STAGES * CV_A_BYTES + (int)sizeof(ConvTail) + 64;

struct ConvTcParams {
  const float* x;              // [NB, Cin, H, W]
  const unsigned char* wimg;   // [ceil(Cout / 128)][Cin / 32][9][32768]
  const float* bias;           // [Cout] or null
  const float* z;              // [NB, Cout, H, W] or null
  float* y;                    // [NB, Cout, H, W]
  int Cin, Cout, H, W;
  int TH, TW, tiles_x;         // spatial tile; tiles per image row
  int N;                       // MMA N = round16(TH * (TW + 2))
  int relu_in, relu_out;
};

__global__ void __launch_bounds__(CV_THREADS, 1) conv3x3_tc_kernel(const ConvTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* Xs = smem;                                   // 2 stages x (hi | lo)
  unsigned char* As = smem + 2 * CV_X_BYTES;                  // CV_A_STAGES x (hi | lo)
  ConvTail& T = *reinterpret_cast<ConvTail*>(smem + 2 * CV_X_BYTES + CV_A_STAGES * CV_A_BYTES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, cot = blockIdx.y, nb = blockIdx.z;
  const int ty0 = (tile / p.tiles_x) * p.TH, tx0 = (tile % p.tiles_x) * p.TW;
  const int TWp = p.TW + 2;
  const int chunks = p.Cin / CV_KC;
  const long long HW = (long long)p.H * p.W;
  if (tid == 0) {
    for (int s = 0; s < CV_A_STAGES; ++s) { mbar_init(smem_u32(&T.a_full[s]), 1); mbar_init(smem_u32(&T.a_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&T.x_full[s]), CV_PROD); mbar_init(smem_u32(&T.x_empty[s]), 1); }
    mbar_init(smem_u32(&T.acc_full), 1);
    mbar_init_fence();
  }
  if (warp == 12) tmem_alloc<512>(smem_u32(&T.tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = T.tmem_base;

  if (warp >= 4 && warp < 12) {
    // ================================== activation producers: thread == tile row ==================================
    const int r = tid - 128;                                  // 0 .. 255; rows >= CV_ROWS do not exist
    const int rows_used = p.N + 2 * TWp + 2;
    const bool row_live = r < CV_ROWS;
    bool valid = false;
    long long poff = 0;
    if (r >= 1 && r < rows_used) {
      const int q = r - 1, ly = q / TWp, lx = q - ly * TWp;
      const int gy = ty0 + ly - 1, gx = tx0 + lx - 1;
      valid = ly < p.TH + 2 && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
      poff = (long long)gy * p.W + gx;
    }
    const float* xb = p.x + (long long)nb * p.Cin * HW + poff;
    float v[CV_KC];
    auto load = [&](int c) {
#pragma unroll
      for (int i = 0; i < CV_KC; ++i) v[i] = valid ? __ldg(xb + (long long)(c * CV_KC + i) * HW) : 0.f;
    };
    load(0);
    for (int c = 0; c < chunks; ++c) {
      const int s = c & 1;
      mbar_wait(smem_u32(&T.x_empty[s]), ((c >> 1) & 1) ^ 1);
      if (row_live) {
        unsigned char* hi = Xs + s * CV_X_BYTES + r * 128;
        unsigned char* lo = hi + CV_X_PLANE;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          float4 f = make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
          if (p.relu_in) f = make_float4(fmaxf(f.x, 0.f), fmaxf(f.y, 0.f), fmaxf(f.z, 0.f), fmaxf(f.w, 0.f));
          const float4 h = make_float4(to_tf32(f.x), to_tf32(f.y), to_tf32(f.z), to_tf32(f.w));
          const float4 l = make_float4(to_tf32(f.x - h.x), to_tf32(f.y - h.y), to_tf32(f.z - h.z), to_tf32(f.w - h.w));
          const int off = (k4 ^ (r & 7)) << 4;
          *reinterpret_cast<float4*>(hi + off) = h;
          *reinterpret_cast<float4*>(lo + off) = l;
        }
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&T.x_full[s]));
      if (c + 1 < chunks) load(c + 1);
    }
  } else if (warp == 13) {
    // ================================== weight loader ==================================
    if (lane == 0) {
      const unsigned char* wsrc = p.wimg + (size_t)cot * chunks * 9 * CV_A_BYTES;
      const int steps = chunks * 9;
      for (int i = 0; i < steps; ++i) {
        const int s = i % CV_A_STAGES;
        mbar_wait(smem_u32(&T.a_empty[s]), ((i / CV_A_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(smem_u32(&T.a_full[s]), CV_A_BYTES);
        bulk_g2s(smem_u32(As + s * CV_A_BYTES), wsrc + (size_t)i * CV_A_BYTES, CV_A_BYTES, smem_u32(&T.a_full[s]));
      }
    }
  } else if (warp == 12) {
    // ================================== MMA issuer ==================================
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(CV_M, p.N, false, false);
      int i = 0;
      for (int c = 0; c < chunks; ++c) {
        const int xs = c & 1;
        mbar_wait(smem_u32(&T.x_full[xs]), (c >> 1) & 1);
        tc_fence_after();
        const uint32_t xb_hi = smem_u32(Xs + xs * CV_X_BYTES), xb_lo = xb_hi + CV_X_PLANE;
        for (int t = 0; t < 9; ++t, ++i) {
          const int s = i % CV_A_STAGES;
          mbar_wait(smem_u32(&T.a_full[s]), (i / CV_A_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(As + s * CV_A_BYTES), a_lo = a_hi + CV_M * 128;
          const uint32_t shift = (uint32_t)((t / 3) * TWp + (t % 3)) * 128u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da_hi = desc_sw128_kmajor(a_hi + ks * 32), da_lo = desc_sw128_kmajor(a_lo + ks * 32);
            const uint64_t db_hi = desc_sw128_kmajor(xb_hi + shift + ks * 32);
            const uint64_t db_lo = desc_sw128_kmajor(xb_lo + shift + ks * 32);
            tc_mma_tf32(tmem + 384, da_lo, db_hi, idesc, (i | ks) != 0 ? 1u : 0u);      // cross terms
            tc_mma_tf32(tmem + 384, da_hi, db_lo, idesc, 1u);
            tc_mma_tf32(tmem + (uint32_t)(c % 3) * 128u, da_hi, db_hi, idesc, (c >= 3 || (t | ks) != 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(&T.a_empty[s]));
        }
        tc_commit(smem_u32(&T.x_empty[xs]));
      }
      tc_commit(smem_u32(&T.acc_full));
    }
  } else {
    // ================================== epilogue: thread == output channel ==================================
    const int co = cot * CV_M + tid;
    const bool co_ok = co < p.Cout;                           // the last channel tile may be padded (zero weight rows)
    const float b = (p.bias && co_ok) ? __ldg(p.bias + co) : 0.f;
    const long long obase = ((long long)nb * p.Cout + co) * HW;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    mbar_wait(smem_u32(&T.acc_full), 0);
    tc_fence_after();
    int ty = 0, lx = 0;                                     // position j = ty * TWp + lx
    const int nacc = chunks < 3 ? chunks : 3;               // hi.hi accumulators in use
    for (int g = 0; g < p.N; g += 32) {
      uint32_t o[32], q[32];
      tmem_ld32(lane_base + g, o);                          // (columns >= N of the last group are never stored)
      for (int a = 1; a < nacc; ++a) {
        tmem_ld32(lane_base + a * 128 + g, q);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) + __uint_as_float(q[j]));
      }
      tmem_ld32(lane_base + 384 + g, q);
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) + __uint_as_float(q[j]));
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int gy = ty0 + ty, gx = tx0 + lx - 1;
        if (co_ok && g + j < p.N && lx >= 1 && lx <= p.TW && ty < p.TH && gy < p.H && gx < p.W) {
          const long long a = obase + (long long)gy * p.W + gx;
          float val = __uint_as_float(o[j]) + b;
          if (p.z) val += __ldg(p.z + a);
          if (p.relu_out) val = fmaxf(val, 0.f);
          p.y[a] = val;
        }
        if (++lx == TWp) { lx = 0; ++ty; }
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

// weight operand image: one thread per (output channel, chunk, tap, 16-byte piece)
__global__ void __launch_bounds__(256) conv3x3_weight_image_kernel(const float* __restrict__ w, int Cout, int Cin,
                                                                   unsigned char* __restrict__ img) {
  const int chunks = Cin / CV_KC;
  const long long total = (long long)((Cout + CV_M - 1) / CV_M * CV_M) * chunks * 9 * 8;
  const long long f = (long long)blockIdx.x * 256 + threadIdx.x;
  if (f >= total) return;
  const int k4 = (int)(f & 7);
  long long g = f >> 3;
  const int row = (int)(g % CV_M); g /= CV_M;
  const int t = (int)(g % 9); g /= 9;
  const int c = (int)(g % chunks);
  const int cot = (int)(g / chunks);
  const int co = cot * CV_M + row;
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = co < Cout ? w[((long long)co * Cin + c * CV_KC + 4 * k4 + i) * 9 + t] : 0.f;
  const float4 h = make_float4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
  const float4 l = make_float4(to_tf32(v[0] - h.x), to_tf32(v[1] - h.y), to_tf32(v[2] - h.z), to_tf32(v[3] - h.w));
  unsigned char* blk = img + (((size_t)cot * chunks + c) * 9 + t) * CV_A_BYTES;
  const int off = row * 128 + ((k4 ^ (row & 7)) << 4);
  *reinterpret_cast<float4*>(blk + off) = h;
  *reinterpret_cast<float4*>(blk + CV_M * 128 + off) = l;
}

}  // namespace

}  // namespace cutie

using namespace cutie;

extern "C" int64_t cutie_conv3x3_weight_image_bytes(int64_t Cout, int64_t Cin) {
  if (Cout < 1 || Cin < CV_KC || Cin % CV_KC) return -1;
  return ((Cout + CV_M - 1) / CV_M) * (Cin / CV_KC) * 9 * (int64_t)CV_A_BYTES;
}

extern "C" int cutie_conv3x3_weight_image(const float* weight, int64_t Cout, int64_t Cin, void* image, void* stream) {
  CUTIE_REQUIRE(weight && image, "null argument");
  CUTIE_REQUIRE(Cout >= 1 && Cin >= CV_KC && Cin % CV_KC == 0, "input channels must be a multiple of 32");
  const long long total = (Cout + CV_M - 1) / CV_M * CV_M * (Cin / CV_KC) * 9 * 8;
  conv3x3_weight_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      weight, (int)Cout, (int)Cin, static_cast<unsigned char*>(image));
  CUTIE_CHECK_LAUNCH();
  return 0;
}

// spatial tile: TW | tile width (full rows when they fit), TH rows, N = round16(TH * (TW + 2)) <= 128 and
// N + 2 (TW + 2) + 2 <= CV_ROWS; maximise useful pixels per MMA column over the whole image (edge tiles included)
static void conv_tile_shape(int H, int W, int* TH, int* TW, int* N) {
  double best = -1;
  for (int parts = 1; parts <= W; ++parts) {
    const int tw = (W + parts - 1) / parts;
    for (int th = 1; th <= H && th * (tw + 2) <= 128; ++th) {
      const int n = (th * (tw + 2) + 15) / 16 * 16;
      if (n + 2 * (tw + 2) + 2 > CV_ROWS) continue;
      const long long tiles = (long long)((H + th - 1) / th) * ((W + tw - 1) / tw);
      const double eff = (double)H * W / ((double)tiles * n);
      // the most useful pixels per MMA column; among equals the larger N (fewer CTAs re-reading the weights)
      if (eff > best + 1e-9 || (eff > best - 1e-9 && n > *N)) { best = eff; *TH = th; *TW = tw; *N = n; }
    }
    if (tw <= 8) break;
  }
}

extern "C" int cutie_conv3x3_tc(const float* x, const void* weight_image, const float* bias, const float* residual,
                                int64_t NB, int64_t Cin, int64_t Cout, int64_t H, int64_t W, int relu_in, int relu_out,
                                float* y, void* stream) {
  CUTIE_REQUIRE(x && weight_image && y, "null argument");
  CUTIE_REQUIRE(Cout >= 1 && Cin >= CV_KC && Cin % CV_KC == 0, "input channels must be a multiple of 32");
  CUTIE_REQUIRE(NB >= 1 && NB <= 65535 && H >= 1 && W >= 1 && H * W < (1ll << 30), "bad geometry");
  ConvTcParams p;
  p.x = x; p.wimg = static_cast<const unsigned char*>(weight_image); p.bias = bias; p.z = residual; p.y = y;
  p.Cin = (int)Cin; p.Cout = (int)Cout; p.H = (int)H; p.W = (int)W;
  p.TH = 0; p.TW = 0; p.N = 0;
  conv_tile_shape(p.H, p.W, &p.TH, &p.TW, &p.N);
  CUTIE_REQUIRE(p.N >= 16, "no tile shape for this geometry");
  p.tiles_x = (p.W + p.TW - 1) / p.TW;
  p.relu_in = relu_in; p.relu_out = relu_out;
  const long long tiles = (long long)p.tiles_x * ((p.H + p.TH - 1) / p.TH);
  CUTIE_REQUIRE(tiles <= 0x7fffffff, "too many tiles");
  static bool attr_done[64] = {};
  if (first_use_on_device(attr_done))
    cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM);
  conv3x3_tc_kernel<<<dim3((unsigned)tiles, (unsigned)((Cout + CV_M - 1) / CV_M), (unsigned)NB), CV_THREADS, CV_SMEM,
                      (cudaStream_t)stream>>>(p);
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_debug_conv_tile_shape(int64_t H, int64_t W, int* out3) {
  int th = 0, tw = 0, n = 0;
  conv_tile_shape((int)H, (int)W, &th, &tw, &n);
  out3[0] = th; out3[1] = tw; out3[2] = n;
  return 0;
}
