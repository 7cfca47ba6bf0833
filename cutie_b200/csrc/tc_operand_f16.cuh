// Shared-memory operand image of the tcgen05 affinity filter in FP16 (sm_100a): one definition used by the memory-bank
// key-image builder (bank.cu) and by the filter (affinity_f16.cu), whose producer is one bulk copy per tile.
//
// Why FP16 and not TF32: both round to 11 significant bits, but kind::f16 issues at twice the kind::tf32 rate and the
// operand bytes halve -- the TF32 filter of round 1 was bound by the L2 -> SM path (69.6 KB per 128-token tile against
// ~42 B/clk/SM), not by the tensor pipe.  Accumulation is fp32 in TMEM either way.  FP16's narrow exponent range is
// handled rigorously: subnormal rounding gets an absolute error term, saturated rows are flagged "always a candidate".
//
// A memory-token tile = 128 rows (tokens) x K = 144 f16, K-major:
//   2 x [128 rows x 128 B] SWIZZLE_128B K-blocks : [shr k_c^2 (c = 0..63)] [shr k_c (c = 0..63)]
//   1 x [128 rows x  32 B] un-swizzled tail block : [shr, shr, -eps P^2, -2 eps P R, -eps R^2, -absB, 1, satflag | 0 x 8]
// with P = sqrt(shr |k|^2), R = sqrt(shr) (rounded up), absB = 2^-25 * ||row||_1 (the subnormal-rounding term).
// The query operand pairs the tail with [b2_hi, b2_lo, s, s v, s v^2, s, -s absA, s] where s = +1 makes the MMA emit a LOWER
// bound of the exact energy (candidate filter) and s = -1 an UPPER bound (threshold sampling).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cutie {

constexpr int F16_KTILE = 128;                   // memory tokens per tile (MMA N)
constexpr int F16_BLK_BYTES = 128 * 128;         // one SW128 K-block: 128 rows x 64 f16
constexpr int F16_TAIL_BYTES = 128 * 32;         // tail block: 128 rows x 16 f16
constexpr int F16_OPER_BYTES = 2 * F16_BLK_BYTES + F16_TAIL_BYTES;   // 36864 bytes per 128-row operand
// Both operands are rounded to nearest f16 (<= 2^-11 each, normal range) => product error <= 2^-10 (1 + 2^-12);
// + the tensor core's fp32 accumulation of 144 terms (aligned adds, <= ~2^-22 per partial sum) + the rounding of the
// bound's own operands (always rounded away from zero).
constexpr float F16_EPS = 1.05e-3f;
constexpr float F16_ABS = 1.1f / 33554432.f;     // 2^-25 (half an f16 subnormal ulp), padded
constexpr float F16_SAT = 60000.f;               // |operand| beyond this does not fit f16: the row is flagged

__device__ __forceinline__ float fsqrt_approx_up(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 1.002f;
}
// round away from zero to f16 (bound terms must never shrink)
__device__ __forceinline__ __half h_up(float x) { return x >= 0.f ? __float2half_ru(x) : __float2half_rd(x); }
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ uint32_t pack_rn(float a, float b) { return pack_h2(__float2half_rn(a), __float2half_rn(b)); }

// byte offsets inside an operand buffer (f16 elements)
__device__ __forceinline__ int f16_off_main(int row, int elem) {     // elem in [0,128): 2 K-blocks of 64
  const int blk = elem >> 6, w = elem & 63;
  return blk * F16_BLK_BYTES + row * 128 + ((((w >> 3) ^ (row & 7))) << 4) + (w & 7) * 2;
}
__device__ __forceinline__ int f16_off_tail(int row, int elem) {     // elem in [0,16)
  return 2 * F16_BLK_BYTES + (elem >> 3) * 2048 + (row >> 3) * 128 + (row & 7) * 16 + (elem & 7) * 2;
}

// One token row handled by 16 consecutive lanes (lane c4 = lane & 15 owns channels 4*c4 .. 4*c4+3; all 32 lanes of the
// warp must call this).  Writes the lane's two 8-byte pieces and, from lane c4 == 0, the row's tail.
__device__ __forceinline__ void store_key_row_operand_f16(unsigned char* tile, int row, int c4, float4 v, float shr,
                                                          bool do_store) {
  const float4 ln = make_float4(shr * v.x, shr * v.y, shr * v.z, shr * v.w);
  const float4 sq = make_float4(ln.x * v.x, ln.y * v.y, ln.z * v.z, ln.w * v.w);
  float n2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
  float n1 = fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w);
  float mx = fmaxf(fmaxf(fabsf(sq.x), fabsf(sq.y)), fmaxf(fabsf(sq.z), fabsf(sq.w)));
  mx = fmaxf(mx, fmaxf(fmaxf(fabsf(ln.x), fabsf(ln.y)), fmaxf(fabsf(ln.z), fabsf(ln.w))));
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) {
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (!do_store) return;
  // A row with an operand outside f16's range (or NaN / Inf) carries NO energy terms, only the flag: the filter then
  // sees D = -F16_SAT < any threshold (always a candidate, decided by the exact re-rank) and the threshold sampler
  // sees +F16_SAT (never a group minimum).
  const bool sat = !(mx <= F16_SAT) || !(shr <= F16_SAT);
  const int piece = row * 128 + ((((c4 >> 1) ^ (row & 7))) << 4) + (c4 & 1) * 8;
  *reinterpret_cast<uint2*>(tile + piece) = sat ? make_uint2(0u, 0u) : make_uint2(pack_rn(sq.x, sq.y), pack_rn(sq.z, sq.w));
  *reinterpret_cast<uint2*>(tile + F16_BLK_BYTES + piece) =
      sat ? make_uint2(0u, 0u) : make_uint2(pack_rn(ln.x, ln.y), pack_rn(ln.z, ln.w));
  if (c4 == 0) {
    const float Pn = fsqrt_approx_up(shr * n2), Rn = fsqrt_approx_up(shr);
    const float absb = F16_ABS * (shr * n2 + shr * n1 + 2.f * shr + 8.f);
    const __half sh = __float2half_rn(shr);
    uint4 t0 = make_uint4(0u, 0u, 0u, pack_h2(__float2half_rn(0.f), __float2half_rn(-F16_SAT)));
    const uint4 t1 = make_uint4(0u, 0u, 0u, 0u);
    if (!sat) {
      t0.x = pack_h2(sh, sh);
      t0.y = pack_h2(h_up(-F16_EPS * Pn * Pn), h_up(-2.f * F16_EPS * Pn * Rn));
      t0.z = pack_h2(h_up(-F16_EPS * Rn * Rn), h_up(-absb));
      t0.w = pack_h2(__float2half_rn(1.f), __float2half_rn(0.f));
    }
    *reinterpret_cast<uint4*>(tile + f16_off_tail(row, 0)) = t0;
    *reinterpret_cast<uint4*>(tile + f16_off_tail(row, 8)) = t1;
  }
}

}  // namespace cutie
