// Pixel-side glue on the frame path that ATen parallelises badly at these shapes (sm_100a).
//
// upsample2x_add: the mask decoder's UpsampleBlock input  out = bilinear_x2(g) + skip  for every object
// (cutie/model/modules.py:8-20: F.interpolate(scale_factor=2, mode='bilinear', align_corners=False), then
// `skip_f + g` broadcast over objects).  ATen's upsample_bilinear2d_out_frame launches one thread per OUTPUT
// PIXEL and loops over batch x channels inside it: 7 CTAs (337 us) for [3 x 256 x 30 x 54] and 26 CTAs (169 us)
// for [3 x 256 x 60 x 108] on a 148-SM part.  Here: one thread per four consecutive output pixels of one
// (object, channel) plane, float4 stores, the add fused -- the kernel is a pure HBM stream.
#include <math_constants.h>

#include "common.cuh"

namespace cutie {

// PyTorch's area_pixel_compute_source_index for align_corners=False, scale = 1/2
__device__ __forceinline__ void src_index(int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = 0.5f * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

template <int V>
__global__ void __launch_bounds__(256) upsample2x_add_kernel(const float* __restrict__ g, const float* __restrict__ skip,
                                                             float* __restrict__ out, long long planes, int K, int C,
                                                             int h, int w) {
  const int W2 = 2 * w, H2 = 2 * h;
  const int xv = W2 / V;                                   // vectors per output row
  const long long total = planes * H2 * xv;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int xq = (int)(t % xv);
  const int Y = (int)((t / xv) % H2);
  const long long plane = t / ((long long)xv * H2);         // (b*K + k)*C + c
  const long long c = plane % C, bk = plane / C, b = bk / K;
  const float* gp = g + plane * (long long)h * w;
  const float* sp = skip + ((b * C + c) * (long long)H2 + Y) * W2 + (long long)xq * V;
  int y0, y1;
  float hy0, hy1;
  src_index(Y, h, y0, y1, hy0, hy1);
  const float* r0 = gp + (long long)y0 * w;
  const float* r1 = gp + (long long)y1 * w;
  float res[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    int x0, x1;
    float wx0, wx1;
    src_index(xq * V + i, w, x0, x1, wx0, wx1);
    const float v = hy0 * (wx0 * __ldg(r0 + x0) + wx1 * __ldg(r0 + x1)) + hy1 * (wx0 * __ldg(r1 + x0) + wx1 * __ldg(r1 + x1));
    res[i] = v + __ldg(sp + i);
  }
  float* op = out + (plane * (long long)H2 + Y) * W2 + (long long)xq * V;
  if (V == 4) {
    *reinterpret_cast<float4*>(op) = make_float4(res[0], res[1], res[2], res[3]);
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) op[i] = res[i];
  }
}

// out[y,x] = lut[argmax_c prob[c,y,x]] (first maximum wins, like torch.argmax): InferenceCore.output_prob_to_mask
// (inference_core.py:377-385 + object_manager.py:99-104) as one pass instead of a strided reduce, a gather and a cast.
__global__ void __launch_bounds__(256) prob_to_mask_kernel(const float* __restrict__ prob, long long plane_stride,
                                                           long long row_stride, int C, int H, int W,
                                                           const long long* __restrict__ lut,
                                                           long long* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)H * W) return;
  const int y = (int)(i / W), x = (int)(i % W);
  const float* p = prob + (long long)y * row_stride + x;
  float best = __ldg(p);
  int arg = 0;
  for (int c = 1; c < C; ++c) {
    const float v = __ldg(p + (long long)c * plane_stride);
    if (v > best) { best = v; arg = c; }
  }
  out[i] = lut[arg];
}

// y = act(y + bias[c] (+ z)) in place -- the epilogue of a cuDNN convolution that was called WITHOUT its bias.
// PyTorch's cudnn_convolution adds the bias with a broadcast TensorIterator kernel (elementwise_kernel<128,2>, offset
// calculator per element, no vector accesses: 6.5 us for a 5 MB map, 29 us for 40 MB) and the ReLU / residual add with
// further launches; this is one float4 grid-stride stream.  Same association as ATen: (y + bias) + z, then the clamp;
// NaN propagates like clamp_min.  CL = channels-last storage [N, HW, C]; otherwise [N, C, HW].
template <bool CL, bool VEC>
__global__ void __launch_bounds__(256) bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias,
                                                       const float* __restrict__ z, long long total, int C,
                                                       long long HW, int relu) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    if (VEC) {
      const long long e = i * 4;
      float4 v = reinterpret_cast<const float4*>(y)[i];
      float4 b;
      if (CL) {
        b = __ldg(reinterpret_cast<const float4*>(bias + (int)(e % C)));
      } else {
        const float bb = __ldg(bias + (int)((e / HW) % C));
        b = make_float4(bb, bb, bb, bb);
      }
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      if (z) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(z) + i);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      if (relu) {
        v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y;
        v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
      }
      reinterpret_cast<float4*>(y)[i] = v;
    } else {
      float v = y[i] + __ldg(bias + (CL ? (int)(i % C) : (int)((i / HW) % C)));
      if (z) v += __ldg(z + i);
      if (relu) v = v < 0.f ? 0.f : v;
      y[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Integer-factor area down-sampling (F.interpolate(mode='area') == adaptive_avg_pool2d when the sizes divide):
// out[p, Y, X] = mean of the f x f window.  ATen's adaptive_average_pool kernel assigns whole planes to a handful of
// CTAs: 15 CTAs / 46 us for the [3, 480, 864] -> [3, 30, 54] mask of cutie.py:149, ~33 us for the decoder's
// [3, 256, 60, 108] / [3, 257, 120, 216] feature maps (modules.py:58-60).  One thread per output pixel here, rows
// of the window read as float4 when the geometry allows; accumulation row-major in fp32, then one division.
template <int F, bool VEC>
__global__ void __launch_bounds__(128) area_pool_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        long long total, int Ho, int Wo, int f_rt) {
  const int f = F > 0 ? F : f_rt;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int X = (int)(t % Wo);
  const int Y = (int)((t / Wo) % Ho);
  const long long p = t / ((long long)Wo * Ho);
  const long long W = (long long)Wo * f;
  const float* src = in + (p * Ho * f + (long long)Y * f) * W + (long long)X * f;
  float s = 0.f;
  if (VEC) {                                               // f % 4 == 0, rows 16-byte aligned
#pragma unroll
    for (int r = 0; r < (F > 0 ? F : 1); ++r) {
#pragma unroll
      for (int c4 = 0; c4 < (F > 0 ? F / 4 : 1); ++c4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * W) + c4);
        s += v.x; s += v.y; s += v.z; s += v.w;
      }
    }
  } else {
    for (int r = 0; r < f; ++r)
      for (int c = 0; c < f; ++c) s += __ldg(src + r * W + c);
  }
  out[t] = s / (float)(f * f);
}

// ------------------------------------------------------------------------------------------------
// Tail of the channel-attention residual block (cutie/model/channel_attn.py:27-38, CAResBlock):
//   gate[n, c] = sigmoid( sum_j w[j] * mean[n, c + j - (k-1)/2] )      (Conv1d over channels, zero padded, no bias)
//   out        = y * gate + x                                           (in place into y)
// ATen: conv1d + sigmoid + mul + add = 4 launches and 5 passes over the feature map; here a (tiny) gate kernel and one
// stream.  y * gate is rounded before the add, as in ATen (no FMA contraction).
__global__ void __launch_bounds__(256) eca_gate_kernel(const float* __restrict__ mean, const float* __restrict__ w,
                                                       float* __restrict__ gate, long long total, int C, int k) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const int pad = (k - 1) / 2;
  float a = 0.f;
  for (int j = 0; j < k; ++j) {
    const int cc = c + j - pad;
    if (cc >= 0 && cc < C) a = fmaf(__ldg(w + j), __ldg(mean + t + (j - pad)), a);
  }
  gate[t] = 1.f / (1.f + expf(-a));
}

template <bool CL, bool VEC>
__global__ void __launch_bounds__(256) scale_add_kernel(float* __restrict__ y, const float* __restrict__ gate,
                                                        const float* __restrict__ x, long long total, int C,
                                                        long long HW) {
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    if (VEC) {
      const long long e = i * 4;
      float4 v = reinterpret_cast<const float4*>(y)[i];
      const float4 r = __ldg(reinterpret_cast<const float4*>(x) + i);
      float4 g;
      if (CL) {
        g = __ldg(reinterpret_cast<const float4*>(gate + (e / (HW * C)) * C + (e % C)));
      } else {
        const float gg = __ldg(gate + e / HW);
        g = make_float4(gg, gg, gg, gg);
      }
      v.x = __fadd_rn(__fmul_rn(v.x, g.x), r.x); v.y = __fadd_rn(__fmul_rn(v.y, g.y), r.y);
      v.z = __fadd_rn(__fmul_rn(v.z, g.z), r.z); v.w = __fadd_rn(__fmul_rn(v.w, g.w), r.w);
      reinterpret_cast<float4*>(y)[i] = v;
    } else {
      const float g = __ldg(gate + (CL ? (i / (HW * C)) * C + (i % C) : i / HW));
      y[i] = __fadd_rn(__fmul_rn(y[i], g), __ldg(x + i));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GRU-like sensory update (cutie/model/modules.py:37-45 _recurrent_update): v [P, 3d, HW] = [forget | update | candidate],
//   out = sigmoid(vf) * h * (1 - sigmoid(vu)) + sigmoid(vu) * tanh(vn)
// ATen: 2 sigmoid + tanh + mul + rsub + mul + mul + add = 8 launches; every product / sum is rounded as ATen rounds it.
template <bool VEC>
__global__ void __launch_bounds__(256) gated_update_kernel(const float* __restrict__ v, const float* __restrict__ h,
                                                           float* __restrict__ out, long long total, int d,
                                                           long long HW) {
  const long long step = (long long)gridDim.x * blockDim.x;
  const long long plane = (long long)d * HW;                   // elements of h per object
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const long long e = VEC ? i * 4 : i;
    const long long p = e / plane, r = e % plane;
    const float* vf = v + p * 3 * plane + r;
    float a[4], b[4], c[4], hh[4], o[4];
    if (VEC) {
      const float4 A = __ldg(reinterpret_cast<const float4*>(vf));
      const float4 B = __ldg(reinterpret_cast<const float4*>(vf + plane));
      const float4 Cc = __ldg(reinterpret_cast<const float4*>(vf + 2 * plane));
      const float4 Hh = __ldg(reinterpret_cast<const float4*>(h) + i);
      a[0] = A.x; a[1] = A.y; a[2] = A.z; a[3] = A.w;
      b[0] = B.x; b[1] = B.y; b[2] = B.z; b[3] = B.w;
      c[0] = Cc.x; c[1] = Cc.y; c[2] = Cc.z; c[3] = Cc.w;
      hh[0] = Hh.x; hh[1] = Hh.y; hh[2] = Hh.z; hh[3] = Hh.w;
    } else {
      a[0] = __ldg(vf); b[0] = __ldg(vf + plane); c[0] = __ldg(vf + 2 * plane); hh[0] = __ldg(h + i);
    }
#pragma unroll
    for (int q = 0; q < (VEC ? 4 : 1); ++q) {
      const float f = 1.f / (1.f + expf(-a[q]));
      const float u = 1.f / (1.f + expf(-b[q]));
      const float n = tanhf(c[q]);
      o[q] = __fadd_rn(__fmul_rn(__fmul_rn(f, hh[q]), __fsub_rn(1.f, u)), __fmul_rn(u, n));
    }
    if (VEC) reinterpret_cast<float4*>(out)[i] = make_float4(o[0], o[1], o[2], o[3]);
    else out[i] = o[0];
  }
}

// ------------------------------------------------------------------------------------------------
// ResNet stem tail: out = relu(maxpool3x3/s2/p1(y) + bias[c]) for a bias-less convolution output y.  Equal, bit for bit,
// to maxpool(relu(y + bias)) (adding a constant and clamping are monotone, so they commute with max) -- i.e. to
// utils/resnet.py:139-142 `relu(bn1(conv1(x)))` -> `maxpool` with the BatchNorm folded into the convolution.  ATen
// spends three launches here (broadcast bias add 18.5 us, clamp, max_pool_forward_nhwc 41 us for [1,64,240,432] at
// 480p); this is one pass: 26.5 MB in, 6.6 MB out.  CL: storage [N,H,W,C], one thread per 4 channels of an output
// pixel; otherwise [N,C,H,W], one thread per output pixel.
template <bool CL>
__global__ void __launch_bounds__(256) bias_relu_maxpool_kernel(const float* __restrict__ y, const float* __restrict__ bias,
                                                                float* __restrict__ out, long long total, int C, int H,
                                                                int W, int Ho, int Wo) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  if (CL) {
    const int C4 = C >> 2;
    const int c4 = (int)(t % C4);
    const int X = (int)((t / C4) % Wo);
    const int Y = (int)((t / ((long long)C4 * Wo)) % Ho);
    const long long n = t / ((long long)C4 * Wo * Ho);
    float4 m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = 2 * Y - 1 + dy;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = 2 * X - 1 + dx;
        if (xx < 0 || xx >= W) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(y + ((n * H + yy) * W + xx) * C) + c4);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    m.x += b.x; m.y += b.y; m.z += b.z; m.w += b.w;
    m.x = m.x < 0.f ? 0.f : m.x; m.y = m.y < 0.f ? 0.f : m.y; m.z = m.z < 0.f ? 0.f : m.z; m.w = m.w < 0.f ? 0.f : m.w;
    reinterpret_cast<float4*>(out)[t] = m;
  } else {
    const int X = (int)(t % Wo);
    const int Y = (int)((t / Wo) % Ho);
    const long long plane = t / ((long long)Wo * Ho);              // n * C + c
    const float* src = y + plane * (long long)H * W;
    float m = -CUDART_INF_F;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = 2 * Y - 1 + dy;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = 2 * X - 1 + dx;
        if (xx < 0 || xx >= W) continue;
        m = fmaxf(m, __ldg(src + (long long)yy * W + xx));
      }
    }
    m += __ldg(bias + (int)(plane % C));
    out[t] = m < 0.f ? 0.f : m;
  }
}

// ------------------------------------------------------------------------------------------------
// CUTIE.segment tail (cutie/model/cutie.py:196-203 + cutie/utils/tensor_utils.py:47-54):
//   prob = sigmoid(x);  bg = prod_k (1 - prob_k);  a = clamp([bg | prob], 1e-7, 1 - 1e-7);  l = log(a / (1 - a))
//   logits = bilinear_x4(l) (align_corners = False);  P = softmax over the 1+K channels
// ATen: 8 tiny launches for the aggregation, upsample_bilinear2d (13.7 us) and a spatial softmax (12.5 us) at 480p.
// Kernel A: one thread per low-resolution pixel; kernel B: one thread per output pixel, all channels in registers.
constexpr int SEG_MAXC = 16;

__global__ void __launch_bounds__(256) aggregate_logits_kernel(const float* __restrict__ x, float* __restrict__ agg,
                                                               long long total, int K, long long hw) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b, pixel)
  if (t >= total) return;
  const long long b = t / hw, px = t % hw;
  const float* xp = x + b * K * hw + px;
  float* ap = agg + b * (K + 1) * hw + px;
  float bg = 1.f;
  for (int k = 0; k < K; ++k) {
    const float p = 1.f / (1.f + expf(-__ldg(xp + (long long)k * hw)));
    bg = __fmul_rn(bg, __fsub_rn(1.f, p));
    const float a = fminf(fmaxf(p, 1e-7f), 0.9999999f);
    ap[(long long)(k + 1) * hw] = logf(__fdiv_rn(a, __fsub_rn(1.f, a)));
  }
  const float a = fminf(fmaxf(bg, 1e-7f), 0.9999999f);
  ap[0] = logf(__fdiv_rn(a, __fsub_rn(1.f, a)));
}

// PyTorch's area_pixel_compute_source_index for align_corners=False, scale = 1/4
__device__ __forceinline__ void src_index4(int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = 0.25f * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) upsample4_softmax_kernel(const float* __restrict__ agg, float* __restrict__ logits,
                                                                float* __restrict__ prob, long long total, int C, int h,
                                                                int w) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b, Y, X)
  if (t >= total) return;
  const int W4 = 4 * w, H4 = 4 * h;
  const int X = (int)(t % W4);
  const int Y = (int)((t / W4) % H4);
  const long long b = t / ((long long)W4 * H4);
  int y0, y1, x0, x1;
  float hy0, hy1, wx0, wx1;
  src_index4(Y, h, y0, y1, hy0, hy1);
  src_index4(X, w, x0, x1, wx0, wx1);
  const long long hw = (long long)h * w, HW4 = (long long)H4 * W4;
  const float* base = agg + b * C * hw;
  float v[SEG_MAXC];
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int c = 0; c < SEG_MAXC; ++c) {
    if (c < C) {
      const float* p = base + (long long)c * hw;
      const float r = hy0 * (wx0 * __ldg(p + (long long)y0 * w + x0) + wx1 * __ldg(p + (long long)y0 * w + x1)) +
                      hy1 * (wx0 * __ldg(p + (long long)y1 * w + x0) + wx1 * __ldg(p + (long long)y1 * w + x1));
      v[c] = r;
      mx = fmaxf(mx, r);
      logits[(b * C + c) * HW4 + (long long)Y * W4 + X] = r;
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < SEG_MAXC; ++c) {
    if (c < C) {
      v[c] = expf(v[c] - mx);
      sum += v[c];
    }
  }
#pragma unroll
  for (int c = 0; c < SEG_MAXC; ++c)
    if (c < C) prob[(b * C + c) * HW4 + (long long)Y * W4 + X] = v[c] / sum;
}

// ------------------------------------------------------------------------------------------------
// The mask decoder's prediction head (cutie/model/big_modules.py:264,300: `self.pred(F.relu(p4))`, Conv2d(C, 1, 3, padding=1)):
// a 3x3 convolution with ONE output channel over [planes, C, H, W].  cuDNN treats it as a GEMM with N = 1 and wraps
// it in NCHW<->NHWC transposes of the 40 MB input and of the weight (18.5 + 10.2 + 39.9 + 6.2 us at 480p, after an
// 11.4 us clamp and before a 3.7 us bias add); it is really a 9-tap weighted sum over channels -- one pass over the
// input.  A CTA owns 32 consecutive pixels of one row; its 8 warps take the channels c = warp, warp + 8, ... (lane ==
// pixel: every load is a coalesced 128-byte row segment), each thread keeps a 9-tap fp32 FMA chain over its channels and
// the 8 partial sums are added in warp order (deterministic).  The ReLU of the input is applied on the fly (zero padding
// is applied to the rectified input, as F.relu -> Conv2d(padding=1) does).  (One thread per pixel over all channels --
// the first version -- was a 1152-step dependent chain per thread: 90 us at 480p for a 40 MB read.)
__global__ void __launch_bounds__(256) conv3x3_c1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         int C, int H, int W, int relu_input) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int X = blockIdx.x * 32 + lane, Y = blockIdx.y;
  const long long p = blockIdx.z;
  const long long hw = (long long)H * W;
  const bool live = X < W;
  const bool y0 = Y > 0, y2 = Y < H - 1, x0 = live && X > 0, x2 = live && X < W - 1;
  float acc = 0.f;
  if (live) {
    const float* xp = x + p * C * hw + (long long)Y * W + X;
    for (int c = g; c < C; c += 8) {
      const float* r = xp + (long long)c * hw;
      const float* k = w + c * 9;
      float v[9];
      v[0] = (y0 && x0) ? __ldg(r - W - 1) : 0.f; v[1] = y0 ? __ldg(r - W) : 0.f; v[2] = (y0 && x2) ? __ldg(r - W + 1) : 0.f;
      v[3] = x0 ? __ldg(r - 1) : 0.f;             v[4] = __ldg(r);                v[5] = x2 ? __ldg(r + 1) : 0.f;
      v[6] = (y2 && x0) ? __ldg(r + W - 1) : 0.f; v[7] = y2 ? __ldg(r + W) : 0.f; v[8] = (y2 && x2) ? __ldg(r + W + 1) : 0.f;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float a = relu_input ? (v[i] < 0.f ? 0.f : v[i]) : v[i];
        acc = fmaf(__ldg(k + i), a, acc);
      }
    }
  }
  part[g][lane] = acc;
  __syncthreads();
  if (g == 0 && live) {
    float s = part[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) s += part[i][lane];
    out[p * hw + (long long)Y * W + X] = s + __ldg(bias);
  }
}

}  // namespace cutie

using namespace cutie;

extern "C" int cutie_bias_act(float* y, const float* bias, const float* z, int64_t N, int64_t C, int64_t HW,
                              int channels_last, int relu, void* stream) {
  CUTIE_REQUIRE(y && bias && N >= 1 && C >= 1 && HW >= 1, "null/empty argument");
  CUTIE_REQUIRE(C < (1LL << 31) && N * C * HW < (1LL << 60), "tensor too large");
  const long long n = (long long)N * C * HW;
  const bool aligned = (((uintptr_t)y | (uintptr_t)(z ? z : y)) & 15) == 0;
  const bool vec = aligned && (channels_last ? (C % 4 == 0 && ((uintptr_t)bias & 15) == 0) : (HW % 4 == 0));
  const long long total = vec ? n / 4 : n;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (channels_last) {
    if (vec) bias_act_kernel<true, true><<<(unsigned)blocks, 256, 0, st>>>(y, bias, z, total, (int)C, HW, relu);
    else bias_act_kernel<true, false><<<(unsigned)blocks, 256, 0, st>>>(y, bias, z, total, (int)C, HW, relu);
  } else {
    if (vec) bias_act_kernel<false, true><<<(unsigned)blocks, 256, 0, st>>>(y, bias, z, total, (int)C, HW, relu);
    else bias_act_kernel<false, false><<<(unsigned)blocks, 256, 0, st>>>(y, bias, z, total, (int)C, HW, relu);
  }
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_prob_to_mask(const float* prob, int64_t plane_stride, int64_t row_stride, int64_t C, int64_t H,
                                  int64_t W, const int64_t* lut, int64_t* out, void* stream) {
  CUTIE_REQUIRE(prob && lut && out && C >= 1 && H >= 1 && W >= 1, "null/empty argument");
  const long long n = (long long)H * W;
  prob_to_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      prob, plane_stride, row_stride, (int)C, (int)H, (int)W, (const long long*)lut, (long long*)out);
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_upsample2x_add(const float* g, const float* skip, float* out, int64_t B, int64_t K, int64_t C,
                                    int64_t h, int64_t w, void* stream) {
  CUTIE_REQUIRE(g && skip && out && B >= 1 && K >= 1 && C >= 1 && h >= 1 && w >= 1, "null/empty argument");
  CUTIE_REQUIRE(h < (1 << 14) && w < (1 << 14), "feature map too large");
  const long long planes = (long long)B * K * C;
  const bool vec = ((2 * w) % 4 == 0) && (((uintptr_t)out & 15) == 0);
  const long long total = planes * 2 * h * (vec ? (2 * w) / 4 : 2 * w);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (vec)
    upsample2x_add_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(g, skip, out, planes, (int)K, (int)C, (int)h, (int)w);
  else
    upsample2x_add_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(g, skip, out, planes, (int)K, (int)C, (int)h, (int)w);
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_area_pool(const float* in, float* out, int64_t planes, int64_t H, int64_t W, int64_t f, void* stream) {
  CUTIE_REQUIRE(in && out && planes >= 1 && H >= 1 && W >= 1 && f >= 1, "null/empty argument");
  CUTIE_REQUIRE(H % f == 0 && W % f == 0, "sizes must be multiples of the factor");
  CUTIE_REQUIRE(H < (1 << 20) && W < (1 << 20) && f <= 64, "feature map / factor too large");
  const int Ho = (int)(H / f), Wo = (int)(W / f);
  const long long total = (long long)planes * Ho * Wo;
  const unsigned blocks = (unsigned)((total + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (f % 4 == 0) && (W % 4 == 0) && (((uintptr_t)in & 15) == 0);
  if (f == 16 && vec) area_pool_kernel<16, true><<<blocks, 128, 0, st>>>(in, out, total, Ho, Wo, (int)f);
  else if (f == 4 && vec) area_pool_kernel<4, true><<<blocks, 128, 0, st>>>(in, out, total, Ho, Wo, (int)f);
  else if (f == 2) area_pool_kernel<2, false><<<blocks, 128, 0, st>>>(in, out, total, Ho, Wo, (int)f);
  else area_pool_kernel<0, false><<<blocks, 128, 0, st>>>(in, out, total, Ho, Wo, (int)f);
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_eca_scale_add(float* y, const float* x, const float* mean, const float* w, float* gate, int64_t N,
                                   int64_t C, int64_t HW, int64_t k, int channels_last, void* stream) {
  CUTIE_REQUIRE(y && x && mean && w && gate && N >= 1 && C >= 1 && HW >= 1, "null/empty argument");
  CUTIE_REQUIRE(k >= 1 && k <= 15 && (k & 1), "kernel size must be odd and <= 15");
  CUTIE_REQUIRE(C < (1LL << 31) && N * C * HW < (1LL << 60), "tensor too large");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nc = (long long)N * C;
  eca_gate_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(mean, w, gate, nc, (int)C, (int)k);
  CUTIE_CHECK_LAUNCH();
  const long long n = nc * HW;
  const bool aligned = (((uintptr_t)y | (uintptr_t)x) & 15) == 0;
  const bool vec = aligned && (channels_last ? (C % 4 == 0 && ((uintptr_t)gate & 15) == 0) : (HW % 4 == 0));
  const long long total = vec ? n / 4 : n;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (channels_last) {
    if (vec) scale_add_kernel<true, true><<<(unsigned)blocks, 256, 0, st>>>(y, gate, x, total, (int)C, HW);
    else scale_add_kernel<true, false><<<(unsigned)blocks, 256, 0, st>>>(y, gate, x, total, (int)C, HW);
  } else {
    if (vec) scale_add_kernel<false, true><<<(unsigned)blocks, 256, 0, st>>>(y, gate, x, total, (int)C, HW);
    else scale_add_kernel<false, false><<<(unsigned)blocks, 256, 0, st>>>(y, gate, x, total, (int)C, HW);
  }
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_gated_update(const float* v, const float* h, float* out, int64_t P, int64_t d, int64_t HW,
                                  void* stream) {
  CUTIE_REQUIRE(v && h && out && P >= 1 && d >= 1 && HW >= 1, "null/empty argument");
  CUTIE_REQUIRE(d < (1LL << 31) && P * d * HW < (1LL << 59), "tensor too large");
  const long long n = (long long)P * d * HW;
  const bool vec = (HW % 4 == 0) && ((((uintptr_t)v | (uintptr_t)h | (uintptr_t)out) & 15) == 0);
  const long long total = vec ? n / 4 : n;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) gated_update_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(v, h, out, total, (int)d, HW);
  else gated_update_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(v, h, out, total, (int)d, HW);
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_bias_relu_maxpool(const float* y, const float* bias, float* out, int64_t N, int64_t C, int64_t H,
                                       int64_t W, int channels_last, void* stream) {
  CUTIE_REQUIRE(y && bias && out && N >= 1 && C >= 1 && H >= 1 && W >= 1, "null/empty argument");
  CUTIE_REQUIRE(H < (1 << 20) && W < (1 << 20) && C < (1 << 20), "feature map too large");
  const int Ho = (int)((H - 1) / 2 + 1), Wo = (int)((W - 1) / 2 + 1);
  cudaStream_t st = (cudaStream_t)stream;
  if (channels_last) {
    CUTIE_REQUIRE(C % 4 == 0 && ((((uintptr_t)y | (uintptr_t)out | (uintptr_t)bias) & 15) == 0),
                  "channels-last path needs C % 4 == 0 and 16-byte aligned pointers");
    const long long total = (long long)N * Ho * Wo * (C / 4);
    bias_relu_maxpool_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(y, bias, out, total, (int)C, (int)H,
                                                                                   (int)W, Ho, Wo);
  } else {
    const long long total = (long long)N * C * Ho * Wo;
    bias_relu_maxpool_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(y, bias, out, total, (int)C, (int)H,
                                                                                    (int)W, Ho, Wo);
  }
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_segment_tail(const float* x, float* agg, float* logits, float* prob, int64_t B, int64_t K, int64_t h,
                                  int64_t w, void* stream) {
  CUTIE_REQUIRE(x && agg && logits && prob && B >= 1 && K >= 1 && h >= 1 && w >= 1, "null/empty argument");
  CUTIE_REQUIRE(K + 1 <= SEG_MAXC, "at most 15 objects per call");
  CUTIE_REQUIRE(h < (1 << 18) && w < (1 << 18), "feature map too large");
  cudaStream_t st = (cudaStream_t)stream;
  const long long lo = (long long)B * h * w;
  aggregate_logits_kernel<<<(unsigned)((lo + 255) / 256), 256, 0, st>>>(x, agg, lo, (int)K, (long long)h * w);
  CUTIE_CHECK_LAUNCH();
  const long long hi = lo * 16;
  upsample4_softmax_kernel<<<(unsigned)((hi + 255) / 256), 256, 0, st>>>(agg, logits, prob, hi, (int)(K + 1), (int)h, (int)w);
  CUTIE_CHECK_LAUNCH();
  return 0;
}

extern "C" int cutie_conv3x3_c1(const float* x, const float* w, const float* bias, float* out, int64_t planes, int64_t C,
                                int64_t H, int64_t W, int relu_input, void* stream) {
  CUTIE_REQUIRE(x && w && bias && out && planes >= 1 && C >= 1 && H >= 1 && W >= 1, "null/empty argument");
  CUTIE_REQUIRE(H < (1 << 20) && W < (1 << 20) && C < (1 << 20), "feature map too large");
  CUTIE_REQUIRE(H <= 65535 && planes <= 65535, "at most 65535 rows / planes");
  conv3x3_c1_kernel<<<dim3((unsigned)((W + 31) / 32), (unsigned)H, (unsigned)planes), 256, 0, (cudaStream_t)stream>>>(
      x, w, bias, out, (int)C, (int)H, (int)W, relu_input);
  CUTIE_CHECK_LAUNCH();
  return 0;
}
