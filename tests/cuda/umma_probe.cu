// Hardware probe (test infrastructure): tcgen05.mma.kind::tf32 shared-memory descriptor encodings for the operand
// layouts the object-transformer kernels use (csrc/qt_tc.cu) -- K-major and MN-major SWIZZLE_128B, N = 64/128/256.
// The host lays the operands out byte-exactly, the kernel copies them to shared memory verbatim, issues the MMAs with
// the given descriptors and returns D; small-integer inputs make every TF32 product and fp32 sum exact.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/umma_probe tests/cuda/umma_probe.cu && gpurun_out/umma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Job {
  uint64_t adesc, bdesc;       // descriptors without the start address
  uint32_t idesc;
  int a_bytes, b_bytes;        // operand image sizes
  int nk;                      // number of MMA instructions
  int a_step[64], b_step[64];  // byte offset of the operand start for instruction i
  int N;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const unsigned char* a_img, const unsigned char* b_img, Job job,
                                                       float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_base;
  unsigned char* A = smem;
  unsigned char* B = smem + ((job.a_bytes + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid * 16; i < job.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(A + i) = *reinterpret_cast<const uint4*>(a_img + i);
  for (int i = tid * 16; i < job.b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(B + i) = *reinterpret_cast<const uint4*>(b_img + i);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    for (int i = 0; i < job.nk; ++i) {
      const uint64_t ad = job.adesc | (uint64_t)(((smem_u32(A) + job.a_step[i]) >> 4) & 0x3FFF);
      const uint64_t bd = job.bdesc | (uint64_t)(((smem_u32(B) + job.b_step[i]) >> 4) & 0x3FFF);
      const uint32_t acc = i ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(ad), "l"(bd), "r"(job.idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < job.N; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(size_t)tid * job.N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// ---- host-side layouts (tf32 elements in fp32 containers) ----
static int off_kmajor(int rows, int row, int k) {              // blocks of 32 k; SW128; 8-row groups 1024 B apart
  const int blk = k / 32, kk = k % 32;
  return blk * rows * 128 + row * 128 + ((((kk >> 2) ^ (row & 7))) << 4) + (kk & 3) * 4;
}
// MN-major: groups of 32 mn; inside a group k rows of 128 B.  variant < 2: SWIZZLE_128B (16-byte chunks XOR k % 8) -- NOT
// valid for tf32 (kept as a negative control); variant >= 2: SWIZZLE_128B_BASE32B (32-byte chunks XOR k % 4), the only
// MN-major layout 32-bit operands have (cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only
// available smem layout").
static int off_mnmajor(int K, int mn, int k, int variant) {
  const int grp = mn / 32, m = mn % 32;
  if (variant < 2) return grp * K * 128 + k * 128 + ((((m >> 2) ^ (k & 7))) << 4) + (m & 3) * 4;
  return grp * K * 128 + k * 128 + ((((m >> 3) ^ (k & 3))) << 5) + (m & 7) * 4;
}
static uint64_t desc(uint32_t lbo_bytes, uint32_t sbo_bytes, int layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

static int run_case(const char* name, int M, int N, int K, bool a_mn, bool b_mn, int variant) {
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(1234 + M + 3 * N + 7 * K + (a_mn ? 100 : 0) + (b_mn ? 1000 : 0));
  for (auto& v : A) v = (float)((rand() % 9) - 4);
  for (auto& v : B) v = (float)((rand() % 9) - 4);
  std::vector<unsigned char> ai((size_t)M * K * 4, 0), bi((size_t)N * K * 4, 0);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k)
      memcpy(&ai[a_mn ? off_mnmajor(K, m, k, variant) : off_kmajor(M, m, k)], &A[(size_t)m * K + k], 4);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k)
      memcpy(&bi[b_mn ? off_mnmajor(K, n, k, variant) : off_kmajor(N, n, k)], &B[(size_t)n * K + k], 4);
  Job job;
  memset(&job, 0, sizeof(job));
  job.N = N;
  job.a_bytes = (int)ai.size();
  job.b_bytes = (int)bi.size();
  job.nk = K / 8;
  // MN-major: LBO = distance between 32-element MN groups, SBO = distance between 8-k groups (variant 1 swaps them)
  auto mn_desc = [&](int K_) {
    switch (variant) {
      case 0: return desc((uint32_t)K_ * 128, 1024);
      case 1: return desc(1024, (uint32_t)K_ * 128);
      case 2: return desc((uint32_t)K_ * 128, 512, 1);     // BASE32B: LBO = next 32 MN elements, SBO = next 4 k rows
      default: return desc(512, (uint32_t)K_ * 128, 1);
    }
  };
  job.adesc = a_mn ? mn_desc(K) : desc(16, 1024);
  job.bdesc = b_mn ? mn_desc(K) : desc(16, 1024);
  job.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
              ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  for (int i = 0; i < job.nk; ++i) {
    const int k0 = i * 8;
    job.a_step[i] = a_mn ? k0 * 128 : (k0 / 32) * M * 128 + (k0 % 32) * 4;
    job.b_step[i] = b_mn ? k0 * 128 : (k0 / 32) * N * 128 + (k0 % 32) * 4;
  }
  unsigned char *da, *db;
  float* dout;
  CK(cudaMalloc(&da, ai.size()));
  CK(cudaMalloc(&db, bi.size()));
  CK(cudaMalloc(&dout, (size_t)M * N * 4));
  CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, (size_t)M * N * 4));
  const size_t smem = ((ai.size() + 1023) / 1024) * 1024 + bi.size() + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(da, db, job, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-60s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(3); }
  std::vector<float> out((size_t)M * N);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  double worst = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < K; ++k) ref += A[(size_t)m * K + k] * B[(size_t)n * K + k];
      const double d = fabs((double)ref - out[(size_t)m * N + n]);
      if (!(d <= 1e-3)) ++bad;
      if (d > worst || d != d) worst = d;
    }
  printf("%-60s M=%d N=%d K=%d variant=%d : %s (%d wrong, worst %.3g)\n", name, M, N, K, variant, bad ? "FAIL" : "PASS", bad, worst);
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return bad == 0;
}


// ---- K-major SWIZZLE_128B operand whose start is shifted by `shift` ROWS (shift * 128 bytes) inside a larger row-major
// buffer swizzled by the ABSOLUTE row index: the implicit-GEMM convolution reads every filter tap as a row-shifted window
// of one shared-memory activation tile.  base_offset_mode 0: descriptor base-offset field 0; 1: (shift & 7) in bits 49..51.
static int run_shift_case(int shift, int base_offset_mode, int N) {
  const int M = 128, K = 64, R = N + 64;
  std::vector<float> A((size_t)M * K), B((size_t)R * K);
  srand(99 + shift);
  for (auto& v : A) v = (float)((rand() % 9) - 4);
  for (auto& v : B) v = (float)((rand() % 9) - 4);
  std::vector<unsigned char> ai((size_t)M * K * 4, 0), bi((size_t)R * K * 4, 0);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) memcpy(&ai[off_kmajor(M, m, k)], &A[(size_t)m * K + k], 4);
  for (int n = 0; n < R; ++n)
    for (int k = 0; k < K; ++k) memcpy(&bi[off_kmajor(R, n, k)], &B[(size_t)n * K + k], 4);
  Job job;
  memset(&job, 0, sizeof(job));
  job.N = N;
  job.a_bytes = (int)ai.size();
  job.b_bytes = (int)bi.size();
  job.nk = K / 8;
  job.adesc = desc(16, 1024);
  job.bdesc = desc(16, 1024) | ((uint64_t)(base_offset_mode ? (shift & 7) : 0) << 49);
  job.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  for (int i = 0; i < job.nk; ++i) {
    const int k0 = i * 8;
    job.a_step[i] = (k0 / 32) * M * 128 + (k0 % 32) * 4;
    job.b_step[i] = (k0 / 32) * R * 128 + (k0 % 32) * 4 + shift * 128;
  }
  unsigned char *da, *db;
  float* dout;
  CK(cudaMalloc(&da, ai.size()));
  CK(cudaMalloc(&db, bi.size()));
  CK(cudaMalloc(&dout, (size_t)M * N * 4));
  CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, (size_t)M * N * 4));
  const size_t smem = ((ai.size() + 1023) / 1024) * 1024 + bi.size() + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(da, db, job, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("row-shifted start: CUDA error: %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<float> out((size_t)M * N);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < K; ++k) ref += A[(size_t)m * K + k] * B[(size_t)(n + shift) * K + k];
      if (!(fabs((double)ref - out[(size_t)m * N + n]) <= 1e-3)) ++bad;
    }
  printf("B K-major SW128, start shifted by %3d rows, base-offset field %s, N=%d : %s (%d wrong)\n", shift,
         base_offset_mode ? "(shift&7)" : "0        ", N, bad ? "FAIL" : "PASS", bad);
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return bad == 0;
}

// ---- TMEM -> register bandwidth: W warps each issue R x tcgen05.ld.32x32b.x32 (4 KB per warp instruction) ----
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) * 32 + it * 128) & 511);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j] & 1u);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 12345.f) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static void tmem_bandwidth() {
  long long* dc;
  float* ds;
  CK(cudaMalloc(&dc, 8 * 148));
  CK(cudaMalloc(&ds, 4));
  for (int warps : {4, 8, 16}) {
    const int iters = 2000;
    tmem_read_kernel<<<1, warps * 32, 0>>>(iters, dc, ds);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
    printf("TMEM read: %2d warps x %d x tcgen05.ld.32x32b.x32 (4 KB each): %lld cycles => %.1f B/clk/SM\n", warps, iters, c,
           (double)warps * iters * 4096.0 / (double)c);
  }
  cudaFree(dc); cudaFree(ds);
}

int main() {
  tmem_bandwidth();
  int ok = 0, n = 0;
  ok += run_case("A K-major, B K-major (control)", 128, 128, 64, false, false, 0); ++n;
  ok += run_case("A K-major, B K-major N=64", 128, 64, 64, false, false, 0); ++n;
  ok += run_case("A K-major, B K-major N=256", 128, 256, 32, false, false, 0); ++n;
  for (int v = 0; v < 4; ++v) {
    ok += run_case("A K-major, B MN-major N=64", 128, 64, 64, false, true, v); ++n;
    ok += run_case("A K-major, B MN-major N=256", 128, 256, 32, false, true, v); ++n;
    ok += run_case("A MN-major, B K-major N=128", 128, 128, 64, true, false, v); ++n;
    ok += run_case("A MN-major, B MN-major N=128", 128, 128, 32, true, true, v); ++n;
  }
  printf("%d / %d cases passed\n", ok, n);
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 8, 1, 3, 7, 13, 55, 57})
      for (int N : {128, 256}) run_shift_case(shift, mode, N);
  return 0;
}
