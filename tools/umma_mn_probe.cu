// Stand-alone probe (not part of libdi_b200): which LBO / SBO convention does a tcgen05.mma B operand in MN-MAJOR,
// 128B-swizzled layout use?  Needed for the planned tcgen05 window-attention kernel, where the
// V halo tile arrives as [key][64 channels] rows, i.e. N (channel) contiguous and K (key) strided.
//
//   D[128 x 128] (fp32, TMEM) = A[128 x 64] (bf16, K-major SW128, smem) * B[64 x 128] (bf16, MN-major SW128, smem)
//
// B image in shared memory: 2 atoms along N (64 channels = 128 B each), each atom column = 64 key rows of 128 B
// (what a TMA box {64 ch, 64 keys} with SWIZZLE_128B writes): offset(k, n) = (n / 64) * 8192 + k * 128 +
// (((n % 64) / 8 ^ (k & 7)) << 4) + (n % 8) * 2.  K advances by 16 keys = 2048 B per instruction.
// Candidates for (LBO, SBO): (8192, 1024) and (1024, 8192); the program prints the max error of each against the
// host product.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_mn_probe tools/umma_mn_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, float* __restrict__ d_out,
             uint32_t lbo, uint32_t sbo, int* status) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  uint8_t* a_s = sm;                 // 128 rows x 128 B = 16 KB
  uint8_t* b_s = sm + 16384;         // 2 x 64 rows x 128 B = 16 KB
  const uint32_t bar = base + 32768, slot = base + 32768 + 16;
  for (int i = threadIdx.x; i < 16384 / 16; i += 128) {
    reinterpret_cast<uint4*>(a_s)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    reinterpret_cast<uint4*>(b_s)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + 32768 + 16);
  if (threadIdx.x == 0) {
    // kind::f16, bf16 x bf16 -> fp32, A K-major, B MN-major (bit 16), M = 128, N = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a_addr = base + ks * 32, b_addr = base + 16384 + ks * 2048;
      const uint64_t da = (uint64_t)((a_addr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
      const uint64_t db = (uint64_t)((b_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
                          (1ull << 46) | (2ull << 61);
      const uint32_t acc = ks != 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
  // bounded wait: never hang the box
  bool done = false;
  for (int spin = 0; spin < 2000000 && !done; ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar)
                 : "memory");
    done = ok != 0;
  }
  if (!done) {
    if (threadIdx.x == 0) *status = 1;
    return;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

int main() {
  const int M = 128, N = 128, K = 64;
  std::vector<float> A(M * K), B(K * N), D(M * N, 0.f);
  srand(7);
  for (auto& x : A) x = bf2f(f2bf((rand() % 2001 - 1000) / 500.f));
  for (auto& x : B) x = bf2f(f2bf((rand() % 2001 - 1000) / 500.f));
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[k * N + n];
      D[m * N + n] = (float)s;
    }
  std::vector<uint8_t> a_img(16384, 0), b_img(16384, 0);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      size_t off = (size_t)m * 128 + (((k / 8) ^ (m & 7)) << 4) + (k % 8) * 2;
      uint16_t v = f2bf(A[m * K + k]);
      memcpy(&a_img[off], &v, 2);
    }
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) {
      size_t off = (size_t)(n / 64) * 8192 + (size_t)k * 128 + ((((n % 64) / 8) ^ (k & 7)) << 4) + (n % 8) * 2;
      uint16_t v = f2bf(B[k * N + n]);
      memcpy(&b_img[off], &v, 2);
    }
  uint8_t *da, *db;
  float* dd;
  int* ds;
  cudaMalloc(&da, 16384);
  cudaMalloc(&db, 16384);
  cudaMalloc(&dd, M * N * 4);
  cudaMalloc(&ds, 4);
  cudaMemcpy(da, a_img.data(), 16384, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img.data(), 16384, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const uint32_t cand[4][2] = {{8192, 1024}, {1024, 8192}, {8192, 2048}, {2048, 8192}};
  for (int c = 0; c < 4; ++c) {
    cudaMemset(dd, 0, M * N * 4);
    cudaMemset(ds, 0, 4);
    probe_kernel<<<1, 128, 40000>>>(da, db, dd, cand[c][0], cand[c][1], ds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("candidate LBO=%u SBO=%u: CUDA error %s\n", cand[c][0], cand[c][1], cudaGetErrorString(e));
      return 1;
    }
    int st = 0;
    std::vector<float> out(M * N);
    cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
    double worst = 0, ref = 0;
    for (int i = 0; i < M * N; ++i) {
      worst = fmax(worst, fabs((double)out[i] - D[i]));
      ref = fmax(ref, fabs((double)D[i]));
    }
    printf("candidate LBO=%5u SBO=%5u: status %d  max|err| %.3e  (max|ref| %.3e)  %s\n", cand[c][0], cand[c][1], st, worst,
           ref, worst < 1e-3 * ref ? "MATCH" : "mismatch");
  }
  return 0;
}
