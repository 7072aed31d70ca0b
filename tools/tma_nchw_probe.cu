// Stand-alone probe: can TMA deliver a CHANNEL-MAJOR box [ke channels][8 rows][16 pixels] of an NCHW fp32 tensor (no
// swizzle), with out-of-image zero fill, for the conv A operand (last paragraph)?
// usage: tma_nchw_probe <variant>   0: {16,8,64,1} no promotion   1: {16,8,64,1} L2 promotion 128B (the config that
// faulted inside the GEMM kernel)   2: {16,8,32,1}   3: rank-3 map {W,H,C*N}, box {16,8,64}   (0-3 start at x = -1)
//   4: {24,8,64,1} starting at x = -4 (16-byte aligned start)   5: {16,8,64,1} starting at x = 0
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map, int rank, int x0, int y0, int c0, int n0, int bytes, float* out,
      int* status) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  float* sm = reinterpret_cast<float*>(raw + (base - smem_u32(raw)));
  const uint32_t bar = base + 65536;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)bytes) : "memory");
    if (rank == 4)
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
              base),
          "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar), "r"(x0), "r"(y0), "r"(c0), "r"(n0)
          : "memory");
    else
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(base),
          "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar), "r"(x0), "r"(y0), "r"(c0)
          : "memory");
  }
  bool done = false;
  for (int spin = 0; spin < 4000000 && !done; ++spin) {
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
  for (int i = threadIdx.x; i < bytes / 4; i += 128) out[i] = sm[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int N = 2, C = 128, H = 37, W = 44;
  std::vector<float> x((size_t)N * C * H * W);
  for (size_t i = 0; i < x.size(); ++i) x[i] = (float)(i % 9973) * 0.25f;
  float *dx, *dout;
  int* ds;
  cudaMalloc(&dx, x.size() * 4);
  cudaMalloc(&dout, 65536);
  cudaMalloc(&ds, 4);
  cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(ds, 0, 4);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  CUtensorMap map;
  const int ke = variant == 2 ? 32 : 64;
  const int bx = variant == 4 ? 24 : 16;
  CUresult r;
  int rank = 4;
  if (variant == 3) {
    rank = 3;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C * N};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {16, 8, (cuuint32_t)ke}, es[3] = {1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)bx, 8, (cuuint32_t)ke, 1}, es[4] = {1, 1, 1, 1};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, variant == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    printf("variant %d: cuTensorMapEncodeTiled failed (%d)\n", variant, (int)r);
    return 2;
  }
  const int bytes = bx * 8 * ke * 4;
  const int x0 = variant == 4 ? -4 : (variant == 5 ? 0 : -1), y0 = 31, c0 = 64, n0 = 1;   // padding left / below
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  probe<<<1, 128, 70000>>>(map, rank, x0, y0, variant == 3 ? n0 * C + c0 : c0, n0, bytes, dout, ds);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e));
    return 1;
  }
  int st = 0;
  cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
  std::vector<float> out(bytes / 4);
  cudaMemcpy(out.data(), dout, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < ke; ++c)
    for (int yy = 0; yy < 8; ++yy)
      for (int xx = 0; xx < bx; ++xx) {
        const int gx = x0 + xx, gy = y0 + yy;
        const float ref = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? x[(((size_t)n0 * C + c0 + c) * H + gy) * W + gx] : 0.f;
        bad += out[(c * 8 + yy) * bx + xx] != ref;
      }
  printf("variant %d: status %d, %d of %d elements differ from [c][y][x] with zero fill -> %s\n", variant, st, bad, ke * 8 * bx,
         (st == 0 && bad == 0) ? "OK" : "FAIL");
  return 0;
}
