// Train-mode BatchNorm over pixel-major rows: batch statistics, normalise (+ReLU), and the backward with the statistics'
// own gradient.  Reference: models/utils/encoder_utils.py:11-34 (ConvBNReLU: conv -> nn.BatchNorm2d -> ReLU, the norm
// in training mode whenever the module is; momentum set in deepinteraction_encoder.py:52-57).  Every BatchNorm of the
// encoder follows a 1x1 convolution, so its input is a [M = N*H*W, C] row matrix here and the statistics are column
// moments.  HBM-bound element-wise / reduction work: float4 channel vectors, fixed-order partials -> deterministic.
#include "common.cuh"

namespace {

constexpr int BN_BLOCKS = 592;   // 4 x 148 row blocks

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// per row block and channel: n, mean, M2 (sum of squared deviations from the block mean; second pass re-reads the block from L1/L2)
__global__ void __launch_bounds__(256)
bn_stats_part_kernel(const float* __restrict__ y, long long M, int C, float* __restrict__ part) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= C) return;
  const long long rows = (M + gridDim.x - 1) / gridDim.x, m0 = blockIdx.x * rows, m1 = min(M, m0 + rows);
  float s = 0.f;
  for (long long m = m0; m < m1; ++m) s += y[m * C + c];
  const float n = (float)max(0LL, m1 - m0), mu = n > 0.f ? s / n : 0.f;
  float q = 0.f;
  for (long long m = m0; m < m1; ++m) {
    const float d = y[m * C + c] - mu;
    q = fmaf(d, d, q);
  }
  float* o = part + ((size_t)blockIdx.x * C + c) * 3;
  o[0] = n; o[1] = mu; o[2] = q;
}

// Chan's pairwise combination in double, fixed block order; optional running-statistics update (unbiased variance, momentum)
__global__ void bn_stats_final_kernel(const float* __restrict__ part, int nblk, int C, long long M, float* __restrict__ mean,
                                      float* __restrict__ var, float* __restrict__ run_mean, float* __restrict__ run_var,
                                      float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double n = 0.0, mu = 0.0, m2 = 0.0;
  for (int b = 0; b < nblk; ++b) {
    const float* p = part + ((size_t)b * C + c) * 3;
    const double nb = p[0];
    if (nb <= 0.0) continue;
    const double d = (double)p[1] - mu, nn = n + nb;
    mu += d * nb / nn;
    m2 += (double)p[2] + d * d * n * nb / nn;
    n = nn;
  }
  mean[c] = (float)mu;
  var[c] = (float)(m2 / (double)M);
  if (run_mean) {
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)mu;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(M > 1 ? m2 / (double)(M - 1) : m2);
  }
}

// z = act((y - mean) * rstd * gamma + beta)
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ y, long long n4, int C4, const float* __restrict__ mean, const float* __restrict__ var,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu, float* __restrict__ z) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)(i % C4) * 4;
  const float4 v = ld4(y + i * 4), mu = ld4(mean + c), va = ld4(var + c);
  const float4 g = gamma ? ld4(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f), b = beta ? ld4(beta + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 o;
  o.x = (v.x - mu.x) * (1.f / sqrtf(va.x + eps)) * g.x + b.x;
  o.y = (v.y - mu.y) * (1.f / sqrtf(va.y + eps)) * g.y + b.y;
  o.z = (v.z - mu.z) * (1.f / sqrtf(va.z + eps)) * g.z + b.z;
  o.w = (v.w - mu.w) * (1.f / sqrtf(va.w + eps)) * g.w + b.w;
  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  *reinterpret_cast<float4*>(z + i * 4) = o;
}

// per row block and channel: sum g, sum g * xhat with g = dz * [z > 0] (z == NULL: no activation), xhat = (y - mean) * rstd
__global__ void __launch_bounds__(256)
bn_bwd_part_kernel(const float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ y, long long M, int C,
                   const float* __restrict__ mean, const float* __restrict__ var, float eps, float* __restrict__ part) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= C) return;
  const long long rows = (M + gridDim.x - 1) / gridDim.x, m0 = blockIdx.x * rows, m1 = min(M, m0 + rows);
  const float mu = mean[c], rstd = 1.f / sqrtf(var[c] + eps);
  float s1 = 0.f, s2 = 0.f;
  for (long long m = m0; m < m1; ++m) {
    float g = dz[m * C + c];
    if (z && !(z[m * C + c] > 0.f)) g = 0.f;
    s1 += g;
    s2 = fmaf(g, (y[m * C + c] - mu) * rstd, s2);
  }
  part[((size_t)blockIdx.x * C + c) * 2] = s1;
  part[((size_t)blockIdx.x * C + c) * 2 + 1] = s2;
}
__global__ void bn_bwd_final_kernel(const float* __restrict__ part, int nblk, int C, float* __restrict__ s1, float* __restrict__ s2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < nblk; ++k) {
    a += (double)part[((size_t)k * C + c) * 2];
    b += (double)part[((size_t)k * C + c) * 2 + 1];
  }
  s1[c] = (float)a;     // = d beta
  s2[c] = (float)b;     // = d gamma
}

// dy = gamma * rstd * (g - s1 / M - xhat * s2 / M)
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ y, long long n4, int C4,
                    const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ gamma, float eps,
                    const float* __restrict__ s1, const float* __restrict__ s2, float inv_m, float* __restrict__ dy) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)(i % C4) * 4;
  float4 g = ld4(dz + i * 4);
  if (z) {
    const float4 zz = ld4(z + i * 4);
    g.x = zz.x > 0.f ? g.x : 0.f; g.y = zz.y > 0.f ? g.y : 0.f; g.z = zz.z > 0.f ? g.z : 0.f; g.w = zz.w > 0.f ? g.w : 0.f;
  }
  const float4 v = ld4(y + i * 4), mu = ld4(mean + c), va = ld4(var + c), a = ld4(s1 + c), b = ld4(s2 + c);
  const float4 ga = gamma ? ld4(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  float4 o;
  float r;
  r = 1.f / sqrtf(va.x + eps); o.x = ga.x * r * (g.x - a.x * inv_m - (v.x - mu.x) * r * b.x * inv_m);
  r = 1.f / sqrtf(va.y + eps); o.y = ga.y * r * (g.y - a.y * inv_m - (v.y - mu.y) * r * b.y * inv_m);
  r = 1.f / sqrtf(va.z + eps); o.z = ga.z * r * (g.z - a.z * inv_m - (v.z - mu.z) * r * b.z * inv_m);
  r = 1.f / sqrtf(va.w + eps); o.w = ga.w * r * (g.w - a.w * inv_m - (v.w - mu.w) * r * b.w * inv_m);
  *reinterpret_cast<float4*>(dy + i * 4) = o;
}

inline int bn_blocks(long long M) { return (int)(M < BN_BLOCKS ? M : BN_BLOCKS); }
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

// y [M, C] contiguous rows -> mean [C], var [C] (biased, as the normalisation uses); run_mean / run_var (optional) updated
// in place with `momentum` and the unbiased variance (torch.nn.BatchNorm2d training semantics).  work: float [592 * C * 3]
int di_bn_stats_f32(const float* y, long long M, int C, float* work, float* mean, float* var, float* run_mean, float* run_var,
                    float momentum, cudaStream_t stream) {
  DI_CHECK_ARG(y && work && mean && var && M > 0 && C > 0 && (!run_mean == !run_var), "di_bn_stats_f32: bad argument");
  const int nblk = bn_blocks(M);
  bn_stats_part_kernel<<<dim3(nblk, di_cdiv(C, 256)), 256, 0, stream>>>(y, M, C, work);
  bn_stats_final_kernel<<<di_cdiv(C, 128), 128, 0, stream>>>(work, nblk, C, M, mean, var, run_mean, run_var, momentum);
  DI_CHECK_LAUNCH("di_bn_stats_f32");
  return DI_OK;
}

// z = act((y - mean) / sqrt(var + eps) * gamma + beta); gamma / beta NULL = affine=False; z may alias y
int di_bn_apply_f32(const float* y, long long M, int C, const float* mean, const float* var, const float* gamma, const float* beta,
                    float eps, int relu, float* z, cudaStream_t stream) {
  DI_CHECK_ARG(y && mean && var && z && M > 0 && C > 0 && C % 4 == 0 && al16(y) && al16(z) && al16(mean) && al16(var) &&
                   al16(gamma) && al16(beta), "di_bn_apply_f32: bad argument");
  const long long n4 = M * (C / 4);
  bn_apply_kernel<<<(unsigned)di_cdiv(n4, 256), 256, 0, stream>>>(y, n4, C / 4, mean, var, gamma, beta, eps, relu, z);
  DI_CHECK_LAUNCH("di_bn_apply_f32");
  return DI_OK;
}

// Backward of z = act(BN_train(y)): dz [M, C], z (the saved output; NULL when there is no ReLU), y (the saved input) ->
// dy [M, C] (may alias dz), dgamma [C], dbeta [C].  work: float [592 * C * 2]
int di_bn_bwd_f32(const float* dz, const float* z, const float* y, long long M, int C, const float* mean, const float* var,
                  const float* gamma, float eps, float* work, float* dy, float* dgamma, float* dbeta, cudaStream_t stream) {
  DI_CHECK_ARG(dz && y && mean && var && work && dy && dgamma && dbeta && M > 0 && C > 0 && C % 4 == 0 && al16(dz) && al16(z) &&
                   al16(y) && al16(dy) && al16(mean) && al16(var) && al16(gamma) && al16(dgamma) && al16(dbeta),
               "di_bn_bwd_f32: bad argument");
  const int nblk = bn_blocks(M);
  bn_bwd_part_kernel<<<dim3(nblk, di_cdiv(C, 256)), 256, 0, stream>>>(dz, z, y, M, C, mean, var, eps, work);
  bn_bwd_final_kernel<<<di_cdiv(C, 128), 128, 0, stream>>>(work, nblk, C, dbeta, dgamma);
  const long long n4 = M * (C / 4);
  bn_bwd_apply_kernel<<<(unsigned)di_cdiv(n4, 256), 256, 0, stream>>>(dz, z, y, n4, C / 4, mean, var, gamma, eps, dbeta, dgamma,
                                                                     1.f / (float)M, dy);
  DI_CHECK_LAUNCH("di_bn_bwd_f32");
  return DI_OK;
}

}  // extern "C"
