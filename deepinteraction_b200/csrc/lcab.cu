// 9x9 (k x k) local-window attention: logits = q . k over the window, softmax(logits/sqrt(C)),
// out = sum_t w_t v_t -- the fused equivalent of the reference's
//   similarFunction -> softmax -> weightingFunction
// (projects/mmdet3d_plugin/models/utils/encoder_utils.py:132-134; kernels
//  models/utils/ops/locatt_ops/kernels.cuh:4-42 `cc2k`, :44-80 `ck2c_ori`).
//
// Semantics kept from the reference kernels: out-of-image taps have logit 0 and STILL take
// softmax mass; their values are skipped.  Both fall out of zero-filling the halo.
//
// Layout: q/k/v/out are pixel-major (NHWC) fp32 with a per-pixel stride (ld*), so C channels of a
// pixel are one contiguous 4*C-byte row.  One CTA = 16x16 queries; the (16+2r)^2 halo tile of k
// (then v) streams through shared memory in 16-channel chunks with cp.async double buffering;
// each thread keeps the k*k logits of its query in registers, so logits/probabilities never
// touch HBM (the reference writes and re-reads 2*N*81*4 bytes per call).
#include "common.cuh"

namespace {

constexpr int TQ = 16;        // query tile edge
constexpr int CC = 16;        // channels per chunk
constexpr int PSTR = CC + 4;  // padded pixel stride in smem (floats): conflict-free LDS.128

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <int KS>
__global__ void __launch_bounds__(TQ * TQ, 1)
lcab_window_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                   const float* __restrict__ v, int ldv, float* __restrict__ out, int ldo, int H, int W, int C,
                   float scale) {
  constexpr int R = KS / 2;
  constexpr int TH = TQ + 2 * R;  // halo tile edge
  constexpr int NPIX = TH * TH;
  extern __shared__ __align__(16) float smem[];  // [2][NPIX][PSTR]
  const int t = threadIdx.x;
  const int tx = t % TQ, ty = t / TQ;
  const int x0 = blockIdx.x * TQ, y0 = blockIdx.y * TQ, n = blockIdx.z;
  const int qx = x0 + tx, qy = y0 + ty;
  const bool qin = qx < W && qy < H;
  const size_t img_off = (size_t)n * H * W;
  const int nchunk = C / CC;
  const int nstage = 2 * nchunk;  // k chunks then v chunks

  auto issue = [&](int stage) {
    const float* src = stage < nchunk ? k : v;
    const int ld = stage < nchunk ? ldk : ldv;
    const int c0 = (stage < nchunk ? stage : stage - nchunk) * CC;
    float* dst = smem + (stage & 1) * (NPIX * PSTR);
    for (int i = t; i < NPIX * (CC / 4); i += TQ * TQ) {
      int px = i / (CC / 4), c4 = i % (CC / 4);
      int gy = y0 - R + px / TH, gx = x0 - R + px % TH;
      bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const float* g = ok ? src + (img_off + (size_t)gy * W + gx) * ld + c0 + c4 * 4 : src;
      cp_async16(dst + px * PSTR + c4 * 4, g, ok);
    }
    cp_async_commit();
  };

  float lg[KS * KS];
#pragma unroll
  for (int i = 0; i < KS * KS; ++i) lg[i] = 0.f;

  issue(0);
  for (int stage = 0; stage < nstage; ++stage) {
    if (stage + 1 < nstage) {
      issue(stage + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* tile = smem + (stage & 1) * (NPIX * PSTR);
    if (stage < nchunk) {
      // ---- logits += q[chunk] . k[chunk] over the window
      float4 qv[CC / 4];
#pragma unroll
      for (int j = 0; j < CC / 4; ++j)
        qv[j] = qin ? ldg4(q + (img_off + (size_t)qy * W + qx) * ldq + stage * CC + j * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int dy = 0; dy < KS; ++dy)
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const float* kp = tile + ((ty + dy) * TH + tx + dx) * PSTR;
          float s = lg[dy * KS + dx];
#pragma unroll
          for (int j = 0; j < CC / 4; ++j) {
            float4 kv = *reinterpret_cast<const float4*>(kp + j * 4);
            s = fmaf(qv[j].x, kv.x, s);
            s = fmaf(qv[j].y, kv.y, s);
            s = fmaf(qv[j].z, kv.z, s);
            s = fmaf(qv[j].w, kv.w, s);
          }
          lg[dy * KS + dx] = s;
        }
      if (stage == nchunk - 1) {
        // ---- softmax over the k*k taps (out-of-image taps carry logit 0)
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < KS * KS; ++i) {
          lg[i] *= scale;
          m = fmaxf(m, lg[i]);
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < KS * KS; ++i) {
          lg[i] = expf(lg[i] - m);
          sum += lg[i];
        }
        float inv = 1.f / sum;
#pragma unroll
        for (int i = 0; i < KS * KS; ++i) lg[i] *= inv;
      }
    } else {
      // ---- out[chunk] = sum_t w_t v_t[chunk]
      float4 acc[CC / 4];
#pragma unroll
      for (int j = 0; j < CC / 4; ++j) acc[j] = make_float4(0, 0, 0, 0);
#pragma unroll
      for (int dy = 0; dy < KS; ++dy)
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const float* vp = tile + ((ty + dy) * TH + tx + dx) * PSTR;
          float w = lg[dy * KS + dx];
#pragma unroll
          for (int j = 0; j < CC / 4; ++j) {
            float4 vv = *reinterpret_cast<const float4*>(vp + j * 4);
            acc[j].x = fmaf(w, vv.x, acc[j].x);
            acc[j].y = fmaf(w, vv.y, acc[j].y);
            acc[j].z = fmaf(w, vv.z, acc[j].z);
            acc[j].w = fmaf(w, vv.w, acc[j].w);
          }
        }
      if (qin) {
        float* o = out + (img_off + (size_t)qy * W + qx) * ldo + (stage - nchunk) * CC;
#pragma unroll
        for (int j = 0; j < CC / 4; ++j) *reinterpret_cast<float4*>(o + j * 4) = acc[j];
      }
    }
    __syncthreads();  // the buffer just consumed is refilled by the next iteration's issue()
  }
}


// ------------------------------------------------------------------------------------------------
// Unfused NCHW window ops with the exact contract of the reference extension `localattention`
// (locatt_ops/kernels.cuh: cc2k :4-42, ck2c_ori :44-80, ck2c_loc :82-119; fp32 data, fp64 accumulate
// as in f_cc2k<float,double>).  They exist for drop-in compatibility (autograd of the unfused path);
// the product forward uses lcab_window_kernel above.
// ------------------------------------------------------------------------------------------------
__global__ void locatt_cc2k_kernel(const float* __restrict__ x_ori, const float* __restrict__ x_loc,
                                   float* __restrict__ y, int C, int H, int W, int kH, int kW, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over N*H*W*patch, w fastest then tap
  if (i >= total) return;
  const int patch = kH * kW, hw = H * W;
  int w = (int)(i % W);
  int t = (int)((i / W) % patch);
  int h = (int)((i / ((long long)W * patch)) % H);
  int n = (int)(i / ((long long)W * patch * H));
  int hh = h - kH / 2 + t / kW, ww = w - kW / 2 + t % kW;
  double acc = 0.0;
  if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
    const float* a = x_ori + (size_t)n * C * hw + h * W + w;
    const float* b = x_loc + (size_t)n * C * hw + hh * W + ww;
    for (int c = 0; c < C; ++c) acc += (double)(__ldg(a + (size_t)c * hw) * __ldg(b + (size_t)c * hw));
  }
  y[((size_t)n * hw + h * W + w) * patch + t] = (float)acc;
}

// y[n,c,h,w] = sum_t x_loc[n,c,h+dy,w+dx] * wgt[n,h,w,t]
__global__ void locatt_ck2c_ori_kernel(const float* __restrict__ x_loc, const float* __restrict__ wgt,
                                       float* __restrict__ y, int C, int H, int W, int kH, int kW, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over N*C*H*W
  if (i >= total) return;
  const int patch = kH * kW, hw = H * W;
  int w = (int)(i % W), h = (int)((i / W) % H);
  long long nc = i / hw;
  int n = (int)(nc / C);
  const float* src = x_loc + nc * hw;
  const float* pw = wgt + ((size_t)n * hw + h * W + w) * patch;
  double acc = 0.0;
  for (int t = 0; t < patch; ++t) {
    int hh = h - kH / 2 + t / kW, ww = w - kW / 2 + t % kW;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) acc += (double)(__ldg(src + hh * W + ww) * __ldg(pw + t));
  }
  y[i] = (float)acc;
}

// y[n,c,h,w] = sum_t x_ori[n,c,h-dy,w-dx] * wgt[n,h-dy,w-dx,t]
__global__ void locatt_ck2c_loc_kernel(const float* __restrict__ x_ori, const float* __restrict__ wgt,
                                       float* __restrict__ y, int C, int H, int W, int kH, int kW, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int patch = kH * kW, hw = H * W;
  int w = (int)(i % W), h = (int)((i / W) % H);
  long long nc = i / hw;
  int n = (int)(nc / C);
  const float* src = x_ori + nc * hw;
  const float* pw = wgt + (size_t)n * hw * patch;
  double acc = 0.0;
  for (int t = 0; t < patch; ++t) {
    int hh = h + kH / 2 - t / kW, ww = w + kW / 2 - t % kW;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W)
      acc += (double)(__ldg(src + hh * W + ww) * __ldg(pw + ((size_t)hh * W + ww) * patch + t));
  }
  y[i] = (float)acc;
}

}  // namespace

extern "C" {

// q,k,v,out: [N,H,W,*] pixel-major fp32 with per-pixel strides ldq/ldk/ldv/ldo (>= C, multiples of 4).
// ksize must be 9 (the only window the reference configs use) or 3 (tests).  C % 16 == 0.
int di_lcab_window_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out, int ldo,
                       int N, int H, int W, int C, int ksize, cudaStream_t stream) {
  DI_CHECK_ARG(q && k && v && out && N > 0 && H > 0 && W > 0, "di_lcab_window_f32: bad argument");
  DI_CHECK_ARG(C > 0 && C % CC == 0, "di_lcab_window_f32: C must be a multiple of 16 (got %d)", C);
  DI_CHECK_ARG(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0, "di_lcab_window_f32: strides must be multiples of 4");
  DI_CHECK_ARG(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "di_lcab_window_f32: pointers must be 16-byte aligned");
  dim3 grid(di_cdiv(W, TQ), di_cdiv(H, TQ), N);
  float scale = 1.0f / sqrtf((float)C);
  if (ksize == 9) {
    size_t smem = 2ull * (TQ + 8) * (TQ + 8) * PSTR * sizeof(float);
    cudaFuncSetAttribute(lcab_window_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lcab_window_kernel<9><<<grid, TQ * TQ, smem, stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, H, W, C, scale);
  } else if (ksize == 3) {
    size_t smem = 2ull * (TQ + 2) * (TQ + 2) * PSTR * sizeof(float);
    cudaFuncSetAttribute(lcab_window_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lcab_window_kernel<3><<<grid, TQ * TQ, smem, stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, H, W, C, scale);
  } else {
    di_set_error("di_lcab_window_f32: unsupported window %d", ksize);
    return DI_ERR_UNSUPPORTED;
  }
  DI_CHECK_LAUNCH("di_lcab_window_f32");
  return DI_OK;
}

// localattention.similar_forward / weighting_backward_weight  (localAttention.cpp:7-15, 51-59):
// x_ori, x_loc [N,C,H,W] -> y [N,H,W,kH*kW]
int di_locatt_cc2k_f32(const float* x_ori, const float* x_loc, float* y, int N, int C, int H, int W, int kH, int kW,
                       cudaStream_t stream) {
  DI_CHECK_ARG(x_ori && x_loc && y && N > 0 && C > 0 && H > 0 && W > 0 && kH > 0 && kW > 0, "di_locatt_cc2k_f32: bad argument");
  long long total = (long long)N * H * W * kH * kW;
  locatt_cc2k_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x_ori, x_loc, y, C, H, W, kH, kW, total);
  DI_CHECK_LAUNCH("di_locatt_cc2k_f32");
  return DI_OK;
}
// localattention.weighting_forward / similar_backward(is_ori=True)  (localAttention.cpp:31-39, 17-28)
int di_locatt_ck2c_ori_f32(const float* x_loc, const float* wgt, float* y, int N, int C, int H, int W, int kH, int kW,
                           cudaStream_t stream) {
  DI_CHECK_ARG(x_loc && wgt && y && N > 0 && C > 0 && H > 0 && W > 0, "di_locatt_ck2c_ori_f32: bad argument");
  long long total = (long long)N * C * H * W;
  locatt_ck2c_ori_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x_loc, wgt, y, C, H, W, kH, kW, total);
  DI_CHECK_LAUNCH("di_locatt_ck2c_ori_f32");
  return DI_OK;
}
// localattention.similar_backward(is_ori=False) / weighting_backward_ori  (localAttention.cpp:17-28, 41-49)
int di_locatt_ck2c_loc_f32(const float* x_ori, const float* wgt, float* y, int N, int C, int H, int W, int kH, int kW,
                           cudaStream_t stream) {
  DI_CHECK_ARG(x_ori && wgt && y && N > 0 && C > 0 && H > 0 && W > 0, "di_locatt_ck2c_loc_f32: bad argument");
  long long total = (long long)N * C * H * W;
  locatt_ck2c_loc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x_ori, wgt, y, C, H, W, kH, kW, total);
  DI_CHECK_LAUNCH("di_locatt_ck2c_loc_f32");
  return DI_OK;
}

}  // extern "C"
