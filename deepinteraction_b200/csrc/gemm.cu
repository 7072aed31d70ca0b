// Dense fp32 layers of the MMRI/MMPI path: pointwise (1x1) convolutions / linear layers
// with up to three K-concatenated sources, and 3x3 convolutions as implicit GEMM.
//
// Replaces (reference, projects/mmdet3d_plugin/):
//   models/utils/encoder_utils.py:11-34   ConvBNReLU (1x1, BN folded on the host)
//   models/necks/deepinteraction_encoder.py:26-27,31-32  cat + 1x1 Conv+BN pairs
//   models/necks/deepinteraction_encoder.py:47-62,80-81  shared 3x3 convs
//   models/dense_heads/deepinteraction_decoder.py:83-101,223-224  heatmap heads (3x3)
//   models/utils/decoder_utils.py  every nn.Linear / Conv1d(k=1) of the decoder
//
// One 128x128x8 register-tiled SGEMM core (8x8 outputs per thread, double-buffered
// shared memory) parameterised by how an A tile is fetched.  fp32 FFMA with fp32
// accumulation: results are fp32-faithful to the reference (no tf32/bf16 rounding).
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8, NT = 256, PADM = 4;

// ---------------------------------------------------------------------------------------------
// A-tile loaders.  Each thread fetches 4 elements of the BMxBK tile per k-step into registers
// (prefetch) and later writes them to shared memory in [k][m] order (commit).
// ---------------------------------------------------------------------------------------------

// Row-major sources, k contiguous; up to 3 sources concatenated along K.
struct RowsLoader {
  const float* p[3];
  int ld[3];
  int kend[3];  // cumulative K
  int M, K;
  bool vec;
  int m;   // this thread's row
  int kq;  // this thread's k offset inside the tile (0 or 4)
  __device__ void init(int m0, int t) {
    m = m0 + (t >> 1);
    kq = (t & 1) * 4;
  }
  __device__ __forceinline__ float at(int k) const {
    if (k >= K) return 0.f;
    int s = (k >= kend[0]) + (k >= kend[1]);
    int kb = s == 0 ? 0 : kend[s - 1];
    return __ldg(p[s] + (size_t)m * ld[s] + (k - kb));
  }
  __device__ __forceinline__ void prefetch(int k0, float (&r)[4]) const {
    int k = k0 + kq;
    if (m >= M) {
      r[0] = r[1] = r[2] = r[3] = 0.f;
      return;
    }
    if (vec) {
      if (k >= K) {
        r[0] = r[1] = r[2] = r[3] = 0.f;
        return;
      }
      int s = (k >= kend[0]) + (k >= kend[1]);
      int kb = s == 0 ? 0 : kend[s - 1];
      float4 v = ldg4(p[s] + (size_t)m * ld[s] + (k - kb));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = at(k + i);
    }
  }
  __device__ __forceinline__ void commit(float* As, int t, const float (&r)[4]) const {
    int row = t >> 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) As[(kq + i) * (BM + PADM) + row] = r[i];
  }
};

// 3x3 conv, NHWC input: k = tap*Cin + ci (ci contiguous).
struct ConvNHWCLoader {
  const float* x;
  int Cin, H, W, M, K;
  int n, y, xx, kq;
  bool inb;
  __device__ void init(int m0, int t) {
    int m = m0 + (t >> 1);
    kq = (t & 1) * 4;
    inb = m < M;
    int hw = H * W;
    n = m / hw;
    int r = m - n * hw;
    y = r / W;
    xx = r - y * W;
  }
  __device__ __forceinline__ void prefetch(int k0, float (&r)[4]) const {
    int k = k0 + kq;
    r[0] = r[1] = r[2] = r[3] = 0.f;
    if (!inb || k >= K) return;
    int tap = k / Cin, ci = k - tap * Cin;
    int yy = y + tap / 3 - 1, xc = xx + tap % 3 - 1;
    if (yy < 0 || yy >= H || xc < 0 || xc >= W) return;
    float4 v = ldg4(x + ((size_t)(n * H + yy) * W + xc) * Cin + ci);
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  }
  __device__ __forceinline__ void commit(float* As, int t, const float (&r)[4]) const {
    int row = t >> 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) As[(kq + i) * (BM + PADM) + row] = r[i];
  }
};

// 3x3 conv, NCHW input: pixels contiguous -> each thread fetches 4 consecutive pixels of one k.
struct ConvNCHWLoader {
  const float* x;
  int Cin, H, W, M, K;
  int n[4], y[4], xx[4];
  int kk, mg;
  __device__ void init(int m0, int t) {
    kk = t >> 5;
    mg = (t & 31) * 4;
    int hw = H * W;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = m0 + mg + i;
      if (m >= M) {
        n[i] = -1; y[i] = 0; xx[i] = 0;
      } else {
        n[i] = m / hw;
        int r = m - n[i] * hw;
        y[i] = r / W;
        xx[i] = r - y[i] * W;
      }
    }
  }
  __device__ __forceinline__ void prefetch(int k0, float (&r)[4]) const {
    int k = k0 + kk;
    r[0] = r[1] = r[2] = r[3] = 0.f;
    if (k >= K) return;
    int tap = k / Cin, ci = k - tap * Cin;
    int dy = tap / 3 - 1, dx = tap % 3 - 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int yy = y[i] + dy, xc = xx[i] + dx;
      if (n[i] >= 0 && yy >= 0 && yy < H && xc >= 0 && xc < W)
        r[i] = __ldg(x + ((size_t)(n[i] * Cin + ci) * H + yy) * W + xc);
    }
  }
  __device__ __forceinline__ void commit(float* As, int t, const float (&r)[4]) const {
    *reinterpret_cast<float4*>(&As[kk * (BM + PADM) + mg]) = make_float4(r[0], r[1], r[2], r[3]);
  }
};

// Weight tile loader: W [N][K] row-major.
struct WLoader {
  const float* w;
  int N, K;
  bool vec;
  int nrow, kq;
  __device__ void init(int n0, int t) {
    nrow = n0 + (t >> 1);
    kq = (t & 1) * 4;
  }
  __device__ __forceinline__ void prefetch(int k0, float (&r)[4]) const {
    int k = k0 + kq;
    r[0] = r[1] = r[2] = r[3] = 0.f;
    if (nrow >= N) return;
    if (vec) {
      if (k >= K) return;
      float4 v = ldg4(w + (size_t)nrow * K + k);
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (k + i < K) r[i] = __ldg(w + (size_t)nrow * K + k + i);
    }
  }
  __device__ __forceinline__ void commit(float* Bs, int t, const float (&r)[4]) const {
    int row = t >> 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[(kq + i) * (BN + PADM) + row] = r[i];
  }
};

struct Epilogue {
  float* C;
  const float* bias;
  int ldc;       // row-major: C[m*ldc + n]
  int act;
  int nchw_hw;   // >0: C is NCHW with this many pixels per image, N channels: C[(img*N + n)*hw + pix]
  size_t split_stride;  // elements between split-K partial outputs
  bool vec_store;       // C base 16-byte aligned and ldc % 4 == 0
  const float* res;     // optional residual / per-row constant: + res[(m % res_mod)*ldres + n]
  int ldres, res_mod;
};

template <class ALoader>
__global__ void __launch_bounds__(NT, 2)
sgemm_kernel(ALoader la, WLoader lb, Epilogue ep, int M, int N, int K, int ktiles_per_split) {
  __shared__ __align__(16) float As[2][BK * (BM + PADM)];
  __shared__ __align__(16) float Bs[2][BK * (BN + PADM)];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = t & 15, ty = t >> 4;
  la.init(m0, t);
  lb.init(n0, t);
  const int nk_total = (K + BK - 1) / BK;
  const int kt_begin = blockIdx.z * ktiles_per_split;
  const int kt_end = min(nk_total, kt_begin + ktiles_per_split);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  if (kt_begin < kt_end) {
    la.prefetch(kt_begin * BK, ra);
    lb.prefetch(kt_begin * BK, rb);
    la.commit(As[0], t, ra);
    lb.commit(Bs[0], t, rb);
  }
  __syncthreads();
  for (int kt = kt_begin; kt < kt_end; ++kt) {
    const int cur = (kt - kt_begin) & 1;
    const bool more = kt + 1 < kt_end;
    if (more) {
      la.prefetch((kt + 1) * BK, ra);
      lb.prefetch((kt + 1) * BK, rb);
    }
    const float* as = As[cur];
    const float* bs = Bs[cur];
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&as[kk * (BM + PADM) + ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&as[kk * (BM + PADM) + 64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&bs[kk * (BN + PADM) + tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&bs[kk * (BN + PADM) + 64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      la.commit(As[cur ^ 1], t, ra);
      lb.commit(Bs[cur ^ 1], t, rb);
    }
    __syncthreads();
  }

  float* C = ep.C + (size_t)blockIdx.z * ep.split_stride;
  const bool final_ep = gridDim.z == 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int nb = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float s = acc[i][jh * 4 + j];
        if (final_ep) {
          if (ep.bias && nb + j < N) s += __ldg(ep.bias + nb + j);
          if (ep.res && nb + j < N) s += __ldg(ep.res + (size_t)(m % ep.res_mod) * ep.ldres + nb + j);
        }
        v[j] = s;
      }
      if (final_ep) {                                  // one uniform branch per 4 values (no per-element if-conversion)
        if (ep.act == DI_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (ep.act == DI_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = di_gelu(v[j]);
        }
      }
      if (ep.nchw_hw > 0) {
        int img = m / ep.nchw_hw, pix = m - img * ep.nchw_hw;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < N) C[((size_t)img * N + nb + j) * ep.nchw_hw + pix] = v[j];
      } else if (ep.vec_store && nb + 3 < N) {
        *reinterpret_cast<float4*>(&C[(size_t)m * ep.ldc + nb]) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < N) C[(size_t)m * ep.ldc + nb + j] = v[j];
      }
    }
  }
}

template <class ALoader>
int launch(const ALoader& la, const WLoader& lb, const Epilogue& ep, int M, int N, int K, int splits,
           cudaStream_t stream, const char* name) {
  int nk = di_cdiv(K, BK);
  if (splits < 1) splits = 1;
  if (splits > nk) splits = nk;
  int per = di_cdiv(nk, splits);
  splits = di_cdiv(nk, per);
  dim3 grid(di_cdiv(M, BM), di_cdiv(N, BN), splits);
  sgemm_kernel<ALoader><<<grid, NT, 0, stream>>>(la, lb, ep, M, N, K, per);
  DI_CHECK_LAUNCH(name);
  return splits;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }


// ------------------------------------------------------------------------------------------------
// Query-level dense layers of the decoder: M = batch * num_proposals (200..600 rows), K <= 512, N <= 512.
// These are latency-bound, not throughput-bound: a 32 x 64 output tile per CTA of 128 threads (4 x 4 per thread)
// gives 14..56 CTAs, the K loop runs on 32-wide chunks with the next chunk's global loads in flight during the
// FMAs.  Plain fp32 FFMA with round-to-nearest in k order: these layers feed softmax logits and LayerNorms of a
// badly conditioned stack, so they stay exact-fp32 rather than tensor-core split products.
// ------------------------------------------------------------------------------------------------
constexpr int SBM = 32, SBN = 64, SBK = 32;

struct SmallSrc {          // scalars, not arrays: dynamically indexed kernel parameters end up in local memory
  const float *p0, *p1, *p2;
  int ld0, ld1, ld2;
  int kend0, kend1;        // cumulative K boundaries of sources 0 and 1
};

__global__ void __launch_bounds__(128)
linear_small_kernel(SmallSrc a, const float* __restrict__ W, const float* __restrict__ bias,
                    const float* __restrict__ res, int ldres, int res_mod, float* __restrict__ C, int ldc, int M,
                    int N, int K, int act) {
  __shared__ float As[SBK][SBM + 1];               // k-major: As[k][row]
  __shared__ float Ws[SBK][SBN + 1];               // Ws[k][col]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  // loader roles: A tile 32 rows x 32 k = 1024 floats -> 8 per thread (row = tid / 4, k = (tid % 4) * 8 .. +7)
  //               W tile 64 cols x 32 k = 2048 floats -> 16 per thread (col = tid / 2, k = (tid % 2) * 16 .. +15)
  const int ar = tid >> 2, ak = (tid & 3) * 8;
  const int wc = tid >> 1, wk = (tid & 1) * 16;
  float ra[8], rw[16];
  auto load = [&](int k0) {
    const int row = m0 + ar;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + ak + j;
      float v = 0.f;
      if (row < M && k < K) {
        const float* src = k < a.kend0 ? a.p0 : (k < a.kend1 ? a.p1 : a.p2);
        const int ld = k < a.kend0 ? a.ld0 : (k < a.kend1 ? a.ld1 : a.ld2);
        const int kb = k < a.kend0 ? 0 : (k < a.kend1 ? a.kend0 : a.kend1);
        v = __ldg(src + (size_t)row * ld + (k - kb));
      }
      ra[j] = v;
    }
    const int col = n0 + wc;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = k0 + wk + j;
      rw[j] = (col < N && k < K) ? __ldg(W + (size_t)col * K + k) : 0.f;
    }
  };
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  load(0);
  for (int k0 = 0; k0 < K; k0 += SBK) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[ak + j][ar] = ra[j];
#pragma unroll
    for (int j = 0; j < 16; ++j) Ws[wk + j][wc] = rw[j];
    __syncthreads();
    if (k0 + SBK < K) load(k0 + SBK);              // next chunk's loads overlap the FMAs below
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
    const float* rr = res ? res + (size_t)(row % res_mod) * ldres : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + col) : 0.f);
      if (rr) v += __ldg(rr + col);
      C[(size_t)row * ldc + col] = di_act(v, act);
    }
  }
}

}  // namespace

extern "C" {

// C[M,N] = act( [A0 | A1 | A2][M, K0+K1+K2] * W[N, K0+K1+K2]^T + bias[N] + res[(m % res_mod), N] )
// (bias, res optional; res is added before the activation; res_mod <= 0 means res_mod = M)
// splits > 1: deterministic split-K; partial s is written (no bias/act) to C + s*split_stride
// and the return value is the number of partials actually produced (>= 1); the caller reduces
// them (di_rows_finish).  Returns < 0 on error.
int di_linear_f32(const float* A0, int lda0, int K0, const float* A1, int lda1, int K1, const float* A2, int lda2,
                  int K2, const float* W, const float* bias, const float* res, int ldres, int res_mod, float* C,
                  int ldc, int M, int N, int act, int splits, long long split_stride, cudaStream_t stream) {
  DI_CHECK_ARG(A0 && W && C && M > 0 && N > 0 && K0 > 0, "di_linear_f32: null pointer or empty shape");
  DI_CHECK_ARG((K1 == 0 || A1) && (K2 == 0 || A2), "di_linear_f32: missing source");
  DI_CHECK_ARG(K2 == 0 || K1 > 0, "di_linear_f32: source 2 without source 1");
  if (M <= 640 && splits == 1 && N <= 2048) {        // query-level layers: latency-optimised small-tile kernel
    SmallSrc sa;
    sa.p0 = A0; sa.p1 = A1 ? A1 : A0; sa.p2 = A2 ? A2 : A0;
    sa.ld0 = lda0; sa.ld1 = lda1; sa.ld2 = lda2;
    sa.kend0 = K0; sa.kend1 = K0 + K1;
    dim3 grid(di_cdiv(N, SBN), di_cdiv(M, SBM));
    linear_small_kernel<<<grid, 128, 0, stream>>>(sa, W, bias, res, ldres, res_mod > 0 ? res_mod : M, C, ldc, M, N,
                                                  K0 + K1 + K2, act);
    DI_CHECK_LAUNCH("di_linear_f32(small)");
    return 1;
  }
  RowsLoader la;
  la.p[0] = A0; la.p[1] = A1 ? A1 : A0; la.p[2] = A2 ? A2 : A0;
  la.ld[0] = lda0; la.ld[1] = lda1; la.ld[2] = lda2;
  la.kend[0] = K0; la.kend[1] = K0 + K1; la.kend[2] = K0 + K1 + K2;
  int K = K0 + K1 + K2;
  if (K1 == 0) la.kend[0] = la.kend[1] = la.kend[2] = K;  // single source: never select s>0
  else if (K2 == 0) la.kend[1] = la.kend[2] = K;
  la.M = M; la.K = K;
  la.vec = (K0 % 4 == 0) && (K1 % 4 == 0) && (K2 % 4 == 0) && (lda0 % 4 == 0) && (K1 == 0 || lda1 % 4 == 0) &&
           (K2 == 0 || lda2 % 4 == 0) && aligned16(A0) && (K1 == 0 || aligned16(A1)) && (K2 == 0 || aligned16(A2));
  WLoader lb;
  lb.w = W; lb.N = N; lb.K = K; lb.vec = (K % 4 == 0) && aligned16(W);
  Epilogue ep;
  ep.C = C; ep.bias = bias; ep.ldc = ldc; ep.act = act; ep.nchw_hw = 0; ep.split_stride = (size_t)split_stride;
  ep.vec_store = aligned16(C) && (ldc % 4 == 0) && (split_stride % 4 == 0);
  ep.res = res; ep.ldres = ldres; ep.res_mod = res_mod > 0 ? res_mod : M;
  DI_CHECK_ARG(!(res && splits > 1), "di_linear_f32: residual is not supported with split-K");
  return launch(la, lb, ep, M, N, K, splits, stream, "di_linear_f32");
}

// 3x3 convolution, stride 1, zero padding 1, as implicit GEMM.
//   x: input, NCHW (x_nhwc=0) or NHWC (x_nhwc=1), fp32
//   w: [Cout][9*Cin], k index = (ky*3+kx)*Cin + ci  (host reorders the torch [Cout,Cin,3,3] weight)
//   y: output NHWC (y_nchw=0) or NCHW (y_nchw=1)
int di_conv3x3_f32(const float* x, int x_nhwc, const float* w, const float* bias, float* y, int y_nchw, int N,
                   int Cin, int H, int W, int Cout, int act, cudaStream_t stream) {
  DI_CHECK_ARG(x && w && y && N > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "di_conv3x3_f32: bad argument");
  long long Mll = (long long)N * H * W;
  DI_CHECK_ARG(Mll < (1ll << 31) && (long long)N * Cin * H * W < (1ll << 40), "di_conv3x3_f32: tensor too large");
  int M = (int)Mll, K = 9 * Cin;
  WLoader lb;
  lb.w = w; lb.N = Cout; lb.K = K; lb.vec = (K % 4 == 0) && aligned16(w);
  Epilogue ep;
  ep.C = y; ep.bias = bias; ep.ldc = Cout; ep.act = act; ep.nchw_hw = y_nchw ? H * W : 0; ep.split_stride = 0;
  ep.vec_store = !y_nchw && aligned16(y) && (Cout % 4 == 0);
  ep.res = nullptr; ep.ldres = 0; ep.res_mod = 1;
  if (x_nhwc) {
    DI_CHECK_ARG(Cin % 4 == 0 && aligned16(x), "di_conv3x3_f32: NHWC input needs Cin %% 4 == 0 and 16-byte alignment");
    ConvNHWCLoader la;
    la.x = x; la.Cin = Cin; la.H = H; la.W = W; la.M = M; la.K = K;
    int r = launch(la, lb, ep, M, Cout, K, 1, stream, "di_conv3x3_f32(nhwc)");
    return r < 0 ? r : DI_OK;
  }
  ConvNCHWLoader la;
  la.x = x; la.Cin = Cin; la.H = H; la.W = W; la.M = M; la.K = K;
  int r = launch(la, lb, ep, M, Cout, K, 1, stream, "di_conv3x3_f32(nchw)");
  return r < 0 ? r : DI_OK;
}

}  // extern "C"
