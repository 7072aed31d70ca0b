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
// Tensor-core version of the 9x9 window attention (the one the product path launches).
// One CTA = 8 query rows x 16 query columns; a warp owns a 2 x 8 patch of queries (M = 16).  Both contractions run
// on mma.sync.m16n8k8 TF32 with error compensation (hi/lo split of both operands, 3 products), so logits and
// outputs stay fp32-faithful:
//   S[16 x (10 rows x 16 cols)] = Q K^T  -> 20 key blocks of 8; accumulated over 32-channel chunks
//   softmax over the 81 in-window keys of each query, entirely in registers (quad shuffles)
//   O[16 x C] = P V                       -> the S accumulator fragments are re-used as the A operand by pairing
//                                            k-index t <-> key 2t and t+4 <-> key 2t+1 of each block
// Band structure: a query uses 9 of the 16 key columns and 9 of the 10 key rows of its patch (1.98x padding).
// The key/value halo tile (16 x 24 pixels x 32 channels) and the query tile stream through shared memory with
// cp.async double buffering; zero-filled halo pixels give the reference's out-of-image rule (logit 0 kept in the
// softmax, value skipped).
// ------------------------------------------------------------------------------------------------
constexpr int MQ_ROWS = 8, MQ_COLS = 16, MCH = 32, MSTR = MCH + 4;   // padded pixel stride: conflict-free fragments
constexpr int MT_ROWS = MQ_ROWS + 8, MT_COLS = MQ_COLS + 8;
constexpr int MKV_FLOATS = MT_ROWS * MT_COLS * MSTR;                  // 13824
constexpr int MQ_FLOATS = MQ_ROWS * MQ_COLS * MSTR;                   // 4608
constexpr int MSTAGE_FLOATS = MKV_FLOATS + MQ_FLOATS;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

__global__ void __launch_bounds__(256, 1)
lcab_window_mma_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                       const float* __restrict__ v, int ldv, float* __restrict__ out, int ldo, int H, int W, int C,
                       float scale) {
  extern __shared__ __align__(16) float smem[];   // [2][MKV_FLOATS + MQ_FLOATS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int x0 = blockIdx.x * MQ_COLS, y0 = blockIdx.y * MQ_ROWS, n = blockIdx.z;
  const size_t img_off = (size_t)n * H * W;
  const int nchunk = C / MCH;
  const int nstage = 2 * nchunk;

  auto issue = [&](int stage) {
    const bool is_k = stage < nchunk;
    const float* src = is_k ? k : v;
    const int ld = is_k ? ldk : ldv;
    const int c0 = (is_k ? stage : stage - nchunk) * MCH;
    float* dst = smem + (stage & 1) * MSTAGE_FLOATS;
    for (int i = tid; i < MT_ROWS * MT_COLS * (MCH / 4); i += 256) {
      int px = i / (MCH / 4), c4 = i % (MCH / 4);
      int gy = y0 - 4 + px / MT_COLS, gx = x0 - 4 + px % MT_COLS;
      bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const float* gp = ok ? src + (img_off + (size_t)gy * W + gx) * ld + c0 + c4 * 4 : src;
      cp_async16(dst + px * MSTR + c4 * 4, gp, ok);
    }
    if (is_k) {
      float* qd = dst + MKV_FLOATS;
      for (int i = tid; i < MQ_ROWS * MQ_COLS * (MCH / 4); i += 256) {
        int px = i / (MCH / 4), c4 = i % (MCH / 4);
        int gy = y0 + px / MQ_COLS, gx = x0 + px % MQ_COLS;
        bool ok = gy < H && gx < W;
        const float* gp = ok ? q + (img_off + (size_t)gy * W + gx) * ldq + c0 + c4 * 4 : q;
        cp_async16(qd + px * MSTR + c4 * 4, gp, ok);
      }
    }
    cp_async_commit();
  };

  // warp -> 2 query rows x 8 query columns (M = 16): fragment rows g = (row wy, col wx+g), g+8 = (row wy+1, col wx+g).
  // Keys: 10 halo rows (wy .. wy+9 in tile coordinates) x 16 halo columns (wx .. wx+15) = 20 blocks of 8.
  const int wy = 2 * (warp >> 1), wx = 8 * (warp & 1);
  float S[20][4];
#pragma unroll
  for (int b = 0; b < 20; ++b) S[b][0] = S[b][1] = S[b][2] = S[b][3] = 0.f;

  issue(0);
  for (int stage = 0; stage < nstage; ++stage) {
    if (stage + 1 < nstage) {
      issue(stage + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* kv = smem + (stage & 1) * MSTAGE_FLOATS;
    if (stage < nchunk) {
      // ---------------- S += Q[chunk] K[chunk]^T ----------------
      const float* qs = kv + MKV_FLOATS + (wy * MQ_COLS + wx) * MSTR;
#pragma unroll
      for (int ks = 0; ks < MCH / 8; ++ks) {
        uint32_t ah[4], al[4];
        split_tf32(qs[g * MSTR + ks * 8 + t], ah[0], al[0]);
        split_tf32(qs[(MQ_COLS + g) * MSTR + ks * 8 + t], ah[1], al[1]);
        split_tf32(qs[g * MSTR + ks * 8 + t + 4], ah[2], al[2]);
        split_tf32(qs[(MQ_COLS + g) * MSTR + ks * 8 + t + 4], ah[3], al[3]);
#pragma unroll
        for (int r = 0; r < 10; ++r)
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
            const float* kp = kv + ((wy + r) * MT_COLS + wx + cb * 8 + g) * MSTR + ks * 8 + t;
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(kp[0], bh0, bl0);
            split_tf32(kp[4], bh1, bl1);
            mma_tf32(S[r * 2 + cb], al, bh0, bh1);
            mma_tf32(S[r * 2 + cb], ah, bl0, bl1);
            mma_tf32(S[r * 2 + cb], ah, bh0, bh1);
          }
      }
      if (stage == nchunk - 1) {
        // ---------------- masked softmax over the 81 in-window keys ----------------
        float m0 = -INFINITY, m1 = -INFINITY;   // fragment rows g (query row wy) and g+8 (query row wy+1)
#pragma unroll
        for (int b = 0; b < 20; ++b) {
          const int r = b >> 1, cb = b & 1;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int d = cb * 8 + 2 * t + j - g;               // key column - query column + 4
            const bool colok = d >= 0 && d <= 8;
            S[b][j] = (colok && r <= 8) ? S[b][j] * scale : -INFINITY;
            S[b][2 + j] = (colok && r >= 1) ? S[b][2 + j] * scale : -INFINITY;
            m0 = fmaxf(m0, S[b][j]);
            m1 = fmaxf(m1, S[b][2 + j]);
          }
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int b = 0; b < 20; ++b)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            S[b][j] = expf(S[b][j] - m0);            // exp(-inf) = 0 for out-of-window entries
            S[b][2 + j] = expf(S[b][2 + j] - m1);
            s0 += S[b][j];
            s1 += S[b][2 + j];
          }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float i0 = 1.f / s0, i1 = 1.f / s1;
#pragma unroll
        for (int b = 0; b < 20; ++b) {
          S[b][0] *= i0; S[b][1] *= i0; S[b][2] *= i1; S[b][3] *= i1;
        }
      }
    } else {
      // ---------------- O[chunk] = P V[chunk] ----------------
      float O[MCH / 8][4];
#pragma unroll
      for (int nb = 0; nb < MCH / 8; ++nb) O[nb][0] = O[nb][1] = O[nb][2] = O[nb][3] = 0.f;
#pragma unroll
      for (int r = 0; r < 10; ++r)
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const int b = r * 2 + cb;
          uint32_t ah[4], al[4];                       // A = P block with k-index t <-> key 2t, t+4 <-> key 2t+1
          split_tf32(S[b][0], ah[0], al[0]);
          split_tf32(S[b][2], ah[1], al[1]);
          split_tf32(S[b][1], ah[2], al[2]);
          split_tf32(S[b][3], ah[3], al[3]);
          const float* vp = kv + ((wy + r) * MT_COLS + wx + cb * 8 + 2 * t) * MSTR + g;
#pragma unroll
          for (int nb = 0; nb < MCH / 8; ++nb) {
            uint32_t bh0, bl0, bh1, bl1;
            split_tf32(vp[nb * 8], bh0, bl0);
            split_tf32(vp[MSTR + nb * 8], bh1, bl1);
            mma_tf32(O[nb], al, bh0, bh1);
            mma_tf32(O[nb], ah, bl0, bl1);
            mma_tf32(O[nb], ah, bh0, bh1);
          }
        }
      const int cbase = (stage - nchunk) * MCH;
      const int qx = x0 + wx + g;
      if (qx < W) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int qy = y0 + wy + half;
          if (qy < H) {
            float* o = out + (img_off + (size_t)qy * W + qx) * ldo + cbase + 2 * t;
#pragma unroll
            for (int nb = 0; nb < MCH / 8; ++nb)
              *reinterpret_cast<float2*>(o + nb * 8) = make_float2(O[nb][half * 2], O[nb][half * 2 + 1]);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// bf16-split version of the kernel above (the default): same tiling, same fragment algebra, but every operand is
// split ONCE per shared-memory stage into hi = bf16(x), mid = bf16(x - hi) (16 mantissa bits kept, product error
// ~1e-5) instead of per fragment load, and the contractions run on mma.sync.m16n8k16 bf16 (3 products per k16
// instead of 3 per k8: half the tensor instructions, ~1/4 of the splitting arithmetic).
//   K / Q chunk: cp.async lands fp32 [pixel][32 ch]; a conversion pass rewrites each channel pair IN PLACE as the
//                64-bit word (hi bf16x2, mid bf16x2)  -> B / A fragments are single LDS.64, conflict-free with a
//                pixel stride of 40 words
//   V chunk:     the k index of P V is the key, so V is re-laid as [key pair along x][channel] -> (hi, mid) words in
//                a separate buffer (pair stride 72 words)
//   P:           after the softmax the 20 S accumulator fragments ARE the A fragments of P V (k16 = the 16 halo
//                columns of one halo row); they are packed to hi / mid once and reused for every V chunk
// ------------------------------------------------------------------------------------------------
constexpr int BSTR = 40;                                             // words per pixel of the K / Q landing rows
constexpr int BKV_WORDS = MT_ROWS * MT_COLS * BSTR;                  // 15360
constexpr int BQ_WORDS = MQ_ROWS * MQ_COLS * BSTR;                   // 5120
constexpr int BSTAGE_WORDS = BKV_WORDS + BQ_WORDS;                   // 20480 (80 KB)
constexpr int VPSTR = 72;                                            // words per key pair of the packed V chunk
constexpr int VPK_WORDS = MT_ROWS * (MT_COLS / 2) * VPSTR;           // 13824 (54 KB)

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// (x0, x1) -> hi = bf16x2(x0, x1), mid = bf16x2(x0 - hi0, x1 - hi1)
__device__ __forceinline__ uint2 split_bf16x2(float x0, float x1) {
  const uint32_t h = pack_bf16x2(x0, x1);
  return make_uint2(h, pack_bf16x2(x0 - __uint_as_float(h << 16), x1 - __uint_as_float(h & 0xFFFF0000u)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 1)
lcab_window_bf16_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                        const float* __restrict__ v, int ldv, float* __restrict__ out, int ldo, int H, int W, int C,
                        float scale) {
  extern __shared__ __align__(16) float smem[];   // [2][BSTAGE_WORDS] landing / converted K,Q | [VPK_WORDS] packed V
  uint32_t* vpk = reinterpret_cast<uint32_t*>(smem + 2 * BSTAGE_WORDS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int x0 = blockIdx.x * MQ_COLS, y0 = blockIdx.y * MQ_ROWS, n = blockIdx.z;
  const size_t img_off = (size_t)n * H * W;
  const int nchunk = C / MCH;
  const int nstage = 2 * nchunk;

  auto issue = [&](int stage) {
    const bool is_k = stage < nchunk;
    const float* src = is_k ? k : v;
    const int ld = is_k ? ldk : ldv;
    const int c0 = (is_k ? stage : stage - nchunk) * MCH;
    float* dst = smem + (stage & 1) * BSTAGE_WORDS;
    for (int i = tid; i < MT_ROWS * MT_COLS * (MCH / 4); i += 256) {
      int px = i / (MCH / 4), c4 = i % (MCH / 4);
      int gy = y0 - 4 + px / MT_COLS, gx = x0 - 4 + px % MT_COLS;
      bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const float* gp = ok ? src + (img_off + (size_t)gy * W + gx) * ld + c0 + c4 * 4 : src;
      cp_async16(dst + px * BSTR + c4 * 4, gp, ok);
    }
    if (is_k) {
      float* qd = dst + BKV_WORDS;
      for (int i = tid; i < MQ_ROWS * MQ_COLS * (MCH / 4); i += 256) {
        int px = i / (MCH / 4), c4 = i % (MCH / 4);
        int gy = y0 + px / MQ_COLS, gx = x0 + px % MQ_COLS;
        bool ok = gy < H && gx < W;
        const float* gp = ok ? q + (img_off + (size_t)gy * W + gx) * ldq + c0 + c4 * 4 : q;
        cp_async16(qd + px * BSTR + c4 * 4, gp, ok);
      }
    }
    cp_async_commit();
  };

  const int wy = 2 * (warp >> 1), wx = 8 * (warp & 1);
  float S[20][4];
#pragma unroll
  for (int b = 0; b < 20; ++b) S[b][0] = S[b][1] = S[b][2] = S[b][3] = 0.f;
  uint32_t Ph[10][4], Pm[10][4];                     // P = softmax(S) as packed bf16 hi / mid A fragments

  issue(0);
  for (int stage = 0; stage < nstage; ++stage) {
    if (stage + 1 < nstage) {
      issue(stage + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float* buf = smem + (stage & 1) * BSTAGE_WORDS;
    if (stage < nchunk) {
      // ---- split K and Q in place: channel pair (2j, 2j+1) of a pixel -> (hi, mid) ----
      for (int i = tid; i < (MT_ROWS * MT_COLS + MQ_ROWS * MQ_COLS) * (MCH / 2); i += 256) {
        float* pp = buf + (i >> 4) * BSTR + (i & 15) * 2;
        const float2 x = *reinterpret_cast<const float2*>(pp);
        *reinterpret_cast<uint2*>(pp) = split_bf16x2(x.x, x.y);
      }
      __syncthreads();
      // ---------------- S += Q[chunk] K[chunk]^T ----------------
      const uint32_t* kw = reinterpret_cast<const uint32_t*>(buf);
      const uint32_t* qw = kw + BKV_WORDS + (wy * MQ_COLS + wx) * BSTR;
#pragma unroll
      for (int ks = 0; ks < MCH / 16; ++ks) {
        uint32_t ah[4], am[4];
        {
          const uint2 a0 = *reinterpret_cast<const uint2*>(qw + g * BSTR + (ks * 8 + t) * 2);
          const uint2 a1 = *reinterpret_cast<const uint2*>(qw + (MQ_COLS + g) * BSTR + (ks * 8 + t) * 2);
          const uint2 a2 = *reinterpret_cast<const uint2*>(qw + g * BSTR + (ks * 8 + t + 4) * 2);
          const uint2 a3 = *reinterpret_cast<const uint2*>(qw + (MQ_COLS + g) * BSTR + (ks * 8 + t + 4) * 2);
          ah[0] = a0.x; am[0] = a0.y; ah[1] = a1.x; am[1] = a1.y;
          ah[2] = a2.x; am[2] = a2.y; ah[3] = a3.x; am[3] = a3.y;
        }
        // 4 key blocks at a time, product-major: consecutive mma.sync never share an accumulator (the legacy
        // tensor path has a long issue-to-result latency; 3 back-to-back products into one fragment serialise on it)
#pragma unroll
        for (int r2 = 0; r2 < 5; ++r2) {
          uint2 b0[4], b1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t* kp = kw + ((wy + 2 * r2 + (u >> 1)) * MT_COLS + wx + (u & 1) * 8 + g) * BSTR + (ks * 8 + t) * 2;
            b0[u] = *reinterpret_cast<const uint2*>(kp);
            b1[u] = *reinterpret_cast<const uint2*>(kp + 8);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) mma_bf16(S[r2 * 4 + u], am, b0[u].x, b1[u].x);
#pragma unroll
          for (int u = 0; u < 4; ++u) mma_bf16(S[r2 * 4 + u], ah, b0[u].y, b1[u].y);
#pragma unroll
          for (int u = 0; u < 4; ++u) mma_bf16(S[r2 * 4 + u], ah, b0[u].x, b1[u].x);
        }
      }
      if (stage == nchunk - 1) {
        // ---------------- masked softmax over the 81 in-window keys ----------------
        float m0 = -INFINITY, m1 = -INFINITY;   // fragment rows g (query row wy) and g+8 (query row wy+1)
#pragma unroll
        for (int b = 0; b < 20; ++b) {
          const int r = b >> 1, cb = b & 1;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int d = cb * 8 + 2 * t + j - g;               // key column - query column + 4
            const bool colok = d >= 0 && d <= 8;
            S[b][j] = (colok && r <= 8) ? S[b][j] * scale : -INFINITY;
            S[b][2 + j] = (colok && r >= 1) ? S[b][2 + j] * scale : -INFINITY;
            m0 = fmaxf(m0, S[b][j]);
            m1 = fmaxf(m1, S[b][2 + j]);
          }
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int b = 0; b < 20; ++b)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            S[b][j] = expf(S[b][j] - m0);            // exp(-inf) = 0 for out-of-window entries
            S[b][2 + j] = expf(S[b][2 + j] - m1);
            s0 += S[b][j];
            s1 += S[b][2 + j];
          }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float i0 = 1.f / s0, i1 = 1.f / s1;
        // A fragments of P V for halo row r: k = the 16 halo columns wx .. wx+15 (block cb=0 -> k 0..7, cb=1 -> 8..15)
#pragma unroll
        for (int r = 0; r < 10; ++r)
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
            const uint2 lo = split_bf16x2(S[r * 2 + cb][0] * i0, S[r * 2 + cb][1] * i0);   // row g
            const uint2 hi = split_bf16x2(S[r * 2 + cb][2] * i1, S[r * 2 + cb][3] * i1);   // row g + 8
            Ph[r][cb * 2] = lo.x; Pm[r][cb * 2] = lo.y;
            Ph[r][cb * 2 + 1] = hi.x; Pm[r][cb * 2 + 1] = hi.y;
          }
      }
    } else {
      // ---- re-lay the V chunk: (key 2i, key 2i+1 of a halo row) x channel -> (hi, mid) ----
      for (int i = tid; i < MT_ROWS * (MT_COLS / 2) * MCH; i += 256) {
        const int c = i & (MCH - 1), pr = i >> 5;               // pr = row * 12 + pair
        const int px = (pr / (MT_COLS / 2)) * MT_COLS + (pr % (MT_COLS / 2)) * 2;
        const float v0 = buf[px * BSTR + c], v1 = buf[(px + 1) * BSTR + c];
        *reinterpret_cast<uint2*>(vpk + pr * VPSTR + 2 * c) = split_bf16x2(v0, v1);
      }
      __syncthreads();
      // ---------------- O[chunk] = P V[chunk] ----------------
      float O[MCH / 8][4];
#pragma unroll
      for (int nb = 0; nb < MCH / 8; ++nb) O[nb][0] = O[nb][1] = O[nb][2] = O[nb][3] = 0.f;
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        const uint32_t* vp = vpk + ((wy + r) * (MT_COLS / 2) + wx / 2 + t) * VPSTR + 2 * g;
        uint2 b0[MCH / 8], b1[MCH / 8];
#pragma unroll
        for (int nb = 0; nb < MCH / 8; ++nb) {
          b0[nb] = *reinterpret_cast<const uint2*>(vp + nb * 16);
          b1[nb] = *reinterpret_cast<const uint2*>(vp + 4 * VPSTR + nb * 16);
        }
#pragma unroll
        for (int nb = 0; nb < MCH / 8; ++nb) mma_bf16(O[nb], Pm[r], b0[nb].x, b1[nb].x);
#pragma unroll
        for (int nb = 0; nb < MCH / 8; ++nb) mma_bf16(O[nb], Ph[r], b0[nb].y, b1[nb].y);
#pragma unroll
        for (int nb = 0; nb < MCH / 8; ++nb) mma_bf16(O[nb], Ph[r], b0[nb].x, b1[nb].x);
      }
      const int cbase = (stage - nchunk) * MCH;
      const int qx = x0 + wx + g;
      if (qx < W) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int qy = y0 + wy + half;
          if (qy < H) {
            float* o = out + (img_off + (size_t)qy * W + qx) * ldo + cbase + 2 * t;
#pragma unroll
            for (int nb = 0; nb < MCH / 8; ++nb)
              *reinterpret_cast<float2*>(o + nb * 8) = make_float2(O[nb][half * 2], O[nb][half * 2 + 1]);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Same kernel for PRE-SPLIT operands: q, k (kind 1) and v (kind 2) were written by the producing dense layer's
// epilogue already as bf16 (hi, mid) words (gemm_tc.cu split_block), 4 bytes per value at the value's own position.
// The two conversion passes, one barrier per stage and the packed-V buffer disappear: stages go straight from
// cp.async to fragments (LDS.64 for Q/K, ldmatrix.x4.trans for V).
// ------------------------------------------------------------------------------------------------
constexpr int VSTR2 = 36;                                            // V chunk pixel stride: 8 ldmatrix rows -> 32 distinct banks

__global__ void __launch_bounds__(256, 1)
lcab_window_pre_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, int ldk,
                        const float* __restrict__ v, int ldv, float* __restrict__ out, int ldo, int H, int W, int C,
                        float scale) {
  extern __shared__ __align__(16) float smem[];   // [2][BSTAGE_WORDS]: K,Q chunk (pixel stride 40 words) or V chunk (36)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int x0 = blockIdx.x * MQ_COLS, y0 = blockIdx.y * MQ_ROWS, n = blockIdx.z;
  const size_t img_off = (size_t)n * H * W;
  const int nchunk = C / MCH;
  const int nstage = 2 * nchunk;

  auto issue = [&](int stage) {
    const bool is_k = stage < nchunk;
    const float* src = is_k ? k : v;
    const int ld = is_k ? ldk : ldv;
    const int c0 = (is_k ? stage : stage - nchunk) * MCH;
    float* dst = smem + (stage & 1) * BSTAGE_WORDS;
    const int pstr = is_k ? BSTR : VSTR2;
    for (int i = tid; i < MT_ROWS * MT_COLS * (MCH / 4); i += 256) {
      int px = i / (MCH / 4), c4 = i % (MCH / 4);
      int gy = y0 - 4 + px / MT_COLS, gx = x0 - 4 + px % MT_COLS;
      bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const float* gp = ok ? src + (img_off + (size_t)gy * W + gx) * ld + c0 + c4 * 4 : src;
      cp_async16(dst + px * pstr + c4 * 4, gp, ok);
    }
    if (is_k) {
      float* qd = dst + BKV_WORDS;
      for (int i = tid; i < MQ_ROWS * MQ_COLS * (MCH / 4); i += 256) {
        int px = i / (MCH / 4), c4 = i % (MCH / 4);
        int gy = y0 + px / MQ_COLS, gx = x0 + px % MQ_COLS;
        bool ok = gy < H && gx < W;
        const float* gp = ok ? q + (img_off + (size_t)gy * W + gx) * ldq + c0 + c4 * 4 : q;
        cp_async16(qd + px * BSTR + c4 * 4, gp, ok);
      }
    }
    cp_async_commit();
  };

  const int wy = 2 * (warp >> 1), wx = 8 * (warp & 1);
  float S[20][4];
#pragma unroll
  for (int b = 0; b < 20; ++b) S[b][0] = S[b][1] = S[b][2] = S[b][3] = 0.f;
  uint32_t Ph[10][4], Pm[10][4];                     // P = softmax(S) as packed bf16 hi / mid A fragments

  issue(0);
  for (int stage = 0; stage < nstage; ++stage) {
    if (stage + 1 < nstage) {
      issue(stage + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float* buf = smem + (stage & 1) * BSTAGE_WORDS;
    if (stage < nchunk) {
      // ---------------- S += Q[chunk] K[chunk]^T ----------------
      const uint32_t* kw = reinterpret_cast<const uint32_t*>(buf);
      const uint32_t* qw = kw + BKV_WORDS + (wy * MQ_COLS + wx) * BSTR;
#pragma unroll
      for (int ks = 0; ks < MCH / 16; ++ks) {
        uint32_t ah[4], am[4];
        {
          const uint2 a0 = *reinterpret_cast<const uint2*>(qw + g * BSTR + (ks * 8 + t) * 2);
          const uint2 a1 = *reinterpret_cast<const uint2*>(qw + (MQ_COLS + g) * BSTR + (ks * 8 + t) * 2);
          const uint2 a2 = *reinterpret_cast<const uint2*>(qw + g * BSTR + (ks * 8 + t + 4) * 2);
          const uint2 a3 = *reinterpret_cast<const uint2*>(qw + (MQ_COLS + g) * BSTR + (ks * 8 + t + 4) * 2);
          ah[0] = a0.x; am[0] = a0.y; ah[1] = a1.x; am[1] = a1.y;
          ah[2] = a2.x; am[2] = a2.y; ah[3] = a3.x; am[3] = a3.y;
        }
        // 4 key blocks at a time, product-major: consecutive mma.sync never share an accumulator (the legacy
        // tensor path has a long issue-to-result latency; 3 back-to-back products into one fragment serialise on it)
#pragma unroll
        for (int r2 = 0; r2 < 5; ++r2) {
          uint2 b0[4], b1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t* kp = kw + ((wy + 2 * r2 + (u >> 1)) * MT_COLS + wx + (u & 1) * 8 + g) * BSTR + (ks * 8 + t) * 2;
            b0[u] = *reinterpret_cast<const uint2*>(kp);
            b1[u] = *reinterpret_cast<const uint2*>(kp + 8);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) mma_bf16(S[r2 * 4 + u], am, b0[u].x, b1[u].x);
#pragma unroll
          for (int u = 0; u < 4; ++u) mma_bf16(S[r2 * 4 + u], ah, b0[u].y, b1[u].y);
#pragma unroll
          for (int u = 0; u < 4; ++u) mma_bf16(S[r2 * 4 + u], ah, b0[u].x, b1[u].x);
        }
      }
      if (stage == nchunk - 1) {
        // ---------------- masked softmax over the 81 in-window keys ----------------
        float m0 = -INFINITY, m1 = -INFINITY;   // fragment rows g (query row wy) and g+8 (query row wy+1)
#pragma unroll
        for (int b = 0; b < 20; ++b) {
          const int r = b >> 1, cb = b & 1;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int d = cb * 8 + 2 * t + j - g;               // key column - query column + 4
            const bool colok = d >= 0 && d <= 8;
            S[b][j] = (colok && r <= 8) ? S[b][j] * scale : -INFINITY;
            S[b][2 + j] = (colok && r >= 1) ? S[b][2 + j] * scale : -INFINITY;
            m0 = fmaxf(m0, S[b][j]);
            m1 = fmaxf(m1, S[b][2 + j]);
          }
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int b = 0; b < 20; ++b)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            S[b][j] = expf(S[b][j] - m0);            // exp(-inf) = 0 for out-of-window entries
            S[b][2 + j] = expf(S[b][2 + j] - m1);
            s0 += S[b][j];
            s1 += S[b][2 + j];
          }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float i0 = 1.f / s0, i1 = 1.f / s1;
        // A fragments of P V for halo row r: k = the 16 halo columns wx .. wx+15 (block cb=0 -> k 0..7, cb=1 -> 8..15)
#pragma unroll
        for (int r = 0; r < 10; ++r)
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) {
            const uint2 lo = split_bf16x2(S[r * 2 + cb][0] * i0, S[r * 2 + cb][1] * i0);   // row g
            const uint2 hi = split_bf16x2(S[r * 2 + cb][2] * i1, S[r * 2 + cb][3] * i1);   // row g + 8
            Ph[r][cb * 2] = lo.x; Pm[r][cb * 2] = lo.y;
            Ph[r][cb * 2 + 1] = hi.x; Pm[r][cb * 2 + 1] = hi.y;
          }
      }
    } else {
      // ---------------- O[chunk] = P V[chunk] ----------------
      // V arrives pre-split in 8-channel groups [4 hi words | 4 mid words]; ldmatrix.trans turns the [key][channel]
      // rows into the (k = key pair, n = channel) B fragments: matrices 0/1 = keys 0-7 / 8-15 hi, 2/3 = the mid parts.
      float O[MCH / 8][4];
#pragma unroll
      for (int nb = 0; nb < MCH / 8; ++nb) O[nb][0] = O[nb][1] = O[nb][2] = O[nb][3] = 0.f;
      const uint32_t vbase = (uint32_t)__cvta_generic_to_shared(buf) +
                             (uint32_t)(((wy * MT_COLS + wx + (lane & 7) + ((lane >> 3) & 1) * 8) * VSTR2 + (lane >> 4) * 4) * 4);
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        uint32_t bh0[MCH / 8], bh1[MCH / 8], bm0[MCH / 8], bm1[MCH / 8];
#pragma unroll
        for (int nb = 0; nb < MCH / 8; ++nb)
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(bh0[nb]), "=r"(bh1[nb]), "=r"(bm0[nb]), "=r"(bm1[nb])
                       : "r"(vbase + (uint32_t)((r * MT_COLS * VSTR2 + nb * 8) * 4)));
#pragma unroll
        for (int nb = 0; nb < MCH / 8; ++nb) mma_bf16(O[nb], Pm[r], bh0[nb], bh1[nb]);
#pragma unroll
        for (int nb = 0; nb < MCH / 8; ++nb) mma_bf16(O[nb], Ph[r], bm0[nb], bm1[nb]);
#pragma unroll
        for (int nb = 0; nb < MCH / 8; ++nb) mma_bf16(O[nb], Ph[r], bh0[nb], bh1[nb]);
      }
      const int cbase = (stage - nchunk) * MCH;
      const int qx = x0 + wx + g;
      if (qx < W) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int qy = y0 + wy + half;
          if (qy < H) {
            float* o = out + (img_off + (size_t)qy * W + qx) * ldo + cbase + 2 * t;
#pragma unroll
            for (int nb = 0; nb < MCH / 8; ++nb)
              *reinterpret_cast<float2*>(o + nb * 8) = make_float2(O[nb][half * 2], O[nb][half * 2 + 1]);
          }
        }
      }
    }
    __syncthreads();
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

static int g_force_ffma_window = 0;   // 1: FFMA kernel, 2: 3xTF32 mma.sync kernel, 0: bf16-split mma.sync kernel

extern "C" {

// test hook: 0 = bf16-split tensor-core kernel (default), 1 = FFMA kernel, 2 = 3xTF32 tensor-core kernel
int di_set_window_ffma(int on) {
  g_force_ffma_window = on;
  return DI_OK;
}

// di_lcab_window_f32 for operands that the producing layers emitted pre-split (di_linear_tcb_split_f32: q, k kind 1,
// v kind 2).  9x9 window, C % 32 == 0.
int di_lcab_window_pre_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out, int ldo,
                           int N, int H, int W, int C, cudaStream_t stream) {
  DI_CHECK_ARG(q && k && v && out && N > 0 && H > 0 && W > 0, "di_lcab_window_pre_f32: bad argument");
  DI_CHECK_ARG(C > 0 && C % MCH == 0, "di_lcab_window_pre_f32: C must be a multiple of 32 (got %d)", C);
  DI_CHECK_ARG(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0, "di_lcab_window_pre_f32: strides must be multiples of 4");
  DI_CHECK_ARG(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "di_lcab_window_pre_f32: pointers must be 16-byte aligned");
  dim3 mgrid(di_cdiv(W, MQ_COLS), di_cdiv(H, MQ_ROWS), N);
  size_t smem = 2ull * BSTAGE_WORDS * sizeof(float);
  static DiSmemOnce once{};
  if (!di_smem_once(once, lcab_window_pre_kernel, (int)smem)) {
    di_set_error("di_lcab_window_pre_f32: cannot reserve shared memory");
    return DI_ERR_LAUNCH;
  }
  lcab_window_pre_kernel<<<mgrid, 256, smem, stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, H, W, C, 1.0f / sqrtf((float)C));
  DI_CHECK_LAUNCH("di_lcab_window_pre_f32");
  return DI_OK;
}

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
  if (ksize == 9 && C % MCH == 0 && g_force_ffma_window == 0) {
    dim3 mgrid(di_cdiv(W, MQ_COLS), di_cdiv(H, MQ_ROWS), N);
    size_t smem = (2ull * BSTAGE_WORDS + VPK_WORDS) * sizeof(float);
    static DiSmemOnce once_bf{};
    if (!di_smem_once(once_bf, lcab_window_bf16_kernel, (int)smem)) {
      di_set_error("di_lcab_window_f32: cannot reserve shared memory");
      return DI_ERR_LAUNCH;
    }
    lcab_window_bf16_kernel<<<mgrid, 256, smem, stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, H, W, C, scale);
  } else if (ksize == 9 && C % MCH == 0 && g_force_ffma_window == 2) {
    dim3 mgrid(di_cdiv(W, MQ_COLS), di_cdiv(H, MQ_ROWS), N);
    size_t smem = 2ull * MSTAGE_FLOATS * sizeof(float);
    static DiSmemOnce once_mma{};
    if (!di_smem_once(once_mma, lcab_window_mma_kernel, (int)smem)) {
      di_set_error("di_lcab_window_f32: cannot reserve shared memory");
      return DI_ERR_LAUNCH;
    }
    lcab_window_mma_kernel<<<mgrid, 256, smem, stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, H, W, C, scale);
  } else if (ksize == 9) {
    size_t smem = 2ull * (TQ + 8) * (TQ + 8) * PSTR * sizeof(float);
    static DiSmemOnce once9{};
    if (!di_smem_once(once9, lcab_window_kernel<9>, (int)smem)) {
      di_set_error("di_lcab_window_f32: cannot reserve shared memory");
      return DI_ERR_LAUNCH;
    }
    lcab_window_kernel<9><<<grid, TQ * TQ, smem, stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, H, W, C, scale);
  } else if (ksize == 3) {
    size_t smem = 2ull * (TQ + 2) * (TQ + 2) * PSTR * sizeof(float);
    static DiSmemOnce once3{};
    if (!di_smem_once(once3, lcab_window_kernel<3>, (int)smem)) {
      di_set_error("di_lcab_window_f32: cannot reserve shared memory");
      return DI_ERR_LAUNCH;
    }
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
