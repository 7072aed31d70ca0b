// Query x BEV cross attention of the MMPI decoder on the Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
// Reference: TransformerDecoderLayer / multi_head_attention_forward, models/utils/decoder_utils.py:101-103, 466-493:
// P = 200 object queries x HW = 32 400 BEV keys, 8 heads of 16 channels, softmax over all keys.  The K / V projection
// (with the folded key positional term) stays a 3xTF32 GEMM; di_attn_planes_f32 then writes the operands as bf16 PLANES:
// K and the pre-scaled queries as hi | mid | lo (x = hi + mid + lo to 24 mantissa bits: the logits go through exp, and a
// two-plane product -- 3 * 2^-18 relative -- moved the decoder outputs by up to 3e-3), V as hi | mid.  A TMA box
// {64 channels, rows} of a plane is a 128B-swizzled tcgen05 operand tile whose 32-byte k-slices are the heads.
//
// CTA = (batch, query tile of 128, channel box of 64 = 4 heads, key split).  Per key tile of 64 keys:
//   S_h[128 q x 64 keys] = Q_h K_h^T    h = 0..3: one k-step (16 channels) per head, SS-mode tcgen05.mma kind::f16,
//                                       six-term split product (lo hi, hi lo, mid mid, mid hi, hi mid, hi hi: error
//                                       ~2^-26) -> TMEM columns [64 h, 64 h + 64)
//   P_h = exp(S_h - m)                  8 softmax warps (two per TMEM lane quarter, two heads each), thread = query:
//                                       running maximum m and sum l per head in registers (online softmax), P written
//                                       IN PLACE over S as bf16 hi | mid words (the A operand of the second product)
//   O_h[128 q x 64 ch] = P_h V          A = P_h from tensor memory, B = the V tile [64 keys x 64 ch] as an MN-major
//                                       128B-swizzled operand; only channels [16 h, 16 h + 16) of O_h belong to head h
//                                       (the other 48 columns are the price of a descriptor convention that is pinned)
//   acc_h = acc_h * corr + O_h          the softmax threads read their head's 16 columns and keep the running
//                                       numerator in registers, so O never needs a rescale in tensor memory
// At the end every CTA writes (m, l, acc[16]) per (query, head) -- flash-decoding partials -- and
// xattn_combine_kernel merges the key splits (one warp per (batch, head, query), lanes over the splits).
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int XT_KEYS = 64;                          // keys per tile
constexpr int XT_BOX = XT_KEYS * 128;                // one K / V box: 64 rows x 128 B
constexpr int XT_STAGE = 5 * XT_BOX;                 // K hi | K mid | K lo | V hi | V mid
constexpr int XT_RING = 4;
constexpr int XT_QBOX = 128 * 128;                   // 128 queries x 128 B
constexpr int XT_Q_OFF = 0, XT_RING_OFF = 3 * XT_QBOX, XT_BAR_OFF = XT_RING_OFF + XT_RING * XT_STAGE;
constexpr int XT_SMEM = XT_BAR_OFF + 256;
constexpr int XT_THREADS = 10 * 32;                  // producer, MMA issuer, 8 softmax warps
static_assert(XT_SMEM <= 232448, "xattn kernel exceeds the shared-memory limit");

// kind::f16: bf16 x bf16 -> fp32, M = 128, N = 64.  S: A and B K-major.  PV: B MN-major (bit 16).
constexpr uint32_t XIDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t XIDESC_PV = XIDESC_S | (1u << 16);

__device__ __forceinline__ void x_umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
          tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t x_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ float x_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct XtParams {
  int P, HW, heads, nsplit, chunk;       // chunk: keys per split (multiple of 64)
  float* part;                           // [B, heads, nsplit, 18, P]
};

__global__ void __launch_bounds__(XT_THREADS, 1)
xattn_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                const __grid_constant__ CUtensorMap mapV, const XtParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t bars = base + XT_BAR_OFF;
  const uint32_t q_full = bars, s_full = bars + 8, p_full = bars + 16, o_full = bars + 24, o_empty = bars + 32;
  auto full = [&](int s) { return bars + 40u + 8u * s; };
  auto empty = [&](int s) { return bars + 40u + 8u * XT_RING + 8u * s; };
  const uint32_t tmem_slot = bars + 40u + 16u * XT_RING;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + XT_BAR_OFF + 40 + 16 * XT_RING);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sp = blockIdx.x, mt = blockIdx.y >> 1, box = blockIdx.y & 1, b = blockIdx.z;
  const int k_begin = sp * p.chunk, k_end = min(p.HW, k_begin + p.chunk);
  const int ntiles = (k_end - k_begin + XT_KEYS - 1) / XT_KEYS;          // >= 1 (the host sizes the splits)

  if (warp == 0 && lane < 3) {
    const CUtensorMap* mp = lane == 0 ? &mapQ : lane == 1 ? &mapK : &mapV;
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(mp)) : "memory");
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 8);
    for (int s = 0; s < XT_RING; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      const int q0 = b * p.P + mt * 128;
      mbar_expect_tx(q_full, 3 * XT_QBOX);
#pragma unroll
      for (int pl = 0; pl < 3; ++pl)                                                 // hi, mid, lo planes
        tma_load_2d(base + XT_Q_OFF + pl * XT_QBOX, &mapQ, q_full, 128 * pl + 64 * box, q0);
      for (int i = 0; i < ntiles; ++i) {
        const int s = i % XT_RING;
        if (i >= XT_RING) mbar_wait(empty(s), ((i / XT_RING) - 1) & 1);
        const uint32_t st = base + XT_RING_OFF + s * XT_STAGE;
        const int row = b * p.HW + k_begin + i * XT_KEYS;
        mbar_expect_tx(full(s), XT_STAGE);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) tma_load_2d(st + pl * XT_BOX, &mapK, full(s), 128 * pl + 64 * box, row);
        tma_load_2d(st + 3 * XT_BOX, &mapV, full(s), 64 * box, row);
        tma_load_2d(st + 4 * XT_BOX, &mapV, full(s), 128 + 64 * box, row);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const uint32_t qb = base + XT_Q_OFF;
    mbar_wait(q_full, 0);
    for (int i = 0; i < ntiles; ++i) {
      const int s = i % XT_RING;
      mbar_wait(full(s), (i / XT_RING) & 1);
      tc_fence_after();
      const uint32_t kb = base + XT_RING_OFF + s * XT_STAGE;
      if (elect_one()) {
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          const uint64_t q_hi = umma_desc(qb + hh * 32), q_mid = umma_desc(qb + XT_QBOX + hh * 32),
                         q_lo = umma_desc(qb + 2 * XT_QBOX + hh * 32);
          const uint64_t k_hi = umma_desc(kb + hh * 32), k_mid = umma_desc(kb + XT_BOX + hh * 32),
                         k_lo = umma_desc(kb + 2 * XT_BOX + hh * 32);
          const uint32_t d = tmem_base + (uint32_t)(64 * hh);
          x_umma_ss(d, q_lo, k_hi, XIDESC_S, 0);              // smallest terms first
          x_umma_ss(d, q_hi, k_lo, XIDESC_S, 1);
          x_umma_ss(d, q_mid, k_mid, XIDESC_S, 1);
          x_umma_ss(d, q_mid, k_hi, XIDESC_S, 1);
          x_umma_ss(d, q_hi, k_mid, XIDESC_S, 1);
          x_umma_ss(d, q_hi, k_hi, XIDESC_S, 1);
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, i & 1);                             // P of this tile is in tensor memory
      if (i > 0) mbar_wait(o_empty, (i - 1) & 1);           // the previous tile's O columns have been read
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vb = kb + 3 * XT_BOX;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          const uint32_t d = tmem_base + 256u + (uint32_t)(64 * hh);
#pragma unroll
          for (int t = 0; t < XT_KEYS / 16; ++t) {
            const uint64_t v_hi = x_desc_mn(vb + t * 2048, XT_BOX), v_mid = x_desc_mn(vb + XT_BOX + t * 2048, XT_BOX);
            const uint32_t p_hi = tmem_base + (uint32_t)(64 * hh + 16 * t), p_mid = p_hi + 8u;
            umma_bf16_ts(d, p_mid, v_hi, XIDESC_PV, t != 0);
            umma_bf16_ts(d, p_hi, v_mid, XIDESC_PV, 1);
            umma_bf16_ts(d, p_hi, v_hi, XIDESC_PV, 1);
          }
        }
        umma_commit(empty(s));
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ---------------- softmax / accumulation (warps 2..9): thread = query = TMEM lane, two heads per warp ----------------
    const int qd = warp & 3, sub = (warp - 2) >> 2;
    const int qrow = mt * 128 + qd * 32 + lane;            // query index inside the sample
    const uint32_t tlane = tmem_base + ((uint32_t)(qd * 32) << 16);
    constexpr float L2E = 1.4426950408889634f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, acc[2][16], corr[2];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[h2][c] = 0.f;
    for (int i = 0; i < ntiles; ++i) {
      const int nvalid = min(XT_KEYS, k_end - (k_begin + i * XT_KEYS));
      mbar_wait(s_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int hh = 2 * sub + h2;
        float v[2][32];
        tmem_ld32_nowait(tlane + (uint32_t)(64 * hh), v[0]);
        tmem_ld32_nowait(tlane + (uint32_t)(64 * hh + 32), v[1]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (u * 32 + j < nvalid) mx = fmaxf(mx, v[u][j]);
        const float mn = fmaxf(m[h2], mx);
        corr[h2] = x_ex2((m[h2] - mn) * L2E);              // exp2(-inf) = 0 on the first tile
        const float mc = mn * L2E;
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[u][j] = (u * 32 + j < nvalid) ? x_ex2(fmaf(v[u][j], L2E, -mc)) : 0.f;
            sum += v[u][j];
          }
          // two k-steps (16 keys each): 8 words of bf16 hi pairs | 8 words of mid pairs
          uint32_t w[32];
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float a = v[u][16 * t + 2 * j], bb = v[u][16 * t + 2 * j + 1];
              const uint32_t hi = pack_bf16x2(a, bb);
              w[16 * t + j] = hi;
              w[16 * t + 8 + j] = pack_bf16x2(a - __uint_as_float(hi << 16), bb - __uint_as_float(hi & 0xFFFF0000u));
            }
          tmem_st32(tlane + (uint32_t)(64 * hh + 32 * u), w);
        }
        l[h2] = l[h2] * corr[h2] + sum;
        m[h2] = mn;
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // numerator of this tile: the head's 16 channels of O
      mbar_wait(o_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int hh = 2 * sub + h2;
        float o[16];
        tmem_ld16(tlane + 256u + (uint32_t)(64 * hh + 16 * hh), o);
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[h2][c] = fmaf(acc[h2][c], corr[h2], o[c]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
    if (qrow < p.P) {
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int head = 4 * box + 2 * sub + h2;
        float* pp = p.part + (((size_t)(b * p.heads + head) * p.nsplit + sp) * 18) * p.P + qrow;
        pp[0] = m[h2];
        pp[p.P] = l[h2];
#pragma unroll
        for (int c = 0; c < 16; ++c) pp[(size_t)(2 + c) * p.P] = acc[h2][c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// merge of the key splits: one warp per (batch, head, query); lanes over the splits
__global__ void __launch_bounds__(256)
xattn_combine_kernel(const float* __restrict__ part, float* __restrict__ out, int B, int P, int C, int Hh, int nsplit) {
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= B * Hh * P) return;
  const int qi = wid % P, head = (wid / P) % Hh, b = wid / (P * Hh);
  const float* pp = part + (size_t)(b * Hh + head) * nsplit * 18 * P + qi;
  const size_t ss = (size_t)18 * P;
  float M = -INFINITY;
  for (int s = lane; s < nsplit; s += 32) M = fmaxf(M, pp[s * ss]);
  M = warp_max(M);
  float L = 0.f, acc[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) acc[d] = 0.f;
  for (int s = lane; s < nsplit; s += 32) {
    const float w = expf(pp[s * ss] - M);
    L += w * pp[s * ss + P];
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[d] += w * pp[s * ss + (size_t)(2 + d) * P];
  }
  L = warp_sum(L);
  const float inv = 1.f / L;
  float mine = 0.f;
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    const float a = warp_sum(acc[d]);
    if (lane == d) mine = a * inv;
  }
  if (lane < 16) out[(size_t)(b * P + qi) * C + head * 16 + lane] = mine;
}

// bf16 planes of fp32 rows: hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid); a plane holds 128 channels as 64
// words of packed pairs (lower channel in the low half)
__device__ __forceinline__ void split3(float a, float b, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
  mid = pack_bf16x2(ra, rb);
  lo = pack_bf16x2(ra - __uint_as_float(mid << 16), rb - __uint_as_float(mid & 0xFFFF0000u));
}
// q rows [M, ld] -> [M, 192 words]: hi | mid | lo
__global__ void planes3_kernel(const float* __restrict__ x, int ld, uint32_t* __restrict__ out, int M) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // (row, word)
  if (i >= M * 64) return;
  const int r = i >> 6, wd = i & 63;
  uint32_t hi, mid, lo;
  split3(x[(size_t)r * ld + 2 * wd], x[(size_t)r * ld + 2 * wd + 1], hi, mid, lo);
  uint32_t* o = out + (size_t)r * 192;
  o[wd] = hi; o[64 + wd] = mid; o[128 + wd] = lo;
}
// kv rows [M, ld >= 256] (K = columns 0..127, V = 128..255) -> kplanes [M, 192 words] hi | mid | lo, vplanes [M, 128 words] hi | mid
__global__ void kv_planes_kernel(const float* __restrict__ kv, int ld, uint32_t* __restrict__ kp, uint32_t* __restrict__ vp,
                                 long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (row, word of 128)
  if (i >= M * 128) return;
  const long long r = i >> 7;
  const int wd = (int)(i & 127);
  const float2 v = *reinterpret_cast<const float2*>(kv + (size_t)r * ld + 2 * wd);
  uint32_t hi, mid, lo;
  split3(v.x, v.y, hi, mid, lo);
  if (wd < 64) {
    uint32_t* o = kp + (size_t)r * 192;
    o[wd] = hi; o[64 + wd] = mid; o[128 + wd] = lo;
  } else {
    uint32_t* o = vp + (size_t)r * 128;
    o[wd - 64] = hi; o[wd] = mid;
  }
}

bool make_rows_map(CUtensorMap* m, const void* ptr, long long rows, int ld_words, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)ld_words * 2, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_words * 4};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool g_xt_attr[64];

}  // namespace

extern "C" {

// Operand planes of the tcgen05 cross attention.  q [M, ld] fp32 (128 channels, pre-scaled) -> qplanes [M, 192 words]
// (bf16 hi | mid | lo); kv [Mk, ld_kv >= 256] fp32 (K | V) -> kplanes [Mk, 192 words], vplanes [Mk, 128 words] (hi | mid).
int di_attn_planes_f32(const float* q, int ld_q, void* qplanes, int M, const float* kv, int ld_kv, void* kplanes,
                       void* vplanes, long long Mk, cudaStream_t stream) {
  DI_CHECK_ARG((q == nullptr || (qplanes && M > 0 && ld_q >= 128)) && (kv == nullptr || (kplanes && vplanes && Mk > 0 && ld_kv >= 256 &&
               ld_kv % 2 == 0 && ((uintptr_t)kv & 7) == 0)), "di_attn_planes_f32: bad argument");
  if (q) planes3_kernel<<<di_cdiv((long long)M * 64, 256), 256, 0, stream>>>(q, ld_q, reinterpret_cast<uint32_t*>(qplanes), M);
  if (kv)
    kv_planes_kernel<<<di_cdiv(Mk * 128, 256), 256, 0, stream>>>(kv, ld_kv, reinterpret_cast<uint32_t*>(kplanes),
                                                                 reinterpret_cast<uint32_t*>(vplanes), Mk);
  DI_CHECK_LAUNCH("di_attn_planes_f32");
  return DI_OK;
}

// Number of key splits di_xattn_tc_f32 uses for (B, HW): the caller sizes `part` as B * heads * nsplit * 18 * P floats.
int di_xattn_tc_splits(int B, int HW) {
  int n = 148 / (4 * (B > 0 ? B : 1));
  if (n < 1) n = 1;
  const int tiles = (HW + XT_KEYS - 1) / XT_KEYS;
  if (n > tiles) n = tiles;
  // every split must own at least one key tile
  while (n > 1 && (long long)(n - 1) * (((tiles + n - 1) / n) * XT_KEYS) >= HW) --n;
  return n;
}

// q [B*P, 192 words], k [B*HW, 192 words], v [B*HW, 128 words]: the planes of di_attn_planes_f32; part: workspace
// (di_xattn_tc_splits), out [B*P, 128] fp32.  8 heads x 16 channels.
int di_xattn_tc_f32(const void* q, const void* k, const void* v, float* part, float* out, int B, int P, int HW, int heads,
                    cudaStream_t stream) {
  DI_CHECK_ARG(q && k && v && part && out && B > 0 && P > 0 && HW > 0, "di_xattn_tc_f32: bad argument");
  if (heads != 8 || P > 256) {
    di_set_error("di_xattn_tc_f32: supported configuration is 8 heads x 16 channels, P <= 256 (got heads=%d P=%d)", heads, P);
    return DI_ERR_UNSUPPORTED;
  }
  DI_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 15) == 0, "di_xattn_tc_f32: operands must be 16-byte aligned");
  int devid = 0;
  cudaGetDevice(&devid);
  if (devid < 0 || devid >= 64) {
    di_set_error("di_xattn_tc_f32: device ordinal %d not supported", devid);
    return DI_ERR_UNSUPPORTED;
  }
  if (!g_xt_attr[devid]) {
    if (cudaFuncSetAttribute(xattn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XT_SMEM) != cudaSuccess) {
      di_set_error("di_xattn_tc_f32: cannot reserve %d bytes of shared memory", XT_SMEM);
      return DI_ERR_LAUNCH;
    }
    g_xt_attr[devid] = true;
  }
  CUtensorMap mq, mk, mv;
  if (!(make_rows_map(&mq, q, (long long)B * P, 192, 128) && make_rows_map(&mk, k, (long long)B * HW, 192, XT_KEYS) &&
        make_rows_map(&mv, v, (long long)B * HW, 128, XT_KEYS))) {
    di_set_error("di_xattn_tc_f32: cuTensorMapEncodeTiled failed");
    return DI_ERR_LAUNCH;
  }
  XtParams p{};
  p.P = P; p.HW = HW; p.heads = heads;
  p.nsplit = di_xattn_tc_splits(B, HW);
  const int tiles = (HW + XT_KEYS - 1) / XT_KEYS;
  p.chunk = ((tiles + p.nsplit - 1) / p.nsplit) * XT_KEYS;
  p.part = part;
  const int mtiles = (P + 127) / 128;
  xattn_tc_kernel<<<dim3(p.nsplit, 2 * mtiles, B), XT_THREADS, XT_SMEM, stream>>>(mq, mk, mv, p);
  DI_CHECK_LAUNCH("di_xattn_tc_f32");
  xattn_combine_kernel<<<di_cdiv((long long)B * heads * P, 8), 256, 0, stream>>>(part, out, B, P, heads * 16, heads, p.nsplit);
  DI_CHECK_LAUNCH("di_xattn_tc_f32 (combine)");
  return DI_OK;
}

}  // extern "C"
