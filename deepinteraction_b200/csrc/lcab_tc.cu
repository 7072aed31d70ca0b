// 9x9 local-window attention on the Blackwell tensor cores (tcgen05 + TMEM + TMA): the fused equivalent of the
// reference's  similarFunction -> softmax(. / sqrt(C)) -> weightingFunction
// (projects/mmdet3d_plugin/models/utils/encoder_utils.py:132-134; CUDA kernels
//  models/utils/ops/locatt_ops/kernels.cuh:4-42 `cc2k`, :44-80 `ck2c_ori`) for C = 128 channels.
// Out-of-image taps keep the reference's rule: logit 0 that STILL takes softmax mass, value skipped -- both fall out
// of the TMA zero fill of the halo.
//
// One persistent CTA per SM; a tile = 16 x 8 queries = the 128 lanes of tensor memory; its halo = 24 x 16 = 384 keys.
//
//   S[128 q x 384 keys] = Q K^T       tcgen05.mma kind::f16, both operands K-major 128B-swizzled tiles that TMA drops
//                                     straight from HBM (the projections emit q/k/v as bf16 hi|mid planes, see
//                                     gemm_tc.cu split kind 3); 4 key groups of 6 halo rows (N = 96); error-compensated
//                                     split product Q_mid K_hi + Q_hi K_mid + Q_hi K_hi, fp32 accumulate in TMEM
//                                     columns [0, 384)
//   P = exp(S/sqrt(C) - max)          4 softmax warps, thread = query = TMEM lane: tcgen05.ld of the 12 halo rows the
//                                     warp's 4 x 8 query patch can see (two passes: masked max, then exp / sum),
//                                     P written back IN PLACE as the A operand of the second product: per halo row
//                                     (16 keys = one k-step) 8 words of bf16 hi pairs | 8 words of bf16 mid pairs,
//                                     zeros for keys outside the window
//   O[128 q x 128 ch] = P V           A = P from tensor memory, B = the V halo stage [key][64 ch] as an MN-major
//                                     128B-swizzled operand (descriptor convention pinned by tools/umma_mn_probe.cu),
//                                     accumulator in TMEM columns [384, 512)
//   out = O / sum                     tcgen05.ld -> registers -> swizzled staging -> TMA store (clips the image edge)
//
// Warp roles (576 threads): warp 0 = tile scheduler + TMA producer (Q tile, then K and V stages through one 3-deep
// ring of 48 KB stages; the next tile's Q is loaded and its first K stages are prefetched into L2 while the current
// tile is in its softmax / P V phase), warp 1 = MMA issuer (elect.sync), warps 2..13 = softmax (three warps per TMEM
// lane quarter, each takes two of the quarter's six halo row pairs -- interleaved so that every key group of P is
// complete after at most two exp passes; maxima / sums exchanged through shared memory), warps 14..17 = epilogue.  All hand-offs through mbarriers; the S product of tile i+1 overlaps the epilogue of tile i, the first
// softmax pass overlaps the S product group by group, and the P V product starts as soon as the first 6 halo rows of
// P are written.
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int QR = 16, QC = 8;                 // query tile (rows x cols)
constexpr int HC = 16;                         // halo columns (halo rows: 24)
constexpr int SG = 6;                          // halo rows per stage (96 keys)
constexpr int NG = 4;                          // stages per operand and tile
constexpr int BOXB = SG * HC * 128;            // one TMA box of a stage: 96 keys x 128 B (64 bf16 channels)
constexpr int STAGE_BYTES = 4 * BOXB;          // hi ch 0-63 | hi ch 64-127 | mid ch 0-63 | mid ch 64-127
constexpr int RING = 3;
constexpr int QBOX = 128 * 128;                // 128 queries x 128 B
constexpr int Q_BYTES = 4 * QBOX;
constexpr int EP_BYTES_W = 4096;               // epilogue staging per warp: 32 rows x 128 B
constexpr int Q_OFF = 0, RING_OFF = Q_BYTES, EP_OFF = RING_OFF + RING * STAGE_BYTES, BAR_OFF = EP_OFF + 4 * EP_BYTES_W;
constexpr int NSW = 3;                         // softmax warps per TMEM lane quarter
constexpr int MX_OFF = BAR_OFF + 512;          // float [NSW][128 queries]: pass-1 maxima, then the partial softmax sums
constexpr int TQD = 4;                         // tile-id queue depth
constexpr int WT_SMEM_BYTES = MX_OFF + NSW * 512;  // the dynamic shared-memory window is 1024-byte aligned (checked at run time)
static_assert(WT_SMEM_BYTES <= 232448, "window kernel exceeds the 227 KB shared-memory limit");
constexpr int WT_THREADS = (2 + 4 * NSW + 4) * 32;
constexpr uint32_t O_COL = 384;

// kind::f16: bf16 x bf16 -> fp32, M = 128.  S: N = 96, A and B K-major.  PV: N = 128, B MN-major (bit 16).
constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_PV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
          tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major 128B-swizzled B operand: atoms of 64 elements (128 B) along N x 8 rows along K; LBO = byte stride between
// the N atoms, SBO = 1024 B between 8-row groups along K (tools/umma_mn_probe.cu: D exact with this convention)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct WtParams {
  int H, W, tiles_x, tiles_y, num_tiles;
  float c_log2;        // log2(e) / sqrt(C)
  int* sched;
  int dbg;
};

// optional pipeline trace of CTA 0 (di_lcab_window_tc_set_debug): clock64 stamps per role, 16 per tile
constexpr int WDBG_SLOTS = 6, WDBG_N = 256;
__device__ long long g_wdbg[WDBG_SLOTS * WDBG_N];
#define WSTAMP(slot, i)                                                                          \
  do {                                                                                          \
    if (p.dbg && blockIdx.x == 0 && (i) < WDBG_N) g_wdbg[(slot) * WDBG_N + (i)] = clock64();     \
  } while (0)

__global__ void __launch_bounds__(WT_THREADS, 1)
lcab_window_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                      const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapO,
                      const WtParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* base_ptr = smem_raw;
  if ((base & 1023u) != 0u) __trap();                 // 128B-swizzled operand tiles need 1024-byte alignment
  const uint32_t bars = base + BAR_OFF;
  const uint32_t q_full = bars, q_empty = bars + 8, o_full = bars + 16, o_empty = bars + 24;
  auto full = [&](int s) { return bars + 32u + 8u * s; };
  auto empty = [&](int s) { return bars + 56u + 8u * s; };
  auto s_full = [&](int g) { return bars + 80u + 8u * g; };
  auto p_full = [&](int g) { return bars + 112u + 8u * g; };
  auto tq_full = [&](int i) { return bars + 144u + 8u * i; };
  auto tq_empty = [&](int i) { return bars + 176u + 8u * i; };
  const uint32_t tmem_slot = bars + 208;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + BAR_OFF + 208);
  volatile int* tq = reinterpret_cast<volatile int*>(base_ptr + BAR_OFF + 224);
  const uint32_t sum_full = bars + 240, sum_empty = bars + 248;
  volatile float* mx = reinterpret_cast<volatile float*>(base_ptr + MX_OFF);   // maxima (pass 1) / partial sums (after pass 2)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane < 4) {
    const CUtensorMap* mp = lane == 0 ? &mapQ : lane == 1 ? &mapK : lane == 2 ? &mapV : &mapO;
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(mp)) : "memory");
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4);
    for (int s = 0; s < RING; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int g = 0; g < NG; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(p_full(g), 4 * NSW);
    }
    mbar_init(sum_full, 4 * NSW);
    mbar_init(sum_empty, 4);
    for (int i = 0; i < TQD; ++i) {
      mbar_init(tq_full(i), 1);
      mbar_init(tq_empty(i), 1 + 4 * NSW + 4);         // MMA warp + softmax warps + 4 epilogue warps
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

  auto tile_coords = [&](int tile, int& img, int& y0, int& x0) {
    const int per_img = p.tiles_x * p.tiles_y;
    img = tile / per_img;
    const int t = tile - img * per_img;
    y0 = (t / p.tiles_x) * QR;
    x0 = (t % p.tiles_x) * QC;
  };
  auto take_tile = [&](int tl) -> int {               // consumer side: id of this CTA's tl-th tile, or -1
    const int i = tl % TQD;
    mbar_wait(tq_full(i), (tl / TQD) & 1);
    const int tile = tq[i];
    __syncwarp();
    if (lane == 0) mbar_arrive(tq_empty(i));
    return tile;
  };

  if (warp == 0) {
    // ---------------- tile scheduler + TMA producer ----------------
    if (lane == 0) {
      int it = 0, published = 0;
      bool exhausted = false;
      auto publish = [&]() {
        const int i = published % TQD;
        if (published >= TQD) mbar_wait(tq_empty(i), ((published / TQD) - 1) & 1);
        int tile = atomicAdd(p.sched, 1);
        if (tile >= p.num_tiles) tile = -1;
        tq[i] = tile;
        mbar_arrive(tq_full(i));
        exhausted = tile < 0;
        ++published;
      };
      publish();
      publish();
      auto load_q = [&](int tile) {
        int img, y0, x0;
        tile_coords(tile, img, y0, x0);
        mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
        for (int b = 0; b < 4; ++b) tma_load_4d(base + Q_OFF + b * QBOX, &mapQ, q_full, b * 64, x0, y0, img);
      };
      if (tq[0] >= 0) load_q(tq[0]);
      for (int n = 0;; ++n) {
        const int tile = tq[n % TQD];
        if (tile < 0) break;
        int img, y0, x0;
        tile_coords(tile, img, y0, x0);
        for (int op = 0; op < 2; ++op)
          for (int g = 0; g < NG; ++g, ++it) {
            if (op == 1 && g == 2) {
              // Between V1 and V2 of this tile (both wait for the end of this tile's S product anyway): the next
              // tile's Q tile, and an L2 prefetch of its first two K stages (their ring slots free up much later).
              const int nxt = tq[(n + 1) % TQD];
              if (nxt >= 0) {
                mbar_wait(q_empty, n & 1);              // the S product of this tile has read Q
                WSTAMP(4, n * 16 + 8);
                load_q(nxt);
                int img2, y2, x2;
                tile_coords(nxt, img2, y2, x2);
#pragma unroll
                for (int g2 = 0; g2 < 2; ++g2)
#pragma unroll
                  for (int b = 0; b < 4; ++b)
                    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                                     reinterpret_cast<uint64_t>(&mapK)),
                                 "r"(b * 64), "r"(x2 - 4), "r"(y2 - 4 + g2 * SG), "r"(img2)
                                 : "memory");
              }
            }
            const int s = it % RING;
            if (it >= RING) mbar_wait(empty(s), ((it / RING) - 1) & 1);
            WSTAMP(4, n * 16 + op * 4 + g);
            const uint32_t st = base + RING_OFF + s * STAGE_BYTES;
            mbar_expect_tx(full(s), STAGE_BYTES);
#pragma unroll
            for (int b = 0; b < 4; ++b)
              tma_load_4d(st + b * BOXB, op == 0 ? &mapK : &mapV, full(s), b * 64, x0 - 4, y0 - 4 + g * SG, img);
          }
        if (!exhausted) publish();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    int it = 0;
    const uint32_t qb = base + Q_OFF;
    for (int tl = 0;; ++tl) {
      if (take_tile(tl) < 0) break;
      if (lane == 0) WSTAMP(0, tl * 16);
      mbar_wait(q_full, tl & 1);
      if (lane == 0) WSTAMP(0, tl * 16 + 1);
      // S = Q K^T, one key group (6 halo rows = 96 keys) per stage
      for (int g = 0; g < NG; ++g, ++it) {
        const int s = it % RING;
        mbar_wait(full(s), (it / RING) & 1);
        tc_fence_after();
        if (lane == 0) WSTAMP(0, tl * 16 + 2 + g);
        if (elect_one()) {
          const uint32_t kb = base + RING_OFF + s * STAGE_BYTES;
          const uint32_t d = tmem_base + (uint32_t)(g * SG * HC);
#pragma unroll
          for (int cb = 0; cb < 2; ++cb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t q_hi = umma_desc(qb + cb * QBOX + ks * 32), q_mid = umma_desc(qb + (2 + cb) * QBOX + ks * 32);
              const uint64_t k_hi = umma_desc(kb + cb * BOXB + ks * 32), k_mid = umma_desc(kb + (2 + cb) * BOXB + ks * 32);
              umma_bf16_ss(d, q_mid, k_hi, IDESC_S, (cb | ks) != 0);
              umma_bf16_ss(d, q_hi, k_mid, IDESC_S, 1);
              umma_bf16_ss(d, q_hi, k_hi, IDESC_S, 1);
            }
          umma_commit(empty(s));
          umma_commit(s_full(g));
          if (g == NG - 1) umma_commit(q_empty);
        }
        __syncwarp();
      }
      // O = P V, one halo row (16 keys) per k-step
      if (lane == 0) WSTAMP(0, tl * 16 + 6);
      if (tl > 0) mbar_wait(o_empty, (tl - 1) & 1);      // the epilogue of the previous tile has drained O
      if (lane == 0) WSTAMP(0, tl * 16 + 7);
      for (int g = 0; g < NG; ++g, ++it) {
        const int s = it % RING;
        mbar_wait(full(s), (it / RING) & 1);
        if (lane == 0 && g == 0) WSTAMP(0, tl * 16 + 13);
        mbar_wait(p_full(g), tl & 1);
        tc_fence_after();
        if (lane == 0) WSTAMP(0, tl * 16 + 8 + g);
        if (elect_one()) {
          const uint32_t vb = base + RING_OFF + s * STAGE_BYTES;
          const uint32_t d = tmem_base + O_COL;
#pragma unroll
          for (int t = 0; t < SG; ++t) {
            const uint64_t v_hi = umma_desc_mn(vb + t * 2048, BOXB), v_mid = umma_desc_mn(vb + 2 * BOXB + t * 2048, BOXB);
            const uint32_t p_hi = tmem_base + (uint32_t)((g * SG + t) * HC), p_mid = p_hi + 8u;
            umma_bf16_ts(d, p_mid, v_hi, IDESC_PV, (g | t) != 0);
            umma_bf16_ts(d, p_hi, v_mid, IDESC_PV, 1);
            umma_bf16_ts(d, p_hi, v_hi, IDESC_PV, 1);
          }
          umma_commit(empty(s));
          if (g == NG - 1) umma_commit(o_full);
        }
        __syncwarp();
      }
      if (lane == 0) WSTAMP(0, tl * 16 + 12);
    }
  } else if (warp < 2 + 4 * NSW) {
    // ---------------- softmax (warps 2..13): thread = query = TMEM lane; NSW warps per lane quarter ----------------
    const int qd = warp & 3;                              // TMEM lane quarter = query rows 4 qd .. 4 qd + 3 of the tile
    const int sub = (warp - 2) >> 2;                      // this warp's index among the quarter's softmax warps
    const int qrow = qd * 32 + lane;                      // query index in the tile
    const int qyl = 4 * qd + (lane >> 3), qx = lane & 7;
    const uint32_t colmask = 0x1FFu << qx;                // halo columns qx .. qx + 8 are inside the window
    const uint32_t tlane = tmem_base + ((uint32_t)(qd * 32) << 16);
    const float c = p.c_log2;
    const int r2_a = 2 * qd + sub, r2_b = r2_a + NSW;     // this warp's two heavy row pairs (rows 2 r2, 2 r2 + 1)
    for (int tl = 0;; ++tl) {
      if (take_tile(tl) < 0) break;
      const int ds = warp == 2 ? 1 : (warp == 13 ? 2 : -1);
      if (lane == 0 && ds > 0) WSTAMP(ds, tl * 16);
      // ---- pass 1: exact maximum over the in-window logits of this warp's 4 halo rows (two rows per load)
      float m = -INFINITY;
      int waited = -1;
#pragma unroll 1
      for (int u = 0; u < 2; ++u) {
        const int r2 = u == 0 ? r2_a : r2_b;
        const int hr = 2 * r2;
        const int g = hr / SG;
        if (g > waited) {
          mbar_wait(s_full(g), tl & 1);
          tc_fence_after();
          waited = g;
        }
        float v[32];
        tmem_ld32_nowait(tlane + (uint32_t)(hr * HC), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const bool ra = (unsigned)(hr - qyl) <= 8u, rb = (unsigned)(hr + 1 - qyl) <= 8u;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool cok = (colmask >> j) & 1u;
          if (ra && cok) m = fmaxf(m, v[j]);
          if (rb && cok) m = fmaxf(m, v[16 + j]);
        }
      }
      if (lane == 0 && ds > 0) WSTAMP(ds, tl * 16 + 1);
      if (tl > 0) mbar_wait(sum_empty, (tl - 1) & 1);      // the epilogue has read the previous tile's sums (same buffer)
      mx[sub * 128 + qrow] = m;                            // exchange with the other warps of this lane quarter
      asm volatile("bar.sync %0, %1;" ::"r"(1 + qd), "r"(32 * NSW) : "memory");
#pragma unroll
      for (int o = 0; o < NSW; ++o) m = fmaxf(m, mx[o * 128 + qrow]);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + qd), "r"(32 * NSW) : "memory");   // all have read before anyone writes a sum
      for (int g = waited + 1; g < NG; ++g) mbar_wait(s_full(g), tl & 1);   // pass 2 writes into every key group
      tc_fence_after();
      if (lane == 0 && ds > 0) WSTAMP(ds, tl * 16 + 2);
      const float mc = m * c;
      // ---- pass 2: p = exp((s - m) / sqrt(C)); P (bf16 hi | mid) written over S, zeros outside the window.  Row pairs
      // outside the quarter's 12 rows (zero fill) are dealt round robin.
      float sum = 0.f;
#pragma unroll 1
      for (int r2 = 0; r2 < 12; ++r2) {
        const int hr = 2 * r2;
        const bool heavy = r2 == r2_a || r2 == r2_b;
        const bool in_quarter = r2 >= 2 * qd && r2 < 2 * qd + 6;
        if (heavy || (!in_quarter && r2 % NSW == sub)) {
          uint32_t w[32];
          if (heavy) {
            float v[32];
            tmem_ld32_nowait(tlane + (uint32_t)(hr * HC), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const bool ra = (unsigned)(hr - qyl) <= 8u, rb = (unsigned)(hr + 1 - qyl) <= 8u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const bool cok = (colmask >> j) & 1u;
              v[j] = (ra && cok) ? ex2_approx(fmaf(v[j], c, -mc)) : 0.f;
              v[16 + j] = (rb && cok) ? ex2_approx(fmaf(v[16 + j], c, -mc)) : 0.f;
              sum += v[j] + v[16 + j];
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float a = v[16 * h + 2 * j], b = v[16 * h + 2 * j + 1];
                const uint32_t hi = pack_bf16x2(a, b);
                w[16 * h + j] = hi;
                w[16 * h + 8 + j] = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
              }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = 0u;
          }
          tmem_st32(tlane + (uint32_t)(hr * HC), w);
        }
        if (r2 % 3 == 2) {                                 // this warp's share of key group r2 / 3 is complete
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(r2 / 3));
          if (lane == 0 && ds > 0) WSTAMP(ds, tl * 16 + 3 + r2 / 3);
        }
      }
      // ---- partial sum of this warp's rows -> epilogue warp of the quarter
      mx[sub * 128 + qrow] = sum;
      __syncwarp();
      if (lane == 0) mbar_arrive(sum_full);
    }
  } else {
    // ---------------- epilogue (warps 14..17): O / sum -> swizzled staging -> TMA store of a 4 x 8 query patch ----
    const int qd = warp & 3;
    const int qrow = qd * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t my_ep = base + EP_OFF + (uint32_t)(warp - (2 + 4 * NSW)) * EP_BYTES_W;
    for (int tl = 0;; ++tl) {
      const int tile = take_tile(tl);
      if (tile < 0) break;
      int img, y0, x0;
      tile_coords(tile, img, y0, x0);
      if (lane == 0 && warp == 2 + 4 * NSW) WSTAMP(3, tl * 16);
      mbar_wait(sum_full, tl & 1);
      float tot = 0.f;
#pragma unroll
      for (int o = 0; o < NSW; ++o) tot += mx[o * 128 + qrow];
      const float inv = 1.f / tot;
      __syncwarp();
      if (lane == 0) mbar_arrive(sum_empty);
      if (lane == 0 && warp == 2 + 4 * NSW) WSTAMP(3, tl * 16 + 1);
      mbar_wait(o_full, tl & 1);
      tc_fence_after();
      if (lane == 0 && warp == 2 + 4 * NSW) WSTAMP(3, tl * 16 + 2);
#pragma unroll 1
      for (int cq = 0; cq < 4; ++cq) {
        float v[32];
        tmem_ld32_nowait(tlane + O_COL + (uint32_t)(cq * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cq == 3) {                                     // last TMEM read of this tile: hand O back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty);
          if (lane == 0 && warp == 2 + 4 * NSW) WSTAMP(3, tl * 16 + 3);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging buffer free again
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my_ep + lane * 128 + ((j ^ (lane & 7)) << 4)),
                       "f"(v[4 * j] * inv), "f"(v[4 * j + 1] * inv), "f"(v[4 * j + 2] * inv), "f"(v[4 * j + 3] * inv)
                       : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&mapO, my_ep, cq * 32, x0, y0 + 4 * qd, img);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores retired before exit
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {                              // the last CTA to finish re-arms the scheduler slot
    __threadfence();
    if (atomicAdd(p.sched + 15, 1) == (int)gridDim.x - 1) {
      for (int i = 0; i < 16; ++i) p.sched[i] = 0;
      __threadfence();
    }
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// 4-D bf16 map over a planar operand [N, H, W, ld words]: per pixel 256 bf16 (hi plane of 128 channels | mid plane)
bool make_planar_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int ld_words, int box_x, int box_y) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {256, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld_words * 4, (cuuint64_t)W * ld_words * 4, (cuuint64_t)H * W * ld_words * 4};
  cuuint32_t box[4] = {64, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// fp32 store map over out [N, H, W, ldo]: box = 32 channels x 8 x 4 pixels (one softmax warp's query patch)
bool make_out_map(CUtensorMap* m, float* ptr, int N, int H, int W, int C, int ldo) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ldo * 4, (cuuint64_t)W * ldo * 4, (cuuint64_t)H * W * ldo * 4};
  cuuint32_t box[4] = {32, QC, 4, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct WtDev {
  int num_sms = 0;
  bool attr_set = false;
};
WtDev g_wt_dev[64];
int g_wt_sm_limit = 0;
int g_wt_debug = 0;

}  // namespace

extern "C" {

// Persistent grids of the window kernel use at most n CTAs (0 = all SMs); see di_tc_set_sm_limit.
int di_lcab_window_tc_set_sm_limit(int n) {
  DI_CHECK_ARG(n >= 0, "di_lcab_window_tc_set_sm_limit: n must be >= 0");
  g_wt_sm_limit = n;
  return DI_OK;
}

// Pipeline trace of CTA 0 (diagnostics, tools/trace_window.py): enable, run ONE launch, read 6 x 256 clock64 stamps
// (rows: 0 MMA issuer, 1 / 2 softmax warps 2 / 9, 3 epilogue warp 10, 4 producer; 16 stamps per tile).
int di_lcab_window_tc_set_debug(int on) {
  g_wt_debug = on;
  return DI_OK;
}
int di_lcab_window_tc_debug_read(long long* host_buf) {
  DI_CHECK_ARG(host_buf, "di_lcab_window_tc_debug_read: null buffer");
  if (cudaMemcpyFromSymbol(host_buf, g_wdbg, sizeof(long long) * WDBG_SLOTS * WDBG_N) != cudaSuccess) {
    di_set_error("di_lcab_window_tc_debug_read: copy failed");
    return DI_ERR_LAUNCH;
  }
  return DI_OK;
}

// 9x9 window attention on tcgen05 for PLANAR pre-split operands (di_linear_tcb_split_f32 split_kind 3): q, k, v point
// at [N*H*W] pixels with a per-pixel stride of ld* 32-bit words, each pixel holding 128 bf16 hi values followed by
// 128 bf16 mid values (x = hi + mid to 16 mantissa bits).  out: fp32 [N, H, W, ldo], C = 128 channels written.
int di_lcab_window_tc_f32(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, float* out, int ldo,
                          int N, int H, int W, int C, cudaStream_t stream) {
  DI_CHECK_ARG(q && k && v && out && N > 0 && H > 0 && W > 0, "di_lcab_window_tc_f32: bad argument");
  if (C != 128) {
    di_set_error("di_lcab_window_tc_f32: C must be 128 (got %d)", C);
    return DI_ERR_UNSUPPORTED;
  }
  DI_CHECK_ARG(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 4 == 0 && ldq >= 128 && ldk >= 128 && ldv >= 128 && ldo >= 128,
               "di_lcab_window_tc_f32: strides must be multiples of 4 words and >= 128");
  DI_CHECK_ARG(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0,
               "di_lcab_window_tc_f32: pointers must be 16-byte aligned");
  int devid = 0;
  cudaGetDevice(&devid);
  if (devid < 0 || devid >= 64) {
    di_set_error("di_lcab_window_tc_f32: device ordinal %d not supported", devid);
    return DI_ERR_UNSUPPORTED;
  }
  WtDev& ds = g_wt_dev[devid];
  if (ds.num_sms == 0) {
    cudaDeviceGetAttribute(&ds.num_sms, cudaDevAttrMultiProcessorCount, devid);
    if (ds.num_sms <= 0) ds.num_sms = 148;
  }
  if (!ds.attr_set) {
    if (cudaFuncSetAttribute(lcab_window_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM_BYTES) != cudaSuccess) {
      di_set_error("di_lcab_window_tc_f32: cannot reserve %d bytes of shared memory", WT_SMEM_BYTES);
      return DI_ERR_LAUNCH;
    }
    ds.attr_set = true;
  }
  CUtensorMap mq, mk, mv, mo;
  if (!(make_planar_map(&mq, q, N, H, W, ldq, QC, QR) && make_planar_map(&mk, k, N, H, W, ldk, HC, SG) &&
        make_planar_map(&mv, v, N, H, W, ldv, HC, SG) && make_out_map(&mo, out, N, H, W, C, ldo))) {
    di_set_error("di_lcab_window_tc_f32: cuTensorMapEncodeTiled failed");
    return DI_ERR_LAUNCH;
  }
  WtParams p{};
  p.H = H; p.W = W;
  p.tiles_x = di_cdiv(W, QC);
  p.tiles_y = di_cdiv(H, QR);
  p.num_tiles = N * p.tiles_x * p.tiles_y;
  p.c_log2 = 1.4426950408889634f / sqrtf((float)C);
  p.dbg = g_wt_debug;
  p.sched = tc::sched_slot(stream);
  if (!p.sched) {
    di_set_error("di_lcab_window_tc_f32: cannot resolve the scheduler buffer");
    return DI_ERR_LAUNCH;
  }
  const int sms = (g_wt_sm_limit > 0 && g_wt_sm_limit < ds.num_sms) ? g_wt_sm_limit : ds.num_sms;
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  lcab_window_tc_kernel<<<grid, WT_THREADS, WT_SMEM_BYTES, stream>>>(mq, mk, mv, mo, p);
  DI_CHECK_LAUNCH("di_lcab_window_tc_f32");
  return DI_OK;
}

}  // extern "C"
