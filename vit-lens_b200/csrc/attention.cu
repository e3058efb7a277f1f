// FlashAttention-style multi-head attention for sm_100a, head_dim = 64, bf16 in / fp32 softmax.
//
// Forward: one CTA per (batch, head, 128-query tile); 192 threads:
//   warp 0    TMA producer: Q tile once, then K/V 128-key blocks through a 2-stage smem ring
//   warp 1    tcgen05 issuer: S = Q K^T into TMEM, later O += P V (P from smem, V as MN-major B)
//   warps 2-5 softmax: one query row per thread (tcgen05.ld 32x32b), online max/sum in fp32,
//             P written to smem in the 128B-swizzled K-major layout, O rescaled in TMEM
// TMEM: S 128 cols + O 64 cols (256 allocated) -> two CTAs per SM overlap each other's phases.
// Scores never touch HBM: bytes moved = Q + K + V + O (+ LSE), the algorithmic minimum.
//
// Backward: one CTA per (batch, head); loops key blocks (outer) x query tiles (inner):
//   S = Q K^T -> P = exp(S*scale - lse) ; dP = dO V^T ; dS = P * (dP - D) * scale
//   dV += P^T dO ; dK += dS^T Q ; dQ += dS K      (all five GEMMs on tcgen05, accumulators in TMEM)
// dQ accumulators of every query tile stay in TMEM across key blocks (no atomics, no fp32 scratch).
#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {

constexpr int kHD = 64;
constexpr int kTQ = 128;  // query tile
constexpr int kTK = 128;  // key block
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  int B, H, nq, nk;
  int nq_main;  // query rows handled by the tensor-core tiles (nq minus a short tail done on CUDA cores)
  int q_rows_per_batch, kv_rows_per_batch;
  int causal;
  float scale;
  __nv_bfloat16* o;
  long long ldo;
  float* lse;  // [B, H, nq]
};

// byte offset of element (row, col) in a [rows x 128 B] 128B-swizzled tile (col in bf16 elements, < 64)
__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return static_cast<uint32_t>(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}

// ============================================================================================ forward
constexpr int kFwdThreads = 192;
constexpr int kFwdSmem = 16384 /*Q*/ + 2 * 32768 /*K,V ring*/ + 32768 /*P*/ + 128;  // 2 CTAs / SM must fit in 228 KB

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0) __trap();  // 128B-swizzle atoms need 1024-byte aligned tiles
  uint8_t* base_ptr = smem_raw;
  const uint32_t sQ = base;
  const uint32_t sKV = base + 16384;  // stage s: K at sKV + s*32768, V at +16384
  const uint32_t sP = base + 16384 + 65536;
  uint8_t* sP_ptr = base_ptr + 16384 + 65536;
  const uint32_t bars = sP + 32768;
  const uint32_t bar_q = bars, bar_kvfull0 = bars + 8, bar_kvempty0 = bars + 24, bar_sfull = bars + 40,
                 bar_pfull = bars + 48, bar_odone = bars + 56, tmem_slot = bars + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + 16384 + 65536 + 32768 + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.nq_main + kTQ - 1) / kTQ;
  const int qt = blockIdx.x % nqt;
  const int bh = blockIdx.x / nqt;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qt * kTQ;
  int nk_eff = p.nk;
  if (p.causal) nk_eff = min(p.nk, q0 + kTQ);
  const int nblk = (nk_eff + kTK - 1) / kTK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_kvfull0 + 8 * s, 1);
      mbar_init(bar_kvempty0 + 8 * s, 1);
    }
    mbar_init(bar_sfull, 1);
    mbar_init(bar_pfull, 128);
    mbar_init(bar_odone, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tO = tmem + 128;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_q, 16384);
      tma_load_2d(sQ, &tmQ, bar_q, h * kHD, b * p.q_rows_per_batch + q0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        mbar_wait(bar_kvempty0 + 8 * s, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(bar_kvfull0 + 8 * s, 32768);
        tma_load_2d(sKV + s * 32768, &tmK, bar_kvfull0 + 8 * s, h * kHD, b * p.kv_rows_per_batch + j * kTK);
        tma_load_2d(sKV + s * 32768 + 16384, &tmV, bar_kvfull0 + 8 * s, h * kHD, b * p.kv_rows_per_batch + j * kTK);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(bar_q, 0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        const int nkb = min(kTK, ((nk_eff - j * kTK) + 15) & ~15);
        mbar_wait(bar_kvfull0 + 8 * s, (j >> 1) & 1);
        tc_fence_after();
        const uint32_t sK = sKV + s * 32768, sV = sK + 16384;
        const uint32_t idesc_s = umma_idesc_bf16(kTQ, nkb, 0, 0);
#pragma unroll
        for (int k = 0; k < kHD / 16; ++k)
          umma_ss(tS, umma_desc_sw128(sQ + k * 32, 16, 1024), umma_desc_sw128(sK + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(bar_sfull);
        mbar_wait(bar_pfull, j & 1);
        tc_fence_after();
        const uint32_t idesc_o = umma_idesc_bf16(kTQ, kHD, 0, 1);
        for (int kk = 0; kk < nkb / 16; ++kk)
          umma_ss(tO, umma_desc_sw128(sP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                  umma_desc_sw128(sV + kk * 2048, 16384, 1024), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(bar_kvempty0 + 8 * s);
        umma_commit(bar_odone);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the tile
    const int qrow = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = p.scale * kLog2e;
    // Online softmax with a lazily updated offset: exponents are taken relative to m_off, which only moves when the running
    // row maximum has grown by more than kLazy (log2 units) since it was set.  softmax is invariant to the offset, P stays
    // <= 2^kLazy (exact in bf16 range, fp32 row sums), and the O-accumulator rescale -- a TMEM read-modify-write that has to
    // wait for the previous P V MMA -- only runs for the rare warps where some row's maximum jumped.
    constexpr float kLazy = 8.0f;
    float m = -INFINITY, l = 0.f, m_off = -INFINITY;  // m, m_off in raw score units
    for (int j = 0; j < nblk; ++j) {
      const int nkb = min(kTK, ((nk_eff - j * kTK) + 15) & ~15);
      const int kmax = p.causal ? min(p.nk, qrow + 1) : p.nk;  // valid keys are < kmax
      const bool need_mask = p.causal || (j * kTK + kTK > p.nk);  // any partial block: tcgen05.ld reads whole 32-column groups
      mbar_wait(bar_sfull, j & 1);
      tc_fence_after();
      // pass 1: block row max
      float bm = -INFINITY;
      for (int c = 0; c < nkb; c += 32) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + c, v);
        tc_wait_ld();
        if (!need_mask) {
#pragma unroll
          for (int t = 0; t < 32; ++t) bm = fmaxf(bm, __uint_as_float(v[t]));
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (j * kTK + c + t < kmax && c + t < nkb) bm = fmaxf(bm, __uint_as_float(v[t]));
        }
      }
      const float m_new = fmaxf(m, bm);
      const bool move = (m_new - m_off) * sl2 > kLazy || m_off == -INFINITY;  // first block: m_off = -inf
      const float off_new = move ? m_new : m_off;
      const float msub = (off_new == -INFINITY) ? 0.f : off_new * sl2;
      const float alpha = move ? ex2_approx(m_off * sl2 - msub) : 1.0f;  // m_off = -inf on the first block -> 0
      float bs = 0.f;
      // pass 2: P = exp2(s*sl2 - offset), bf16 into swizzled smem (first nkb columns; masked -> 0)
      for (int c = 0; c < nkb; c += 32) {
        uint32_t v[32];
        tmem_ld32(tS + lane_off + c, v);
        tc_wait_ld();
        uint32_t pk[16];
        if (!need_mask) {
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(v[t]), sl2, -msub));
            const float e1 = ex2_approx(fmaf(__uint_as_float(v[t + 1]), sl2, -msub));
            bs += e0 + e1;
            pk[t >> 1] = pack_bf16(e0, e1);
          }
        } else {
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            float e0 = 0.f, e1 = 0.f;
            if (j * kTK + c + t < kmax && c + t < nkb) e0 = ex2_approx(fmaf(__uint_as_float(v[t]), sl2, -msub));
            if (j * kTK + c + t + 1 < kmax && c + t + 1 < nkb) e1 = ex2_approx(fmaf(__uint_as_float(v[t + 1]), sl2, -msub));
            bs += e0 + e1;
            pk[t >> 1] = pack_bf16(e0, e1);
          }
        }
        uint8_t* chunk = sP_ptr + (c >> 6) * 16384;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int col = (c & 63) + u * 8;
          *reinterpret_cast<uint4*>(chunk + sw128_off(r, col)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
      }
      l = l * alpha + bs;
      m = m_new;
      m_off = off_new;
      if (j > 0 && __any_sync(0xffffffffu, move)) {
        // O *= alpha for the rows whose offset moved (previous PV must have retired)
        mbar_wait(bar_odone, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kHD; c += 16) {
          uint32_t v[16];
          tmem_ld16(tO + lane_off + c, v);
          tc_wait_ld();
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = __float_as_uint(__uint_as_float(v[t]) * alpha);
          tmem_st16(tO + lane_off + c, v);
        }
        tc_wait_st();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pfull);
    }
    // epilogue: O / l -> bf16 -> global
    mbar_wait(bar_odone, (nblk - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    const bool ok = qrow < p.nq_main;
    __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * p.q_rows_per_batch + qrow) * p.ldo + h * kHD;
#pragma unroll
    for (int c = 0; c < kHD; c += 32) {
      uint32_t v[32];
      tmem_ld32(tO + lane_off + c, v);
      tc_wait_ld();
      if (ok) {
#pragma unroll
        for (int t = 0; t < 32; t += 8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[t]) * inv, __uint_as_float(v[t + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(v[t + 2]) * inv, __uint_as_float(v[t + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(v[t + 4]) * inv, __uint_as_float(v[t + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(v[t + 6]) * inv, __uint_as_float(v[t + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c + t) = u;
        }
      }
    }
    if (ok && p.lse) p.lse[(static_cast<long long>(b) * p.H + h) * p.nq + qrow] = m_off * p.scale + __logf(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem, 256);
  }
}

// dot(row `row` of a [rows x 64] bf16 128B-swizzled tile, fp32 vector in smem)
__device__ __forceinline__ float dot_row_sw128(const uint8_t* tile, int row, const float* vec) {
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 w = *reinterpret_cast<const uint4*>(tile + sw128_off(row, u * 8));
    const float4 a = *reinterpret_cast<const float4*>(vec + u * 8), c = *reinterpret_cast<const float4*>(vec + u * 8 + 4);
    acc += bf16_lo(w.x) * a.x + bf16_hi(w.x) * a.y + bf16_lo(w.y) * a.z + bf16_hi(w.y) * a.w + bf16_lo(w.z) * c.x + bf16_hi(w.z) * c.y +
           bf16_lo(w.w) * c.z + bf16_hi(w.w) * c.w;
  }
  return acc;
}

// ============================================================================================ backward
constexpr int kBwdThreads = 192;
constexpr int kBwdMaxQT = 3;  // nq <= 384
// Q tiles | dO tiles | K | V | P | dS | lse,D | barriers
constexpr int kMaxTail = 4;  // tail rows / keys (n % 128) folded in on CUDA cores when 1..kMaxTail
constexpr int kTailFloats = 2 * kMaxTail * 2 * kHD + 2 * kMaxTail;  // q|dO, k|v vectors + (lse2, D) per tail query
constexpr int kBwdSmem = 2 * kBwdMaxQT * 16384 + 2 * 16384 + 2 * 32768 + 2 * kBwdMaxQT * kTQ * 4 + kTailFloats * 4 + 1024 + 128;

struct AttnBwdParams {
  int B, H, nq, nk;
  int nq_main, nk_main;  // rows / keys covered by the tensor-core tiles; the short tails [n_main, n) are
  int tq, tk;            // folded in on CUDA cores (epilogue corrections here + attn_bwd_tail_kernel)
  long long* dbg;        // optional clock64 timeline (first 8 CTAs)
  const __nv_bfloat16 *q, *k, *v;
  long long ldq, ldk, ldv;
  int q_rows_per_batch, kv_rows_per_batch;
  int causal;
  float scale;
  const __nv_bfloat16* o;
  const __nv_bfloat16* dout;
  long long ldo, lddo;
  const float* lse;
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base;                          // kBwdMaxQT x 16 KB
  const uint32_t sDO = sQ + kBwdMaxQT * 16384;       // kBwdMaxQT x 16 KB
  const uint32_t sK = sDO + kBwdMaxQT * 16384;       // 16 KB
  const uint32_t sV = sK + 16384;                    // 16 KB
  const uint32_t sP = sV + 16384;                    // 32 KB
  const uint32_t sDS = sP + 32768;                   // 32 KB
  uint8_t* sP_ptr = base_ptr + (sP - base);
  uint8_t* sDS_ptr = base_ptr + (sDS - base);
  float* s_lse = reinterpret_cast<float*>(base_ptr + (sDS - base) + 32768);  // [kBwdMaxQT*128] (lse * log2e)
  float* s_D = s_lse + kBwdMaxQT * kTQ;
  float* s_tail = s_D + kBwdMaxQT * kTQ;               // [(tq + tk) * 2][64] fp32
  float* s_tail_stat = s_tail + 2 * kMaxTail * 2 * kHD;  // [tq][2]
  uint8_t* sQ_ptr = base_ptr;
  uint8_t* sDO_ptr = base_ptr + (sDO - base);
  uint8_t* sK_ptr = base_ptr + (sK - base);
  uint8_t* sV_ptr = base_ptr + (sV - base);
  const uint32_t bars = sDS + 32768 + 2 * kBwdMaxQT * kTQ * 4 + kTailFloats * 4;
  const uint32_t bar_qdo = bars, bar_kvfull = bars + 8, bar_kvempty = bars + 16, bar_sfull = bars + 24,
                 bar_pfull = bars + 32, bar_dpfull = bars + 40, bar_dsfull = bars + 48, bar_pairdone = bars + 56,
                 bar_dkvfull = bars + 64, bar_dkvfree = bars + 72, bar_dqfull = bars + 80, tmem_slot = bars + 88;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + (bars - base) + 88);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
  const int nqt = (p.nq_main + kTQ - 1) / kTQ;
  const int nkblk = (p.nk_main + kTK - 1) / kTK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_qdo, 1);
    mbar_init(bar_kvfull, 1);
    mbar_init(bar_kvempty, 1 + 128);  // MMA commit + the compute warps (they read K/V rows for the tail corrections)
    mbar_init(bar_sfull, 1);
    mbar_init(bar_pfull, 128);
    mbar_init(bar_dpfull, 1);
    mbar_init(bar_dsfull, 128);
    mbar_init(bar_pairdone, 1);
    mbar_init(bar_dkvfull, 1);
    mbar_init(bar_dkvfree, 128);
    mbar_init(bar_dqfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tDV = tmem + 128, tDK = tmem + 192, tDQ = tmem + 256;  // tDQ + 64*i

  // first query tile that sees key block j at all (causal: query >= key)
  auto first_qt = [&](int j) { return p.causal ? (j * kTK) / kTQ : 0; };

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_qdo, nqt * 2 * 16384);
      for (int i = 0; i < nqt; ++i) {
        tma_load_2d(sQ + i * 16384, &tmQ, bar_qdo, h * kHD, b * p.q_rows_per_batch + i * kTQ);
        tma_load_2d(sDO + i * 16384, &tmDO, bar_qdo, h * kHD, b * p.q_rows_per_batch + i * kTQ);
      }
      for (int j = 0; j < nkblk; ++j) {
        mbar_wait(bar_kvempty, (j & 1) ^ 1);
        mbar_expect_tx(bar_kvfull, 32768);
        tma_load_2d(sK, &tmK, bar_kvfull, h * kHD, b * p.kv_rows_per_batch + j * kTK);
        tma_load_2d(sV, &tmV, bar_kvfull, h * kHD, b * p.kv_rows_per_batch + j * kTK);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(bar_qdo, 0);
      uint32_t pair = 0;
      uint32_t dq_started = 0;  // bit i set once dQ_i has been written (accumulate afterwards)
      for (int j = 0; j < nkblk; ++j) {
        const int nkb = min(kTK, ((p.nk_main - j * kTK) + 15) & ~15);
        mbar_wait(bar_kvfull, j & 1);
        mbar_wait(bar_dkvfree, (j & 1) ^ 1);  // previous block's dK/dV drained by the compute warps
        tc_fence_after();
        const uint32_t idesc_s = umma_idesc_bf16(kTQ, nkb, 0, 0);     // S / dP: [128 q x nkb]
        const uint32_t idesc_kv = umma_idesc_bf16(kTK, kHD, 1, 1);    // dV/dK: A = P^T / dS^T (MN-major), B MN-major
        const uint32_t idesc_q = umma_idesc_bf16(kTQ, kHD, 0, 1);     // dQ: A = dS (K-major), B = K (MN-major)
        bool first = true;
        for (int i = first_qt(j); i < nqt; ++i, ++pair) {
          const uint32_t sQi = sQ + i * 16384, sDOi = sDO + i * 16384;
          // S = Q_i K_j^T      (S/dP region free: the compute warps signalled ds_full of the previous pair)
          if (pair > 0) mbar_wait(bar_dsfull, (pair - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k)
            umma_ss(tS, umma_desc_sw128(sQi + k * 32, 16, 1024), umma_desc_sw128(sK + k * 32, 16, 1024), idesc_s, k > 0);
          umma_commit(bar_sfull);
          // P ready -> dP = dO_i V_j^T (overwrites S), dV_j += P^T dO_i
          mbar_wait(bar_pfull, pair & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k)
            umma_ss(tS, umma_desc_sw128(sDOi + k * 32, 16, 1024), umma_desc_sw128(sV + k * 32, 16, 1024), idesc_s, k > 0);
          umma_commit(bar_dpfull);
#pragma unroll
          for (int kq = 0; kq < kTQ / 16; ++kq)
            umma_ss(tDV, umma_desc_sw128(sP + kq * 2048, 16384, 1024), umma_desc_sw128(sDOi + kq * 2048, 16384, 1024),
                    idesc_kv, (!first || kq > 0) ? 1u : 0u);
          // dS ready -> dK_j += dS^T Q_i ; dQ_i += dS K_j
          mbar_wait(bar_dsfull, pair & 1);
          tc_fence_after();
#pragma unroll
          for (int kq = 0; kq < kTQ / 16; ++kq)
            umma_ss(tDK, umma_desc_sw128(sDS + kq * 2048, 16384, 1024), umma_desc_sw128(sQi + kq * 2048, 16384, 1024),
                    idesc_kv, (!first || kq > 0) ? 1u : 0u);
          const bool dq_acc = (dq_started >> i) & 1u;
          for (int kk = 0; kk < nkb / 16; ++kk)
            umma_ss(tDQ + 64 * i, umma_desc_sw128(sDS + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                    umma_desc_sw128(sK + kk * 2048, 16384, 1024), idesc_q, (dq_acc || kk > 0) ? 1u : 0u);
          dq_started |= 1u << i;
          umma_commit(bar_pairdone);  // P / dS smem reusable
          first = false;
        }
        umma_commit(bar_dkvfull);  // dK_j, dV_j complete
        umma_commit(bar_kvempty);  // K_j, V_j smem reusable
      }
      umma_commit(bar_dqfull);
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = p.scale * kLog2e;
    int dbg_n = 0;
    const bool dbg_on = p.dbg != nullptr && blockIdx.x < 8 && r == 0;
#define VL_STAMP()                                             \
  do {                                                         \
    if (dbg_on && dbg_n < 63) p.dbg[blockIdx.x * 64 + (dbg_n++)] = clock64(); \
  } while (0)
    VL_STAMP();  // 0: compute warps start (after TMEM alloc + barrier init + __syncthreads)
    // prologue: lse (pre-multiplied by log2e) and D = rowsum(dO * O) for every query row of this head
    for (int i = 0; i < nqt; ++i) {
      const int qrow = i * kTQ + r;
      float lse2 = 0.f, dsum = 0.f;
      if (qrow < p.nq_main) {
        lse2 = p.lse[(static_cast<long long>(b) * p.H + h) * p.nq + qrow] * kLog2e;
        const long long grow = static_cast<long long>(b) * p.q_rows_per_batch + qrow;
        const uint4* orow = reinterpret_cast<const uint4*>(p.o + grow * p.ldo + h * kHD);
        const uint4* grad = reinterpret_cast<const uint4*>(p.dout + grow * p.lddo + h * kHD);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 a = orow[u], g = grad[u];
          dsum += bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) + bf16_hi(a.y) * bf16_hi(g.y) +
                  bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) + bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
        }
      }
      s_lse[i * kTQ + r] = lse2;
      s_D[i * kTQ + r] = dsum;
    }
    // tail rows (CUDA-core corrections): stage q/dO of the tail queries and k/v of the tail keys as fp32 in smem
    if (p.tq + p.tk > 0) {
      const int nvec = (p.tq * 2 + p.tk * 2) * kHD;  // [tq][q|dO][64] then [tk][k|v][64]
      for (int idx = r; idx < nvec; idx += 128) {
        const int d = idx % kHD;
        const int which = (idx / kHD) & 1;
        const int t = idx / (2 * kHD);
        float val;
        if (t < p.tq) {
          const long long grow = static_cast<long long>(b) * p.q_rows_per_batch + p.nq_main + t;
          val = which == 0 ? __bfloat162float(p.q[grow * p.ldq + h * kHD + d]) : __bfloat162float(p.dout[grow * p.lddo + h * kHD + d]);
        } else {
          const long long grow = static_cast<long long>(b) * p.kv_rows_per_batch + p.nk_main + (t - p.tq);
          val = which == 0 ? __bfloat162float(p.k[grow * p.ldk + h * kHD + d]) : __bfloat162float(p.v[grow * p.ldv + h * kHD + d]);
        }
        s_tail[idx] = val;
      }
      if (r < p.tq) {  // lse (log2 domain) and D of the tail queries
        const int qrow = p.nq_main + r;
        const long long grow = static_cast<long long>(b) * p.q_rows_per_batch + qrow;
        float dsum = 0.f;
        for (int d = 0; d < kHD; ++d)
          dsum += __bfloat162float(p.o[grow * p.ldo + h * kHD + d]) * __bfloat162float(p.dout[grow * p.lddo + h * kHD + d]);
        s_tail_stat[2 * r] = p.lse[(static_cast<long long>(b) * p.H + h) * p.nq + qrow] * kLog2e;
        s_tail_stat[2 * r + 1] = dsum;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four compute warps only
    }
    VL_STAMP();  // 1: prologue done
    uint32_t pair = 0;
    for (int j = 0; j < nkblk; ++j) {
      const int nkb = min(kTK, ((p.nk_main - j * kTK) + 15) & ~15);
      for (int i = first_qt(j); i < nqt; ++i, ++pair) {
        const int qrow = i * kTQ + r;
        const bool row_ok = qrow < p.nq_main;
        const int kmax = p.causal ? min(p.nk_main, qrow + 1) : p.nk_main;
        const float lse2 = s_lse[i * kTQ + r], Di = s_D[i * kTQ + r];
        mbar_wait(bar_sfull, pair & 1);
        tc_fence_after();
        VL_STAMP();  // S ready
        if (pair > 0) mbar_wait(bar_pairdone, (pair - 1) & 1);  // previous pair's MMAs finished reading P / dS smem
        VL_STAMP();  // previous pair retired
        uint32_t pk[64];  // P of this row, packed bf16 (128 keys)
#pragma unroll
        for (int c = 0; c < kTK; c += 32) {
          uint32_t v[32];
          if (c < nkb) {
            tmem_ld32(tS + lane_off + c, v);
            tc_wait_ld();
          }
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            float e0 = 0.f, e1 = 0.f;
            if (c < nkb && row_ok) {
              if (j * kTK + c + t < kmax) e0 = exp2f(fmaf(__uint_as_float(v[t]), sl2, -lse2));
              if (j * kTK + c + t + 1 < kmax) e1 = exp2f(fmaf(__uint_as_float(v[t + 1]), sl2, -lse2));
            }
            pk[(c + t) >> 1] = pack_bf16(e0, e1);
          }
          uint8_t* chunk = sP_ptr + (c >> 6) * 16384;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int col = (c & 63) + u * 8;
            const int q4 = ((c + u * 8) >> 1);
            *reinterpret_cast<uint4*>(chunk + sw128_off(r, col)) = make_uint4(pk[q4], pk[q4 + 1], pk[q4 + 2], pk[q4 + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_pfull);
        VL_STAMP();  // P written
        // dS = P * (dP - D) * scale
        mbar_wait(bar_dpfull, pair & 1);
        tc_fence_after();
        VL_STAMP();  // dP ready
#pragma unroll
        for (int c = 0; c < kTK; c += 32) {
          uint32_t v[32];
          if (c < nkb) {
            tmem_ld32(tS + lane_off + c, v);
            tc_wait_ld();
          }
          uint32_t dk[16];
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            float d0 = 0.f, d1 = 0.f;
            if (c < nkb) {
              const uint32_t pp = pk[(c + t) >> 1];
              d0 = bf16_lo(pp) * (__uint_as_float(v[t]) - Di) * p.scale;
              d1 = bf16_hi(pp) * (__uint_as_float(v[t + 1]) - Di) * p.scale;
            }
            dk[t >> 1] = pack_bf16(d0, d1);
          }
          uint8_t* chunk = sDS_ptr + (c >> 6) * 16384;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int col = (c & 63) + u * 8;
            *reinterpret_cast<uint4*>(chunk + sw128_off(r, col)) = make_uint4(dk[4 * u], dk[4 * u + 1], dk[4 * u + 2], dk[4 * u + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_dsfull);
        VL_STAMP();  // dS written
      }
      // dK_j / dV_j -> global (thread r owns key row j*128 + r)
      mbar_wait(bar_dkvfull, j & 1);
      tc_fence_after();
      VL_STAMP();  // dK/dV accumulators complete
      {
        const int krow = j * kTK + r;
        const bool ok = krow < p.nk_main;
        const long long grow = static_cast<long long>(b) * p.kv_rows_per_batch + krow;
        // tail queries' contributions to this key row: dV_j += p_tj dO_t ; dK_j += ds_tj q_t
        float pt[kMaxTail], dst_[kMaxTail];
        for (int t = 0; t < p.tq; ++t) {
          const float* qv = s_tail + (2 * t) * kHD;
          const float* gv = qv + kHD;
          const float sdot = dot_row_sw128(sK_ptr, r, qv);
          const float dpv = dot_row_sw128(sV_ptr, r, gv);
          const float pv = ok ? exp2f(fmaf(sdot, sl2, -s_tail_stat[2 * t])) : 0.f;
          pt[t] = pv;
          dst_[t] = pv * (dpv - s_tail_stat[2 * t + 1]) * p.scale;
        }
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          __nv_bfloat16* dst = which == 0 ? p.dv + grow * p.lddv + h * kHD : p.dk + grow * p.lddk + h * kHD;
          const uint32_t t0 = which == 0 ? tDV : tDK;
#pragma unroll
          for (int c = 0; c < kHD; c += 32) {
            uint32_t v[32];
            tmem_ld32(t0 + lane_off + c, v);
            tc_wait_ld();
            for (int t = 0; t < p.tq; ++t) {
              const float coef = which == 0 ? pt[t] : dst_[t];
              const float* vec = s_tail + (2 * t + (which == 0 ? 1 : 0)) * kHD + c;
#pragma unroll
              for (int e2 = 0; e2 < 32; ++e2) v[e2] = __float_as_uint(fmaf(coef, vec[e2], __uint_as_float(v[e2])));
            }
            if (ok) {
#pragma unroll
              for (int t = 0; t < 32; t += 8) {
                uint4 u;
                u.x = pack_bf16(__uint_as_float(v[t]), __uint_as_float(v[t + 1]));
                u.y = pack_bf16(__uint_as_float(v[t + 2]), __uint_as_float(v[t + 3]));
                u.z = pack_bf16(__uint_as_float(v[t + 4]), __uint_as_float(v[t + 5]));
                u.w = pack_bf16(__uint_as_float(v[t + 6]), __uint_as_float(v[t + 7]));
                *reinterpret_cast<uint4*>(dst + c + t) = u;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_dkvfree);
      mbar_arrive(bar_kvempty);
      VL_STAMP();  // dK/dV written
    }
    // dQ tiles -> global
    mbar_wait(bar_dqfull, 0);
    tc_fence_after();
    VL_STAMP();  // dQ complete
    for (int i = 0; i < nqt; ++i) {
      const int qrow = i * kTQ + r;
      const bool ok = qrow < p.nq_main;
      __nv_bfloat16* dst = p.dq + (static_cast<long long>(b) * p.q_rows_per_batch + qrow) * p.lddq + h * kHD;
      // tail keys' contribution to this query row: dQ_i += ds_it k_t
      float dsk[kMaxTail];
      for (int t = 0; t < p.tk; ++t) {
        const float* kv = s_tail + (2 * (p.tq + t)) * kHD;
        const float sdot = dot_row_sw128(sQ_ptr + i * 16384, r, kv);
        const float dpv = dot_row_sw128(sDO_ptr + i * 16384, r, kv + kHD);
        const float pv = ok ? exp2f(fmaf(sdot, sl2, -s_lse[i * kTQ + r])) : 0.f;
        dsk[t] = pv * (dpv - s_D[i * kTQ + r]) * p.scale;
      }
#pragma unroll
      for (int c = 0; c < kHD; c += 32) {
        uint32_t v[32];
        tmem_ld32(tDQ + 64 * i + lane_off + c, v);
        tc_wait_ld();
        for (int t = 0; t < p.tk; ++t) {
          const float* vec = s_tail + (2 * (p.tq + t)) * kHD + c;
#pragma unroll
          for (int e2 = 0; e2 < 32; ++e2) v[e2] = __float_as_uint(fmaf(dsk[t], vec[e2], __uint_as_float(v[e2])));
        }
        if (ok) {
#pragma unroll
          for (int t = 0; t < 32; t += 8) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(v[t]), __uint_as_float(v[t + 1]));
            u.y = pack_bf16(__uint_as_float(v[t + 2]), __uint_as_float(v[t + 3]));
            u.z = pack_bf16(__uint_as_float(v[t + 4]), __uint_as_float(v[t + 5]));
            u.w = pack_bf16(__uint_as_float(v[t + 6]), __uint_as_float(v[t + 7]));
            *reinterpret_cast<uint4*>(dst + c + t) = u;
          }
        }
      }
    }
  }

  if (p.dbg != nullptr && blockIdx.x < 8 && threadIdx.x == 64) p.dbg[blockIdx.x * 64 + 63] = clock64();  // compute done
#undef VL_STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// ============================================================================================ tail rows on CUDA cores
// N = 257 (ViT-L/14: 256 patches + cls) leaves one query row / key past the last full 128-tile.  Padding it to a
// whole tensor-core tile wastes 1/3 (fwd) to 5/9 (bwd) of the MMA work, so rows [n_main, n) are handled by these
// small warp-per-head kernels (O(N * 64) work each) plus the epilogue corrections inside attn_bwd_kernel.
constexpr int kTailMaxIter = 10;  // keys (queries) per lane: n <= 320 (covers N = 257); longer rows use the padded tiles

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void load_row64(const __nv_bfloat16* p, float (&f)[64]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 w = reinterpret_cast<const uint4*>(p)[u];
    f[8 * u] = bf16_lo(w.x); f[8 * u + 1] = bf16_hi(w.x); f[8 * u + 2] = bf16_lo(w.y); f[8 * u + 3] = bf16_hi(w.y);
    f[8 * u + 4] = bf16_lo(w.z); f[8 * u + 5] = bf16_hi(w.z); f[8 * u + 6] = bf16_lo(w.w); f[8 * u + 7] = bf16_hi(w.w);
  }
}
__device__ __forceinline__ float dot_row64(const __nv_bfloat16* p, const float (&f)[64]) {
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 w = reinterpret_cast<const uint4*>(p)[u];
    acc += bf16_lo(w.x) * f[8 * u] + bf16_hi(w.x) * f[8 * u + 1] + bf16_lo(w.y) * f[8 * u + 2] + bf16_hi(w.y) * f[8 * u + 3] +
           bf16_lo(w.z) * f[8 * u + 4] + bf16_hi(w.z) * f[8 * u + 5] + bf16_lo(w.w) * f[8 * u + 6] + bf16_hi(w.w) * f[8 * u + 7];
  }
  return acc;
}

struct AttnTailParams {
  const __nv_bfloat16 *q, *k, *v, *o, *dout;
  __nv_bfloat16 *out_o, *dq, *dk, *dv;
  float* lse;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int B, H, nq, nk, tq, tk;
  float scale;
};

// dot(bf16 row in global memory, fp32 vector in shared memory)
__device__ __forceinline__ float dot_row64_sm(const __nv_bfloat16* p, const float* vec) {
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 w = reinterpret_cast<const uint4*>(p)[u];
    const float4 a = *reinterpret_cast<const float4*>(vec + 8 * u), c = *reinterpret_cast<const float4*>(vec + 8 * u + 4);
    acc += bf16_lo(w.x) * a.x + bf16_hi(w.x) * a.y + bf16_lo(w.y) * a.z + bf16_hi(w.y) * a.w + bf16_lo(w.z) * c.x + bf16_hi(w.z) * c.y +
           bf16_lo(w.w) * c.z + bf16_hi(w.w) * c.w;
  }
  return acc;
}

constexpr int kTailMaxN = 32 * kTailMaxIter;
constexpr int kTailThreads = 128;
constexpr int kTailSegs = kTailThreads / 8;  // 16 key segments x 8 dim groups (8 dims = one 16-byte load each)

// acc[e] += sum over this thread's key segment of w[j] * row_j[dv*8 + e]   (rows of 64 bf16, stride ld)
__device__ __forceinline__ void weighted_rows_segment(const float* w, const __nv_bfloat16* rows, long long ld, int n, float (&acc)[8]) {
  const int dv = threadIdx.x & 7, seg = threadIdx.x >> 3;
  const int per = (n + kTailSegs - 1) / kTailSegs;
  const int j0 = seg * per, j1 = min(n, j0 + per);
#pragma unroll 4
  for (int j = j0; j < j1; ++j) {
    const uint4 u = *reinterpret_cast<const uint4*>(rows + static_cast<long long>(j) * ld + dv * 8);
    const float wj = w[j];
    acc[0] = fmaf(wj, bf16_lo(u.x), acc[0]); acc[1] = fmaf(wj, bf16_hi(u.x), acc[1]);
    acc[2] = fmaf(wj, bf16_lo(u.y), acc[2]); acc[3] = fmaf(wj, bf16_hi(u.y), acc[3]);
    acc[4] = fmaf(wj, bf16_lo(u.z), acc[4]); acc[5] = fmaf(wj, bf16_hi(u.z), acc[5]);
    acc[6] = fmaf(wj, bf16_lo(u.w), acc[6]); acc[7] = fmaf(wj, bf16_hi(u.w), acc[7]);
  }
}
// part[seg][64] <- acc ; then threads < 64 reduce over the segments
__device__ __forceinline__ float reduce_segments(float (*part)[kHD], const float (&acc)[8]) {
  const int dv = threadIdx.x & 7, seg = threadIdx.x >> 3;
  __syncthreads();
#pragma unroll
  for (int e2 = 0; e2 < 8; ++e2) part[seg][dv * 8 + e2] = acc[e2];
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < kHD) {
#pragma unroll
    for (int sg = 0; sg < kTailSegs; ++sg) r += part[sg][threadIdx.x];
  }
  return r;
}

__device__ __forceinline__ float block_reduce_128(float v, float* red, bool is_max) {
  v = is_max ? warp_max_f(v) : warp_sum_f(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < kTailThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// forward: one 128-thread CTA per (b, h, tail query): thread-per-key scores, then 64 dims x 2 key halves for P @ V.
__global__ void __launch_bounds__(kTailThreads, 4) attn_fwd_tail_kernel(const AttnTailParams p) {
  __shared__ float sc[kTailMaxN];
  __shared__ __align__(16) float qv[kHD];
  __shared__ float red[4];
  __shared__ float part[kTailSegs][kHD];
  const int tid = threadIdx.x;
  const int t = blockIdx.x % p.tq;
  const int h = (blockIdx.x / p.tq) % p.H;
  const int b = blockIdx.x / (p.tq * p.H);
  const int qrow = p.nq - p.tq + t;
  const float sl2 = p.scale * kLog2e;
  if (tid < kHD) qv[tid] = __bfloat162float(p.q[(static_cast<long long>(b) * p.nq + qrow) * p.ldq + h * kHD + tid]);
  __syncthreads();
  const __nv_bfloat16* kb = p.k + static_cast<long long>(b) * p.nk * p.ldk + h * kHD;
  const __nv_bfloat16* vb = p.v + static_cast<long long>(b) * p.nk * p.ldv + h * kHD;
  float m = -INFINITY;
  for (int j = tid; j < p.nk; j += kTailThreads) {
    const float sv = dot_row64_sm(kb + static_cast<long long>(j) * p.ldk, qv) * sl2;
    sc[j] = sv;
    m = fmaxf(m, sv);
  }
  m = block_reduce_128(m, red, true);
  float l = 0.f;
  for (int j = tid; j < p.nk; j += kTailThreads) {
    const float e = exp2f(sc[j] - m);
    sc[j] = e;
    l += e;
  }
  l = block_reduce_128(l, red, false);  // also orders the sc[] writes before the reads below
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  weighted_rows_segment(sc, vb, p.ldv, p.nk, acc);
  const float osum = reduce_segments(part, acc);
  if (tid < kHD) {
    p.out_o[(static_cast<long long>(b) * p.nq + qrow) * p.ldo + h * kHD + tid] = __float2bfloat16(osum / l);
    if (tid == 0 && p.lse) p.lse[(static_cast<long long>(b) * p.H + h) * p.nq + qrow] = (m + log2f(l)) * 0.6931471805599453f;
  }
}

// backward: one 128-thread CTA per (b, h, part, t).  part 0: dQ of tail query t (sum over all keys);
// part 1: dK, dV of tail key t (sum over all queries).
__global__ void __launch_bounds__(kTailThreads, 3) attn_bwd_tail_kernel(const AttnTailParams p) {
  __shared__ float s_ds[kTailMaxN];
  __shared__ float s_p[kTailMaxN];
  __shared__ __align__(16) float va[kHD];
  __shared__ __align__(16) float vb_[kHD];
  __shared__ float part[kTailSegs][kHD];
  __shared__ float stat[2];
  const int tid = threadIdx.x;
  const int nper = p.tq + p.tk;
  const int unit = blockIdx.x % nper;
  const int h = (blockIdx.x / nper) % p.H;
  const int b = blockIdx.x / (nper * p.H);
  const float sl2 = p.scale * kLog2e;
  const __nv_bfloat16* qb = p.q + static_cast<long long>(b) * p.nq * p.ldq + h * kHD;
  const __nv_bfloat16* kb = p.k + static_cast<long long>(b) * p.nk * p.ldk + h * kHD;
  const __nv_bfloat16* vb = p.v + static_cast<long long>(b) * p.nk * p.ldv + h * kHD;
  const __nv_bfloat16* ob = p.o + static_cast<long long>(b) * p.nq * p.ldo + h * kHD;
  const __nv_bfloat16* gb = p.dout + static_cast<long long>(b) * p.nq * p.lddo + h * kHD;
  const float* lse = p.lse + (static_cast<long long>(b) * p.H + h) * p.nq;
  if (unit < p.tq) {
    // ---------------- part 0: tail query
    const int qrow = p.nq - p.tq + unit;
    if (tid < kHD) {
      va[tid] = __bfloat162float(qb[static_cast<long long>(qrow) * p.ldq + tid]);
      vb_[tid] = __bfloat162float(gb[static_cast<long long>(qrow) * p.lddo + tid]);
    }
    __syncthreads();
    if (tid == 0) {
      stat[0] = lse[qrow] * kLog2e;
      stat[1] = dot_row64_sm(ob + static_cast<long long>(qrow) * p.ldo, vb_);
    }
    __syncthreads();
    const float lse2 = stat[0], Dq = stat[1];
    for (int j = tid; j < p.nk; j += kTailThreads) {
      const float sdot = dot_row64_sm(kb + static_cast<long long>(j) * p.ldk, va);
      const float dpv = dot_row64_sm(vb + static_cast<long long>(j) * p.ldv, vb_);
      s_ds[j] = exp2f(fmaf(sdot, sl2, -lse2)) * (dpv - Dq) * p.scale;
    }
    __syncthreads();
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    weighted_rows_segment(s_ds, kb, p.ldk, p.nk, acc);
    const float r = reduce_segments(part, acc);
    if (tid < kHD) p.dq[(static_cast<long long>(b) * p.nq + qrow) * p.lddq + h * kHD + tid] = __float2bfloat16(r);
  } else {
    // ---------------- part 1: tail key
    const int krow = p.nk - p.tk + (unit - p.tq);
    if (tid < kHD) {
      va[tid] = __bfloat162float(kb[static_cast<long long>(krow) * p.ldk + tid]);
      vb_[tid] = __bfloat162float(vb[static_cast<long long>(krow) * p.ldv + tid]);
    }
    __syncthreads();
    for (int i = tid; i < p.nq; i += kTailThreads) {
      const float sdot = dot_row64_sm(qb + static_cast<long long>(i) * p.ldq, va);
      const float dpv = dot_row64_sm(gb + static_cast<long long>(i) * p.lddo, vb_);
      float Di = 0.f;  // D_i = dO_i . O_i
      const uint4* orow = reinterpret_cast<const uint4*>(ob + static_cast<long long>(i) * p.ldo);
      const uint4* grow = reinterpret_cast<const uint4*>(gb + static_cast<long long>(i) * p.lddo);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint4 x = orow[u], g = grow[u];
        Di += bf16_lo(x.x) * bf16_lo(g.x) + bf16_hi(x.x) * bf16_hi(g.x) + bf16_lo(x.y) * bf16_lo(g.y) + bf16_hi(x.y) * bf16_hi(g.y) +
              bf16_lo(x.z) * bf16_lo(g.z) + bf16_hi(x.z) * bf16_hi(g.z) + bf16_lo(x.w) * bf16_lo(g.w) + bf16_hi(x.w) * bf16_hi(g.w);
      }
      const float pi = exp2f(fmaf(sdot, sl2, -lse[i] * kLog2e));
      s_p[i] = pi;
      s_ds[i] = pi * (dpv - Di) * p.scale;
    }
    __syncthreads();
    float acck[8] = {0, 0, 0, 0, 0, 0, 0, 0}, accv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    weighted_rows_segment(s_ds, qb, p.ldq, p.nq, acck);
    weighted_rows_segment(s_p, gb, p.lddo, p.nq, accv);
    const float rk = reduce_segments(part, acck);
    const float rv = reduce_segments(part, accv);
    if (tid < kHD) {
      p.dk[(static_cast<long long>(b) * p.nk + krow) * p.lddk + h * kHD + tid] = __float2bfloat16(rk);
      p.dv[(static_cast<long long>(b) * p.nk + krow) * p.lddv + h * kHD + tid] = __float2bfloat16(rv);
    }
  }
}

static int tail_rows(int n, int causal) {
  const int t = n % kTQ;
  return (!causal && n > kTQ && t > 0 && t <= kMaxTail && n <= 32 * kTailMaxIter) ? t : 0;
}

static int check_common(const char* who, int B, int H, int nq, int nk, int64_t ldq, int64_t ldk, int64_t ldv) {
  VL_CHECK_ARG(B > 0 && H > 0 && nq > 0 && nk > 0, "%s: non-positive dims", who);
  VL_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0, "%s: leading dims must be multiples of 8", who);
  VL_CHECK_ARG(ldq >= (int64_t)H * kHD && ldk >= (int64_t)H * kHD && ldv >= (int64_t)H * kHD, "%s: leading dim < H*64", who);
  return 0;
}

}  // namespace vl

using namespace vl;

extern "C" {

int vl_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t B, int32_t H, int32_t nq, int32_t nk,
                     int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float scale, int32_t causal, void* stream) {
  VL_CHECK_ARG(q && k && v && o, "vl_attention_fwd: null pointer");
  if (int rc = check_common("vl_attention_fwd", B, H, nq, nk, ldq, ldk, ldv)) return rc;
  VL_CHECK_ARG(ldo % 8 == 0 && ldo >= (int64_t)H * kHD, "vl_attention_fwd: bad ldo");
  VL_CHECK_ARG(!causal || nq == nk, "vl_attention_fwd: causal requires nq == nk");
  {
    // two or more full query tiles, non-causal: the persistent two-group kernel (attention_fwd2.cu) ...
    // debug knob 13: 1 = keep the one-tile-per-CTA kernel below for every shape
    const int t2 = nq % kTQ;
    const int tq2 = (nq > kTQ && t2 == 1) ? 1 : 0;  // the kernel folds one tail row (N = 128 k + 1) in on CUDA cores
    // ... and 128 + 1 rows (the 128-latent Lens + cls of the audio / depth / point recipes): one softmax group idles, but the tail row
    // rides inside the kernel instead of a second launch (measured at batch 512 x 16 heads: 0.227 vs 0.292 ms; at exactly 128 rows
    // the one-tile kernel is the faster one, 0.163 vs 0.202 ms).  knob 17: 1 = the persistent kernel for every single-tile length
    if (!causal && (nq - tq2 > kTQ || tq2 == 1 || (debug_get(17) == 1 && nq - tq2 >= 16)) && debug_get(13) != 1)
      return launch_attn_fwd2(q, k, v, o, lse, B, H, nq, nk, ldq, ldk, ldv, ldo, scale, reinterpret_cast<cudaStream_t>(stream));
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmQ, q, (uint64_t)H * kHD, (uint64_t)B * nq, ldq, kHD, kTQ))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, k, (uint64_t)H * kHD, (uint64_t)B * nk, ldk, kHD, kTK))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, v, (uint64_t)H * kHD, (uint64_t)B * nk, ldv, kHD, kTK))) return rc;
  AttnParams p;
  p.B = B; p.H = H; p.nq = nq; p.nk = nk;
  const int tq = tail_rows(nq, causal);
  p.nq_main = nq - tq;
  p.q_rows_per_batch = nq; p.kv_rows_per_batch = nk;
  p.causal = causal; p.scale = scale;
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.ldo = ldo; p.lse = lse;
  static bool attr = false;
  if (!attr) {
    VL_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    attr = true;
  }
  const int nqt = (p.nq_main + kTQ - 1) / kTQ;
  const long long grid = (long long)B * H * nqt;
  VL_CHECK_ARG(grid < (1ll << 31), "vl_attention_fwd: grid too large");
  attn_fwd_kernel<<<(unsigned)grid, kFwdThreads, kFwdSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  if (int rc2 = launch_check("attn_fwd_kernel")) return rc2;
  if (tq > 0) {
    AttnTailParams t{};
    t.q = reinterpret_cast<const __nv_bfloat16*>(q); t.k = reinterpret_cast<const __nv_bfloat16*>(k); t.v = reinterpret_cast<const __nv_bfloat16*>(v);
    t.out_o = reinterpret_cast<__nv_bfloat16*>(o); t.lse = lse;
    t.ldq = ldq; t.ldk = ldk; t.ldv = ldv; t.ldo = ldo;
    t.B = B; t.H = H; t.nq = nq; t.nk = nk; t.tq = tq; t.tk = 0; t.scale = scale;
    attn_fwd_tail_kernel<<<(unsigned)(B * H * tq), kTailThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t);
    return launch_check("attn_fwd_tail_kernel");
  }
  return 0;
}

int vl_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq, void* dk,
                     void* dv, int32_t B, int32_t H, int32_t nq, int32_t nk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                     int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv, float scale, int32_t causal, void* stream) {
  VL_CHECK_ARG(q && k && v && o && dout && lse && dq && dk && dv, "vl_attention_bwd: null pointer");
  if (int rc = check_common("vl_attention_bwd", B, H, nq, nk, ldq, ldk, ldv)) return rc;
  VL_CHECK_ARG(ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0, "vl_attention_bwd: bad leading dims");
  VL_CHECK_ARG(!causal || nq == nk, "vl_attention_bwd: causal requires nq == nk");
  // debug knob 12: 0 = default (non-causal: third-generation kernel, keys on the TMEM lanes; causal: second generation),
  // 2 = second generation for everything, 1 = the first-generation kernel (one compute group + separate tail kernel, nq <= 384)
  if (debug_get(12) == 0 && !causal)
    return launch_attn_bwd3(q, k, v, o, dout, lse, dq, dk, dv, B, H, nq, nk, ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, scale,
                            reinterpret_cast<cudaStream_t>(stream));
  if (debug_get(12) != 1)
    return launch_attn_bwd2(q, k, v, o, dout, lse, dq, dk, dv, B, H, nq, nk, ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, scale, causal,
                            reinterpret_cast<cudaStream_t>(stream));
  const int tq = tail_rows(nq, causal), tk = tail_rows(nk, causal);
  if (nq - tq > kBwdMaxQT * kTQ) {
    set_error("vl_attention_bwd: nq=%d > %d not supported", nq, kBwdMaxQT * kTQ);
    return VL_ENOTSUP;
  }
  CUtensorMap tmQ, tmK, tmV, tmDO;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmQ, q, (uint64_t)H * kHD, (uint64_t)B * nq, ldq, kHD, kTQ))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, k, (uint64_t)H * kHD, (uint64_t)B * nk, ldk, kHD, kTK))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, v, (uint64_t)H * kHD, (uint64_t)B * nk, ldv, kHD, kTK))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDO, dout, (uint64_t)H * kHD, (uint64_t)B * nq, lddo, kHD, kTQ))) return rc;
  AttnBwdParams p;
  p.B = B; p.H = H; p.nq = nq; p.nk = nk;
  p.nq_main = nq - tq; p.nk_main = nk - tk; p.tq = tq; p.tk = tk;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k); p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv;
  p.q_rows_per_batch = nq; p.kv_rows_per_batch = nk;
  p.causal = causal; p.scale = scale;
  p.o = reinterpret_cast<const __nv_bfloat16*>(o); p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.ldo = ldo; p.lddo = lddo; p.lse = lse;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.dbg = debug_buffer();
  static bool attr = false;
  if (!attr) {
    VL_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    attr = true;
  }
  attn_bwd_kernel<<<(unsigned)(B * H), kBwdThreads, kBwdSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, tmDO, p);
  if (int rc2 = launch_check("attn_bwd_kernel")) return rc2;
  if (tq + tk > 0) {
    AttnTailParams t{};
    t.q = p.q; t.k = p.k; t.v = p.v; t.o = p.o; t.dout = p.dout;
    t.dq = p.dq; t.dk = p.dk; t.dv = p.dv; t.lse = const_cast<float*>(lse);
    t.ldq = ldq; t.ldk = ldk; t.ldv = ldv; t.ldo = ldo; t.lddo = lddo; t.lddq = lddq; t.lddk = lddk; t.lddv = lddv;
    t.B = B; t.H = H; t.nq = nq; t.nk = nk; t.tq = tq; t.tk = tk; t.scale = scale;
    attn_bwd_tail_kernel<<<(unsigned)(B * H * (tq + tk)), kTailThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t);
    return launch_check("attn_bwd_tail_kernel");
  }
  return 0;
}
}
