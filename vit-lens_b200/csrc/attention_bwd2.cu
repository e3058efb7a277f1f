// FlashAttention backward for sm_100a, head_dim = 64 -- two-group pipelined version.
//
// One CTA per (batch, head) and per launch a range of at most 256 query rows (+ up to 4 "tail" query rows) against every
// key.  384 threads (three warpgroups; setmaxnreg moves registers from the producer / issuer warpgroup to the compute groups):
//   warp 0      TMA producer   Q / dO / O tiles once, K / V 128-key blocks through a 2-stage ring
//   warp 1      tcgen05 issuer five GEMMs per (key block, query tile): S = Q K^T, dP = dO V^T, dV += P^T dO,
//                              dK += dS^T Q, dQ += dS K -- accumulators in TMEM
//   warps 4-7   compute group 0 (query tile 0): thread r owns query row r of its tile
//   warps 8-11  compute group 1 (query tile 1)
// The two groups work on the two query tiles of a key block at the same time (own S/dP TMEM region, own P/dS staging
// block), so the exp / dS arithmetic of one tile overlaps the MMAs and TMEM traffic of the other, and every SM
// sub-partition has two resident compute warps.  The dK / dV / dQ epilogues are split between the groups.
// TMEM (512 columns): S/dP of group 0 | S/dP of group 1 | dV | dK | dQ tile 0 | dQ tile 1.
//
// Sequence lengths that leave 1..4 rows past the last 128-row tile (ViT-L/14: 257 = 2 x 128 + cls) are not padded to
// a third tile: those tail queries / tail keys are folded in on CUDA cores inside this kernel (rank-1 corrections of
// the accumulators plus small smem mat-vecs), so no second kernel re-reads Q / K / V / dO.
// Longer sequences (nq > 256 + tail) are covered by several launches over query ranges; launches after the first add
// into dK / dV (accum_kv).
#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {
namespace bwd2 {

constexpr int kHD = 64;
constexpr int kT = 128;  // query tile / key block
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kThreads = 384;  // warpgroup 0: TMA producer, MMA issuer (+ two idle warps); warpgroups 1, 2: the two compute groups
constexpr int kMaxTail = 4;

// shared memory map (bytes)
constexpr int kOffQ = 0;                        // 2 x 16 KB
constexpr int kOffDO = kOffQ + 2 * 16384;       // 2 x 16 KB
constexpr int kOffK = kOffDO + 2 * 16384;       // 2 stages x 16 KB
constexpr int kOffV = kOffK + 2 * 16384;        // 2 stages x 16 KB
constexpr int kOffPD = kOffV + 2 * 16384;       // 2 groups x 32 KB (P, then dS in place; O tile during the prologue)
constexpr int kOffF = kOffPD + 2 * 32768;       // fp32 scratch
// fp32 scratch layout (floats)
constexpr int kFTq = 0;                               // [kMaxTail][2][64]  q_t | dO_t
constexpr int kFTk = kFTq + kMaxTail * 2 * kHD;       // [kMaxTail][2][64]  k_t | v_t
constexpr int kFStat = kFTk + kMaxTail * 2 * kHD;     // [kMaxTail][2]      lse2_t, D_t
constexpr int kFDq = kFStat + kMaxTail * 2;           // [kMaxTail][64]     dQ of the tail queries
constexpr int kFDk = kFDq + kMaxTail * kHD;           // [2 groups][kMaxTail][64] dK of the tail keys (one copy per compute group: no
constexpr int kFDv = kFDk + 2 * kMaxTail * kHD;       // [2 groups][kMaxTail][64] dV   cross-group accumulation order to depend on)
constexpr int kFCoef = kFDv + 2 * kMaxTail * kHD;     // [2 groups][2][128] per-row coefficients for the mat-vecs
constexpr int kFPart = kFCoef + 2 * 2 * kT;           // [2 groups][4 warps][64] per-warp partial sums of a mat-vec
constexpr int kFEnd = kFPart + 2 * 4 * kHD;
constexpr int kOffBar = kOffF + kFEnd * 4;
constexpr int kSmem = kOffBar + 256 + 1024;

struct Params {
  int B, H, nq, nk;     // full sequence lengths (rows per batch)
  int q0, nq_main, tq;  // this launch: rows [q0, q0 + nq_main) on tensor cores, [q0 + nq_main, +tq) on CUDA cores
  int nk_main, tk;      // keys [0, nk_main) in 128-key blocks, tail keys [nk_main, nk_main + tk)
  int causal, accum_kv;
  float scale;
  const __nv_bfloat16 *q, *k, *v, *o, *dout;
  long long ldq, ldk, ldv, ldo, lddo;
  const float* lse;
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  long long* dbg;
};

__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return static_cast<uint32_t>(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}
// dot(row `row` of a [128 x 64] bf16 128B-swizzled tile, fp32 vector in smem)
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
__device__ __forceinline__ float dot_rows_sw128(const uint8_t* a, const uint8_t* b, int row) {
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 x = *reinterpret_cast<const uint4*>(a + sw128_off(row, u * 8));
    const uint4 y = *reinterpret_cast<const uint4*>(b + sw128_off(row, u * 8));
    acc += bf16_lo(x.x) * bf16_lo(y.x) + bf16_hi(x.x) * bf16_hi(y.x) + bf16_lo(x.y) * bf16_lo(y.y) + bf16_hi(x.y) * bf16_hi(y.y) +
           bf16_lo(x.z) * bf16_lo(y.z) + bf16_hi(x.z) * bf16_hi(y.z) + bf16_lo(x.w) * bf16_lo(y.w) + bf16_hi(x.w) * bf16_hi(y.w);
  }
  return acc;
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void both_groups_sync() { asm volatile("bar.sync 3, 256;" ::: "memory"); }

// out[d] += sum_r coef[r] * tile[r][d] over the 128 rows of a swizzled [128 x 64] bf16 tile; executed by the 128 threads
// of one compute group g (x = thread index in the group): thread x covers dims (2*(x&31), +1) of rows [32*(x>>5), +32).
// The four warps' partial sums meet in `part` ([4][64] floats of this group) and are added in warp order by the first 64
// threads -- a fixed order, no floating-point atomics, so the kernel is bit-reproducible.  Ends with a group barrier.
__device__ __forceinline__ void matvec_rows(const float* coef, const uint8_t* tile, float* out, int x, float* part, int g) {
  const int dp = x & 31, r0 = (x >> 5) * 32;
  float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
  for (int rr = 0; rr < 32; ++rr) {
    const int row = r0 + rr;
    const uint32_t w = *reinterpret_cast<const uint32_t*>(tile + sw128_off(row, 2 * dp));
    const float c = coef[row];
    a0 = fmaf(c, bf16_lo(w), a0);
    a1 = fmaf(c, bf16_hi(w), a1);
  }
  part[(x >> 5) * kHD + 2 * dp] = a0;
  part[(x >> 5) * kHD + 2 * dp + 1] = a1;
  group_sync(g);
  if (x < kHD) out[x] += ((part[x] + part[kHD + x]) + part[2 * kHD + x]) + part[3 * kHD + x];
  group_sync(g);
}

__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const uint32_t (&v)[32], bool accum) {
#pragma unroll
  for (int t = 0; t < 32; t += 8) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[t + e]);
    if (accum) {
      const uint4 old = *reinterpret_cast<const uint4*>(dst + t);
      f[0] += bf16_lo(old.x); f[1] += bf16_hi(old.x); f[2] += bf16_lo(old.y); f[3] += bf16_hi(old.y);
      f[4] += bf16_lo(old.z); f[5] += bf16_hi(old.z); f[6] += bf16_lo(old.w); f[7] += bf16_hi(old.w);
    }
    *reinterpret_cast<uint4*>(dst + t) = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

// row r, columns [c, c + 32) of a 128B-swizzled [128 x 64] bf16 staging tile <- 32 fp32 accumulators
__device__ __forceinline__ void stage_row32(uint8_t* tile, int r, int c, const uint32_t (&v)[32]) {
#pragma unroll
  for (int t = 0; t < 32; t += 8)
    *reinterpret_cast<uint4*>(tile + sw128_off(r, c + t)) =
        make_uint4(pack_bf16(__uint_as_float(v[t]), __uint_as_float(v[t + 1])), pack_bf16(__uint_as_float(v[t + 2]), __uint_as_float(v[t + 3])),
                   pack_bf16(__uint_as_float(v[t + 4]), __uint_as_float(v[t + 5])), pack_bf16(__uint_as_float(v[t + 6]), __uint_as_float(v[t + 7])));
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmDQ,
                 const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + kOffQ, sDO = base + kOffDO, sK = base + kOffK, sV = base + kOffV, sPD = base + kOffPD;
  float* sf = reinterpret_cast<float*>(bp + kOffF);
  const uint32_t bars = base + kOffBar;
  // barrier slots (8 bytes each)
  const uint32_t bar_qdo = bars;
  auto bar_kvfull = [&](int s) { return bars + 8u * (1 + s); };
  auto bar_kvempty = [&](int s) { return bars + 8u * (3 + s); };
  auto bar_sfull = [&](int g) { return bars + 8u * (5 + g); };
  auto bar_pfull = [&](int g) { return bars + 8u * (7 + g); };
  auto bar_dpfull = [&](int g) { return bars + 8u * (9 + g); };
  auto bar_dsfull = [&](int g) { return bars + 8u * (11 + g); };
  auto bar_pairdone = [&](int g) { return bars + 8u * (13 + g); };
  const uint32_t bar_dkvfull = bars + 8u * 15, bar_dkvfree = bars + 8u * 16, bar_dqfull = bars + 8u * 17, tmem_slot = bars + 8u * 18,
                 bar_do = bars + 8u * 19;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(bp + kOffBar + 8 * 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
  const int nqt = (p.nq_main + kT - 1) / kT;  // 1 or 2 (0 when only tail queries remain)
  const int nkblk = (p.nk_main + kT - 1) / kT;
  const long long qrow0 = static_cast<long long>(b) * p.nq + p.q0;  // global row of this launch's first query
  const long long krow0 = static_cast<long long>(b) * p.nk;
  // causal: key block j is needed by query tile g iff its first key <= the tile's last query
  auto pair_active = [&](int j, int g) { return g < nqt && (!p.causal || j * kT <= p.q0 + g * kT + kT - 1); };
  auto block_active = [&](int j) { return pair_active(j, 0) || pair_active(j, 1); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmO);
    mbar_init(bar_qdo, 1);
    mbar_init(bar_do, 1);
    tma_prefetch_desc(&tmDQ);
    tma_prefetch_desc(&tmDK);
    tma_prefetch_desc(&tmDV);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_kvfull(s), 1);
      mbar_init(bar_kvempty(s), 1 + 8);  // MMA commit + the eight compute warps (tail corrections read K / V rows)
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(bar_sfull(g), 1);
      mbar_init(bar_pfull(g), 4);
      mbar_init(bar_dpfull(g), 1);
      mbar_init(bar_dsfull(g), 4);
      mbar_init(bar_pairdone(g), 1);
    }
    mbar_init(bar_dkvfull, 1);
    mbar_init(bar_dkvfree, 8);
    mbar_init(bar_dqfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // zero the tail accumulators
  for (int i = threadIdx.x; i < 5 * kMaxTail * kHD; i += kThreads) sf[kFDq + i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  auto tS = [&](int g) { return tmem + 128u * g; };
  const uint32_t tDV = tmem + 256, tDK = tmem + 320;
  auto tDQ = [&](int g) { return tmem + 384u + 64u * g; };

  // Register budget per warpgroup (384 threads x 168 at launch): the compute groups hold a row's 128 packed probabilities plus
  // two in-flight TMEM chunks; the producer / issuer need little.  56 + 2 x 224 = 504 <= 3 x 168.
  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
   if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      // issue order = need order: Q tiles and the first K/V block feed S; dO and O are first needed for dP / D
      if (nqt > 0) {
        mbar_expect_tx(bar_qdo, nqt * 16384);
        for (int g = 0; g < nqt; ++g) tma_load_2d(sQ + g * 16384, &tmQ, bar_qdo, h * kHD, static_cast<int>(qrow0) + g * kT);
      }
      int n = 0;
      bool rest_issued = false;
      for (int j = 0; j < nkblk; ++j) {
        if (!block_active(j)) continue;
        const int s = n & 1;
        mbar_wait(bar_kvempty(s), ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(bar_kvfull(s), 32768);
        tma_load_2d(sK + s * 16384, &tmK, bar_kvfull(s), h * kHD, static_cast<int>(krow0) + j * kT);
        tma_load_2d(sV + s * 16384, &tmV, bar_kvfull(s), h * kHD, static_cast<int>(krow0) + j * kT);
        ++n;
        if (!rest_issued && nqt > 0) {
          rest_issued = true;
          mbar_expect_tx(bar_do, nqt * 2 * 16384);
          for (int g = 0; g < nqt; ++g) {
            tma_load_2d(sDO + g * 16384, &tmDO, bar_do, h * kHD, static_cast<int>(qrow0) + g * kT);
            tma_load_2d(sPD + g * 32768 + 16384, &tmO, bar_do, h * kHD, static_cast<int>(qrow0) + g * kT);  // O: second half of the block
          }
        }
      }
    }
   } else if (warp == 1) {
    // ================================================================== MMA issuer
    // One thread issues 32 MMAs per (key block, query tile); its instruction stream is on the critical path of every
    // hand-off (S -> P -> dP -> dS -> dK/dQ), so descriptors are not re-encoded per MMA: the high word is shared and the low
    // word (14-bit start address + leading-dimension field) is a base plus a multiple of 16 bytes.
    if (elect_one() && nqt > 0) {
      mbar_wait(bar_qdo, 0);
      constexpr uint32_t kHi = 0x40004040u;      // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
      constexpr uint32_t kLoK = 1u << 16;        // K-major operands: LBO field unused (16 B)
      constexpr uint32_t kLoMN = 1024u << 16;    // MN-major operands: LBO = 16 KB between 64-wide chunks
      const uint32_t q_k = (sQ >> 4) | kLoK, q_mn = (sQ >> 4) | kLoMN;        // + g * 1024 (16 KB tiles)
      const uint32_t do_k = (sDO >> 4) | kLoK, do_mn = (sDO >> 4) | kLoMN;    // + g * 1024
      const uint32_t pd_k = (sPD >> 4) | kLoK, pd_mn = (sPD >> 4) | kLoMN;    // + g * 2048 (32 KB blocks)
      const uint32_t k_k = (sK >> 4) | kLoK, k_mn = (sK >> 4) | kLoMN;        // + s * 1024
      const uint32_t v_k = (sV >> 4) | kLoK;                                  // + s * 1024
      const uint32_t idesc_kv = umma_idesc_bf16(kT, kHD, 1, 1);  // dV / dK: A = P^T / dS^T (MN-major), B MN-major
      const uint32_t idesc_q = umma_idesc_bf16(kT, kHD, 0, 1);   // dQ: A = dS (K-major), B = K (MN-major)
      uint32_t np[2] = {0, 0};  // pairs issued per group (barrier phases)
      uint32_t dq_started = 0;
      int n = 0;
      for (int j = 0; j < nkblk; ++j) {
        if (!block_active(j)) continue;
        const int s = n & 1;
        const int nkb = min(kT, ((p.nk_main - j * kT) + 15) & ~15);
        const uint32_t ks_k = k_k + s * 1024u, ks_mn = k_mn + s * 1024u, vs_k = v_k + s * 1024u;
        const uint32_t idesc_s = umma_idesc_bf16(kT, nkb, 0, 0);
        mbar_wait(bar_kvfull(s), (n >> 1) & 1);
        tc_fence_after();
        // S_g = Q_g K_j^T  (the S/dP region of group g is free: the group signalled ds_full of its previous pair)
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (!pair_active(j, g)) continue;
          if (np[g] > 0) {
            mbar_wait(bar_dsfull(g), (np[g] - 1) & 1);
            tc_fence_after();
          }
          const uint32_t a = q_k + g * 1024u;
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) umma_ss_lohi(tS(g), a + 2 * k, ks_k + 2 * k, kHi, idesc_s, k > 0);
          umma_commit(bar_sfull(g));
        }
        // dK_j / dV_j accumulators: drained by the compute groups after the previous block
        if (n > 0) {
          mbar_wait(bar_dkvfree, (n - 1) & 1);
          tc_fence_after();
        }
        bool first = true;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (!pair_active(j, g)) continue;
          if (n == 0 && first) mbar_wait(bar_do, 0);
          mbar_wait(bar_pfull(g), np[g] & 1);
          tc_fence_after();
          const uint32_t a = do_k + g * 1024u;
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) umma_ss_lohi(tS(g), a + 2 * k, vs_k + 2 * k, kHi, idesc_s, k > 0);
          const uint32_t pa = pd_mn + g * 2048u, db = do_mn + g * 1024u;
#pragma unroll
          for (int kq = 0; kq < kT / 16; ++kq) umma_ss_lohi(tDV, pa + 128 * kq, db + 128 * kq, kHi, idesc_kv, (!first || kq > 0) ? 1u : 0u);
          umma_commit(bar_dpfull(g));  // dP ready and P consumed: the group may overwrite P with dS
          first = false;
        }
        first = true;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (!pair_active(j, g)) continue;
          mbar_wait(bar_dsfull(g), np[g] & 1);
          tc_fence_after();
          const uint32_t da = pd_mn + g * 2048u, qb = q_mn + g * 1024u;
#pragma unroll
          for (int kq = 0; kq < kT / 16; ++kq) umma_ss_lohi(tDK, da + 128 * kq, qb + 128 * kq, kHi, idesc_kv, (!first || kq > 0) ? 1u : 0u);
          const bool dq_acc = (dq_started >> g) & 1u;
          const uint32_t dsa = pd_k + g * 2048u;
#pragma unroll
          for (int kk = 0; kk < kT / 16; ++kk)
            if (kk < nkb / 16)
              umma_ss_lohi(tDQ(g), dsa + (kk >> 2) * 1024u + (kk & 3) * 2u, ks_mn + 128 * kk, kHi, idesc_q, (dq_acc || kk > 0) ? 1u : 0u);
          dq_started |= 1u << g;
          umma_commit(bar_pairdone(g));  // dS staging block reusable
          ++np[g];
          first = false;
        }
        umma_commit(bar_dkvfull);     // dK_j, dV_j complete
        umma_commit(bar_kvempty(s));  // K_j, V_j smem reusable (the compute warps add their own arrivals)
        ++n;
      }
      umma_commit(bar_dqfull);
    }
   }
  } else {
    // ================================================================== compute groups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int g = (warp - 4) >> 2;      // 0 or 1
    const int quarter = warp & 3;       // TMEM lane quarter of this warp
    const int r = quarter * 32 + lane;  // row inside the tile
    const int x = ((warp - 4) & 3) * 32 + lane;  // thread index inside the group
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = p.scale * kLog2e;
    uint8_t* sQg = bp + kOffQ + g * 16384;
    uint8_t* sDOg = bp + kOffDO + g * 16384;
    uint8_t* sPDg = bp + kOffPD + g * 32768;
    float* coef0 = sf + kFCoef + g * 2 * kT;
    float* coef1 = coef0 + kT;
    float* mvpart = sf + kFPart + g * 4 * kHD;
    float* sDkg = sf + kFDk + g * kMaxTail * kHD;  // this group's share of the tail keys' dK / dV
    float* sDvg = sf + kFDv + g * kMaxTail * kHD;
    const int dbg_cta = static_cast<int>(blockIdx.x) - 4 * static_cast<int>(gridDim.x) / 7;  // a CTA of a later wave (warm caches)
    const bool dbg_on = p.dbg != nullptr && dbg_cta >= 0 && dbg_cta < 8 && x == 0;
    int dbg_n = 0;
#define VL_STAMP()                                                                          \
  do {                                                                                      \
    if (dbg_on && dbg_n < 31) p.dbg[dbg_cta * 64 + g * 32 + (dbg_n++)] = clock64();      \
  } while (0)
    VL_STAMP();

    // ---- prologue: tail vectors (fp32) and per-row statistics
    {
      const int nvec = (p.tq + p.tk) * 2 * kHD;
      for (int idx = threadIdx.x - 128; idx < nvec; idx += 256) {
        const int d = idx % kHD, which = (idx / kHD) & 1, t = idx / (2 * kHD);
        float val;
        if (t < p.tq) {
          const long long grow = qrow0 + p.nq_main + t;
          val = which == 0 ? __bfloat162float(p.q[grow * p.ldq + h * kHD + d]) : __bfloat162float(p.dout[grow * p.lddo + h * kHD + d]);
        } else {
          const long long grow = krow0 + p.nk_main + (t - p.tq);
          val = which == 0 ? __bfloat162float(p.k[grow * p.ldk + h * kHD + d]) : __bfloat162float(p.v[grow * p.ldv + h * kHD + d]);
        }
        sf[kFTq + idx] = val;  // kFTk follows kFTq contiguously only when tq == kMaxTail; place explicitly below
      }
    }
    // (the loop above wrote [tq + tk][2][64] contiguously from kFTq; tail-key vectors are addressed relative to it)
    const float* s_tq = sf + kFTq;                       // [t][q|dO][64]
    const float* s_tk = sf + kFTq + p.tq * 2 * kHD;      // [t'][k|v][64]
    float* s_stat = sf + kFStat;
    if (warp == 4) {  // lse and D = rowsum(dO * O) of the tail queries: one warp, two dims per lane, coalesced 128-byte rows
      for (int t = 0; t < p.tq; ++t) {
        const long long grow = qrow0 + p.nq_main + t;
        const uint32_t ow = *reinterpret_cast<const uint32_t*>(p.o + grow * p.ldo + h * kHD + 2 * lane);
        const uint32_t gw = *reinterpret_cast<const uint32_t*>(p.dout + grow * p.lddo + h * kHD + 2 * lane);
        float dsum = bf16_lo(ow) * bf16_lo(gw) + bf16_hi(ow) * bf16_hi(gw);
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o2);
        if (lane == 0) {
          s_stat[2 * t] = p.lse[(static_cast<long long>(b) * p.H + h) * p.nq + p.q0 + p.nq_main + t] * kLog2e;
          s_stat[2 * t + 1] = dsum;
        }
      }
    }
    const bool tile_ok = g < nqt;
    const int qrow = g * kT + r;                       // row inside this launch's main range
    const bool row_ok = tile_ok && qrow < p.nq_main;
    float lse2 = INFINITY, Di = 0.f;                   // invalid rows: exp2(s - inf) = 0
    if (row_ok) lse2 = p.lse[(static_cast<long long>(b) * p.H + h) * p.nq + p.q0 + qrow] * kLog2e;
    both_groups_sync();  // tail vectors / stats visible to everyone
    bool have_D = false;
    bool store_pending = false;  // a TMA store issued by this group may still be reading its staging block
    VL_STAMP();

    // Tail keys (the cls key of N = 257): per query row r of this tile  p_rt = exp(s_rt - lse_r), ds_rt = p_rt (dp_rt - D_r) scale.
    // dsk[] corrects this row's dQ in the epilogue; dV_t += sum_r p_rt dO_r and dK_t += sum_r ds_rt q_r are small mat-vecs
    // over the tile (shared-memory atomics into the tail accumulators).  None of it depends on the MMAs.
    float dsk[kMaxTail] = {0.f, 0.f, 0.f, 0.f};
    bool tailk_done = false;
    auto tail_keys = [&]() {
      tailk_done = true;
      if (!tile_ok || p.tk == 0) return;
      float ptk[kMaxTail];
      for (int t = 0; t < p.tk; ++t) {
        const float* kv = s_tk + (2 * t) * kHD;
        const float sdot = dot_row_sw128(sQg, r, kv);
        const float dpv = dot_row_sw128(sDOg, r, kv + kHD);
        const float pv = row_ok ? ex2_approx(fmaf(sdot, sl2, -lse2)) : 0.f;
        ptk[t] = pv;
        dsk[t] = pv * (dpv - Di) * p.scale;
      }
      for (int t = 0; t < p.tk; ++t) {
        coef0[r] = ptk[t];
        coef1[r] = dsk[t];
        group_sync(g);
        matvec_rows(coef0, sDOg, sDvg + t * kHD, x, mvpart, g);
        matvec_rows(coef1, sQg, sDkg + t * kHD, x, mvpart, g);
      }
    };

    uint32_t np = 0;  // pairs processed by this group
    int n = 0;        // key blocks processed
    for (int j = 0; j < nkblk; ++j) {
      if (!block_active(j)) continue;
      const int s = n & 1;
      const int nkb = min(kT, ((p.nk_main - j * kT) + 15) & ~15);
      uint8_t* sKs = bp + kOffK + s * 16384;
      uint8_t* sVs = bp + kOffV + s * 16384;
      mbar_wait(bar_kvfull(s), (n >> 1) & 1);  // K / V rows are also read with ordinary loads (tail corrections)
      // tail queries' coefficients for this thread's key row (dV_r += p_tr dO_t ; dK_r += ds_tr q_t ; dQ_t += ds_tr k_r): they
      // depend on K / V and the tail vectors only, so they are computed here, under the S MMA, not after the dK / dV MMAs
      const int krow = j * kT + r;
      const bool krow_ok = krow < p.nk_main;
      float cf[kMaxTail];
      for (int t = 0; t < p.tq; ++t) {
        const float* qv = s_tq + (2 * t) * kHD;
        const float sdot = dot_row_sw128(sKs, r, qv);
        const float pv = krow_ok ? ex2_approx(fmaf(sdot, sl2, -s_stat[2 * t])) : 0.f;
        if (g == 0) {
          cf[t] = pv;
        } else {
          const float dpv = dot_row_sw128(sVs, r, qv + kHD);
          cf[t] = pv * (dpv - s_stat[2 * t + 1]) * p.scale;
        }
      }
      if (pair_active(j, g)) {
        const int kmax = p.causal ? min(p.nk_main, p.q0 + qrow + 1) : p.nk_main;
        const bool need_mask = p.causal || (j * kT + kT > p.nk_main);
        mbar_wait(bar_sfull(g), np & 1);
        tc_fence_after();
        VL_STAMP();
        if (np > 0) mbar_wait(bar_pairdone(g), (np - 1) & 1);  // previous pair's MMAs finished reading the staging block
        if (store_pending) {  // ... and so has the bulk store of the previous block's dK / dV rows
          if (x == 0) tma_store_wait_read<0>();
          group_sync(g);
          store_pending = false;
        }
        uint32_t pk[64];  // P of this row, packed bf16 (128 keys)
#pragma unroll
        for (int c = 0; c < kT; c += 32) {
          if (c == 64 && !have_D) {
            // D = rowsum(dO * O): the O tile sits in the second half of this group's staging block, which P is about to
            // overwrite (each thread reads and then writes only its own row)
            mbar_wait(bar_do, 0);
            Di = row_ok ? dot_rows_sw128(sDOg, sPDg + 16384, r) : 0.f;
            have_D = true;
          }
          uint32_t v[32];
          if (c < nkb) {
            tmem_ld32(tS(g) + lane_off + c, v);
            tc_wait_ld();
          }
          if (c >= nkb) {
#pragma unroll
            for (int t = 0; t < 16; ++t) pk[(c >> 1) + t] = 0u;
          } else if (!need_mask) {
            const float2 sl22 = make_float2(sl2, sl2), nl2 = make_float2(-lse2, -lse2);
#pragma unroll
            for (int t = 0; t < 32; t += 2) {  // two scores per FMA-pipe instruction (FFMA2)
              const float2 a = __ffma2_rn(make_float2(__uint_as_float(v[t]), __uint_as_float(v[t + 1])), sl22, nl2);
              pk[(c + t) >> 1] = pack_bf16(ex2_approx(a.x), ex2_approx(a.y));
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; t += 2) {
              const float e0 = (j * kT + c + t < kmax) ? ex2_approx(fmaf(__uint_as_float(v[t]), sl2, -lse2)) : 0.f;
              const float e1 = (j * kT + c + t + 1 < kmax) ? ex2_approx(fmaf(__uint_as_float(v[t + 1]), sl2, -lse2)) : 0.f;
              pk[(c + t) >> 1] = pack_bf16(e0, e1);
            }
          }
          uint8_t* chunk = sPDg + (c >> 6) * 16384;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int col = (c & 63) + u * 8;
            const int q4 = ((c + u * 8) >> 1);
            *reinterpret_cast<uint4*>(chunk + sw128_off(r, col)) = make_uint4(pk[q4], pk[q4 + 1], pk[q4 + 2], pk[q4 + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pfull(g));
        VL_STAMP();
        // dS = P * (dP - D) * scale, written over P (the dV MMA that read P has retired when dp_full fires)
        mbar_wait(bar_dpfull(g), np & 1);
        tc_fence_after();
        VL_STAMP();
        const float nDs = -Di * p.scale;
#pragma unroll
        for (int c = 0; c < kT; c += 32) {
          uint32_t v[32];
          if (c < nkb) {
            tmem_ld32(tS(g) + lane_off + c, v);
            tc_wait_ld();
          }
          uint32_t dk[16];
          const float2 sc2 = make_float2(p.scale, p.scale), nDs2 = make_float2(nDs, nDs);
#pragma unroll
          for (int t = 0; t < 32; t += 2) {
            float2 d = make_float2(0.f, 0.f);
            if (c < nkb) {  // FFMA2 / FMUL2: two elements per FMA-pipe instruction
              const uint32_t pp = pk[(c + t) >> 1];
              const float2 u = __ffma2_rn(make_float2(__uint_as_float(v[t]), __uint_as_float(v[t + 1])), sc2, nDs2);
              d = __fmul2_rn(make_float2(bf16_lo(pp), bf16_hi(pp)), u);
            }
            dk[t >> 1] = pack_bf16(d.x, d.y);
          }
          uint8_t* chunk = sPDg + (c >> 6) * 16384;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int col = (c & 63) + u * 8;
            *reinterpret_cast<uint4*>(chunk + sw128_off(r, col)) = make_uint4(dk[4 * u], dk[4 * u + 1], dk[4 * u + 2], dk[4 * u + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dsfull(g));
        VL_STAMP();
        ++np;
        if (!tailk_done) tail_keys();  // needs D (known since this pair); runs under this pair's dK / dQ MMAs
      }
      // ---- dK_j / dV_j -> global: group 0 writes dV, group 1 writes dK (thread r owns key row j*128 + r)
      mbar_wait(bar_dkvfull, n & 1);
      tc_fence_after();
      VL_STAMP();
      {
        const bool ok = krow_ok;
        const long long grow = krow0 + krow;
        VL_STAMP();  // (tail-query dots: moved to the top of the block)
        __nv_bfloat16* dst = g == 0 ? p.dv + grow * p.lddv + h * kHD : p.dk + grow * p.lddk + h * kHD;
        const uint32_t t0 = g == 0 ? tDV : tDK;
#pragma unroll
        for (int c = 0; c < kHD; c += 32) {
          uint32_t v[32];
          tmem_ld32(t0 + lane_off + c, v);
          tc_wait_ld();
          for (int t = 0; t < p.tq; ++t) {
            const float* vec = s_tq + (2 * t + (g == 0 ? 1 : 0)) * kHD + c;
#pragma unroll
            for (int e2 = 0; e2 < 32; ++e2) v[e2] = __float_as_uint(fmaf(cf[t], vec[e2], __uint_as_float(v[e2])));
          }
          if (p.accum_kv) {
            if (ok) store_row32(dst + c, v, true);
          } else {
            stage_row32(sPDg, r, c, v);  // the staging block is idle here: every MMA of this key block has retired
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dkvfree);
        if (!p.accum_kv) {  // one bulk tensor store per group: whole 128-byte lines, rows past nk_main clipped by the map
          fence_proxy_async_smem();
          group_sync(g);
          if (x == 0) {
            tma_store_3d(g == 0 ? &tmDV : &tmDK, sPD + g * 32768, h * kHD, j * kT, b);
            tma_store_commit();
          }
          store_pending = true;
        }
        VL_STAMP();  // dK / dV rows stored
        if (g == 1 && p.tq > 0) {  // dQ_t += sum_r ds_tr k_r
          for (int t = 0; t < p.tq; ++t) {
            coef0[r] = cf[t];
            group_sync(g);
            matvec_rows(coef0, sKs, sf + kFDq + t * kHD, x, mvpart, g);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_kvempty(s));
      VL_STAMP();
      ++n;
    }

    // ---- dQ tile of this group -> global (+ tail keys' contributions), tail keys' dK / dV from this tile's rows
    if (tile_ok) {
      mbar_wait(bar_dqfull, 0);
      mbar_wait(bar_qdo, 0);
      tc_fence_after();
      VL_STAMP();
      if (store_pending) {
        if (x == 0) tma_store_wait_read<0>();
        group_sync(g);
        store_pending = false;
      }
      if (!tailk_done) {  // (a group without any active pair, causal launches: D was never needed before)
        mbar_wait(bar_do, 0);
        Di = row_ok ? dot_rows_sw128(sDOg, sPDg + 16384, r) : 0.f;
        tail_keys();
      }
      VL_STAMP();  // (tail-key dots: done after the group's first pair)
#pragma unroll
      for (int c = 0; c < kHD; c += 32) {
        uint32_t v[32];
        tmem_ld32(tDQ(g) + lane_off + c, v);
        tc_wait_ld();
        for (int t = 0; t < p.tk; ++t) {
          const float* vec = s_tk + (2 * t) * kHD + c;
#pragma unroll
          for (int e2 = 0; e2 < 32; ++e2) v[e2] = __float_as_uint(fmaf(dsk[t], vec[e2], __uint_as_float(v[e2])));
        }
        stage_row32(sPDg, r, c, v);
      }
      fence_proxy_async_smem();
      group_sync(g);
      if (x == 0) {  // rows past the last main query are clipped by the tensor map
        tma_store_3d(&tmDQ, sPD + g * 32768, h * kHD, p.q0 + g * kT, b);
        tma_store_commit();
      }
      VL_STAMP();  // dQ rows stored
    }
    VL_STAMP();
    // ---- tail x tail and the tail rows' outputs (one warp; 2 dims per lane)
    if (p.tq + p.tk > 0) {
      both_groups_sync();
      if (warp == 4) {
        const int d0 = 2 * lane;
        for (int t = 0; t < p.tq; ++t) {
          float dq0 = sf[kFDq + t * kHD + d0], dq1 = sf[kFDq + t * kHD + d0 + 1];
          const float* qv = s_tq + (2 * t) * kHD;
          const float* gv = qv + kHD;
          for (int u = 0; u < p.tk; ++u) {
            const float* kv = s_tk + (2 * u) * kHD;
            const float* vv = kv + kHD;
            float sd = qv[d0] * kv[d0] + qv[d0 + 1] * kv[d0 + 1];
            float dp = gv[d0] * vv[d0] + gv[d0 + 1] * vv[d0 + 1];
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) {
              sd += __shfl_xor_sync(0xffffffffu, sd, o2);
              dp += __shfl_xor_sync(0xffffffffu, dp, o2);
            }
            const float pv = ex2_approx(fmaf(sd, sl2, -s_stat[2 * t]));
            const float ds = pv * (dp - s_stat[2 * t + 1]) * p.scale;
            dq0 = fmaf(ds, kv[d0], dq0);
            dq1 = fmaf(ds, kv[d0 + 1], dq1);
            sf[kFDk + u * kHD + d0] += ds * qv[d0];  // (group 0's copy collects the tail x tail terms)
            sf[kFDk + u * kHD + d0 + 1] += ds * qv[d0 + 1];
            sf[kFDv + u * kHD + d0] += pv * gv[d0];
            sf[kFDv + u * kHD + d0 + 1] += pv * gv[d0 + 1];
          }
          __nv_bfloat16* dst = p.dq + (qrow0 + p.nq_main + t) * p.lddq + h * kHD + d0;
          *reinterpret_cast<uint32_t*>(dst) = pack_bf16(dq0, dq1);
        }
        __syncwarp();
        for (int u = 0; u < p.tk; ++u) {
          const long long grow = krow0 + p.nk_main + u;
          __nv_bfloat16* dkp = p.dk + grow * p.lddk + h * kHD + d0;
          __nv_bfloat16* dvp = p.dv + grow * p.lddv + h * kHD + d0;
          constexpr int kG1 = kMaxTail * kHD;  // offset of group 1's copy
          float k0 = sf[kFDk + u * kHD + d0] + sf[kFDk + kG1 + u * kHD + d0], k1 = sf[kFDk + u * kHD + d0 + 1] + sf[kFDk + kG1 + u * kHD + d0 + 1];
          float v0 = sf[kFDv + u * kHD + d0] + sf[kFDv + kG1 + u * kHD + d0], v1 = sf[kFDv + u * kHD + d0 + 1] + sf[kFDv + kG1 + u * kHD + d0 + 1];
          if (p.accum_kv) {
            const uint32_t ok_ = *reinterpret_cast<const uint32_t*>(dkp), ov_ = *reinterpret_cast<const uint32_t*>(dvp);
            k0 += bf16_lo(ok_); k1 += bf16_hi(ok_); v0 += bf16_lo(ov_); v1 += bf16_hi(ov_);
          }
          *reinterpret_cast<uint32_t*>(dkp) = pack_bf16(k0, k1);
          *reinterpret_cast<uint32_t*>(dvp) = pack_bf16(v0, v1);
        }
      }
    }
    if (x == 0) tma_store_wait<0>();  // bulk stores complete before the CTA retires
    VL_STAMP();
#undef VL_STAMP
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace bwd2

// Host entry used by vl_attention_bwd (attention.cu).  Query ranges of at most 256 (+tail) rows per launch.
int launch_attn_bwd2(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                     int B, int H, int nq, int nk, long long ldq, long long ldk, long long ldv, long long ldo, long long lddo, long long lddq,
                     long long lddk, long long lddv, float scale, int causal, cudaStream_t stream) {
  using namespace bwd2;
  CUtensorMap tmQ, tmK, tmV, tmDO, tmO;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmQ, q, (uint64_t)H * kHD, (uint64_t)B * nq, ldq, kHD, kT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, k, (uint64_t)H * kHD, (uint64_t)B * nk, ldk, kHD, kT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, v, (uint64_t)H * kHD, (uint64_t)B * nk, ldv, kHD, kT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDO, dout, (uint64_t)H * kHD, (uint64_t)B * nq, lddo, kHD, kT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmO, o, (uint64_t)H * kHD, (uint64_t)B * nq, ldo, kHD, kT))) return rc;
  auto tail = [&](int n) {
    const int t = n % kT;
    return (!causal && n > kT && t > 0 && t <= kMaxTail) ? t : 0;
  };
  const int tq_all = tail(nq), tk = tail(nk);
  const int nq_main_all = nq - tq_all;
  // outputs as [cols, main rows of one batch element, batch]: the row box is clipped per batch element, so partial tiles
  // never spill into the tail rows (written separately) or the next batch element
  CUtensorMap tmDQ, tmDK, tmDV;
  auto out_map = [&](CUtensorMap* m, const void* ptr, long long ld, int rows_main, int rows_all) {
    VL_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "vl_attention_bwd: output pointers must be 16-byte aligned");
    const uint64_t dims[3] = {(uint64_t)H * kHD, (uint64_t)rows_main, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)rows_all * (uint64_t)ld * 2};
    const uint32_t box[3] = {kHD, kT, 1};
    return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, strides, box, true);
  };
  if ((rc = out_map(&tmDQ, dq, lddq, nq_main_all, nq))) return rc;
  if ((rc = out_map(&tmDK, dk, lddk, nk - tk, nk))) return rc;
  if ((rc = out_map(&tmDV, dv, lddv, nk - tk, nk))) return rc;
  static bool attr = false;
  if (!attr) {
    VL_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr = true;
  }
  Params p;
  p.B = B; p.H = H; p.nq = nq; p.nk = nk;
  p.nk_main = nk - tk; p.tk = tk;
  p.causal = causal; p.scale = scale;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k); p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.o = reinterpret_cast<const __nv_bfloat16*>(o); p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lddo = lddo;
  p.lse = lse;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.dbg = debug_buffer();
  for (int q0 = 0; q0 < nq_main_all || q0 == 0; q0 += 2 * kT) {
    p.q0 = q0;
    p.nq_main = nq_main_all - q0 < 2 * kT ? nq_main_all - q0 : 2 * kT;
    p.tq = (q0 + 2 * kT >= nq_main_all) ? tq_all : 0;  // the tail rides with the last range
    p.accum_kv = q0 > 0;
    attn_bwd2_kernel<<<(unsigned)(B * H), kThreads, kSmem, stream>>>(tmQ, tmK, tmV, tmDO, tmO, tmDQ, tmDK, tmDV, p);
    if (int rc2 = launch_check("attn_bwd2_kernel")) return rc2;
  }
  return 0;
}

}  // namespace vl
