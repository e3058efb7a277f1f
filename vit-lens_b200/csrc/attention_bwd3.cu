// FlashAttention backward for sm_100a, head_dim = 64, non-causal -- third generation: keys on the TMEM lanes.
//
// The reference op is nn.MultiheadAttention's backward inside ResidualAttentionBlock (open_clip/transformer.py:241-252) and the
// Perceiver attention of the Lens (perceiver.py:104-154).  One CTA per (batch, head) and per launch a range of at most 256 query
// rows (+ up to 4 "tail" query rows) against every key.
//
// What changed against attention_bwd2.cu (measured there: 37 k cycles per CTA, one serial chain S -> P -> dP -> dS -> dK/dQ with a
// shared-memory hand-off of P and of dS per 128 x 128 pair, the tensor pipe 15 % busy):
//   * transposed scores.  S^T = K Q^T and dP^T = V dO^T put the 128 keys of a block on the TMEM lanes; a compute thread owns one
//     KEY row.  P^T and dS^T are then exactly the A operands of dV += P^T dO and dK += dS^T Q, so they go back into TMEM (bf16
//     pairs over their own fp32 inputs, tcgen05.st) and are consumed from there -- no staging, no async-proxy fence for them;
//   * S^T and dP^T do not depend on each other: both are issued up front and ONE compute pass produces P^T and dS^T (one hand-off
//     per unit instead of two);
//   * a unit is 128 keys x 64 queries: the two compute groups own the two 64-query halves of a query tile, run one after the other
//     through the issuer, and the S^T / dP^T MMAs of a group's NEXT unit are queued right behind its dV / dK MMAs, so one group
//     computes while the tensor pipe works for the other;
//   * only dS goes through shared memory (as dS^T, [keys x queries], the MN-major A operand of dQ += dS K; double-buffered);
//   * tail keys (257 = 2 x 128 + cls) are simply a third, TMA-zero-filled key block whose idle warps skip the math; the tail
//     QUERY rides as an N = 16 unit of group 0 (its S^T / dP^T / dV / dK all on the tensor cores), and only its dQ row -- a
//     mat-vec over the keys -- is left to CUDA cores: the two otherwise idle warps of the producer warpgroup do it concurrently.
// TMEM (512 columns): [S^T | dP^T] of group 0 (64 + 64) | [S^T | dP^T] of group 1 | dV | dK | dQ tile 0 | dQ tile 1.
// Deterministic: no atomics anywhere; every sum has a fixed order.
#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {
namespace bwd3 {

constexpr int kHD = 64;
constexpr int kT = 128;   // keys per block / queries per tile
constexpr int kQH = 64;   // queries per group and tile (half a tile)
constexpr int kUPT = 1;   // units per group and tile: 1 = 64-query units, one TMEM buffer per group; 2 = 32-query units, double-buffered
                          // (measured: 0.518 vs 0.553 ms at ViT-L/14 batch 256 -- every SS MMA re-reads its 4 KB A operand from shared
                          // memory, ~40 cycles whatever N is, so halving N doubles the tensor-pipe time of S^T / dP^T)
constexpr int kQU = kQH / kUPT;  // queries per unit
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kThreads = 384;  // warpgroup 0: TMA producer, MMA issuer, two helper warps; warpgroups 1, 2: the compute groups
constexpr int kMaxTail = 4;

// shared memory map (bytes)
constexpr int kOffQ = 0;                         // 2 tiles x 16 KB
constexpr int kOffDO = kOffQ + 2 * 16384;        // 2 tiles x 16 KB
constexpr int kOffK = kOffDO + 2 * 16384;        // 2 stages x 16 KB
constexpr int kOffV = kOffK + 2 * 16384;         // 2 stages x 16 KB
constexpr int kOffST = kOffV + 2 * 16384;        // 2 buffers x 32 KB: dS^T [128 keys x (64 | 64) queries]; O tiles in the prologue
constexpr int kOffQt = kOffST + 2 * 32768;       // tail queries  [16 x 64] (2 KB)
constexpr int kOffDOt = kOffQt + 2048;           // tail dO rows  [16 x 64] (2 KB)
constexpr int kOffKt = kOffDOt + 2048;           // tail keys     [16 x 64] (2 KB)
constexpr int kOffVt = kOffKt + 2048;            // tail values   [16 x 64] (2 KB)
constexpr int kOffF = kOffVt + 2048;             // fp32 scratch
constexpr int kNStat = 2 * kT + 16;              // main rows of the range + the N = 16 tail unit
constexpr int kFNl = 0;                          // [kNStat]  -lse * log2(e)   (-inf for rows that do not exist)
constexpr int kFNd = kFNl + kNStat;              // [kNStat]  -D * scale
constexpr int kFCoef = kFNd + kNStat;            // [2][kMaxTail][128]  dS of the tail queries per key (helpers' mat-vec input)
constexpr int kFTk = kFCoef + 2 * kMaxTail * kT;  // [2][kMaxTail][256]  P | dS of the tail keys per query of the range (helpers' mat-vec input)
constexpr int kFXch = kFTk + 2 * kMaxTail * 2 * kT;  // [3][kMaxTail][32] float2: helper warp 1's partial sums
constexpr int kFEnd = kFXch + 3 * kMaxTail * 64;
constexpr int kOffBar = kOffF + kFEnd * 4;
constexpr int kSmem = kOffBar + 256 + 1024;
static_assert(kSmem <= 227 * 1024, "shared memory budget");

struct Params {
  int B, H, nq, nk;     // full sequence lengths (rows per batch element)
  int q0, nq_main, tq;  // this launch: query rows [q0, q0 + nq_main) as tiles, [q0 + nq_main, +tq) as the tail unit
  int nk_main, tk;      // keys [0, nk_main) in 128-key blocks on the TMEM lanes, tail keys [nk_main, nk_main + tk) as N = 16 columns
  int accum_kv;
  int pf_dist;          // L2 prefetch distance in CTAs (0 = off): the inputs of CTA blockIdx + pf_dist are prefetched by this one
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
__device__ __forceinline__ float dot_rows_sw128(const uint8_t* a, const uint8_t* b, int row) {
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const uint4 x = *reinterpret_cast<const uint4*>(a + sw128_off(row, u * 8));
    const uint4 y = *reinterpret_cast<const uint4*>(b + sw128_off(row, u * 8));
    acc0 += bf16_lo(x.x) * bf16_lo(y.x) + bf16_hi(x.x) * bf16_hi(y.x) + bf16_lo(x.y) * bf16_lo(y.y) + bf16_hi(x.y) * bf16_hi(y.y);
    acc1 += bf16_lo(x.z) * bf16_lo(y.z) + bf16_hi(x.z) * bf16_hi(y.z) + bf16_lo(x.w) * bf16_lo(y.w) + bf16_hi(x.w) * bf16_hi(y.w);
  }
  return acc0 + acc1;
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void both_groups_sync() { asm volatile("bar.sync 3, 256;" ::: "memory"); }

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
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmQt,
                 const __grid_constant__ CUtensorMap tmDOt, const __grid_constant__ CUtensorMap tmKt,
                 const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                 const __grid_constant__ CUtensorMap tmDV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + kOffQ, sDO = base + kOffDO, sK = base + kOffK, sV = base + kOffV, sST = base + kOffST, sQt = base + kOffQt,
                 sDOt = base + kOffDOt, sKt = base + kOffKt, sVt = base + kOffVt;
  float* sf = reinterpret_cast<float*>(bp + kOffF);
  const uint32_t bars = base + kOffBar;
  // barrier slots (8 bytes each)
  // inputs per query tile, so the first tile's units start while the second tile is still on its way: bar_q(i): Q_i (+ the tail
  // keys with tile 0, the tail queries with the last tile); bar_do(i): dO_i, O_i (+ the tail queries' dO with the last tile)
  auto bar_q = [&](int i) { return bars + (i ? 8u * 26 : 0u); };
  auto bar_do = [&](int i) { return bars + (i ? 8u * 27 : 8u); };
  auto bar_kvfull = [&](int s) { return bars + 8u * (2 + s); };
  auto bar_kvempty = [&](int s) { return bars + 8u * (4 + s); };
  auto bar_sfull = [&](int g, uint32_t bsel) { return bars + 8u * (6 + 2 * g + bsel); };  // S^T and dP^T of a unit are in TMEM buffer bsel
  // the group wrote P^T / dS^T (TMEM buffer bsel) and its dS^T columns (smem).  One barrier per buffer: a group runs up to two
  // units ahead of the issuer, and a single barrier would then advance two phases under a waiter (parity aliasing)
  auto bar_pfull = [&](int g, uint32_t bsel) { return bars + 8u * (10 + 2 * g + bsel); };
  auto bar_dsfree = [&](int b) { return bars + 8u * (14 + b); };  // the dQ MMAs reading staging buffer b have retired
  auto bar_cfull = [&](int b) { return bars + 8u * (16 + b); };   // tail-query dS coefficients of a key block written
  auto bar_cfree = [&](int b) { return bars + 8u * (18 + b); };   // ... and consumed by the helper warps
  const uint32_t bar_dkvfull = bars + 8u * 20, bar_dkvfree = bars + 8u * 21, bar_dqfull = bars + 8u * 22, tmem_slot = bars + 8u * 23;
  auto bar_tks = [&](int i) { return bars + 8u * (i ? 28 : 24); };  // scores of tile i against the tail keys (Q K_t^T, dO V_t^T) are in TMEM
  const uint32_t bar_tkc = bars + 8u * 25;  // ... and their P / dS coefficients in shared memory (eight compute warps)
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(bp + kOffBar + 8 * 23);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
  const int nqt = (p.nq_main + kT - 1) / kT;  // 1 or 2
  const int nkblk = (p.nk_main + kT - 1) / kT;
  const long long qrow0 = static_cast<long long>(b) * p.nq + p.q0;  // global row of this launch's first query
  const long long krow0 = static_cast<long long>(b) * p.nk;
  const bool has_tail = p.tq > 0;
  const bool has_tk = p.tk > 0;

  const int dbg_cta0 = static_cast<int>(blockIdx.x) - 4 * static_cast<int>(gridDim.x) / 7;
  const bool dbg0 = p.dbg != nullptr && dbg_cta0 >= 0 && dbg_cta0 < 8;
  if (dbg0 && threadIdx.x == 0) p.dbg[1024 + dbg_cta0 * 8 + 0] = clock64();  // kernel entry
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_q(i), 1);
      mbar_init(bar_do(i), 1);
      mbar_init(bar_tks(i), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_kvfull(s), 1);
      mbar_init(bar_kvempty(s), has_tail ? 3 : 1);  // MMA commit (+ the two helper warps, which read K rows)
      mbar_init(bar_sfull(s, 0), 1);
      mbar_init(bar_sfull(s, 1), 1);
      mbar_init(bar_pfull(s, 0), 4);
      mbar_init(bar_pfull(s, 1), 4);
      mbar_init(bar_dsfree(s), 1);
      mbar_init(bar_cfull(s), 4);
      mbar_init(bar_cfree(s), 2);
    }
    mbar_init(bar_dkvfull, 1);
    mbar_init(bar_dkvfree, 8);
    mbar_init(bar_dqfull, 1);
    mbar_init(bar_tkc, 8);
    fence_mbar_init();
    tma_prefetch_desc(&tmDQ);
    tma_prefetch_desc(&tmDK);
    tma_prefetch_desc(&tmDV);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  // group g, buffer b: S^T (kQU columns; then P^T as bf16 pairs in its first half) | dP^T (kQU columns; then dS^T in its first half)
  auto tS = [&](int g, uint32_t bsel) { return tmem + 128u * g + (128u / kUPT) * bsel; };
  auto tP = [&](int g, uint32_t bsel) { return tmem + 128u * g + (128u / kUPT) * bsel + kQU; };
  const uint32_t tDV = tmem + 256, tDK = tmem + 320;
  auto tDQ = [&](int i) { return tmem + 384u + 64u * i; };

  // Register budget: setmaxnreg moves registers inside the CTA's own allocation (384 threads x 168 at launch = 64512), so
  // 128 x R0 + 256 x R1 must not exceed 64512 (asking for more blocks forever): 128 x 88 + 256 x 208 = 64512.
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    if (warp == 0) {
      // ================================================================== TMA producer
      if (elect_one()) {
        // issue order = need order: tile 0's Q, the first K / V block, tile 0's dO and O (for D); then the second tile; then
        // the remaining K / V blocks through the ring.  O_0 lands in staging buffer 0 (chunk 0), O_1 in staging buffer 1 (chunk 1):
        // both are consumed before the first dS^T is written there.
        const int qr = static_cast<int>(qrow0), kr = static_cast<int>(krow0);
        auto load_tile = [&](int i) {
          const bool last = i == nqt - 1;
          mbar_expect_tx(bar_q(i), 16384 + ((last && has_tail) ? 2048 : 0) + ((i == 0 && has_tk) ? 4096 : 0));
          tma_load_2d(sQ + i * 16384, &tmQ, bar_q(i), h * kHD, qr + i * kT);
          if (i == 0 && has_tk) {
            tma_load_2d(sKt, &tmKt, bar_q(0), h * kHD, kr + p.nk_main);
            tma_load_2d(sVt, &tmVt, bar_q(0), h * kHD, kr + p.nk_main);
          }
          if (last && has_tail) tma_load_2d(sQt, &tmQt, bar_q(i), h * kHD, qr + p.nq_main);
        };
        auto load_do = [&](int i) {
          const bool last = i == nqt - 1;
          mbar_expect_tx(bar_do(i), 2 * 16384 + ((last && has_tail) ? 2048 : 0));
          tma_load_2d(sDO + i * 16384, &tmDO, bar_do(i), h * kHD, qr + i * kT);
          tma_load_2d(sST + i * (32768 + 16384), &tmO, bar_do(i), h * kHD, qr + i * kT);
          if (last && has_tail) tma_load_2d(sDOt, &tmDOt, bar_do(i), h * kHD, qr + p.nq_main);
        };
        if (dbg0) p.dbg[1024 + dbg_cta0 * 8 + 1] = clock64();  // producer starts issuing
        load_tile(0);
        for (int j = 0; j < nkblk; ++j) {
          const int s = j & 1;
          mbar_wait_trap(bar_kvempty(s), ((j >> 1) & 1) ^ 1);
          mbar_expect_tx(bar_kvfull(s), 32768);
          tma_load_2d(sK + s * 16384, &tmK, bar_kvfull(s), h * kHD, kr + j * kT);
          tma_load_2d(sV + s * 16384, &tmV, bar_kvfull(s), h * kHD, kr + j * kT);
          if (j == 0) {
            load_do(0);
            if (nqt > 1) {
              load_tile(1);
              load_do(1);
            }
            if (dbg0) p.dbg[1024 + dbg_cta0 * 8 + 2] = clock64();  // first batch of loads issued
          }
          if (j == min(1, nkblk - 1)) {
            // Measured: the first bytes of a CTA's loads arrive ~4.6 k cycles after they are issued (more than 1 GB of tensors behind
            // one CTA per SM: every new CTA starts on cold TLB entries and cold DRAM pages), 15 % of the CTA's life with nothing to
            // overlap it.  So every CTA pulls the inputs of the CTA one wave ahead of it into L2 while it computes itself.
            const int nb = static_cast<int>(blockIdx.x) + p.pf_dist;
            if (p.pf_dist > 0 && nb < static_cast<int>(gridDim.x)) {
              const int h2 = nb % p.H, b2 = nb / p.H;
              const int qr2 = b2 * p.nq + p.q0, kr2 = b2 * p.nk;
              for (int i = 0; i < nqt; ++i) {
                tma_prefetch_2d(&tmQ, h2 * kHD, qr2 + i * kT);
                tma_prefetch_2d(&tmDO, h2 * kHD, qr2 + i * kT);
                tma_prefetch_2d(&tmO, h2 * kHD, qr2 + i * kT);
              }
              for (int jj = 0; jj < nkblk; ++jj) {
                tma_prefetch_2d(&tmK, h2 * kHD, kr2 + jj * kT);
                tma_prefetch_2d(&tmV, h2 * kHD, kr2 + jj * kT);
              }
              if (has_tail) {
                tma_prefetch_2d(&tmQt, h2 * kHD, qr2 + p.nq_main);
                tma_prefetch_2d(&tmDOt, h2 * kHD, qr2 + p.nq_main);
              }
              if (has_tk) {
                tma_prefetch_2d(&tmKt, h2 * kHD, kr2 + p.nk_main);
                tma_prefetch_2d(&tmVt, h2 * kHD, kr2 + p.nk_main);
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ================================================================== MMA issuer
      if (elect_one()) {
        constexpr uint32_t kHi = 0x40004040u;    // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
        constexpr uint32_t kLoK = 1u << 16;      // K-major operands: LBO field unused
        constexpr uint32_t kLoMN = 1024u << 16;  // MN-major operands: LBO = 16 KB between 64-wide chunks
        const uint32_t q_k = (sQ >> 4) | kLoK, q_mn = (sQ >> 4) | kLoMN;      // + i * 1024 (tile) + g * 512 (64 rows) + h * 256 (32 rows)
        const uint32_t do_k = (sDO >> 4) | kLoK, do_mn = (sDO >> 4) | kLoMN;  // idem
        const uint32_t k_k = (sK >> 4) | kLoK, k_mn = (sK >> 4) | kLoMN;      // + s * 1024
        const uint32_t v_k = (sV >> 4) | kLoK;                                // + s * 1024
        const uint32_t st_mn = (sST >> 4) | kLoMN;                            // + buffer * 2048
        const uint32_t qt_k = (sQt >> 4) | kLoK, qt_mn = (sQt >> 4) | kLoMN, dot_k = (sDOt >> 4) | kLoK, dot_mn = (sDOt >> 4) | kLoMN;
        const uint32_t idesc_s = umma_idesc_bf16(kT, kQU, 0, 0);    // S^T / dP^T: A = K / V (K-major), B = Q / dO rows (K-major)
        const uint32_t idesc_st = umma_idesc_bf16(kT, 16, 0, 0);    // ... of the tail unit
        const uint32_t idesc_acc = umma_idesc_bf16(kT, kHD, 0, 1);  // dV / dK: A = P^T / dS^T in TMEM, B = dO / Q rows (MN-major)
        const uint32_t idesc_dq = umma_idesc_bf16(kT, kHD, 1, 1);   // dQ: A = dS^T staging (MN-major), B = K (MN-major)
        // Per group: the next unit whose S^T / dP^T get issued -- (key block, index inside the group's units of a block: 2 per
        // query tile, then group 0's tail unit) -- and how many were issued (TMEM buffer = count & 1).  S^T / dP^T run two units
        // ahead of the group's compute, so a group finds its next scores waiting when it hands a unit over.
        int sblk0 = 0, sblk1 = 0, sidx0 = 0, sidx1 = 0;
        uint32_t ns0 = 0, ns1 = 0;
        // Inputs of query tile i have landed.  Tail keys (the cls key of
        // 257 = 2 x 128 + 1) ride as N = 16 COLUMNS with the queries on the lanes: S = Q_i K_t^T, dP = dO_i V_t^T go into the still
        // unused dQ accumulator of the tile (read by the tile's compute group before the first dQ MMA can be issued).
        auto tile_ready = [&](int i) {
          mbar_wait_trap(bar_q(i), 0);
          if (dbg0 && i == 0) p.dbg[1024 + dbg_cta0 * 8 + 3] = clock64();  // Q_0 (+ tail keys) landed
          mbar_wait_trap(bar_do(i), 0);
          if (dbg0 && i == 0) p.dbg[1024 + dbg_cta0 * 8 + 4] = clock64();  // dO_0, O_0 landed
          tc_fence_after();
          if (has_tk) {
            const uint32_t kt_k = (sKt >> 4) | kLoK, vt_k = (sVt >> 4) | kLoK;
#pragma unroll
            for (int k = 0; k < kHD / 16; ++k) umma_ss_lohi(tDQ(i), q_k + i * 1024u + 2 * k, kt_k + 2 * k, kHi, idesc_st, k > 0);
#pragma unroll
            for (int k = 0; k < kHD / 16; ++k) umma_ss_lohi(tDQ(i) + 16u, do_k + i * 1024u + 2 * k, vt_k + 2 * k, kHi, idesc_st, k > 0);
            umma_commit(bar_tks(i));
          }
        };
        auto issue_next = [&](int g) {
          int& sblk = g ? sblk1 : sblk0;
          int& sidx = g ? sidx1 : sidx0;
          uint32_t& ns = g ? ns1 : ns0;
          if (sblk >= nkblk) return;
          const int ng = kUPT * nqt + ((g == 0 && has_tail) ? 1 : 0);
          const int s = sblk & 1;
          if (sidx == 0) {  // first unit of a key block for this group: K / V must have landed
            mbar_wait_trap(bar_kvfull(s), (sblk >> 1) & 1);
            tc_fence_after();
            if (dbg0 && sblk == 0 && g == 0) p.dbg[1024 + dbg_cta0 * 8 + 5] = clock64();  // K_0, V_0 landed
          }
          const bool tail = sidx == kUPT * nqt;
          const uint32_t rows = static_cast<uint32_t>(sidx / kUPT) * 1024u + static_cast<uint32_t>(g) * 512u + static_cast<uint32_t>(sidx % kUPT) * (kQU * 8u);
          const uint32_t ka = k_k + s * 1024u, va = v_k + s * 1024u;
          const uint32_t qb = tail ? qt_k : q_k + rows, gb = tail ? dot_k : do_k + rows;
          const uint32_t idesc = tail ? idesc_st : idesc_s;
          const uint32_t ts = tS(g, ns % kUPT), tp = tP(g, ns % kUPT);
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) umma_ss_lohi(ts, ka + 2 * k, qb + 2 * k, kHi, idesc, k > 0);
#pragma unroll
          for (int k = 0; k < kHD / 16; ++k) umma_ss_lohi(tp, va + 2 * k, gb + 2 * k, kHi, idesc, k > 0);
          umma_commit(bar_sfull(g, ns % kUPT));
          ++ns;
          if (++sidx == ng) {
            sidx = 0;
            ++sblk;
          }
        };
        const int dbg_cta = static_cast<int>(blockIdx.x) - 4 * static_cast<int>(gridDim.x) / 7;
        const bool dbg_on = p.dbg != nullptr && dbg_cta >= 0 && dbg_cta < 8;
        int dbg_n = 0;
#define VL_ISTAMP()                                                                         \
  do {                                                                                      \
    if (dbg_on && dbg_n < 63) p.dbg[512 + dbg_cta * 64 + (dbg_n++)] = clock64();           \
  } while (0)
        VL_ISTAMP();
        for (int i = 0; i < nqt; ++i) tile_ready(i);
#pragma unroll 1
        for (int n = 0; n < 2 * kUPT; ++n) issue_next(n & 1);  // kUPT units ahead for both groups
        VL_ISTAMP();
        uint32_t np0 = 0, np1 = 0;  // units handed over per group (pfull phases, TMEM buffer = count & 1)
        uint32_t ntile = 0;         // query tiles completed (staging buffer = ntile & 1)
        const int nunits = 2 * kUPT * nqt + (has_tail ? 1 : 0);  // per key block, in issue order: (tile, half, group) ..., then the tail unit
#pragma unroll 1
        for (int j = 0; j < nkblk; ++j) {
          const int s = j & 1;
          const int nkb = min(kT, ((p.nk_main - j * kT) + 15) & ~15);  // key rows of this block the dQ MMAs have to read
#pragma unroll 1
          for (int u = 0; u < nunits; ++u) {
            const bool tail = u == 2 * kUPT * nqt;
            const int i = tail ? 0 : u / (2 * kUPT), hq = tail ? 0 : (u >> 1) % kUPT, g = tail ? 0 : u & 1;
            uint32_t& np = g ? np1 : np0;
            const uint32_t bsel = np % kUPT;
            mbar_wait_trap(bar_pfull(g, bsel), (np / kUPT) & 1);
            ++np;
            if (u == 0 && j > 0) mbar_wait_trap(bar_dkvfree, (j - 1) & 1);  // dV / dK of the previous block drained
            tc_fence_after();
            VL_ISTAMP();
            // dV_j += P^T dO, dK_j += dS^T Q over the unit's queries (A operands straight from TMEM); the first MMA of a key block
            // overwrites the accumulators.  Tail unit: K = 16 queries (rows past tq are zero in P^T / dS^T).
            const uint32_t rows = static_cast<uint32_t>(i) * 1024u + static_cast<uint32_t>(g) * 512u + static_cast<uint32_t>(hq) * (kQU * 8u);
            const uint32_t gb = tail ? dot_mn : do_mn + rows, qb = tail ? qt_mn : q_mn + rows;
            const int nkk = tail ? 1 : kQU / 16;
#pragma unroll
            for (int kk = 0; kk < kQU / 16; ++kk)
              if (kk < nkk) umma_ts_lohi(tDV, tS(g, bsel) + 8 * kk, gb + 128 * kk, kHi, idesc_acc, (u == 0 && kk == 0) ? 0u : 1u);
#pragma unroll
            for (int kk = 0; kk < kQU / 16; ++kk)
              if (kk < nkk) umma_ts_lohi(tDK, tP(g, bsel) + 8 * kk, qb + 128 * kk, kHi, idesc_acc, (u == 0 && kk == 0) ? 0u : 1u);
            if (!tail && (u % (2 * kUPT)) == 2 * kUPT - 1) {
              // dQ_i += dS_i K_j over the tile's four units (dS^T staging buffer ntile & 1)
              const uint32_t da = st_mn + (ntile & 1u) * 2048u, kb = k_mn + s * 1024u;
#pragma unroll
              for (int kk = 0; kk < kT / 16; ++kk)
                if (kk < nkb / 16) umma_ss_lohi(tDQ(i), da + 128 * kk, kb + 128 * kk, kHi, idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
              umma_commit(bar_dsfree(ntile & 1u));
              ++ntile;
            }
            if (u == nunits - 1) {
              umma_commit(bar_dkvfull);     // dV_j, dK_j complete
              umma_commit(bar_kvempty(s));  // K_j, V_j smem reusable (a later block's S^T units only read the other stage)
            }
            // refill the TMEM buffer just consumed: queued behind the MMAs above (same thread: executed in issue order)
            issue_next(g);
            VL_ISTAMP();
          }
        }
        umma_commit(bar_dqfull);
#undef VL_ISTAMP
      }
    } else if (has_tail || has_tk) {
      // ================================================================== helper warps (2 warps; lane = head dims 2 lane, 2 lane + 1)
      // What is left on CUDA cores, off the compute groups' critical path:
      //  * tail keys: dV_u[d] = sum_q P[q, u] dO[q, d], dK_u[d] = sum_q dS[q, u] Q[q, d] over the range's queries (coefficients from the
      //    compute groups' prologue), plus the tail-query x tail-key corner;
      //  * tail queries: dQ_t[d] = sum over keys dS[t, key] K[key, d] (coefficients from compute group 0's tail unit).
      // The two warps split the rows of every mat-vec in halves and meet once at the end (partials through shared memory, added
      // in a fixed order: deterministic).
      const int hw = warp - 2;
      const int d0 = 2 * lane;
      const float sl2 = p.scale * kLog2e;
      float2 acc[kMaxTail], ak[kMaxTail], av[kMaxTail];  // dQ of the tail queries; dK / dV of the tail keys (this warp's share)
#pragma unroll
      for (int t = 0; t < kMaxTail; ++t) acc[t] = ak[t] = av[t] = make_float2(0.f, 0.f);
      if (has_tk) {
        for (int i = 0; i < nqt; ++i) {
          mbar_wait_trap(bar_q(i), 0);
          mbar_wait_trap(bar_do(i), 0);
        }
        mbar_wait_trap(bar_tkc, 0);
        const float* cp = sf + kFTk;                      // [u][q] P
        const float* cd = sf + kFTk + kMaxTail * 2 * kT;  // [u][q] dS
        if (hw < nqt) {  // warp hw takes query tile hw (rows past nq_main carry zero coefficients)
          const uint8_t* gt = bp + kOffDO + hw * 16384;
          const uint8_t* qt = bp + kOffQ + hw * 16384;
#pragma unroll 4
          for (int q = 0; q < kT; ++q) {
            const uint32_t off = sw128_off(q, d0);
            const uint32_t gw = *reinterpret_cast<const uint32_t*>(gt + off), qw = *reinterpret_cast<const uint32_t*>(qt + off);
            const float2 gv = make_float2(bf16_lo(gw), bf16_hi(gw)), qv = make_float2(bf16_lo(qw), bf16_hi(qw));
#pragma unroll
            for (int u = 0; u < kMaxTail; ++u) {
              if (u < p.tk) {
                const float cpv = cp[u * 2 * kT + hw * kT + q], cdv = cd[u * 2 * kT + hw * kT + q];
                av[u] = __ffma2_rn(make_float2(cpv, cpv), gv, av[u]);
                ak[u] = __ffma2_rn(make_float2(cdv, cdv), qv, ak[u]);
              }
            }
          }
        }
        if (hw == 0) {
          // tail query x tail key: the few scalar scores by warp reduction, then every lane updates its two dims
          for (int t = 0; t < p.tq; ++t) {
            const uint32_t qtw = *reinterpret_cast<const uint32_t*>(bp + kOffQt + sw128_off(t, d0));
            const uint32_t gtw = *reinterpret_cast<const uint32_t*>(bp + kOffDOt + sw128_off(t, d0));
#pragma unroll
            for (int u = 0; u < kMaxTail; ++u) {
              if (u >= p.tk) break;
              const uint32_t kuw = *reinterpret_cast<const uint32_t*>(bp + kOffKt + sw128_off(u, d0));
              const uint32_t vuw = *reinterpret_cast<const uint32_t*>(bp + kOffVt + sw128_off(u, d0));
              float sd = bf16_lo(qtw) * bf16_lo(kuw) + bf16_hi(qtw) * bf16_hi(kuw);
              float dp = bf16_lo(gtw) * bf16_lo(vuw) + bf16_hi(gtw) * bf16_hi(vuw);
#pragma unroll
              for (int o2 = 16; o2 > 0; o2 >>= 1) {
                sd += __shfl_xor_sync(0xffffffffu, sd, o2);
                dp += __shfl_xor_sync(0xffffffffu, dp, o2);
              }
              const float pv = ex2_approx(fmaf(sd, sl2, sf[kFNl + 2 * kT + t]));
              const float ds = pv * fmaf(dp, p.scale, sf[kFNd + 2 * kT + t]);
              av[u] = __ffma2_rn(make_float2(pv, pv), make_float2(bf16_lo(gtw), bf16_hi(gtw)), av[u]);
              ak[u] = __ffma2_rn(make_float2(ds, ds), make_float2(bf16_lo(qtw), bf16_hi(qtw)), ak[u]);
#pragma unroll
              for (int t2 = 0; t2 < kMaxTail; ++t2)
                if (t2 == t) acc[t2] = __ffma2_rn(make_float2(ds, ds), make_float2(bf16_lo(kuw), bf16_hi(kuw)), acc[t2]);
            }
          }
        }
      }
      if (has_tail) {
        for (int j = 0; j < nkblk; ++j) {
          const int s = j & 1, cb = j & 1;
          mbar_wait_trap(bar_kvfull(s), (j >> 1) & 1);
          mbar_wait_trap(bar_cfull(cb), (j >> 1) & 1);
          const uint8_t* kt = bp + kOffK + s * 16384;
          const float* cf = sf + kFCoef + cb * kMaxTail * kT;
          const int r1 = min(kT / 2 * (hw + 1), p.nk_main - j * kT);  // warp hw takes key rows [64 hw, 64 hw + 64) of the block
#pragma unroll 4
          for (int r = kT / 2 * hw; r < r1; ++r) {
            const uint32_t kw = *reinterpret_cast<const uint32_t*>(kt + sw128_off(r, d0));
            const float2 kv = make_float2(bf16_lo(kw), bf16_hi(kw));
#pragma unroll
            for (int t = 0; t < kMaxTail; ++t) {
              if (t < p.tq) {
                const float c = cf[t * kT + r];
                acc[t] = __ffma2_rn(make_float2(c, c), kv, acc[t]);
              }
            }
          }
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_cfree(cb));
            mbar_arrive(bar_kvempty(s));
          }
        }
      }
      // the two warps' partial sums meet: warp 1 hands its registers over through shared memory (the coefficient area of the tail
      // unit is dead by now for the tail keys' part; a private slice is used to stay clear of it)
      float2* xch = reinterpret_cast<float2*>(sf + kFXch);  // [3][kMaxTail][32] float2
      if (hw == 1) {
#pragma unroll
        for (int t = 0; t < kMaxTail; ++t) {
          xch[(0 * kMaxTail + t) * 32 + lane] = acc[t];
          xch[(1 * kMaxTail + t) * 32 + lane] = ak[t];
          xch[(2 * kMaxTail + t) * 32 + lane] = av[t];
        }
      }
      asm volatile("bar.sync 4, 64;" ::: "memory");
      if (hw == 0) {
#pragma unroll
        for (int t = 0; t < kMaxTail; ++t) {
          if (t < p.tq) {
            const float2 o = xch[(0 * kMaxTail + t) * 32 + lane];
            *reinterpret_cast<uint32_t*>(p.dq + (qrow0 + p.nq_main + t) * p.lddq + h * kHD + d0) = pack_bf16(acc[t].x + o.x, acc[t].y + o.y);
          }
          if (t < p.tk) {
            const float2 ok_ = xch[(1 * kMaxTail + t) * 32 + lane], ov_ = xch[(2 * kMaxTail + t) * 32 + lane];
            __nv_bfloat16* dkp = p.dk + (krow0 + p.nk_main + t) * p.lddk + h * kHD + d0;
            __nv_bfloat16* dvp = p.dv + (krow0 + p.nk_main + t) * p.lddv + h * kHD + d0;
            float k0 = ak[t].x + ok_.x, k1 = ak[t].y + ok_.y, v0 = av[t].x + ov_.x, v1 = av[t].y + ov_.y;
            if (p.accum_kv) {
              const uint32_t okw = *reinterpret_cast<const uint32_t*>(dkp), ovw = *reinterpret_cast<const uint32_t*>(dvp);
              k0 += bf16_lo(okw); k1 += bf16_hi(okw); v0 += bf16_lo(ovw); v1 += bf16_hi(ovw);
            }
            *reinterpret_cast<uint32_t*>(dkp) = pack_bf16(k0, k1);
            *reinterpret_cast<uint32_t*>(dvp) = pack_bf16(v0, v1);
          }
        }
      }
    }
  } else {
    // ================================================================== compute groups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int g = (warp - 4) >> 2;      // 0 or 1: query half [64 g, 64 g + 64) of every tile
    const int quarter = warp & 3;       // TMEM lane quarter of this warp
    const int r = quarter * 32 + lane;  // key row inside the block = TMEM lane
    const int x = r;                    // thread index inside the group
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = p.scale * kLog2e;
    float* sNl = sf + kFNl;
    float* sNd = sf + kFNd;
    const int dbg_cta = static_cast<int>(blockIdx.x) - 4 * static_cast<int>(gridDim.x) / 7;  // a CTA of a later wave (warm caches)
    const bool dbg_on = p.dbg != nullptr && dbg_cta >= 0 && dbg_cta < 8 && x == 0;
    int dbg_n = 0;
#define VL_STAMP()                                                                     \
  do {                                                                                 \
    if (dbg_on && dbg_n < 31) p.dbg[dbg_cta * 64 + g * 32 + (dbg_n++)] = clock64();    \
  } while (0)
    VL_STAMP();

    // ---- per-query statistics  nl = -lse log2 e,  nd = -D scale  with D = rowsum(dO * O): group g's threads own the rows of query
    // tile g (group 0 also the tail queries).  (Tried: group 1 computing tile 1's statistics after its first unit, so that the first
    // tile's units start before the second tile has landed -- measured slower, 0.49 vs 0.46 ms: the extra two-group barrier and
    // the delayed first hand-over of group 1 cost more than the ~1 k cycles the earlier start saves.)
    float dsk[kMaxTail] = {0.f, 0.f, 0.f, 0.f};  // dS of this thread's query row against the tail keys (for its dQ row)
    auto tile_stats = [&]() {
      const int qrow = g * kT + x;  // row inside this launch's main range
      const bool ok = qrow < p.nq_main;
      float lse = 0.f;
      if (ok) lse = p.lse[(static_cast<long long>(b) * p.H + h) * p.nq + p.q0 + qrow];
      if (warp == 4) {  // tail queries: one warp, two dims per lane, coalesced 128-byte rows
        for (int t = 0; t < 16; ++t) {
          float nl = -INFINITY, nd = 0.f;
          if (t < p.tq) {
            const long long grow = qrow0 + p.nq_main + t;
            const uint32_t ow = *reinterpret_cast<const uint32_t*>(p.o + grow * p.ldo + h * kHD + 2 * lane);
            const uint32_t gw = *reinterpret_cast<const uint32_t*>(p.dout + grow * p.lddo + h * kHD + 2 * lane);
            float dsum = bf16_lo(ow) * bf16_lo(gw) + bf16_hi(ow) * bf16_hi(gw);
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o2);
            nl = -p.lse[(static_cast<long long>(b) * p.H + h) * p.nq + p.q0 + p.nq_main + t] * kLog2e;
            nd = -dsum * p.scale;
          }
          if (lane == 0) {
            sNl[2 * kT + t] = nl;
            sNd[2 * kT + t] = nd;
          }
        }
      }
      mbar_wait_trap(bar_do(g), 0);
      const float D = ok ? dot_rows_sw128(bp + kOffDO + g * 16384, bp + kOffST + g * (32768 + 16384), x) : 0.f;
      sNl[qrow] = ok ? -lse * kLog2e : -INFINITY;  // rows that do not exist: exp2(s - inf) = 0 -> P = dS = 0
      sNd[qrow] = -D * p.scale;
      if (has_tk) {
        // tail keys: this thread's query row against the <= 4 tail keys (scores sit in the still unused dQ accumulator of the
        // row's tile).  P / dS go to the helper warps (dV / dK of the tail keys); dS stays here for this row's dQ.
        float* cp = sf + kFTk;
        float* cd = sf + kFTk + kMaxTail * 2 * kT;
        mbar_wait_trap(bar_tks(g), 0);
        tc_fence_after();
        uint32_t sv[16], dpv[16];
        tmem_ld16(tDQ(g) + lane_off, sv);
        tmem_ld16(tDQ(g) + 16u + lane_off, dpv);
        tc_wait_ld();
#pragma unroll
        for (int u = 0; u < kMaxTail; ++u) {
          const float pv = (ok && u < p.tk) ? ex2_approx(fmaf(__uint_as_float(sv[u]), sl2, -lse * kLog2e)) : 0.f;
          const float ds = pv * fmaf(__uint_as_float(dpv[u]), p.scale, -D * p.scale);
          dsk[u] = ds;
          cp[u * 2 * kT + qrow] = pv;
          cd[u * 2 * kT + qrow] = ds;
        }
        tc_fence_before();
      }
    };
    if (g < nqt) tile_stats();
    if (has_tk) {
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tkc);
    }
    both_groups_sync();  // statistics visible; the O tiles (staging buffers) are dead from here on
    VL_STAMP();

    uint32_t cs = 0;        // units received (TMEM buffer = cs & 1, sfull phase = (cs >> 1) & 1)
    uint32_t ntile = 0;     // query tiles done (staging buffer = ntile & 1)
    int store_buf = -1;     // staging buffer a bulk store of this group may still be reading
    for (int j = 0; j < nkblk; ++j) {
      const int krow = j * kT + r;
      const bool krow_ok = krow < p.nk_main;
      const bool warp_on = j * kT + quarter * 32 < p.nk_main;  // this warp owns at least one existing key row
      for (int i = 0; i < nqt; ++i) {
        const uint32_t sb = ntile & 1u;
        uint8_t* chunk = bp + kOffST + sb * 32768 + g * 16384;
#pragma unroll 1
        for (int hq = 0; hq < kUPT; ++hq) {
          const uint32_t tb = cs % kUPT;
          mbar_wait_trap(bar_sfull(g, tb), (cs / kUPT) & 1);
          ++cs;
          tc_fence_after();
          if (hq == 0) {
            if (ntile >= 2) mbar_wait_trap(bar_dsfree(sb), ((ntile >> 1) - 1) & 1);  // dQ MMAs of tile ntile - 2 have read the buffer
            if (store_buf == static_cast<int>(sb)) {                          // ... and so has this group's last bulk store
              if (x == 0) tma_store_wait_read<0>();
              group_sync(g);
              store_buf = -1;
            }
          }
          if (warp_on) {
            const uint32_t tSg = tS(g, tb) + lane_off, tPg = tP(g, tb) + lane_off;
            const float* nlp = sNl + i * kT + g * kQH + hq * kQU;
            const float* ndp = sNd + i * kT + g * kQH + hq * kQU;
            const float2 sl22 = make_float2(sl2, sl2), sc2 = make_float2(p.scale, p.scale);
#pragma unroll
            for (int cc = 0; cc < kQU / 32; ++cc) {
              uint32_t sv[32], dpv[32];
              tmem_ld32(tSg + 32 * cc, sv);
              tmem_ld32(tPg + 32 * cc, dpv);
              tc_wait_ld();
              uint32_t pw[16], dw[16];
              if (krow_ok) {
#pragma unroll
                for (int t = 0; t < 32; t += 4) {
                  const float4 l4 = *reinterpret_cast<const float4*>(nlp + 32 * cc + t);
                  const float4 d4 = *reinterpret_cast<const float4*>(ndp + 32 * cc + t);
                  const float2 a0 = __ffma2_rn(make_float2(__uint_as_float(sv[t]), __uint_as_float(sv[t + 1])), sl22, make_float2(l4.x, l4.y));
                  const float2 a1 = __ffma2_rn(make_float2(__uint_as_float(sv[t + 2]), __uint_as_float(sv[t + 3])), sl22, make_float2(l4.z, l4.w));
                  const float2 p0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y)), p1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
                  const float2 u0 = __ffma2_rn(make_float2(__uint_as_float(dpv[t]), __uint_as_float(dpv[t + 1])), sc2, make_float2(d4.x, d4.y));
                  const float2 u1 = __ffma2_rn(make_float2(__uint_as_float(dpv[t + 2]), __uint_as_float(dpv[t + 3])), sc2, make_float2(d4.z, d4.w));
                  const float2 e0 = __fmul2_rn(p0, u0), e1 = __fmul2_rn(p1, u1);
                  pw[t >> 1] = pack_bf16(p0.x, p0.y);
                  pw[(t >> 1) + 1] = pack_bf16(p1.x, p1.y);
                  dw[t >> 1] = pack_bf16(e0.x, e0.y);
                  dw[(t >> 1) + 1] = pack_bf16(e1.x, e1.y);
                }
              } else {
#pragma unroll
                for (int t = 0; t < 16; ++t) pw[t] = dw[t] = 0u;
              }
              // P^T / dS^T back into TMEM over their own inputs: key row in its lane, queries (2c, 2c + 1) in 32-bit column c
              // (columns 16 cc .. 16 cc + 15 were read one chunk ago at the latest)
              tmem_st16(tSg + 16 * cc, pw);
              tmem_st16(tPg + 16 * cc, dw);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                *reinterpret_cast<uint4*>(chunk + sw128_off(r, kQU * hq + 32 * cc + 8 * u)) = make_uint4(dw[4 * u], dw[4 * u + 1], dw[4 * u + 2], dw[4 * u + 3]);
            }
            tc_wait_st();
            fence_proxy_async_smem();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_pfull(g, tb));
          VL_STAMP();
        }
        ++ntile;
      }
      if (has_tail && g == 0) {
        // tail unit: N = 16 query columns, the first tq exist
        const int cb = j & 1;
        const uint32_t tb = cs % kUPT;
        mbar_wait_trap(bar_sfull(0, tb), (cs / kUPT) & 1);
        ++cs;
        tc_fence_after();
        if (j >= 2) mbar_wait_trap(bar_cfree(cb), ((j >> 1) - 1) & 1);
        float* cf = sf + kFCoef + cb * kMaxTail * kT;
        if (warp_on) {
          const uint32_t tSg = tS(0, tb) + lane_off, tPg = tP(0, tb) + lane_off;
          uint32_t sv[16], dpv[16];
          tmem_ld16(tSg, sv);
          tmem_ld16(tPg, dpv);
          tc_wait_ld();
          float pt[kMaxTail], dt[kMaxTail];
#pragma unroll
          for (int t = 0; t < kMaxTail; ++t) {
            const float pv = krow_ok ? ex2_approx(fmaf(__uint_as_float(sv[t]), sl2, sNl[2 * kT + t])) : 0.f;  // nl = -inf past tq
            pt[t] = pv;
            dt[t] = pv * fmaf(__uint_as_float(dpv[t]), p.scale, sNd[2 * kT + t]);
            cf[t * kT + r] = dt[t];
          }
          const uint32_t pw[8] = {pack_bf16(pt[0], pt[1]), pack_bf16(pt[2], pt[3]), 0u, 0u, 0u, 0u, 0u, 0u};
          const uint32_t dw[8] = {pack_bf16(dt[0], dt[1]), pack_bf16(dt[2], dt[3]), 0u, 0u, 0u, 0u, 0u, 0u};
          tmem_st8(tSg, pw);
          tmem_st8(tPg, dw);
          tc_wait_st();
        } else {
#pragma unroll
          for (int t = 0; t < kMaxTail; ++t) cf[t * kT + r] = 0.f;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_pfull(0, tb));
          mbar_arrive(bar_cfull(cb));
        }
      }
      // ---- dV_j (group 0) / dK_j (group 1) -> global; thread r owns key row j * 128 + r
      mbar_wait_trap(bar_dkvfull, j & 1);
      tc_fence_after();
      VL_STAMP();
      {
        const int db = static_cast<int>((ntile - 1) & 1u);  // the staging buffer of the block's last tile: idle until tile ntile + 1
        uint8_t* stage = bp + kOffST + db * 32768 + g * 16384;
        const long long grow = krow0 + krow;
        __nv_bfloat16* dst = g == 0 ? p.dv + grow * p.lddv + h * kHD : p.dk + grow * p.lddk + h * kHD;
        const uint32_t t0 = (g == 0 ? tDV : tDK) + lane_off;
        if (warp_on) {
#pragma unroll
          for (int c = 0; c < kHD; c += 32) {
            uint32_t v[32];
            tmem_ld32(t0 + c, v);
            tc_wait_ld();
            if (p.accum_kv) {
              if (krow_ok) store_row32(dst + c, v, true);
            } else {
              stage_row32(stage, r, c, v);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dkvfree);
        if (!p.accum_kv) {  // one bulk tensor store per group: whole 128-byte lines, rows past nk clipped by the map
          fence_proxy_async_smem();
          group_sync(g);
          if (x == 0) {
            tma_store_3d(g == 0 ? &tmDV : &tmDK, sST + db * 32768 + g * 16384, h * kHD, j * kT, b);
            tma_store_commit();
          }
          store_buf = db;
        }
      }
      VL_STAMP();
    }

    // ---- dQ tile g -> global (thread = query row of tile g)
    if (g < nqt) {
      mbar_wait_trap(bar_dqfull, 0);
      tc_fence_after();
      VL_STAMP();
      if (store_buf >= 0) {
        if (x == 0) tma_store_wait_read<0>();
        group_sync(g);
        store_buf = -1;
      }
      uint8_t* stage = bp + kOffST + g * 16384;
#pragma unroll
      for (int c = 0; c < kHD; c += 32) {
        uint32_t v[32];
        tmem_ld32(tDQ(g) + lane_off + c, v);
        tc_wait_ld();
        if (has_tk) {  // dQ[q, :] += sum over tail keys dS[q, u] K_u
#pragma unroll
          for (int u = 0; u < kMaxTail; ++u) {
            if (u < p.tk) {
#pragma unroll
              for (int e = 0; e < 32; e += 8) {
                const uint4 kw = *reinterpret_cast<const uint4*>(bp + kOffKt + sw128_off(u, c + e));
                v[e] = __float_as_uint(fmaf(dsk[u], bf16_lo(kw.x), __uint_as_float(v[e])));
                v[e + 1] = __float_as_uint(fmaf(dsk[u], bf16_hi(kw.x), __uint_as_float(v[e + 1])));
                v[e + 2] = __float_as_uint(fmaf(dsk[u], bf16_lo(kw.y), __uint_as_float(v[e + 2])));
                v[e + 3] = __float_as_uint(fmaf(dsk[u], bf16_hi(kw.y), __uint_as_float(v[e + 3])));
                v[e + 4] = __float_as_uint(fmaf(dsk[u], bf16_lo(kw.z), __uint_as_float(v[e + 4])));
                v[e + 5] = __float_as_uint(fmaf(dsk[u], bf16_hi(kw.z), __uint_as_float(v[e + 5])));
                v[e + 6] = __float_as_uint(fmaf(dsk[u], bf16_lo(kw.w), __uint_as_float(v[e + 6])));
                v[e + 7] = __float_as_uint(fmaf(dsk[u], bf16_hi(kw.w), __uint_as_float(v[e + 7])));
              }
            }
          }
        }
        stage_row32(stage, r, c, v);
      }
      fence_proxy_async_smem();
      group_sync(g);
      if (x == 0) {  // rows past the last main query are clipped by the tensor map
        tma_store_3d(&tmDQ, sST + g * 16384, h * kHD, p.q0 + g * kT, b);
        tma_store_commit();
      }
    }
    if (x == 0) tma_store_wait_read<0>();  // the bulk stores have read their staging tiles before the CTA (and its shared memory) retires
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

}  // namespace bwd3

// Host entry used by vl_attention_bwd (attention.cu) for non-causal attention.  Query ranges of at most 256 (+ tail) rows per launch.
int launch_attn_bwd3(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                     int B, int H, int nq, int nk, long long ldq, long long ldk, long long ldv, long long ldo, long long lddo, long long lddq,
                     long long lddk, long long lddv, float scale, cudaStream_t stream) {
  using namespace bwd3;
  CUtensorMap tmQ, tmK, tmV, tmDO, tmO, tmQt, tmDOt, tmKt, tmVt;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmQ, q, (uint64_t)H * kHD, (uint64_t)B * nq, ldq, kHD, kT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, k, (uint64_t)H * kHD, (uint64_t)B * nk, ldk, kHD, kT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, v, (uint64_t)H * kHD, (uint64_t)B * nk, ldv, kHD, kT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDO, dout, (uint64_t)H * kHD, (uint64_t)B * nq, lddo, kHD, kT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmO, o, (uint64_t)H * kHD, (uint64_t)B * nq, ldo, kHD, kT))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmQt, q, (uint64_t)H * kHD, (uint64_t)B * nq, ldq, kHD, 16))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDOt, dout, (uint64_t)H * kHD, (uint64_t)B * nq, lddo, kHD, 16))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmKt, k, (uint64_t)H * kHD, (uint64_t)B * nk, ldk, kHD, 16))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmVt, v, (uint64_t)H * kHD, (uint64_t)B * nk, ldv, kHD, 16))) return rc;
  auto tail = [&](int n) {
    const int t = n % kT;
    return (n > kT && t > 0 && t <= kMaxTail) ? t : 0;
  };
  const int tq_all = tail(nq), tk = tail(nk);
  const int nq_main_all = nq - tq_all;
  // outputs as [cols, rows of one batch element, batch]: the row box is clipped per batch element, so partial tiles never spill
  // into the tail rows (written separately) or the next batch element
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
    VL_CUDA(cudaFuncSetAttribute(attn_bwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr = true;
  }
  Params p;
  p.B = B; p.H = H; p.nq = nq; p.nk = nk;
  p.scale = scale;
  p.nk_main = nk - tk; p.tk = tk;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k); p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.o = reinterpret_cast<const __nv_bfloat16*>(o); p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lddo = lddo;
  p.lse = lse;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.dbg = debug_buffer();
  // debug knob 15: L2 prefetch distance in CTAs (0 = default: one wave = the number of SMs, -1 = off)
  const int knob = debug_get(15);
  p.pf_dist = knob < 0 ? 0 : (knob > 0 ? knob : num_sms());
  for (int q0 = 0; q0 < nq_main_all; q0 += 2 * kT) {
    p.q0 = q0;
    p.nq_main = nq_main_all - q0 < 2 * kT ? nq_main_all - q0 : 2 * kT;
    p.tq = (q0 + 2 * kT >= nq_main_all) ? tq_all : 0;  // the tail rides with the last range
    p.accum_kv = q0 > 0;
    attn_bwd3_kernel<<<(unsigned)(B * H), kThreads, kSmem, stream>>>(tmQ, tmK, tmV, tmDO, tmO, tmQt, tmDOt, tmKt, tmVt, tmDQ, tmDK, tmDV, p);
    if (int rc2 = launch_check("attn_bwd3_kernel")) return rc2;
  }
  return 0;
}

}  // namespace vl
