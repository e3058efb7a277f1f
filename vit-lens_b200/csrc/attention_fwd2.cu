// FlashAttention forward for sm_100a, head_dim = 64 -- persistent two-group version for short, non-causal sequences
// (ViT-L/14: N = 257 = 2 x 128 + cls; Lens self-attention over 256 latents).
//
// One CTA per SM loops over work items (batch, head, pair of 128-query tiles).  512 threads:
//   warp 0       TMA producer: K / V 64-key blocks through a 5-stage ring that runs ahead into the next item -- no
//                per-item load bubble;  warp 3: TMA producer of the items' two Q tiles (double-buffered across items)
//   warps 1-2    tcgen05 issuers, one per group: S_g = Q_g K_j^T (N = 64) into a double-buffered TMEM region, then
//                O_g += P_g V_j; S of block j+1 is issued before P V of block j, so the tensor pipe never waits on softmax
//   warps 4-7    softmax group 0 (query tile 0), warps 8-11 group 1 (query tile 1): one query row per thread; the
//                64 scores of a block are read from TMEM ONCE into registers (the S buffer is released at once),
//                online max / sum in fp32 with a lazily moved offset, P (bf16) into a double-buffered swizzled
//                staging block.  The O accumulators are double-buffered in TMEM as well, so the epilogue of item i
//                (O / l -> bf16 -> global, LSE) runs after the first block of item i+1 has been handed to the MMA warp.
//   warps 12-15  tail rows: sequence lengths that leave 1..4 query rows past the last 128-row tile (257 = 256 + cls)
//                are not padded to a third tensor-core tile; these four warps compute those rows on CUDA cores from the
//                K / V blocks already in shared memory (each warp keeps its own online-softmax state over its quarter
//                of the keys, merged once per item) -- no second kernel re-reading K / V.
// A partial last key block (257 = 4 x 64 + 1) is a 16-key TMA box and an N = 16 MMA, masked in the softmax.
// TMEM (512 columns): S[group][2] 4 x 64 | O[group][2] 4 x 64.
// Scores never touch HBM: bytes moved = Q + K + V + O (+ LSE), the algorithmic minimum of SURVEY 8(d).
#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {
namespace fwd2 {

constexpr int kHD = 64;
constexpr int kTQ = 128;  // query tile (one per softmax group)
constexpr int kBK = 64;   // key block
constexpr int kNS = 9;    // K / V ring stages
constexpr int kThreads = 512;
constexpr int kMaxTail = 1;  // nq = 128 k + 1 (ViT: patches + cls); other remainders run as a padded tile
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// shared memory map (bytes, from a 1024-aligned base)
constexpr int kOffQ = 0;                       // [2 item stages][2 tiles] x 16 KB
constexpr int kOffKV = kOffQ + 4 * 16384;      // kNS x (K 8 KB | V 8 KB)
constexpr int kOffF = kOffKV + kNS * 16384;    // fp32 scratch of the tail warps
constexpr int kFQ = 0;                 // [4 warps][64]   q of the tail row (each tail warp its own copy)
constexpr int kFMerge = kFQ + 4 * kHD;   // [4 warps][66]   (offset, sum, out[64]) per tail warp
constexpr int kFEnd = kFMerge + 4 * 66;
constexpr int kOffBar = kOffF + ((kFEnd * 4 + 15) & ~15);
// barrier slots (8 bytes each)
constexpr int kBarQFull = 0, kBarQEmpty = 2, kBarKVFull = 4, kBarKVEmpty = kBarKVFull + kNS, kBarGrp = kBarKVEmpty + kNS;
// per (group, buffer): s_full, s_free, p_full, p_free, o_full, o_free
constexpr int kBarEnd = kBarGrp + 6 * 4;
constexpr int kSmem = kOffBar + kBarEnd * 8 + 16 + 1024;  // + TMEM slot + alignment slack
static_assert(kSmem <= 232448, "shared memory budget");

struct Params {
  int B, H, nq, nk;
  int nq_main, tq;  // rows [0, nq_main) on tensor cores, [nq_main, nq_main + tq) on CUDA cores
  int npairs;       // 256-row query ranges per (b, h)
  int nblk;         // 64-key blocks
  int nkb_last;     // MMA width of the last key block (multiple of 16)
  int valid_last;   // valid keys in the last block
  int n_items;
  float scale;
  const __nv_bfloat16* q;
  long long ldq;
  __nv_bfloat16* o;
  long long ldo;
  float* lse;  // [B, H, nq]
  long long* dbg;  // optional clock64 timeline (bring-up only): [4 CTAs][4 roles][64]
};

struct Item {
  int b, h, qp, nt;
};
__device__ __forceinline__ Item decode(const Params& p, int item) {
  Item it;
  it.qp = item % p.npairs;
  const int bh = item / p.npairs;
  it.h = bh % p.H;
  it.b = bh / p.H;
  it.nt = (p.nq_main - it.qp * 2 * kTQ) > kTQ ? 2 : 1;
  return it;
}

__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return static_cast<uint32_t>(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}
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

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmK16, const __grid_constant__ CUtensorMap tmV16, const __grid_constant__ CUtensorMap tmO,
                 const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + kOffBar;
  auto bar = [&](int slot) { return bars + 8u * static_cast<uint32_t>(slot); };
  auto bar_grp = [&](int kind, int g, int buf) { return bars + 8u * static_cast<uint32_t>(kBarGrp + kind * 4 + g * 2 + buf); };
  enum { S_FULL = 0, S_FREE = 1, P_FULL = 2, P_FREE = 3, O_FULL = 4, O_FREE = 5 };
  const uint32_t tmem_slot = bars + 8u * kBarEnd;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(bp + kOffBar + 8 * kBarEnd);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_local = (static_cast<int>(blockIdx.x) < p.n_items) ? (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  // bring-up timeline (-DVL_FWD2_TIMELINE): lane 0 of one warp per role stamps clock64 from its 4th item on (steady state)
#ifdef VL_FWD2_TIMELINE
  int dbg_n = 0;
#define VL_STAMP(role, n)                                                                                   \
  do {                                                                                                      \
    if (p.dbg != nullptr && blockIdx.x < 4 && lane == 0 && (n) >= 3 && dbg_n < 64) p.dbg[(blockIdx.x * 4 + (role)) * 64 + (dbg_n++)] = clock64(); \
  } while (0)
#else
#define VL_STAMP(role, n) \
  do {                    \
  } while (0)
#endif
  auto item_at = [&](int n) { return decode(p, static_cast<int>(blockIdx.x) + n * static_cast<int>(gridDim.x)); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmK16);
    tma_prefetch_desc(&tmV16);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kBarQFull + s), 1);
      mbar_init(bar(kBarQEmpty + s), 8);  // one arrival per softmax warp, after its O store has read its slab of the tile (the issuer stands in, x4, for a group without a tile)
    }
    for (int s = 0; s < kNS; ++s) {
      mbar_init(bar(kBarKVFull + s), 1);
      mbar_init(bar(kBarKVEmpty + s), 2 + 4);  // the two MMA issuers + the four tail warps
    }
    for (int g = 0; g < 2; ++g)
      for (int b2 = 0; b2 < 2; ++b2) {
        mbar_init(bar_grp(S_FULL, g, b2), 1);
        mbar_init(bar_grp(P_FULL, g, b2), 4);
        mbar_init(bar_grp(P_FREE, g, b2), 1);
        mbar_init(bar_grp(O_FULL, g, b2), 1);
        mbar_init(bar_grp(O_FREE, g, b2), 4);
      }
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
  auto tS = [&](int g, int buf) { return tmem + 64u * static_cast<uint32_t>(g * 2 + buf); };
  auto tO = [&](int g, int buf) { return tmem + 256u + 64u * static_cast<uint32_t>(g * 2 + buf); };
  auto sQ = [&](int qs, int g) { return base + kOffQ + static_cast<uint32_t>(qs * 2 + g) * 16384u; };
  auto sK = [&](int st) { return base + kOffKV + static_cast<uint32_t>(st) * 16384u; };

  // Register budget per warpgroup (512 threads x 128 at launch): the softmax groups hold a block's 64 scores, its packed P
  // and the deferred epilogue state per thread; producer / issuers / tail warps need little.  48 + 80 + 2 x 192 = 512.

  if (warp < 4) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
   if (warp == 0) {
    // ================================================================== TMA producer: K / V ring
    if (elect_one()) {
      int st = 0;
      uint32_t st_ph = 0;  // phase of the ring pass
      for (int n = 0; n < n_local; ++n) {
        const Item it = item_at(n);
        for (int j = 0; j < p.nblk; ++j) {
          mbar_wait(bar(kBarKVEmpty + st), st_ph ^ 1);
          VL_STAMP(0, n);  // stage free
          const bool small = (j == p.nblk - 1) && p.nkb_last == 16;
          mbar_expect_tx(bar(kBarKVFull + st), small ? 4096 : 16384);
          const int row = it.b * p.nk + j * kBK;
          tma_load_2d(sK(st), small ? &tmK16 : &tmK, bar(kBarKVFull + st), it.h * kHD, row);
          tma_load_2d(sK(st) + 8192, small ? &tmV16 : &tmV, bar(kBarKVFull + st), it.h * kHD, row);
          if (++st == kNS) {
            st = 0;
            st_ph ^= 1;
          }
        }
      }
    }
   } else if (warp == 3) {
    // ================================================================== TMA producer: Q tiles (own warp: a Q stage comes back
    // from the softmax groups only after their O store, which must not hold up the K / V prefetch)
    if (elect_one()) {
      for (int n = 0; n < n_local; ++n) {
        const Item it = item_at(n);
        const int qs = n & 1;
        mbar_wait(bar(kBarQEmpty + qs), ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(bar(kBarQFull + qs), it.nt * 16384);
        for (int g = 0; g < it.nt; ++g)
          tma_load_2d(sQ(qs, g), &tmQ, bar(kBarQFull + qs), it.h * kHD, it.b * p.nq + it.qp * 2 * kTQ + g * kTQ);
      }
    }
   } else if (warp == 1 || warp == 2) {
    // ================================================================== MMA issuers (one warp per softmax group)
    // One elected thread per group issues S_g(c) and then P_g V(c-1): the tensor pipe always has the next block's scores
    // ready when the group finishes a block.  Two issuers keep the single-thread instruction stream per block short (at
    // N = 64 the issue loop, not the tensor pipe, would otherwise set the pace).  Descriptors are built once: only the
    // 14-bit start-address field (low word) changes between operands, by multiples of 16 bytes.
    const int g = warp - 1;
    if (elect_one()) {
      constexpr uint32_t kDescHi = 0x40004040u;  // SBO = 1024 B, descriptor version 1, SWIZZLE_128B
      const uint32_t q_lo = ((base + kOffQ + g * 16384) >> 4) | (1u << 16);       // K-major operands: LBO field = 1 (unused)
      const uint32_t k_lo = ((base + kOffKV) >> 4) | (1u << 16);
      const uint32_t v_lo = ((base + kOffKV + 8192) >> 4) | (1024u << 16);        // MN-major operand: LBO = 16 KB (single 64-wide chunk: unused)
      const uint32_t idesc_o = umma_idesc_bf16(kTQ, kHD, 0, 1);
      const uint32_t idesc_s64 = umma_idesc_bf16(kTQ, kBK, 0, 0);
      const uint32_t idesc_slast = umma_idesc_bf16(kTQ, p.nkb_last, 0, 0);
      const int kk_last = p.nkb_last >> 4;
      const int qp_step = static_cast<int>(gridDim.x) % p.npairs;
      const int nt_last = (p.nq_main - (p.npairs - 1) * 2 * kTQ) > kTQ ? 2 : 1;  // tiles of the last query range of a (b, h)
      const uint32_t ts0 = tS(g, 0), to0 = tO(g, 0);
      const uint32_t b_sfull = bar_grp(S_FULL, g, 0), b_pfull = bar_grp(P_FULL, g, 0),
                     b_pfree = bar_grp(P_FREE, g, 0), b_ofull = bar_grp(O_FULL, g, 0), b_ofree = bar_grp(O_FREE, g, 0);
      uint32_t cs = 0, cp = 0, no = 0;
      const int total = n_local * p.nblk;
      // (item, block, stage, phase) of the S step and of the PV step, which trails it by one block
      int nS = 0, jS = 0, stS = 0, jP = 0, stP = 0;
      int qpS = static_cast<int>(blockIdx.x) % p.npairs, qpP = qpS;
      uint32_t phS = 0;
      for (int c = 0; c <= total; ++c) {
        if (c < total) {
          const bool active = g == 0 || qpS != p.npairs - 1 || nt_last > 1;
          const bool lastS = jS == p.nblk - 1;
          if (jS == 0) mbar_wait(bar(kBarQFull + (nS & 1)), (nS >> 1) & 1);
          mbar_wait(bar(kBarKVFull + stS), phS);
          if (g == 0) VL_STAMP(1, nS);  // K / V block landed
          if (active) {
            // S buffer cs & 1 last held P of block cs - 2, read by P V(cs - 2): issued by this thread earlier, and a thread's
            // tcgen05.mma execute in issue order -- no barrier needed for the reuse
            const uint32_t buf = cs & 1;
            tc_fence_after();
            const uint32_t a_lo = q_lo + static_cast<uint32_t>(nS & 1) * 2048u;
            const uint32_t b_lo = k_lo + static_cast<uint32_t>(stS) * 1024u;
            const uint32_t idesc_s = lastS ? idesc_slast : idesc_s64;
#pragma unroll
            for (int k = 0; k < kHD / 16; ++k) umma_ss_lohi(ts0 + 64 * buf, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc_s, k > 0);
            umma_commit(b_sfull + 8 * buf);
            if (g == 0) VL_STAMP(1, nS);  // S issued
            ++cs;
          } else {
            // this group has no tile in the item: release its share of the Q stage and of the K / V stage (both loaded, so
            // the arrivals land in the right barrier phase)
            if (lastS) mbar_arrive_cnt(bar(kBarQEmpty + (nS & 1)), 4);
            mbar_arrive(bar(kBarKVEmpty + stS));
          }
          if (++jS == p.nblk) {
            jS = 0;
            ++nS;
            qpS += qp_step;
            if (qpS >= p.npairs) qpS -= p.npairs;
          }
          if (++stS == kNS) {
            stS = 0;
            phS ^= 1;
          }
        }
        if (c >= 1) {
          const bool active = g == 0 || qpP != p.npairs - 1 || nt_last > 1;
          const bool lastP = jP == p.nblk - 1;
          if (active) {
            const uint32_t buf = cp & 1, ob = no & 1;
            if (jP == 0) mbar_wait(b_ofree + 8 * ob, ((no >> 1) & 1) ^ 1);
            if (g == 0) VL_STAMP(1, nS);  // S issued, waiting for P
            mbar_wait(b_pfull + 8 * buf, (cp >> 1) & 1);
            tc_fence_after();
            if (g == 0) VL_STAMP(1, nS);  // P ready -> issue P V
            const uint32_t a_tm = ts0 + 64 * buf;  // P (bf16 pairs) sits in the first 32 columns of the block's S buffer
            const uint32_t b_lo = v_lo + static_cast<uint32_t>(stP) * 1024u;
            const int nkk = lastP ? kk_last : kBK / 16;
#pragma unroll
            for (int kk = 0; kk < kBK / 16; ++kk)
              if (kk < nkk) umma_ts_lohi(to0 + 64 * ob, a_tm + 8 * kk, b_lo + 128 * kk, kDescHi, idesc_o, (jP > 0 || kk > 0) ? 1u : 0u);
            umma_commit(b_pfree + 8 * buf);
            if (lastP) {
              umma_commit(b_ofull + 8 * ob);
              ++no;
            }
            umma_commit(bar(kBarKVEmpty + stP));
            ++cp;
          }
          if (++jP == p.nblk) {
            jP = 0;
            qpP += qp_step;
            if (qpP >= p.npairs) qpP -= p.npairs;
          }
          if (++stP == kNS) stP = 0;
        }
      }
    }
   }
  } else if (warp < 12) {
    // ================================================================== softmax groups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float sl2 = p.scale * kLog2e;
    constexpr float kLazy = 8.0f;
    uint32_t cg = 0, ng = 0;
    // deferred epilogue of the previous item
    bool pend = false;
    float pend_l = 1.f, pend_off = 0.f;
    long long pend_lse = 0;
    bool pend_ok = false;
    uint32_t pend_ob = 0, pend_ph = 0;
    int pend_qs = 0, pend_c0 = 0, pend_c1 = 0, pend_c2 = 0;
    // the group's bulk store of an item's O tile reads the item's (retired) Q tile; the Q stage goes back to the producer
    // once that read has finished
    bool store_pending = false;
    int store_qs = 0;

    // O tile of the finished item: TMEM -> registers -> (1 / l) -> bf16 -> the item's own Q tile in shared memory (its S MMAs
    // retired before o_full) -> one bulk tensor store per warp (32 rows of whole 128-byte lines; rows past nq_main clipped by the map)
    auto epilogue = [&]() {
      mbar_wait(bar_grp(O_FULL, g, pend_ob), pend_ph);
      tc_fence_after();
      uint32_t a[32], c2[32];
      tmem_ld32(tO(g, pend_ob) + lane_off, a);
      tmem_ld32(tO(g, pend_ob) + lane_off + 32, c2);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_grp(O_FREE, g, pend_ob));
      const float inv = 1.0f / pend_l;
      uint8_t* tile = bp + kOffQ + (pend_qs * 2 + g) * 16384;
#pragma unroll
      for (int t = 0; t < 32; t += 8)
        *reinterpret_cast<uint4*>(tile + sw128_off(r, t)) =
            make_uint4(pack_bf16(__uint_as_float(a[t]) * inv, __uint_as_float(a[t + 1]) * inv), pack_bf16(__uint_as_float(a[t + 2]) * inv, __uint_as_float(a[t + 3]) * inv),
                       pack_bf16(__uint_as_float(a[t + 4]) * inv, __uint_as_float(a[t + 5]) * inv), pack_bf16(__uint_as_float(a[t + 6]) * inv, __uint_as_float(a[t + 7]) * inv));
#pragma unroll
      for (int t = 0; t < 32; t += 8)
        *reinterpret_cast<uint4*>(tile + sw128_off(r, 32 + t)) =
            make_uint4(pack_bf16(__uint_as_float(c2[t]) * inv, __uint_as_float(c2[t + 1]) * inv), pack_bf16(__uint_as_float(c2[t + 2]) * inv, __uint_as_float(c2[t + 3]) * inv),
                       pack_bf16(__uint_as_float(c2[t + 4]) * inv, __uint_as_float(c2[t + 5]) * inv), pack_bf16(__uint_as_float(c2[t + 6]) * inv, __uint_as_float(c2[t + 7]) * inv));
      if (pend_ok && p.lse) p.lse[pend_lse] = pend_off * p.scale + __logf(pend_l);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {  // each warp stores its own 32-row slab: no group-wide barrier
        tma_store_3d(&tmO, sQ(pend_qs, g) + quarter * 4096, pend_c0, pend_c1 + quarter * 32, pend_c2);
        tma_store_commit();
      }
      store_pending = true;
      store_qs = pend_qs;
      pend = false;
    };
    auto release_q = [&]() {
      if (lane == 0) {
        tma_store_wait_read<0>();
        mbar_arrive(bar(kBarQEmpty + store_qs));
      }
      store_pending = false;
    };
    // The epilogue of an item is deferred into the group's next item (after that item's first block has been handed to the MMA
    // warp) only when the group has a next item right behind this one and more than one block per item -- otherwise the Q
    // stage would be released too late for the producer (it holds the tile of the item after next).

    for (int n = 0; n < n_local; ++n) {
      const Item it = item_at(n);
      if (g >= it.nt) continue;
      const int qrow = it.qp * 2 * kTQ + g * kTQ + r;
      const bool defer = p.nblk > 1 && n + 1 < n_local && g < item_at(n + 1).nt;
      float m = -INFINITY, l = 0.f, m_off = -INFINITY;  // raw score units
      for (int j = 0; j < p.nblk; ++j) {
        const int buf = cg & 1;
        const uint32_t ph = (cg >> 1) & 1;
        const bool last = j == p.nblk - 1;
        const int nkb = last ? p.nkb_last : kBK;
        const int valid = last ? p.valid_last : kBK;
        const bool full = valid == kBK;
        if (warp == 4) VL_STAMP(2, n);  // block start
        mbar_wait(bar_grp(S_FULL, g, buf), ph);
        tc_fence_after();
        if (warp == 4) VL_STAMP(2, n);  // S ready
        uint32_t va[32], vb[32];
        float bm = -INFINITY;
        if (full) {
          tmem_ld32(tS(g, buf) + lane_off, va);
          tmem_ld32(tS(g, buf) + lane_off + 32, vb);
          tc_wait_ld();  // the scores live in registers from here on
#pragma unroll
          for (int t = 0; t < 32; ++t) bm = fmaxf(bm, fmaxf(__uint_as_float(va[t]), __uint_as_float(vb[t])));
        } else {
          // partial last block (257 = 4 x 64 + 1): 16-column groups, first pass = masked row maximum
          for (int c = 0; c < nkb; c += 16) {
            uint32_t v[16];
            tmem_ld16(tS(g, buf) + lane_off + c, v);
            tc_wait_ld();
#pragma unroll
            for (int t = 0; t < 16; ++t)
              if (c + t < valid) bm = fmaxf(bm, __uint_as_float(v[t]));
          }
        }
        if (warp == 4) VL_STAMP(2, n);  // scores in registers
        const float m_new = fmaxf(m, bm);
        const bool move = (m_new - m_off) * sl2 > kLazy || m_off == -INFINITY;
        const float off_new = move ? m_new : m_off;
        const float msub = off_new * sl2;
        const float alpha = move ? ex2_approx(m_off * sl2 - msub) : 1.0f;  // first block: m_off = -inf -> 0
        float bs0 = 0.f, bs1 = 0.f;
        uint32_t pk[32];
        if (full) {
          const float2 sl22 = make_float2(sl2, sl2), nms = make_float2(-msub, -msub);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int t = 0; t < 32; t += 2) {  // FFMA2 / FADD2: two scores per FMA-pipe instruction
            const float2 a = __ffma2_rn(make_float2(__uint_as_float(va[t]), __uint_as_float(va[t + 1])), sl22, nms);
            const float2 b2 = __ffma2_rn(make_float2(__uint_as_float(vb[t]), __uint_as_float(vb[t + 1])), sl22, nms);
            const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y)), f = make_float2(ex2_approx(b2.x), ex2_approx(b2.y));
            acc0 = __fadd2_rn(acc0, e);
            acc1 = __fadd2_rn(acc1, f);
            pk[t >> 1] = pack_bf16(e.x, e.y);
            pk[16 + (t >> 1)] = pack_bf16(f.x, f.y);
          }
          bs0 = acc0.x + acc0.y;
          bs1 = acc1.x + acc1.y;
        }
        m = m_new;
        m_off = off_new;
        if (warp == 4) VL_STAMP(2, n);  // exps done
        if (j > 0 && __any_sync(0xffffffffu, move)) {
          // O *= alpha for rows whose offset moved; the previous P V of this item must have retired
          mbar_wait(bar_grp(P_FREE, g, buf ^ 1), ((cg - 1) >> 1) & 1);
          tc_fence_after();
          const uint32_t to = tO(g, ng & 1) + lane_off;
#pragma unroll
          for (int c = 0; c < kHD; c += 16) {
            uint32_t v[16];
            tmem_ld16(to + c, v);
            tc_wait_ld();
#pragma unroll
            for (int t = 0; t < 16; ++t) v[t] = __float_as_uint(__uint_as_float(v[t]) * alpha);
            tmem_st16(to + c, v);
          }
          tc_wait_st();
        }
        // P (bf16) goes back into TMEM, over the block's own scores, as the A operand of P V: row r in lane r, keys (2i, 2i + 1)
        // in 32-bit column i -- no shared-memory round trip, no async-proxy fence
        if (full) {
          tmem_st32(tS(g, buf) + lane_off, pk);
        } else {
          // second pass over the partial block: P = exp2(.) (masked)
          uint4 wq[4], wq2[4];
#pragma unroll
          for (int c = 0; c < kBK; c += 16) {
            if (c >= nkb) break;
            uint32_t v[16];
            tmem_ld16(tS(g, buf) + lane_off + c, v);
            tc_wait_ld();
            uint32_t w[8];
#pragma unroll
            for (int t = 0; t < 16; t += 2) {
              const float e0 = (c + t < valid) ? ex2_approx(fmaf(__uint_as_float(v[t]), sl2, -msub)) : 0.f;
              const float e1 = (c + t + 1 < valid) ? ex2_approx(fmaf(__uint_as_float(v[t + 1]), sl2, -msub)) : 0.f;
              bs0 += e0 + e1;
              w[t >> 1] = pack_bf16(e0, e1);
            }
            wq[c >> 4] = make_uint4(w[0], w[1], w[2], w[3]);
            wq2[c >> 4] = make_uint4(w[4], w[5], w[6], w[7]);
          }
          // all groups read before any is overwritten (P group c lands on the columns of score groups c/2)
#pragma unroll
          for (int c = 0; c < kBK; c += 16) {
            if (c >= nkb) break;
            const uint32_t w8[8] = {wq[c >> 4].x, wq[c >> 4].y, wq[c >> 4].z, wq[c >> 4].w, wq2[c >> 4].x, wq2[c >> 4].y, wq2[c >> 4].z, wq2[c >> 4].w};
            tmem_st8(tS(g, buf) + lane_off + (c >> 1), w8);
          }
        }
        l = l * alpha + (bs0 + bs1);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_grp(P_FULL, g, buf));
        if (warp == 4) VL_STAMP(2, n);  // P handed over
        ++cg;
        if (store_pending) release_q();
        // one epilogue call site, two occasions: e = 0 the previous item's deferred epilogue (after this item's first block),
        // e = 1 this item's own, right after its last block, when it cannot be deferred
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
          if (e == 0) {
            if (!(pend && j == 0)) continue;
          } else {
            if (!last) break;
            pend = true;
            pend_l = l;
            pend_off = m_off;
            pend_ok = qrow < p.nq_main;
            pend_lse = (static_cast<long long>(it.b) * p.H + it.h) * p.nq + qrow;
            pend_ob = ng & 1;
            pend_ph = (ng >> 1) & 1;
            pend_qs = n & 1;
            pend_c0 = it.h * kHD;
            pend_c1 = it.qp * 2 * kTQ + g * kTQ;
            pend_c2 = it.b;
            ++ng;
            if (defer) break;
          }
          epilogue();
          if (e == 1) release_q();
          if (warp == 4) VL_STAMP(2, n);  // an item's rows stored
        }
      }
    }
    if (store_pending) release_q();
    if (lane == 0) tma_store_wait<0>();  // bulk stores complete before the CTA retires
  } else {
    // ================================================================== the tail query row on CUDA cores
    // Lane (kl, dh) of tail warp tw owns key 16 tw + kl of every block and the dim half dh: it keeps its OWN online-softmax
    // state over the keys it sees (running offset, sum, 32 un-normalised output dims), so a block costs one shuffle (the
    // two halves of q . k) and no reductions; the 16 x 4 partial states are merged once per item.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    const int tw = warp - 12;
    const int kl = lane & 15, dh = lane >> 4;
    const float sl2 = p.scale * kLog2e;
    constexpr float kLazy = 8.0f;
    float* sf = reinterpret_cast<float*>(bp + kOffF);
    float* sq = sf + kFQ + tw * kHD;
    float* mg = sf + kFMerge;
    int st = 0;
    uint32_t st_ph = 0;
    for (int n = 0; n < n_local; ++n) {
      const Item it = item_at(n);
      const bool has_tail = p.tq > 0 && it.qp == p.npairs - 1;
      float off = -INFINITY, l = 0.f;  // log2 units
      float acc[32];
#pragma unroll
      for (int d = 0; d < 32; ++d) acc[d] = 0.f;
      if (has_tail) {
        __syncwarp();
        const uint32_t w = *reinterpret_cast<const uint32_t*>(p.q + (static_cast<long long>(it.b) * p.nq + p.nq_main) * p.ldq + it.h * kHD + 2 * lane);
        sq[2 * lane] = bf16_lo(w);
        sq[2 * lane + 1] = bf16_hi(w);
        __syncwarp();
      }
      for (int j = 0; j < p.nblk; ++j) {
        mbar_wait(bar(kBarKVFull + st), st_ph);
        if (warp == 12) VL_STAMP(3, n);  // K / V block landed
        if (has_tail) {
          const int valid = (j == p.nblk - 1) ? p.valid_last : kBK;
          const int key = tw * 16 + kl;
          if (tw * 16 < valid) {  // warp-uniform: some key of this warp is valid
            const uint8_t* Kt = bp + kOffKV + st * 16384;
            const uint8_t* Vt = Kt + 8192;
            const float* qv = sq + dh * 32;
            float ac[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 kr = *reinterpret_cast<const uint4*>(Kt + sw128_off(key, (dh * 4 + u) * 8));
              const float4 x = *reinterpret_cast<const float4*>(qv + u * 8), y = *reinterpret_cast<const float4*>(qv + u * 8 + 4);
              ac[u] = bf16_lo(kr.x) * x.x + bf16_hi(kr.x) * x.y + bf16_lo(kr.y) * x.z + bf16_hi(kr.y) * x.w + bf16_lo(kr.z) * y.x +
                      bf16_hi(kr.z) * y.y + bf16_lo(kr.w) * y.z + bf16_hi(kr.w) * y.w;
            }
            float dot = (ac[0] + ac[1]) + (ac[2] + ac[3]);
            dot += __shfl_xor_sync(0xffffffffu, dot, 16);  // the two dim halves
            if (key < valid) {                             // keys past the end hold stale rows: skipped entirely
              const float s2 = dot * sl2;
              if (s2 - off > kLazy) {  // also the first block (off = -inf): move the offset, rescale what has been summed
                const float alpha = ex2_approx(off - s2);
                off = s2;
                l *= alpha;
#pragma unroll
                for (int d = 0; d < 32; ++d) acc[d] *= alpha;
              }
              const float pv = ex2_approx(s2 - off);
              l += pv;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint4 vr = *reinterpret_cast<const uint4*>(Vt + sw128_off(key, (dh * 4 + u) * 8));
                acc[8 * u + 0] = fmaf(pv, bf16_lo(vr.x), acc[8 * u + 0]);
                acc[8 * u + 1] = fmaf(pv, bf16_hi(vr.x), acc[8 * u + 1]);
                acc[8 * u + 2] = fmaf(pv, bf16_lo(vr.y), acc[8 * u + 2]);
                acc[8 * u + 3] = fmaf(pv, bf16_hi(vr.y), acc[8 * u + 3]);
                acc[8 * u + 4] = fmaf(pv, bf16_lo(vr.z), acc[8 * u + 4]);
                acc[8 * u + 5] = fmaf(pv, bf16_hi(vr.z), acc[8 * u + 5]);
                acc[8 * u + 6] = fmaf(pv, bf16_lo(vr.w), acc[8 * u + 6]);
                acc[8 * u + 7] = fmaf(pv, bf16_hi(vr.w), acc[8 * u + 7]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBarKVEmpty + st));
        if (warp == 12) VL_STAMP(3, n);  // tail row done with the block
        if (++st == kNS) {
          st = 0;
          st_ph ^= 1;
        }
      }
      if (has_tail) {
        // merge: the 16 key lanes of each dim half (butterfly), then the four warps through shared memory
        float M = off;
#pragma unroll
        for (int o2 = 8; o2 > 0; o2 >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o2));
        const float wgt = (off == -INFINITY) ? 0.f : ex2_approx(off - M);  // lanes that saw no key contribute nothing
        l *= wgt;
#pragma unroll
        for (int d = 0; d < 32; ++d) acc[d] *= wgt;
#pragma unroll
        for (int o2 = 8; o2 > 0; o2 >>= 1) {
          l += __shfl_xor_sync(0xffffffffu, l, o2);
#pragma unroll
          for (int d = 0; d < 32; ++d) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o2);
        }
        // lanes kl == 0 (dh = 0, 1) hold the warp's state: [4 warps][2 + 64]
        if (kl == 0) {
          float* d0 = mg + tw * 66;
          if (dh == 0) {
            d0[0] = M;
            d0[1] = l;
          }
#pragma unroll
          for (int d = 0; d < 32; ++d) d0[2 + dh * 32 + d] = acc[d];
        }
        asm volatile("bar.sync 4, 128;" ::: "memory");
        if (tw == 0) {
          float MM = -INFINITY;
#pragma unroll
          for (int w2 = 0; w2 < 4; ++w2) MM = fmaxf(MM, mg[w2 * 66]);
          float L = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
          for (int w2 = 0; w2 < 4; ++w2) {
            const float* d0 = mg + w2 * 66;
            const float w8 = ex2_approx(d0[0] - MM);  // a warp that saw no key: -inf -> 0
            L = fmaf(d0[1], w8, L);
            o0 = fmaf(d0[2 + 2 * lane], w8, o0);
            o1 = fmaf(d0[3 + 2 * lane], w8, o1);
          }
          const float inv = 1.0f / L;
          const long long row = static_cast<long long>(it.b) * p.nq + p.nq_main;
          *reinterpret_cast<uint32_t*>(p.o + row * p.ldo + it.h * kHD + 2 * lane) = pack_bf16(o0 * inv, o1 * inv);
          if (lane == 0 && p.lse) p.lse[(static_cast<long long>(it.b) * p.H + it.h) * p.nq + p.nq_main] = (MM + log2f(L)) * kLn2;
        }
        asm volatile("bar.sync 4, 128;" ::: "memory");  // merge area free for the next item
      }
    }
  }

#undef VL_STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace fwd2

// Host entry used by vl_attention_fwd (attention.cu): non-causal, at least two 128-row query tiles.
int launch_attn_fwd2(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int nq, int nk, long long ldq, long long ldk,
                     long long ldv, long long ldo, float scale, cudaStream_t stream) {
  using namespace fwd2;
  CUtensorMap tmQ, tmK, tmV, tmK16, tmV16;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmQ, q, (uint64_t)H * kHD, (uint64_t)B * nq, ldq, kHD, kTQ))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, k, (uint64_t)H * kHD, (uint64_t)B * nk, ldk, kHD, kBK))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, v, (uint64_t)H * kHD, (uint64_t)B * nk, ldv, kHD, kBK))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK16, k, (uint64_t)H * kHD, (uint64_t)B * nk, ldk, kHD, 16))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV16, v, (uint64_t)H * kHD, (uint64_t)B * nk, ldv, kHD, 16))) return rc;
  Params p;
  p.B = B; p.H = H; p.nq = nq; p.nk = nk;
  const int t = nq % kTQ;
  p.tq = (nq > kTQ && t > 0 && t <= kMaxTail) ? t : 0;
  p.nq_main = nq - p.tq;
  p.npairs = (p.nq_main + 2 * kTQ - 1) / (2 * kTQ);
  p.nblk = (nk + kBK - 1) / kBK;
  p.valid_last = nk - (p.nblk - 1) * kBK;
  p.nkb_last = (p.valid_last + 15) & ~15;
  const long long items = (long long)B * H * p.npairs;
  VL_CHECK_ARG(items < (1ll << 30), "vl_attention_fwd: too many work items");
  p.n_items = (int)items;
  p.scale = scale;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.ldq = ldq;
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.ldo = ldo; p.lse = lse;
  p.dbg = debug_buffer();
  // output as [cols, main rows of one batch element, batch]: the row box is clipped per batch element, so a partial last tile
  // never spills into the tail rows (written by the tail warps) or the next batch element
  CUtensorMap tmO;
  {
    VL_CHECK_ARG((reinterpret_cast<uintptr_t>(o) & 15) == 0, "vl_attention_fwd: output pointer must be 16-byte aligned");
    const uint64_t dims[3] = {(uint64_t)H * kHD, (uint64_t)p.nq_main, (uint64_t)B};
    const uint64_t strides[2] = {(uint64_t)ldo * 2, (uint64_t)nq * (uint64_t)ldo * 2};
    const uint32_t box[3] = {kHD, 32, 1};  // one softmax warp's rows
    if ((rc = make_tmap(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, o, dims, strides, box, true))) return rc;
  }
  static bool attr = false;
  if (!attr) {
    VL_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr = true;
  }
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  attn_fwd2_kernel<<<grid, kThreads, kSmem, stream>>>(tmQ, tmK, tmV, tmK16, tmV16, tmO, p);
  return launch_check("attn_fwd2_kernel");
}

}  // namespace vl
