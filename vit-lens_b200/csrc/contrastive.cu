// Fused backward of the gathered InfoNCE loss, one launch per direction (reference: gather_features + ClipLoss / TriClipLoss,
// open_clip/loss.py:55-76, 116-138, 158-163; the mask variants loss.py:485-903).
//
//   dX[i, :] = s * sum_j g[i, j] * Y[j, :],      g = gs * (softmax_row(z) + softmax_col(z) - k * onehot) (* mask),   z = s * X Y^T
//
// The two-GEMM path (VL_EPI_CLIPGRAD, then a plain GEMM) writes g[B_loc x B_all] to HBM and reads it back.  Here a CTA owns 128
// rows of X and a 256-column slice of dX and walks over the 128-row blocks of Y -- local rows, or every rank's rows read in place
// from the peer arenas over NVLink (one tensor map per peer), which is the all-gather of the reference fused into the kernel:
//   S = X_m Y_j^T  (tcgen05, K = E streamed through a TMA ring)  ->  TMEM
//   g from S, the row / column log-sum-exps of the forward pass and the label diagonal (one thread per row), written back over S
//   as bf16 pairs (tcgen05.st)  ->  the A operand of
//   dX_slice += g Y_j[:, slice]  (B = the slice's 64-column tiles of Y_j, MN-major)
// so logits and their gradient never leave the SM.  d(loss) / d(scale) = sum g * (X Y^T) is reduced per warp and summed in a
// fixed order afterwards (deterministic).  S is recomputed by each of the E / 256 slice CTAs of a row block (the loss is < 0.1 %
// of a step; what matters here is that nothing B_all-sized touches HBM).
#include <type_traits>

#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {
namespace clipbwd {

constexpr int kT = 128;       // rows of X per CTA / rows of Y per block
constexpr int kSlice = 256;   // dX columns per CTA
constexpr int kStages = 4;    // ring of (X tile, Y tile) pairs, 64 columns of E each
constexpr int kThreads = 192; // warp 0 TMA producer, warp 1 tcgen05 issuer, warps 2-5 one thread per row

constexpr int kOffRing = 0;                          // kStages x (16 KB X + 16 KB Y)
constexpr int kOffYc = kOffRing + kStages * 32768;   // 4 x 16 KB: Y_j[:, slice] as 64-column tiles
constexpr int kOffBar = kOffYc + 4 * 16384;
constexpr int kSmem = kOffBar + 256 + 1024;

struct PeerMaps {
  CUtensorMap m[8];
};

struct Params {
  int M, N, E;
  int peer_rows;         // rows of Y per peer map (N when Y is local)
  int label_off;         // the label of row i is column i + label_off
  int ds_row_only;
  float gscale;
  const float* alpha_dev;   // logit scale s (device scalar)
  const float* gscale_dev;  // optional extra factor (upstream gradient), device scalar
  const float* row_lse;     // [M]
  const float* col_lse;     // [N] or null
  const uint8_t* mask;      // [M, N] or null
  long long ldmask;
  float* dx;                // [M, E] fp32
  long long lddx;
  float* ds_part;           // [tiles_m * 4] partial sums of g * (X Y^T) (slice 0 CTAs only)
};

__device__ __forceinline__ int ld_acquire_sys_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreads, 1)
clip_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ PeerMaps pmY, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* bp = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + kOffBar;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (kStages + s); };
  const uint32_t bar_ycfull = bars + 8u * (2 * kStages), bar_ycempty = bars + 8u * (2 * kStages + 1), bar_sfull = bars + 8u * (2 * kStages + 2),
                 bar_gfull = bars + 8u * (2 * kStages + 3), bar_dxfull = bars + 8u * (2 * kStages + 4), tmem_slot = bars + 8u * (2 * kStages + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(bp + kOffBar + 8 * (2 * kStages + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kT;
  const int c0 = blockIdx.y * kSlice;                      // first dX column of this CTA
  const int ncw = min(kSlice, p.E - c0);                   // slice width (multiple of 64)
  const int nyc = ncw / 64;
  const int ne = p.E / 64;                                 // 64-column steps of the S reduction
  const int nblk = (p.N + kT - 1) / kT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_ycfull, 1);
    mbar_init(bar_ycempty, 1);
    mbar_init(bar_sfull, 1);
    mbar_init(bar_gfull, 4);
    mbar_init(bar_dxfull, 1);
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
  const uint32_t tSc = tmem, tDX = tmem + 128;  // S / g: 128 columns; dX slice: up to 256 columns

  if (warp == 0) {
    // ================================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        const int row = j * kT;
        const int pr = row / p.peer_rows;  // (host: a block never straddles two peers)
        const CUtensorMap* my = &pmY.m[pr];
        const int lrow = row - pr * p.peer_rows;
        for (int e = 0; e < ne; ++e) {
          mbar_wait_trap(bar_empty(stage), phase ^ 1);
          const uint32_t sx = base + kOffRing + stage * 32768;
          mbar_expect_tx(bar_full(stage), 32768);
          tma_load_2d(sx, &tmX, bar_full(stage), e * 64, m0);
          tma_load_2d(sx + 16384, my, bar_full(stage), e * 64, lrow);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        mbar_wait_trap(bar_ycempty, (j & 1) ^ 1);  // the dX MMAs of block j - 1 have read the slice tiles
        mbar_expect_tx(bar_ycfull, nyc * 16384);
        for (int c = 0; c < nyc; ++c) tma_load_2d(base + kOffYc + c * 16384, my, bar_ycfull, c0 + c * 64, lrow);
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_bf16(kT, kT, 0, 0);    // S: A = X tile (K-major), B = Y tile (K-major), N = 128
      const uint32_t idesc_dx = umma_idesc_bf16(kT, ncw, 0, 1);  // dX: A = g in TMEM, B = Y slice tiles (MN-major), N = slice width
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nblk; ++j) {
        // S_j (overwrites g of block j - 1: its dX MMAs were issued before, same thread -> executed in order)
        for (int e = 0; e < ne; ++e) {
          mbar_wait_trap(bar_full(stage), phase);
          tc_fence_after();
          const uint32_t sx = base + kOffRing + stage * 32768;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss(tSc, umma_desc_sw128(sx + k * 32, 16, 1024), umma_desc_sw128(sx + 16384 + k * 32, 16, 1024), idesc_s, (e > 0 || k > 0) ? 1u : 0u);
          umma_commit(bar_empty(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(bar_sfull);
        // dX_slice += g_j Y_j[:, slice]
        mbar_wait_trap(bar_gfull, j & 1);
        mbar_wait_trap(bar_ycfull, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kT / 16; ++kk)
          umma_ts(tDX, tSc + 8 * kk, umma_desc_sw128(base + kOffYc + kk * 2048, 16384, 1024), idesc_dx, (j > 0 || kk > 0) ? 1u : 0u);
        umma_commit(bar_ycempty);
      }
      umma_commit(bar_dxfull);
    }
  } else {
    // ================================================================== one thread per row of X
    const int quarter = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int row = m0 + quarter * 32 + lane;
    const bool row_ok = row < p.M;
    const float alpha = __ldg(p.alpha_dev);
    const float gs = p.gscale * (p.gscale_dev ? __ldg(p.gscale_dev) : 1.0f);
    const float rl = row_ok ? __ldg(p.row_lse + row) : 0.f;
    const int diag = row + p.label_off;
    const float kdiag = p.col_lse ? 2.f : 1.f;
    float dsum = 0.f;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait_trap(bar_sfull, j & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kT; c += 32) {
        uint32_t v[32];
        tmem_ld32(tSc + lane_off + c, v);
        tc_wait_ld();
        const int col0 = j * kT + c;
        uint32_t mbits = 0xffffffffu;
        if (p.mask != nullptr && row_ok) {
          mbits = 0u;
          const uint8_t* mr = p.mask + static_cast<long long>(row) * p.ldmask + col0;
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (col0 + t < p.N && mr[t] != 0) mbits |= 1u << t;
        }
        uint32_t gw[16];
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
          float g2[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = col0 + t + u;
            const float accv = __uint_as_float(v[t + u]);
            const bool keep = (mbits >> (t + u)) & 1u;
            const float z = keep ? accv * alpha : 0.f;
            float gval = 0.f;
            if (row_ok && col < p.N) {
              gval = __expf(z - rl);
              const bool on_diag = col == diag;
              const float grow = gval - (on_diag ? 1.f : 0.f);  // this row's own cross-entropy term
              if (p.col_lse) gval += __expf(z - __ldg(p.col_lse + col));
              if (on_diag) gval -= kdiag;
              gval *= gs;
              if (!keep) gval = 0.f;  // d(logit * mask) / d(logit) = mask
              else dsum += (p.ds_row_only ? grow * gs : gval) * accv;
            }
            g2[u] = gval;
          }
          gw[t >> 1] = pack_bf16(g2[0], g2[1]);
        }
        // g (bf16 pairs) over the scores already read: row in its lane, columns (2 i, 2 i + 1) in 32-bit column i
        tmem_st16(tSc + lane_off + (c >> 1), gw);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_gfull);
    }
    // dX slice -> global (fp32), scaled by s
    mbar_wait_trap(bar_dxfull, 0);
    tc_fence_after();
    for (int c = 0; c < ncw; c += 32) {
      uint32_t v[32];
      tmem_ld32(tDX + lane_off + c, v);
      tc_wait_ld();
      if (row_ok) {
        float* dst = p.dx + static_cast<long long>(row) * p.lddx + c0 + c;
#pragma unroll
        for (int t = 0; t < 32; t += 4)
          *reinterpret_cast<float4*>(dst + t) = make_float4(__uint_as_float(v[t]) * alpha, __uint_as_float(v[t + 1]) * alpha, __uint_as_float(v[t + 2]) * alpha,
                                                            __uint_as_float(v[t + 3]) * alpha);
      }
    }
    if (blockIdx.y == 0 && p.ds_part != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
      if (lane == 0) p.ds_part[blockIdx.x * 4 + quarter] = dsum;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace clipbwd
}  // namespace vl

extern "C" int vl_clip_backward(const VlClipBwdArgs* a, void* stream_) {
  using namespace vl;
  using namespace vl::clipbwd;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VL_CHECK_ARG(a != nullptr && a->x && (a->y || a->y_peers) && a->dx && a->row_lse && a->alpha_dev, "vl_clip_backward: null pointer");
  VL_CHECK_ARG(a->M > 0 && a->N > 0 && a->E >= 64 && a->E % 64 == 0, "vl_clip_backward: M=%d N=%d E=%d (E must be a multiple of 64)", a->M, a->N, a->E);
  VL_CHECK_ARG(a->ldx % 8 == 0 && a->ldy % 8 == 0 && a->lddx % 4 == 0 && (reinterpret_cast<uintptr_t>(a->dx) & 15) == 0, "vl_clip_backward: misaligned operand");
  VL_CHECK_ARG(a->mask == nullptr || a->ldmask >= a->N, "vl_clip_backward: ldmask < N");
  Params p;
  p.M = a->M; p.N = a->N; p.E = a->E;
  p.label_off = a->label_off;
  p.ds_row_only = a->ds_row_only;
  p.gscale = a->gscale;
  p.alpha_dev = a->alpha_dev;
  p.gscale_dev = a->gscale_dev;
  p.row_lse = a->row_lse;
  p.col_lse = a->col_lse;
  p.mask = a->mask;
  p.ldmask = a->ldmask;
  p.dx = a->dx;
  p.lddx = a->lddx;
  CUtensorMap tmX;
  PeerMaps pm;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmX, a->x, a->E, a->M, a->ldx, 64, kT))) return rc;
  if (a->y_peers == nullptr) {
    p.peer_rows = a->N > kT ? ((a->N + kT - 1) / kT) * kT : kT;  // one map over all rows
    if ((rc = make_tmap_bf16_2d(&pm.m[0], a->y, a->E, a->N, a->ldy, 64, kT))) return rc;
    for (int q = 1; q < 8; ++q) pm.m[q] = pm.m[0];
  } else {
    VL_CHECK_ARG(a->y_npeers >= 1 && a->y_npeers <= 8 && a->y_peer_rows > 0, "vl_clip_backward: y_npeers / y_peer_rows invalid");
    VL_CHECK_ARG(static_cast<long long>(a->y_npeers) * a->y_peer_rows == a->N, "vl_clip_backward: y_npeers * y_peer_rows must equal N");
    VL_CHECK_ARG(a->y_npeers == 1 || a->y_peer_rows % kT == 0, "vl_clip_backward: y_peer_rows must be a multiple of 128 (a block must not straddle two peers)");
    p.peer_rows = a->y_peer_rows;
    // one map per peer over ITS rows only: rows past a ragged last block read as zero, never the next allocation
    for (int q = 0; q < a->y_npeers; ++q) {
      VL_CHECK_ARG(a->y_peers[q] != nullptr, "vl_clip_backward: null peer pointer");
      if ((rc = make_tmap_bf16_2d(&pm.m[q], a->y_peers[q], a->E, a->y_peer_rows, a->ldy, 64, kT))) return rc;
    }
    for (int q = a->y_npeers; q < 8; ++q) pm.m[q] = pm.m[0];
  }
  const int tiles_m = (a->M + kT - 1) / kT;
  const int slices = (a->E + kSlice - 1) / kSlice;
  float* ds_part = nullptr;
  if (a->ds_out != nullptr) {
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&ds_part), (size_t)tiles_m * 4 * sizeof(float), stream))) return rc;
  }
  p.ds_part = ds_part;
  static bool attr = false;
  if (!attr) {
    VL_CUDA(cudaFuncSetAttribute(clip_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr = true;
  }
  clip_bwd_kernel<<<dim3(tiles_m, slices), kThreads, kSmem, stream>>>(tmX, pm, p);
  if ((rc = launch_check("clip_bwd_kernel"))) return rc;
  if (ds_part != nullptr) {
    if ((rc = launch_colreduce(ds_part, tiles_m * 4, 1, a->ds_out, nullptr, nullptr, stream))) return rc;
    return scratch_free(ds_part, stream);
  }
  return 0;
}
