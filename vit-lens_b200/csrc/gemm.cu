// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue(alpha * A[M,K] * B[N,K]^T),  bf16 operands, fp32 accumulation in TMEM.
//
// Roles (320 threads, 1 CTA / SM, grid = min(#tiles, #SMs)):
//   warp 0      TMA producer   (one elected lane): global -> 128B-swizzled smem ring, mbarrier tx-count
//   warp 1      MMA issuer     (one elected lane): tcgen05.mma 128 x BN x 16, accumulators in TMEM,
//                               tcgen05.commit releases smem stages / publishes accumulators
//   warps 2..9  epilogue       tcgen05.ld TMEM -> registers -> bias / GELU / residual / GELU' -> global
// TMEM holds two accumulator buffers (2 x BN columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  Operands may be K-major ([rows, K] row-major) or MN-major ([K, rows]
// row-major, used by the weight-gradient GEMMs where K is the token dimension).
#include <type_traits>

#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kGroupM = 16;  // m-tiles per L2 reuse group

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;
  int tiles_m, tiles_n, split_k, kb_total, kb_per_split;
  int epi, act_quick, d_f32, accumulate;
  float alpha;
  void* d;
  long long ldd;
  const float* bias;
  const __nv_bfloat16* aux_in;
  __nv_bfloat16* aux_out;
  long long ldaux;
  const float* row_vec;
  const float* col_vec;
  float* out_vec0;
  float* out_vec1;
  float* out_vec2;
  float* scalar_out;
  int iparam;
  float fparam;
  const float* alpha_dev;
  const float* fparam_dev;
  int aux_row_div, relu;
  int loss_flags;
  const uint8_t* mask;           // loss epilogues: optional [M, N] bytes (ld = ldmask); 0 -> the logit is replaced by 0 (loss.py:540-575)
  long long ldmask;
  int b_peer_rows;               // PEER kernels: B's global row r lives in peer r / b_peer_rows at local row r % b_peer_rows
  const int* peer_flags;         // PEER kernels: int32 [n peers] tickets in THIS rank's arena, or null
  int peer_flag_value;
  float* rowsum_out;  // fp32 [tiles_n * split_k][M] partial row sums of A (one row per (n-tile, split); launch_colreduce adds them in
                      // order -> the bias gradient of a weight-gradient GEMM), or null
  int dbg;  // bring-up knob 9: 1 skip epilogue, 2 no global traffic in the epilogue, 4 sleeping epilogue waits, 8 MMA ignores full barriers
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep;  // bytes
};

// TE = true: the epilogue stages each warp's 32 x 64 output block in shared memory (128B-swizzled, 4 KB per warp) and
// moves it with TMA (bulk tensor store; the residual / pre-activation block arrives the same way), so global traffic is
// whole 128-byte lines issued by the copy engine instead of 16-byte per-thread stores.  One pipeline stage pays for it.
constexpr int kEpiBufBytes = 32 * 64 * 2;
template <int BN, bool TE = false>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? (TE ? 3 : 4) : 6;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = TE ? (BN / 16) * kEpiBufBytes : 0;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;
};

__device__ __forceinline__ void tile_coords(const GemmParams& p, int t, int& m_blk, int& n_blk, int& ks) {
  int mn = t / p.split_k;
  ks = t - mn * p.split_k;
  int group_sz = kGroupM * p.tiles_n;
  int g = mn / group_sz;
  int r = mn - g * group_sz;
  int m_first = g * kGroupM;
  int gm = min(kGroupM, p.tiles_m - m_first);
  n_blk = r / gm;
  m_blk = m_first + (r - n_blk * gm);
}

// Epilogue geometry: 4 TMEM lane quarters x kSlices column slices of 64 accumulator columns each.  BN = 256 -> 16
// epilogue warps (4 per SM sub-partition: the GELU / GELU' math is issue-latency bound, more resident warps hide it),
// BN = 128 -> 8 warps.
template <int BN>
struct EpiCfg {
  static constexpr int kSlices = BN / 64;
  static constexpr int kWarps = 4 * kSlices;
  static constexpr int kThreads = 64 + 32 * kWarps;
  static constexpr int kChunk = (kWarps > 8) ? 16 : 32;  // accumulator columns per step (register budget: 65536 / kThreads)
};

// Kernel flavours: EW = 0 -> per-thread global stores, EpiCfg<BN>::kWarps epilogue warps; EW = 8 -> TMA-staged epilogue with
// 8 warps that own two 64-column slices each.  (Measured: 16 warps at the 96-register cap of an 18-warp CTA serialise the
// activation math through one register; 8 warps with 168 registers keep 16 independent element streams in flight and
// are faster on every shape - profiles/r01_gemm_epilogue_attribution.log.)
template <int BN, int EW>
struct KernelCfg {
  static constexpr bool kTE = EW > 0;
  static constexpr int kEpiWarps = EW > 0 ? EW : EpiCfg<BN>::kWarps;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
};

// One output tile's epilogue for the calling warp: rows [row_base + quarter*32, +32), columns of slice (e >> 2) of n-tile
// n_blk; thread t owns row quarter*32 + t and processes it 32 accumulator columns at a time.
template <int BN, int CW>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tmem_base, uint32_t acc_col, int row_base, int n_blk, int ks,
                                              int e, int quarter, int lane, int part_idx) {
  const int slice = e >> 2;
  const float alpha_eff = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.0f);
  const float fparam_eff = p.fparam * (p.fparam_dev ? __ldg(p.fparam_dev) : 1.0f);
  const int row = row_base + quarter * 32 + lane;
  const bool row_ok = row < p.M && !(p.dbg & 2);
  const bool lead_split = (ks == 0);  // bias / aux terms are added by split 0 only
  float lse_m = -INFINITY, lse_s = 0.f, clip_ds = 0.f;
  const long long arow = p.aux_row_div > 1 ? row / p.aux_row_div : row;
#pragma unroll 1
  for (int c = 0; c < 64 / CW; ++c) {
    const int col0 = n_blk * BN + slice * 64 + c * CW;
    if (col0 >= p.N) break;
    const bool full = col0 + CW <= p.N;
    uint32_t v[CW];
    if constexpr (CW == 32)
      tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc_col + slice * 64 + c * CW, v);
    else
      tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc_col + slice * 64 + c * CW, v);
    // issue the independent global loads (bias, residual / pre-activation row) before waiting on the TMEM read
    uint4 ax[CW / 8];
    const bool need_aux = row_ok && ((p.epi == VL_EPI_RESIDUAL && lead_split) || p.epi == VL_EPI_GELU_BWD);
    if (need_aux) {
      const uint4* ap = reinterpret_cast<const uint4*>(p.aux_in + arow * p.ldaux + col0);
#pragma unroll
      for (int j = 0; j < CW / 8; ++j) ax[j] = (col0 + 8 * j < p.N) ? ap[j] : make_uint4(0, 0, 0, 0);
    }
    float bv[CW];
    const bool need_bias = p.bias != nullptr && lead_split && p.epi != VL_EPI_GELU_BWD && p.epi != VL_EPI_ROWLSE && p.epi != VL_EPI_CLIPGRAD;
    if (need_bias) {
#pragma unroll
      for (int j = 0; j < CW; j += 4) {
        const float4 b4 = (col0 + j < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j)) : make_float4(0, 0, 0, 0);
        bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
      }
    }
    tc_wait_ld();
    float f[CW];
    // mask variants of the contrastive loss (ClipLossSimMask / ClipLossLabelMask, loss.py:485-748): `logits * mask` -- a masked
    // logit is the constant 0 (it keeps its softmax mass exp(0 - lse) but carries no gradient)
    uint32_t mbits = 0xffffffffu;
    if (p.mask != nullptr && row_ok && (p.epi == VL_EPI_ROWLSE || p.epi == VL_EPI_CLIPGRAD)) {
      mbits = 0u;
      const uint8_t* mr = p.mask + static_cast<long long>(row) * p.ldmask + col0;
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (col0 + j < p.N && mr[j] != 0) mbits |= 1u << j;
    }
    if (p.epi == VL_EPI_ROWLSE) {
#pragma unroll
      for (int j = 0; j < CW; ++j) f[j] = ((mbits >> j) & 1u) ? __uint_as_float(v[j]) * alpha_eff : 0.f;
      if (c == 0) {
        lse_m = -INFINITY;
        lse_s = 0.f;
      }
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (col0 + j < p.N) cm = fmaxf(cm, f[j]);
      const float nm = fmaxf(lse_m, cm);
      float add = 0.f;
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (col0 + j < p.N) add += __expf(f[j] - nm);
      lse_s = lse_s * __expf(lse_m - nm) + add;
      lse_m = nm;
      if (row_ok) {
        const int dj = row + p.iparam - col0;
        if (dj >= 0 && dj < CW) {
          float dv = 0.f;
#pragma unroll
          for (int j = 0; j < CW; ++j)
            if (j == dj) dv = f[j];
          p.out_vec2[row] = dv;
        }
        if (c == 64 / CW - 1 || col0 + CW >= p.N) {
          const long long po = static_cast<long long>(row) * (p.tiles_n * (BN / 64)) + n_blk * (BN / 64) + slice;
          p.out_vec0[po] = lse_m;
          p.out_vec1[po] = lse_s;
        }
      }
      continue;
    }
    if (p.epi == VL_EPI_CLIPGRAD) {
      const float rl = row_ok ? __ldg(p.row_vec + row) : 0.f;
      float dsum = 0.f;
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const float accv = __uint_as_float(v[j]);
        const bool keep = (mbits >> j) & 1u;
        const float z = keep ? accv * alpha_eff : 0.f;
        float gval = 0.f;
        if (row_ok && col0 + j < p.N) {
          gval = __expf(z - rl);
          const bool on_diag = col0 + j == row + p.iparam;
          const float grow = gval - (on_diag ? 1.f : 0.f);  // this row's own cross-entropy term
          if (p.col_vec) gval += __expf(z - __ldg(p.col_vec + col0 + j));
          if (on_diag) gval -= p.col_vec ? 2.f : 1.f;
          gval *= fparam_eff;
          if (!keep) gval = 0.f;  // d(logit * mask) / d(logit) = mask
          else dsum += ((p.loss_flags & 1) ? grow * fparam_eff : gval) * accv;
        }
        f[j] = gval;
      }
      clip_ds += dsum;
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j) f[j] = __uint_as_float(v[j]) * alpha_eff;
      if (need_bias) {
#pragma unroll
        for (int j = 0; j < CW; ++j) f[j] += bv[j];
      }
      if (p.epi == VL_EPI_GELU) {
        if (p.aux_out != nullptr && row_ok) {
          uint4* up = reinterpret_cast<uint4*>(p.aux_out + static_cast<long long>(row) * p.ldaux + col0);
#pragma unroll
          for (int j = 0; j < CW / 8; ++j)
            if (full || col0 + 8 * j < p.N)
              up[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]), pack_bf16(f[8 * j + 4], f[8 * j + 5]),
                                 pack_bf16(f[8 * j + 6], f[8 * j + 7]));
        }
        if (p.act_quick == 2) {
#pragma unroll
          for (int j = 0; j < CW; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (p.act_quick) {
#pragma unroll
          for (int j = 0; j < CW; ++j) f[j] = gelu_quick_fwd(f[j]);
        } else {
#pragma unroll
          for (int j = 0; j < CW; ++j) f[j] = gelu_erf_fwd(f[j]);
        }
      } else if (need_aux) {
        const uint32_t* aw = reinterpret_cast<const uint32_t*>(ax);
        if (p.epi == VL_EPI_RESIDUAL) {
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) {
            f[2 * j] += bf16_lo(aw[j]);
            f[2 * j + 1] += bf16_hi(aw[j]);
          }
        } else if (p.act_quick == 2) {
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) {
            f[2 * j] = bf16_lo(aw[j]) > 0.f ? f[2 * j] : 0.f;
            f[2 * j + 1] = bf16_hi(aw[j]) > 0.f ? f[2 * j + 1] : 0.f;
          }
        } else if (p.act_quick) {
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) {
            f[2 * j] *= gelu_quick_grad(bf16_lo(aw[j]));
            f[2 * j + 1] *= gelu_quick_grad(bf16_hi(aw[j]));
          }
        } else {
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) {
            f[2 * j] *= gelu_erf_grad(bf16_lo(aw[j]));
            f[2 * j + 1] *= gelu_erf_grad(bf16_hi(aw[j]));
          }
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < CW; ++j) f[j] = fmaxf(f[j], 0.f);
      }
    }
    // ---- store
    if (row_ok) {
      const long long doff = static_cast<long long>(row) * p.ldd + col0;
      if (p.d_f32) {
        float* dp = reinterpret_cast<float*>(p.d) + doff;
        if (p.accumulate) {
#pragma unroll
          for (int j = 0; j < CW; j += 4)
            if (col0 + j < p.N) red_add_v4_f32(dp + j, f[j], f[j + 1], f[j + 2], f[j + 3]);  // N % 4 == 0 and 16-byte aligned rows (checked at launch)
        } else {
#pragma unroll
          for (int j = 0; j < CW; j += 4)
            if (col0 + j < p.N) *reinterpret_cast<float4*>(dp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        }
      } else {
        uint4* dp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.d) + doff);
#pragma unroll
        for (int j = 0; j < CW / 8; ++j)
          if (full || col0 + 8 * j < p.N)
            dp[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]), pack_bf16(f[8 * j + 4], f[8 * j + 5]),
                               pack_bf16(f[8 * j + 6], f[8 * j + 7]));
      }
    }
  }
  if (p.epi == VL_EPI_CLIPGRAD && p.scalar_out != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) clip_ds += __shfl_xor_sync(0xffffffffu, clip_ds, o);
    if (lane == 0) p.scalar_out[part_idx] = clip_ds;  // one partial per (tile, epilogue warp); summed in order by launch_colreduce
  }
}


// TMA-staged epilogue for bf16 outputs (LINEAR / GELU / RESIDUAL / GELU_BWD).  The warp owns rows [row_w, row_w + 32) x
// columns [col_s, col_s + 64) of the output; `buf` is its private 4 KB staging block (layout = TMA SWIZZLE_128B: 16-byte
// chunk j of row r lives at r*128 + ((j ^ (r & 7)) << 4), conflict-free for per-row 16-byte accesses).  When the mode
// needs an aux block the caller has already issued its TMA load into `buf` (completion on aux_bar / aux_phase).
// `release` hands the TMEM accumulator back as soon as the last tcgen05.ld has landed, before any store is issued.
template <int BN, int EPI, typename Release>
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, const CUtensorMap* tmD, const CUtensorMap* tmX, uint32_t tmem_base,
                                                  uint32_t acc_col, int row_w, int col_s, int slice, int quarter, int lane, uint32_t buf,
                                                  uint32_t buf2, uint32_t aux_bar, uint32_t aux_phase, bool aux_loaded, bool last_slice,
                                                  Release release) {
  const int epi = EPI >= 0 ? EPI : p.epi;  // compile-time constant in the specialised kernels
  const float alpha_eff = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.0f);
  const bool need_aux = epi == VL_EPI_RESIDUAL || epi == VL_EPI_GELU_BWD;
  const bool need_bias = p.bias != nullptr && epi != VL_EPI_GELU_BWD;
  const bool keep_pre = epi == VL_EPI_GELU && p.aux_out != nullptr;
  const uint32_t sw = lane & 7;
  // GELU forward issues two stores per slice; when the warp owns a second staging block (buf2 != buf) the stores alternate
  // between the two so that a block is only rewritten two stores later (no wait on the store that was just issued).
  const bool ring = keep_pre && buf2 != buf;
  if (ring) {
    if (lane == 0) tma_store_wait_read<1>();
    __syncwarp();
  }
  const uint32_t row_addr = buf + lane * 128;
  uint32_t pre[32];  // packed pre-activations of the 64 columns (GELU forward keeps them for the second store)
  float bnext[16];
  const bool full64 = col_s + 64 <= p.N;
  auto load_bias = [&](int c, float (&bv)[16]) {
    const int col0 = col_s + c * 16;
    if (full64) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b4 = (col0 + j < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j)) : make_float4(0, 0, 0, 0);
        bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
      }
    }
  };
  if (need_bias) load_bias(0, bnext);
  const uint32_t tsrc = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc_col + slice * 64;
  uint32_t vnext[16];
  tmem_ld16(tsrc, vnext);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t v[16];
    float bv[16];
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = vnext[j];
    if (c < 3) tmem_ld16(tsrc + (c + 1) * 16, vnext);  // next chunk's accumulators stream in during this chunk's math
    if (c == 3 && last_slice) release();
    if (need_bias) {
#pragma unroll
      for (int j = 0; j < 16; ++j) bv[j] = bnext[j];
      if (c < 3) load_bias(c + 1, bnext);  // likewise the next chunk's bias
    }
    float f[16];
    if (need_bias) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 r = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), make_float2(alpha_eff, alpha_eff), make_float2(bv[j], bv[j + 1]));
        f[j] = r.x;
        f[j + 1] = r.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 r = __fmul2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), make_float2(alpha_eff, alpha_eff));
        f[j] = r.x;
        f[j + 1] = r.y;
      }
    }
    const uint32_t a0 = row_addr + (((2 * c) ^ sw) << 4), a1 = row_addr + (((2 * c + 1) ^ sw) << 4);
    if (epi == VL_EPI_GELU) {
      if (keep_pre) {
#pragma unroll
        for (int j = 0; j < 8; ++j) pre[8 * c + j] = pack_bf16(f[2 * j], f[2 * j + 1]);
      }
      if (p.act_quick == 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
      } else if (p.act_quick) {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = gelu_quick_fwd(f[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {  // two elements per FMA-pipe instruction
          const float2 r = gelu_erf_fwd2(make_float2(f[j], f[j + 1]));
          f[j] = r.x;
          f[j + 1] = r.y;
        }
      }
    } else if (need_aux) {
      if (c == 0 && aux_loaded) mbar_wait(aux_bar, aux_phase);
      const uint4 x0 = ld_shared_v4(a0), x1 = ld_shared_v4(a1);
      const uint32_t aw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      if (epi == VL_EPI_RESIDUAL) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[2 * j] += bf16_lo(aw[j]);
          f[2 * j + 1] += bf16_hi(aw[j]);
        }
      } else if (p.act_quick == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[2 * j] = bf16_lo(aw[j]) > 0.f ? f[2 * j] : 0.f;
          f[2 * j + 1] = bf16_hi(aw[j]) > 0.f ? f[2 * j + 1] : 0.f;
        }
      } else if (p.act_quick) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[2 * j] *= gelu_quick_grad(bf16_lo(aw[j]));
          f[2 * j + 1] *= gelu_quick_grad(bf16_hi(aw[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 r = __fmul2_rn(make_float2(f[2 * j], f[2 * j + 1]), gelu_erf_grad2(make_float2(bf16_lo(aw[j]), bf16_hi(aw[j]))));
          f[2 * j] = r.x;
          f[2 * j + 1] = r.y;
        }
      }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    st_shared_v4(a0, pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    st_shared_v4(a1, pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0 && !(p.dbg & 2)) {
    tma_store_2d(tmD, buf, col_s, row_w);
    tma_store_commit();
  }
  if (keep_pre) {
    if (lane == 0) {
      if (ring) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
    }
    __syncwarp();
    const uint32_t row2 = (ring ? buf2 : buf) + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) st_shared_v4(row2 + ((j ^ sw) << 4), pre[4 * j], pre[4 * j + 1], pre[4 * j + 2], pre[4 * j + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && !(p.dbg & 2)) {
      tma_store_2d(tmX, ring ? buf2 : buf, col_s, row_w);
      tma_store_commit();
    }
  }
}

// Fused GEGLU epilogue (Lens FeedForward, reference perceiver.py:85-102: Linear(d, 2F) -> chunk -> value * gelu(gate)).  The
// weight rows are PERMUTED on the host so that every 256-column tile holds the value columns [n*128, n*128 + 128) in its first
// half and the matching gate columns in its second half; the warp that owns value slice s (64 columns) also owns gate slice
// s + 2.  It writes three 32 x 64 blocks: the two pre-activation blocks back to h (the [M, 2F] tensor the backward needs, in
// the ORIGINAL column order: value at column c, gate at F + c) and value * gelu(gate) to the [M, F] output.
template <int BN, typename Release>
__device__ __forceinline__ void epilogue_tile_geglu(const GemmParams& p, const CUtensorMap* tmD, const CUtensorMap* tmX, uint32_t tmem_base,
                                                    uint32_t acc_col, int row_w, int n_blk, int s, int quarter, int lane, uint32_t bufA,
                                                    uint32_t bufB, Release release) {
  const int F = p.N / 2;
  const int out_col = n_blk * (BN / 2) + s * 64;  // column in the output and in the value half of h
  const int gcol = n_blk * BN + s * 64;           // GEMM column of the value slice (gate slice: + BN / 2)
  const float alpha_eff = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.0f);
  const uint32_t sw = lane & 7;
  const uint32_t rowA = bufA + lane * 128, rowB = bufB + lane * 128;
  const uint32_t tval = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc_col + s * 64;
  const uint32_t tgate = tval + BN / 2;
  uint32_t outp[32];  // value * gelu(gate) of the 64 columns, packed bf16
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t v[16], g[16];
    tmem_ld16(tval + c * 16, v);
    tmem_ld16(tgate + c * 16, g);
    float bv[16], bg[16];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 x4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + gcol + c * 16 + j)) : make_float4(0, 0, 0, 0);
      const float4 y4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + gcol + BN / 2 + c * 16 + j)) : make_float4(0, 0, 0, 0);
      bv[j] = x4.x; bv[j + 1] = x4.y; bv[j + 2] = x4.z; bv[j + 3] = x4.w;
      bg[j] = y4.x; bg[j + 1] = y4.y; bg[j + 2] = y4.z; bg[j + 3] = y4.w;
    }
    tc_wait_ld();
    if (c == 3) release();
    float fv[16], fg[16];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float2 a2 = make_float2(alpha_eff, alpha_eff);
      const float2 r = __ffma2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), a2, make_float2(bv[j], bv[j + 1]));
      const float2 q = __ffma2_rn(make_float2(__uint_as_float(g[j]), __uint_as_float(g[j + 1])), a2, make_float2(bg[j], bg[j + 1]));
      fv[j] = r.x; fv[j + 1] = r.y;
      fg[j] = q.x; fg[j + 1] = q.y;
    }
    const uint32_t o0 = ((2 * c) ^ sw) << 4, o1 = ((2 * c + 1) ^ sw) << 4;
    st_shared_v4(rowA + o0, pack_bf16(fv[0], fv[1]), pack_bf16(fv[2], fv[3]), pack_bf16(fv[4], fv[5]), pack_bf16(fv[6], fv[7]));
    st_shared_v4(rowA + o1, pack_bf16(fv[8], fv[9]), pack_bf16(fv[10], fv[11]), pack_bf16(fv[12], fv[13]), pack_bf16(fv[14], fv[15]));
    st_shared_v4(rowB + o0, pack_bf16(fg[0], fg[1]), pack_bf16(fg[2], fg[3]), pack_bf16(fg[4], fg[5]), pack_bf16(fg[6], fg[7]));
    st_shared_v4(rowB + o1, pack_bf16(fg[8], fg[9]), pack_bf16(fg[10], fg[11]), pack_bf16(fg[12], fg[13]), pack_bf16(fg[14], fg[15]));
#pragma unroll
    for (int j = 0; j < 16; j += 2) {  // value * gelu(gate), two elements per FMA-pipe instruction
      const float2 act = gelu_erf_fwd2(make_float2(fg[j], fg[j + 1]));
      const float2 o = __fmul2_rn(make_float2(fv[j], fv[j + 1]), act);
      outp[8 * c + (j >> 1)] = pack_bf16(o.x, o.y);
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0 && !(p.dbg & 2)) {
    tma_store_2d(tmX, bufA, out_col, row_w);      // value pre-activations  -> h[:, c]
    tma_store_commit();
    tma_store_2d(tmX, bufB, F + out_col, row_w);  // gate pre-activations   -> h[:, F + c]
    tma_store_commit();
    tma_store_wait_read<1>();                     // the first store has finished reading block A: reuse it for the output
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) st_shared_v4(rowA + ((j ^ sw) << 4), outp[4 * j], outp[4 * j + 1], outp[4 * j + 2], outp[4 * j + 3]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0 && !(p.dbg & 2)) {
    tma_store_2d(tmD, bufA, out_col, row_w);
    tma_store_commit();
  }
}

// PEER = true: the B operand is sharded by rows over the ranks of the box (the all-gathered feature matrix of the contrastive
// loss, reference loss.py:55-76): one tensor map per peer arena, and the producer loads each tile straight from its owner's
// memory over NVLink after seeing that peer's ticket -- the all-gather happens inside the GEMM, tile by tile.
struct PeerMaps {
  CUtensorMap m[8];
};
struct NoPeerMaps {
  int unused;
};
__device__ __forceinline__ int ld_acquire_sys_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int BN, int EW, int EPI = -1, bool PEER = false>
__global__ void __launch_bounds__((KernelCfg<BN, EW>::kThreads), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                 const __grid_constant__ CUtensorMap tmX, const GemmParams p,
                 const __grid_constant__ typename std::conditional<PEER, PeerMaps, NoPeerMaps>::type pmB) {
  constexpr bool TE = EW > 0;
  constexpr int kEpiWarps = KernelCfg<BN, EW>::kEpiWarps;
  using Cfg = GemmCfg<BN, TE>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = stg_base + Cfg::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  auto aux_bar = [&](int e) { return bar_base + 8u * (2 * Cfg::kStages + 5 + e); };
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n * p.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TE) {
      tma_prefetch_desc(&tmD);
      tma_prefetch_desc(&tmX);
      for (int e = 0; e < 16; ++e) mbar_init(aux_bar(e), 1);
    }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      [[maybe_unused]] uint32_t peers_seen = 0;
      // PEER: the tensor map and local row of B's global row `row`; the first touch of a peer waits for its ticket
      [[maybe_unused]] auto peer_map = [&](int row, int& local_row) -> const CUtensorMap* {
        if constexpr (PEER) {
          const int pr = row / p.b_peer_rows;
          local_row = row - pr * p.b_peer_rows;
          if (p.peer_flags != nullptr && !((peers_seen >> pr) & 1u)) {
            const long long t0 = clock64();
            unsigned spins = 0;
            while (ld_acquire_sys_s32(p.peer_flags + pr) < p.peer_flag_value) {
              __nanosleep(100);
              if ((++spins & 1023u) == 0 && clock64() - t0 > 40000000000ll) __trap();  // a peer that never publishes fails the launch
            }
            asm volatile("fence.proxy.async.global;" ::: "memory");  // the TMA (async proxy) reads that follow see the payload
            peers_seen |= 1u << pr;
          }
          return &pmB.m[pr];
        } else {
          local_row = row;
          return &tmB;
        }
      };
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m_blk, n_blk, ks;
        tile_coords(p, t, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
          if (!p.a_mn) {
            tma_load_2d(sa, &tmA, full_bar(stage), kb * kBK, m_blk * kBM);
          } else {
#pragma unroll
            for (int c = 0; c < kBM / 64; ++c)
              tma_load_2d(sa + c * (kBK * 128), &tmA, full_bar(stage), m_blk * kBM + c * 64, kb * kBK);
          }
          if (!p.b_mn) {
            int lrow;
            const CUtensorMap* mb = peer_map(n_blk * BN, lrow);  // (host: b_peer_rows % BN == 0, a tile never straddles two peers)
            tma_load_2d(sb, mb, full_bar(stage), kb * kBK, lrow);
          } else {
            int lrow;
            const CUtensorMap* mb = peer_map(kb * kBK, lrow);    // rows of an MN-major B are K indices (host: b_peer_rows % 64 == 0)
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * (kBK * 128), mb, full_bar(stage), n_blk * BN + c * 64, lrow);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(kBM, BN, p.a_mn, p.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int m_blk, n_blk, ks;
        tile_coords(p, t, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!(p.dbg & 8)) mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t ad = umma_desc_sw128(sa + k * p.a_kstep, p.a_lbo, p.a_sbo);
            const uint64_t bd = umma_desc_sw128(sb + k * p.b_kstep, p.b_lbo, p.b_sbo);
            umma_ss(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // smem stage reusable once these MMAs retire
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int e = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    [[maybe_unused]] uint32_t aux_phase[4] = {0, 0, 0, 0};  // per owned slice: flips only when that slice's block was loaded
    [[maybe_unused]] const int epi = EPI >= 0 ? EPI : p.epi;
    [[maybe_unused]] const bool need_aux = epi == VL_EPI_RESIDUAL || epi == VL_EPI_GELU_BWD;
    constexpr int kGroups = kEpiWarps / 4;             // warps per TMEM lane quarter
    [[maybe_unused]] constexpr int kSlicesPerWarp = (BN / 64) / kGroups;  // 64-column slices each warp owns (TE path)
    // GELU forward (two stores per slice): the warp's two staging blocks form a ring, activation -> first, pre-activation ->
    // second, for both of its slices (see epilogue_tile_tma).
    [[maybe_unused]] const bool ring2 = kSlicesPerWarp > 1 && epi == VL_EPI_GELU && p.aux_out != nullptr;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int m_blk, n_blk, ks;
      tile_coords(p, t, m_blk, n_blk, ks);
      const int row_base = m_blk * kBM;
      [[maybe_unused]] const int row_w = row_base + quarter * 32;
      if constexpr (TE) {
        if (lane == 0 && row_w < p.M) {
          tma_store_wait_read<0>();  // the previous tile's stores have finished reading the staging blocks
          if (need_aux && !(p.dbg & 3)) {
#pragma unroll
            for (int si = 0; si < kSlicesPerWarp; ++si) {
              const int slice = (e >> 2) + si * kGroups;
              const int col_s = n_blk * BN + slice * 64;
              if (col_s < p.N) {
                mbar_expect_tx(aux_bar(slice * 4 + quarter), kEpiBufBytes);
                tma_load_2d(stg_base + (slice * 4 + quarter) * kEpiBufBytes, &tmX, aux_bar(slice * 4 + quarter), col_s, row_w);
              }
            }
          }
        }
        __syncwarp();
      }
      if (p.dbg & 4) {
        while (!mbar_try_wait(tfull_bar(acc), acc_phase)) __nanosleep(64);
      } else {
        mbar_wait(tfull_bar(acc), acc_phase);
      }
      tc_fence_after();
      auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      };
      if constexpr (TE) {
        bool released = false;
        if (row_w < p.M && !(p.dbg & 1)) {
#pragma unroll
          for (int si = 0; si < kSlicesPerWarp; ++si) {
            const int slice = (e >> 2) + si * kGroups;
            const int col_s = n_blk * BN + slice * 64;
            if (col_s < p.N) {
              const bool last = (si == kSlicesPerWarp - 1) || (col_s + kGroups * 64 >= p.N);
              epilogue_tile_tma<BN, EPI>(p, &tmD, &tmX, tmem_base, acc * BN, row_w, col_s, slice, quarter, lane,
                                    stg_base + ((ring2 ? (e >> 2) : slice) * 4 + quarter) * kEpiBufBytes,
                                    stg_base + ((ring2 ? (e >> 2) + kGroups : slice) * 4 + quarter) * kEpiBufBytes,
                                    aux_bar(slice * 4 + quarter), aux_phase[si],
                                    need_aux && !(p.dbg & 3), last, release);
              released |= last;
              if (need_aux && !(p.dbg & 3)) aux_phase[si] ^= 1;
            }
          }
        }
        if (!released) release();
      } else {
        if (!(p.dbg & 1)) epilogue_tile<BN, EpiCfg<BN>::kChunk>(p, tmem_base, acc * BN, row_base, n_blk, ks, e, quarter, lane, t * kEpiWarps + e);
        release();  // accumulator drained -> hand the TMEM buffer back to the MMA warp
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if constexpr (TE) {
      if (lane == 0) tma_store_wait<0>();  // all bulk stores complete before the CTA exits
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
static int launch_gemm2(const VlGemmArgs& a, GemmParams p, cudaStream_t stream);

// The TMA-staged epilogue applies to bf16 outputs of the four activation-path modes.  debug knob 10: 1 = force the
// per-thread global-store epilogue.
template <int BN>
static bool tma_epilogue_ok(const VlGemmArgs& a, const GemmParams& p) {
  return BN == 256 && !a.d_f32 && !a.accumulate && p.split_k == 1 && p.aux_row_div == 1 &&
         (a.epilogue == VL_EPI_LINEAR || a.epilogue == VL_EPI_GELU || a.epilogue == VL_EPI_RESIDUAL || a.epilogue == VL_EPI_GELU_BWD ||
          a.epilogue == VL_EPI_GEGLU) &&
         (reinterpret_cast<uintptr_t>(a.d) & 15) == 0 && debug_get(10) != 1;
}

static int make_epilogue_maps(const VlGemmArgs& a, CUtensorMap* tmD, CUtensorMap* tmX) {
  if (a.epilogue == VL_EPI_GEGLU) {  // output [M, N / 2]; the pre-activations go to aux_out [M, N]
    int rc = make_tmap_bf16_2d(tmD, a.d, a.N / 2, a.M, a.ldd, 64, 32);
    if (rc) return rc;
    VL_CHECK_ARG((reinterpret_cast<uintptr_t>(a.aux_out) & 15) == 0, "vl_gemm_bf16: aux pointer must be 16-byte aligned");
    return make_tmap_bf16_2d(tmX, a.aux_out, a.N, a.M, a.ldaux, 64, 32);
  }
  int rc = make_tmap_bf16_2d(tmD, a.d, a.N, a.M, a.ldd, 64, 32);
  if (rc) return rc;
  const void* xptr = (a.epilogue == VL_EPI_GELU) ? a.aux_out : a.aux_in;
  if (xptr != nullptr && a.epilogue != VL_EPI_LINEAR) {
    VL_CHECK_ARG((reinterpret_cast<uintptr_t>(xptr) & 15) == 0, "vl_gemm_bf16: aux pointer must be 16-byte aligned");
    return make_tmap_bf16_2d(tmX, xptr, a.N, a.M, a.ldaux, 64, 32);
  }
  *tmX = *tmD;
  return 0;
}

// debug knob 8: 0 = default kernel choice, 1 = force the single-CTA kernel, 2 = force the CTA-pair kernel.
// Measured on B200 (profiles/r01_gemm_epilogue_attribution.log): with the TMA-staged epilogue the pair kernel wins on every
// ViT-L shape (its main loop alone sustains 1.55-1.62 PFLOP/s against 1.44-1.49 for the single-CTA kernel), so it is the
// default whenever there are at least two 256-row tiles; the contrastive-loss epilogues stay on the single-CTA kernel.

template <int BN>
static int launch_gemm(const VlGemmArgs& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  GemmParams p;
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.a_mn = a.a_mn ? 1 : 0;
  p.b_mn = a.b_mn ? 1 : 0;
  p.tiles_m = (a.M + kBM - 1) / kBM;
  p.tiles_n = (a.N + BN - 1) / BN;
  p.split_k = a.split_k < 1 ? 1 : a.split_k;
  p.kb_total = (a.K + kBK - 1) / kBK;
  if (p.split_k > p.kb_total) p.split_k = p.kb_total;
  p.kb_per_split = (p.kb_total + p.split_k - 1) / p.split_k;
  p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.epi = a.epilogue;
  p.act_quick = a.act_quick;
  p.d_f32 = a.d_f32;
  p.accumulate = a.accumulate;
  p.alpha = a.alpha;
  p.d = a.d;
  p.ldd = a.ldd;
  p.bias = a.bias;
  p.aux_in = reinterpret_cast<const __nv_bfloat16*>(a.aux_in);
  p.aux_out = reinterpret_cast<__nv_bfloat16*>(a.aux_out);
  p.ldaux = a.ldaux;
  p.row_vec = a.row_vec;
  p.col_vec = a.col_vec;
  p.out_vec0 = a.out_vec0;
  p.out_vec1 = a.out_vec1;
  p.out_vec2 = a.out_vec2;
  p.scalar_out = a.scalar_out;
  p.iparam = a.iparam;
  p.fparam = a.fparam;
  p.alpha_dev = a.alpha_dev;
  p.fparam_dev = a.fparam_dev;
  p.aux_row_div = a.aux_row_div > 1 ? a.aux_row_div : 1;
  p.relu = a.relu;
  p.rowsum_out = a.rowsum_out;
  p.loss_flags = a.loss_flags;
  p.mask = a.mask;
  p.ldmask = a.ldmask;
  p.b_peer_rows = a.b_peer_rows;
  p.peer_flags = a.peer_flags;
  p.peer_flag_value = a.peer_flag_value;
  p.dbg = debug_get(9);
  // K-major: 8-row groups 1024 B apart, +32 B per UMMA_K inside the swizzle row.
  // MN-major: 64-wide chunks kBK*128 B apart (LBO), 8-K groups 1024 B apart (SBO), +2048 B per UMMA_K.
  p.a_lbo = p.a_mn ? kBK * 128 : 16;
  p.a_sbo = 1024;
  p.a_kstep = p.a_mn ? 2048 : 32;
  p.b_lbo = p.b_mn ? kBK * 128 : 16;
  p.b_sbo = 1024;
  p.b_kstep = p.b_mn ? 2048 : 32;
  if (debug_get(1)) p.a_lbo = debug_get(1);
  if (debug_get(2)) p.a_sbo = debug_get(2);
  if (debug_get(3)) p.a_kstep = debug_get(3);
  if (debug_get(4)) p.b_lbo = debug_get(4);
  if (debug_get(5)) p.b_sbo = debug_get(5);
  if (debug_get(6)) p.b_kstep = debug_get(6);

  {
    const int mode = debug_get(8);
    const bool pair_ok = a.M >= 4 * kBM && BN == 256;
    const bool pair_default = a.epilogue != VL_EPI_ROWLSE && a.epilogue != VL_EPI_CLIPGRAD && a.b_peers == nullptr;
    if (mode == 2 || (mode == 0 && pair_default && pair_ok)) return launch_gemm2<BN>(a, p, stream);
  }
  CUtensorMap tmA, tmB, tmD, tmX;
  int rc;
  if (!p.a_mn)
    rc = make_tmap_bf16_2d(&tmA, a.a, a.K, a.M, a.lda, kBK, kBM);
  else
    rc = make_tmap_bf16_2d(&tmA, a.a, a.M, a.K, a.lda, 64, kBK);
  if (rc) return rc;
  PeerMaps pm;
  if (a.b_peers == nullptr) {
    if (!p.b_mn)
      rc = make_tmap_bf16_2d(&tmB, a.b, a.K, a.N, a.ldb, kBK, BN);
    else
      rc = make_tmap_bf16_2d(&tmB, a.b, a.N, a.K, a.ldb, 64, kBK);
    if (rc) return rc;
  } else {
    // one map per peer over ITS rows only: out-of-range rows of a ragged last tile read as zero, never the next allocation
    for (int q = 0; q < a.b_npeers; ++q) {
      if (!p.b_mn)
        rc = make_tmap_bf16_2d(&pm.m[q], a.b_peers[q], a.K, a.b_peer_rows, a.ldb, kBK, BN);
      else
        rc = make_tmap_bf16_2d(&pm.m[q], a.b_peers[q], a.N, a.b_peer_rows, a.ldb, 64, kBK);
      if (rc) return rc;
    }
    for (int q = a.b_npeers; q < 8; ++q) pm.m[q] = pm.m[0];
    tmB = pm.m[0];
  }

  const bool te = tma_epilogue_ok<BN>(a, p) && a.b_peers == nullptr;
  const int total = p.tiles_m * p.tiles_n * p.split_k;
  int grid = total < num_sms() ? total : num_sms();
  if (debug_get(7) > 0 && debug_get(7) < grid) grid = debug_get(7);
  if constexpr (BN == 256) {
    if (te) {
      using CfgT = GemmCfg<BN, true>;
      rc = make_epilogue_maps(a, &tmD, &tmX);
      if (rc) return rc;
      // one kernel per epilogue mode: the mode tests and the dead activation variants drop out of the epilogue's inner loop
#define VL_LAUNCH_TE1(EPI_)                                                                                                          \
  do {                                                                                                                              \
    static bool attr_ = false;                                                                                                      \
    if (!attr_) {                                                                                                                   \
      VL_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, 8, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgT::kSmemBytes));   \
      attr_ = true;                                                                                                                 \
    }                                                                                                                               \
    gemm_bf16_kernel<BN, 8, EPI_><<<grid, KernelCfg<BN, 8>::kThreads, CfgT::kSmemBytes, stream>>>(tmA, tmB, tmD, tmX, p, NoPeerMaps{0}); \
  } while (0)
      switch (a.epilogue) {
        case VL_EPI_LINEAR: VL_LAUNCH_TE1(VL_EPI_LINEAR); break;
        case VL_EPI_GELU: VL_LAUNCH_TE1(VL_EPI_GELU); break;
        case VL_EPI_RESIDUAL: VL_LAUNCH_TE1(VL_EPI_RESIDUAL); break;
        default: VL_LAUNCH_TE1(VL_EPI_GELU_BWD); break;
      }
#undef VL_LAUNCH_TE1
      return launch_check("gemm_bf16_kernel<tma epilogue>");
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    VL_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  // d(loss)/d(alpha) of the CLIPGRAD epilogue: one partial per (tile, epilogue warp), added in order afterwards (deterministic)
  float* ds_part = nullptr;
  const int ds_n = total * KernelCfg<BN, 0>::kEpiWarps;
  if (a.epilogue == VL_EPI_CLIPGRAD && a.scalar_out != nullptr) {
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&ds_part), (size_t)ds_n * sizeof(float), stream))) return rc;
    p.scalar_out = ds_part;
  }
  if (a.b_peers != nullptr) {
    static bool attr_peer = false;
    if (!attr_peer) {
      VL_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, 0, -1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
      attr_peer = true;
    }
    gemm_bf16_kernel<BN, 0, -1, true><<<grid, KernelCfg<BN, 0>::kThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmA, tmA, p, pm);
  } else {
    gemm_bf16_kernel<BN, 0><<<grid, KernelCfg<BN, 0>::kThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmA, tmA, p, NoPeerMaps{0});
  }
  if ((rc = launch_check("gemm_bf16_kernel"))) return rc;
  if (ds_part != nullptr) {
    if ((rc = launch_colreduce(ds_part, ds_n, 1, a.scalar_out, nullptr, nullptr, stream))) return rc;
    return scratch_free(ds_part, stream);
  }
  return 0;
}


// =============================================================================================== CTA-pair variant
// Same roles, but two CTAs of a cluster (one per SM of a TPC) share a 256 x BN tile: each loads its own 128 rows of A
// and only HALF of B (BN/2 rows); the leader issues tcgen05.mma.cta_group::2 (M = 256) which reads both halves and
// writes 128 accumulator rows into each CTA's TMEM.  Per-SM operand traffic (smem fill + MMA reads) drops by 1/3, which
// is what lets the tensor pipe run closer to its peak.  mbarrier protocol:
//   full[s]   (leader's, 1 arrival + tx bytes of both CTAs)   <- leader's producer arms it, both CTAs' TMA complete it
//   empty[s]  (one per CTA)                                   <- leader's tcgen05.commit, multicast to both CTAs
//   tfull[a]  (one per CTA)                                   <- leader's tcgen05.commit, multicast
//   tempty[a] (leader's, 2 x 8 arrivals)                      <- epilogue warps of both CTAs
template <int BN, bool TE = false>
struct Gemm2Cfg {
  static constexpr int kStages = (BN == 256) ? (TE ? 5 : 6) : 8;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / 2) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = TE ? (BN / 16) * kEpiBufBytes : 1024;  // direct-store flavour: 8 x 64 tile of ones (row sums)
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 512;
  static constexpr int kTmemCols = 2 * BN;
};

__device__ __forceinline__ void mbar_wait_guarded(uint32_t bar, uint32_t parity) {
  // bring-up guard: a protocol bug must fail the launch instead of hanging the GPU
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 8000000000ll) __trap();
  }
}

__device__ __forceinline__ void tile_coords2(const GemmParams& p, int tiles_m2, int t, int& m_blk, int& n_blk, int& ks) {
  int mn = t / p.split_k;
  ks = t - mn * p.split_k;
  constexpr int kGroup = kGroupM / 2;
  int group_sz = kGroup * p.tiles_n;
  int g = mn / group_sz;
  int r = mn - g * group_sz;
  int m_first = g * kGroup;
  int gm = min(kGroup, tiles_m2 - m_first);
  n_blk = r / gm;
  m_blk = m_first + (r - n_blk * gm);
}

template <int BN, int EW, int EPI = -1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((KernelCfg<BN, EW>::kThreads), 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                  const __grid_constant__ CUtensorMap tmX, const GemmParams p) {
  constexpr bool TE = EW > 0;
  constexpr int kEpiWarps = KernelCfg<BN, EW>::kEpiWarps;
  using Cfg = Gemm2Cfg<BN, TE>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();  // both CTAs must use identical offsets (the MMA addresses the peer by offset)
  const uint32_t stg_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = stg_base + Cfg::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);
  auto aux_bar = [&](int e) { return bar_base + 8u * (2 * Cfg::kStages + 5 + e); };
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int tiles_m2 = (p.M + 2 * kBM - 1) / (2 * kBM);
  const int total_tiles = tiles_m2 * p.tiles_n * p.split_k;
  // Row sums of A (bias gradient riding on a weight-gradient GEMM): tiles of the first n-column also run an N = 16 MMA of the
  // A tiles against a tile of ones into TMEM columns [BN, BN + 16); the accumulators are then single-buffered.
  const bool rs = !TE && p.rowsum_out != nullptr;
  const int nacc = rs ? 1 : 2;
  if (rs && threadIdx.x >= 64 && threadIdx.x < 128) {
    st_shared_v4(stg_base + (threadIdx.x - 64) * 16, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);  // bf16 1.0 x 8
    fence_proxy_async_smem();
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (TE) {
      tma_prefetch_desc(&tmD);
      tma_prefetch_desc(&tmX);
      for (int e = 0; e < 16; ++e) mbar_init(aux_bar(e), 1);
    }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);  // the leader's arrive.expect_tx; the peer only contributes transaction bytes
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        int m_blk, n_blk, ks;
        tile_coords2(p, tiles_m2, t, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int row_a = m_blk * 2 * kBM + rank * kBM;
        const int row_b = n_blk * BN + rank * (BN / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_guarded(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
          // No arrival from the peer: a cluster-scope release per stage would serialise its producer.  The peer can
          // only refill stage s after the leader's MMA consumed it (empty[s] is signalled by the leader's commit), so
          // its bytes always land in the phase the leader is arming.
          if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
          if (!p.a_mn) {
            tma_load_2d_pair(sa, &tmA, lead_full, kb * kBK, row_a);
          } else {
#pragma unroll
            for (int c = 0; c < kBM / 64; ++c) tma_load_2d_pair(sa + c * (kBK * 128), &tmA, lead_full, row_a + c * 64, kb * kBK);
          }
          if (!p.b_mn) {
            tma_load_2d_pair(sb, &tmB, lead_full, kb * kBK, row_b);
          } else {
#pragma unroll
            for (int c = 0; c < (BN / 2) / 64; ++c) tma_load_2d_pair(sb + c * (kBK * 128), &tmB, lead_full, row_b + c * 64, kb * kBK);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(2 * kBM, BN, p.a_mn, p.b_mn);
      const uint32_t idesc_ones = umma_idesc_bf16(2 * kBM, 16, p.a_mn, 0);
      const uint64_t ones_desc = umma_desc_sw128(stg_base, 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        int m_blk, n_blk, ks;
        tile_coords2(p, tiles_m2, t, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        mbar_wait_guarded(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        bool rs_started = false;
        int rs_wait = rs ? ((n_blk - kb0) % p.tiles_n + p.tiles_n) % p.tiles_n : -1;  // k-blocks until this tile's next row-sum turn
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!(p.dbg & 8)) mbar_wait_guarded(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t ad = umma_desc_sw128(sa + k * p.a_kstep, p.a_lbo, p.a_sbo);
            const uint64_t bd = umma_desc_sw128(sb + k * p.b_kstep, p.b_lbo, p.b_sbo);
            umma_ss_pair(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // Row sums of A: the n-tiles of one m-row share the work (k-block kb belongs to n-tile kb % tiles_n), four MMAs back
          // to back so the pipe switches instruction shape twice per k-block; partial sums meet in rowsum_out (atomics).
          if (rs_wait-- == 0) {
            rs_wait = p.tiles_n - 1;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t ad = umma_desc_sw128(sa + k * p.a_kstep, p.a_lbo, p.a_sbo);
              umma_ss_pair(tmem_base + BN, ad, ones_desc, idesc_ones, (rs_started || k > 0) ? 1u : 0u);
            }
            rs_started = true;
          }
          umma_commit_pair_mc(empty_bar(stage), 3);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair_mc(tfull_bar(acc), 3);
        if (++acc == nacc) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows)
    const int e = warp - 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    [[maybe_unused]] uint32_t aux_phase[4] = {0, 0, 0, 0};  // per owned slice: flips only when that slice's block was loaded
    [[maybe_unused]] const int epi = EPI >= 0 ? EPI : p.epi;
    [[maybe_unused]] const bool need_aux = epi == VL_EPI_RESIDUAL || epi == VL_EPI_GELU_BWD;
    constexpr int kGroups = kEpiWarps / 4;             // warps per TMEM lane quarter
    [[maybe_unused]] constexpr int kSlicesPerWarp = (BN / 64) / kGroups;  // 64-column slices each warp owns (TE path)
    // GELU forward (two stores per slice): the warp's two staging blocks form a ring, activation -> first, pre-activation ->
    // second, for both of its slices (see epilogue_tile_tma).
    [[maybe_unused]] const bool ring2 = kSlicesPerWarp > 1 && epi == VL_EPI_GELU && p.aux_out != nullptr;
    for (int t = cluster_id; t < total_tiles; t += num_clusters) {
      int m_blk, n_blk, ks;
      tile_coords2(p, tiles_m2, t, m_blk, n_blk, ks);
      const int row_base = m_blk * 2 * kBM + static_cast<int>(rank) * kBM;
      [[maybe_unused]] const int row_w = row_base + quarter * 32;
      if constexpr (TE) {
        if (lane == 0 && row_w < p.M) {
          tma_store_wait_read<0>();  // the previous tile's stores have finished reading the staging blocks
          if (need_aux && !(p.dbg & 3)) {
#pragma unroll
            for (int si = 0; si < kSlicesPerWarp; ++si) {
              const int slice = (e >> 2) + si * kGroups;
              const int col_s = n_blk * BN + slice * 64;
              if (col_s < p.N) {
                mbar_expect_tx(aux_bar(slice * 4 + quarter), kEpiBufBytes);
                tma_load_2d(stg_base + (slice * 4 + quarter) * kEpiBufBytes, &tmX, aux_bar(slice * 4 + quarter), col_s, row_w);
              }
            }
          }
        }
        __syncwarp();
      }
      if (p.dbg & 4) {
        while (!mbar_try_wait(tfull_bar(acc), acc_phase)) __nanosleep(64);
      } else {
        mbar_wait_guarded(tfull_bar(acc), acc_phase);
      }
      tc_fence_after();
      auto release = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(acc), 0));
      };
      if constexpr (TE && EPI == VL_EPI_GEGLU) {
        static_assert(!TE || (kSlicesPerWarp == 2 && kGroups == 2), "GEGLU epilogue: 8 warps, value slice s and gate slice s + 2 per warp");
        bool released = false;
        if (row_w < p.M && !(p.dbg & 1)) {
          const int s = e >> 2;
          epilogue_tile_geglu<BN>(p, &tmD, &tmX, tmem_base, acc * BN, row_w, n_blk, s, quarter, lane, stg_base + (s * 4 + quarter) * kEpiBufBytes,
                                  stg_base + ((s + kGroups) * 4 + quarter) * kEpiBufBytes, release);
          released = true;
        }
        if (!released) release();
      } else if constexpr (TE) {
        bool released = false;
        if (row_w < p.M && !(p.dbg & 1)) {
#pragma unroll
          for (int si = 0; si < kSlicesPerWarp; ++si) {
            const int slice = (e >> 2) + si * kGroups;
            const int col_s = n_blk * BN + slice * 64;
            if (col_s < p.N) {
              const bool last = (si == kSlicesPerWarp - 1) || (col_s + kGroups * 64 >= p.N);
              epilogue_tile_tma<BN, EPI>(p, &tmD, &tmX, tmem_base, acc * BN, row_w, col_s, slice, quarter, lane,
                                    stg_base + ((ring2 ? (e >> 2) : slice) * 4 + quarter) * kEpiBufBytes,
                                    stg_base + ((ring2 ? (e >> 2) + kGroups : slice) * 4 + quarter) * kEpiBufBytes,
                                    aux_bar(slice * 4 + quarter), aux_phase[si],
                                    need_aux && !(p.dbg & 3), last, release);
              released |= last;
              if (need_aux && !(p.dbg & 3)) aux_phase[si] ^= 1;
            }
          }
        }
        if (!released) release();
      } else {
        if (!(p.dbg & 1)) epilogue_tile<BN, EpiCfg<BN>::kChunk>(p, tmem_base, acc * BN, row_base, n_blk, ks, e, quarter, lane, t * kEpiWarps + e);
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        const int kb_first = kb0 + ((n_blk - kb0) % p.tiles_n + p.tiles_n) % p.tiles_n;  // first k-block this tile summed
        if (rs && (e >> 2) == 0) {  // one warp per lane quarter stores this tile's partial row sums of A (0 if it summed no k-block)
          uint32_t v[16];
          v[0] = 0u;
          if (kb_first < kb1) {
            tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + BN, v);
            tc_wait_ld();
          }
          const int row = row_base + quarter * 32 + lane;
          if (row < p.M) p.rowsum_out[static_cast<long long>(n_blk * p.split_k + ks) * p.M + row] = __uint_as_float(v[0]);
        }
        release();  // accumulator drained -> hand the TMEM buffer back to the MMA warp
      }
      if (++acc == nacc) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if constexpr (TE) {
      if (lane == 0) tma_store_wait<0>();  // all bulk stores complete before the CTA exits
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be signalling barriers / reading operands in this CTA's smem
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
static int launch_gemm2(const VlGemmArgs& a, GemmParams p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN, false>;
  // MN-major chunk strides are the same; only B's per-CTA row count halves
  CUtensorMap tmA, tmB, tmD, tmX;
  int rc;
  if (!p.a_mn)
    rc = make_tmap_bf16_2d(&tmA, a.a, a.K, a.M, a.lda, kBK, kBM);
  else
    rc = make_tmap_bf16_2d(&tmA, a.a, a.M, a.K, a.lda, 64, kBK);
  if (rc) return rc;
  if (!p.b_mn)
    rc = make_tmap_bf16_2d(&tmB, a.b, a.K, a.N, a.ldb, kBK, BN / 2);
  else
    rc = make_tmap_bf16_2d(&tmB, a.b, a.N, a.K, a.ldb, 64, kBK);
  if (rc) return rc;
  const int tiles_m2 = (a.M + 2 * kBM - 1) / (2 * kBM);
  const int total = tiles_m2 * p.tiles_n * p.split_k;
  int clusters = num_sms() / 2;
  if (total < clusters) clusters = total;
  if (debug_get(7) > 0 && debug_get(7) < clusters) clusters = debug_get(7);
  if constexpr (BN == 256) {
    if (tma_epilogue_ok<BN>(a, p)) {
      using CfgT = Gemm2Cfg<BN, true>;
      rc = make_epilogue_maps(a, &tmD, &tmX);
      if (rc) return rc;
#define VL_LAUNCH_TE2(EPI_)                                                                                                          \
  do {                                                                                                                              \
    static bool attr_ = false;                                                                                                      \
    if (!attr_) {                                                                                                                   \
      VL_CUDA(cudaFuncSetAttribute(gemm2_bf16_kernel<BN, 8, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgT::kSmemBytes));  \
      attr_ = true;                                                                                                                 \
    }                                                                                                                               \
    gemm2_bf16_kernel<BN, 8, EPI_><<<2 * clusters, KernelCfg<BN, 8>::kThreads, CfgT::kSmemBytes, stream>>>(tmA, tmB, tmD, tmX, p);  \
  } while (0)
      switch (a.epilogue) {
        case VL_EPI_LINEAR: VL_LAUNCH_TE2(VL_EPI_LINEAR); break;
        case VL_EPI_GELU: VL_LAUNCH_TE2(VL_EPI_GELU); break;
        case VL_EPI_RESIDUAL: VL_LAUNCH_TE2(VL_EPI_RESIDUAL); break;
        case VL_EPI_GEGLU: VL_LAUNCH_TE2(VL_EPI_GEGLU); break;
        default: VL_LAUNCH_TE2(VL_EPI_GELU_BWD); break;
      }
#undef VL_LAUNCH_TE2
      return launch_check("gemm2_bf16_kernel<tma epilogue>");
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    VL_CUDA(cudaFuncSetAttribute(gemm2_bf16_kernel<BN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  // row sums of A (bias gradient): one partial row per (n-tile, split), added in order afterwards (deterministic)
  float* rs_part = nullptr;
  const int rs_n = p.tiles_n * p.split_k;
  if (a.rowsum_out != nullptr) {
    if ((rc = scratch_alloc(reinterpret_cast<void**>(&rs_part), (size_t)rs_n * a.M * sizeof(float), stream))) return rc;
    p.rowsum_out = rs_part;
  }
  gemm2_bf16_kernel<BN, 0><<<2 * clusters, KernelCfg<BN, 0>::kThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, tmA, tmA, p);
  if ((rc = launch_check("gemm2_bf16_kernel"))) return rc;
  if (rs_part != nullptr) {
    if ((rc = launch_colreduce(rs_part, rs_n, a.M, a.rowsum_out, nullptr, nullptr, stream))) return rc;
    return scratch_free(rs_part, stream);
  }
  return 0;
}

}  // namespace vl

extern "C" int vl_gemm_bf16(const VlGemmArgs* a, void* stream) {
  using namespace vl;
  VL_CHECK_ARG(a != nullptr && a->a && (a->b || a->b_peers) && (a->d || a->epilogue == VL_EPI_ROWLSE), "vl_gemm_bf16: null pointer");
  VL_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "vl_gemm_bf16: non-positive dims M=%d N=%d K=%d", a->M, a->N, a->K);
  const bool loss_epi = a->epilogue == VL_EPI_ROWLSE || a->epilogue == VL_EPI_CLIPGRAD;
  VL_CHECK_ARG(a->N % 8 == 0 || loss_epi, "vl_gemm_bf16: N=%d must be a multiple of 8", a->N);
  VL_CHECK_ARG(a->lda % 8 == 0 && a->ldb % 8 == 0, "vl_gemm_bf16: lda/ldb must be multiples of 8 elements");
  VL_CHECK_ARG(a->ldd % (a->d_f32 ? 4 : 8) == 0, "vl_gemm_bf16: ldd misaligned");
  VL_CHECK_ARG(a->lda >= (a->a_mn ? a->M : a->K) && a->ldb >= (a->b_mn ? a->N : a->K) &&
                   (a->epilogue == VL_EPI_ROWLSE || a->ldd >= (((a->epilogue == VL_EPI_GEGLU ? a->N / 2 : a->N) + 7) / 8) * 8),
               "vl_gemm_bf16: leading dimension smaller than the (8-padded) row length");
  VL_CHECK_ARG(!(a->accumulate && !a->d_f32), "vl_gemm_bf16: accumulate needs fp32 output");
  VL_CHECK_ARG(!(a->split_k > 1 && !(a->accumulate && a->d_f32)), "vl_gemm_bf16: split_k needs accumulate fp32 output");
  VL_CHECK_ARG(!(a->split_k > 1 && a->epilogue != VL_EPI_LINEAR), "vl_gemm_bf16: split_k only with the linear epilogue");
  if (a->epilogue == VL_EPI_RESIDUAL || a->epilogue == VL_EPI_GELU_BWD)
    VL_CHECK_ARG(a->aux_in != nullptr && a->ldaux % 8 == 0 && a->ldaux >= a->N, "vl_gemm_bf16: aux_in / ldaux invalid");
  if (a->epilogue == VL_EPI_GELU && a->aux_out)
    VL_CHECK_ARG(a->ldaux % 8 == 0 && a->ldaux >= a->N, "vl_gemm_bf16: ldaux invalid");
  if (a->epilogue < 0 || a->epilogue > VL_EPI_CLIPGRAD) {
    set_error("vl_gemm_bf16: epilogue %d not supported", a->epilogue);
    return VL_ENOTSUP;
  }
  if (a->epilogue == VL_EPI_GEGLU) {
    VL_CHECK_ARG(a->aux_out != nullptr && a->ldaux % 8 == 0 && a->ldaux >= a->N && !a->d_f32 && !a->accumulate && a->split_k <= 1 && a->ldd >= a->N / 2,
                 "vl_gemm_bf16: GEGLU needs a bf16 output [M, N/2], aux_out [M, N] (ldaux >= N) and no split_k");
    if (!(a->M >= 4 * kBM && a->N % 256 == 0 && a->b_peers == nullptr && (reinterpret_cast<uintptr_t>(a->d) & 15) == 0 && debug_get(8) != 1 &&
          debug_get(10) != 1)) {
      set_error("vl_gemm_bf16: the fused GEGLU epilogue runs on the CTA-pair kernel: M >= 512 and N %% 256 == 0 (use vl_geglu_fwd otherwise)");
      return VL_ENOTSUP;
    }
  }
  VL_CHECK_ARG(!((a->relu || a->aux_row_div > 1) && (a->d_f32 || (a->epilogue != VL_EPI_LINEAR && a->epilogue != VL_EPI_RESIDUAL))),
               "vl_gemm_bf16: relu / aux_row_div need a bf16 LINEAR or RESIDUAL epilogue");
  if (a->epilogue == VL_EPI_ROWLSE)
    VL_CHECK_ARG(a->out_vec0 && a->out_vec1 && a->out_vec2 && a->split_k <= 1, "vl_gemm_bf16: ROWLSE needs out_vec0/1/2 and split_k == 1");
  if (a->epilogue == VL_EPI_CLIPGRAD)
    VL_CHECK_ARG(a->row_vec && !a->d_f32 && a->split_k <= 1, "vl_gemm_bf16: CLIPGRAD needs row_vec, bf16 output, split_k == 1");
  if (a->rowsum_out != nullptr && !(a->M >= 4 * kBM && a->N > 128 && a->epilogue == VL_EPI_LINEAR && a->d_f32 && debug_get(8) != 1)) {
    set_error("vl_gemm_bf16: rowsum_out needs the CTA-pair kernel (M >= 512, N > 128), the LINEAR epilogue and fp32 output");
    return VL_ENOTSUP;
  }
  if (a->b_peers != nullptr) {
    VL_CHECK_ARG(a->b_npeers >= 1 && a->b_npeers <= 8 && a->b_peer_rows > 0, "vl_gemm_bf16: b_npeers / b_peer_rows invalid");
    VL_CHECK_ARG(a->split_k <= 1 && a->rowsum_out == nullptr, "vl_gemm_bf16: b_peers excludes split_k and rowsum_out");
    const long long rows = static_cast<long long>(a->b_npeers) * a->b_peer_rows;
    VL_CHECK_ARG(rows == (a->b_mn ? a->K : a->N), "vl_gemm_bf16: b_npeers * b_peer_rows must equal B's row count (%lld)", rows);
    if (a->b_npeers > 1 && a->b_peer_rows % (a->b_mn ? 64 : 256) != 0) {
      set_error("vl_gemm_bf16: b_peer_rows=%d must be a multiple of %d (a tile must not straddle two peers); gather into one buffer instead",
                a->b_peer_rows, a->b_mn ? 64 : 256);
      return VL_ENOTSUP;
    }
    for (int q = 0; q < a->b_npeers; ++q) VL_CHECK_ARG(a->b_peers[q] != nullptr, "vl_gemm_bf16: null peer pointer");
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a->b_peers != nullptr) return launch_gemm<256>(*a, s);
  if (a->epilogue == VL_EPI_ROWLSE || a->epilogue == VL_EPI_CLIPGRAD) return launch_gemm<256>(*a, s);  // fixed part geometry
  if (a->N <= 128) return launch_gemm<128>(*a, s);
  return launch_gemm<256>(*a, s);
}

extern "C" int vl_gemm_rowlse_parts(int32_t N) { return ((N + 255) / 256) * 4; }  // one part per 64-column slice
