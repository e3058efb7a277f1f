// HBM-bound row kernels of the ViT-Lens hot path: LayerNorm fwd/bwd, bias-gradient column sums,
// patch gather (conv-as-GEMM A operand), token assembly (cls + positional), embedding lookup,
// L2 normalisation, GEGLU, casts/transposes and fused AdamW.  All are single-pass, 16-byte
// vectorised, one warp per row where a row reduction is needed; grids are sized in multiples of
// the SM count and grid-stride over rows.
#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {

constexpr int kMaxLnChunks = 8;  // D <= 32 lanes * 8 chunks * 8 elems = 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------ LayerNorm
// y[i,:] = (x[r,:] - mean) * rstd * w + b with r = row_index ? row_index[i] : i.
// Warp per row; lane l owns the 16-byte vectors l, l + 32, ... of every row, so its slice of w / b is loop-invariant and lives
// in registers (re-reading it per row made the kernel L1-bound: 8 KB of weight traffic per 2 KB row, l1tex 80 % busy in ncu).
// The row itself stays packed (bf16) in registers and is unpacked in each of the three passes.  The next kLnPF rows of the warp
// are in flight as cp.async copies into a private shared-memory ring (each lane copies and later reads only its own 16-byte
// chunks: no barrier at all): with 16 warps per SM and one register-held row ahead the kernel had ~32 KB in flight per SM and
// sat at 0.6 of the HBM peak (latency-bound by Little's law: 6.5 TB/s x ~1 us needs ~44 KB per SM); the ring holds 4x that.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int CH, int kLnPF>
__global__ void __launch_bounds__(256, 2) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                               const long long* __restrict__ row_index,
                                                               const float* __restrict__ w, const float* __restrict__ b,
                                                               __nv_bfloat16* __restrict__ y, long long ldy,
                                                               float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                               int T, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int nvec = D >> 3;
  const long long stride = (long long)gridDim.x * wpb;
  long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5);
  constexpr bool kRegW = CH <= 4;  // D <= 1024: 64 registers of weights; wider rows re-read them per row (L1 hits)
  constexpr int WCH = kRegW ? CH : 1;
  float wr[WCH][8], br[WCH][8];
#pragma unroll
  for (int c = 0; c < WCH; ++c) {
    const int i = lane + c * 32;
    if (kRegW && i < nvec) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * i), w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * i + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + 2 * i), b1 = __ldg(reinterpret_cast<const float4*>(b) + 2 * i + 1);
      wr[c][0] = w0.x; wr[c][1] = w0.y; wr[c][2] = w0.z; wr[c][3] = w0.w; wr[c][4] = w1.x; wr[c][5] = w1.y; wr[c][6] = w1.z; wr[c][7] = w1.w;
      br[c][0] = b0.x; br[c][1] = b0.y; br[c][2] = b0.z; br[c][3] = b0.w; br[c][4] = b1.x; br[c][5] = b1.y; br[c][6] = b1.z; br[c][7] = b1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) wr[c][j] = br[c][j] = 0.f;
    }
  }
  extern __shared__ uint4 ln_ring[];  // [warps per CTA][kLnPF][CH * 32] 16-byte chunks
  uint4* my = ln_ring + static_cast<size_t>(threadIdx.x >> 5) * kLnPF * CH * 32 + lane;
  auto fetch = [&](long long r, int slot) {  // (always commits a group, possibly empty: the wait below counts groups)
    if (r < T) {
      const long long src = row_index ? row_index[r] : r;
      const uint4* xr = reinterpret_cast<const uint4*>(x + src * ldx);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int i = lane + c * 32;
        if (i < nvec) cp_async16(smem_u32(my + (slot * CH + c) * 32), xr + i);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int k = 0; k < kLnPF; ++k) fetch(row + k * stride, k);
  const float invD = 1.0f / D;
  int slot = 0;
  for (; row < T; row += stride) {
    cp_async_wait<kLnPF - 1>();  // this row's copies (the oldest group) have landed
    uint4 cur[CH];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      cur[c] = (lane + c * 32 < nvec) ? my[(slot * CH + c) * 32] : make_uint4(0u, 0u, 0u, 0u);
      float v[8];
      unpack8(cur[c], v);  // vectors past the row are zero
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
    }
    fetch(row + kLnPF * stride, slot);  // refill the slot just read (the sum above depends on every chunk of it)
    slot = slot + 1 == kLnPF ? 0 : slot + 1;
    const float mu = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = lane + c * 32;
      if (i < nvec) {
        float v[8];
        unpack8(cur[c], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[j] - mu;
          q = fmaf(d, d, q);
        }
      }
    }
    const float rs = rsqrtf(warp_sum(q) * invD + eps);
    const float nmr = -mu * rs;
    uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = lane + c * 32;
      if (i < nvec) {
        float v[8], o[8];
        unpack8(cur[c], v);
        if constexpr (kRegW) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(fmaf(v[j], rs, nmr), wr[c][j], br[c][j]);
        } else {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * i), w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * i + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + 2 * i), b1 = __ldg(reinterpret_cast<const float4*>(b) + 2 * i + 1);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(fmaf(v[j], rs, nmr), wv[j], bv[j]);
        }
        yr[i] = pack8(o);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rs;
    }
  }
}

// dx = rstd * (w*dy - mean(w*dy) - xhat * mean(w*dy*xhat)) [+ dres];  dw += sum_t dy*xhat; db += sum_t dy;
// dres_sum += sum_t dres (template RS).  Rows may be scattered back through row_index (dx row r = row_index[i]).
// Two passes over the row: pass 1 reduces the two row statistics, pass 2 re-reads x / dy (L1 hits: a row is 2-4 KB) and
// writes dx.  Not holding the row in registers keeps the kernel at <= 128 registers -> 16 warps / SM, which is what a
// streaming kernel needs to cover HBM latency.
template <int CH, bool RS>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                               const __nv_bfloat16* __restrict__ x, long long ldx,
                                                               const long long* __restrict__ row_index,
                                                               const float* __restrict__ w, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               const __nv_bfloat16* __restrict__ dres, long long lddres,
                                                               __nv_bfloat16* __restrict__ dx, long long lddx,
                                                               float* __restrict__ part, int want_w, int want_r, int T, int D) {
  // part: [gridDim.x][(2 * want_w + want_r) * D] per-CTA partial sums of dw | db | dres_sum (second stage: launch_colreduce)
  extern __shared__ float red[];  // [warps][D] reused for dw, db, dres_sum
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int nvec = D >> 3;
  float aw[CH][8], ab[CH][8], ar[RS ? CH : 1][8];
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      aw[c][j] = ab[c][j] = 0.f;
      if (RS) ar[c][j] = 0.f;
    }

  for (long long row = (long long)blockIdx.x * wpb + warp; row < T; row += (long long)gridDim.x * wpb) {
    const long long src = row_index ? row_index[row] : row;
    const uint4* xr = reinterpret_cast<const uint4*>(x + src * ldx);
    const uint4* gr = reinterpret_cast<const uint4*>(dy + row * lddy);
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = lane + c * 32;
      if (i < nvec) {
        float xv[8], gv[8];
        unpack8(xr[i], xv);
        unpack8(gr[i], gv);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * i), w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * i + 1);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float wg = gv[j] * wv[j];
          s1 += wg;
          s2 += wg * ((xv[j] - mu) * rs);
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    uint4* dr = reinterpret_cast<uint4*>(dx + src * lddx);
    const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + src * lddres) : nullptr;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = lane + c * 32;
      if (i < nvec) {
        float xv[8], gv[8], o[8];
        unpack8(xr[i], xv);
        unpack8(gr[i], gv);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + 2 * i), w1 = __ldg(reinterpret_cast<const float4*>(w) + 2 * i + 1);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mu) * rs;
          if (want_w) {
            aw[c][j] += gv[j] * xh;
            ab[c][j] += gv[j];
          }
          o[j] = rs * (gv[j] * wv[j] - s1 - xh * s2);
        }
        if (rr) {
          float r[8];
          unpack8(rr[i], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] += r[j];
            if (RS) ar[c][j] += r[j];
          }
        }
        dr[i] = pack8(o);
      }
    }
  }
  if (part == nullptr) return;
  // block reduction of the per-warp partial column sums, then one partial row per CTA (plain stores)
  const int nout = 2 * want_w + want_r;
  for (int pass = 0; pass < (RS ? 3 : 2); ++pass) {
    if (pass < 2 ? !want_w : !want_r) continue;
    float* dst = part + (long long)blockIdx.x * nout * D + (pass < 2 ? pass : 2 * want_w) * D;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = lane + c * 32;
      if (i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp * D + i * 8 + j] = pass == 0 ? aw[c][j] : (pass == 1 ? ab[c][j] : ar[RS ? c : 0][j]);
      }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < D; col += blockDim.x) {
      float s = 0.f;
      for (int ww = 0; ww < wpb; ++ww) s += red[ww * D + col];
      dst[col] = s;
    }
  }
}

// D = 1024 fast path (every ViT-L LayerNorm): the CTA owns R rows at a time and thread t the four columns [4t, 4t + 4) of each,
// so the dgamma / dbeta accumulators are 8 registers per thread (the warp-per-row kernel above needs 64) and each thread
// keeps 2 x R independent 8-byte loads in flight.  Row statistics: per-thread partials -> warp shuffles -> one smem hop.
template <bool HAS_RES>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_d1024_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                                     const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                     const float* __restrict__ w, const float* __restrict__ mean,
                                                                     const float* __restrict__ rstd,
                                                                     const __nv_bfloat16* __restrict__ dres, long long lddres,
                                                                     __nv_bfloat16* __restrict__ dx, long long lddx,
                                                                     float* __restrict__ part, int T) {
  // part: [gridDim.x][2048] per-CTA partial dw | db, or null
  // rows per step.  Measured on B200 (65792 x 1024): R = 8 without register prefetch 136.6 us, R = 4 with the next step
  // prefetched 147.6 us (twice the barriers), the warp-per-row kernel 144.8 us.
  constexpr int R = 8;
  __shared__ float red[8][2 * R];
  __shared__ float tot[2 * R];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int c0 = 4 * t;
  const float4 w4 = *reinterpret_cast<const float4*>(w + c0);
  const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
  float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  uint2 nx[R], ng[R], nr[R];
  float nmu[R], nrs[R];
  auto fetch = [&](long long r0) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const long long row = r0 + i;
      const bool ok = row < T;
      nx[i] = ok ? *reinterpret_cast<const uint2*>(x + row * ldx + c0) : make_uint2(0u, 0u);
      ng[i] = ok ? *reinterpret_cast<const uint2*>(dy + row * lddy + c0) : make_uint2(0u, 0u);
      if (HAS_RES) nr[i] = ok ? *reinterpret_cast<const uint2*>(dres + row * lddres + c0) : make_uint2(0u, 0u);
      nmu[i] = ok ? mean[row] : 0.f;
      nrs[i] = ok ? rstd[row] : 0.f;
    }
  };
  const long long step = (long long)gridDim.x * R;
  long long r0 = (long long)blockIdx.x * R;
  for (; r0 < T; r0 += step) {
    fetch(r0);
    uint2 xv[R], gv[R], rv[R];
    float mu[R], rs[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      xv[i] = nx[i]; gv[i] = ng[i]; mu[i] = nmu[i]; rs[i] = nrs[i];
      if (HAS_RES) rv[i] = nr[i];
    }
    float s[2 * R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float xf[4] = {bf16_lo(xv[i].x), bf16_hi(xv[i].x), bf16_lo(xv[i].y), bf16_hi(xv[i].y)};
      const float gf[4] = {bf16_lo(gv[i].x), bf16_hi(gv[i].x), bf16_lo(gv[i].y), bf16_hi(gv[i].y)};
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float wg = gf[j] * wv[j];
        s1 += wg;
        s2 = fmaf(wg, (xf[j] - mu[i]) * rs[i], s2);
      }
      s[2 * i] = s1;
      s[2 * i + 1] = s2;
    }
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) s[k] = warp_sum(s[k]);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 2 * R; ++k) red[warp][k] = s[k];
    }
    __syncthreads();
    if (t < 2 * R) {
      float a = 0.f;
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) a += red[ww][t];
      tot[t] = a * (1.0f / 1024.0f);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const long long row = r0 + i;
      if (row >= T) break;
      const float s1 = tot[2 * i], s2 = tot[2 * i + 1];
      const float xf[4] = {bf16_lo(xv[i].x), bf16_hi(xv[i].x), bf16_lo(xv[i].y), bf16_hi(xv[i].y)};
      const float gf[4] = {bf16_lo(gv[i].x), bf16_hi(gv[i].x), bf16_lo(gv[i].y), bf16_hi(gv[i].y)};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (xf[j] - mu[i]) * rs[i];
        aw[j] = fmaf(gf[j], xh, aw[j]);
        ab[j] += gf[j];
        o[j] = rs[i] * (gf[j] * wv[j] - s1 - xh * s2);
      }
      if (HAS_RES) {
        o[0] += bf16_lo(rv[i].x); o[1] += bf16_hi(rv[i].x); o[2] += bf16_lo(rv[i].y); o[3] += bf16_hi(rv[i].y);
      }
      *reinterpret_cast<uint2*>(dx + row * lddx + c0) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    }
  }
  if (part != nullptr) {
    float* pw = part + (long long)blockIdx.x * 2048;
    *reinterpret_cast<float4*>(pw + c0) = make_float4(aw[0], aw[1], aw[2], aw[3]);
    *reinterpret_cast<float4*>(pw + 1024 + c0) = make_float4(ab[0], ab[1], ab[2], ab[3]);
  }
}

// D = 1024, second generation.  ncu on the kernel above: 1332 warp-instructions per row, issue slots 60 % busy at 16 warps / SM,
// i.e. instruction-bound before it is HBM-bound.  Here a thread owns EIGHT columns (16-byte loads; 128 threads per row, two row
// slots per CTA, R rows per slot and step), the fp32 math runs two elements per instruction (FFMA2 / FMUL2 / FADD2), the
// normalised x and the gradient stay unpacked in registers between the statistics pass and the output pass, and the 2 R row
// statistics are reduced across the warp with a transposing butterfly (4 + 2 + 1 + 2 shuffles instead of 8 x 5).
template <bool HAS_RES>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_d1024_v2_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                                        const __nv_bfloat16* __restrict__ x, long long ldx,
                                                                        const float* __restrict__ w, const float* __restrict__ mean,
                                                                        const float* __restrict__ rstd,
                                                                        const __nv_bfloat16* __restrict__ dres, long long lddres,
                                                                        __nv_bfloat16* __restrict__ dx, long long lddx,
                                                                        float* __restrict__ part, int T) {
  constexpr int R = 4;  // rows per slot and step (8 rows per CTA step)
  __shared__ float red[2][4][2 * R];
  __shared__ float tot[2][2 * R];
  __shared__ float comb[128][16];
  const int t = threadIdx.x, lane = t & 31;
  const int slot = t >> 7, tt = t & 127, wslot = (t >> 5) & 3;
  const int c0 = 8 * tt;
  float2 wv[4];
  {
    const float4 a = *reinterpret_cast<const float4*>(w + c0), b = *reinterpret_cast<const float4*>(w + c0 + 4);
    wv[0] = make_float2(a.x, a.y); wv[1] = make_float2(a.z, a.w); wv[2] = make_float2(b.x, b.y); wv[3] = make_float2(b.z, b.w);
  }
  float2 aw[4], ab[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) aw[j] = ab[j] = make_float2(0.f, 0.f);
  const long long step = (long long)gridDim.x * (2 * R);
  // The rows of the NEXT step are in flight (cp.async into a two-stage ring; every thread copies and later reads only its own
  // 16-byte chunks, so no barrier is involved) while this step is reduced and written: without it the kernel alternated between
  // a load phase and a compute phase with nothing in flight (0.59 of the HBM peak inside the step).
  extern __shared__ uint4 lnb_ring[];  // [2 stages][x | dy | dres][R][256 threads]
  auto fetch = [&](long long r0n, int st) {  // always commits a group (possibly empty)
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const long long row = r0n + i;
      if (row < T) {
        cp_async16(smem_u32(lnb_ring + ((st * 3 + 0) * R + i) * 256 + t), x + row * ldx + c0);
        cp_async16(smem_u32(lnb_ring + ((st * 3 + 1) * R + i) * 256 + t), dy + row * lddy + c0);
        if (HAS_RES) cp_async16(smem_u32(lnb_ring + ((st * 3 + 2) * R + i) * 256 + t), dres + row * lddres + c0);
      }
    }
    cp_async_commit();
  };
  int st = 0;
  fetch((long long)blockIdx.x * (2 * R) + slot * R, 0);
  for (long long r0 = (long long)blockIdx.x * (2 * R) + slot * R; r0 - slot * R < T; r0 += step, st ^= 1) {
    fetch(r0 + step, st ^ 1);
    uint4 xr[R], gr[R], rr[R];
    float mu[R], rs[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const long long row = r0 + i;
      const bool ok = row < T;
      mu[i] = ok ? mean[row] : 0.f;
      rs[i] = ok ? rstd[row] : 0.f;
    }
    cp_async_wait<1>();  // this step's copies have landed (the next step's may still be in flight)
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const bool ok = r0 + i < T;
      xr[i] = ok ? lnb_ring[((st * 3 + 0) * R + i) * 256 + t] : make_uint4(0u, 0u, 0u, 0u);
      gr[i] = ok ? lnb_ring[((st * 3 + 1) * R + i) * 256 + t] : make_uint4(0u, 0u, 0u, 0u);
      if (HAS_RES) rr[i] = ok ? lnb_ring[((st * 3 + 2) * R + i) * 256 + t] : make_uint4(0u, 0u, 0u, 0u);
    }
    float2 xh[R][4], g[R][4];
    float s[2 * R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const uint32_t xw[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w}, gw[4] = {gr[i].x, gr[i].y, gr[i].z, gr[i].w};
      const float2 a2 = make_float2(rs[i], rs[i]), b2 = make_float2(-mu[i] * rs[i], -mu[i] * rs[i]);
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xh[i][j] = __ffma2_rn(make_float2(bf16_lo(xw[j]), bf16_hi(xw[j])), a2, b2);  // (x - mean) * rstd
        g[i][j] = make_float2(bf16_lo(gw[j]), bf16_hi(gw[j]));
        const float2 wg = __fmul2_rn(g[i][j], wv[j]);
        s1 = __fadd2_rn(s1, wg);
        s2 = __ffma2_rn(wg, xh[i][j], s2);
      }
      s[2 * i] = s1.x + s1.y;
      s[2 * i + 1] = s2.x + s2.y;
    }
    // transposing butterfly: 8 values x 32 lanes -> lane l ends with the warp sum of value ((l >> 4) & 1) * 4 + ((l >> 3) & 1) * 2 + ((l >> 2) & 1)
    {
      const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
      float a[4], b2[2], c;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float keep = h16 ? s[4 + k] : s[k], send = h16 ? s[k] : s[4 + k];
        a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float keep = h8 ? a[2 + k] : a[k], send = h8 ? a[k] : a[2 + k];
        b2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      {
        const float keep = h4 ? b2[1] : b2[0], send = h4 ? b2[0] : b2[1];
        c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      c += __shfl_xor_sync(0xffffffffu, c, 2);
      c += __shfl_xor_sync(0xffffffffu, c, 1);
      if ((lane & 3) == 0) red[slot][wslot][lane >> 2] = c;
    }
    __syncthreads();
    if (tt < 2 * R) tot[slot][tt] = (red[slot][0][tt] + red[slot][1][tt] + red[slot][2][tt] + red[slot][3][tt]) * (1.0f / 1024.0f);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const long long row = r0 + i;
      if (row < T) {
        const float s1 = tot[slot][2 * i], s2 = tot[slot][2 * i + 1];
        const float2 ns1 = make_float2(-s1, -s1), ns2 = make_float2(-s2, -s2), r2 = make_float2(rs[i], rs[i]);
        const uint32_t rw[4] = {HAS_RES ? rr[i].x : 0u, HAS_RES ? rr[i].y : 0u, HAS_RES ? rr[i].z : 0u, HAS_RES ? rr[i].w : 0u};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          aw[j] = __ffma2_rn(g[i][j], xh[i][j], aw[j]);
          ab[j] = __fadd2_rn(ab[j], g[i][j]);
          float2 v = __ffma2_rn(g[i][j], wv[j], ns1);  // w dy - mean(w dy)
          v = __ffma2_rn(xh[i][j], ns2, v);            // - xhat mean(w dy xhat)
          if (HAS_RES) v = __ffma2_rn(v, r2, make_float2(bf16_lo(rw[j]), bf16_hi(rw[j])));
          else v = __fmul2_rn(v, r2);
          o[j] = pack_bf16(v.x, v.y);
        }
        *reinterpret_cast<uint4*>(dx + row * lddx + c0) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  if (part != nullptr) {  // the two row slots own the same columns: combine through shared memory, one partial row per CTA
    if (slot == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        comb[tt][4 * j] = aw[j].x; comb[tt][4 * j + 1] = aw[j].y; comb[tt][4 * j + 2] = ab[j].x; comb[tt][4 * j + 3] = ab[j].y;
      }
    }
    __syncthreads();
    if (slot == 0) {
      float* pw = part + (long long)blockIdx.x * 2048 + c0;
      float o[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[2 * j] = aw[j].x + comb[tt][4 * j];
        o[2 * j + 1] = aw[j].y + comb[tt][4 * j + 1];
        o[8 + 2 * j] = ab[j].x + comb[tt][4 * j + 2];
        o[8 + 2 * j + 1] = ab[j].y + comb[tt][4 * j + 3];
      }
      *reinterpret_cast<float4*>(pw) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(pw + 4) = make_float4(o[4], o[5], o[6], o[7]);
      *reinterpret_cast<float4*>(pw + 1024) = make_float4(o[8], o[9], o[10], o[11]);
      *reinterpret_cast<float4*>(pw + 1028) = make_float4(o[12], o[13], o[14], o[15]);
    }
  }
}

// ------------------------------------------------------------------------------------ column sums (bias grads)
// part[blockIdx.y][n] = sum over this CTA's rows of dy[t, n]   (second stage: launch_colreduce)
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ dy, long long ld, float* __restrict__ part,
                                                     int T, int N, int rows_per_cta) {
  __shared__ float red[8][256];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min((long long)T, r0 + rows_per_cta);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    for (long long r = r0 + ty; r < r1; r += 8) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(dy + r * ld + col), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;  // 256 columns per CTA
  float s = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) s += red[y][c];
  const int gc = blockIdx.x * 256 + c;
  if (gc < N) part[(long long)blockIdx.y * N + gc] = s;
}

// ------------------------------------------------------------------------------------ patch gather (im2col)
// out[(b*OH + oh)*OW + ow, (c*kh + i)*kw + j] = in[b*sb + c*sc + (oh*sth+i)*sh + (ow*stw+j)*sw]; columns >= C*kh*kw are 0.
template <typename TIn>
__global__ void __launch_bounds__(256) patchify_kernel(const TIn* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int C,
                                                       int OH, int OW, int kh, int kw, int sth, int stw, long long sb,
                                                       long long sc, long long sh, long long sw, int Kpad) {
  const int kvec = Kpad >> 3;
  const long long total = (long long)B * OH * OW * kvec;
  const int K = C * kh * kw;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int kv = (int)(idx % kvec);
    const long long row = idx / kvec;
    const int ow = (int)(row % OW);
    const int oh = (int)((row / OW) % OH);
    const int b = (int)(row / ((long long)OW * OH));
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kv * 8 + e;
      float val = 0.f;
      if (k < K) {
        const int j = k % kw;
        const int i = (k / kw) % kh;
        const int c = k / (kw * kh);
        val = static_cast<float>(in[b * sb + c * sc + (long long)(oh * sth + i) * sh + (long long)(ow * stw + j) * sw]);
      }
      f[e] = val;
    }
    *reinterpret_cast<uint4*>(out + row * Kpad + kv * 8) = pack8(f);
  }
}

// ------------------------------------------------------------------------------------ token assembly
// out[b, off + l, :] = tok[b, l, :] + pos[off + l, :]  (off = has_cls);  out[b, 0, :] = cls + pos[0] when has_cls.
// pos may be NULL (treated as 0).
__global__ void __launch_bounds__(256) assemble_kernel(const __nv_bfloat16* __restrict__ tok, const float* __restrict__ cls,
                                                       const float* __restrict__ pos, __nv_bfloat16* __restrict__ out, int B, int L,
                                                       int D, int has_cls) {
  const int dvec = D >> 3;
  const int Lo = L + has_cls;
  const long long total = (long long)B * Lo * dvec;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int dv = (int)(idx % dvec);
    const long long r = idx / dvec;
    const int lo = (int)(r % Lo);
    const long long b = r / Lo;
    float f[8];
    if (has_cls && lo == 0) {
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(cls) + 2 * dv), c1 = __ldg(reinterpret_cast<const float4*>(cls) + 2 * dv + 1);
      f[0] = c0.x; f[1] = c0.y; f[2] = c0.z; f[3] = c0.w; f[4] = c1.x; f[5] = c1.y; f[6] = c1.z; f[7] = c1.w;
    } else {
      unpack8(*reinterpret_cast<const uint4*>(tok + (b * L + (lo - has_cls)) * D + dv * 8), f);
    }
    if (pos) {
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + (long long)lo * D) + 2 * dv);
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + (long long)lo * D) + 2 * dv + 1);
      f[0] += p0.x; f[1] += p0.y; f[2] += p0.z; f[3] += p0.w; f[4] += p1.x; f[5] += p1.y; f[6] += p1.z; f[7] += p1.w;
    }
    *reinterpret_cast<uint4*>(out + r * D + dv * 8) = pack8(f);
  }
}

// Backward of assemble: dtok[b,l,:] = dx[b, off+l, :] (optional);  part[blockIdx.y][lo,:] = sum over this CTA's batch chunk of
// dx[b,lo,:] for lo < part_rows (optional; dpos = all rows, dcls = row 0; second stage: launch_colreduce).  One thread owns
// (lo, 8 columns) and loops over the batch chunk.
__global__ void __launch_bounds__(256) assemble_bwd_kernel(const __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dtok,
                                                           float* __restrict__ part, int part_rows, int B, int L, int D,
                                                           int has_cls, int bchunk) {
  const int dvec = D >> 3;
  const int Lo = L + has_cls;
  const long long total = (long long)Lo * dvec;
  const int b0 = blockIdx.y * bchunk, b1 = min(B, b0 + bchunk);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int dv = (int)(idx % dvec);
    const int lo = (int)(idx / dvec);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = b0; b < b1; ++b) {
      const uint4 u = *reinterpret_cast<const uint4*>(dx + ((long long)b * Lo + lo) * D + dv * 8);
      if (dtok && lo >= has_cls) *reinterpret_cast<uint4*>(dtok + ((long long)b * L + lo - has_cls) * D + dv * 8) = u;
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    if (part != nullptr && lo < part_rows) {
      float* pp = part + ((long long)blockIdx.y * part_rows + lo) * D + dv * 8;
      *reinterpret_cast<float4*>(pp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(pp + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// ------------------------------------------------------------------------------------ text embedding
// out[b*ctx + t, :] = table[ids[b,t], :] + pos[t, :]
__global__ void __launch_bounds__(256) embed_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                                    const float* __restrict__ pos, __nv_bfloat16* __restrict__ out, long long rows,
                                                    int ctx, int D) {
  const int dvec = D >> 3;
  const long long total = rows * dvec;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int dv = (int)(idx % dvec);
    const long long r = idx / dvec;
    const long long id = ids[r];
    const int t = (int)(r % ctx);
    const float4* tr = reinterpret_cast<const float4*>(table + id * D) + 2 * dv;
    const float4* pr = reinterpret_cast<const float4*>(pos + (long long)t * D) + 2 * dv;
    const float4 a0 = __ldg(tr), a1 = __ldg(tr + 1), p0 = __ldg(pr), p1 = __ldg(pr + 1);
    const float f[8] = {a0.x + p0.x, a0.y + p0.y, a0.z + p0.z, a0.w + p0.w, a1.x + p1.x, a1.y + p1.y, a1.z + p1.z, a1.w + p1.w};
    *reinterpret_cast<uint4*>(out + r * D + dv * 8) = pack8(f);
  }
}
// dtable[ids[r], :] += dx[r, :];  dpos[t, :] += dx[r, :]
__global__ void __launch_bounds__(256) embed_bwd_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ dx,
                                                        float* __restrict__ dtable, float* __restrict__ dpos, long long rows, int ctx, int D) {
  const int dvec = D >> 3;
  const long long total = rows * dvec;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int dv = (int)(idx % dvec);
    const long long r = idx / dvec;
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(dx + r * D + dv * 8), f);
    const long long id = ids[r];
    const int t = (int)(r % ctx);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (dtable) atomicAdd(dtable + id * D + dv * 8 + j, f[j]);
      if (dpos) atomicAdd(dpos + (long long)t * D + dv * 8 + j, f[j]);
    }
  }
}

// ------------------------------------------------------------------------------------ L2 normalise (fp32 rows)
// y = x / max(||x||, eps); inv_norm saved.   bwd: dx = (dy - y * <dy, y>) * inv_norm
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ inv_norm,
                                                         int B, int E, float eps) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < B; row += gridDim.x * wpb) {
    float s = 0.f;
    for (int i = lane; i < E; i += 32) {
      const float v = x[(long long)row * E + i];
      s += v * v;
    }
    const float inv = 1.0f / fmaxf(sqrtf(warp_sum(s)), eps);
    for (int i = lane; i < E; i += 32) y[(long long)row * E + i] = x[(long long)row * E + i] * inv;
    if (lane == 0 && inv_norm) inv_norm[row] = inv;
  }
}
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                         const float* __restrict__ inv_norm, float* __restrict__ dx, int B, int E) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < B; row += gridDim.x * wpb) {
    float s = 0.f;
    for (int i = lane; i < E; i += 32) s += dy[(long long)row * E + i] * y[(long long)row * E + i];
    s = warp_sum(s);
    const float inv = inv_norm[row];
    for (int i = lane; i < E; i += 32)
      dx[(long long)row * E + i] = (dy[(long long)row * E + i] - y[(long long)row * E + i] * s) * inv;
  }
}

// ------------------------------------------------------------------------------------ GEGLU (Lens FeedForward)
// h[M, 2F] = [val | gate];  out[M, F] = val * gelu(gate)
__global__ void __launch_bounds__(256) geglu_fwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ out, long long M, int F) {
  const int fvec = F >> 3;
  const long long total = M * fvec;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int fv = (int)(idx % fvec);
    const long long r = idx / fvec;
    float a[8], g[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(h + r * 2 * F + fv * 8), a);
    unpack8(*reinterpret_cast<const uint4*>(h + r * 2 * F + F + fv * 8), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] * gelu_fwd(g[j], 0);
    *reinterpret_cast<uint4*>(out + r * F + fv * 8) = pack8(o);
  }
}
// dh[M, 2F]: dval = dout * gelu(gate); dgate = dout * val * gelu'(gate)
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ dout,
                                                        __nv_bfloat16* __restrict__ dh, long long M, int F) {
  const int fvec = F >> 3;
  const long long total = M * fvec;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int fv = (int)(idx % fvec);
    const long long r = idx / fvec;
    float a[8], g[8], d[8], da[8], dg[8];
    unpack8(*reinterpret_cast<const uint4*>(h + r * 2 * F + fv * 8), a);
    unpack8(*reinterpret_cast<const uint4*>(h + r * 2 * F + F + fv * 8), g);
    unpack8(*reinterpret_cast<const uint4*>(dout + r * F + fv * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      da[j] = d[j] * gelu_fwd(g[j], 0);
      dg[j] = d[j] * a[j] * gelu_grad(g[j], 0);
    }
    *reinterpret_cast<uint4*>(dh + r * 2 * F + fv * 8) = pack8(da);
    *reinterpret_cast<uint4*>(dh + r * 2 * F + F + fv * 8) = pack8(dg);
  }
}

// ------------------------------------------------------------------------------------ casts / adds
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long nv = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * i], b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    reinterpret_cast<uint4*>(out)[i] = pack8(f);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = nv << 3; i < n; ++i) out[i] = __float2bfloat16(in[i]);
}
// out = a + b  (bf16, n % 8 == 0)
__global__ void __launch_bounds__(256) add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                       __nv_bfloat16* __restrict__ out, long long n) {
  const long long nv = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(reinterpret_cast<const uint4*>(a)[i], x);
    unpack8(reinterpret_cast<const uint4*>(b)[i], y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    reinterpret_cast<uint4*>(out)[i] = pack8(x);
  }
}

// ------------------------------------------------------------------------------------ fused AdamW (one tensor per launch)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2, float grad_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    p[i] = pi;
  }
}

// ------------------------------------------------------------------------------------ contrastive loss: combine per-part LSEs
__global__ void __launch_bounds__(256) lse_combine_kernel(const float* __restrict__ pm, const float* __restrict__ ps,
                                                          const float* __restrict__ diag, int M, int nparts, float* __restrict__ lse,
                                                          float* __restrict__ loss_part) {
  __shared__ float red[8];
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (row < M) {
    float mx = -INFINITY;
    for (int q = 0; q < nparts; ++q) mx = fmaxf(mx, pm[(long long)row * nparts + q]);
    float sum = 0.f;
    for (int q = 0; q < nparts; ++q) {
      const float m = pm[(long long)row * nparts + q];
      if (m > -INFINITY) sum += ps[(long long)row * nparts + q] * __expf(m - mx);
    }
    const float l = mx + __logf(sum);
    lse[row] = l;
    contrib = l - diag[row];
  }
  contrib = warp_sum(contrib);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x == 0 && loss_part) {  // one partial per CTA; summed in CTA order by launch_colreduce
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    loss_part[blockIdx.x] = t;
  }
}

// Multi-tensor AdamW: one launch for every parameter.  chunk_tab[c] = (tensor id, element offset); per tensor a row of
// five pointers (p, g, m, v, p16 | NULL) and a weight-decay value.  Also refreshes the bf16 operand copy of the weight.
constexpr int kAdamChunk = 16384;
// sumsq and max_norm: gradient-norm clipping (torch.nn.utils.clip_grad_norm_, training/train.py:212-240) folded into the
// update: with *sumsq = sum over ALL gradients of g^2 (vl_multi_sqnorm), every gradient is scaled by
// min(1, max_norm / (grad_scale * sqrt(*sumsq) + 1e-6)) on its way into the moments -- no extra pass over the gradients.
__global__ void __launch_bounds__(256) adamw_multi_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ sizes,
                                                          const float* __restrict__ wds, const float* __restrict__ lrs,
                                                          const int2* __restrict__ chunk_tab, float lr, float b1,
                                                          float b2, float eps, float bc1, float bc2, float grad_scale,
                                                          const float* __restrict__ sumsq, float max_norm, int n_src, long long src_stride) {
  // n_src > 1: the gradient is the sum, in source order, of n_src copies src_stride elements apart (the ranks' slots of the
  // peer-memory gradient exchange, vl_allreduce_grads): every rank adds the same values in the same order
  if (sumsq != nullptr) grad_scale *= fminf(1.0f, max_norm / (grad_scale * sqrtf(__ldg(sumsq)) + 1e-6f));
  const int2 ct = chunk_tab[blockIdx.x];
  const long long* pr = ptrs + 5ll * ct.x;
  float* p = reinterpret_cast<float*>(pr[0]);
  const float* g = reinterpret_cast<const float*>(pr[1]);
  if (g == nullptr) return;  // no gradient this step: the parameter (and its moments) stay untouched, as in torch.optim
  if (lrs != nullptr) lr = lrs[ct.x];  // per-tensor learning rate (param_groups)
  float* m = reinterpret_cast<float*>(pr[2]);
  float* v = reinterpret_cast<float*>(pr[3]);
  __nv_bfloat16* p16 = reinterpret_cast<__nv_bfloat16*>(pr[4]);
  const float wd = wds[ct.x];
  const long long n = sizes[ct.x];
  const long long base = static_cast<long long>(ct.y) * kAdamChunk;
  const long long end = min(n, base + kAdamChunk);
  for (long long i = base + threadIdx.x; i < end; i += 256) {
    float gsum = g[i];
    for (int q = 1; q < n_src; ++q) gsum += g[i + q * src_stride];
    const float gi = gsum * grad_scale;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    p[i] = pi;
    if (p16) p16[i] = __float2bfloat16(pi);
  }
}

// out[blockIdx.x] = sum of g^2 over one 16384-element chunk of one gradient tensor (same tables as adamw_multi_kernel)
__global__ void __launch_bounds__(256) multi_sqnorm_kernel(const long long* __restrict__ ptrs, const long long* __restrict__ sizes,
                                                           const int2* __restrict__ chunk_tab, float* __restrict__ out, int n_src,
                                                           long long src_stride) {
  const int2 ct = chunk_tab[blockIdx.x];
  const float* g = reinterpret_cast<const float*>(ptrs[5ll * ct.x + 1]);
  const long long n = sizes[ct.x];
  const long long base = static_cast<long long>(ct.y) * kAdamChunk;
  const long long end = g != nullptr ? min(n, base + kAdamChunk) : base;  // a tensor without a gradient contributes 0
  float acc = 0.f;
  for (long long i = base + threadIdx.x; i < end; i += 256) {
    float gs = g[i];
    for (int q = 1; q < n_src; ++q) gs += g[i + q * src_stride];
    acc = fmaf(gs, gs, acc);
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) t += red[w2];
    out[blockIdx.x] = t;
  }
}

static inline int grid_for(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace vl

using namespace vl;

extern "C" {

int vl_layernorm_fwd(const void* x, int64_t ldx, const int64_t* row_index, const float* w, const float* b, void* y, int64_t ldy,
                     float* mean, float* rstd, int32_t T, int32_t D, float eps, void* stream) {
  VL_CHECK_ARG(x && w && b && y, "vl_layernorm_fwd: null pointer");
  VL_CHECK_ARG(T > 0 && D > 0 && D % 8 == 0 && D <= 256 * kMaxLnChunks, "vl_layernorm_fwd: D=%d must be a multiple of 8 and <= 2048", D);
  VL_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0, "vl_layernorm_fwd: ld must be a multiple of 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int grid = grid_for((long long)T * 32, 256);
  if (grid > 2 * num_sms()) grid = 2 * num_sms();  // persistent: two resident CTAs per SM, each warp walks its rows through the ring
  const int ch = (D / 8 + 31) / 32;
  // rows in flight per warp (cp.async ring): 4 (6 measured the same: 52.1 vs 50.7 us at 65792 x 1024), 3 for 4 KB rows
#define VL_LN_FWD(CH, PF)                                                                                                      \
  do {                                                                                                                         \
    constexpr int smem_ = 8 * PF * CH * 32 * 16;                                                                               \
    static bool attr_ = false;                                                                                                 \
    if (!attr_) {                                                                                                              \
      VL_CUDA(cudaFuncSetAttribute(layernorm_fwd_kernel<CH, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_));         \
      attr_ = true;                                                                                                            \
    }                                                                                                                          \
    layernorm_fwd_kernel<CH, PF><<<grid, 256, smem_, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), ldx, (const long long*)row_index, w, b, \
                                                          reinterpret_cast<__nv_bfloat16*>(y), ldy, mean, rstd, T, D, eps);    \
  } while (0)
  if (ch <= 1) VL_LN_FWD(1, 4); else if (ch <= 2) VL_LN_FWD(2, 4); else if (ch <= 4) VL_LN_FWD(4, 4); else VL_LN_FWD(8, 3);
#undef VL_LN_FWD
  return launch_check("layernorm_fwd");
}

int vl_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, const int64_t* row_index, const float* w,
                     const float* mean, const float* rstd, const void* dres, int64_t lddres, void* dx, int64_t lddx, float* dw,
                     float* db, float* dres_sum, int32_t T, int32_t D, void* stream) {
  VL_CHECK_ARG(dy && x && w && mean && rstd && dx, "vl_layernorm_bwd: null pointer");
  VL_CHECK_ARG((dw == nullptr) == (db == nullptr), "vl_layernorm_bwd: dw and db must both be given or both be NULL");
  VL_CHECK_ARG(dres_sum == nullptr || dres != nullptr, "vl_layernorm_bwd: dres_sum needs dres");
  VL_CHECK_ARG(T > 0 && D > 0 && D % 8 == 0 && D <= 256 * kMaxLnChunks, "vl_layernorm_bwd: D=%d unsupported", D);
  VL_CHECK_ARG(ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && lddres % 8 == 0, "vl_layernorm_bwd: ld must be a multiple of 8");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // Parameter gradients: every CTA writes one row of partial column sums into stream-ordered scratch memory and a second
  // launch adds the rows in CTA order -- deterministic, no floating-point atomics, outputs need no zero fill.
  float* part = nullptr;
  auto finish = [&](int nparts) -> int {
    if (part == nullptr) return 0;
    int rc = dw ? launch_colreduce(part, nparts, D, dw, db, dres_sum, s) : launch_colreduce(part, nparts, D, dres_sum, nullptr, nullptr, s);
    if (rc) return rc;
    return scratch_free(part, s);
  };
  if (D == 1024 && row_index == nullptr && dres_sum == nullptr && T >= 2048 && ldx % 4 == 0 && debug_get(13) != 1) {  // knob 13: 1 = generic kernel
    int g2 = num_sms() * 2;
    if ((long long)g2 * 8 > T) g2 = (T + 7) / 8;
    if (dw) {
      if (int rc = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)g2 * 2048 * sizeof(float), s)) return rc;
    }
    const bool al16 = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dres)) & 15) == 0;
    if (al16 && debug_get(14) != 1) {  // knob 14: 1 = first-generation D = 1024 kernel (four columns per thread)
      constexpr int kRing = 2 * 3 * 4 * 256 * 16;  // two stages of (x | dy | dres) x 4 rows x 256 threads x 16 bytes
      static bool attr_ = false;
      if (!attr_) {
        VL_CUDA(cudaFuncSetAttribute(layernorm_bwd_d1024_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRing));
        VL_CUDA(cudaFuncSetAttribute(layernorm_bwd_d1024_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRing));
        attr_ = true;
      }
      if (dres)
        layernorm_bwd_d1024_v2_kernel<true><<<g2, 256, kRing, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(x), ldx, w,
                                                               mean, rstd, reinterpret_cast<const __nv_bfloat16*>(dres), lddres,
                                                               reinterpret_cast<__nv_bfloat16*>(dx), lddx, part, T);
      else
        layernorm_bwd_d1024_v2_kernel<false><<<g2, 256, kRing, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(x), ldx,
                                                                w, mean, rstd, nullptr, 0, reinterpret_cast<__nv_bfloat16*>(dx), lddx, part, T);
      if (int rc = launch_check("layernorm_bwd_d1024_v2")) return rc;
      return finish(g2);
    }
    if (dres)
      layernorm_bwd_d1024_kernel<true><<<g2, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(x), ldx, w,
                                                          mean, rstd, reinterpret_cast<const __nv_bfloat16*>(dres), lddres,
                                                          reinterpret_cast<__nv_bfloat16*>(dx), lddx, part, T);
    else
      layernorm_bwd_d1024_kernel<false><<<g2, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(x), ldx, w,
                                                           mean, rstd, nullptr, 0, reinterpret_cast<__nv_bfloat16*>(dx), lddx, part, T);
    if (int rc = launch_check("layernorm_bwd_d1024")) return rc;
    return finish(g2);
  }
  int grid = num_sms() * 4;
  if ((long long)grid * 8 > T) grid = (T + 7) / 8;
  const size_t smem = (dw || dres_sum) ? (size_t)8 * D * sizeof(float) : 0;
  const int ch = (D / 8 + 31) / 32;
  const int want_w = dw ? 1 : 0, want_r = dres_sum ? 1 : 0;
  if (want_w || want_r) {
    if (int rc = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)grid * (2 * want_w + want_r) * D * sizeof(float), s)) return rc;
  }
#define VL_LN_BWD2(CH, RS)                                                                                               \
  do {                                                                                                                    \
    if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_bwd_kernel<CH, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    layernorm_bwd_kernel<CH, RS><<<grid, 256, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), lddy,                \
        reinterpret_cast<const __nv_bfloat16*>(x), ldx, (const long long*)row_index, w, mean, rstd,                       \
        reinterpret_cast<const __nv_bfloat16*>(dres), lddres, reinterpret_cast<__nv_bfloat16*>(dx), lddx, part, want_w, want_r, T, D); \
  } while (0)
#define VL_LN_BWD(CH)                  \
  do {                                 \
    if (dres_sum) VL_LN_BWD2(CH, true); \
    else VL_LN_BWD2(CH, false);        \
  } while (0)
  if (ch <= 1) VL_LN_BWD(1); else if (ch <= 2) VL_LN_BWD(2); else if (ch <= 4) VL_LN_BWD(4); else VL_LN_BWD(8);
#undef VL_LN_BWD
#undef VL_LN_BWD2
  if (int rc = launch_check("layernorm_bwd")) return rc;
  return finish(grid);
}

int vl_colsum_bf16(const void* dy, int64_t ld, float* db, int32_t T, int32_t N, void* stream) {
  VL_CHECK_ARG(dy && db && T > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0, "vl_colsum_bf16: bad arguments");
  const int gx = (N + 255) / 256;
  int gy = (num_sms() * 4 + gx - 1) / gx;
  if (gy > (T + 63) / 64) gy = (T + 63) / 64;
  if (gy < 1) gy = 1;
  const int rows_per = (T + gy - 1) / gy;
  gy = (T + rows_per - 1) / rows_per;  // no empty row ranges: every partial row is written
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* part = nullptr;
  if (int rc = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)gy * N * sizeof(float), s)) return rc;
  colsum_kernel<<<dim3(gx, gy), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), ld, part, T, N, rows_per);
  if (int rc = launch_check("colsum")) return rc;
  if (int rc = launch_colreduce(part, gy, N, db, nullptr, nullptr, s)) return rc;
  return scratch_free(part, s);
}

int vl_patchify(const void* in, int32_t in_is_bf16, void* out, int32_t B, int32_t C, int32_t OH, int32_t OW, int32_t kh, int32_t kw,
                int32_t stride_h, int32_t stride_w, int64_t sb, int64_t sc, int64_t sh, int64_t sw, int32_t Kpad, void* stream) {
  VL_CHECK_ARG(in && out && B > 0 && C > 0 && OH > 0 && OW > 0 && Kpad % 8 == 0 && Kpad >= C * kh * kw, "vl_patchify: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)B * OH * OW * (Kpad / 8);
  const int grid = grid_for(total, 256);
  if (in_is_bf16)
    patchify_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B, C, OH, OW, kh, kw, stride_h, stride_w, sb, sc, sh, sw, Kpad);
  else
    patchify_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(in), reinterpret_cast<__nv_bfloat16*>(out), B, C, OH, OW, kh, kw, stride_h, stride_w, sb, sc, sh, sw, Kpad);
  return launch_check("patchify");
}

int vl_assemble_tokens(const void* tok, const float* cls, const float* pos, void* out, int32_t B, int32_t L, int32_t D, int32_t has_cls,
                       void* stream) {
  VL_CHECK_ARG(tok && out && B > 0 && L > 0 && D % 8 == 0 && (!has_cls || cls), "vl_assemble_tokens: bad arguments");
  const long long total = (long long)B * (L + (has_cls ? 1 : 0)) * (D / 8);
  assemble_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(tok), cls, pos, reinterpret_cast<__nv_bfloat16*>(out), B, L, D, has_cls ? 1 : 0);
  return launch_check("assemble_tokens");
}

int vl_assemble_tokens_bwd(const void* dx, void* dtok, float* dpos, float* dcls, int32_t B, int32_t L, int32_t D, int32_t has_cls,
                           void* stream) {
  VL_CHECK_ARG(dx && B > 0 && L > 0 && D % 8 == 0, "vl_assemble_tokens_bwd: bad arguments");
  const long long total = (long long)(L + (has_cls ? 1 : 0)) * (D / 8);
  const int gx = (int)((total + 255) / 256);
  int gy = (num_sms() * 2 + gx - 1) / gx;
  if (gy > B) gy = B;
  if (gy < 1) gy = 1;
  const int bchunk = (B + gy - 1) / gy;
  gy = (B + bchunk - 1) / bchunk;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool want_cls = dcls != nullptr && has_cls;
  const int part_rows = dpos ? L + (has_cls ? 1 : 0) : (want_cls ? 1 : 0);  // dcls is row 0 of the same sums
  float* part = nullptr;
  if (part_rows > 0) {
    if (int rc = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)gy * part_rows * D * sizeof(float), s)) return rc;
  }
  assemble_bwd_kernel<<<dim3(gx, gy), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dx), reinterpret_cast<__nv_bfloat16*>(dtok), part, part_rows, B, L, D, has_cls ? 1 : 0, bchunk);
  if (int rc = launch_check("assemble_tokens_bwd")) return rc;
  if (part != nullptr) {
    if (dpos) {
      if (int rc = launch_colreduce(part, gy, (long long)part_rows * D, dpos, nullptr, nullptr, s)) return rc;
    }
    if (want_cls) {
      if (dpos) {
        VL_CUDA(cudaMemcpyAsync(dcls, dpos, (size_t)D * sizeof(float), cudaMemcpyDeviceToDevice, s));
      } else {
        if (int rc = launch_colreduce(part, gy, D, dcls, nullptr, nullptr, s)) return rc;
      }
    }
    return scratch_free(part, s);
  }
  return 0;
}

int vl_embed_tokens(const int64_t* ids, const float* table, const float* pos, void* out, int64_t rows, int32_t ctx, int32_t D, void* stream) {
  VL_CHECK_ARG(ids && table && pos && out && rows > 0 && D % 8 == 0, "vl_embed_tokens: bad arguments");
  embed_kernel<<<grid_for(rows * (D / 8), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const long long*)ids, table, pos, reinterpret_cast<__nv_bfloat16*>(out), rows, ctx, D);
  return launch_check("embed_tokens");
}

int vl_embed_tokens_bwd(const int64_t* ids, const void* dx, float* dtable, float* dpos, int64_t rows, int32_t ctx, int32_t D, void* stream) {
  VL_CHECK_ARG(ids && dx && rows > 0 && D % 8 == 0, "vl_embed_tokens_bwd: bad arguments");
  embed_bwd_kernel<<<grid_for(rows * (D / 8), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const long long*)ids, reinterpret_cast<const __nv_bfloat16*>(dx), dtable, dpos, rows, ctx, D);
  return launch_check("embed_tokens_bwd");
}

int vl_l2norm_fwd(const float* x, float* y, float* inv_norm, int32_t B, int32_t E, float eps, void* stream) {
  VL_CHECK_ARG(x && y && B > 0 && E > 0, "vl_l2norm_fwd: bad arguments");
  l2norm_fwd_kernel<<<grid_for((long long)B * 32, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, inv_norm, B, E, eps);
  return launch_check("l2norm_fwd");
}

int vl_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, int32_t B, int32_t E, void* stream) {
  VL_CHECK_ARG(dy && y && inv_norm && dx && B > 0 && E > 0, "vl_l2norm_bwd: bad arguments");
  l2norm_bwd_kernel<<<grid_for((long long)B * 32, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, y, inv_norm, dx, B, E);
  return launch_check("l2norm_bwd");
}

int vl_geglu_fwd(const void* h, void* out, int64_t M, int32_t F, void* stream) {
  VL_CHECK_ARG(h && out && M > 0 && F % 8 == 0, "vl_geglu_fwd: bad arguments");
  geglu_fwd_kernel<<<grid_for(M * (F / 8), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(h), reinterpret_cast<__nv_bfloat16*>(out), M, F);
  return launch_check("geglu_fwd");
}

int vl_geglu_bwd(const void* h, const void* dout, void* dh, int64_t M, int32_t F, void* stream) {
  VL_CHECK_ARG(h && dout && dh && M > 0 && F % 8 == 0, "vl_geglu_bwd: bad arguments");
  geglu_bwd_kernel<<<grid_for(M * (F / 8), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(h), reinterpret_cast<const __nv_bfloat16*>(dout), reinterpret_cast<__nv_bfloat16*>(dh), M, F);
  return launch_check("geglu_bwd");
}

int vl_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  VL_CHECK_ARG(in && out && n > 0, "vl_cast_f32_bf16: bad arguments");
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, reinterpret_cast<__nv_bfloat16*>(out), n);
  return launch_check("cast_f32_bf16");
}

int vl_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  VL_CHECK_ARG(a && b && out && n > 0 && n % 8 == 0, "vl_add_bf16: bad arguments");
  add_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b), reinterpret_cast<__nv_bfloat16*>(out), n);
  return launch_check("add_bf16");
}

int vl_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int32_t step, float grad_scale, void* stream) {
  VL_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "vl_adamw_step: bad arguments");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_kernel<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
  return launch_check("adamw_step");
}

int vl_lse_combine(const float* part_max, const float* part_sum, const float* diag, int32_t M, int32_t nparts, float* lse,
                   float* loss_sum, void* stream) {
  VL_CHECK_ARG(part_max && part_sum && diag && lse && M > 0 && nparts > 0, "vl_lse_combine: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int g = (M + 255) / 256;
  float* part = nullptr;
  if (loss_sum) {
    if (int rc = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)g * sizeof(float), s)) return rc;
  }
  lse_combine_kernel<<<g, 256, 0, s>>>(part_max, part_sum, diag, M, nparts, lse, part);
  if (int rc = launch_check("lse_combine")) return rc;
  if (part == nullptr) return 0;
  if (int rc = launch_colreduce(part, g, 1, loss_sum, nullptr, nullptr, s)) return rc;
  return scratch_free(part, s);
}

int vl_adamw_multi(const int64_t* ptrs, const int64_t* sizes, const float* wds, const float* lrs, const int32_t* chunk_tab, int32_t n_chunks,
                   float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale, int32_t n_src, int64_t src_stride, void* stream) {
  VL_CHECK_ARG(ptrs && sizes && wds && chunk_tab && n_chunks > 0 && step >= 1 && n_src >= 1, "vl_adamw_multi: bad arguments");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_multi_kernel<<<n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const long long*)ptrs, (const long long*)sizes, wds,
                                                                                     lrs, reinterpret_cast<const int2*>(chunk_tab), lr, beta1, beta2,
                                                                                     eps, bc1, bc2, grad_scale, nullptr, 0.f, n_src, src_stride);
  return launch_check("adamw_multi");
}

int vl_multi_sqnorm(const int64_t* ptrs, const int64_t* sizes, const int32_t* chunk_tab, int32_t n_chunks, float* sumsq, int32_t n_src,
                    int64_t src_stride, void* stream) {
  VL_CHECK_ARG(ptrs && sizes && chunk_tab && sumsq && n_chunks > 0 && n_src >= 1, "vl_multi_sqnorm: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* part = nullptr;
  if (int rc = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)n_chunks * sizeof(float), s)) return rc;
  multi_sqnorm_kernel<<<n_chunks, 256, 0, s>>>((const long long*)ptrs, (const long long*)sizes, reinterpret_cast<const int2*>(chunk_tab), part, n_src,
                                               src_stride);
  if (int rc = launch_check("multi_sqnorm")) return rc;
  if (int rc = launch_colreduce(part, n_chunks, 1, sumsq, nullptr, nullptr, s)) return rc;
  return scratch_free(part, s);
}

int vl_adamw_multi_clip(const int64_t* ptrs, const int64_t* sizes, const float* wds, const float* lrs, const int32_t* chunk_tab,
                        int32_t n_chunks, float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale, const float* sumsq,
                        float max_norm, int32_t n_src, int64_t src_stride, void* stream) {
  VL_CHECK_ARG(ptrs && sizes && wds && chunk_tab && sumsq && n_chunks > 0 && step >= 1 && max_norm > 0.f && n_src >= 1, "vl_adamw_multi_clip: bad arguments");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_multi_kernel<<<n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const long long*)ptrs, (const long long*)sizes, wds,
                                                                                     lrs, reinterpret_cast<const int2*>(chunk_tab), lr, beta1, beta2,
                                                                                     eps, bc1, bc2, grad_scale, sumsq, max_norm, n_src, src_stride);
  return launch_check("adamw_multi_clip");
}
}
