// Point-cloud tokenizer kernels (reference modal_3d/models/pointbert: misc.py:48-68 fps, dvae.py:107-176 knn + grouping,
// dvae.py:196-212 mini-PointNet).  Index work (FPS order, kNN sets) is exact integer/compare logic on fp32 distances
// computed without FMA contraction, in the reference's operation order, so the selected indices match the CPU oracle.
#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {

// ------------------------------------------------------------------------------------ farthest point sampling
// One CTA per sample, 1024 threads, points in registers (<= 16 per thread).  Each of the `npoint` iterations: update the
// running min-distance to the chosen set, block-wide arg-max (first index wins ties), broadcast the winner.
constexpr int kFpsThreads = 1024;
constexpr int kFpsMaxPer = 8;  // N <= 8192 (vitlensL point clouds)

__device__ __forceinline__ void argmax_combine(float& v, int& i, float v2, int i2) {
  if (v2 > v || (v2 == v && i2 < i)) {
    v = v2;
    i = i2;
  }
}

__global__ void __launch_bounds__(kFpsThreads) fps_kernel(const float* __restrict__ xyz, const long long* __restrict__ start, int N, int npoint,
                                                          long long* __restrict__ idx_out, float* __restrict__ centers) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ float s_c[3];
  __shared__ int s_far;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* p = xyz + static_cast<long long>(b) * N * 3;
  float px[kFpsMaxPer], py[kFpsMaxPer], pz[kFpsMaxPer], dist[kFpsMaxPer];
#pragma unroll
  for (int k = 0; k < kFpsMaxPer; ++k) {
    const int i = tid + k * kFpsThreads;
    if (i < N) {
      px[k] = p[3 * i];
      py[k] = p[3 * i + 1];
      pz[k] = p[3 * i + 2];
    } else {
      px[k] = py[k] = pz[k] = 0.f;
    }
    dist[k] = 1e10f;
  }
  int far = static_cast<int>(start[b]);
  for (int it = 0; it < npoint; ++it) {
    if (tid == 0) {
      idx_out[static_cast<long long>(b) * npoint + it] = far;
      s_c[0] = p[3 * far];
      s_c[1] = p[3 * far + 1];
      s_c[2] = p[3 * far + 2];
      centers[(static_cast<long long>(b) * npoint + it) * 3] = s_c[0];
      centers[(static_cast<long long>(b) * npoint + it) * 3 + 1] = s_c[1];
      centers[(static_cast<long long>(b) * npoint + it) * 3 + 2] = s_c[2];
    }
    __syncthreads();
    const float cx = s_c[0], cy = s_c[1], cz = s_c[2];
    float bv = -1.f;
    int bi = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < kFpsMaxPer; ++k) {
      const int i = tid + k * kFpsThreads;
      if (i < N) {
        const float dx = __fsub_rn(px[k], cx), dy = __fsub_rn(py[k], cy), dz = __fsub_rn(pz[k], cz);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        dist[k] = fminf(dist[k], d);
        argmax_combine(bv, bi, dist[k], i);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v2 = __shfl_xor_sync(0xffffffffu, bv, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      argmax_combine(bv, bi, v2, i2);
    }
    if (lane == 0) {
      s_val[warp] = bv;
      s_idx[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
      bv = s_val[lane];
      bi = s_idx[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_xor_sync(0xffffffffu, bv, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
        argmax_combine(bv, bi, v2, i2);
      }
      if (lane == 0) s_far = bi;
    }
    __syncthreads();
    far = s_far;
  }
}

// ------------------------------------------------------------------------------------ kNN grouping
// One warp per centre: keeps the `k` (<= 32) smallest squared distances, one per lane; scans the points 32 at a time
// and replaces the current worst entry whenever a closer point shows up.  Writes neighbours minus the centre
// (dvae.py:162-176) as [B*G*k, 3] fp32 and, padded with zeros to 8 columns, as bf16 rows for the first GEMM.
__global__ void __launch_bounds__(256) knn_group_kernel(const float* __restrict__ xyz, const float* __restrict__ centers, int N, int G, int k,
                                                        long long total_centers, float* __restrict__ nb_out, long long* __restrict__ idx_out) {
  const int lane = threadIdx.x & 31;
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (w >= total_centers) return;
  const long long b = w / G;
  const float* p = xyz + b * N * 3;
  const float cx = centers[w * 3], cy = centers[w * 3 + 1], cz = centers[w * 3 + 2];
  const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
  float best = (lane < k) ? INFINITY : -INFINITY;  // lanes >= k never hold candidates
  int best_i = -1;
  float worst = INFINITY;
  int worst_lane = 0;
  for (int base = 0; base < N; base += 32) {
    const int i = base + lane;
    float d = INFINITY;
    if (i < N) {
      const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
      // square_distance (dvae.py:121-140): -2 * <c, x> + |c|^2 + |x|^2
      const float dot = __fadd_rn(__fadd_rn(__fmul_rn(cx, x), __fmul_rn(cy, y)), __fmul_rn(cz, z));
      const float x2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
      d = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), c2), x2);
    }
    unsigned m = __ballot_sync(0xffffffffu, d < worst);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float dv = __shfl_sync(0xffffffffu, d, src);
      if (dv < worst) {  // worst may have tightened since the ballot
        if (lane == worst_lane) {
          best = dv;
          best_i = base + src;
        }
        // recompute the worst entry among the k candidate lanes
        float wv = (lane < k) ? best : -INFINITY;
        int wl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float v2 = __shfl_xor_sync(0xffffffffu, wv, o);
          const int l2 = __shfl_xor_sync(0xffffffffu, wl, o);
          if (v2 > wv || (v2 == wv && l2 < wl)) {
            wv = v2;
            wl = l2;
          }
        }
        worst = wv;
        worst_lane = wl;
      }
    }
  }
  if (lane < k) {
    const long long o = (w * k + lane);
    nb_out[o * 3] = p[3 * best_i] - cx;
    nb_out[o * 3 + 1] = p[3 * best_i + 1] - cy;
    nb_out[o * 3 + 2] = p[3 * best_i + 2] - cz;
    if (idx_out) idx_out[o] = best_i;
  }
}

// ------------------------------------------------------------------------------------ tiny-K linear: out = act(x[R,3] @ W[C,3]^T * scale + shift)
// first_conv.0 (+ folded BatchNorm + ReLU, dvae.py:200-203) and pos_embed.0 (+ GELU, point_encoder.py:325-327).
// act: 0 none, 1 relu, 2 gelu(erf).  scale/shift are per-output-channel fp32 (bias and BN already folded in).
__global__ void __launch_bounds__(256) linear3_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, __nv_bfloat16* __restrict__ out,
                                                      __nv_bfloat16* __restrict__ pre_out, long long R, int C, int act) {
  const int cvec = C >> 3;
  const long long total = R * cvec;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    const long long r = idx / cvec;
    const float x0 = x[3 * r], x1 = x[3 * r + 1], x2 = x[3 * r + 2];
    float f[8], pre[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cv * 8 + e;
      float v = x0 * __ldg(w + 3 * c) + x1 * __ldg(w + 3 * c + 1) + x2 * __ldg(w + 3 * c + 2);
      v = v * __ldg(scale + c) + __ldg(shift + c);
      pre[e] = v;
      if (act == 1) v = fmaxf(v, 0.f);
      if (act == 2) v = gelu_erf_fwd(v);
      f[e] = v;
    }
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(out + r * C + cv * 8) = u;
    if (pre_out) {
      uint4 q;
      q.x = pack_bf16(pre[0], pre[1]); q.y = pack_bf16(pre[2], pre[3]); q.z = pack_bf16(pre[4], pre[5]); q.w = pack_bf16(pre[6], pre[7]);
      *reinterpret_cast<uint4*>(pre_out + r * C + cv * 8) = q;
    }
  }
}

// ------------------------------------------------------------------------------------ max over groups of G consecutive rows
// out[g, c] = max_{r in group g} x[g*G + r, c]  (torch.max(feature, dim=2), dvae.py:205,210); optional arg-max rows.
__global__ void __launch_bounds__(256) group_max_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int* __restrict__ arg,
                                                        long long groups, int G, int C) {
  const int cvec = C >> 3;
  const long long total = groups * cvec;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    const long long g = idx / cvec;
    float best[8];
    int bi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      best[e] = -INFINITY;
      bi[e] = 0;
    }
    for (int r = 0; r < G; ++r) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (g * G + r) * C + cv * 8);
      const float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (f[e] > best[e]) {
          best[e] = f[e];
          bi[e] = r;
        }
    }
    uint4 o;
    o.x = pack_bf16(best[0], best[1]); o.y = pack_bf16(best[2], best[3]); o.z = pack_bf16(best[4], best[5]); o.w = pack_bf16(best[6], best[7]);
    *reinterpret_cast<uint4*>(out + g * C + cv * 8) = o;
    if (arg) {
#pragma unroll
      for (int e = 0; e < 8; ++e) arg[g * C + cv * 8 + e] = bi[e];
    }
  }
}


// ------------------------------------------------------------------------------------ backward helpers
// dx[g*G + r, c] = (arg[g, c] == r) ? dout[g, c] : 0   (backward of the per-group max)
__global__ void __launch_bounds__(256) group_max_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const int* __restrict__ arg,
                                                            __nv_bfloat16* __restrict__ dx, long long groups, int G, int C) {
  const int cvec = C >> 3;
  const long long total = groups * cvec;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    const long long g = idx / cvec;
    const uint4 u = *reinterpret_cast<const uint4*>(dout + g * C + cv * 8);
    const uint16_t* h = reinterpret_cast<const uint16_t*>(&u);
    int a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = arg[g * C + cv * 8 + e];
    for (int r = 0; r < G; ++r) {
      uint16_t o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (a[e] == r) ? h[e] : static_cast<uint16_t>(0);
      *reinterpret_cast<uint4*>(dx + (g * G + r) * C + cv * 8) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

// out[g, c] = sum_r x[g*G + r, c]
__global__ void __launch_bounds__(256) group_sum_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, long long groups, int G,
                                                        int C) {
  const int cvec = C >> 3;
  const long long total = groups * cvec;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % cvec);
    const long long g = idx / cvec;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < G; ++r) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (g * G + r) * C + cv * 8);
      acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
      acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
    }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]); o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + g * C + cv * 8) = o;
  }
}

// part[blockIdx.y][0][n] = sum_t a[t,n], part[blockIdx.y][1][n] = sum_t a[t,n] * b[t,n] over this CTA's rows (BatchNorm / bias
// gradients behind a ReLU; second stage: launch_colreduce)
__global__ void __launch_bounds__(256) colsum2_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, float* __restrict__ part,
                                                      long long T, int N, long long rows_per_cta) {
  __shared__ float red[2][8][256];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(T, r0 + rows_per_cta);
  float a1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    for (long long r = r0 + ty; r < r1; r += 8) {
      const uint4 ua = *reinterpret_cast<const uint4*>(a + r * N + col);
      const uint4 ub = *reinterpret_cast<const uint4*>(b + r * N + col);
      const float fa[8] = {bf16_lo(ua.x), bf16_hi(ua.x), bf16_lo(ua.y), bf16_hi(ua.y), bf16_lo(ua.z), bf16_hi(ua.z), bf16_lo(ua.w), bf16_hi(ua.w)};
      const float fb[8] = {bf16_lo(ub.x), bf16_hi(ub.x), bf16_lo(ub.y), bf16_hi(ub.y), bf16_lo(ub.z), bf16_hi(ub.z), bf16_lo(ub.w), bf16_hi(ub.w)};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a1[e] += fa[e];
        a2[e] += fa[e] * fb[e];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[0][ty][tx * 8 + e] = a1[e];
    red[1][ty][tx * 8 + e] = a2[e];
  }
  __syncthreads();
  const int c = threadIdx.x;
  float v1 = 0.f, v2 = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    v1 += red[0][y][c];
    v2 += red[1][y][c];
  }
  const int gc = blockIdx.x * 256 + c;
  if (gc < N) {
    float* pp = part + static_cast<long long>(blockIdx.y) * 2 * N;
    pp[gc] = v1;
    pp[N + gc] = v2;
  }
}

// part[blockIdx.y][c, j] = sum over this CTA's rows of dy[r, c] * x[r, j]   (weight gradient of a 3-input linear layer)
__global__ void __launch_bounds__(256) wgrad3_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, float* __restrict__ part, long long R,
                                                     int C, long long rows_per_cta) {
  // thread = one channel (blockIdx.x * 256 + tid), loops over the CTA's row range
  const int c = blockIdx.x * 256 + threadIdx.x;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(R, r0 + rows_per_cta);
  if (c >= C) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float g = __bfloat162float(dy[r * C + c]);
    a0 = fmaf(g, __ldg(x + 3 * r), a0);
    a1 = fmaf(g, __ldg(x + 3 * r + 1), a1);
    a2 = fmaf(g, __ldg(x + 3 * r + 2), a2);
  }
  float* dw = part + static_cast<long long>(blockIdx.y) * 3 * C;
  dw[3 * c] = a0;
  dw[3 * c + 1] = a1;
  dw[3 * c + 2] = a2;
}


// part[blockIdx.x][0..2] = sum_r x[r, :],  part[blockIdx.x][3 + 3 i + j] = sum_r x[r, i] * x[r, j]   (x: [R, 3] fp32, this CTA's rows)
// First and second moments of the centre-normalised neighbourhood points: first_conv.0 is linear in them, so the batch
// statistics of its 128 outputs (BatchNorm in training mode, dvae.py:185-188) follow in closed form from these 12 numbers.
__global__ void __launch_bounds__(256) moments3_kernel(const float* __restrict__ x, float* __restrict__ part, long long R) {
  float* out = part + static_cast<long long>(blockIdx.x) * 12;
  float s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // x y z xx xy xz yy yz zz
  for (long long r = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; r < R; r += static_cast<long long>(gridDim.x) * 256) {
    const float a = x[3 * r], b = x[3 * r + 1], c = x[3 * r + 2];
    s[0] += a; s[1] += b; s[2] += c;
    s[3] = fmaf(a, a, s[3]); s[4] = fmaf(a, b, s[4]); s[5] = fmaf(a, c, s[5]);
    s[6] = fmaf(b, b, s[6]); s[7] = fmaf(b, c, s[7]); s[8] = fmaf(c, c, s[8]);
  }
  __shared__ float red[8][9];
#pragma unroll
  for (int e = 0; e < 9; ++e) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[e] += __shfl_xor_sync(0xffffffffu, s[e], o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
#pragma unroll
    for (int e = 0; e < 9; ++e) red[w][e] = s[e];
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float v = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) v += red[y][threadIdx.x];
    // x y z | xx xy xz | xy yy yz | xz yz zz
    const int e = threadIdx.x;
    if (e < 3) {
      out[e] = v;
    } else {
      const int ij[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
      const int i = ij[e - 3][0], j = ij[e - 3][1];
      out[3 + 3 * i + j] = v;
      if (i != j) out[3 + 3 * j + i] = v;
    }
  }
}

// out[r, c] = act(p0[c] * a[r, c] + p1[c] * b[r, c] + p2[c])   (b / p1 optional; act: 0 none, 1 relu)
// BatchNorm with batch statistics around the grouped GEMMs: forward normalise + ReLU (a = pre-BN activations), backward
// dz = s (dy - mean dy - xhat mean(dy xhat)) (a = dy, b = pre-BN activations) -- one read of each operand, one write.
__global__ void __launch_bounds__(256) col_affine_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                         const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                                                         __nv_bfloat16* __restrict__ out, long long nvec, int cvec, int act) {
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < nvec; i += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % cvec) * 8;
    const uint4 ua = reinterpret_cast<const uint4*>(a)[i];
    float v[8] = {bf16_lo(ua.x), bf16_hi(ua.x), bf16_lo(ua.y), bf16_hi(ua.y), bf16_lo(ua.z), bf16_hi(ua.z), bf16_lo(ua.w), bf16_hi(ua.w)};
    const float4 q0 = *reinterpret_cast<const float4*>(p0 + c), q1 = *reinterpret_cast<const float4*>(p0 + c + 4);
    const float4 r0 = *reinterpret_cast<const float4*>(p2 + c), r1 = *reinterpret_cast<const float4*>(p2 + c + 4);
    const float m0[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    const float m2[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf(m0[e], v[e], m2[e]);
    if (b != nullptr) {
      const uint4 ub = reinterpret_cast<const uint4*>(b)[i];
      const float w[8] = {bf16_lo(ub.x), bf16_hi(ub.x), bf16_lo(ub.y), bf16_hi(ub.y), bf16_lo(ub.z), bf16_hi(ub.z), bf16_lo(ub.w), bf16_hi(ub.w)};
      const float4 t0 = *reinterpret_cast<const float4*>(p1 + c), t1 = *reinterpret_cast<const float4*>(p1 + c + 4);
      const float m1[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(m1[e], w[e], v[e]);
    }
    if (act == 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

}  // namespace vl

using namespace vl;

extern "C" {

int vl_fps(const float* xyz, const int64_t* start, int32_t B, int32_t N, int32_t npoint, int64_t* idx_out, float* centers, void* stream) {
  VL_CHECK_ARG(xyz && start && idx_out && centers && B > 0 && N > 0 && npoint > 0, "vl_fps: bad arguments");
  if (N > kFpsThreads * kFpsMaxPer) {
    set_error("vl_fps: N=%d > %d not supported", N, kFpsThreads * kFpsMaxPer);
    return VL_ENOTSUP;
  }
  fps_kernel<<<B, kFpsThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(xyz, (const long long*)start, N, npoint, (long long*)idx_out, centers);
  return launch_check("fps");
}

int vl_knn_group(const float* xyz, const float* centers, int32_t B, int32_t N, int32_t G, int32_t k, float* nb_out, int64_t* idx_out,
                 void* stream) {
  VL_CHECK_ARG(xyz && centers && nb_out && B > 0 && N > 0 && G > 0 && k > 0 && k <= 32 && k <= N, "vl_knn_group: bad arguments (k must be <= 32)");
  const long long total = static_cast<long long>(B) * G;
  const long long blocks = (total * 32 + 255) / 256;
  knn_group_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(xyz, centers, N, G, k, total, nb_out, (long long*)idx_out);
  return launch_check("knn_group");
}

int vl_linear3(const float* x, const float* w, const float* scale, const float* shift, void* out, void* pre_out, int64_t R, int32_t C, int32_t act,
               void* stream) {
  VL_CHECK_ARG(x && w && scale && shift && out && R > 0 && C > 0 && C % 8 == 0 && act >= 0 && act <= 2, "vl_linear3: bad arguments");
  long long g = (R * (C / 8) + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  linear3_kernel<<<(unsigned)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, w, scale, shift, reinterpret_cast<__nv_bfloat16*>(out),
                                                                                   reinterpret_cast<__nv_bfloat16*>(pre_out), R, C, act);
  return launch_check("linear3");
}

int vl_group_max(const void* x, void* out, int32_t* arg, int64_t groups, int32_t G, int32_t C, void* stream) {
  VL_CHECK_ARG(x && out && groups > 0 && G > 0 && C > 0 && C % 8 == 0, "vl_group_max: bad arguments");
  long long g = (groups * (C / 8) + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  group_max_kernel<<<(unsigned)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                                     reinterpret_cast<__nv_bfloat16*>(out), arg, groups, G, C);
  return launch_check("group_max");
}

static inline unsigned pc_grid(long long items) {
  long long g = (items + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  return static_cast<unsigned>(g < 1 ? 1 : g);
}

int vl_group_max_bwd(const void* dout, const int32_t* arg, void* dx, int64_t groups, int32_t G, int32_t C, void* stream) {
  VL_CHECK_ARG(dout && arg && dx && groups > 0 && G > 0 && C > 0 && C % 8 == 0, "vl_group_max_bwd: bad arguments");
  group_max_bwd_kernel<<<pc_grid(groups * (C / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dout), arg, reinterpret_cast<__nv_bfloat16*>(dx), groups, G, C);
  return launch_check("group_max_bwd");
}

int vl_group_sum(const void* x, void* out, int64_t groups, int32_t G, int32_t C, void* stream) {
  VL_CHECK_ARG(x && out && groups > 0 && G > 0 && C > 0 && C % 8 == 0, "vl_group_sum: bad arguments");
  group_sum_kernel<<<pc_grid(groups * (C / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), groups, G, C);
  return launch_check("group_sum");
}

int vl_colsum2_bf16(const void* a, const void* b, float* s1, float* s2, int64_t T, int32_t N, void* stream) {
  VL_CHECK_ARG(a && b && s1 && s2 && T > 0 && N > 0 && N % 8 == 0, "vl_colsum2_bf16: bad arguments");
  const int gx = (N + 255) / 256;
  long long gy = (static_cast<long long>(num_sms()) * 4 + gx - 1) / gx;
  if (gy > (T + 63) / 64) gy = (T + 63) / 64;
  if (gy < 1) gy = 1;
  const long long rows_per = (T + gy - 1) / gy;
  gy = (T + rows_per - 1) / rows_per;  // no empty row ranges: every partial row is written
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* part = nullptr;
  if (int rc = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)gy * 2 * N * sizeof(float), s)) return rc;
  colsum2_kernel<<<dim3(gx, (unsigned)gy), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b), part, T, N, rows_per);
  if (int rc = launch_check("colsum2")) return rc;
  if (int rc = launch_colreduce(part, (int)gy, N, s1, s2, nullptr, s)) return rc;
  return scratch_free(part, s);
}

int vl_wgrad3(const void* dy, const float* x, float* dw, int64_t R, int32_t C, void* stream) {
  VL_CHECK_ARG(dy && x && dw && R > 0 && C > 0, "vl_wgrad3: bad arguments");
  const int gx = (C + 255) / 256;
  long long gy = (static_cast<long long>(num_sms()) * 8 + gx - 1) / gx;
  if (gy > (R + 255) / 256) gy = (R + 255) / 256;
  if (gy < 1) gy = 1;
  const long long rows_per = (R + gy - 1) / gy;
  gy = (R + rows_per - 1) / rows_per;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* part = nullptr;
  if (int rc = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)gy * 3 * C * sizeof(float), s)) return rc;
  wgrad3_kernel<<<dim3(gx, (unsigned)gy), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy), x, part, R, C, rows_per);
  if (int rc = launch_check("wgrad3")) return rc;
  if (int rc = launch_colreduce(part, (int)gy, 3ll * C, dw, nullptr, nullptr, s)) return rc;
  return scratch_free(part, s);
}

int vl_moments3(const float* x, float* out12, int64_t R, void* stream) {
  VL_CHECK_ARG(x && out12 && R > 0, "vl_moments3: bad arguments");
  long long g = (R + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 4;
  if (g > cap) g = cap;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* part = nullptr;
  if (int rc = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)g * 12 * sizeof(float), s)) return rc;
  moments3_kernel<<<(unsigned)g, 256, 0, s>>>(x, part, R);
  if (int rc = launch_check("moments3")) return rc;
  if (int rc = launch_colreduce(part, (int)g, 12, out12, nullptr, nullptr, s)) return rc;
  return scratch_free(part, s);
}

int vl_col_affine_bf16(const void* a, const void* b, const float* p0, const float* p1, const float* p2, void* out, int64_t R, int32_t C,
                       int32_t act, void* stream) {
  VL_CHECK_ARG(a && p0 && p2 && out && R > 0 && C > 0 && C % 8 == 0, "vl_col_affine_bf16: bad arguments (C must be a multiple of 8)");
  VL_CHECK_ARG((b == nullptr) == (p1 == nullptr), "vl_col_affine_bf16: b and p1 go together");
  VL_CHECK_ARG(act == 0 || act == 1, "vl_col_affine_bf16: act must be 0 (none) or 1 (relu)");
  const long long nvec = static_cast<long long>(R) * (C / 8);
  long long g = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (g > cap) g = cap;
  col_affine_kernel<<<(unsigned)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b), p0, p1, p2, reinterpret_cast<__nv_bfloat16*>(out), nvec,
      C / 8, act);
  return launch_check("col_affine");
}
}
