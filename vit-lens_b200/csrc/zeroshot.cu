// Zero-shot evaluation on the device (reference training/zero_shot.py:36-60,155-257,572-789; open_clip/zero_shot_classifier.py:27-88;
// open_clip/metrics/{accuracy,map,recall}.py): the text-template classifier reduction, row-wise top-k over a large gallery of
// similarities, and per-class average precision.  Integer results (indices, counts) are exact; every floating-point reduction
// has a fixed order, so repeated runs agree bit for bit.
#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {

__device__ __forceinline__ float zs_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[g, :] = normalize(mean_t normalize(x[g * T + t, :]))   (zero_shot_classifier.py:68-75: per-template L2 normalisation,
// mean over the class's templates, L2 normalisation of the mean).  One warp per class; E <= 32 * 32.
__global__ void __launch_bounds__(256) template_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int G, int T, int E,
                                                            long long ldo, int transpose_out) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= G) return;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int t = 0; t < T; ++t) {
    const float* row = x + (static_cast<long long>(g) * T + t) * E;
    float v[32], s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < E ? row[c] : 0.f;
      s = fmaf(v[i], v[i], s);
    }
    const float inv = 1.0f / fmaxf(sqrtf(zs_warp_sum(s)), 1e-12f);
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = fmaf(v[i], inv, acc[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    acc[i] *= 1.0f / T;
    s = fmaf(acc[i], acc[i], s);
  }
  const float inv = 1.0f / sqrtf(zs_warp_sum(s));  // the reference divides by the plain norm here (no eps)
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < E) {
      if (transpose_out) out[static_cast<long long>(c) * ldo + g] = acc[i] * inv;
      else out[static_cast<long long>(g) * ldo + c] = acc[i] * inv;
    }
  }
}

// Row-wise top-k (k <= 16) of scores[rows, cols]; ties go to the smaller column index.  One warp per row: every lane keeps the
// sorted top-k of its strided column subset in registers, then k rounds of a warp arg-max pop the global winners in order.
constexpr int kMaxTopK = 16;
__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ scores, long long ld, int rows, int cols, int k,
                                                        int* __restrict__ idx_out, float* __restrict__ val_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float bv[kMaxTopK];
  int bi[kMaxTopK];
#pragma unroll
  for (int j = 0; j < kMaxTopK; ++j) {
    bv[j] = -INFINITY;
    bi[j] = 0x7fffffff;
  }
  const float* r = scores + static_cast<long long>(row) * ld;
  for (int c = lane; c < cols; c += 32) {
    float v = r[c];
    if (v != v) v = -INFINITY;  // NaN never wins
    int ci = c;
    // insertion into the sorted (descending; equal values ordered by ascending index) list -- columns arrive in ascending
    // order, so an equal value never displaces an earlier one; once inserted, the displaced entries shift down one slot each
    bool ins = false;
#pragma unroll
    for (int j = 0; j < kMaxTopK; ++j) {
      if (j < k && (ins || v > bv[j])) {
        ins = true;
        const float tv = bv[j];
        const int ti = bi[j];
        bv[j] = v;
        bi[j] = ci;
        v = tv;
        ci = ti;
      }
    }
  }
  for (int out = 0; out < k; ++out) {
    float hv = bv[0];
    int hi = bi[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, hv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, hi, o);
      if (ov > hv || (ov == hv && oi < hi)) {
        hv = ov;
        hi = oi;
      }
    }
    if (lane == 0) {
      idx_out[static_cast<long long>(row) * k + out] = hi == 0x7fffffff ? -1 : hi;
      if (val_out) val_out[static_cast<long long>(row) * k + out] = hv;
    }
    if (bi[0] == hi && hi != 0x7fffffff) {  // the winning lane pops its head
#pragma unroll
      for (int j = 0; j + 1 < kMaxTopK; ++j) {
        bv[j] = bv[j + 1];
        bi[j] = bi[j + 1];
      }
      bv[kMaxTopK - 1] = -INFINITY;
      bi[kMaxTopK - 1] = 0x7fffffff;
    }
  }
}

// Average precision of one class per CTA, exactly sklearn.metrics.average_precision_score's definition (what
// open_clip/metrics/map.py:50 calls): AP = sum over distinct thresholds of (R_n - R_{n-1}) P_n, which with ties equals
//   (1 / P) * sum over positives i of TP(score >= s_i) / N(score >= s_i).
// Thread t takes positives t, t + 256, ... and counts over all N samples; the per-thread sums are combined in a fixed order.
__global__ void __launch_bounds__(256) average_precision_kernel(const float* __restrict__ scores, long long lds, const float* __restrict__ targets,
                                                                long long ldt, int N, int apply_sigmoid, float* __restrict__ ap_out,
                                                                int* __restrict__ npos_out) {
  extern __shared__ float sm[];  // [N] scores, then [N] targets as 0/1 floats (when they fit), else global reads
  const int c = blockIdx.x;
  float* ss = sm;
  float* st = sm + N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float s = scores[static_cast<long long>(i) * lds + c];
    if (apply_sigmoid) s = 1.0f / (1.0f + expf(-s));
    ss[i] = s;
    st[i] = targets[static_cast<long long>(i) * ldt + c] > 0.5f ? 1.f : 0.f;
  }
  __syncthreads();
  double local = 0.0;
  int npos = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    if (st[i] == 0.f) continue;
    ++npos;
    const float si = ss[i];
    int n_ge = 0, tp_ge = 0;
    for (int j = 0; j < N; ++j) {
      const bool ge = ss[j] >= si;
      n_ge += ge;
      tp_ge += ge && st[j] != 0.f;
    }
    local += static_cast<double>(tp_ge) / static_cast<double>(n_ge);
  }
  __shared__ double red[256];
  __shared__ int redn[256];
  red[threadIdx.x] = local;
  redn[threadIdx.x] = npos;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      red[threadIdx.x] += red[threadIdx.x + o];
      redn[threadIdx.x] += redn[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ap_out[c] = redn[0] > 0 ? static_cast<float>(red[0] / redn[0]) : 0.f;
    if (npos_out) npos_out[c] = redn[0];
  }
}

}  // namespace vl

using namespace vl;

extern "C" {

int vl_template_mean(const float* x, float* out, int32_t G, int32_t T, int32_t E, int64_t ldo, int32_t transpose_out, void* stream) {
  VL_CHECK_ARG(x && out && G > 0 && T > 0 && E > 0 && E <= 1024, "vl_template_mean: bad arguments (E <= 1024)");
  template_mean_kernel<<<(G + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, G, T, E, ldo, transpose_out);
  return launch_check("template_mean");
}

int vl_topk_rows(const float* scores, int64_t ld, int32_t rows, int32_t cols, int32_t k, int32_t* idx_out, float* val_out, void* stream) {
  VL_CHECK_ARG(scores && idx_out && rows > 0 && cols > 0 && k > 0 && k <= kMaxTopK && ld >= cols, "vl_topk_rows: bad arguments (k <= 16)");
  topk_rows_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(scores, ld, rows, cols, k, idx_out, val_out);
  return launch_check("topk_rows");
}

int vl_average_precision(const float* scores, int64_t lds, const float* targets, int64_t ldt, int32_t N, int32_t C, int32_t apply_sigmoid,
                         float* ap_out, int32_t* npos_out, void* stream) {
  VL_CHECK_ARG(scores && targets && ap_out && N > 0 && C > 0, "vl_average_precision: bad arguments");
  const size_t smem = static_cast<size_t>(N) * 2 * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("vl_average_precision: N=%d samples exceed the shared-memory column cache (max 25600)", N);
    return VL_ENOTSUP;
  }
  if (smem > 48 * 1024) VL_CUDA(cudaFuncSetAttribute(average_precision_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  average_precision_kernel<<<C, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(scores, lds, targets, ldt, N, apply_sigmoid, ap_out, npos_out);
  return launch_check("average_precision");
}
}
