// Input pipelines feeding the towers, on the device:
//   * Kaldi-compatible log-mel filterbank of a waveform (torchaudio.compliance.kaldi.fbank as the reference calls it,
//     modal_audio/processors/at_processor.py:854-872: htk_compat, hanning window, 25 ms / 10 ms frames, no dither, power
//     spectrum, log) + pad / crop to the target length + AST mean / std normalisation (at_processor.py:845-852);
//   * point-cloud normalisation pc_norm (modal_3d/processors/pc_processor.py:32-38);
//   * DepthNorm + Normalize of the depth channel (modal_depth/processors/transforms_rgbd.py:393-413, vt_processor.py:311-322).
#include "vl_host.h"
#include "vl_sm100.cuh"

namespace vl {

constexpr int kFftN = 512;
constexpr int kFftBins = kFftN / 2 + 1;

// One CTA per (clip, frame).  Steps follow kaldi's feature-window.cc / torchaudio _get_window + fbank:
// remove DC offset, pre-emphasis (first sample against itself), window, zero-pad to 512, radix-2 FFT in shared memory,
// power spectrum, dense mel projection (the filter matrix is built on the host exactly as get_mel_banks does), log(max(., eps)).
__global__ void __launch_bounds__(256) fbank_kernel(const float* __restrict__ wav, long long clip_stride, int n_samples, int frame_len, int frame_shift,
                                                    int n_frames, const float* __restrict__ window, const float* __restrict__ mel /*[n_mel][257]*/,
                                                    int n_mel, float preemph, int target_len, float mean, float inv_std, float* __restrict__ out) {
  __shared__ float re[kFftN], im[kFftN], pw[kFftBins + 3], red[8];
  const int frame = blockIdx.x, clip = blockIdx.y, t = threadIdx.x;
  float* orow = out + (static_cast<long long>(clip) * target_len + frame) * n_mel;
  if (frame >= n_frames) {  // zero padding up to target_len, normalised like every other frame (at_processor.py:863-866,849)
    for (int m = t; m < n_mel; m += blockDim.x) orow[m] = (0.f - mean) * inv_std;
    return;
  }
  const float* src = wav + clip * clip_stride + static_cast<long long>(frame) * frame_shift;
  float s = 0.f;
  for (int i = t; i < frame_len; i += blockDim.x) s += src[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((t & 31) == 0) red[t >> 5] = s;
  __syncthreads();
  const float dc = (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) / frame_len;
  for (int i = t; i < kFftN; i += blockDim.x) {
    float v = 0.f;
    if (i < frame_len) {
      const float cur = src[i] - dc;
      const float prev = src[i > 0 ? i - 1 : 0] - dc;
      v = (cur - preemph * prev) * window[i];
    }
    // bit-reversed placement for the in-place decimation-in-time FFT
    const unsigned r = __brev(static_cast<unsigned>(i)) >> (32 - 9);
    re[r] = v;
    im[r] = 0.f;
  }
  __syncthreads();
#pragma unroll 1
  for (int len = 2; len <= kFftN; len <<= 1) {
    const int half = len >> 1;
    const int k = t & (half - 1);          // butterfly index inside its group
    const int base = (t / half) * len;     // 256 butterflies per stage, one per thread
    float sn, cs;
    sincospif(-2.0f * k / len, &sn, &cs);
    const int a = base + k, b = a + half;
    const float xr = re[b] * cs - im[b] * sn, xi = re[b] * sn + im[b] * cs;
    const float ar = re[a], ai = im[a];
    __syncthreads();
    re[a] = ar + xr; im[a] = ai + xi;
    re[b] = ar - xr; im[b] = ai - xi;
    __syncthreads();
  }
  for (int i = t; i < kFftBins; i += blockDim.x) pw[i] = re[i] * re[i] + im[i] * im[i];
  __syncthreads();
  for (int m = t; m < n_mel; m += blockDim.x) {
    const float* f = mel + static_cast<long long>(m) * kFftBins;
    float e = 0.f;
    for (int i = 0; i < kFftBins; ++i) e = fmaf(pw[i], f[i], e);
    orow[m] = (logf(fmaxf(e, 1.1920929e-07f)) - mean) * inv_std;
  }
}

// pc_norm: subtract the centroid, divide by the largest distance from it.  One CTA per cloud; xyz = first 3 of C channels, the
// remaining channels (colours) are copied through.
__global__ void __launch_bounds__(256) pc_norm_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int C) {
  __shared__ float red[8][3];
  __shared__ float cen[3], inv;
  const int t = threadIdx.x;
  const float* p = in + static_cast<long long>(blockIdx.x) * N * C;
  float* q = out + static_cast<long long>(blockIdx.x) * N * C;
  float s[3] = {0.f, 0.f, 0.f};
  for (int i = t; i < N; i += blockDim.x) {
    s[0] += p[i * C]; s[1] += p[i * C + 1]; s[2] += p[i * C + 2];
  }
#pragma unroll
  for (int e = 0; e < 3; ++e) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[e] += __shfl_xor_sync(0xffffffffu, s[e], o);
    if ((t & 31) == 0) red[t >> 5][e] = s[e];
  }
  __syncthreads();
  if (t < 3) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][t];
    cen[t] = v / N;
  }
  __syncthreads();
  float m = 0.f;
  for (int i = t; i < N; i += blockDim.x) {
    const float x = p[i * C] - cen[0], y = p[i * C + 1] - cen[1], z = p[i * C + 2] - cen[2];
    m = fmaxf(m, x * x + y * y + z * z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5][0] = m;
  __syncthreads();
  if (t == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v = fmaxf(v, red[w][0]);
    inv = 1.0f / sqrtf(v);
  }
  __syncthreads();
  for (long long i = t; i < static_cast<long long>(N) * C; i += blockDim.x) {
    const int c = static_cast<int>(i % C);
    q[i] = c < 3 ? (p[i] - cen[c]) * inv : p[i];
  }
}

__global__ void __launch_bounds__(256) depth_norm_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, float min_depth,
                                                         float max_depth, int clamp_max, float mean, float inv_std) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float d = fmaxf(in[i], min_depth);
    if (clamp_max) d = fminf(d, max_depth);
    out[i] = (d / max_depth - mean) * inv_std;
  }
}

}  // namespace vl

using namespace vl;

extern "C" {

int vl_fbank(const float* wav, int64_t clip_stride, int32_t n_clips, int32_t n_samples, int32_t frame_len, int32_t frame_shift, const float* window,
             const float* mel, int32_t n_mel, float preemph, int32_t target_len, float mean, float std, float* out, void* stream) {
  VL_CHECK_ARG(wav && window && mel && out && n_clips > 0 && n_mel > 0 && target_len > 0 && std > 0.f, "vl_fbank: bad arguments");
  VL_CHECK_ARG(frame_len > 0 && frame_len <= kFftN && frame_shift > 0 && n_samples >= frame_len, "vl_fbank: frame_len must be <= 512 and <= n_samples");
  const int n_frames = 1 + (n_samples - frame_len) / frame_shift;  // snip_edges = true
  fbank_kernel<<<dim3(target_len, n_clips), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      wav, clip_stride, n_samples, frame_len, frame_shift, n_frames < target_len ? n_frames : target_len, window, mel, n_mel, preemph, target_len,
      mean, 1.0f / std, out);
  return launch_check("fbank");
}

int vl_pc_norm(const float* in, float* out, int32_t B, int32_t N, int32_t C, void* stream) {
  VL_CHECK_ARG(in && out && B > 0 && N > 0 && C >= 3, "vl_pc_norm: bad arguments");
  pc_norm_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, N, C);
  return launch_check("pc_norm");
}

int vl_depth_norm(const float* in, float* out, int64_t n, float min_depth, float max_depth, int32_t clamp_max, float mean, float std, void* stream) {
  VL_CHECK_ARG(in && out && n > 0 && max_depth > 0.f && std > 0.f, "vl_depth_norm: bad arguments");
  long long g = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (g > cap) g = cap;
  depth_norm_kernel<<<(unsigned)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, n, min_depth, max_depth, clamp_max, mean, 1.0f / std);
  return launch_check("depth_norm");
}
}
