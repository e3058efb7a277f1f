// Host-side helpers shared by the .cu files: error reporting, tensor-map encode (driver entry point
// resolved at run time so the library loads on a machine without libcuda), device properties.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vitlens_b200.h"

namespace vl {

void set_error(const char* fmt, ...);
int num_sms();
int debug_get(int key);
long long* debug_buffer();  // optional device buffer for clock64 timelines (bring-up only)

// 2-D / 3-D bf16 tensor map, row-major global tensor, 128B swizzle (or none), zero OOB fill.
// dims[0] is the contiguous dimension.  strides_bytes[i] is the stride of dims[i+1].
int make_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);

inline int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                             uint64_t outer_stride_elems, uint32_t box_inner, uint32_t box_outer,
                             bool swizzle128 = true) {
  uint64_t dims[2] = {inner, outer};
  uint64_t strides[1] = {outer_stride_elems * 2};
  uint32_t box[2] = {box_inner, box_outer};
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, swizzle128);
}

#define VL_CHECK_ARG(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      vl::set_error(__VA_ARGS__);    \
      return VL_EINVAL;              \
    }                                \
  } while (0)

#define VL_CUDA(expr)                                                           \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      vl::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));            \
      return static_cast<int>(_e);                                              \
    }                                                                           \
  } while (0)

// Stream-ordered scratch memory for the per-CTA partial results of the deterministic reductions below: cudaMallocAsync /
// cudaFreeAsync on the device's default memory pool (release threshold raised once so blocks are recycled instead of handed
// back to the driver).  Ordered on `s` like the kernels that use it, so concurrent callers on different streams never share a
// buffer.
int scratch_alloc(void** p, size_t bytes, cudaStream_t s);
int scratch_free(void* p, cudaStream_t s);
// Second stage of every cross-CTA reduction on the path (LayerNorm / bias / BatchNorm parameter gradients, loss sums, gradient
// norms): out_k[c] = sum over p = 0 .. nparts-1, IN THAT ORDER, of part[p * (nout * n_each) + k * n_each + c].  First stages
// write one partial row per CTA with plain stores; results are bit-identical from run to run (no floating-point atomics) and
// the outputs need no zero fill.
int launch_colreduce(const float* part, int nparts, long long n_each, float* out0, float* out1, float* out2, cudaStream_t s);

// attention_bwd2.cu
int launch_attn_bwd2(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                     int B, int H, int nq, int nk, long long ldq, long long ldk, long long ldv, long long ldo, long long lddo, long long lddq,
                     long long lddk, long long lddv, float scale, int causal, cudaStream_t stream);

// attention_bwd3.cu (non-causal)
int launch_attn_bwd3(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq, void* dk, void* dv,
                     int B, int H, int nq, int nk, long long ldq, long long ldk, long long ldv, long long ldo, long long lddo, long long lddq,
                     long long lddk, long long lddv, float scale, cudaStream_t stream);

// attention_fwd2.cu
int launch_attn_fwd2(const void* q, const void* k, const void* v, void* o, float* lse, int B, int H, int nq, int nk, long long ldq, long long ldk,
                     long long ldv, long long ldo, float scale, cudaStream_t stream);

inline int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace vl
