/* vitlens_b200 -- C ABI of the B200 (sm_100a) kernels behind the ViT-Lens hot path.
 *
 * The reference (TencentARC/ViT-Lens) is 100% Python/PyTorch: its hot path has no FFI of its
 * own; every FLOP runs in PyTorch library kernels.  This header is therefore
 * the seam a maintainer would bind from Python (ctypes); each entry point
 * names the reference call(s) it replaces (paths relative to vitlens/src/open_clip/).
 *
 * Conventions: plain pointers + sizes, no torch types.  All pointers are DEVICE pointers unless
 * a name ends in _host.  `stream` is a cudaStream_t passed as void*.  bf16 = __nv_bfloat16 bits.
 * Every function returns 0 on success, a negative VL_E* code on argument errors, or a positive
 * cudaError_t; vl_last_error() returns a static description of the last failure on this thread.
 * Matrices are row-major with an explicit leading dimension (elements).
 */
#ifndef VITLENS_B200_H
#define VITLENS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VL_ABI_VERSION 1
#define VL_EINVAL (-1)   /* bad argument (shape / alignment / null pointer) */
#define VL_ENOTSUP (-2)  /* shape outside what the kernel supports */
#define VL_EDRIVER (-3)  /* CUDA driver entry point (tensor-map encode) unavailable */

int vl_abi_version(void);
const char* vl_last_error(void);
/* debug/bring-up knobs (descriptor overrides etc.); key/value are kernel-specific, 0 resets. */
int vl_debug_set(int key, int value);

/* ---------------------------------------------------------------------------------------------
 * GEMM (tcgen05 / TMEM / TMA, bf16 x bf16 -> fp32 accumulate):
 *     D[M,N] = epilogue(alpha * sum_k A(m,k) * B(n,k))
 * A is [M,K] row-major (a_mn=0, "K-major") or given transposed as [K,M] row-major (a_mn=1);
 * B is [N,K] row-major (b_mn=0; an nn.Linear weight [out,in]) or [K,N] row-major (b_mn=1).
 * Replaces: every nn.Linear / F.linear / `@` on the path -- in_proj & out_proj of
 * nn.MultiheadAttention (transformer.py:215,252), mlp.c_fc / c_proj (transformer.py:226-234),
 * Lens to_q/to_kv/to_out and FeedForward (perceiver.py:85-154), conv1-as-GEMM (transformer.py:464-470),
 * `pooled @ self.proj` (transformer.py:786-787), and their autograd backward (dgrad / wgrad).
 */
enum {
  VL_EPI_LINEAR = 0,   /* d = alpha*acc + bias                                                   */
  VL_EPI_GELU = 1,     /* u = alpha*acc + bias; aux_out(bf16) = u if given; d = act(u)           */
  VL_EPI_RESIDUAL = 2, /* d = alpha*acc + bias + aux_in                                          */
  VL_EPI_GELU_BWD = 3, /* d = alpha*acc * act'(aux_in)     (dgrad of c_proj fused with GELU bwd) */
  VL_EPI_GEGLU = 4,    /* weight rows interleaved in 32-column groups (value|gate):
                          u = alpha*acc + bias; aux_out = u (bf16, [M,N]); d[M,N/2] = val * gelu(gate) */
};

typedef struct {
  const void* a; /* bf16 */
  const void* b; /* bf16 */
  void* d;       /* bf16, or fp32 when d_f32 */
  int32_t M, N, K;
  int64_t lda, ldb, ldd; /* elements */
  int32_t a_mn, b_mn;
  int32_t d_f32;
  int32_t accumulate; /* fp32 output only: D += result (atomic; required when split_k > 1) */
  int32_t split_k;    /* >= 1 */
  int32_t epilogue;   /* VL_EPI_* */
  int32_t act_quick;  /* 0 = erf GELU (nn.GELU), 1 = QuickGELU (transformer.py:37-40) */
  float alpha;
  const float* bias;  /* fp32 [N] or NULL */
  const void* aux_in; /* bf16 [M,N] (ld = ldaux) or NULL */
  void* aux_out;      /* bf16 [M,N] (ld = ldaux) or NULL */
  int64_t ldaux;
} VlGemmArgs;

int vl_gemm_bf16(const VlGemmArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITLENS_B200_H */
