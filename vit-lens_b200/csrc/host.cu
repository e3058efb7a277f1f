// Host-side runtime glue: error string, debug knobs, tensor-map encoding.
#include <stdarg.h>
#include <string.h>

#include "vl_host.h"

namespace vl {

static thread_local char g_err[512] = "";
static int g_debug[64] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int debug_get(int key) { return (key >= 0 && key < 64) ? g_debug[key] : 0; }
static long long* g_dbg_buf = nullptr;
long long* debug_buffer() { return g_dbg_buf; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  // cuTensorMapEncodeTiled is a DRIVER entry point: it needs a current context on the calling thread.  Autograd runs backward on
  // its own threads, and the first call of a backward pass may be this one (seen: CUDA_ERROR_INVALID_CONTEXT when the loss
  // backward was the first kernel of the thread) -- a runtime call binds the device's primary context.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return VL_EDRIVER;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("tensor map base %p not 16-byte aligned", base);
    return VL_EINVAL;
  }
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) {
      set_error("tensor map stride %llu not a multiple of 16 bytes", (unsigned long long)gstr[i]);
      return VL_EINVAL;
    }
  CUresult r = fn(out, dtype, rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return VL_EDRIVER;
  }
  return 0;
}

int scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
  static bool pool_set[64] = {false};
  int dev = 0;
  VL_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !pool_set[dev]) {
    cudaMemPool_t pool;
    VL_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ull;
    VL_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    pool_set[dev] = true;
  }
  VL_CUDA(cudaMallocAsync(p, bytes, s));
  return 0;
}

int scratch_free(void* p, cudaStream_t s) {
  VL_CUDA(cudaFreeAsync(p, s));
  return 0;
}

// CB = 32: 32 columns x 8 part slices per CTA (slice y sums parts y, y + 8, ...; the eight slice sums are then added in slice
// order).  CB = 1: a handful of columns with many parts (scalars): 256 part slices per column, fixed-shape tree over the slices.
template <int CB>
__global__ void __launch_bounds__(256) colreduce_kernel(const float* __restrict__ part, int nparts, long long ncols, long long n_each,
                                                        float* __restrict__ out0, float* __restrict__ out1, float* __restrict__ out2) {
  __shared__ float red[256];
  constexpr int kSlices = 256 / CB;
  const int tx = threadIdx.x % CB, ty = threadIdx.x / CB;
  const long long col = (long long)blockIdx.x * CB + tx;
  float acc = 0.f;
  if (col < ncols)
    for (int p = ty; p < nparts; p += kSlices) acc += part[(long long)p * ncols + col];
  red[threadIdx.x] = acc;
  __syncthreads();
  if constexpr (CB == 1) {
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    acc = red[0];
  } else {
    if (ty == 0) {
      acc = 0.f;
#pragma unroll
      for (int y = 0; y < kSlices; ++y) acc += red[y * CB + tx];
    }
  }
  if (ty == 0 && col < ncols) {
    const long long k = col / n_each, c = col - k * n_each;
    float* out = k == 0 ? out0 : (k == 1 ? out1 : out2);
    out[c] = acc;
  }
}

int launch_colreduce(const float* part, int nparts, long long n_each, float* out0, float* out1, float* out2, cudaStream_t s) {
  const int nout = 1 + (out1 != nullptr) + (out2 != nullptr);
  const long long ncols = n_each * nout;
  if (ncols < 32)
    colreduce_kernel<1><<<(unsigned)ncols, 256, 0, s>>>(part, nparts, ncols, n_each, out0, out1, out2);
  else
    colreduce_kernel<32><<<(unsigned)((ncols + 31) / 32), 256, 0, s>>>(part, nparts, ncols, n_each, out0, out1, out2);
  return launch_check("colreduce");
}

}  // namespace vl

extern "C" {

int vl_abi_version(void) { return VL_ABI_VERSION; }
const char* vl_last_error(void) { return vl::g_err; }
int vl_debug_buffer(void* dev_ptr) {
  vl::g_dbg_buf = reinterpret_cast<long long*>(dev_ptr);
  return 0;
}
int vl_debug_set(int key, int value) {
  if (key < 0 || key >= 64) return VL_EINVAL;
  vl::g_debug[key] = value;
  return 0;
}
}
