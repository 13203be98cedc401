// LSE layout converters for packed (varlen) batches - the CUDA counterpart of the two Triton kernels the
// reference ships (/root/reference/ring_flash_attn/triton_utils.py:6-36,70-100).
//   flatten:   (batch, H, max_seqlen)  ->  (H, total)
//   unflatten: (total, H)              ->  (batch, H, max_seqlen)   (padding untouched)
#include "attn_common.h"

namespace rfa {
namespace lse {

// grid: (ceil(max_seqlen / 256), H, batch)
__global__ void flatten_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ cu,
                               int heads, int max_seqlen, int total) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int start = cu[b], len = cu[b + 1] - start;
  if (s < len) out[static_cast<size_t>(h) * total + start + s] = in[(static_cast<size_t>(b) * heads + h) * max_seqlen + s];
}

__global__ void unflatten_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ cu,
                                 int heads, int max_seqlen) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int start = cu[b], len = cu[b + 1] - start;
  if (s < len) out[(static_cast<size_t>(b) * heads + h) * max_seqlen + s] = in[static_cast<size_t>(start + s) * heads + h];
}

}  // namespace lse

const char* lse_flatten_launch(const float* in, float* out, const int* cu, int batch, int heads, int max_seqlen,
                               int total, cudaStream_t stream) {
  if (batch == 0 || heads == 0 || max_seqlen == 0) return nullptr;
  dim3 grid((max_seqlen + 255) / 256, heads, batch);
  lse::flatten_kernel<<<grid, 256, 0, stream>>>(in, out, cu, heads, max_seqlen, total);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

const char* lse_unflatten_launch(const float* in, float* out, const int* cu, int batch, int heads, int max_seqlen,
                                 cudaStream_t stream) {
  if (batch == 0 || heads == 0 || max_seqlen == 0) return nullptr;
  dim3 grid((max_seqlen + 255) / 256, heads, batch);
  lse::unflatten_kernel<<<grid, 256, 0, stream>>>(in, out, cu, heads, max_seqlen);
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

}  // namespace rfa
