// Owner-side reduction of the backward "inbox".
//
// During the fused backward every rank stores its dK/dV partials for a shard (rounded once to the model dtype)
// directly into the shard owner's inbox slot over NVLink (attn_bwd_sm100.cu epilogue).  This kernel runs on the owner afterwards:
// it waits until every contributing rank has raised its "gradients landed" epoch, sums the slots that hold a
// partial for each row range in a fixed order (so the result is deterministic), writes dK/dV in the model
// dtype and finally tells every peer that the inbox may be reused.  It replaces the reference's W-hop fp32
// dK/dV ring (/root/reference/ring_flash_attn/ring_flash_attn.py:139-152) and llama3's blocking
// reduce_scatter (llama3_flash_attn_varlen.py:292-293).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "attn_common.h"
#include "comm_device.cuh"

namespace rfa {
namespace comm {

template <typename T>
__device__ __forceinline__ uint2 pack4(const float4& a);
template <>
__device__ __forceinline__ uint2 pack4<__nv_bfloat16>(const float4& a) {
  return make_uint2(Pack2<__nv_bfloat16>::pack(a.x, a.y), Pack2<__nv_bfloat16>::pack(a.z, a.w));
}
template <>
__device__ __forceinline__ uint2 pack4<__half>(const float4& a) {
  return make_uint2(Pack2<__half>::pack(a.x, a.y), Pack2<__half>::pack(a.z, a.w));
}

template <typename T>
__device__ __forceinline__ float4 unpack4(const uint2& a);
template <>
__device__ __forceinline__ float4 unpack4<__nv_bfloat16>(const uint2& a) {
  // bf16 -> fp32 is a 16-bit shift
  return make_float4(__uint_as_float(a.x << 16), __uint_as_float(a.x & 0xffff0000u), __uint_as_float(a.y << 16),
                     __uint_as_float(a.y & 0xffff0000u));
}
template <>
__device__ __forceinline__ float4 unpack4<__half>(const uint2& a) {
  const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&a.x));
  const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ uint2 ldcv_u2(const uint2* p) {
  uint2 r;
  // written by a peer over NVLink: always read from L2, never from a stale L1 line
  asm volatile("ld.global.cv.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// grid: (blocks_per_task, n_tasks).  The partials arrive in the model dtype (the backward epilogue rounds each
// rank's fp32 tile once - what flash-attn does for every block in the reference, ring_flash_attn.py:131-141) and
// are summed here in fp32 in fixed rank order.
template <typename T>
__global__ void __launch_bounds__(256) reduce_dkv_kernel(const __grid_constant__ ReduceParams p) {
  const ReduceTask t = p.tasks[blockIdx.y];
  if (threadIdx.x == 0) {
    for (int r = 0; r < p.world; ++r) {
      if ((t.src_mask >> r) & 1u) {
        if (r != p.my_rank) wait_epoch(p.my_pad + kPadDkvReady + r, p.epoch, "dkv landed", p.my_rank, r);
      }
    }
  }
  __syncthreads();
  const long long n4 = static_cast<long long>(t.rows) * p.row_elems / 4;
  const long long off4 = static_cast<long long>(t.row0) * p.row_elems / 4;
  const long long slot4 = p.slot_stride / 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  constexpr int U = 4;
  for (int which = 0; which < 2; ++which) {
    const uint2* in = reinterpret_cast<const uint2*>(static_cast<const T*>(p.inbox) + which * p.kv_stride) + off4;
    uint2* out = reinterpret_cast<uint2*>(which == 0 ? p.dk : p.dv) + off4;
    for (long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
      float4 acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < p.world; ++r) {
        if (!((t.src_mask >> r) & 1u)) continue;
        uint2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + u * stride;
          if (i < n4) v[u] = ldcv_u2(in + r * slot4 + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + u * stride;
          if (i < n4) {
            const float4 f = unpack4<T>(v[u]);
            acc[u].x += f.x;
            acc[u].y += f.y;
            acc[u].z += f.z;
            acc[u].w += f.w;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * stride;
        if (i < n4) out[i] = pack4<T>(acc[u]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t old = atomicAdd(p.ticket, 1u);
    if (old + 1u == p.ticket_target) {
      __threadfence_system();
      for (int r = 0; r < p.world; ++r) st_release_sys(p.peer_pads[r] + kPadInboxFree + p.my_rank, p.epoch);
    }
  }
}

// dQ leaves the backward kernel as an fp32 accumulator (key-tile CTAs add their partial tiles with TMA
// reduce-add).  This kernel turns it into the model dtype AND zeroes the accumulator again, so the workspace is
// ready for the next backward: one pass of our own instead of a memset before and a cast after every call.
template <typename T>
__global__ void __launch_bounds__(256) dq_finalize_kernel(float4* __restrict__ acc, uint2* __restrict__ out,
                                                          long long n4) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldcs(acc + i);
    out[i] = pack4<T>(v);
    acc[i] = zero;
  }
}

}  // namespace comm

const char* reduce_dkv_launch(int dtype, const ReduceParams& p, cudaStream_t stream) {
  if (p.n_tasks <= 0) return nullptr;
  dim3 grid(kReduceBlocksPerTask, p.n_tasks, 1);
  if (dtype == kDtypeBF16) {
    comm::reduce_dkv_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p);
  } else {
    comm::reduce_dkv_kernel<__half><<<grid, 256, 0, stream>>>(p);
  }
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

const char* dq_finalize_launch(int dtype, float* acc, void* out, long long numel, cudaStream_t stream) {
  if (numel <= 0) return nullptr;
  if (numel % 4) return "dq_finalize: element count must be a multiple of 4";
  const long long n4 = numel / 4;
  const long long want = (n4 + 255) / 256;
  const int blocks = static_cast<int>(want < 148 * 8 ? want : 148 * 8);
  if (dtype == kDtypeBF16) {
    comm::dq_finalize_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(reinterpret_cast<float4*>(acc),
                                                                        static_cast<uint2*>(out), n4);
  } else {
    comm::dq_finalize_kernel<__half><<<blocks, 256, 0, stream>>>(reinterpret_cast<float4*>(acc),
                                                                 static_cast<uint2*>(out), n4);
  }
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

}  // namespace rfa
