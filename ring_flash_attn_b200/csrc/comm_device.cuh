// Device-side NVLink communication roles that live INSIDE the attention kernels.
//
// The reference rotates K/V (and fp32 dK/dV) around the ring with NCCL isend/irecv between flash_attn calls
// (/root/reference/ring_flash_attn/utils.py:98-151).  Here every GPU is one NVSwitch hop from every other, so:
//
//  * "push" CTAs (the first blocks of the attention grid) stream the rows of the local K/V shard that each
//    peer will need straight into that peer's staging buffer with posted stores over NVLink, in ring order
//    (next neighbour first), and raise a per-source "landed" flag in the peer's signal pad;
//  * compute CTAs wait on those flags only when they reach a remote segment, so the transfer of shard s+1
//    overlaps the tensor-core work on shard s inside one kernel launch;
//  * when the last compute CTA retires it tells every source that its staging buffer may be overwritten;
//  * backward compute CTAs store their dK/dV tiles directly into the owner's fp32 inbox and the last tile for
//    an owner raises that owner's "gradients landed" flag.
//
// All cross-GPU flags are monotonically increasing epochs (one per collective call), so nothing is reset.
#pragma once
#include "attn_common.h"
#include "sm100_ptx.cuh"

namespace rfa {

// Epochs are compared with wrap-around safe signed distance.
__device__ __forceinline__ bool epoch_reached(uint32_t have, uint32_t want) {
  return static_cast<int32_t>(have - want) >= 0;
}
__device__ __forceinline__ void wait_epoch(const uint32_t* flag, uint32_t want, const char* what, int who = -1,
                                           int idx = -1) {
  if (epoch_reached(ld_acquire_sys(flag), want)) return;
  const uint64_t t0 = global_timer_ns();
  while (!epoch_reached(ld_acquire_sys(flag), want)) {
    if (global_timer_ns() - t0 > RFA_WATCHDOG_NS) {
      printf("rfa: epoch wait timeout (%s) rank %d peer/flag %d block %d want %u have %u\n", what, who, idx,
             blockIdx.x, want, ld_acquire_sys(flag));
      __trap();
    }
  }
}

__device__ __forceinline__ uint4 ld_nc_v4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_v4(uint4* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Body of a push CTA.  Tasks are sorted by destination in ring order; CTA b takes tasks b, b + n, ...
__device__ __forceinline__ void push_role(const PushParams& pp) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  __shared__ int s_last;
  for (int ti = blockIdx.x; ti < pp.n_tasks; ti += pp.n_ctas) {
    const PushTask t = pp.tasks[ti];
    if (tid == 0 && pp.epoch > 2) {
      // the destination must have finished reading what we pushed into this staging parity two calls ago
      wait_epoch(pp.my_pad + kPadConsumed + t.dst, pp.epoch - 2, "staging reuse", pp.my_rank, t.dst);
    }
    __syncthreads();
    const int vec_per_row = pp.row_bytes >> 4;
    const long long total = static_cast<long long>(t.rows) * vec_per_row;
    const long long src_pitch = pp.src_row_bytes[t.which];
    const char* src = pp.src_base[t.which] + t.src_row * src_pitch;
    char* dst = pp.stage_ptrs[t.dst] + pp.parity_off + t.dst_off;
    constexpr int U = 8;
    for (long long base = static_cast<long long>(tid); base < total; base += static_cast<long long>(nthr) * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = base + static_cast<long long>(u) * nthr;
        if (e < total) {
          const long long r = e / vec_per_row, c = e - r * vec_per_row;
          v[u] = ld_nc_v4(reinterpret_cast<const uint4*>(src + r * src_pitch + c * 16));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = base + static_cast<long long>(u) * nthr;
        if (e < total) st_v4(reinterpret_cast<uint4*>(dst + e * 16), v[u]);
      }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const uint32_t old = atomicAdd(pp.sent_count + t.dst, 1u);
      s_last = (old + 1u == pp.sent_target[t.dst]) ? 1 : 0;
      if (s_last) {
        __threadfence_system();
        st_release_sys(pp.peer_pads[t.dst] + kPadKvReady + pp.my_rank, pp.epoch);
      }
    }
    __syncthreads();
  }
}

// Called by one thread of every compute CTA after the CTA has finished reading staged K/V.
__device__ __forceinline__ void consumer_done(const SignalParams& sp) {
  if (sp.peer_pads == nullptr) return;
  __threadfence();
  const uint32_t old = atomicAdd(sp.done_count, 1u);
  if (old + 1u == sp.done_target) {
    __threadfence_system();
    for (int r = 0; r < sp.world; ++r) {
      if (r == sp.my_rank) continue;
      st_release_sys(sp.peer_pads[r] + kPadConsumed + sp.my_rank, sp.epoch);
    }
  }
}

}  // namespace rfa
