// Descriptor probe: single-CTA tcgen05 GEMMs in exactly the operand forms the attention kernels use.
// Each mode multiplies small matrices whose fp32 product is checked against torch on the host side, so a
// wrong UMMA shared-memory / instruction descriptor is localised without touching the attention kernels.
// LBO/SBO/k-step values can be overridden at run time to sweep alternatives in one GPU session.
#include <stdio.h>

#include "attn_common.h"
#include "sm100_ptx.cuh"

namespace rfa {
namespace probe {

struct Smem {
  uint64_t full;
  uint64_t done;
  uint32_t tmem_base;
};

// Throughput loop: warp-uniform control flow, constant-step descriptors, one elected lane issues - exactly the
// issue pattern of the attention kernels - so the measured cycles are the tensor pipe's, not the issuer's.
template <int A_KIND, int B_KIND, int N, int KSTEPS>
__device__ __forceinline__ void rate_loop(uint32_t tmem, uint32_t sa_u, uint32_t sb_u, int reps, bool leader) {
  constexpr uint32_t hi = umma_desc_hi(1024, kSwizzle128B);
  constexpr int b_rows = B_KIND == 0 ? N : KSTEPS * 16;
  constexpr uint32_t idesc = umma_idesc_f16(1, 128, N, A_KIND == 3, B_KIND != 0);
  const uint32_t b0 = umma_desc_lo(sb_u, B_KIND == 1 ? b_rows * 128 : 16);
  const uint32_t a0 = A_KIND == 1 ? tmem + 256 : umma_desc_lo(sa_u, A_KIND == 3 ? 16384 : 16);
  for (int rep = 0; rep < reps; ++rep) {
    if (leader) {
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k) {
        const uint32_t bo = B_KIND == 0 ? (((k >> 2) * (b_rows * 128) + (k & 3) * 32) >> 4) : (k * 2048) >> 4;
        if constexpr (A_KIND == 1) {
          umma_ts2(tmem, a0 + k * 8, b0 + bo, hi, idesc, 1u);
        } else {
          const uint32_t ao = A_KIND == 0 ? (((k >> 2) * 16384 + (k & 3) * 32) >> 4)
                                          : (A_KIND == 2 ? (k * 32) >> 4 : (k * 2048) >> 4);
          umma_ss2(tmem, a0 + ao, hi, b0 + bo, hi, idesc, 1u);
        }
      }
    }
    __syncwarp();
  }
}

// operand kinds
//  A: 0 = smem K-major [128 rows][128 k] via TMA      (Q/K/V as A)
//     1 = TMEM bf16 [128 rows][kdim]                   (P, P^T)
//     2 = smem K-major [128 rows][64 k] written by threads with the 128B swizzle (dS^T as A)
//     3 = smem MN-major from a [128 k][128 m] TMA tile (K as A of dQ^T)
//  B: 0 = smem K-major [n rows][128 k] via TMA         (K as B, Q/dO as B, n = 128 or 64)
//     1 = smem MN-major from a [kdim rows][128 n] TMA tile (V, Q, dO as B; kdim = 128 or 64)
//     2 = smem MN-major from thread-written swizzled [128 k][64 n] (dS^T as B)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
             const __nv_bfloat16* __restrict__ a_raw, const __nv_bfloat16* __restrict__ b_raw,
             float* __restrict__ out, const ProbeConfig c) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;             // up to 32 KB
  uint8_t* sb = smem + 32768;     // up to 32 KB
  Smem* sm = reinterpret_cast<Smem*>(smem + 65536);
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(&sm->full, 1);
    mbar_init(&sm->done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(&sm->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;

  // ---- operands written by threads
  if (c.a_kind == 1) {
    // row `tid`: kdim bf16 values -> kdim/2 packed columns starting at column 256
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_raw + static_cast<size_t>(tid) * c.kdim);
    for (int ch = 0; ch < c.kdim / 2; ch += 16) {
      uint32_t r[16];
      for (int i = 0; i < 16; ++i) r[i] = src[ch + i];
      tmem_st16(tmem + 256 + ch + lane_addr, r);
    }
    tmem_st_wait();
  }
  if (c.a_kind == 2) {
    const uint4* src = reinterpret_cast<const uint4*>(a_raw + static_cast<size_t>(tid) * 64);
    for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(sa + tid * 128 + ((ch ^ (tid & 7)) << 4)) = src[ch];
    fence_proxy_async_smem();
  }
  if (c.b_kind == 2) {
    const uint4* src = reinterpret_cast<const uint4*>(b_raw + static_cast<size_t>(tid) * 64);
    for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(sb + tid * 128 + ((ch ^ (tid & 7)) << 4)) = src[ch];
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (c.reps > 1 && warp == 0) {
    // all operands were staged by threads / are garbage-tolerant: only timing matters here
    const bool leader = elect_one();
    const uint32_t sa_u = smem_u32(sa), sb_u = smem_u32(sb);
    const long long t0 = clock64();
    const int form = c.a_kind * 100 + c.b_kind * 10 + (c.n == 128 ? 1 : 0) + (c.kdim == 128 ? 2 : 0);
    switch (form) {
      case 3: rate_loop<0, 0, 128, 8>(tmem, sa_u, sb_u, c.reps, leader); break;    // QK
      case 2: rate_loop<0, 0, 64, 8>(tmem, sa_u, sb_u, c.reps, leader); break;     // S^T
      case 113: rate_loop<1, 1, 128, 8>(tmem, sa_u, sb_u, c.reps, leader); break;  // PV
      case 111: rate_loop<1, 1, 128, 4>(tmem, sa_u, sb_u, c.reps, leader); break;  // dV
      case 211: rate_loop<2, 1, 128, 4>(tmem, sa_u, sb_u, c.reps, leader); break;  // dK
      case 322: rate_loop<3, 2, 64, 8>(tmem, sa_u, sb_u, c.reps, leader); break;   // dQ^T
      default: break;
    }
    const long long t1 = clock64();
    if (leader) umma_commit(&sm->done);
    __syncwarp();
    mbar_wait(&sm->done, 0);
    if (leader && c.cycles != nullptr) {
      c.cycles[0] = static_cast<unsigned long long>(clock64() - t0);
      c.cycles[1] = static_cast<unsigned long long>(t1 - t0);
    }
  } else if (c.reps > 1) {
    // other warps just wait for the epilogue
  } else if (tid == 0) {
    uint32_t bytes = 0;
    if (c.a_kind == 0 || c.a_kind == 3) bytes += 32768;
    if (c.b_kind == 0) bytes += c.n * 256;
    if (c.b_kind == 1) bytes += c.kdim * 256;
    if (bytes) {
      mbar_arrive_expect_tx(&sm->full, bytes);
      if (c.a_kind == 0 || c.a_kind == 3) {
        tma_load_3d(sa, &tm_a, &sm->full, 0, 0, 0);
        tma_load_3d(sa + 16384, &tm_a, &sm->full, 64, 0, 0);
      }
      if (c.b_kind == 0 || c.b_kind == 1) {
        const int rows = c.b_kind == 0 ? c.n : c.kdim;
        tma_load_3d(sb, &tm_b, &sm->full, 0, 0, 0);
        tma_load_3d(sb + rows * 128, &tm_b, &sm->full, 64, 0, 0);
      }
      mbar_wait(&sm->full, 0);
    }
    tc_fence_after();
    const uint32_t a_mn = c.a_kind == 3, b_mn = c.b_kind != 0;
    const uint32_t idesc = umma_idesc_f16(1, 128, c.n, a_mn, b_mn);
    const uint32_t sa_u = smem_u32(sa), sb_u = smem_u32(sb);
    const int b_rows = c.b_kind == 0 ? c.n : c.kdim;  // rows of the B tile as stored
    if (c.reps > 1) {
      // throughput mode: descriptors precomputed, instructions issued back to back (what the kernels do)
      constexpr uint32_t hi = umma_desc_hi(1024, kSwizzle128B);
      uint32_t a_lo[8], b_lo[8];
      const int ksteps = c.kdim / 16;
      for (int k = 0; k < ksteps; ++k) {
        if (c.b_kind == 0) b_lo[k] = umma_desc_lo(sb_u + (k >> 2) * (b_rows * 128) + (k & 3) * 32, 16);
        else if (c.b_kind == 1) b_lo[k] = umma_desc_lo(sb_u + k * 2048, b_rows * 128);
        else b_lo[k] = umma_desc_lo(sb_u + k * 2048, 16);
        if (c.a_kind == 0) a_lo[k] = umma_desc_lo(sa_u + (k >> 2) * 16384 + (k & 3) * 32, 16);
        else if (c.a_kind == 2) a_lo[k] = umma_desc_lo(sa_u + k * 32, 16);
        else if (c.a_kind == 3) a_lo[k] = umma_desc_lo(sa_u + k * 2048, 16384);
        else a_lo[k] = tmem + 256 + k * 8;
      }
      const long long t0 = clock64();
      for (int rep = 0; rep < c.reps; ++rep) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k < ksteps) {
            if (c.a_kind == 1) umma_ts2(tmem, a_lo[k], b_lo[k], hi, idesc, 1u);
            else umma_ss2(tmem, a_lo[k], hi, b_lo[k], hi, idesc, 1u);
          }
        }
      }
      const long long t1 = clock64();
      umma_commit(&sm->done);
      mbar_wait(&sm->done, 0);
      if (c.cycles != nullptr) {
        c.cycles[0] = static_cast<unsigned long long>(clock64() - t0);
        c.cycles[1] = static_cast<unsigned long long>(t1 - t0);
      }
    } else {
    const long long t_start = clock64();
    for (int rep = 0; rep < (c.reps > 0 ? c.reps : 1); ++rep)
    for (int k = 0; k < c.kdim / 16; ++k) {
      uint64_t bd;
      if (c.b_kind == 0) {
        bd = umma_smem_desc(sb_u + (k >> 2) * (b_rows * 128) + (k & 3) * 32, c.lbo_b >= 0 ? c.lbo_b : 16,
                            c.sbo_b >= 0 ? c.sbo_b : 1024, kSwizzle128B);
      } else if (c.b_kind == 1) {
        bd = umma_smem_desc(sb_u + k * (c.kstep_b >= 0 ? c.kstep_b : 2048), c.lbo_b >= 0 ? c.lbo_b : b_rows * 128,
                            c.sbo_b >= 0 ? c.sbo_b : 1024, kSwizzle128B);
      } else {
        bd = umma_smem_desc(sb_u + k * (c.kstep_b >= 0 ? c.kstep_b : 2048), c.lbo_b >= 0 ? c.lbo_b : 16,
                            c.sbo_b >= 0 ? c.sbo_b : 1024, kSwizzle128B);
      }
      if (c.a_kind == 1) {
        umma_ts(tmem, tmem + 256 + k * 8, bd, idesc, (k > 0 || rep > 0) ? 1u : 0u);
      } else {
        uint64_t ad;
        if (c.a_kind == 0) {
          ad = umma_smem_desc(sa_u + (k >> 2) * 16384 + (k & 3) * 32, c.lbo_a >= 0 ? c.lbo_a : 16,
                              c.sbo_a >= 0 ? c.sbo_a : 1024, kSwizzle128B);
        } else if (c.a_kind == 2) {
          ad = umma_smem_desc(sa_u + k * 32, c.lbo_a >= 0 ? c.lbo_a : 16, c.sbo_a >= 0 ? c.sbo_a : 1024, kSwizzle128B);
        } else {
          ad = umma_smem_desc(sa_u + k * (c.kstep_a >= 0 ? c.kstep_a : 2048), c.lbo_a >= 0 ? c.lbo_a : 16384,
                              c.sbo_a >= 0 ? c.sbo_a : 1024, kSwizzle128B);
        }
        umma_ss(tmem, ad, bd, idesc, (k > 0 || rep > 0) ? 1u : 0u);
      }
    }
    const long long t_issued = clock64();
    umma_commit(&sm->done);
    mbar_wait(&sm->done, 0);
    if (c.cycles != nullptr) {
      c.cycles[0] = static_cast<unsigned long long>(clock64() - t_start);
      c.cycles[1] = static_cast<unsigned long long>(t_issued - t_start);
    }
    }
  }
  __syncwarp();
  mbar_wait(&sm->done, 0);
  tc_fence_after();
  for (int cc = 0; cc < c.n; cc += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + cc + lane_addr, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[static_cast<size_t>(tid) * c.n + cc + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace probe

const char* probe_launch(const TensorView& a, const TensorView& b, const void* a_raw, const void* b_raw, float* out,
                         const ProbeConfig& c, cudaStream_t stream) {
  CUtensorMap ta, tb;
  TensorView dummy = a.ptr ? a : b;
  if (const char* e = make_tensor_map(&ta, a.ptr ? a : dummy, 2, 128, 128)) return e;
  const int b_rows = c.b_kind == 0 ? c.n : c.kdim;
  if (const char* e = make_tensor_map(&tb, b.ptr ? b : dummy, 2, b_rows, 128)) return e;
  const int smem = 65536 + 1024 + 1024;
  cudaError_t err = cudaFuncSetAttribute(probe::probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (err != cudaSuccess) return cudaGetErrorString(err);
  probe::probe_kernel<<<1, 128, smem, stream>>>(ta, tb, static_cast<const __nv_bfloat16*>(a_raw),
                                                static_cast<const __nv_bfloat16*>(b_raw), out, c);
  err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

}  // namespace rfa
