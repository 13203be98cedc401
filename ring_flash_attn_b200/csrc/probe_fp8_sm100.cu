// Descriptor probe for kind::f8f6f4 in the two operand forms an fp8 attention forward needs:
//   form 0 (Q K^T):  D = A B^T, A = smem K-major [128 rows][128 k] e4m3, B = smem K-major [128 n][128 k] e4m3,
//                    both TMA tiles with one 128-byte swizzle span per row; 4 instructions of K = 32.
//   form 1 (P V):    D = A B,   A = TMEM e4m3 [128 rows][128 k] (four elements per 32-bit column, written by
//                    threads), B = smem MN-major from a [128 k][128 n] TMA tile; 4 instructions of K = 32.
// LBO / SBO / k-step of the B descriptor and the byte order inside a packed TMEM word can be overridden at run
// time so that one GPU session can sweep the alternatives (benchmark/probe_descriptors.py --fp8).
#include <stdio.h>

#include "attn_common.h"
#include "sm100_ptx.cuh"

namespace rfa {
namespace probe8 {

struct Smem {
  uint64_t full;
  uint64_t done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1)
probe_fp8_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const uint8_t* __restrict__ a_raw, float* __restrict__ out, const ProbeConfig c) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;          // 16 KB
  uint8_t* sb = smem + 16384;  // 16 KB
  Smem* sm = reinterpret_cast<Smem*>(smem + 32768);
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

  if (c.a_kind == 1) {
    // row `tid`: 128 e4m3 bytes -> 32 packed columns starting at column 256.  c.lbo_a selects the byte order
    // inside a word: 0 (default) = element 4c in the least significant byte, 1 = reversed.
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_raw + static_cast<size_t>(tid) * 128);
    for (int ch = 0; ch < 32; ch += 8) {
      uint32_t r[8];
      for (int i = 0; i < 8; ++i) {
        uint32_t w = src[ch + i];
        if (c.lbo_a == 1) w = __byte_perm(w, 0, 0x0123);
        r[i] = w;
      }
      tmem_st8(tmem + 256 + ch + lane_addr, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    uint32_t bytes = 16384;
    if (c.a_kind == 0) bytes += 16384;
    mbar_arrive_expect_tx(&sm->full, bytes);
    if (c.a_kind == 0) tma_load_3d(sa, &tm_a, &sm->full, 0, 0, 0);
    tma_load_3d(sb, &tm_b, &sm->full, 0, 0, 0);
    mbar_wait(&sm->full, 0);
    tc_fence_after();
    const uint32_t sa_u = smem_u32(sa), sb_u = smem_u32(sb);
    const uint32_t idesc = umma_idesc_f8(0, 0, 128, 128, 0, c.b_kind == 1);
    const uint32_t hi_a = umma_desc_hi(1024, kSwizzle128B);
    const uint32_t hi_b = umma_desc_hi(c.sbo_b >= 0 ? c.sbo_b : 1024, kSwizzle128B);
    for (int k = 0; k < 4; ++k) {
      // K-major: 32 one-byte elements = 32 bytes inside the 128-byte row; MN-major: 32 key rows of 128 bytes
      const uint32_t b_step = c.kstep_b >= 0 ? c.kstep_b : (c.b_kind == 0 ? 32 : 4096);
      const uint32_t b_lo = umma_desc_lo(sb_u + k * b_step, c.lbo_b >= 0 ? c.lbo_b : 16);
      if (c.a_kind == 1) {
        umma_ts2_f8(tmem, tmem + 256 + k * 8, b_lo, hi_b, idesc, k > 0);
      } else {
        umma_ss2_f8(tmem, umma_desc_lo(sa_u + k * 32, 16), hi_a, b_lo, hi_b, idesc, k > 0);
      }
    }
    umma_commit(&sm->done);
  }
  __syncwarp();
  mbar_wait(&sm->done, 0);
  tc_fence_after();
  for (int cc = 0; cc < 128; cc += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + cc + lane_addr, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[static_cast<size_t>(tid) * 128 + cc + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace probe8

// a, b: (128, 1, 128) one-byte tensors.  c.a_kind: 0 = smem K-major, 1 = TMEM; c.b_kind: 0 = K-major, 1 = MN-major.
const char* probe_fp8_launch(const TensorView& a, const TensorView& b, float* out, const ProbeConfig& c,
                             cudaStream_t stream) {
  CUtensorMap ta, tb;
  if (const char* e = make_tensor_map(&ta, a, 1, 128, 128)) return e;
  if (const char* e = make_tensor_map(&tb, b, 1, 128, 128)) return e;
  const int smem = 32768 + 1024 + 1024;
  cudaError_t err = cudaFuncSetAttribute(probe8::probe_fp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (err != cudaSuccess) return cudaGetErrorString(err);
  probe8::probe_fp8_kernel<<<1, 128, smem, stream>>>(ta, tb, static_cast<const uint8_t*>(a.ptr), out, c);
  err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

}  // namespace rfa
