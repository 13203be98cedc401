// Blockwise flash-attention backward for sm_100a (head_dim 128, bf16/fp16).
//
// One CTA = one 128-key tile of one kv head.  dK and dV of that tile stay in tensor memory while the CTA
// sweeps every query tile (64 rows) of every query head of the GQA group and every query segment that can
// see the keys; dQ tiles are produced transposed, staged through shared memory and added into an fp32
// accumulator with TMA reduce-add.  This replaces flash_attn's backward + the reference's fp32
// dq/dk/dv bookkeeping per ring step (/root/reference/ring_flash_attn/ring_flash_attn.py:97-152).
//
// Per query tile i (all GEMMs on tcgen05, accumulators in TMEM):
//   S^T  = K  Q_i^T        (128 keys x 64 queries)        SS, both K-major
//   dP^T = V  dO_i^T                                      SS, both K-major
//   P^T  = exp2(S^T * c - lse_i)          -> TMEM (bf16), dS^T = P^T o (dP^T - delta_i) * scale -> smem
//   dV  += P^T  dO_i       (A from TMEM, B = dO MN-major)
//   dK  += dS^T Q_i        (A = dS^T smem K-major, B = Q MN-major)
//   dQ_i^T = K^T dS_i      (A = K MN-major, B = dS^T MN-major)  -> fp32 smem -> TMA reduce-add
//
// Warp roles (512 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warp 3 lse/delta
// prefetch, warps 4-7 and 8-11 softmax (thread == key row; each warpgroup owns 32 of the 64 query columns, so
// two warps per scheduler hide each other's latencies), warps 12-15 dQ drain (thread == head-dim lane).
// TMEM (512 columns): S^T x2 [0,128) (P^T halves parked at +0 and +32 of a buffer)  dP^T [128,192)  dQ^T [192,256)  dV [256,384)  dK [384,512).
#include <math_constants.h>
#include <stdio.h>

#include "attn_common.h"
#include "comm_device.cuh"
#include "sm100_ptx.cuh"

namespace rfa {
namespace bwd {

#ifdef RFA_TRACE
// per-tile stamps of CTA 0 (16 slots): 0-5 MMA warp, 6-10 softmax half 0, 11-12 drain
#define RFA_STAMP(cond, iter, slot)                                                              \
  do {                                                                                            \
    if ((cond) && cta == 0 && p.trace != nullptr && (iter) < 64) p.trace[(iter)*16 + (slot)] = clock64(); \
  } while (0)
#else
#define RFA_STAMP(cond, iter, slot) \
  do {                              \
  } while (0)
#endif

constexpr int kTileK = 128;  // keys per CTA
constexpr int kTileQ = 64;   // queries per inner iteration
constexpr int kStages = 3;   // Q/dO ring
constexpr int kThreads = 512;
constexpr int kQHalf = kTileQ * 128;               // 64-wide swizzled sub-tile of a 64-row tile (8 KB)
constexpr int kKVHalf = kTileK * 128;              // 64-wide swizzled sub-tile of a 128-row tile (16 KB)
constexpr int kDSBytes = kTileK * kTileQ * 2;      // 16 KB
// Head dim kD (64 or 128) is a template parameter: tiles are kD / 64 sub-tiles wide, dK / dV take kD columns.
template <int kD>
struct Geo {
  static constexpr int kKVBytes = kTileK * kD * 2;  // K, V tile: 32 KB at kD = 128
  static constexpr int kQBytes = kTileQ * kD * 2;   // Q, dO tile: 16 KB
  static constexpr int kDQBytes = kTileQ * kD * 4;  // fp32 dQ staging: 32 KB
  static constexpr int kSmemBytes = 2 * kKVBytes + kStages * 2 * kQBytes + kDSBytes + kDQBytes +
                                    kStages * 2 * kTileQ * 4 /*stats*/ + 1024 /*barriers*/ + 1024 /*slack*/;
};
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColDP = 128, kColDQ = 192, kColDV = 256, kColDK = 384;

struct Barriers {
  uint64_t kv_full;
  uint64_t qdo_full[kStages];
  uint64_t qdo_empty[kStages];
  uint64_t stat_full[kStages];
  uint64_t s_full[2];
  uint64_t p_ready[2];  // one per S^T buffer: with S^T look-ahead a fast warp may reach tile i+1 before a slow
                        // warp has arrived for tile i, and must not be counted in tile i's phase
  uint64_t dp_full;
  uint64_t ds_ready;
  uint64_t dq_full;
  uint64_t dq_free;
  uint64_t dkv_done;
  uint32_t tmem_base;
  uint32_t pad;
};

constexpr int kStatBytes = kStages * 2 * kTileQ * 4;  // lse / delta ring, same slots as Q/dO

// Query tiles of one segment that can see this key tile.
struct QGeom {
  int q_row0, q_len, diag, lo;
  int t_begin, t_end;  // 64-row tile range
};
__device__ __forceinline__ QGeom q_geom(const BwdQSegment& s) {
  QGeom g;
  g.q_row0 = s.q_row0;
  g.q_len = s.q_len;
  g.diag = s.diag;
  g.lo = s.lo;
  const int first = s.diag >= 0 ? 0 : -s.diag;  // first chunk row that sees key 0 of the tile
  g.t_begin = first / kTileQ;
  g.t_end = (s.q_len + kTileQ - 1) / kTileQ;
  if (g.t_begin > g.t_end) g.t_begin = g.t_end;
  return g;
}

template <typename T, bool kWindow, int kD>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                const __grid_constant__ CUtensorMap tm_ks, const __grid_constant__ CUtensorMap tm_vs,
                const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ BwdParams p) {
  // communication CTAs first (see comm_device.cuh): they re-publish this rank's K/V rows to the peers
  if (static_cast<int>(blockIdx.x) < p.push.n_ctas) {
    if (p.push.use_tma) {
      extern __shared__ uint8_t push_smem_raw[];
      push_role_tma(p.push, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(push_smem_raw) + 1023) & ~uintptr_t(1023)));
    } else {
      push_role(p.push);
    }
    return;
  }
  static_assert(kD == 64 || kD == 128, "head dim 64 or 128");
  constexpr int kKVBytes = Geo<kD>::kKVBytes, kQBytes = Geo<kD>::kQBytes, kDQBytes = Geo<kD>::kDQBytes;
  constexpr int kSubTiles = kD / 64;
  const int cta = static_cast<int>(blockIdx.x) - p.push.n_ctas;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_k = smem;
  uint8_t* smem_v = smem_k + kKVBytes;
  uint8_t* smem_qdo = smem_v + kKVBytes;                 // kStages x (Q 16 KB, dO 16 KB)
  uint8_t* smem_ds = smem_qdo + kStages * 2 * kQBytes;   // 16 KB, [128 keys][64 queries] swizzled
  float* smem_dq = reinterpret_cast<float*>(smem_ds + kDSBytes);   // [64 queries][128 dims] fp32
  float* smem_stat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem_dq) + kDQBytes);  // [2][2][64]
  Barriers* bars = reinterpret_cast<Barriers*>(reinterpret_cast<uint8_t*>(smem_stat) + kStatBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA -> (key tile, kv head).  One GPU: head-major, so that concurrently running CTAs re-read the Q / dO of ONE
  // head out of L2 (at S = 32K all heads together do not fit).  Fused multi-GPU launches: tile-major - the table
  // is sorted by ring step (local keys first, then the sources in the order their K/V arrives), and with head-major
  // numbering the tiles of the LAST source of head 0 would be scheduled, and sit waiting for their data, before any
  // tile of head 1 that could already run.
  const int kv_head = p.item_major ? cta % p.hkv : cta / p.n_items;
  const int group = p.hq / p.hkv;
  BwdItem it = p.items[p.item_major ? cta / p.hkv : cta % p.n_items];
  if (p.flags & 4) it.seg_count = 0;  // timing experiment: K/V push + dK/dV return only

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_dq);
    tma_prefetch_desc(&tm_ks);
    tma_prefetch_desc(&tm_vs);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bars->kv_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars->qdo_full[i], 1);
      mbar_init(&bars->qdo_empty[i], 1);
      mbar_init(&bars->stat_full[i], 1);
    }
    mbar_init(&bars->s_full[0], 1);
    mbar_init(&bars->s_full[1], 1);
    mbar_init(&bars->p_ready[0], 256);
    mbar_init(&bars->p_ready[1], 256);
    mbar_init(&bars->dp_full, 1);
    mbar_init(&bars->ds_ready, 256);
    mbar_init(&bars->dq_full, 1);
    mbar_init(&bars->dq_free, 128);
    mbar_init(&bars->dkv_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // total number of query tiles this CTA sweeps (identical in every role)
  int tiles_per_head = 0;
  for (int si = 0; si < it.seg_count; ++si) {
    const QGeom g = q_geom(p.qsegs[it.seg_begin + si]);
    tiles_per_head += g.t_end - g.t_begin;
  }
  const int total_tiles = tiles_per_head * group;

  if (warp < 4) {
    reg_dealloc<72>();  // 128*72 + 384*144 = 512*126 <= 512*128
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      // (a tile no local query reaches only sends zero rows: nothing is loaded, and in particular no TMA write can
      // land in the K | V area after the epilogue has started to stage those rows there)
      if (lane == 0 && total_tiles > 0) {
        const bool staged = it.flag >= 0 && p.ready_flags != nullptr;
        if (staged && !(p.flags & 2)) {
          wait_epoch(p.ready_flags + it.flag, p.ready_epoch, "bwd kv ready", p.sig.my_rank, it.flag);
          fence_proxy_async_all();
        }
        const CUtensorMap* mk = staged ? &tm_ks : &tm_k;
        const CUtensorMap* mv = staged ? &tm_vs : &tm_v;
        mbar_arrive_expect_tx(&bars->kv_full, 2 * kKVBytes);
#pragma unroll
        for (int h = 0; h < kSubTiles; ++h) {
          tma_load_3d(smem_k + h * kKVHalf, mk, &bars->kv_full, 64 * h, kv_head, it.kv_row0);
          tma_load_3d(smem_v + h * kKVHalf, mv, &bars->kv_full, 64 * h, kv_head, it.kv_row0);
        }
        uint32_t slot = 0, phase = 0;
        for (int gq = 0; gq < group; ++gq) {
          const int head = kv_head * group + gq;
          for (int si = 0; si < it.seg_count; ++si) {
            const QGeom g = q_geom(p.qsegs[it.seg_begin + si]);
            for (int ti = g.t_begin; ti < g.t_end; ++ti) {
              const int row = g.q_row0 + ti * kTileQ;
              mbar_wait(&bars->qdo_empty[slot], phase ^ 1);
              uint8_t* dq_ = smem_qdo + slot * 2 * kQBytes;
              uint8_t* ddo = dq_ + kQBytes;
              mbar_arrive_expect_tx(&bars->qdo_full[slot], 2 * kQBytes);
#pragma unroll
              for (int h = 0; h < kSubTiles; ++h) {
                tma_load_3d(dq_ + h * kQHalf, &tm_q, &bars->qdo_full[slot], 64 * h, head, row);
                tma_load_3d(ddo + h * kQHalf, &tm_do, &bars->qdo_full[slot], 64 * h, head, row);
              }
              if (++slot == kStages) {
                slot = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      // Warp-uniform loop (descriptors live in uniform registers); one elected lane issues and commits.
      if (total_tiles > 0) {
        const bool leader = elect_one();
        constexpr uint32_t fmt = Pack2<T>::kFmt;
        constexpr uint32_t idesc_st = umma_idesc_f16(fmt, kTileK, kTileQ, 0, 0);  // S^T, dP^T
        constexpr uint32_t idesc_dv = umma_idesc_f16(fmt, kTileK, kD, 0, 1);      // dV (A tmem), dK (A smem)
        // dQ^T has M = head dims.  For kD = 64 the instruction still runs with M = 128: the upper 64 "dims" of the
        // MN-major A operand fall into whatever follows the K tile in shared memory (the V tile) and produce
        // garbage accumulator LANES 64..127, which the drain warps never read - rows of D are independent, and
        // it avoids the different tensor-memory layout of M = 64 accumulators.
        constexpr uint32_t idesc_dq = umma_idesc_f16(fmt, 128, kTileQ, 1, 1);     // dQ^T
        const uint32_t k_base = smem_u32(smem_k), v_base = smem_u32(smem_v);
        const uint32_t qdo_base = smem_u32(smem_qdo), ds_base = smem_u32(smem_ds);

        // Descriptors as (lo, hi) halves: hi is a constant, lo = (base >> 4) + compile-time k step.
        constexpr uint32_t hi = umma_desc_hi(1024, kSwizzle128B);
        const uint32_t k_km = umma_desc_lo(k_base, 16);        // K tile, K-major A (S^T)
        const uint32_t v_km = umma_desc_lo(v_base, 16);        // V tile, K-major A (dP^T)
        const uint32_t k_mn = umma_desc_lo(k_base, kKVHalf);   // K tile, MN-major A (dQ^T)
        const uint32_t ds_km = umma_desc_lo(ds_base, 16);      // dS^T, K-major A (dK) / MN-major B (dQ^T)
        const uint32_t q_km0 = umma_desc_lo(qdo_base, 16);     // stage 0 Q, K-major B
        const uint32_t q_mn0 = umma_desc_lo(qdo_base, kQHalf); // stage 0 Q, MN-major B
        constexpr uint32_t stage_step = (2 * kQBytes) >> 4, do_step = kQBytes >> 4;

        // S^T / dP^T: A = K or V tile (128 rows, K-major), B = Q or dO tile (64 rows, K-major)
        auto issue_kq = [&](uint32_t d_col, uint32_t a_lo, uint32_t b_lo) {
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            const uint32_t oa = ((k >> 2) * kKVHalf + (k & 3) * 32) >> 4;
            const uint32_t ob = ((k >> 2) * kQHalf + (k & 3) * 32) >> 4;
            umma_ss2(d_col, a_lo + oa, hi, b_lo + ob, hi, idesc_st, k > 0);
          }
        };

        mbar_wait(&bars->kv_full, 0);
        tc_fence_after();

        // Issue order (round 2).  The longest dependency chain of a tile runs through the softmax warpgroups:
        //   dP^T(i) complete -> dS^T(i) written -> [dK(i), dQ^T(i)] ... -> dP^T(i+1) complete -> dS^T(i+1) ...
        // (dP^T has ONE tensor-memory buffer, so dP^T(i+1) can only be issued once dS^T(i) exists).  Round 1 issued
        // dP^T(i+1) LAST in iteration i, behind S^T(i+1), dV(i), dK(i) and dQ^T(i): 3660 cycles per tile for 1664
        // cycles of tensor work (profiles/trace_bwd_cta0.log).  Now dP^T(i+1) is the FIRST thing issued when dS^T(i)
        // arrives, and everything that does not depend on the softmax of tile i+1 (dK(i), dQ^T(i), S^T(i+2),
        // dV(i+1)) fills the pipe behind it while the softmax warps turn dP^T(i+1) into dS^T(i+1).
        // S^T runs two tiles ahead (three Q/dO stages are in flight: i, i+1, i+2).
        uint32_t ph_p[2] = {0, 0};
        uint32_t ph_ds = 0, ph_dqfree = 0;
        int issued_s = 0;                  // tiles whose S^T has been issued
        uint32_t s_slot = 0, s_phase = 0;  // ring position used by the next S^T issue
        auto issue_s = [&]() {
          mbar_wait(&bars->qdo_full[s_slot], s_phase);
          tc_fence_after();
          if (leader) {
            issue_kq(tmem + kColS + (issued_s & 1) * 64, k_km, q_km0 + s_slot * stage_step);
            umma_commit(&bars->s_full[issued_s & 1]);
          }
          ++issued_s;
          if (++s_slot == kStages) {
            s_slot = 0;
            s_phase ^= 1;
          }
        };
        // dV(t) += P^T(t) dO(t)   (P^T: bf16 in the first 32 columns of tile t's S^T buffer)
        auto issue_dv = [&](int t, uint32_t t_slot) {
          mbar_wait(&bars->p_ready[t & 1], ph_p[t & 1]);
          ph_p[t & 1] ^= 1;
          tc_fence_after();
          const uint32_t do_mn = q_mn0 + t_slot * stage_step + do_step;  // dO_t read MN-major
          if (leader)
#pragma unroll
            for (int k = 0; k < kTileQ / 16; ++k)
              umma_ts2(tmem + kColDV, tmem + kColS + (t & 1) * 64 + (k >> 1) * 32 + (k & 1) * 8,
                       do_mn + k * (2048 >> 4), hi, idesc_dv, (t > 0 || k > 0) ? 1u : 0u);
        };
        auto next_slot = [](uint32_t sl) { return sl + 1 == kStages ? 0u : sl + 1; };

        issue_s();  // S^T(0)
        if (leader) {  // dP^T(0): its Q/dO stage has landed (S^T(0) waited for it)
          issue_kq(tmem + kColDP, v_km, q_km0 + do_step);
          umma_commit(&bars->dp_full);
        }
        if (total_tiles > 1) issue_s();  // S^T(1)
        issue_dv(0, 0);
        uint32_t slot = 0;  // stage of tile i
        for (int i = 0; i < total_tiles; ++i) {
          const uint32_t q_mn = q_mn0 + slot * stage_step;  // Q_i read MN-major (B of dK)
          const uint32_t slot1 = next_slot(slot);
          RFA_STAMP(leader, i, 0);
          mbar_wait(&bars->ds_ready, ph_ds);  // dS^T(i) is in shared memory; dP^T(i) has been consumed
          ph_ds ^= 1;
          tc_fence_after();
          RFA_STAMP(leader, i, 1);
          if (i + 1 < total_tiles && leader) {  // dP^T(i+1) first: it heads the critical chain
            issue_kq(tmem + kColDP, v_km, q_km0 + slot1 * stage_step + do_step);
            umma_commit(&bars->dp_full);
          }
          RFA_STAMP(leader, i, 2);
          if (leader) {  // dK += dS^T Q
#pragma unroll
            for (int k = 0; k < kTileQ / 16; ++k)
              umma_ss2(tmem + kColDK, ds_km + k * (32 >> 4), hi, q_mn + k * (2048 >> 4), hi, idesc_dv,
                       (i > 0 || k > 0) ? 1u : 0u);
          }
          if (i > 0) {  // previous dQ^T must have been drained out of TMEM
            mbar_wait(&bars->dq_free, ph_dqfree);
            ph_dqfree ^= 1;
            tc_fence_after();
          }
          RFA_STAMP(leader, i, 3);
          if (leader) {  // dQ^T = K^T dS
#pragma unroll
            for (int k = 0; k < kTileK / 16; ++k)
              umma_ss2(tmem + kColDQ, k_mn + k * (2048 >> 4), hi, ds_km + k * (2048 >> 4), hi, idesc_dq, k > 0);
            umma_commit(&bars->dq_full);            // also tells the softmax warps that dS^T(i) may be overwritten
            umma_commit(&bars->qdo_empty[slot]);    // Q_i / dO_i no longer needed once everything above retires
          }
          RFA_STAMP(leader, i, 4);
          if (i + 1 < total_tiles) issue_dv(i + 1, slot1);  // dV(i+1): P^T(i+1) has been ready for a while
          // S^T(i+2) last: its stage was released by tile i-1 one iteration ago, so this is the only wait of the
          // loop that may have to sit out a TMA latency (its buffer held P^T(i), and dV(i) has been issued)
          if (i + 2 < total_tiles) issue_s();
          RFA_STAMP(leader, i, 5);
          slot = slot1;
          __syncwarp();
        }
        if (leader) umma_commit(&bars->dkv_done);
        __syncwarp();
      }
    } else if (warp == 3) {
      // ---------------------------------------------------------------- per-query statistics producer
      // lse (converted to log2 units; +inf for rows that do not exist or saw no key => P = 0) and delta.
      uint32_t slot = 0, phase = 0;
      for (int gq = 0; gq < group; ++gq) {
        const int head = kv_head * group + gq;
        for (int si = 0; si < it.seg_count; ++si) {
          const QGeom g = q_geom(p.qsegs[it.seg_begin + si]);
          for (int ti = g.t_begin; ti < g.t_end; ++ti) {
            float l2[2], dl[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int qi = ti * kTileQ + lane + 32 * h;
              l2[h] = CUDART_INF_F;
              dl[h] = 0.f;
              if (qi < g.q_len) {
                const int row = g.q_row0 + qi;
                const size_t b = row / p.lse_S, sidx = row % p.lse_S;
                const size_t idx = (b * p.hq + head) * static_cast<size_t>(p.lse_S) + sidx;
                const float x = p.lse[idx];
                l2[h] = x == -CUDART_INF_F ? CUDART_INF_F : x * 1.4426950408889634f;
                dl[h] = p.delta[idx];
              }
            }
            mbar_wait(&bars->qdo_empty[slot], phase ^ 1);
            float* st = smem_stat + slot * 2 * kTileQ;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              st[lane + 32 * h] = l2[h];
              st[kTileQ + lane + 32 * h] = dl[h];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->stat_full[slot]);
            if (++slot == kStages) {
              slot = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ softmax warpgroups (thread == key row)
    reg_alloc<144>();
    const int half = (warp - 4) >> 2;             // which 32 query columns of every tile this warpgroup owns
    const int c0 = half * 32;
    const int wg_tid = (threadIdx.x - 128) & 127;
    const int key = wg_tid;                       // row in the key tile
    const bool key_ok = key < it.kv_rows;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t ph_s[2] = {0, 0};
    uint32_t ph_dp = 0;
    uint32_t st_slot = 0, st_phase = 0;
    int i = 0;
    for (int gq = 0; gq < group; ++gq) {
      for (int si = 0; si < it.seg_count; ++si) {
        const QGeom g = q_geom(p.qsegs[it.seg_begin + si]);
        for (int ti = g.t_begin; ti < g.t_end; ++ti, ++i) {
          const float* st = smem_stat + st_slot * 2 * kTileQ + c0;
          mbar_wait(&bars->stat_full[st_slot], st_phase);
          if (++st_slot == kStages) {
            st_slot = 0;
            st_phase ^= 1;
          }
          float lse2[32];
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(st + c);
            lse2[c] = v4.x; lse2[c + 1] = v4.y; lse2[c + 2] = v4.z; lse2[c + 3] = v4.w;
          }

          const int buf = i & 1;
          const uint32_t t_s = tmem + kColS + buf * 64 + lane_addr;
          mbar_wait(&bars->s_full[buf], ph_s[buf]);
          ph_s[buf] ^= 1;
          tc_fence_after();
          RFA_STAMP(threadIdx.x == 128, i, 6);
          uint32_t sr[32];
          tmem_ld32(t_s + c0, sr);
          tmem_ld_wait();
          // causal boundary: key visible to query qi iff key <= qi + diag  <=>  qi >= key - diag
          const long long first_q = static_cast<long long>(key) - g.diag - static_cast<long long>(ti) * kTileQ;
          const int q_lo = !key_ok ? kTileQ : (first_q < 0 ? 0 : (first_q > kTileQ ? kTileQ : static_cast<int>(first_q)));
          float pr[32];
          if constexpr (kWindow) {
            // sliding window: ... and iff key >= qi + lo  <=>  qi <= key - lo (the host already dropped the query
            // tiles past the last such row)
            const long long last_q = static_cast<long long>(key) - g.lo - static_cast<long long>(ti) * kTileQ;
            const int q_hi = last_q < 0 ? 0 : (last_q >= kTileQ ? kTileQ : static_cast<int>(last_q) + 1);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float e = fast_exp2(fmaf(__uint_as_float(sr[c]), p.scale_log2, -lse2[c]));
              pr[c] = ((c0 + c) >= q_lo && (c0 + c) < q_hi) ? e : 0.f;
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float e = fast_exp2(fmaf(__uint_as_float(sr[c]), p.scale_log2, -lse2[c]));
              pr[c] = (c0 + c) >= q_lo ? e : 0.f;
            }
          }
          {
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) pk[c] = Pack2<T>::pack(pr[2 * c], pr[2 * c + 1]);
            // each warpgroup parks its half of P^T inside the S^T columns it has already consumed itself
            // (half 0 -> columns [0,16), half 1 -> [32,48)); writing into the other group's columns would race
            tmem_st16(t_s + c0, pk);
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&bars->p_ready[buf]);
          RFA_STAMP(threadIdx.x == 128, i, 7);

          // dS^T = P^T o (dP^T - delta) * scale  -> shared memory, 128-byte swizzled rows of 64 bf16
          float dlt[32];
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(st + kTileQ + c);
            dlt[c] = v4.x; dlt[c + 1] = v4.y; dlt[c + 2] = v4.z; dlt[c + 3] = v4.w;
          }

          mbar_wait(&bars->dp_full, ph_dp);
          ph_dp ^= 1;
          tc_fence_after();
          RFA_STAMP(threadIdx.x == 128, i, 8);
          uint32_t dpr[32];
          tmem_ld32(tmem + kColDP + lane_addr + c0, dpr);
          tmem_ld_wait();
          uint8_t* ds_row = smem_ds + key * 128;
          uint32_t ds_pk[16];
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float d0 = pr[c] * (__uint_as_float(dpr[c]) - dlt[c]) * p.scale;
            const float d1 = pr[c + 1] * (__uint_as_float(dpr[c + 1]) - dlt[c + 1]) * p.scale;
            ds_pk[c >> 1] = Pack2<T>::pack(d0, d1);
          }
          // dS^T has ONE shared-memory buffer and dP^T(i) is now issued ahead of dK(i-1) / dQ^T(i-1): those two
          // GEMMs must have finished reading dS^T(i-1) before it is overwritten (dq_full = "dQ^T(i-1) complete",
          // which by in-order completion covers dK(i-1) as well)
          if (i > 0) mbar_wait(&bars->dq_full, (i - 1) & 1);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            *reinterpret_cast<uint4*>(ds_row + (((half * 4 + ch) ^ (key & 7)) << 4)) =
                make_uint4(ds_pk[ch * 4], ds_pk[ch * 4 + 1], ds_pk[ch * 4 + 2], ds_pk[ch * 4 + 3]);
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&bars->ds_ready);
          RFA_STAMP(threadIdx.x == 128, i, 9);
        }
      }
    }
    // ---------------------------------------------------------------- epilogue: dK (half 0) / dV (half 1) -> global
    const bool remote = p.dkv.world > 0;
    const int row = (remote ? it.out_row0 : it.kv_row0) + key;
    if (remote && wg_tid == 0) {
      // the owner must have drained what we stored into its inbox during the previous backward call
      wait_epoch(p.dkv.my_pad + kPadInboxFree + it.owner, p.dkv.wait_epoch, "inbox reuse", p.dkv.my_rank, it.owner);
    }
    if (remote) named_bar_sync(1 + half, 128);
    if (total_tiles > 0) {
      mbar_wait(&bars->dkv_done, 0);
      tc_fence_after();
    }
    // Every thread (== key row) parks its row in shared memory that is idle by now (all GEMMs of this CTA have
    // retired and every TMA load has landed: dK uses K | V plus the head of the Q/dO ring, dV the rest of the ring)
    // and ships it itself with ONE bulk copy - no cross-thread hand-off is needed because a thread only sends
    // what it wrote.  Rows are pitched 16 bytes past their length so that the 16-byte stores of a warp fall into 8
    // distinct bank groups.  The row is written in the model dtype (fused multi-GPU launches and world size 1:
    // half the NVLink / HBM bytes of round 1's fp32 rows, and no cast pass afterwards; the owner-side reduction
    // accumulates the partials of different ranks in fp32) or in fp32 (torch.distributed fallback transports,
    // which add the partials of successive ring steps).  Round 1 sent one 16-byte store per lane and row:
    // measured 1.91 -> 1.64 ms per fwd+bwd on 2 GPUs for the headline shard (profiles/r2/breakdown_bulk*.log).
    auto store_rows = [&](auto tag) {
      using O = decltype(tag);
      constexpr int kRowBytes = kD * static_cast<int>(sizeof(O));
      constexpr int kKVBytes = Geo<kD>::kKVBytes, kQBytes = Geo<kD>::kQBytes;
      constexpr int kPitch = kRowBytes + 16;
      static_assert(2 * kTileK * kPitch <= 2 * kKVBytes + kStages * 2 * kQBytes, "staging must fit in K|V|Q/dO");
      uint8_t* mine = smem_k + half * (kTileK * kPitch) + key * kPitch;
      const uint32_t col = tmem + (half == 0 ? kColDK : kColDV) + lane_addr;
#pragma unroll
      for (int c = 0; c < kD; c += 32) {
        uint32_t r[32];
        if (total_tiles > 0) {
          tmem_ld32(col + c, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = 0;  // tile no local query reaches: the owner still needs defined rows
        }
        if constexpr (sizeof(O) == 4) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<uint4*>(mine + (c + e) * 4) = make_uint4(r[e], r[e + 1], r[e + 2], r[e + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 v;
            v.x = Pack2<T>::pack(__uint_as_float(r[e + 0]), __uint_as_float(r[e + 1]));
            v.y = Pack2<T>::pack(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
            v.z = Pack2<T>::pack(__uint_as_float(r[e + 4]), __uint_as_float(r[e + 5]));
            v.w = Pack2<T>::pack(__uint_as_float(r[e + 6]), __uint_as_float(r[e + 7]));
            *reinterpret_cast<uint4*>(mine + (c + e) * 2) = v;
          }
        }
      }
      fence_proxy_async_smem();
      if (key_ok) {
        uint8_t* base = static_cast<uint8_t*>(remote ? (half == 0 ? p.dkv.dk_ptrs[it.owner] : p.dkv.dv_ptrs[it.owner])
                                                     : (half == 0 ? p.dk : p.dv));
        bulk_store(base + (static_cast<size_t>(row) * p.hkv + kv_head) * kRowBytes, mine, kRowBytes);
        tma_store_commit();
        tma_store_wait<0>();  // written, not just read: the flag below must not overtake the data
        fence_proxy_async_all();
      }
    };
    if (p.dkv_fp32) store_rows(float{});
    else store_rows(T{});
    if (remote) {
      // publish: the last tile destined for an owner raises that owner's "gradients landed" flag
      __threadfence_system();
      named_bar_sync(3, 256);  // both halves (dK and dV) have been stored and fenced
      if (half == 0 && wg_tid == 0) {
        const uint32_t old = atomicAdd(p.dkv.sent_count + it.owner, 1u);
        if (old + 1u == p.dkv.sent_target[it.owner]) {
          __threadfence_system();
          st_release_sys(p.dkv.peer_pads[it.owner] + kPadDkvReady + p.dkv.my_rank, p.dkv.epoch);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ dQ drain warpgroup (thread == dim)
    reg_alloc<144>();
    const int wg_tid = threadIdx.x - 384;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t ph = 0;
    int i = 0;
    for (int gq = 0; gq < group; ++gq) {
      const int head = kv_head * group + gq;
      for (int si = 0; si < it.seg_count; ++si) {
        const QGeom g = q_geom(p.qsegs[it.seg_begin + si]);
        for (int ti = g.t_begin; ti < g.t_end; ++ti, ++i) {
          mbar_wait(&bars->dq_full, ph);
          ph ^= 1;
          tc_fence_after();
          RFA_STAMP(wg_tid == 0, i, 11);
          uint32_t r[64];
          tmem_ld32(tmem + kColDQ + lane_addr, r);
          tmem_ld32(tmem + kColDQ + lane_addr + 32, r + 32);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&bars->dq_free);
          // the previous tile's reduce must have finished reading the staging buffer
          if (wg_tid == 0) tma_store_wait_read<0>();
          named_bar_sync(4, 128);
          if (wg_tid < kD) {  // (kD = 64: lanes 64..127 of the M = 128 accumulator are not dQ)
#pragma unroll
            for (int q = 0; q < kTileQ; ++q) smem_dq[q * kD + wg_tid] = __uint_as_float(r[q]);
          }
          fence_proxy_async_smem();
          named_bar_sync(4, 128);
          if (wg_tid == 0) {
            tma_reduce_add_3d(&tm_dq, smem_dq, 0, head, g.q_row0 + ti * kTileQ);
            tma_store_commit();
          }
          RFA_STAMP(wg_tid == 0, i, 12);
        }
      }
    }
    if (wg_tid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem);
  if (threadIdx.x == 0 && p.sig.world > 0) consumer_done(p.sig);
}

// delta[h, t] = sum_d out[t,h,d] * dout[t,h,d]; one warp per (row, head).
template <typename T>
__global__ void bwd_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ delta,
                                 int rows, int hq, int lse_S, int64_t o_rs, int64_t o_hs, int64_t do_rs,
                                 int64_t do_hs, int head_dim) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= rows * hq) return;
  const int row = gw / hq, head = gw % hq;
  float s = 0.f;
  if (lane * 4 < head_dim) {  // 4 elements per lane: 32 lanes cover head_dim 128, 16 lanes head_dim 64
    const uint2 a = *reinterpret_cast<const uint2*>(out + row * o_rs + head * o_hs + lane * 4);
    const uint2 b = *reinterpret_cast<const uint2*>(dout + row * do_rs + head * do_hs + lane * 4);
    const T* pa = reinterpret_cast<const T*>(&a);
    const T* pb = reinterpret_cast<const T*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) s += static_cast<float>(pa[i]) * static_cast<float>(pb[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const size_t b_ = row / lse_S, sidx = row % lse_S;
    delta[(b_ * hq + head) * static_cast<size_t>(lse_S) + sidx] = s;
  }
}

}  // namespace bwd

const char* attn_bwd_delta_launch(int dtype, const TensorView& out, const TensorView& dout, float* delta, int lse_S,
                                  int head_dim, cudaStream_t stream) {
  const int rows = static_cast<int>(out.rows), hq = out.heads;
  const long long warps = static_cast<long long>(rows) * hq;
  if (warps == 0) return nullptr;
  const int threads = 256;
  const int blocks = static_cast<int>((warps * 32 + threads - 1) / threads);
  if (dtype == kDtypeBF16) {
    bwd::bwd_delta_kernel<__nv_bfloat16><<<blocks, threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(out.ptr), static_cast<const __nv_bfloat16*>(dout.ptr), delta, rows, hq, lse_S,
        out.row_stride, out.head_stride, dout.row_stride, dout.head_stride, head_dim);
  } else {
    bwd::bwd_delta_kernel<__half><<<blocks, threads, 0, stream>>>(
        static_cast<const __half*>(out.ptr), static_cast<const __half*>(dout.ptr), delta, rows, hq, lse_S,
        out.row_stride, out.head_stride, dout.row_stride, dout.head_stride, head_dim);
  }
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

const char* attn_bwd_launch(int dtype, const TensorView& q, const TensorView& dout, const TensorView& k,
                            const TensorView& v, const TensorView& k_stage, const TensorView& v_stage,
                            const TensorView& dq_accum, const BwdParams& p, cudaStream_t stream) {
  const int n_blocks = p.push.n_ctas + p.n_items * p.hkv;
  if (n_blocks <= 0) return nullptr;
  if (p.sig.world > 0) set_peer_timeout_from_env();
  CUtensorMap tq, tdo, tk, tv, tks, tvs, tdq;
  const int d = p.head_dim;
  if (d != 64 && d != 128) return "the sm_100a backward is instantiated for head_dim 64 and 128";
  if (const char* e = make_tensor_map(&tq, q, 2, bwd::kTileQ, d)) return e;
  if (const char* e = make_tensor_map(&tdo, dout, 2, bwd::kTileQ, d)) return e;
  if (const char* e = make_tensor_map(&tk, k, 2, bwd::kTileK, d)) return e;
  if (const char* e = make_tensor_map(&tv, v, 2, bwd::kTileK, d)) return e;
  if (const char* e = make_tensor_map(&tks, k_stage, 2, bwd::kTileK, d)) return e;
  if (const char* e = make_tensor_map(&tvs, v_stage, 2, bwd::kTileK, d)) return e;
  if (const char* e = make_plain_tensor_map(&tdq, dq_accum, 4, bwd::kTileQ, d)) return e;
  dim3 grid(n_blocks, 1, 1), block(bwd::kThreads, 1, 1);
  cudaError_t err;
  err = cudaSuccess;
  auto launch = [&](auto kern, int smem) {
    if (p.push.n_ctas > 0 && smem < kPushSmemBytes) smem = kPushSmemBytes;  // (head dim 64: the push ring is larger)
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err == cudaSuccess) kern<<<grid, block, smem, stream>>>(tq, tdo, tk, tv, tks, tvs, tdq, p);
  };
  auto pick = [&](auto tag) {
    using T = decltype(tag);
    if (d == 128) {
      if (p.window) launch(bwd::attn_bwd_kernel<T, true, 128>, bwd::Geo<128>::kSmemBytes);
      else launch(bwd::attn_bwd_kernel<T, false, 128>, bwd::Geo<128>::kSmemBytes);
    } else {
      if (p.window) launch(bwd::attn_bwd_kernel<T, true, 64>, bwd::Geo<64>::kSmemBytes);
      else launch(bwd::attn_bwd_kernel<T, false, 64>, bwd::Geo<64>::kSmemBytes);
    }
  };
  if (dtype == kDtypeBF16) pick(__nv_bfloat16{});
  else pick(__half{});
  if (err != cudaSuccess) return cudaGetErrorString(err);
  err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

}  // namespace rfa
