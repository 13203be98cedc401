// Blockwise flash-attention forward for sm_100a (head_dim 128, bf16/fp16).
//
// One CTA = one work item (<= 256 query rows of one chunk = two 128-row MMA tiles) x one query head.
// The CTA walks a *segment list*: runs of keys - possibly from several source ranks - each with its own
// causal diagonal.  The output accumulator never leaves tensor memory between segments, so a whole
// ring / zigzag / stripe / llama3 forward is ordinary online softmax over the concatenated segments
// (this replaces the reference's flash_attn call + fp32 out/lse merge per ring step,
// /root/reference/ring_flash_attn/ring_flash_attn.py:26-63 and utils.py:32-73).
//
// Warp roles (384 threads):
//   warp 0      TMA producer: Q tiles once, then K / V tiles through a ring of 32 KB stages
//   warp 1      tcgen05.mma issuer: S_t = Q_t K^T (SS), O_t += P_t V (P from TMEM, V MN-major from smem)
//   warp 2      TMEM allocator
//   warps 4-7   softmax warpgroup for tile 0 (thread == row): S -> regs, mask, max, exp2, P -> TMEM,
//               lazy O rescale, final normalisation + store of out / lse
//   warps 8-11  the same for tile 1
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_t aliases the first 64
// columns of S_t.  The MMA warp ping-pongs the two tiles so one tile's softmax overlaps the other's MMAs.
#include <math_constants.h>
#include <stdio.h>

#include <cstdlib>
#include <type_traits>

#include "attn_common.h"
#include "comm_device.cuh"
#include "sm100_ptx.cuh"

namespace rfa {

namespace fwd {

#ifdef RFA_TRACE
// slot layout per key-tile iteration (16 stamps): 0-5 MMA thread, 6-10 softmax tile 0, 11-15 softmax tile 1
#define RFA_STAMP(cond, iter, slot)                                                              \
  do {                                                                                            \
    if ((cond) && cta == 0 && p.trace != nullptr && (iter) < 64) p.trace[(iter)*16 + (slot)] = clock64(); \
  } while (0)
#else
#define RFA_STAMP(cond, iter, slot) \
  do {                              \
  } while (0)
#endif

constexpr int kTile = 128;          // rows per MMA tile (queries and keys)
constexpr int kStages = 4;          // K/V ring slots
constexpr int kHalfBytes = kTile * 128;  // one 64-element-wide (128-byte) swizzled sub-tile: 16 KB
constexpr int kThreads = 384;
// Head dim kD is a template parameter (64 or 128): a tile is kD / 64 sub-tiles of 128 rows x 128 bytes, the output
// accumulator kD tensor-memory columns.  (The reference inherits its head sizes from flash-attn; smaller sizes are
// zero-padded to the next supported one in parallel/api.py.)
template <int kD>
constexpr int tile_bytes() {
  return kTile * kD * 2;
}
template <int kD>
constexpr int smem_bytes() {
  return 2 * tile_bytes<kD>() + kStages * tile_bytes<kD>() + 1024 /*barriers*/ + 1024 /*align slack*/;
}
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 384;
constexpr float kRescaleThreshold = 8.0f;  // log2 units; P stays below 2^8

struct Barriers {
  uint64_t q_full[2];
  uint64_t kv_full[kStages];
  uint64_t kv_empty[kStages];
  uint64_t s_full[2];
  uint64_t p_ready[2];
  uint64_t o_done[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// Element traits.  Fp8E4M3 is the tag of the experimental fp8 forward (RFA_B200_FP8_KERNEL=1): q / k / v are e4m3
// bytes, a tile is ONE 128-byte swizzle span per row (16 KB, loaded into the same 32 KB slots), both GEMMs are
// kind::f8f6f4 with K = 32 per instruction, P is written back to tensor memory as e4m3 (four per column) and the
// output is bf16.  Block descales (q: per token block x head, k / v: per 128-token block x kv head, or coarser) come
// in through FwdParams::q_scale / k_scale / v_scale; they are applied to the fp32 scores and folded into P, at no
// extra tensor-core or memory cost.
struct Fp8E4M3 {};
template <typename T>
struct Elem {
  static constexpr bool kFp8 = false;
  static constexpr uint32_t kBytes = 2;
  static constexpr uint32_t kFmt = Pack2<T>::kFmt;
  using Out = T;
};
template <>
struct Elem<Fp8E4M3> {
  static constexpr bool kFp8 = true;
  static constexpr uint32_t kBytes = 1;
  static constexpr uint32_t kFmt = 0;  // e4m3
  using Out = __nv_bfloat16;
};

// Per-(item, segment) geometry shared by all roles so that they agree on the iteration space.
struct SegGeom {
  int kv_row0, kv_len, diag;
  int n_tiles;  // key tiles to visit
};

__device__ __forceinline__ SegGeom seg_geom(const KVSegment& s, const WorkItem& it) {
  SegGeom g;
  g.kv_row0 = s.kv_row0;
  g.kv_len = s.kv_len;
  g.diag = s.diag;
  const int last_row = it.q_off + it.q_rows - 1;
  long long lim = static_cast<long long>(last_row) + s.diag + 1;  // keys [0, lim) reachable by the last row
  int reach = lim < 0 ? 0 : (lim > s.kv_len ? s.kv_len : static_cast<int>(lim));
  g.n_tiles = (reach + kTile - 1) / kTile;
  return g;
}
// Is tile t (rows [q_off + 128 t, ...)) touching key tile jj at all?
__device__ __forceinline__ bool tile_active(const SegGeom& g, const WorkItem& it, int t, int jj) {
  const int n_t = t == 0 ? (it.q_rows < kTile ? it.q_rows : kTile) : it.q_rows - kTile;
  if (n_t <= 0) return false;
  const int last_row = it.q_off + t * kTile + n_t - 1;
  return static_cast<long long>(jj) * kTile <= static_cast<long long>(last_row) + g.diag;
}
__device__ __forceinline__ bool tile_needs_mask(const SegGeom& g, const WorkItem& it, int t, int jj) {
  const int first_row = it.q_off + t * kTile;
  const bool ragged = (jj + 1) * kTile > g.kv_len;
  const bool diagonal = static_cast<long long>(jj) * kTile + (kTile - 1) > static_cast<long long>(first_row) + g.diag;
  return ragged || diagonal;
}
// Sliding window (kWindow kernels): key j of the segment is visible to chunk row i only if j >= i + lo.  The host
// trims every segment per work item so that the loop still starts at key tile 0 (ops/attn_cuda.py); what remains
// for the kernel is the slanted lower edge of the band.
__device__ __forceinline__ bool tile_needs_lower_mask(int lo, const WorkItem& it, int t, int jj) {
  const long long last_row = static_cast<long long>(it.q_off) + t * kTile + (kTile - 1);
  return last_row + lo > static_cast<long long>(jj) * kTile;
}

template <typename T, bool kWindow, int kD>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_ks,
                const __grid_constant__ CUtensorMap tm_vs, const __grid_constant__ FwdParams p) {
  // The first blocks of the grid are communication CTAs: they push this rank's K/V rows to the peers that
  // need them while the remaining (compute) CTAs already work on the local shard.
  if (static_cast<int>(blockIdx.x) < p.push.n_ctas) {
    if (p.push.use_tma) {
      extern __shared__ uint8_t push_smem_raw[];
      push_role_tma(p.push, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(push_smem_raw) + 1023) & ~uintptr_t(1023)));
    } else {
      push_role(p.push);
    }
    return;
  }
  using E = Elem<T>;
  static_assert(kD == 64 || kD == 128, "head dim 64 or 128");
  static_assert(!E::kFp8 || kD == 128, "the fp8 forward is instantiated for head dim 128 only");
  constexpr int kTileBytes = tile_bytes<kD>();   // smem slot of one tile (fp8 tiles use the first half of it)
  constexpr int kSubTiles = E::kFp8 ? 1 : kD / 64;  // 128-byte spans per row
  constexpr uint32_t kTxBytes = kTile * kD * E::kBytes;
  const int cta = static_cast<int>(blockIdx.x) - p.push.n_ctas;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                       // 2 x 32 KB
  uint8_t* smem_kv = smem + 2 * kTileBytes;     // kStages x 32 KB
  Barriers* bars = reinterpret_cast<Barriers*>(smem_kv + kStages * kTileBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = cta / p.n_items;  // consecutive CTAs share a head => K/V tiles are reused out of L2
  const int kv_head = head / (p.hq / p.hkv);
  WorkItem it = p.items[cta % p.n_items];
  if (p.flags & 4) it.seg_count = 0;  // timing experiment: communication only (compute CTAs have nothing to do)
  const int n_rows0 = it.q_rows < kTile ? it.q_rows : kTile;
  const int n_rows1 = it.q_rows - n_rows0;
  const bool has_t1 = n_rows1 > 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_ks);
    tma_prefetch_desc(&tm_vs);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->q_full[i], 1);
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_ready[i], 128);
      mbar_init(&bars->o_done[i], 1);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars->kv_full[i], 1);
      mbar_init(&bars->kv_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(&bars->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  if (warp < 4) {
   reg_dealloc<88>();
   if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->q_full[0], kTxBytes);
#pragma unroll
      for (int h = 0; h < kSubTiles; ++h)
        tma_load_3d(smem_q + h * kHalfBytes, &tm_q, &bars->q_full[0], 64 * h, head, it.q_row0);
      if (has_t1) {
        mbar_arrive_expect_tx(&bars->q_full[1], kTxBytes);
#pragma unroll
        for (int h = 0; h < kSubTiles; ++h)
          tma_load_3d(smem_q + kTileBytes + h * kHalfBytes, &tm_q, &bars->q_full[1], 64 * h, head, it.q_row0 + kTile);
      }
      uint32_t slot = 0, phase = 0;
      for (int si = 0; si < it.seg_count; ++si) {
        const KVSegment sg = p.segs[it.seg_begin + si];
        const SegGeom g = seg_geom(sg, it);
        const bool staged = sg.flag >= 0 && p.ready_flags != nullptr;
        if (g.n_tiles > 0 && staged && !(p.flags & 2)) {  // (flags bit 1: timing experiment, compute only)
          wait_epoch(p.ready_flags + sg.flag, p.ready_epoch, "fwd kv ready", p.sig.my_rank, sg.flag);
          fence_proxy_async_all();
        }
        for (int jj = 0; jj < g.n_tiles; ++jj) {
          const int row = g.kv_row0 + jj * kTile;
          for (int kv = 0; kv < 2; ++kv) {
            mbar_wait(&bars->kv_empty[slot], phase ^ 1);
            uint8_t* dst = smem_kv + slot * kTileBytes;
            const CUtensorMap* tm = staged ? (kv == 0 ? &tm_ks : &tm_vs) : (kv == 0 ? &tm_k : &tm_v);
            mbar_arrive_expect_tx(&bars->kv_full[slot], kTxBytes);
#pragma unroll
            for (int h = 0; h < kSubTiles; ++h)
              tma_load_3d(dst + h * kHalfBytes, tm, &bars->kv_full[slot], 64 * h, kv_head, row);
            if (++slot == kStages) {
              slot = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loop (so addresses / descriptors stay in uniform registers); one elected lane
    // issues the tcgen05 instructions and commits.
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_qk = E::kFp8 ? umma_idesc_f8(E::kFmt, E::kFmt, kTile, kTile, 0, 0)
                                            : umma_idesc_f16(E::kFmt, kTile, kTile, 0, 0);
      constexpr uint32_t idesc_pv = E::kFp8 ? umma_idesc_f8(E::kFmt, E::kFmt, kTile, kD, 0, 1)
                                            : umma_idesc_f16(E::kFmt, kTile, kD, 0, 1);
      const uint32_t q_base = smem_u32(smem_q);
      const uint32_t kv_base = smem_u32(smem_kv);
      const uint32_t col_s[2] = {tmem + kColS0, tmem + kColS1};
      const uint32_t col_o[2] = {tmem + kColO0, tmem + kColO1};

      // Descriptors: high words are constants, low words are (base >> 4) + a compile-time step per k.
      constexpr uint32_t hi = umma_desc_hi(1024, kSwizzle128B);
      const uint32_t q_lo[2] = {umma_desc_lo(q_base, 16), umma_desc_lo(q_base + kTileBytes, 16)};
      const uint32_t k_lo0 = umma_desc_lo(kv_base, 16);           // K tile read K-major
      const uint32_t v_lo0 = umma_desc_lo(kv_base, kHalfBytes);   // V tile read MN-major (LBO = 64-dim half)
      constexpr uint32_t slot_step = kTileBytes >> 4;

      auto issue_qk = [&](int t, uint32_t k_slot) {
        const uint32_t a0 = q_lo[t], b0 = k_lo0 + k_slot * slot_step;
        if constexpr (E::kFp8) {
          // one 128-byte span per row holds all 128 head dims: 4 instructions of K = 32 bytes
#pragma unroll
          for (int k = 0; k < kD / 32; ++k)
            umma_ss2_f8(col_s[t], a0 + ((k * 32) >> 4), hi, b0 + ((k * 32) >> 4), hi, idesc_qk, k > 0);
        } else {
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            const uint32_t off = ((k >> 2) * kHalfBytes + (k & 3) * 32) >> 4;
            umma_ss2(col_s[t], a0 + off, hi, b0 + off, hi, idesc_qk, k > 0);
          }
        }
        umma_commit(&bars->s_full[t]);
      };
      auto issue_pv = [&](int t, uint32_t v_slot, bool accumulate) {
        // V tile is [128 keys][2 x 64 dims] -> MN-major B: LBO = stride between the two 64-wide halves,
        // SBO = stride between 8-key groups; one MMA consumes 16 keys = 2048 bytes.
        const uint32_t b0 = v_lo0 + v_slot * slot_step;
        if constexpr (E::kFp8) {
          // e4m3: the 128 dims of a key are one span (no second half, LBO unused); one instruction consumes 32
          // keys = 4096 bytes of V and 32 packed P values = 8 TMEM columns
#pragma unroll
          for (int k = 0; k < kTile / 32; ++k)
            umma_ts2_f8(col_o[t], col_s[t] + k * 8, b0 + k * (4096 >> 4), hi, idesc_pv,
                        (accumulate || k > 0) ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < kTile / 16; ++k)
            umma_ts2(col_o[t], col_s[t] + k * 8, b0 + k * (2048 >> 4), hi, idesc_pv,
                     (accumulate || k > 0) ? 1u : 0u);
        }
      };

      mbar_wait(&bars->q_full[0], 0);
      if (has_t1) mbar_wait(&bars->q_full[1], 0);
      tc_fence_after();

      // Software pipeline (round 2).  Per q-tile t the dependent chain of one key tile is
      //   S_t(j) complete -> softmax -> P_t(j) -> PV_t(j) -> QK_t(j+1) -> S_t(j+1) complete
      // (S and P share tensor-memory columns, so QK_t(j+1) has to queue behind PV_t(j)).  Round 1 issued
      // QK_0(j+1) only after the loop bookkeeping and the K wait of iteration j+1 (~530 cycles of a 3940-cycle
      // iteration, profiles/trace_fwd_cta0.log); now the geometry of step j+1 is computed and its K tile awaited
      // BEFORE P_t(j) is awaited, and QK_t(j+1) is issued back to back with PV_t(j) for both tiles.
      struct Step {
        bool valid, a0, a1;
        uint32_t k_slot, k_phase, v_slot, v_phase;
      };
      uint32_t slot = 0, phase = 0;  // ring position of the next K tile
      int it_si = 0, it_jj = 0;
      SegGeom it_g = it.seg_count > 0 ? seg_geom(p.segs[it.seg_begin], it) : SegGeom{0, 0, 0, 0};
      auto next_step = [&]() {
        Step st{};
        while (it_si < it.seg_count && it_jj >= it_g.n_tiles) {
          ++it_si;
          it_jj = 0;
          if (it_si < it.seg_count) it_g = seg_geom(p.segs[it.seg_begin + it_si], it);
        }
        if (it_si >= it.seg_count) return st;
        st.valid = true;
        st.a0 = tile_active(it_g, it, 0, it_jj);
        st.a1 = tile_active(it_g, it, 1, it_jj);
        st.k_slot = slot;
        st.k_phase = phase;
        if (++slot == kStages) {
          slot = 0;
          phase ^= 1;
        }
        st.v_slot = slot;
        st.v_phase = phase;
        if (++slot == kStages) {
          slot = 0;
          phase ^= 1;
        }
        ++it_jj;
        return st;
      };
      uint32_t p_phase[2] = {0, 0};
      bool o_started[2] = {false, false};
      Step cur = next_step();
      if (cur.valid) {
        mbar_wait(&bars->kv_full[cur.k_slot], cur.k_phase);
        tc_fence_after();
        if (leader) {
          if (cur.a0) issue_qk(0, cur.k_slot);
          if (cur.a1) issue_qk(1, cur.k_slot);
          umma_commit(&bars->kv_empty[cur.k_slot]);
        }
      }
      int xi = 0;
      while (cur.valid) {
        const Step nx = next_step();
        RFA_STAMP(true, xi, 0);
        mbar_wait(&bars->kv_full[cur.v_slot], cur.v_phase);
        if (nx.valid) mbar_wait(&bars->kv_full[nx.k_slot], nx.k_phase);
        tc_fence_after();
        RFA_STAMP(true, xi, 1);
        if (cur.a0) {
          mbar_wait(&bars->p_ready[0], p_phase[0]);
          p_phase[0] ^= 1;
          tc_fence_after();
          RFA_STAMP(true, xi, 2);
          if (leader) issue_pv(0, cur.v_slot, o_started[0]);
          o_started[0] = true;
        }
        if (leader && nx.valid && nx.a0) issue_qk(0, nx.k_slot);
        RFA_STAMP(true, xi, 3);
        if (cur.a1) {
          mbar_wait(&bars->p_ready[1], p_phase[1]);
          p_phase[1] ^= 1;
          tc_fence_after();
          RFA_STAMP(true, xi, 4);
          if (leader) issue_pv(1, cur.v_slot, o_started[1]);
          o_started[1] = true;
        }
        if (leader) {
          umma_commit(&bars->kv_empty[cur.v_slot]);
          if (nx.valid) {
            if (nx.a1) issue_qk(1, nx.k_slot);
            umma_commit(&bars->kv_empty[nx.k_slot]);
          }
        }
        RFA_STAMP(true, xi, 5);
        cur = nx;
        ++xi;
        __syncwarp();
      }
      if (leader) {
        umma_commit(&bars->o_done[0]);
        umma_commit(&bars->o_done[1]);
      }
      __syncwarp();
    }
   }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    reg_alloc<208>();
    const int t = (warp - 4) >> 2;                 // tile handled by this warpgroup
    const int row_in_tile = ((warp & 3) << 5) | lane;
    const int n_rows = t == 0 ? n_rows0 : n_rows1;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem + (t == 0 ? kColS0 : kColS1) + lane_addr;
    const uint32_t t_o = tmem + (t == 0 ? kColO0 : kColO1) + lane_addr;
    const int chunk_row = it.q_off + t * kTile + row_in_tile;  // row index inside the chunk (diagonal space)

    // Scores in log2 units are s * t2 with t2 = softmax_scale * log2(e); the running reference max m2 is kept in
    // those units.  fp8 (block-scaled inputs): the descale of this query row (per token block x head) rides on t2
    // for the whole row, the descale of the key block (128-token blocks x kv head, or per head) is multiplied in per
    // key tile, and the V descale of the key block enters P as a ratio to the head's largest V descale (so that P
    // stays in e4m3's normal range) with the reference value applied once in the epilogue.
    float base_l2 = p.scale_log2;
    [[maybe_unused]] float inv_vref = 1.f;
    if constexpr (E::kFp8) {
      const int qrow = it.q_row0 + t * kTile + (row_in_tile < n_rows ? row_in_tile : 0);
      base_l2 *= p.q_scale[static_cast<size_t>(qrow / p.q_scale_block) * p.hq + head];
      inv_vref = 1.0f / p.v_ref[kv_head];
    }
    float m2 = -CUDART_INF_F;  // reference max (log2 units)
    float l = 0.f;
    bool first = true;
    uint32_t s_phase = 0;

    // The two softmax warpgroups take turns in the MUFU-heavy exp section (named barriers 1 + t).  Without
    // the hand-off both tiles drift into phase: their exp sections then share the 4 SFU lanes per scheduler,
    // and the tensor pipe idles while both run.  Strict alternation keeps one tile's softmax under the other
    // tile's MMAs.  Both groups execute one hand-off per key tile, active or not, so the counts always match.
    int handoffs_left = 0;
    const bool take_turns = has_t1 && !(p.flags & 1);
    if (take_turns) {
      for (int si = 0; si < it.seg_count; ++si) handoffs_left += seg_geom(p.segs[it.seg_begin + si], it).n_tiles;
      if (t == 1 && handoffs_left > 0) named_bar_arrive(1, 256);  // tile 0 goes first
    }
    auto turn_wait = [&]() {
      if (take_turns) named_bar_sync(1 + t, 256);
    };
    auto turn_pass = [&]() {
      if (take_turns) {
        --handoffs_left;
        if (!(t == 1 && handoffs_left == 0)) named_bar_arrive(1 + (1 - t), 256);  // nobody waits after the last one
      }
    };

    if (n_rows > 0) {
      int xi = 0;
      [[maybe_unused]] const bool stamper = (threadIdx.x & 127) == 0;  // RFA_TRACE builds only
      for (int si = 0; si < it.seg_count; ++si) {
        const KVSegment sg = p.segs[it.seg_begin + si];
        const SegGeom g = seg_geom(sg, it);
        int seg_lo = 0;
        if constexpr (kWindow) seg_lo = p.seg_lo[it.seg_begin + si];
        // fp8: row of the key tile in the scale tables (staged rows are indexed like the staging buffer, local rows
        // are offset to this rank's slot)
        [[maybe_unused]] const long long scale_row0 =
            g.kv_row0 + ((sg.flag >= 0 && p.ready_flags != nullptr) ? 0ll : static_cast<long long>(p.kv_scale_row0));
        for (int jj = 0; jj < g.n_tiles; ++jj, ++xi) {
          if (!tile_active(g, it, t, jj)) {
            turn_wait();
            turn_pass();
            continue;
          }
          mbar_wait(&bars->s_full[t], s_phase);
          s_phase ^= 1;
          tc_fence_after();
          RFA_STAMP(stamper, xi, 6 + 5 * t);

          uint32_t sr[128];
          tmem_ld32(t_s + 0, sr + 0);
          tmem_ld32(t_s + 32, sr + 32);
          tmem_ld32(t_s + 64, sr + 64);
          tmem_ld32(t_s + 96, sr + 96);
          tmem_ld_wait();
          RFA_STAMP(stamper, xi, 7 + 5 * t);
          float s[128];
#pragma unroll
          for (int c = 0; c < 128; ++c) s[c] = __uint_as_float(sr[c]);

          bool masked = tile_needs_mask(g, it, t, jj);
          if constexpr (kWindow) masked = masked || tile_needs_lower_mask(seg_lo, it, t, jj);
          if (masked) {
            long long lim_ll = static_cast<long long>(chunk_row) + g.diag;
            if (lim_ll > g.kv_len - 1) lim_ll = g.kv_len - 1;
            lim_ll -= static_cast<long long>(jj) * kTile;
            const int lim = lim_ll < -1 ? -1 : (lim_ll > 127 ? 127 : static_cast<int>(lim_ll));
            if constexpr (kWindow) {
              const long long lo_ll = static_cast<long long>(chunk_row) + seg_lo - static_cast<long long>(jj) * kTile;
              const int lo_lim = lo_ll < 0 ? 0 : (lo_ll > 128 ? 128 : static_cast<int>(lo_ll));
#pragma unroll
              for (int c = 0; c < 128; ++c) s[c] = (c <= lim && c >= lo_lim) ? s[c] : -CUDART_INF_F;
            } else {
#pragma unroll
              for (int c = 0; c < 128; ++c) s[c] = c <= lim ? s[c] : -CUDART_INF_F;
            }
          }

          float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
          for (int c = 4; c < 128; c += 4) {
            mx0 = fmaxf(mx0, s[c]);
            mx1 = fmaxf(mx1, s[c + 1]);
            mx2 = fmaxf(mx2, s[c + 2]);
            mx3 = fmaxf(mx3, s[c + 3]);
          }
          float t2 = base_l2;
          [[maybe_unused]] float log2_ratio = 0.f, inv_ratio = 1.f;
          if constexpr (E::kFp8) {
            const size_t sb = static_cast<size_t>((scale_row0 + static_cast<long long>(jj) * kTile) / p.kv_scale_block) *
                                  p.hkv + kv_head;
            t2 *= p.k_scale[sb];
            const float ratio = p.v_scale[sb] * inv_vref;
            log2_ratio = __log2f(ratio);
            inv_ratio = 1.0f / ratio;
          }
          const float m_new = fmaxf(m2, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * t2);
          // lazy rescale: only move the reference max when it grew by more than the threshold
          const bool need = (m_new - m2) > kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {
            const float f = need ? fast_exp2(m2 - m_new) : 1.0f;
            if (need) {
              l *= f;
              m2 = m_new;
            }
            if (!first) {
              // S_full for this key tile implies the previous PV of this tile has completed, so O is quiescent.
#pragma unroll
              for (int c = 0; c < kD; c += 32) {
                uint32_t orr[32];
                tmem_ld32(t_o + c, orr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
                tmem_st32(t_o + c, orr);
              }
            }
          }
          first = false;
          const float mc = (m2 == -CUDART_INF_F ? 0.f : m2) - log2_ratio;
          turn_wait();
          RFA_STAMP(stamper, xi, 8 + 5 * t);
          // exp2(s * c - m * c) on packed pairs (FFMA2 for the scaling, MUFU.EX2 for the exponential; a polynomial
          // exp2 on the FMA pipes for a share of the elements was measured twice on B200 - round 1 and again after
          // the round-2 pipeline change, profiles/r2/trip_fwd_tuning.log - and never beat MUFU-only, so it is gone).
          const uint64_t sc2 = pack2(t2, t2), nmc2 = pack2(-mc, -mc);
          uint64_t lsum = pack2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 128; c += 32) {
            uint32_t pk[16];
            [[maybe_unused]] float pe0 = 0.f, pe1 = 0.f;  // fp8 only: the even pair waiting for its odd partner
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint64_t x = ffma2(pack2(s[c + 2 * i], s[c + 2 * i + 1]), sc2, nmc2);
              float x0, x1;
              unpack2(x, x0, x1);
              const float e0 = fast_exp2(x0), e1 = fast_exp2(x1);
              lsum = fadd2(lsum, pack2(e0, e1));
              if constexpr (E::kFp8) {
                if (i & 1) {
                  pk[i >> 1] = pack4_e4m3(pe0, pe1, e0, e1);
                } else {
                  pe0 = e0;
                  pe1 = e1;
                }
              } else {
                pk[i] = Pack2<T>::pack(e0, e1);
              }
            }
            if constexpr (E::kFp8) {
              tmem_st8(t_s + (c >> 2), pk);  // 32 e4m3 = 8 columns per 32 scores
            } else {
              tmem_st16(t_s + (c >> 1), pk);
            }
          }
          float l0, l1;
          unpack2(lsum, l0, l1);
          turn_pass();
          RFA_STAMP(stamper, xi, 9 + 5 * t);
          l += (l0 + l1) * inv_ratio;
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&bars->p_ready[t]);
          RFA_STAMP(stamper, xi, 10 + 5 * t);
        }
      }

      // ---------------------------------------------------------------- epilogue: O / l -> out, lse
      const int row = it.q_row0 + t * kTile + row_in_tile;
      const bool row_ok = row_in_tile < n_rows;
      using OutT = typename E::Out;
      OutT* out_row = reinterpret_cast<OutT*>(p.out) + (static_cast<size_t>(row) * p.hq + head) * kD;
      if (!first) {
        mbar_wait(&bars->o_done[t], 0);
        tc_fence_after();
        float inv = l > 0.f ? 1.0f / l : 0.f;
        if constexpr (E::kFp8) inv *= p.v_ref[kv_head];
#pragma unroll
        for (int c = 0; c < kD; c += 32) {
          uint32_t orr[32];
          tmem_ld32(t_o + c, orr);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 v;
              v.x = Pack2<OutT>::pack(__uint_as_float(orr[i + 0]) * inv, __uint_as_float(orr[i + 1]) * inv);
              v.y = Pack2<OutT>::pack(__uint_as_float(orr[i + 2]) * inv, __uint_as_float(orr[i + 3]) * inv);
              v.z = Pack2<OutT>::pack(__uint_as_float(orr[i + 4]) * inv, __uint_as_float(orr[i + 5]) * inv);
              v.w = Pack2<OutT>::pack(__uint_as_float(orr[i + 6]) * inv, __uint_as_float(orr[i + 7]) * inv);
              *reinterpret_cast<uint4*>(out_row + c + i) = v;
            }
          }
        }
      } else if (row_ok) {
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < kD; c += 8) *reinterpret_cast<uint4*>(out_row + c) = z;
      }
      if (row_ok) {
        const float lse = l > 0.f ? (m2 + __log2f(l)) * 0.6931471805599453f : -CUDART_INF_F;
        const size_t b = row / p.lse_S, sidx = row % p.lse_S;
        p.lse[(b * p.hq + head) * static_cast<size_t>(p.lse_S) + sidx] = lse;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<kTmemCols>(tmem);
  if (threadIdx.x == 0 && p.sig.world > 0) consumer_done(p.sig);
}

}  // namespace fwd

const char* attn_fwd_launch(int dtype, const TensorView& q, const TensorView& k, const TensorView& v,
                            const TensorView& k_stage, const TensorView& v_stage, const FwdParams& p,
                            cudaStream_t stream) {
  const int n_blocks = p.push.n_ctas + p.n_items * p.hq;
  if (n_blocks <= 0) return nullptr;
  if (p.sig.world > 0) set_peer_timeout_from_env();
  CUtensorMap tq, tk, tv, tks, tvs;
  const int eb = dtype == kDtypeE4M3 ? 1 : 2;
  const int d = p.head_dim;
  if (d != 64 && d != 128) return "the sm_100a forward is instantiated for head_dim 64 and 128";
  if (dtype == kDtypeE4M3 && (p.q_scale == nullptr || p.k_scale == nullptr || p.v_scale == nullptr ||
                              p.v_ref == nullptr || p.q_scale_block < 1 || p.kv_scale_block < 1 ||
                              p.seg_lo != nullptr || d != 128))
    return "fp8 forward needs q / k / v descale tables and head_dim 128, and does not support sliding windows yet";
  if (const char* e = make_tensor_map(&tq, q, eb, fwd::kTile, d)) return e;
  if (const char* e = make_tensor_map(&tk, k, eb, fwd::kTile, d)) return e;
  if (const char* e = make_tensor_map(&tv, v, eb, fwd::kTile, d)) return e;
  if (const char* e = make_tensor_map(&tks, k_stage, eb, fwd::kTile, d)) return e;
  if (const char* e = make_tensor_map(&tvs, v_stage, eb, fwd::kTile, d)) return e;
  dim3 grid(n_blocks, 1, 1), block(fwd::kThreads, 1, 1);
  cudaError_t err = cudaSuccess;
  auto launch = [&](auto kern, int smem) {
    if (p.push.n_ctas > 0 && smem < kPushSmemBytes) smem = kPushSmemBytes;  // (head dim 64: the push ring is larger)
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err == cudaSuccess) kern<<<grid, block, smem, stream>>>(tq, tk, tv, tks, tvs, p);
  };
  auto pick = [&](auto tag, auto window) {
    using T = decltype(tag);
    constexpr bool kW = decltype(window)::value;
    if (d == 128) launch(fwd::attn_fwd_kernel<T, kW, 128>, fwd::smem_bytes<128>());
    else launch(fwd::attn_fwd_kernel<T, kW, 64>, fwd::smem_bytes<64>());
  };
  if (dtype == kDtypeE4M3) {  // fp8 forward (e4m3 in, bf16 out)
    launch(fwd::attn_fwd_kernel<fwd::Fp8E4M3, false, 128>, fwd::smem_bytes<128>());
  } else if (p.seg_lo != nullptr) {  // sliding-window tables: the lower band edge is masked in-kernel
    if (dtype == kDtypeBF16) pick(__nv_bfloat16{}, std::true_type{});
    else pick(__half{}, std::true_type{});
  } else {
    if (dtype == kDtypeBF16) pick(__nv_bfloat16{}, std::false_type{});
    else pick(__half{}, std::false_type{});
  }
  if (err != cudaSuccess) return cudaGetErrorString(err);
  err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

}  // namespace rfa
