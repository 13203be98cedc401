// Forward variant with 64-key softmax steps ("h64", opt-in: RFA_B200_FWD_H64=1; written after the last round-1
// hardware session, validated in round 2).
//
// Motivation (profiles/trace_fwd_cta0.log): in attn_fwd_sm100.cu the S tile of a q-tile and its P alias the same
// tensor-memory columns, so Q K^T of key tile j+1 cannot be issued before P V of tile j - the softmax warpgroup
// idles for a whole MMA round trip (~1700 of ~3900 cycles per key tile).  Here the 128 S columns of a q-tile are
// two 64-key halves S_a | S_b.  While the softmax works on one half the tensor pipe already produces the other
// (and, behind its P V, the first half of the next key tile), so a warpgroup always finds its next scores ready:
//
//   tensor pipe, per q-tile t:   ... PV(j-1,a) QK(j,a) PV(j-1,b) QK(j,b) PV(j,a) QK(j+1,a) ...
//   softmax warpgroup t:         ...        sm(j-1,b)   |  sm(j,a)   |   sm(j,b)   | ...
//
// Everything else - work items, segment tables, TMA producer, K/V ring, epilogue, the push CTAs of the fused
// multi-GPU mode - is the layout of attn_fwd_sm100.cu, and the launch takes the same FwdParams.
// Costs: Q K^T instructions are N = 64 (48 instead of 2 x 32 cycles per 128 keys, measured), twice as many
// barrier round trips.  Not supported here (use the default kernel): sliding windows, fp8.
#include <math_constants.h>
#include <stdio.h>

#include <cstdlib>
#include <type_traits>

#include "attn_common.h"
#include "comm_device.cuh"
#include "sm100_ptx.cuh"

namespace rfa {
namespace fwd64 {

constexpr int kD = 128;
constexpr int kTile = 128;  // query rows per MMA tile, keys per K/V smem tile
constexpr int kHalf = 64;   // keys per softmax step
constexpr int kStages = 4;
constexpr int kTileBytes = kTile * kD * 2;
constexpr int kHalfBytes = kTileBytes / 2;  // one 64-dim-wide swizzled sub-tile
constexpr int kThreads = 384;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 384;
constexpr float kRescaleThreshold = 8.0f;

struct Barriers {
  uint64_t q_full[2];
  uint64_t kv_full[kStages];
  uint64_t kv_empty[kStages];
  uint64_t s_full[2][2];   // [q-tile][half]
  uint64_t p_ready[2][2];  // [q-tile][half], 128 arrivals
  uint64_t pv_done[2];     // one phase per P V of a q-tile (the softmax waits on it only before an O rescale)
  uint64_t o_done[2];
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int kSmemBytes = 2 * kTileBytes + kStages * kTileBytes + 1024 + 1024;

struct SegGeom {
  int kv_row0, kv_len, diag;
  int n_tiles;
};
__device__ __forceinline__ SegGeom seg_geom(const KVSegment& s, const WorkItem& it) {
  SegGeom g;
  g.kv_row0 = s.kv_row0;
  g.kv_len = s.kv_len;
  g.diag = s.diag;
  const int last_row = it.q_off + it.q_rows - 1;
  long long lim = static_cast<long long>(last_row) + s.diag + 1;
  int reach = lim < 0 ? 0 : (lim > s.kv_len ? s.kv_len : static_cast<int>(lim));
  g.n_tiles = (reach + kTile - 1) / kTile;
  return g;
}
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

// kPolyOf4: of every 4 element pairs of an unmasked step, this many take the polynomial exp2 on the FMA pipes
// (RFA_B200_POLY_EXP, as in the default kernel; the 64-column steps leave the registers for it).
template <typename T, int kPolyOf4>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_h64_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_ks,
                    const __grid_constant__ CUtensorMap tm_vs, const __grid_constant__ FwdParams p) {
  if (static_cast<int>(blockIdx.x) < p.push.n_ctas) {
    if (p.push.use_tma) {
      extern __shared__ uint8_t push_smem_raw[];
      push_role_tma(p.push, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(push_smem_raw) + 1023) & ~uintptr_t(1023)));
    } else {
      push_role(p.push);
    }
    return;
  }
  const int cta = static_cast<int>(blockIdx.x) - p.push.n_ctas;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem + 2 * kTileBytes;
  Barriers* bars = reinterpret_cast<Barriers*>(smem_kv + kStages * kTileBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = cta / p.n_items;
  const int kv_head = head / (p.hq / p.hkv);
  const WorkItem it = p.items[cta % p.n_items];
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
      mbar_init(&bars->pv_done[i], 1);
      mbar_init(&bars->o_done[i], 1);
      for (int h = 0; h < 2; ++h) {
        mbar_init(&bars->s_full[i][h], 1);
        mbar_init(&bars->p_ready[i][h], 128);
      }
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
      // ------------------------------------------------------------------ TMA producer (as in attn_fwd_sm100.cu)
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars->q_full[0], kTileBytes);
        tma_load_3d(smem_q, &tm_q, &bars->q_full[0], 0, head, it.q_row0);
        tma_load_3d(smem_q + kHalfBytes, &tm_q, &bars->q_full[0], 64, head, it.q_row0);
        if (has_t1) {
          mbar_arrive_expect_tx(&bars->q_full[1], kTileBytes);
          tma_load_3d(smem_q + kTileBytes, &tm_q, &bars->q_full[1], 0, head, it.q_row0 + kTile);
          tma_load_3d(smem_q + kTileBytes + kHalfBytes, &tm_q, &bars->q_full[1], 64, head, it.q_row0 + kTile);
        }
        uint32_t slot = 0, phase = 0;
        for (int si = 0; si < it.seg_count; ++si) {
          const KVSegment sg = p.segs[it.seg_begin + si];
          const SegGeom g = seg_geom(sg, it);
          const bool staged = sg.flag >= 0 && p.ready_flags != nullptr;
          if (g.n_tiles > 0 && staged) {
            wait_epoch(p.ready_flags + sg.flag, p.ready_epoch, "fwd kv ready", p.sig.my_rank, sg.flag);
            fence_proxy_async_all();
          }
          for (int jj = 0; jj < g.n_tiles; ++jj) {
            const int row = g.kv_row0 + jj * kTile;
            for (int kv = 0; kv < 2; ++kv) {
              mbar_wait(&bars->kv_empty[slot], phase ^ 1);
              uint8_t* dst = smem_kv + slot * kTileBytes;
              const CUtensorMap* tm = staged ? (kv == 0 ? &tm_ks : &tm_vs) : (kv == 0 ? &tm_k : &tm_v);
              mbar_arrive_expect_tx(&bars->kv_full[slot], kTileBytes);
              tma_load_3d(dst, tm, &bars->kv_full[slot], 0, kv_head, row);
              tma_load_3d(dst + kHalfBytes, tm, &bars->kv_full[slot], 64, kv_head, row);
              if (++slot == kStages) {
                slot = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer (warp-uniform, one leader)
      const bool leader = elect_one();
      constexpr uint32_t idesc_qk = umma_idesc_f16(Pack2<T>::kFmt, kTile, kHalf, 0, 0);  // 128 x 64 scores
      constexpr uint32_t idesc_pv = umma_idesc_f16(Pack2<T>::kFmt, kTile, kD, 0, 1);
      const uint32_t q_base = smem_u32(smem_q);
      const uint32_t kv_base = smem_u32(smem_kv);
      const uint32_t col_s[2] = {tmem + kColS0, tmem + kColS1};
      const uint32_t col_o[2] = {tmem + kColO0, tmem + kColO1};
      constexpr uint32_t hi = umma_desc_hi(1024, kSwizzle128B);
      const uint32_t q_lo[2] = {umma_desc_lo(q_base, 16), umma_desc_lo(q_base + kTileBytes, 16)};
      const uint32_t k_lo0 = umma_desc_lo(kv_base, 16);
      const uint32_t v_lo0 = umma_desc_lo(kv_base, kHalfBytes);
      constexpr uint32_t slot_step = kTileBytes >> 4;
      constexpr uint32_t half_rows = (kHalf * 128) >> 4;  // 64 key rows of 128 bytes inside a 64-dim sub-tile

      // S_h(t) = Q_t K_h^T: B = key rows [64 h, 64 h + 64) of the K tile, K-major, N = 64
      auto issue_qk = [&](int t, uint32_t k_slot, int h) {
        const uint32_t a0 = q_lo[t], b0 = k_lo0 + k_slot * slot_step + h * half_rows;
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint32_t off = ((k >> 2) * kHalfBytes + (k & 3) * 32) >> 4;
          umma_ss2(col_s[t] + h * kHalf, a0 + off, hi, b0 + off, hi, idesc_qk, k > 0);
        }
        umma_commit(&bars->s_full[t][h]);
      };
      // O_t += P_h V_h: A = P (bf16, 32 columns at the start of S_h), B = key rows [64 h, +64) of V, MN-major
      auto issue_pv = [&](int t, uint32_t v_slot, int h, bool accumulate) {
        const uint32_t b0 = v_lo0 + v_slot * slot_step + h * half_rows;
#pragma unroll
        for (int k = 0; k < kHalf / 16; ++k)
          umma_ts2(col_o[t], col_s[t] + h * kHalf + k * 8, b0 + k * (2048 >> 4), hi, idesc_pv,
                   (accumulate || k > 0) ? 1u : 0u);
        umma_commit(&bars->pv_done[t]);
      };

      mbar_wait(&bars->q_full[0], 0);
      if (has_t1) mbar_wait(&bars->q_full[1], 0);
      tc_fence_after();

      uint32_t slot = 0, phase = 0;
      uint32_t p_phase[2][2] = {{0, 0}, {0, 0}};
      bool o_started[2] = {false, false};
      bool pend[2][2] = {{false, false}, {false, false}};  // P V of the previous key tile still to be issued
      uint32_t pend_slot = 0;      // V slot of the previous key tile
      bool pend_slot_live = false;  // ... and whether it still has to be released
      auto advance = [&]() {
        if (++slot == kStages) {
          slot = 0;
          phase ^= 1;
        }
      };
      // P V of half h of the previous key tile for q-tile t, if still owed
      auto flush_pv = [&](int t, int h) {
        if (!pend[t][h]) return;
        mbar_wait(&bars->p_ready[t][h], p_phase[t][h]);
        p_phase[t][h] ^= 1;
        tc_fence_after();
        if (leader) issue_pv(t, pend_slot, h, o_started[t]);
        o_started[t] = true;
        pend[t][h] = false;
      };
      for (int si = 0; si < it.seg_count; ++si) {
        const SegGeom g = seg_geom(p.segs[it.seg_begin + si], it);
        for (int jj = 0; jj < g.n_tiles; ++jj) {
          const bool act[2] = {tile_active(g, it, 0, jj), tile_active(g, it, 1, jj)};
          const uint32_t k_slot = slot, k_phase = phase;
          advance();
          const uint32_t v_slot = slot, v_phase = phase;
          advance();
          mbar_wait(&bars->kv_full[k_slot], k_phase);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              flush_pv(t, h);  // frees S_h(t) (in-order pipe: the Q K^T below runs after this P V)
              if (act[t] && leader) issue_qk(t, k_slot, h);
            }
          }
          if (leader) {
            umma_commit(&bars->kv_empty[k_slot]);
            if (pend_slot_live) umma_commit(&bars->kv_empty[pend_slot]);  // every P V of the previous tile is issued
          }
          pend_slot_live = false;
          mbar_wait(&bars->kv_full[v_slot], v_phase);
          tc_fence_after();
          if (act[0] || act[1]) {
            pend[0][0] = pend[0][1] = act[0];
            pend[1][0] = pend[1][1] = act[1];
            pend_slot = v_slot;
            pend_slot_live = true;
          } else if (leader) {
            umma_commit(&bars->kv_empty[v_slot]);
          }
          __syncwarp();
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int t = 0; t < 2; ++t) flush_pv(t, h);
      }
      if (leader) {
        if (pend_slot_live) umma_commit(&bars->kv_empty[pend_slot]);
        umma_commit(&bars->o_done[0]);
        umma_commit(&bars->o_done[1]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue
    reg_alloc<208>();
    const int t = (warp - 4) >> 2;
    const int row_in_tile = ((warp & 3) << 5) | lane;
    const int n_rows = t == 0 ? n_rows0 : n_rows1;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem + (t == 0 ? kColS0 : kColS1) + lane_addr;
    const uint32_t t_o = tmem + (t == 0 ? kColO0 : kColO1) + lane_addr;
    const int chunk_row = it.q_off + t * kTile + row_in_tile;

    float m_ref = -CUDART_INF_F;
    float l = 0.f;
    uint32_t s_phase0 = 0, s_phase1 = 0;
    int n_pv = 0;  // halves handed to the MMA warp so far == P V GEMMs of this q-tile issued or about to be

    // The exp sections of the two warpgroups take turns (see attn_fwd_sm100.cu), one hand-off per 64-key step.
    // With the next scores always ready the alternation may no longer pay: FwdParams::flags bit 0 disables it.
    const bool turns = has_t1 && !(p.flags & 1);
    int handoffs_left = 0;
    if (turns) {
      for (int si = 0; si < it.seg_count; ++si)
        handoffs_left += 2 * seg_geom(p.segs[it.seg_begin + si], it).n_tiles;
      if (t == 1 && handoffs_left > 0) named_bar_arrive(1, 256);
    }
    auto turn_wait = [&]() {
      if (turns) named_bar_sync(1 + t, 256);
    };
    auto turn_pass = [&]() {
      if (turns) {
        --handoffs_left;
        if (!(t == 1 && handoffs_left == 0)) named_bar_arrive(1 + (1 - t), 256);
      }
    };

    if (n_rows > 0) {
      for (int si = 0; si < it.seg_count; ++si) {
        const SegGeom g = seg_geom(p.segs[it.seg_begin + si], it);
        for (int jj = 0; jj < g.n_tiles; ++jj) {
          if (!tile_active(g, it, t, jj)) {
            turn_wait();
            turn_pass();
            turn_wait();
            turn_pass();
            continue;
          }
          const bool masked = tile_needs_mask(g, it, t, jj);
          long long lim_ll = static_cast<long long>(chunk_row) + g.diag;
          if (lim_ll > g.kv_len - 1) lim_ll = g.kv_len - 1;
          lim_ll -= static_cast<long long>(jj) * kTile;
          const int lim128 = lim_ll < -1 ? -1 : (lim_ll > 127 ? 127 : static_cast<int>(lim_ll));
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const uint32_t t_sh = t_s + h * kHalf;
            mbar_wait(&bars->s_full[t][h], h == 0 ? s_phase0 : s_phase1);
            if (h == 0) s_phase0 ^= 1;
            else s_phase1 ^= 1;
            tc_fence_after();
            uint32_t sr[64];
            tmem_ld32(t_sh + 0, sr + 0);
            tmem_ld32(t_sh + 32, sr + 32);
            tmem_ld_wait();
            float s[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) s[c] = __uint_as_float(sr[c]);
            if (masked) {
              const int lim = lim128 - h * kHalf;  // last visible column of this half (may be < 0 or > 63)
#pragma unroll
              for (int c = 0; c < 64; ++c) s[c] = c <= lim ? s[c] : -CUDART_INF_F;
            }
            float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
            for (int c = 4; c < 64; c += 4) {
              mx0 = fmaxf(mx0, s[c]);
              mx1 = fmaxf(mx1, s[c + 1]);
              mx2 = fmaxf(mx2, s[c + 2]);
              mx3 = fmaxf(mx3, s[c + 3]);
            }
            const float m_new = fmaxf(fmaxf(m_ref, fmaxf(mx0, mx1)), fmaxf(mx2, mx3));
            const bool need = (m_new - m_ref) * p.scale_log2 > kRescaleThreshold;
            if (__any_sync(0xffffffffu, need)) {
              const float f = need ? fast_exp2((m_ref - m_new) * p.scale_log2) : 1.0f;
              if (need) {
                l *= f;
                m_ref = m_new;
              }
              if (n_pv > 0) {
                // O must be quiescent: wait for the most recent P V of this q-tile.  S_full of the current half
                // already implies that the one before it has completed, so the barrier is at most one phase
                // behind and the parity wait cannot alias.
                mbar_wait(&bars->pv_done[t], static_cast<uint32_t>(n_pv - 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 128; c += 32) {
                  uint32_t orr[32];
                  tmem_ld32(t_o + c, orr);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
                  tmem_st32(t_o + c, orr);
                }
              }
            }
            const float mc = (m_ref == -CUDART_INF_F ? 0.f : m_ref) * p.scale_log2;
            turn_wait();
            const uint64_t sc2 = pack2(p.scale_log2, p.scale_log2), nmc2 = pack2(-mc, -mc);
            uint64_t lsum = pack2(0.f, 0.f);
            auto exp_chunks = [&](auto use_poly) {
#pragma unroll
              for (int c = 0; c < 64; c += 32) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint64_t x = ffma2(pack2(s[c + 2 * i], s[c + 2 * i + 1]), sc2, nmc2);
                  float e0, e1;
                  if (decltype(use_poly)::value && (i & 3) < kPolyOf4) {
                    exp2_poly2(x, e0, e1);
                  } else {
                    float x0, x1;
                    unpack2(x, x0, x1);
                    e0 = fast_exp2(x0);
                    e1 = fast_exp2(x1);
                  }
                  lsum = fadd2(lsum, pack2(e0, e1));
                  pk[i] = Pack2<T>::pack(e0, e1);
                }
                tmem_st16(t_sh + (c >> 1), pk);  // P_h: 32 columns at the start of S_h
              }
            };
            if (masked) {  // masked steps stay on MUFU so that masked entries are exact zeros
              exp_chunks(std::false_type{});
            } else {
              exp_chunks(std::true_type{});
            }
            float l0, l1;
            unpack2(lsum, l0, l1);
            turn_pass();
            l += l0 + l1;
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->p_ready[t][h]);
            ++n_pv;
          }
        }
      }

      // ---------------------------------------------------------------- epilogue: O / l -> out, lse
      const int row = it.q_row0 + t * kTile + row_in_tile;
      const bool row_ok = row_in_tile < n_rows;
      T* out_row = reinterpret_cast<T*>(p.out) + (static_cast<size_t>(row) * p.hq + head) * kD;
      if (n_pv > 0) {
        mbar_wait(&bars->o_done[t], 0);
        tc_fence_after();
        const float inv = l > 0.f ? 1.0f / l : 0.f;
#pragma unroll
        for (int c = 0; c < 128; c += 32) {
          uint32_t orr[32];
          tmem_ld32(t_o + c, orr);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 v;
              v.x = Pack2<T>::pack(__uint_as_float(orr[i + 0]) * inv, __uint_as_float(orr[i + 1]) * inv);
              v.y = Pack2<T>::pack(__uint_as_float(orr[i + 2]) * inv, __uint_as_float(orr[i + 3]) * inv);
              v.z = Pack2<T>::pack(__uint_as_float(orr[i + 4]) * inv, __uint_as_float(orr[i + 5]) * inv);
              v.w = Pack2<T>::pack(__uint_as_float(orr[i + 6]) * inv, __uint_as_float(orr[i + 7]) * inv);
              *reinterpret_cast<uint4*>(out_row + c + i) = v;
            }
          }
        }
      } else if (row_ok) {
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < 128; c += 8) *reinterpret_cast<uint4*>(out_row + c) = z;
      }
      if (row_ok) {
        const float lse = l > 0.f ? m_ref * p.scale + __logf(l) : -CUDART_INF_F;
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

}  // namespace fwd64

const char* attn_fwd_h64_launch(int dtype, const TensorView& q, const TensorView& k, const TensorView& v,
                                const TensorView& k_stage, const TensorView& v_stage, const FwdParams& p,
                                cudaStream_t stream) {
  const int n_blocks = p.push.n_ctas + p.n_items * p.hq;
  if (n_blocks <= 0) return nullptr;
  if (dtype != kDtypeBF16 && dtype != kDtypeFP16) return "the h64 forward supports bf16 / fp16 only";
  if (p.seg_lo != nullptr) return "the h64 forward does not support sliding windows";
  CUtensorMap tq, tk, tv, tks, tvs;
  if (const char* e = make_tensor_map(&tq, q, 2, fwd64::kTile, fwd64::kD)) return e;
  if (const char* e = make_tensor_map(&tk, k, 2, fwd64::kTile, fwd64::kD)) return e;
  if (const char* e = make_tensor_map(&tv, v, 2, fwd64::kTile, fwd64::kD)) return e;
  if (const char* e = make_tensor_map(&tks, k_stage, 2, fwd64::kTile, fwd64::kD)) return e;
  if (const char* e = make_tensor_map(&tvs, v_stage, 2, fwd64::kTile, fwd64::kD)) return e;
  dim3 grid(n_blocks, 1, 1), block(fwd64::kThreads, 1, 1);
  cudaError_t err = cudaSuccess;
  auto launch = [&](auto kern) {
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd64::kSmemBytes);
    if (err == cudaSuccess) kern<<<grid, block, fwd64::kSmemBytes, stream>>>(tq, tk, tv, tks, tvs, p);
  };
  static const int poly = [] {
    const char* e = std::getenv("RFA_B200_POLY_EXP");
    const int v = e ? std::atoi(e) : 0;
    return v < 0 ? 0 : (v > 2 ? 2 : v);
  }();
  if (dtype == kDtypeBF16) {
    if (poly == 0) launch(fwd64::attn_fwd_h64_kernel<__nv_bfloat16, 0>);
    else if (poly == 1) launch(fwd64::attn_fwd_h64_kernel<__nv_bfloat16, 1>);
    else launch(fwd64::attn_fwd_h64_kernel<__nv_bfloat16, 2>);
  } else {
    if (poly == 0) launch(fwd64::attn_fwd_h64_kernel<__half, 0>);
    else if (poly == 1) launch(fwd64::attn_fwd_h64_kernel<__half, 1>);
    else launch(fwd64::attn_fwd_h64_kernel<__half, 2>);
  }
  if (err != cudaSuccess) return cudaGetErrorString(err);
  err = cudaGetLastError();
  return err == cudaSuccess ? nullptr : cudaGetErrorString(err);
}

}  // namespace rfa
