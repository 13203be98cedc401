// Host/device shared structures for the sm_100a attention kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rfa {

constexpr int kDiagFull = 1 << 29;  // "no causal boundary": every key of the segment is visible

// One CTA work item: up to 256 consecutive query rows of one query chunk (two 128-row MMA tiles).
struct alignas(16) WorkItem {
  int q_row0;     // first local query row
  int q_rows;     // 1..256
  int q_off;      // offset of q_row0 inside its chunk (the diagonal is expressed in chunk rows)
  int seg_begin;  // first entry in the segment table
  int seg_count;
  int pad0, pad1, pad2;
};

// A run of keys a chunk may see.  key j visible to chunk-row i  iff  j <= i + diag  and  j < kv_len.
struct alignas(16) KVSegment {
  int kv_row0;  // row inside the K/V tensor handed to this launch
  int kv_len;
  int diag;
  int flag;  // fused multi-GPU mode: index of the "rows have landed" flag to wait on, -1 = local data
};

// Tensor view handed to the launchers: (rows, heads, 128) with element strides.
struct TensorView {
  void* ptr;
  int64_t rows;
  int heads;
  int64_t row_stride;
  int64_t head_stride;
};

struct FwdParams {
  const WorkItem* items;
  const KVSegment* segs;
  void* out;  // (rows, hq, 128) contiguous, input dtype
  float* lse;  // index = (row / lse_S) * hq * lse_S + head * lse_S + row % lse_S
  int lse_S;
  int hq, hkv;
  float scale;       // softmax scale
  float scale_log2;  // scale * log2(e)
  const uint32_t* ready_flags;  // fused mode only
  uint32_t ready_epoch;
};

// Backward work item: one tile of <= 128 keys (exclusive owner of those dK/dV rows in this launch).
struct alignas(16) BwdItem {
  int kv_row0;    // first row inside the K/V tensors handed to this launch
  int kv_rows;    // 1..128
  int seg_begin;  // first entry in the query-segment table
  int seg_count;
  int flag;  // fused mode: index of the "keys have landed" flag, -1 = local data
  int pad0, pad1, pad2;
};
// A query chunk that can see the key tile.  tile key j visible to chunk row i  iff  j <= i + diag.
struct alignas(16) BwdQSegment {
  int q_row0;  // first local row of the chunk
  int q_len;
  int diag;  // already relative to the key tile's first key
  int pad;
};

struct BwdParams {
  const BwdItem* items;
  const BwdQSegment* qsegs;
  const float* lse;    // same indexing as FwdParams::lse
  const float* delta;  // rowsum(out * dout), same indexing
  float* dk;           // (kv rows, hkv, 128) fp32, rows of each item are written (not accumulated)
  float* dv;
  int lse_S;
  int hq, hkv;
  float scale, scale_log2;
  const uint32_t* ready_flags;
  uint32_t ready_epoch;
};

// Descriptor probe (csrc/probe_sm100.cu): operand forms + optional run-time overrides (-1 = kernel default).
struct ProbeConfig {
  int a_kind, b_kind;
  int n;     // 128 or 64
  int kdim;  // 128 or 64
  int lbo_a, sbo_a, kstep_a;
  int lbo_b, sbo_b, kstep_b;
};

// dtype codes shared with the Python side
enum : int { kDtypeBF16 = 0, kDtypeFP16 = 1 };

// ---- launchers (implemented in the .cu files; plain C++ so bindings.cpp needs no CUDA headers beyond runtime)
const char* attn_fwd_launch(int dtype, const TensorView& q, const TensorView& k, const TensorView& v,
                            const FwdParams& p, int n_items, cudaStream_t stream);

const char* attn_bwd_delta_launch(int dtype, const TensorView& out, const TensorView& dout, float* delta, int lse_S,
                                  cudaStream_t stream);
const char* attn_bwd_launch(int dtype, const TensorView& q, const TensorView& dout, const TensorView& k,
                            const TensorView& v, const TensorView& dq_accum, const BwdParams& p, int n_items,
                            cudaStream_t stream);
const char* probe_launch(const TensorView& a, const TensorView& b, const void* a_raw, const void* b_raw, float* out,
                         const ProbeConfig& c, cudaStream_t stream);
const char* lse_flatten_launch(const float* in, float* out, const int* cu, int batch, int heads, int max_seqlen,
                               int total, cudaStream_t stream);
const char* lse_unflatten_launch(const float* in, float* out, const int* cu, int batch, int heads, int max_seqlen,
                                 cudaStream_t stream);
// Non-swizzled map with a (dim_inner, 1, box_rows) box (used for the fp32 dQ reduce-add).
const char* make_plain_tensor_map(CUtensorMap* out, const TensorView& t, int elem_bytes, int box_rows, int dim_inner);

// Encode a (rows, heads, 128) tensor as a 3-D TMA map with a (64, 1, box_rows) box and 128-byte swizzle.
// elem_bytes is 2 (bf16/fp16) or 4 (fp32; then the box is 32 elements wide).
const char* make_tensor_map(CUtensorMap* out, const TensorView& t, int elem_bytes, int box_rows, int dim_inner);

}  // namespace rfa
