// Python bindings for the sm_100a kernels and the peer-memory runtime (module ring_flash_attn_b200._C).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "attn_common.h"
#include "peer_mem.h"

namespace {

using rfa::TensorView;

int dtype_code(const at::Tensor& t) {
  if (t.scalar_type() == at::kBFloat16) return rfa::kDtypeBF16;
  if (t.scalar_type() == at::kHalf) return rfa::kDtypeFP16;
  TORCH_CHECK(false, "ring_flash_attn_b200: only bf16 / fp16 inputs are supported by the sm_100a kernels");
}

TensorView view3(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.dim() == 3, name, " must be (rows, heads, head_dim)");
  TORCH_CHECK(t.size(2) == 128, name, ": the sm_100a kernels support head_dim == 128 only");
  TORCH_CHECK(t.stride(2) == 1, name, ": last dimension must be contiguous");
  return TensorView{t.data_ptr(), t.size(0), static_cast<int>(t.size(1)), t.stride(0), t.stride(1)};
}

void check(const char* err) { TORCH_CHECK(err == nullptr, "ring_flash_attn_b200 kernel launch failed: ", err); }

const uint32_t* flag_ptr(const c10::optional<at::Tensor>& flags) {
  return flags.has_value() ? reinterpret_cast<const uint32_t*>(flags->data_ptr()) : nullptr;
}

void attn_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& items,
              const at::Tensor& segs, at::Tensor& out, at::Tensor& lse, int64_t lse_S, double scale,
              const c10::optional<at::Tensor>& ready_flags, int64_t ready_epoch) {
  const c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(items.scalar_type() == at::kInt && items.is_cuda() && items.is_contiguous() && items.size(1) == 8);
  TORCH_CHECK(segs.scalar_type() == at::kInt && segs.is_cuda() && segs.is_contiguous() && segs.size(1) == 4);
  TORCH_CHECK(out.is_contiguous() && out.scalar_type() == q.scalar_type());
  TORCH_CHECK(lse.is_contiguous() && lse.scalar_type() == at::kFloat);
  TORCH_CHECK(k.scalar_type() == q.scalar_type() && v.scalar_type() == q.scalar_type());
  TORCH_CHECK(q.size(1) % k.size(1) == 0, "query heads must be a multiple of kv heads");
  rfa::FwdParams p{};
  p.items = reinterpret_cast<const rfa::WorkItem*>(items.data_ptr());
  p.segs = reinterpret_cast<const rfa::KVSegment*>(segs.data_ptr());
  p.out = out.data_ptr();
  p.lse = lse.data_ptr<float>();
  p.lse_S = static_cast<int>(lse_S);
  p.hq = static_cast<int>(q.size(1));
  p.hkv = static_cast<int>(k.size(1));
  p.scale = static_cast<float>(scale);
  p.scale_log2 = static_cast<float>(scale * 1.4426950408889634);
  p.ready_flags = flag_ptr(ready_flags);
  p.ready_epoch = static_cast<uint32_t>(ready_epoch);
  check(rfa::attn_fwd_launch(dtype_code(q), view3(q, "q"), view3(k, "k"), view3(v, "v"), p,
                             static_cast<int>(items.size(0)), at::cuda::getCurrentCUDAStream()));
}

void attn_bwd_delta(const at::Tensor& out, const at::Tensor& dout, at::Tensor& delta, int64_t lse_S) {
  const c10::cuda::CUDAGuard guard(out.device());
  TORCH_CHECK(delta.is_contiguous() && delta.scalar_type() == at::kFloat);
  check(rfa::attn_bwd_delta_launch(dtype_code(out), view3(out, "out"), view3(dout, "dout"), delta.data_ptr<float>(),
                                   static_cast<int>(lse_S), at::cuda::getCurrentCUDAStream()));
}

void attn_bwd(const at::Tensor& q, const at::Tensor& dout, const at::Tensor& k, const at::Tensor& v,
              at::Tensor& dq_accum, const at::Tensor& items, const at::Tensor& qsegs, const at::Tensor& lse,
              const at::Tensor& delta, at::Tensor& dk, at::Tensor& dv, int64_t lse_S, double scale,
              const c10::optional<at::Tensor>& ready_flags, int64_t ready_epoch) {
  const c10::cuda::CUDAGuard guard(q.device());
  TORCH_CHECK(items.scalar_type() == at::kInt && items.is_cuda() && items.is_contiguous() && items.size(1) == 8);
  TORCH_CHECK(qsegs.scalar_type() == at::kInt && qsegs.is_cuda() && qsegs.is_contiguous() && qsegs.size(1) == 4);
  TORCH_CHECK(dq_accum.scalar_type() == at::kFloat && dq_accum.dim() == 3 && dq_accum.size(2) == 128 &&
              dq_accum.stride(2) == 1);
  TORCH_CHECK(dk.scalar_type() == at::kFloat && dk.is_contiguous() && dv.scalar_type() == at::kFloat &&
              dv.is_contiguous());
  TORCH_CHECK(lse.scalar_type() == at::kFloat && delta.scalar_type() == at::kFloat);
  rfa::BwdParams p{};
  p.items = reinterpret_cast<const rfa::BwdItem*>(items.data_ptr());
  p.qsegs = reinterpret_cast<const rfa::BwdQSegment*>(qsegs.data_ptr());
  p.lse = lse.data_ptr<float>();
  p.delta = delta.data_ptr<float>();
  p.dk = dk.data_ptr<float>();
  p.dv = dv.data_ptr<float>();
  p.lse_S = static_cast<int>(lse_S);
  p.hq = static_cast<int>(q.size(1));
  p.hkv = static_cast<int>(k.size(1));
  p.scale = static_cast<float>(scale);
  p.scale_log2 = static_cast<float>(scale * 1.4426950408889634);
  p.ready_flags = flag_ptr(ready_flags);
  p.ready_epoch = static_cast<uint32_t>(ready_epoch);
  TensorView dqv{dq_accum.data_ptr(), dq_accum.size(0), static_cast<int>(dq_accum.size(1)), dq_accum.stride(0),
                 dq_accum.stride(1)};
  check(rfa::attn_bwd_launch(dtype_code(q), view3(q, "q"), view3(dout, "dout"), view3(k, "k"), view3(v, "v"), dqv, p,
                             static_cast<int>(items.size(0)), at::cuda::getCurrentCUDAStream()));
}

at::Tensor probe(const at::Tensor& a, const at::Tensor& b, std::vector<int64_t> cfg) {
  // a: (rows, 128) or (128, kdim) bf16 ; b likewise; cfg = {a_kind, b_kind, n, kdim, lbo_a, sbo_a, kstep_a, lbo_b, sbo_b, kstep_b}
  const c10::cuda::CUDAGuard guard(a.device());
  TORCH_CHECK(cfg.size() == 10);
  rfa::ProbeConfig c{static_cast<int>(cfg[0]), static_cast<int>(cfg[1]), static_cast<int>(cfg[2]),
                     static_cast<int>(cfg[3]), static_cast<int>(cfg[4]), static_cast<int>(cfg[5]),
                     static_cast<int>(cfg[6]), static_cast<int>(cfg[7]), static_cast<int>(cfg[8]),
                     static_cast<int>(cfg[9])};
  TORCH_CHECK(a.is_contiguous() && b.is_contiguous() && a.scalar_type() == at::kBFloat16 &&
              b.scalar_type() == at::kBFloat16);
  TensorView va{nullptr, 0, 1, 128, 128}, vb{nullptr, 0, 1, 128, 128};
  if (c.a_kind == 0 || c.a_kind == 3) va = TensorView{a.data_ptr(), a.size(0), 1, 128, 128};
  if (c.b_kind == 0 || c.b_kind == 1) vb = TensorView{b.data_ptr(), b.size(0), 1, 128, 128};
  at::Tensor out = at::zeros({128, c.n}, a.options().dtype(at::kFloat));
  check(rfa::probe_launch(va, vb, a.data_ptr(), b.data_ptr(), out.data_ptr<float>(), c,
                          at::cuda::getCurrentCUDAStream()));
  return out;
}

at::Tensor lse_flatten(const at::Tensor& lse, const at::Tensor& cu) {
  const c10::cuda::CUDAGuard guard(lse.device());
  TORCH_CHECK(lse.dim() == 3 && lse.scalar_type() == at::kFloat && lse.is_contiguous());
  TORCH_CHECK(cu.scalar_type() == at::kInt && cu.is_cuda());
  const int batch = static_cast<int>(cu.numel()) - 1;
  const int64_t total = cu[batch].item<int>();
  at::Tensor out = at::empty({lse.size(1), total}, lse.options());
  check(rfa::lse_flatten_launch(lse.data_ptr<float>(), out.data_ptr<float>(), cu.data_ptr<int>(), batch,
                                static_cast<int>(lse.size(1)), static_cast<int>(lse.size(2)), static_cast<int>(total),
                                at::cuda::getCurrentCUDAStream()));
  return out;
}

at::Tensor lse_unflatten(const at::Tensor& lse, const at::Tensor& cu, int64_t max_seqlen) {
  const c10::cuda::CUDAGuard guard(lse.device());
  TORCH_CHECK(lse.dim() == 3 && lse.size(2) == 1 && lse.scalar_type() == at::kFloat && lse.is_contiguous());
  TORCH_CHECK(cu.scalar_type() == at::kInt && cu.is_cuda());
  const int batch = static_cast<int>(cu.numel()) - 1;
  at::Tensor out = at::empty({batch, lse.size(1), max_seqlen}, lse.options());
  check(rfa::lse_unflatten_launch(lse.data_ptr<float>(), out.data_ptr<float>(), cu.data_ptr<int>(), batch,
                                  static_cast<int>(lse.size(1)), static_cast<int>(max_seqlen),
                                  at::cuda::getCurrentCUDAStream()));
  return out;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "ring_flash_attn_b200 sm_100a kernels";
  m.def("attn_fwd", &attn_fwd);
  m.def("attn_bwd_delta", &attn_bwd_delta);
  m.def("attn_bwd", &attn_bwd);
  m.def("probe", &probe);
  m.def("lse_flatten", &lse_flatten);
  m.def("lse_unflatten", &lse_unflatten);
  rfa::bind_peer_mem(m);
}
