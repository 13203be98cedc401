"""torch.compile around the public functions (the reference's tests have a `compile` mode, test/test.sh:23-25).

The context-parallel op is a ``torch.library`` custom op with a fake implementation and an autograd formula
(parallel/ops.py), so the public functions trace with ``fullgraph=True``: no graph break, for every scheme."""
import pytest
import torch

import ring_flash_attn_b200 as rfa
from ring_flash_attn_b200.ops.dense import attention_oracle, varlen_attention_oracle


def _compile(f):
    # aot_eager: dynamo + AOT autograd (fake tensors through forward AND backward) without inductor's C++ toolchain
    return torch.compile(f, backend="aot_eager", fullgraph=True)


@pytest.mark.parametrize("name", ["ring", "zigzag_ring", "stripe"])
def test_batch_schemes_fullgraph(name):
    torch.manual_seed(0)
    qkv = torch.randn(1, 32, 3, 2, 16, requires_grad=True)
    fn = getattr(rfa, f"{name}_flash_attn_qkvpacked_func")

    def f(x):
        return fn(x * 1.0, causal=True) * 2.0

    ref, _ = attention_oracle(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], True)
    out = _compile(f)(qkv)
    torch.testing.assert_close(out, 2.0 * ref, atol=2e-5, rtol=2e-4)
    out.sum().backward()
    g = qkv.grad.clone()
    qkv.grad = None
    f(qkv).sum().backward()
    torch.testing.assert_close(g, qkv.grad, atol=1e-5, rtol=1e-5)


def test_varlen_and_llama3_fullgraph():
    torch.manual_seed(0)
    T, H, D = 48, 2, 16
    cu = torch.tensor([0, 10, 30, T], dtype=torch.int32)
    q, k, v = (torch.randn(T, H, D, requires_grad=True) for _ in range(3))
    ref, ref_lse = varlen_attention_oracle(q, k, v, cu, True)

    def f_varlen(q, k, v, cu):
        out, lse, _ = rfa.zigzag_ring_flash_attn_varlen_func(q, k, v, cu, 20, causal=True, return_attn_probs=True)
        return out + 0.0, lse

    out, lse = _compile(f_varlen)(q, k, v, cu)
    torch.testing.assert_close(out, ref, atol=2e-5, rtol=2e-4)
    torch.testing.assert_close(lse, ref_lse, atol=1e-4, rtol=1e-4)

    cq, ck, mq, mk, ks = rfa.llama3_flash_attn_prepare_cu_seqlens(cu, True, 0, 1)

    def f_llama3(q, k, v, cq, ck):
        return rfa.llama3_flash_attn_varlen_func(q, k, v, cq, ck, mq, mk, heads_k_stride=1, local_k_slice=ks,
                                                 causal=True) * 1.0

    out = _compile(f_llama3)(q, k, v, cq, ck)
    torch.testing.assert_close(out, ref, atol=2e-5, rtol=2e-4)
    out.sum().backward()
    assert q.grad is not None and torch.isfinite(q.grad).all()


def test_op_is_registered_with_fake_and_autograd():
    op = torch.ops.rfa_b200.cp_attn_fwd.default
    q = torch.randn(8, 2, 16)
    torch.library.opcheck(op, (q, q.clone(), q.clone(), None, None, None, None, None, "ring", "", [1, 8, 1, -1, -1],
                               0.25, False), test_utils=("test_schema", "test_faketensor"))
