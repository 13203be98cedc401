"""Reference-compatible utility surface: merge, LSE layout converters, comm guards, get_default_args."""
import pytest
import torch

from ring_flash_attn_b200 import utils
from ring_flash_attn_b200.ops import lse_layout
from ring_flash_attn_b200.ops.dense import block_fwd
from ring_flash_attn_b200.ops.merge import merge_partial, update_out_and_lse
from ring_flash_attn_b200.parallel.comm import RingComm


def test_update_out_and_lse_matches_exact_logaddexp_and_reference_formula():
    torch.manual_seed(0)
    b, s, h, d = 2, 16, 3, 8
    out = lse = None
    outs, lses = [], []
    for _ in range(4):
        bo, bl = torch.randn(b, s, h, d), torch.randn(b, h, s)
        outs.append(bo)
        lses.append(bl)
        out, lse = update_out_and_lse(out, lse, bo, bl)
    stack_l = torch.stack(lses)  # (n,b,h,s)
    tot = torch.logsumexp(stack_l, dim=0)
    w = torch.exp(stack_l - tot).permute(0, 1, 3, 2).unsqueeze(-1)
    want = (torch.stack(outs) * w).sum(0)
    torch.testing.assert_close(lse.squeeze(-1).transpose(1, 2), tot, atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(out, want, atol=1e-6, rtol=1e-5)
    # the reference's sigmoid / logsigmoid form (utils.py:32-50) is algebraically the same update
    o2 = outs[0] - torch.sigmoid(lses[1].transpose(1, 2).unsqueeze(-1) - lses[0].transpose(1, 2).unsqueeze(-1)) * (outs[0] - outs[1])
    o, l = update_out_and_lse(None, None, outs[0], lses[0])
    o, l = update_out_and_lse(o, l, outs[1], lses[1])
    torch.testing.assert_close(o, o2, atol=1e-6, rtol=1e-5)


def test_update_out_and_lse_slice_and_guard():
    out, lse = update_out_and_lse(None, None, torch.zeros(1, 8, 2, 4), torch.zeros(1, 2, 8))
    sl = (slice(None), slice(4, None))
    out, lse = update_out_and_lse(out, lse, torch.ones(1, 4, 2, 4), torch.zeros(1, 2, 4), slice_=sl)
    assert torch.allclose(out[:, 4:], torch.full((1, 4, 2, 4), 0.5)) and torch.allclose(out[:, :4], torch.zeros(1, 4, 2, 4))
    with pytest.raises(RuntimeError):
        update_out_and_lse(None, None, torch.zeros(1, 4, 2, 4), torch.zeros(1, 2, 4), slice_=sl)


def test_merge_partial_blocks_equal_one_shot_attention():
    torch.manual_seed(0)
    q, k, v = torch.randn(10, 2, 8), torch.randn(24, 2, 8), torch.randn(24, 2, 8)
    ref_o, ref_l = block_fwd(q, k, v, 0.3, None)
    out = lse = None
    for a in range(0, 24, 8):
        o, l = block_fwd(q, k[a:a + 8], v[a:a + 8], 0.3, None)
        out, lse = merge_partial(out, lse, o, l)
    torch.testing.assert_close(out, ref_o, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(lse, ref_l, atol=1e-5, rtol=1e-5)
    # rows that have seen nothing (lse = -inf) are the identity of the merge
    empty_o, empty_l = torch.zeros(10, 2, 8), torch.full((2, 10), float("-inf"))
    o2, l2 = merge_partial(empty_o.clone(), empty_l.clone(), ref_o, ref_l)
    torch.testing.assert_close(o2, ref_o)
    torch.testing.assert_close(l2, ref_l)


def test_lse_layout_roundtrip_cpu():
    cu = torch.tensor([0, 3, 10, 12], dtype=torch.int32)
    padded = torch.randn(3, 4, 7)
    flat = lse_layout.flatten_varlen_lse(padded, cu)
    assert flat.shape == (4, 12)
    back = lse_layout.unflatten_varlen_lse(flat.transpose(0, 1).unsqueeze(-1).contiguous(), cu, 7)
    for b, (a, e) in enumerate([(0, 3), (3, 10), (10, 12)]):
        assert torch.equal(back[b, :, : e - a], padded[b, :, : e - a])


def test_ringcomm_protocol_guards_single_process():
    comm = RingComm(None)
    x = torch.arange(6.0)
    y = comm.send_recv(x)
    assert torch.equal(y, x)  # world size 1: the ring is a self loop
    comm.commit()
    with pytest.raises(RuntimeError):
        comm.commit()
    comm.wait()
    with pytest.raises(RuntimeError):
        comm.wait()


def test_get_default_args_returns_fresh_dict():
    def f(a, b=2, softcap=5.0, *, c=None):
        return a

    d = utils.get_default_args(f)
    assert d == {"b": 2, "softcap": 0.0, "c": None}
    d["b"] = 7
    assert utils.get_default_args(f)["b"] == 2
    assert set(utils.__all__) >= {"update_out_and_lse", "RingComm", "AllGatherComm", "flatten_varlen_lse",
                                  "unflatten_varlen_lse", "get_default_args"}


def test_fused_heads_per_pass_budget(monkeypatch):
    """llama3 on the fused path: heads_k_stride is the granularity, RFA_B200_STAGE_BUDGET_MB the cap of one launch."""
    import torch

    from ring_flash_attn_b200.ops.plan import CPPlan
    from ring_flash_attn_b200.parallel import engine

    plan = CPPlan(world=8, rank=0, q_rows=8192, kv_rows=8192)
    k = torch.empty(8192, 8, 128, dtype=torch.bfloat16)
    per_head_mb = 2 * 2 * 8 * 8192 * 128 * 2 / 2 ** 20  # 64 MiB of staging per kv head
    assert per_head_mb == 64
    monkeypatch.delenv("RFA_B200_LLAMA3_HEAD_GROUPS", raising=False)
    monkeypatch.setenv("RFA_B200_STAGE_BUDGET_MB", "8192")
    assert engine.fused_heads_per_pass(plan, k, 1) == 8  # everything fits: one launch over all heads
    monkeypatch.setenv("RFA_B200_STAGE_BUDGET_MB", "300")
    assert engine.fused_heads_per_pass(plan, k, 1) == 4  # largest divisor of 8 whose staging fits 300 MiB
    assert engine.fused_heads_per_pass(plan, k, 4) == 4
    monkeypatch.setenv("RFA_B200_STAGE_BUDGET_MB", "1")
    assert engine.fused_heads_per_pass(plan, k, 2) == 2  # never below the caller's stride
    monkeypatch.setenv("RFA_B200_LLAMA3_HEAD_GROUPS", "strict")
    monkeypatch.setenv("RFA_B200_STAGE_BUDGET_MB", "8192")
    assert engine.fused_heads_per_pass(plan, k, 2) == 2  # the reference's memory behaviour
    import pytest

    with pytest.raises(ValueError):
        engine.fused_heads_per_pass(plan, k, 3)


def test_ring_schemes_head_groups_under_budget(monkeypatch):
    """zigzag / ring / stripe: all heads in one fused launch unless their staging exceeds the budget; then groups of
    kv heads (granularity 1), never governed by the llama3-only strict switch."""
    import torch

    from ring_flash_attn_b200.ops.plan import CPPlan
    from ring_flash_attn_b200.parallel import engine

    plan = CPPlan(world=8, rank=0, q_rows=8192, kv_rows=8192)
    k = torch.empty(8192, 8, 128, dtype=torch.bfloat16)  # 64 MiB of staging per kv head
    monkeypatch.setenv("RFA_B200_LLAMA3_HEAD_GROUPS", "strict")
    monkeypatch.setenv("RFA_B200_STAGE_BUDGET_MB", "8192")
    assert engine._fused_by_head_groups(plan, k, 1, "ring") is None
    assert engine._fused_by_head_groups(plan, k, 1, "allgather") == 1
    monkeypatch.setenv("RFA_B200_STAGE_BUDGET_MB", "200")
    assert engine._fused_by_head_groups(plan, k, 1, "ring") == 2
    monkeypatch.setenv("RFA_B200_STAGE_BUDGET_MB", "0")
    assert engine._fused_by_head_groups(plan, k, 1, "ring") == 1
    assert engine._fused_by_head_groups(CPPlan(world=1, rank=0, q_rows=64, kv_rows=64), k, 1, "ring") is None


def test_fp8_scale_tables_follow_rows():
    """Fp8Scales: per-source views and head slices index the gathered tables like the staging buffer."""
    import torch

    from ring_flash_attn_b200.ops.attn_cuda import Fp8Scales

    q = torch.arange(8.0).view(4, 2)            # 4 query blocks x 2 heads
    k = torch.arange(6.0).view(3, 2) + 1.0      # 3 key blocks of 128 rows x 2 kv heads
    v = torch.tensor([[1.0, 5.0], [2.0, 4.0], [3.0, 3.0]])
    sc = Fp8Scales(q, 96, k, v, 128)
    assert sc.v_ref.tolist() == [3.0, 5.0] and sc.kv_row0 == 0 and sc.world_rows == 0
    assert sc.gathered(None, 0, 1, 384) is sc    # world 1: nothing to gather
    allr = Fp8Scales(q, 96, torch.cat([k, k + 10]), torch.cat([v, v]), 128, sc.v_ref, world_rows=384, kv_row0=384)
    src0 = allr.for_source(0)
    assert src0.kv_row0 == 0 and src0.k is allr.k
    h = allr.heads(slice(1, 2), slice(1, 2))
    assert h.q.shape == (4, 1) and h.k.shape == (6, 1) and h.v_ref.tolist() == [5.0] and h.kv_row0 == 384
    assert len(allr.args()) == 7


def test_nvtx_ranges_are_free_on_cpu(monkeypatch):
    import torch

    from ring_flash_attn_b200.utils import trace

    monkeypatch.setenv("RFA_B200_NVTX", "1")
    assert trace.enabled()
    with trace.nvtx("rfa.test", torch.zeros(1)):  # CPU tensor: no CUDA call is made
        pass
    monkeypatch.delenv("RFA_B200_NVTX")
    assert not trace.enabled()


def test_packed_inputs_get_one_concatenated_gradient():
    """qkv / kv packed entry points: the slices are views, their gradient is a single stack (no zero-fill + add per
    slice), and it equals the gradient of the unpacked call."""
    import torch

    import ring_flash_attn_b200 as rfa

    torch.manual_seed(0)
    qkv = torch.randn(2, 96, 3, 2, 32)
    dout = torch.randn(2, 96, 2, 32)
    a = qkv.clone().requires_grad_(True)
    out = rfa.ring_flash_attn_qkvpacked_func(a, causal=True)
    node = out.grad_fn
    names = set()
    stack = [node]
    while stack:
        n = stack.pop()
        if n is None or n.name() in names and n.name() != "torch::autograd::AccumulateGrad":
            continue
        names.add(n.name())
        stack.extend(f for f, _ in n.next_functions)
    assert any("_Unpack" in n for n in names) and not any("Select" in n or "Unbind" in n for n in names), names
    out.backward(dout)
    q, k, v = (qkv[:, :, i].clone().requires_grad_(True) for i in range(3))
    rfa.ring_flash_attn_func(q, k, v, causal=True).backward(dout)
    torch.testing.assert_close(a.grad, torch.stack((q.grad, k.grad, v.grad), dim=2))
    # kv packed, varlen layout, only kv requires grad
    kv = torch.randn(96, 2, 2, 32, requires_grad=True)
    qq = torch.randn(96, 2, 32)
    cu = torch.tensor([0, 40, 96], dtype=torch.int32)
    rfa.ring_flash_attn_varlen_kvpacked_func(qq, kv, cu, 56, causal=True).sum().backward()
    assert kv.grad.shape == kv.shape and torch.isfinite(kv.grad).all()
    with torch.no_grad():  # no autograd: plain views
        assert rfa.ring_flash_attn_qkvpacked_func(qkv, causal=True).shape == (2, 96, 2, 32)
