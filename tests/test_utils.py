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
