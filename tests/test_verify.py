"""The sampled full-size correctness witness (utils/verify.py, used by ``bench.py --check``) on CPU/gloo:
it must accept a correct step for every batch scheme and flag a corrupted output / gradient."""
import pytest
import torch

import ring_flash_attn_b200 as rfa
from ring_flash_attn_b200.utils.verify import sampled_check
from dist_utils import run_distributed


def _case(rank, world, scheme, hq, hkv):
    torch.manual_seed(5 + rank)
    s_l, d = 64, 16
    q = torch.randn(1, s_l, hq, d, requires_grad=True)
    k = torch.randn(1, s_l, hkv, d, requires_grad=True)
    v = torch.randn(1, s_l, hkv, d, requires_grad=True)
    dout = torch.randn(1, s_l, hq, d)
    prefix = {"ring": "ring", "zigzag": "zigzag_ring", "stripe": "stripe"}[scheme]
    out, lse, _ = getattr(rfa, f"{prefix}_flash_attn_func")(q, k, v, causal=True, return_attn_probs=True)
    out.backward(dout)
    args = (scheme, q.detach()[0], k.detach()[0], v.detach()[0], dout[0])
    good = sampled_check(*args, out.detach()[0], lse[0], q.grad[0], k.grad[0], v.grad[0], n_rows=16, key_rows=32)
    assert good["ok"], good
    assert set(good["max_err"]) == {"out", "lse", "dq", "dk", "dv"}
    assert max(good["max_err"].values()) < 1e-4, good
    # a wrong dk on ONE rank must fail the (collective) verdict on every rank
    bad_dk = k.grad[0].clone()
    if rank == world - 1:
        bad_dk[s_l // 2 // 128 * 128 + 3, 0, 5] += 1.0
    bad = sampled_check(*args, out.detach()[0], lse[0], q.grad[0], bad_dk, v.grad[0], n_rows=16, key_rows=32)
    assert not bad["ok"] and bad["max_err"]["dk"] > 0.5, bad
    fwd_only = sampled_check(scheme, q.detach()[0], k.detach()[0], v.detach()[0], None, out.detach()[0], lse[0],
                             None, None, None, n_rows=8)
    assert fwd_only["ok"] and set(fwd_only["max_err"]) == {"out", "lse"}


def _all(rank, world):
    for scheme, hq, hkv in (("zigzag", 4, 4), ("ring", 4, 2), ("stripe", 2, 1)):
        _case(rank, world, scheme, hq, hkv)


@pytest.mark.parametrize("world", [1, 2])
def test_sampled_check(world):
    run_distributed(_all, world)
