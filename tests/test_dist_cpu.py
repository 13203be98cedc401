"""All public entry points on CPU/gloo (world_size 2 and 4) against the dense fp32 oracle.

This is the reference's test strategy (/root/reference/test/test_*_func.py: full problem on every
rank as oracle, per-rank shard comparison of out / lse / dq / dk / dv) with real tolerances, plus
the coverage the reference lacks: kvpacked/unpacked entry points, GQA, non-causal, world sizes.
"""
import pytest
import torch
import torch.distributed as dist

import ring_flash_attn_b200 as rfa
from ring_flash_attn_b200.ops.dense import attention_oracle, varlen_attention_oracle
from ring_flash_attn_b200.parallel import layouts
from dist_utils import run_distributed

TOL = dict(atol=2e-5, rtol=2e-4)


def _bcast(t):
    dist.broadcast(t, src=0)
    return t


def _batch_case(rank, world, scheme, causal, packing, hq, hkv, dtype, window=(-1, -1)):
    torch.manual_seed(0)
    b, d = 2, 16
    s = 8 * world * 2
    q = _bcast(torch.randn(b, s, hq, d, dtype=dtype))
    k = _bcast(torch.randn(b, s, hkv, d, dtype=dtype))
    v = _bcast(torch.randn(b, s, hkv, d, dtype=dtype))
    dout = _bcast(torch.randn(b, s, hq, d, dtype=dtype))
    shard = getattr(layouts, f"shard_{scheme}")
    qr, kr, vr = (x.detach().clone().requires_grad_(True) for x in (q, k, v))
    ref_out, ref_lse = attention_oracle(qr, kr, vr, causal, window_size=window)
    ref_out.backward(dout.float())

    lq, lk, lv = (shard(x, rank, world).detach().requires_grad_(True) for x in (q, k, v))
    prefix = {"ring": "ring", "zigzag": "zigzag_ring", "stripe": "stripe"}[scheme]
    if packing == "qkv":
        qkv = torch.stack([lq, lk, lv], dim=2).detach().requires_grad_(True)
        fn = getattr(rfa, f"{prefix}_flash_attn_qkvpacked_func")
        out, lse, _ = fn(qkv, causal=causal, window_size=window, return_attn_probs=True)
    elif packing == "kv":
        kv = torch.stack([lk, lv], dim=2).detach().requires_grad_(True)
        fn = getattr(rfa, f"{prefix}_flash_attn_kvpacked_func")
        out, lse, _ = fn(lq, kv, causal=causal, window_size=window, return_attn_probs=True)
    else:
        fn = getattr(rfa, f"{prefix}_flash_attn_func")
        out, lse, _ = fn(lq, lk, lv, causal=causal, window_size=window, return_attn_probs=True)
    assert out.dtype == dtype and lse.dtype == torch.float32
    out.backward(shard(dout, rank, world))
    if packing == "qkv":
        gq, gk, gv = qkv.grad[:, :, 0], qkv.grad[:, :, 1], qkv.grad[:, :, 2]
    elif packing == "kv":
        gq, gk, gv = lq.grad, kv.grad[:, :, 0], kv.grad[:, :, 1]
    else:
        gq, gk, gv = lq.grad, lk.grad, lv.grad
    tol = TOL if dtype == torch.float32 else dict(atol=3e-2, rtol=3e-2)
    torch.testing.assert_close(out.float(), shard(ref_out, rank, world), **tol)
    torch.testing.assert_close(lse, shard(ref_lse, rank, world, dim=2), atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(gq.float(), shard(qr.grad, rank, world).float(), **tol)
    torch.testing.assert_close(gk.float(), shard(kr.grad, rank, world).float(), **tol)
    torch.testing.assert_close(gv.float(), shard(vr.grad, rank, world).float(), **tol)


BATCH_CASES = [
    ("ring", True, "qkv", 4, 4), ("ring", False, "kv", 4, 2), ("ring", True, "none", 4, 1),
    ("zigzag", True, "qkv", 4, 4), ("zigzag", True, "kv", 6, 2), ("zigzag", True, "none", 2, 2),
    ("stripe", True, "qkv", 4, 4), ("stripe", True, "kv", 4, 2), ("stripe", True, "none", 3, 3),
]


def _all_batch_cases(rank, world):
    # one process group per world size: spawning + rendezvous dominates the run time otherwise
    for scheme, causal, packing, hq, hkv in BATCH_CASES:
        _batch_case(rank, world, scheme, causal, packing, hq, hkv, torch.float32)


@pytest.mark.parametrize("world", [2, 4])
def test_batch_schemes(world):
    """9 cases: {ring, zigzag, stripe} x {qkvpacked, kvpacked, unpacked} incl. GQA/MQA and non-causal ring."""
    run_distributed(_all_batch_cases, world)


def test_batch_bf16_dtype_roundtrip():
    run_distributed(_batch_case, 2, "zigzag", True, "qkv", 4, 4, torch.bfloat16)


def test_baseline_config_1():
    """BASELINE.json config 1: ring qkvpacked, world 2, bs=1 seq=256 nheads=4 d=64, fwd numerics."""
    run_distributed(_config1, 2)


def _config1(rank, world):
    torch.manual_seed(1)
    qkv = _bcast(torch.randn(1, 256, 3, 4, 64))
    ref, ref_lse = attention_oracle(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], True)
    out, lse, _ = rfa.ring_flash_attn_qkvpacked_func(layouts.shard_ring(qkv, rank, world), causal=True,
                                                     return_attn_probs=True)
    torch.testing.assert_close(out, layouts.shard_ring(ref, rank, world), **TOL)
    torch.testing.assert_close(lse, layouts.shard_ring(ref_lse, rank, world, dim=2), atol=1e-4, rtol=1e-4)


def _varlen_case(rank, world, scheme, causal, packing, hq, hkv, window=(-1, -1)):
    torch.manual_seed(0)
    d = 16
    unit = 2 * world
    lens = [unit * 2, unit * 5, unit * 3]
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    total = cu[-1]
    q = _bcast(torch.randn(total, hq, d))
    k = _bcast(torch.randn(total, hkv, d))
    v = _bcast(torch.randn(total, hkv, d))
    dout = _bcast(torch.randn(total, hq, d))
    cu_t = torch.tensor(cu, dtype=torch.int32)
    qr, kr, vr = (x.detach().clone().requires_grad_(True) for x in (q, k, v))
    ref_out, ref_lse = varlen_attention_oracle(qr, kr, vr, cu_t, causal, window_size=window)
    ref_out.backward(dout)
    shard = layouts.shard_ring_varlen if scheme == "ring" else layouts.shard_zigzag_varlen
    lq, lk, lv = (shard(x, cu, rank, world).detach().requires_grad_(True) for x in (q, k, v))
    local_cu = cu_t // world
    max_s = int((local_cu[1:] - local_cu[:-1]).max())
    prefix = "ring" if scheme == "ring" else "zigzag_ring"
    if packing == "qkv":
        qkv = torch.stack([lq, lk, lv], dim=1).detach().requires_grad_(True)
        out, lse, _ = getattr(rfa, f"{prefix}_flash_attn_varlen_qkvpacked_func")(
            qkv, local_cu, max_s, causal=causal, window_size=window, return_attn_probs=True)
    elif packing == "kv":
        kv = torch.stack([lk, lv], dim=1).detach().requires_grad_(True)
        out, lse, _ = getattr(rfa, f"{prefix}_flash_attn_varlen_kvpacked_func")(
            lq, kv, local_cu, max_s, causal=causal, window_size=window, return_attn_probs=True)
    else:
        out, lse, _ = getattr(rfa, f"{prefix}_flash_attn_varlen_func")(
            lq, lk, lv, local_cu, max_s, causal=causal, window_size=window, return_attn_probs=True)
    out.backward(shard(dout, cu, rank, world))
    if packing == "qkv":
        gq, gk, gv = qkv.grad[:, 0], qkv.grad[:, 1], qkv.grad[:, 2]
    elif packing == "kv":
        gq, gk, gv = lq.grad, kv.grad[:, 0], kv.grad[:, 1]
    else:
        gq, gk, gv = lq.grad, lk.grad, lv.grad
    torch.testing.assert_close(out, shard(ref_out, cu, rank, world), **TOL)
    torch.testing.assert_close(lse, shard(ref_lse.transpose(0, 1), cu, rank, world).transpose(0, 1),
                               atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(gq, shard(qr.grad, cu, rank, world), **TOL)
    torch.testing.assert_close(gk, shard(kr.grad, cu, rank, world), **TOL)
    torch.testing.assert_close(gv, shard(vr.grad, cu, rank, world), **TOL)


VARLEN_CASES = [
    ("ring", True, "qkv", 4, 4), ("ring", False, "kv", 4, 2), ("ring", True, "none", 2, 1),
    ("zigzag", True, "qkv", 4, 4), ("zigzag", True, "kv", 4, 2), ("zigzag", True, "none", 2, 2),
]


def _all_varlen_cases(rank, world):
    for scheme, causal, packing, hq, hkv in VARLEN_CASES:
        _varlen_case(rank, world, scheme, causal, packing, hq, hkv)


@pytest.mark.parametrize("world", [2, 4])
def test_varlen_schemes(world):
    run_distributed(_all_varlen_cases, world)


def _llama3_case(rank, world, causal, packing, hq, hkv, stride, window=(-1, -1)):
    torch.manual_seed(0)
    d = 8
    cu = [0, 3 * world + 1, 7 * world - 1, 12 * world]
    total = cu[-1]
    q = _bcast(torch.randn(total, hq, d))
    k = _bcast(torch.randn(total, hkv, d))
    v = _bcast(torch.randn(total, hkv, d))
    dout = _bcast(torch.randn(total, hq, d))
    cu_t = torch.tensor(cu, dtype=torch.int32)
    qr, kr, vr = (x.detach().clone().requires_grad_(True) for x in (q, k, v))
    ref_out, ref_lse = varlen_attention_oracle(qr, kr, vr, cu_t, causal, window_size=window)
    ref_out.backward(dout)
    sh = lambda x: layouts.shard_llama3(x, rank, world)  # noqa: E731
    lq, lk, lv = (sh(x).detach().requires_grad_(True) for x in (q, k, v))
    cq, ck, mq, mk, ks = rfa.llama3_flash_attn_prepare_cu_seqlens(cu_t, causal, rank, world)
    if packing == "qkv":
        qkv = torch.stack([lq, lk, lv], dim=1).detach().requires_grad_(True)
        out, lse, _ = rfa.llama3_flash_attn_varlen_qkvpacked_func(
            qkv, cq, ck, mq, mk, heads_k_stride=stride, local_k_slice=ks, causal=causal, window_size=window,
            return_attn_probs=True)
    elif packing == "kv":
        kv = torch.stack([lk, lv], dim=1).detach().requires_grad_(True)
        out, lse, _ = rfa.llama3_flash_attn_varlen_kvpacked_func(
            lq, kv, cq, ck, mq, mk, heads_k_stride=stride, local_k_slice=ks, causal=causal, window_size=window,
            return_attn_probs=True)
    else:
        out, lse, _ = rfa.llama3_flash_attn_varlen_func(
            lq, lk, lv, cq, ck, mq, mk, heads_k_stride=stride, local_k_slice=ks, causal=causal,
            window_size=window, return_attn_probs=True)
    out.backward(sh(dout))
    if packing == "qkv":
        gq, gk, gv = qkv.grad[:, 0], qkv.grad[:, 1], qkv.grad[:, 2]
    elif packing == "kv":
        gq, gk, gv = lq.grad, kv.grad[:, 0], kv.grad[:, 1]
    else:
        gq, gk, gv = lq.grad, lk.grad, lv.grad
    torch.testing.assert_close(out, sh(ref_out), **TOL)
    torch.testing.assert_close(lse, sh(ref_lse.transpose(0, 1)).transpose(0, 1), atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(gq, sh(qr.grad), **TOL)
    torch.testing.assert_close(gk, sh(kr.grad), **TOL)
    torch.testing.assert_close(gv, sh(vr.grad), **TOL)


LLAMA3_CASES = [(True, "qkv", 4, 4, 1), (True, "kv", 4, 2, 2), (False, "none", 4, 2, 1), (True, "none", 8, 4, 4)]


def _all_llama3_cases(rank, world):
    for causal, packing, hq, hkv, stride in LLAMA3_CASES:
        _llama3_case(rank, world, causal, packing, hq, hkv, stride)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_llama3(world):
    run_distributed(_all_llama3_cases, world)


def _all_window_cases(rank, world):
    """Sliding window (left, right) applied to *global* positions, every scheme, windows smaller and larger
    than a shard (the reference forwards window_size to each per-block flash_attn call, which is only
    meaningful for its llama3 path: /root/reference/ring_flash_attn/llama3_flash_attn_varlen.py:147)."""
    for win in ((5, 0), (8 * world + 3, 0)):
        _batch_case(rank, world, "ring", True, "qkv", 4, 4, torch.float32, win)
        _batch_case(rank, world, "zigzag", True, "kv", 4, 2, torch.float32, win)
        _batch_case(rank, world, "stripe", True, "none", 2, 1, torch.float32, win)
        _varlen_case(rank, world, "ring", True, "kv", 4, 2, win)
        _varlen_case(rank, world, "zigzag", True, "none", 2, 2, win)
        _llama3_case(rank, world, True, "kv", 4, 2, 2, win)
    # non-causal two-sided windows (ring and llama3 only: zigzag/stripe are causal-only layouts)
    for win in ((3, 2), (0, 7), (-1, 4), (6, -1)):
        _batch_case(rank, world, "ring", False, "none", 2, 2, torch.float32, win)
        _varlen_case(rank, world, "ring", False, "qkv", 2, 2, win)
        _llama3_case(rank, world, False, "none", 2, 1, 1, win)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_sliding_window(world):
    run_distributed(_all_window_cases, world)


def _zigzag_llama3_case(rank, world, causal, packing, hq, hkv, window=(-1, -1)):
    torch.manual_seed(0)
    d = 8
    total = 2 * world * 11
    cu = [0, 7, 8, total // 2 + 3, total]  # arbitrary document lengths, cut anywhere by the chunk boundaries
    q = _bcast(torch.randn(total, hq, d))
    k = _bcast(torch.randn(total, hkv, d))
    v = _bcast(torch.randn(total, hkv, d))
    dout = _bcast(torch.randn(total, hq, d))
    cu_t = torch.tensor(cu, dtype=torch.int32)
    qr, kr, vr = (x.detach().clone().requires_grad_(True) for x in (q, k, v))
    ref_out, ref_lse = varlen_attention_oracle(qr, kr, vr, cu_t, causal, window_size=window)
    ref_out.backward(dout)
    sh = lambda x: layouts.shard_zigzag_llama3(x, rank, world)  # noqa: E731
    lq, lk, lv = (sh(x).detach().requires_grad_(True) for x in (q, k, v))
    if packing == "qkv":
        qkv = torch.stack([lq, lk, lv], dim=1).detach().requires_grad_(True)
        out, lse, _ = rfa.zigzag_llama3_flash_attn_varlen_qkvpacked_func(qkv, cu_t, causal=causal, window_size=window,
                                                                         return_attn_probs=True)
    elif packing == "kv":
        kv = torch.stack([lk, lv], dim=1).detach().requires_grad_(True)
        out, lse, _ = rfa.zigzag_llama3_flash_attn_varlen_kvpacked_func(lq, kv, cu_t, causal=causal,
                                                                        window_size=window, return_attn_probs=True)
    else:
        out, lse, _ = rfa.zigzag_llama3_flash_attn_varlen_func(lq, lk, lv, cu_t, causal=causal, window_size=window,
                                                               return_attn_probs=True)
    out.backward(sh(dout))
    if packing == "qkv":
        gq, gk, gv = qkv.grad[:, 0], qkv.grad[:, 1], qkv.grad[:, 2]
    elif packing == "kv":
        gq, gk, gv = lq.grad, kv.grad[:, 0], kv.grad[:, 1]
    else:
        gq, gk, gv = lq.grad, lk.grad, lv.grad
    torch.testing.assert_close(out, sh(ref_out), **TOL)
    torch.testing.assert_close(lse, sh(ref_lse.transpose(0, 1)).transpose(0, 1), atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(gq, sh(qr.grad), **TOL)
    torch.testing.assert_close(gk, sh(kr.grad), **TOL)
    torch.testing.assert_close(gv, sh(vr.grad), **TOL)
    # RoPE helper: position of every local token inside its document
    pos = torch.cat([torch.arange(b - a) for a, b in zip(cu[:-1], cu[1:])])
    assert torch.equal(layouts.positions_zigzag_llama3(cu, rank, world), sh(pos))


def _all_zigzag_llama3_cases(rank, world):
    _zigzag_llama3_case(rank, world, True, "qkv", 4, 4)
    _zigzag_llama3_case(rank, world, True, "kv", 4, 2)
    _zigzag_llama3_case(rank, world, False, "none", 2, 1)
    _zigzag_llama3_case(rank, world, True, "none", 2, 2, (6, 0))
    _zigzag_llama3_case(rank, world, False, "kv", 2, 2, (4, 3))


@pytest.mark.parametrize("world", [1, 2, 4])
def test_zigzag_llama3(world):
    """Beyond the reference (its README TODO): zigzag over the flat packed stream, arbitrary document lengths."""
    run_distributed(_all_zigzag_llama3_cases, world)


def _world3_cases(rank, world):
    # a world size that is not a power of two: one case per scheme family
    _batch_case(rank, world, "ring", True, "kv", 4, 2, torch.float32)
    _batch_case(rank, world, "zigzag", True, "qkv", 2, 2, torch.float32)
    _batch_case(rank, world, "stripe", True, "none", 4, 1, torch.float32, (7, 0))
    _varlen_case(rank, world, "zigzag", True, "kv", 4, 2)
    _llama3_case(rank, world, True, "none", 4, 2, 2)
    _zigzag_llama3_case(rank, world, True, "kv", 4, 2)


def test_world_size_three():
    run_distributed(_world3_cases, 3)


def _single_process_case():
    # world_size 1 without any process group: every scheme degenerates to plain attention
    torch.manual_seed(0)
    qkv = torch.randn(1, 32, 3, 2, 16, requires_grad=True)
    ref, _ = attention_oracle(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], True)
    for fn in (rfa.ring_flash_attn_qkvpacked_func, rfa.zigzag_ring_flash_attn_qkvpacked_func,
               rfa.stripe_flash_attn_qkvpacked_func):
        torch.testing.assert_close(fn(qkv, causal=True), ref, **TOL)


def _padded_head_dim_cases(rank, world):
    # d=16 padded to the kernels' 128: identical results, gradients of the padding never leak out
    _batch_case(rank, world, "zigzag", True, "qkv", 2, 2, torch.float32)
    _varlen_case(rank, world, "ring", True, "kv", 4, 2)
    _llama3_case(rank, world, True, "none", 4, 2, 1)


def _tiled_dense_cases(rank, world):
    _batch_case(rank, world, "zigzag", True, "kv", 4, 2, torch.float32)
    _batch_case(rank, world, "stripe", True, "qkv", 2, 2, torch.float32, (9, 0))
    _varlen_case(rank, world, "ring", False, "none", 2, 1, (5, 3))
    _llama3_case(rank, world, True, "qkv", 2, 2, 1)


def test_dense_fallback_is_tiled(monkeypatch):
    """RFA_B200_DENSE_TILE=3: every dense block is cut into 3 x 12 sub-blocks with re-based band offsets."""
    monkeypatch.setenv("RFA_B200_DENSE_TILE", "3")
    run_distributed(_tiled_dense_cases, 2)


def test_head_dim_padding(monkeypatch):
    monkeypatch.setenv("RFA_B200_PAD_HEAD_DIM", "force")
    run_distributed(_padded_head_dim_cases, 2)


def test_head_dim_padding_reaches_engine(monkeypatch):
    from ring_flash_attn_b200.parallel import api, engine

    monkeypatch.setenv("RFA_B200_PAD_HEAD_DIM", "force")
    seen = []
    real = engine.cp_forward
    monkeypatch.setattr(engine, "cp_forward", lambda plan, q, *a, **kw: (seen.append(q.shape[-1]), real(plan, q, *a, **kw))[1])
    qkv = torch.randn(1, 32, 3, 2, 24, requires_grad=True)
    ref, _ = attention_oracle(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], True)
    out = rfa.ring_flash_attn_qkvpacked_func(qkv, causal=True)
    assert seen == [64] and 64 in api.KERNEL_HEAD_DIMS and out.shape[-1] == 24 and out.is_contiguous()
    torch.testing.assert_close(out, ref, **TOL)
    g, = torch.autograd.grad(out.sum(), qkv)
    gr, = torch.autograd.grad(ref.sum(), qkv)
    torch.testing.assert_close(g, gr, **TOL)
    monkeypatch.setenv("RFA_B200_PAD_HEAD_DIM", "0")
    rfa.ring_flash_attn_qkvpacked_func(qkv, causal=True)
    assert seen[-1] == 24


def _subgroup_case(rank, world):
    """Two independent context-parallel groups inside one job (ranks {0,1} and {2,3}), as when CP is combined with
    data parallelism: group ranks are translated to global ranks for the point-to-point transport
    (/root/reference/ring_flash_attn/utils.py:109-111)."""
    groups = [dist.new_group([0, 1]), dist.new_group([2, 3])]  # collective: every rank creates both
    g = groups[rank // 2]
    grank, gworld = dist.get_rank(g), dist.get_world_size(g)
    assert (grank, gworld) == (rank % 2, 2)
    torch.manual_seed(100 + rank // 2)  # different data per group
    qkv = torch.randn(1, 32, 3, 2, 16)
    dist.broadcast(qkv, src=(rank // 2) * 2, group=g)
    ref, _ = attention_oracle(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], True)
    for scheme, fn in (("zigzag", rfa.zigzag_ring_flash_attn_qkvpacked_func), ("ring", rfa.ring_flash_attn_qkvpacked_func)):
        shard = getattr(layouts, f"shard_{scheme}")
        out = fn(shard(qkv, grank, gworld), causal=True, group=g)
        torch.testing.assert_close(out, shard(ref, grank, gworld), **TOL)
    cu = torch.tensor([0, 13, 32], dtype=torch.int32)
    refv, _ = varlen_attention_oracle(qkv[0, :, 0], qkv[0, :, 1], qkv[0, :, 2], cu, True)
    cq, ck, mq, mk, ks = rfa.llama3_flash_attn_prepare_cu_seqlens(cu, True, grank, gworld)
    out = rfa.llama3_flash_attn_varlen_qkvpacked_func(layouts.shard_llama3(qkv[0], grank, gworld), cq, ck, mq, mk,
                                                      heads_k_stride=1, local_k_slice=ks, causal=True, group=g)
    torch.testing.assert_close(out, layouts.shard_llama3(refv, grank, gworld), **TOL)


def test_process_subgroups():
    run_distributed(_subgroup_case, 4)


def test_custom_softmax_scale_and_lse_shapes():
    """softmax_scale is forwarded (not the 1/sqrt(d) default); lse shapes follow flash-attn: (B, H, S) batch,
    (H, T) varlen."""
    torch.manual_seed(0)
    q = torch.randn(2, 24, 4, 16, requires_grad=True)
    kv = torch.randn(2, 24, 2, 2, 16)
    ref, ref_lse = attention_oracle(q, kv[:, :, 0], kv[:, :, 1], True, softmax_scale=0.37)
    out, lse, none = rfa.zigzag_ring_flash_attn_kvpacked_func(q, kv, softmax_scale=0.37, causal=True,
                                                             return_attn_probs=True)
    assert none is None and lse.shape == (2, 4, 24)
    torch.testing.assert_close(out, ref, **TOL)
    torch.testing.assert_close(lse, ref_lse, atol=1e-4, rtol=1e-4)
    default = rfa.zigzag_ring_flash_attn_kvpacked_func(q, kv, causal=True)
    assert not torch.allclose(default, out)
    cu = torch.tensor([0, 10, 24], dtype=torch.int32)
    qv, kvv = q[0].detach(), kv[0]
    refv, refv_lse = varlen_attention_oracle(qv, kvv[:, 0], kvv[:, 1], cu, False, softmax_scale=1.3)
    outv, lsev, _ = rfa.ring_flash_attn_varlen_kvpacked_func(qv, kvv, cu, 14, softmax_scale=1.3, causal=False,
                                                            return_attn_probs=True)
    assert lsev.shape == (4, 24)
    torch.testing.assert_close(outv, refv, **TOL)
    torch.testing.assert_close(lsev, refv_lse, atol=1e-4, rtol=1e-4)
    # deterministic=True is accepted everywhere and does not change the result on this path
    torch.testing.assert_close(rfa.ring_flash_attn_varlen_kvpacked_func(qv, kvv, cu, 14, softmax_scale=1.3,
                                                                      deterministic=True), outv, **TOL)


def test_single_process_no_group():
    _single_process_case()


def test_argument_guards():
    q = torch.randn(1, 8, 2, 16)
    with pytest.raises(AssertionError):
        rfa.zigzag_ring_flash_attn_func(q, q, q, causal=False)
    with pytest.raises(AssertionError):
        rfa.stripe_flash_attn_func(q, q, q, causal=False)
    with pytest.raises(NotImplementedError):
        rfa.ring_flash_attn_func(q, q, q, alibi_slopes=torch.ones(2))
    with pytest.raises(NotImplementedError):
        rfa.ring_flash_attn_func(q, q, q, dropout_p=0.1)
