"""The fused multi-GPU launch replayed in ONE process: W ranks' plans, push tables, forward / backward work tables
(ready flags, staging offsets, owners, rows in the owner's shard) and owner-side reduce tables are executed with
torch tensors standing in for the peer-mapped staging buffers and dK/dV inboxes, and compared with the dense
oracle.  This checks what the tables *mean* together (parallel/symm.py is the real driver of the same tables on
NVLink peer memory), with and without sliding windows."""
import pytest
import torch

from ring_flash_attn_b200.ops import attn_cuda, plan as P
from ring_flash_attn_b200.ops.dense import attention_oracle
from ring_flash_attn_b200.parallel import layouts, symm

from fake_ext import LO_NONE, FakeExt, _visible

D = 128


class _Ctx:
    group = None


def _plans(scheme, world, L, window):
    if scheme == "ring":
        ps = [P.plan_ring(r, world, 1, L, True, window) for r in range(world)]
    elif scheme == "zigzag":
        ps = [P.plan_zigzag(r, world, 1, L, window) for r in range(world)]
    else:
        ps = [P.plan_stripe(r, world, 1, L, window) for r in range(world)]
    for p in ps:
        p.peer = (lambda rr, _ps=ps: _ps[rr])
    return ps


def _fwd_rank(fake, q, k, v, k_stage, v_stage, items, segs, seg_lo, scale, hq):
    """FakeExt forward with segments that read either the local tensors (flag < 0) or the staging buffers."""
    tq = q.shape[0]
    out = torch.zeros(tq, hq, D)
    lse = torch.full((hq, tq), float("-inf"))
    # route every segment through one concatenated key space: [local rows | staging rows]
    kk, vv = torch.cat([k, k_stage]), torch.cat([v, v_stage])
    base = k.shape[0]
    segs2 = [[(r0 if flag < 0 else base + r0), n, d, flag] for r0, n, d, flag in segs]
    fake._fwd(q, kk, vv, torch.tensor(items), torch.tensor(segs2 if segs2 else [[0, 0, 0, -1]]), seg_lo, out, lse, scale)
    return out, lse


def _replay(plans, world, L, hq, hkv, lq, lk, lv, ldo, ref_out, ref_lse, ref_dq, ref_dk, ref_dv):
    """Run forward + backward of every rank from its tables; ``l*`` / ``ref_*`` are per-rank lists."""
    fake = FakeExt()
    scale = D ** -0.5
    row_bytes = hkv * D * 4
    # ---- push: every rank copies the rows its peers need into their staging slot [src = me]
    stage_k = [torch.full((world * L, hkv, D), float("nan")) for _ in range(world)]
    stage_v = [torch.full((world * L, hkv, D), float("nan")) for _ in range(world)]
    region = world * L * row_bytes
    dynamic = symm.is_dynamic(plans[0])
    # llama3-style plans: every rank only knows its own needs; the all-gather of the local tables is what
    # symm.needs_gathered does on the device, and the push / reduce roles derive their work from it
    needs_all = torch.stack([symm.local_needs_table(p) for p in plans]).tolist() if dynamic else None
    for me, p in enumerate(plans):
        if dynamic:
            assert symm.dynamic_ok(p)
            chunk = symm.push_chunk_rows(row_bytes)
            chunks = -(-L // chunk)
            n_tasks = (world - 1) * 2 * symm.NEED_RANGES * chunks
            tasks = []
            for ti in range(n_tasks):
                src_row, dst_off, rows, dst, which = symm.dynamic_push_task(needs_all, me, world, ti, chunk, chunks, L,
                                                                           row_bytes, region)
                assert dst != me
                if rows:
                    tasks.append([src_row, dst_off, rows | (dst << 32), which])
        else:
            t, per_dst = symm.push_tasks(p, _Ctx(), row_bytes, torch.device("cpu"))
            assert per_dst[me] == 0
            tasks = t.tolist()
        for src_row, dst_off, packed, which in tasks:
            rows, dst = packed & 0xFFFFFFFF, packed >> 32
            assert dst_off % row_bytes == 0
            dst_row = (dst_off - which * region) // row_bytes
            assert me * L <= dst_row and dst_row + rows <= (me + 1) * L
            (stage_k if which == 0 else stage_v)[dst][dst_row:dst_row + rows] = (lk if which == 0 else lv)[me][src_row:src_row + rows]
    # ---- forward
    outs, lses = [], []
    for me, p in enumerate(plans):
        offsets = {s: (0 if s == me else s * L) for s in range(world)}
        flags = {s: s for s in range(world) if s != me}
        if attn_cuda.has_window(p.segments):
            items, segs, seg_lo, _cov = attn_cuda.fwd_tables_window_host(p, p.segments, offsets, flags)
        else:
            items, segs, _cov = attn_cuda.fwd_tables_host(p, p.segments, offsets, flags)
            seg_lo = None
        o, l = _fwd_rank(fake, lq[me], lk[me], lv[me], stage_k[me], stage_v[me], items, segs, seg_lo, scale, hq)
        assert not torch.isnan(o).any(), "a forward segment read staging rows nobody pushed"
        torch.testing.assert_close(o, ref_out[me], atol=2e-5, rtol=2e-4)
        torch.testing.assert_close(l, ref_lse[me], atol=1e-4, rtol=1e-4)
        outs.append(o)
        lses.append(l)
    # ---- backward: tiles write into the owner's inbox slot [src = me]; owners reduce
    inbox = [torch.full((world, 2, L, hkv, D), float("nan")) for _ in range(world)]  # [owner][slot][dK|dV]
    for me, p in enumerate(plans):
        offsets = {s: (0 if s == me else s * L) for s in range(world)}
        flags = {s: s for s in range(world) if s != me}
        if attn_cuda.has_window(p.segments):
            items, qsegs, per_owner = attn_cuda.bwd_tables_window_host(p, p.segments, offsets, flags, fused=True)
            win = True
        else:
            it_t, qs_t, per_owner = attn_cuda.bwd_tables_fused(p, offsets, torch.device("cpu"), flags)
            items, qsegs, win = it_t.tolist(), qs_t.tolist(), False
        assert sum(per_owner) == len(items)
        delta = (outs[me] * ldo[me]).sum(-1).transpose(0, 1).contiguous()
        dq = torch.zeros(L, hq, D)
        kk, vv = torch.cat([lk[me], stage_k[me]]), torch.cat([lv[me], stage_v[me]])
        for kv_row0, kv_rows, b, c, flag, owner, out_row0, _z in items:
            src_row = kv_row0 if flag < 0 else L + kv_row0
            dk_t = torch.zeros(L + world * L, hkv, D)
            dv_t = torch.zeros(L + world * L, hkv, D)
            one = torch.tensor([[src_row, kv_rows, 0, c, flag, owner, out_row0, 0]])
            qs = torch.tensor(qsegs[b:b + c] if c else [[0, 0, 0, 0]])
            if c:
                fake._bwd(lq[me], ldo[me], kk, vv, dq, one, qs, lses[me], delta, dk_t, dv_t, scale, win)
            inbox[owner][me, 0, out_row0:out_row0 + kv_rows] = dk_t[src_row:src_row + kv_rows]
            inbox[owner][me, 1, out_row0:out_row0 + kv_rows] = dv_t[src_row:src_row + kv_rows]
        torch.testing.assert_close(dq, ref_dq[me], atol=5e-5, rtol=5e-4)
    for me, p in enumerate(plans):
        tasks = symm.reduce_tasks(p, _Ctx(), torch.device("cpu")).tolist()
        dk, dv = torch.zeros(L, hkv, D), torch.zeros(L, hkv, D)
        if dynamic:  # fixed row blocks; the kernel computes the contributing ranks per row from the needs table
            tasks = [[r, 1, symm.dynamic_row_mask(needs_all, me, world, r), 0]
                     for row0, rows, _m, _p in tasks for r in range(row0, row0 + rows)]
        for row0, rows, mask, _pad in tasks:
            for s in range(world):
                if (mask >> s) & 1:
                    part_k, part_v = inbox[me][s, 0, row0:row0 + rows], inbox[me][s, 1, row0:row0 + rows]
                    assert not torch.isnan(part_k).any(), "the reduction reads inbox rows no peer wrote"
                    dk[row0:row0 + rows] += part_k
                    dv[row0:row0 + rows] += part_v
        torch.testing.assert_close(dk, ref_dk[me], atol=5e-5, rtol=5e-4)
        torch.testing.assert_close(dv, ref_dv[me], atol=5e-5, rtol=5e-4)


@pytest.mark.parametrize("scheme", ["zigzag", "ring", "stripe"])
@pytest.mark.parametrize("window", [(-1, -1), (150, 0), (700, 0)])
def test_fused_tables_end_to_end(scheme, window):
    world, L, hq, hkv = 4, 384, 4, 2
    S = world * L
    torch.manual_seed(0)
    q, k, v, dout = (torch.randn(1, S, h, D) for h in (hq, hkv, hkv, hq))
    rq, rk, rv = (t.clone().requires_grad_(True) for t in (q, k, v))
    ref, ref_lse = attention_oracle(rq, rk, rv, True, window_size=window)
    ref.backward(dout)
    shard = getattr(layouts, f"shard_{scheme}")
    R = range(world)
    _replay(_plans(scheme, world, L, window), world, L, hq, hkv,
            [shard(q, r, world)[0] for r in R], [shard(k, r, world)[0] for r in R], [shard(v, r, world)[0] for r in R],
            [shard(dout, r, world)[0] for r in R], [shard(ref, r, world)[0] for r in R],
            [shard(ref_lse, r, world, dim=2)[0] for r in R], [shard(rq.grad, r, world)[0] for r in R],
            [shard(rk.grad, r, world)[0] for r in R], [shard(rv.grad, r, world)[0] for r in R])


@pytest.mark.parametrize("window", [(-1, -1), (200, 0)])
def test_fused_tables_llama3(window):
    """llama3 layout (contiguous split of packed documents): every rank's plan is built from its OWN prepare()
    outputs only; what the peers need of a shard comes from their plans (on hardware: the device-side needs
    exchange, parallel/symm.py:needs_gathered)."""
    from ring_flash_attn_b200.ops.dense import varlen_attention_oracle
    from ring_flash_attn_b200.parallel import api, ops

    world, L, hq, hkv = 4, 256, 4, 2
    S = world * L
    cu = (0, 300, 301, 777, S)
    torch.manual_seed(0)
    q, k, v, dout = (torch.randn(S, h, D) for h in (hq, hkv, hkv, hq))
    rq, rk, rv = (t.clone().requires_grad_(True) for t in (q, k, v))
    ref, ref_lse = varlen_attention_oracle(rq, rk, rv, torch.tensor(cu), True, window_size=window)
    ref.backward(dout)
    plans = []
    for r in range(world):
        cq, ck, _mq, _mk, ks = api.llama3_flash_attn_prepare_cu_seqlens(torch.tensor(cu, dtype=torch.int32), True, r,
                                                                       world)
        plans.append(ops.llama3_plan(r, world, L, tuple(cq.tolist()), tuple(ck.tolist()), int(ks.start), True, window))
    sh = lambda x, r: layouts.shard_llama3(x, r, world)  # noqa: E731
    R = range(world)
    _replay(plans, world, L, hq, hkv, [sh(q, r) for r in R], [sh(k, r) for r in R], [sh(v, r) for r in R],
            [sh(dout, r) for r in R], [sh(ref, r) for r in R],
            [sh(ref_lse.transpose(0, 1), r).transpose(0, 1) for r in R], [sh(rq.grad, r) for r in R],
            [sh(rk.grad, r) for r in R], [sh(rv.grad, r) for r in R])


@pytest.mark.parametrize("window", [(-1, -1), (200, 0)])
def test_fused_tables_zigzag_llama3(window):
    from ring_flash_attn_b200.ops.dense import varlen_attention_oracle
    from ring_flash_attn_b200.parallel import ops

    world, L, hq, hkv = 4, 256, 4, 2
    S = world * L
    cu = (0, 130, 131, 700, S)
    torch.manual_seed(0)
    q, k, v, dout = (torch.randn(S, h, D) for h in (hq, hkv, hkv, hq))
    rq, rk, rv = (t.clone().requires_grad_(True) for t in (q, k, v))
    ref, ref_lse = varlen_attention_oracle(rq, rk, rv, torch.tensor(cu), True, window_size=window)
    ref.backward(dout)
    plans = [ops.zigzag_llama3_plan(r, world, cu, True, window) for r in range(world)]
    sh = lambda x, r: layouts.shard_zigzag_llama3(x, r, world)  # noqa: E731
    R = range(world)
    _replay(plans, world, L, hq, hkv, [sh(q, r) for r in R], [sh(k, r) for r in R], [sh(v, r) for r in R],
            [sh(dout, r) for r in R], [sh(ref, r) for r in R],
            [sh(ref_lse.transpose(0, 1), r).transpose(0, 1) for r in R], [sh(rq.grad, r) for r in R],
            [sh(rk.grad, r) for r in R], [sh(rv.grad, r) for r in R])


def test_visible_helper_matches_lo_none():
    assert bool(_visible(4, 0, 6, 1 << 29, LO_NONE).all())
