"""Host-side work tables consumed by the sm_100a kernels (no GPU needed): coverage and exclusivity invariants."""
import pytest
import torch

from ring_flash_attn_b200.ops import attn_cuda, plan as P
from ring_flash_attn_b200.parallel import api, symm


def _plans(scheme, world, L=600):
    if scheme == "ring":
        return [P.plan_ring(r, world, 1, L, True) for r in range(world)]
    if scheme == "zigzag":
        return [P.plan_zigzag(r, world, 2, L) for r in range(world)]
    if scheme == "stripe":
        return [P.plan_stripe(r, world, 1, L) for r in range(world)]
    cu = [0, 200, 520, 600]
    if scheme == "ring_varlen":
        return [P.plan_ring_varlen(r, world, cu, True) for r in range(world)]
    return [P.plan_zigzag_varlen(r, world, cu) for r in range(world)]


@pytest.mark.parametrize("scheme", ["ring", "zigzag", "stripe", "ring_varlen", "zigzag_varlen"])
@pytest.mark.parametrize("world", [1, 4])
def test_forward_items_cover_every_query_row_once(scheme, world):
    for plan in _plans(scheme, world):
        offsets = {s: (0 if s == plan.rank else s * plan.kv_rows) for s in range(world)}
        items, segs, covered = attn_cuda.fwd_tables_host(plan, plan.segments, offsets, {s: s for s in range(world)})
        assert covered
        seen = torch.zeros(plan.q_rows, dtype=torch.int32)
        for q_row0, q_rows, q_off, seg_begin, seg_count, *_ in items:
            assert 1 <= q_rows <= attn_cuda.Q_ITEM_ROWS and seg_count >= 1
            seen[q_row0:q_row0 + q_rows] += 1
            for kv_row0, kv_len, diag, flag in segs[seg_begin:seg_begin + seg_count]:
                assert kv_len > 0 and kv_row0 >= 0
        assert torch.all(seen == 1)


def _visible_pairs_from_plan(plan):
    total = 0
    for s in plan.segments:
        total += P.visible_area(plan.q_chunks[s.chunk].rows, s.kv_len, s.diag)
    return total


@pytest.mark.parametrize("scheme", ["ring", "zigzag", "stripe", "ring_varlen", "zigzag_varlen"])
@pytest.mark.parametrize("world", [1, 4])
def test_backward_tiles_are_exclusive_and_complete(scheme, world):
    for plan in _plans(scheme, world):
        offsets = {s: s * plan.kv_rows for s in range(world)}
        items, qsegs = attn_cuda.bwd_tables_host(plan, plan.segments, offsets)
        owner_rows = torch.zeros(world * plan.kv_rows, dtype=torch.int32)
        pairs = 0
        for kv_row0, kv_rows, seg_begin, seg_count, flag, *_ in items:
            assert 1 <= kv_rows <= attn_cuda.K_TILE_ROWS
            owner_rows[kv_row0:kv_row0 + kv_rows] += 1
            for q_row0, q_len, diag, _pad in qsegs[seg_begin:seg_begin + seg_count]:
                d = None if diag >= attn_cuda.DIAG_FULL else diag
                pairs += P.visible_area(q_len, kv_rows, d)
        assert int(owner_rows.max()) <= 1, "a key row has two writers in one launch"
        # every (query, key) pair of the plan is visited exactly once by the backward tiles
        assert pairs == _visible_pairs_from_plan(plan)


@pytest.mark.parametrize("scheme", ["zigzag", "ring"])
def test_fused_tables_push_and_reduce_are_consistent(scheme):
    world = 4
    plans = _plans(scheme, world, L=512)
    for r, p in enumerate(plans):
        p.peer = (lambda rr, _ps=plans: _ps[rr])

    class Ctx:
        group = None

    per = {}
    for p in plans:
        t, per_dst = symm.push_tasks(p, Ctx(), 1024, torch.device("cpu"))
        per[p.rank] = per_dst
        assert per_dst[p.rank] == 0
        for src_row, dst_off, packed, which in t.tolist():
            rows, dst = packed & 0xFFFFFFFF, packed >> 32
            assert 0 < rows and 0 <= dst < world and dst != p.rank and which in (0, 1)
            assert dst_off % 16 == 0
    for p in plans:
        offsets = {s: (0 if s == p.rank else s * p.kv_rows) for s in range(world)}
        flags = {s: s for s in range(world) if s != p.rank}
        _items, segs, _c = attn_cuda.fwd_tables_host(p, p.segments, offsets, flags)
        for s in {g[3] for g in segs if g[3] >= 0}:
            assert per[s][p.rank] > 0, f"rank {p.rank} waits for {s} but {s} pushes nothing to it"
        items, qsegs, per_owner = attn_cuda.bwd_tables_fused(p, offsets, torch.device("cpu"), flags)
        tasks = symm.reduce_tasks(p, Ctx(), torch.device("cpu")).tolist()
        covered = sorted((row0, row0 + rows) for row0, rows, _m, _p in tasks)
        assert covered[0][0] == 0 and covered[-1][1] == p.kv_rows
        assert all(a[1] == b[0] for a, b in zip(covered[:-1], covered[1:])), "reduce tasks must tile the shard"
        for row0, rows, mask, _p in tasks:
            for s in range(world):
                contributes = any(lo <= row0 and row0 + rows <= hi for lo, hi in symm._ranges(plans[s], p.rank))
                assert bool((mask >> s) & 1) == contributes
        assert sum(per_owner) == items.shape[0]


def test_llama3_same_local_slice_different_global_layouts_do_not_share_caches():
    """Two global packings can give a rank the same local description; its peers' needs still differ."""
    world, T = 8, 64
    a = torch.tensor([0, 8 * T], dtype=torch.int32)
    b = torch.tensor([0, 5 * T, 8 * T], dtype=torch.int32)
    pa = api.llama3_flash_attn_prepare_cu_seqlens(a, True, 0, world)
    pb = api.llama3_flash_attn_prepare_cu_seqlens(b, True, 0, world)
    assert pa[0].tolist() == pb[0].tolist() and pa[1].tolist() == pb[1].tolist()  # identical for rank 0
    assert pa[0]._rfa_llama3[0] != pb[0]._rfa_llama3[0]
    plan_a = api._llama3_plan(0, world, T, tuple(pa[0].tolist()), tuple(pa[1].tolist()), 0, True, pa[0]._rfa_llama3[0])
    plan_b = api._llama3_plan(0, world, T, tuple(pb[0].tolist()), tuple(pb[1].tolist()), 0, True, pb[0]._rfa_llama3[0])
    assert plan_a is not plan_b
    peers_a = api._llama3_peer_plan(pa[0]._rfa_llama3[0], True, 6, world, T)
    peers_b = api._llama3_peer_plan(pb[0]._rfa_llama3[0], True, 6, world, T)
    assert symm._ranges(peers_a, 0) and not symm._ranges(peers_b, 0)  # rank 6 needs rank 0's keys only in layout a
