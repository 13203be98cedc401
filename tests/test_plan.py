"""Plans must reproduce the true mask on global positions for every scheme, rank and world size."""
import itertools

import pytest
import torch

from ring_flash_attn_b200.ops import plan as P
from ring_flash_attn_b200.parallel import layouts


def _mask_from_plans(plans, pos_of_rank, total):
    """Rebuild the global (q_pos, k_pos) visibility matrix claimed by a set of per-rank plans."""
    m = torch.zeros(total, total, dtype=torch.int32)
    for plan in plans:
        qpos = pos_of_rank[plan.rank]
        for seg in plan.segments:
            ch = plan.q_chunks[seg.chunk]
            kpos = pos_of_rank[seg.src]
            for i in range(ch.rows):
                hi = seg.kv_len if seg.diag is None else min(seg.kv_len, i + seg.diag + 1)
                for j in range(max(hi, 0)):
                    m[qpos[ch.row0 + i], kpos[seg.kv_row0 + j]] += 1
    return m


@pytest.mark.parametrize("scheme,world,causal", [
    ("ring", 1, True), ("ring", 2, True), ("ring", 4, True), ("ring", 3, False),
    ("zigzag", 1, True), ("zigzag", 2, True), ("zigzag", 4, True),
    ("stripe", 1, True), ("stripe", 2, True), ("stripe", 4, True),
])
def test_batch_plans_cover_exact_mask(scheme, world, causal):
    L = 8
    total = L * world
    pos = {r: layouts.positions(scheme, r, world, L).tolist() for r in range(world)}
    plans = []
    for r in range(world):
        if scheme == "ring":
            plans.append(P.plan_ring(r, world, 1, L, causal))
        elif scheme == "zigzag":
            plans.append(P.plan_zigzag(r, world, 1, L))
        else:
            plans.append(P.plan_stripe(r, world, 1, L))
    got = _mask_from_plans(plans, pos, total)
    i = torch.arange(total).unsqueeze(1)
    j = torch.arange(total).unsqueeze(0)
    want = (j <= i).int() if causal else torch.ones(total, total, dtype=torch.int32)
    assert torch.equal(got, want)


@pytest.mark.parametrize("scheme,world,causal", [("ring", 2, True), ("ring", 4, False), ("zigzag", 2, True),
                                                 ("zigzag", 4, True)])
def test_varlen_plans_cover_exact_mask(scheme, world, causal):
    unit = 2 * world
    doc_lens = [unit, 3 * unit, 2 * unit]
    cu = [0] + list(itertools.accumulate(doc_lens))
    total = cu[-1]
    local_cu = [c // world for c in cu]
    ids = torch.arange(total)
    shard = layouts.shard_ring_varlen if scheme == "ring" else layouts.shard_zigzag_varlen
    pos = {r: shard(ids, cu, r, world).tolist() for r in range(world)}
    plans = [P.plan_ring_varlen(r, world, local_cu, causal) if scheme == "ring"
             else P.plan_zigzag_varlen(r, world, local_cu) for r in range(world)]
    got = _mask_from_plans(plans, pos, total)
    want = torch.zeros(total, total, dtype=torch.int32)
    for a, b in zip(cu[:-1], cu[1:]):
        n = b - a
        blk = torch.tril(torch.ones(n, n, dtype=torch.int32)) if causal else torch.ones(n, n, dtype=torch.int32)
        want[a:b, a:b] = blk
    assert torch.equal(got, want)


@pytest.mark.parametrize("world,causal", [(1, True), (2, True), (4, True), (8, True), (4, False)])
def test_llama3_plan_covers_exact_mask(world, causal):
    from ring_flash_attn_b200 import llama3_flash_attn_prepare_cu_seqlens

    cu = [0, 5, 5 + 11, 5 + 11 + 8]
    total = cu[-1]
    assert total % world == 0
    L = total // world
    pos = {r: list(range(r * L, (r + 1) * L)) for r in range(world)}
    plans = []
    for r in range(world):
        cq, ck, mq, mk, ks = llama3_flash_attn_prepare_cu_seqlens(torch.tensor(cu, dtype=torch.int32), causal, r, world)
        plans.append(P.plan_llama3(r, world, L, cq.tolist(), ck.tolist(), ks.start, causal))
    got = _mask_from_plans(plans, pos, total)
    want = torch.zeros(total, total, dtype=torch.int32)
    for a, b in zip(cu[:-1], cu[1:]):
        n = b - a
        want[a:b, a:b] = torch.tril(torch.ones(n, n, dtype=torch.int32)) if causal else 1
    assert torch.equal(got, want)


def test_zigzag_plan_is_balanced():
    world, L = 8, 64
    flops = [P.plan_zigzag(r, world, 1, L).flops(4, 16) for r in range(world)]
    assert max(flops) - min(flops) <= 0.05 * max(flops)
    ring = [P.plan_ring(r, world, 1, L, True).flops(4, 16) for r in range(world)]
    assert max(ring) > 4 * min(ring)  # plain ring is badly imbalanced under a causal mask


def test_visible_area_matches_bruteforce():
    for n, m, d in itertools.product([1, 3, 8], [1, 4, 9], [-9, -3, -1, 0, 1, 4, 20, None]):
        want = sum(1 for i in range(n) for j in range(m) if d is None or j <= i + d)
        assert P.visible_area(n, m, d) == want
