"""Sampled correctness witness for a context-parallel attention step at FULL problem size.

A dense fp32 oracle of the whole problem is out of reach at benchmark shapes (S = 32768, 32 heads), so the
check recomputes in fp32, from all-gathered inputs, exactly the quantities of

* a random sample of this rank's query rows: ``out``, ``lse`` and ``dq`` (needs every rank's K / V only), and
* one slice of this rank's key rows: ``dk`` and ``dv`` (needs every rank's Q / dO / out / lse),

for a few heads, and compares them with what the library produced.  ``bench.py --check`` prints the result
inside its JSON line so that every timed number comes with a correctness witness at the same shape and
world size; ``tests/test_verify.py`` runs the same code on gloo.

The reference has no such tool: its tests print max / mean errors of small shapes for a human to read
(/root/reference/test/utils.py:15-38, test/test_zigzag_ring_flash_attn_func.py:60-92).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from ..parallel import layouts
from ..parallel.comm import group_info


def _gather(x: torch.Tensor, group, world: int) -> List[torch.Tensor]:
    x = x.contiguous()
    if world == 1:
        return [x]
    outs = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(outs, x, group=group)
    return outs


def sampled_check(scheme: str, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, dout: Optional[torch.Tensor],
                  out: torch.Tensor, lse: Optional[torch.Tensor], dq: Optional[torch.Tensor],
                  dk: Optional[torch.Tensor], dv: Optional[torch.Tensor], group=None, softmax_scale=None,
                  kv_heads: Optional[Sequence[int]] = None, n_rows: int = 256, key_rows: int = 128,
                  seed: int = 0) -> Dict[str, object]:
    """Compare one causal CP step against an fp32 oracle on sampled rows.

    Local tensors of ONE batch element in the scheme's layout: q / dout / out / dq ``(S_l, Hq, D)``,
    k / v / dk / dv ``(S_l, Hkv, D)``, lse ``(Hq, S_l)``.  ``scheme`` is ``ring`` / ``zigzag`` / ``stripe``.
    Collective over ``group`` (every rank must call it).  Returns max abs errors, the tolerances used and ``ok``.
    Gradient entries are skipped when ``dout`` is None (forward-only steps).
    """
    rank, world = group_info(group)
    s_l, hq, d = q.shape
    hkv = k.shape[1]
    rep = hq // hkv
    if kv_heads is None:
        kv_heads = sorted({0, hkv - 1})
    kvh = list(kv_heads)
    qh = [h * rep + i for h in kvh for i in range(rep)]
    scale = d ** -0.5 if softmax_scale is None else float(softmax_scale)
    dev = q.device
    with_grad = dout is not None

    def glob(x, heads):  # local (S_l, H, D) -> global order (S, len(heads), D) fp32
        return layouts.unshard(scheme, _gather(x[:, heads], group, world), dim=0).float()

    kg, vg = glob(k, kvh), glob(v, kvh)
    pos_local = layouts.positions(scheme, rank, world, s_l, device=dev)
    pos_global = layouts.unshard(scheme, [layouts.positions(scheme, r, world, s_l, device=dev) for r in range(world)],
                                 dim=0)
    gen = torch.Generator(device="cpu").manual_seed(seed + 7919 * rank)
    rows = torch.randperm(s_l, generator=gen)[:min(n_rows, s_l)].sort().values.to(dev)
    res: Dict[str, object] = {"rows": int(rows.numel()), "heads_q": qh, "world": world}
    err: Dict[str, float] = {}
    tol: Dict[str, float] = {}

    def note(name, got, ref, rel):
        got, ref = got.detach(), ref.detach().float()
        err[name] = float((got.float() - ref).abs().max())
        finite = ref[torch.isfinite(ref)]
        tol[name] = rel * float(finite.abs().max() if finite.numel() else 1.0) + 2e-3

    # ---- sampled query rows: out, lse, dq ------------------------------------------------------------
    qr = q[rows][:, qh].float().transpose(0, 1)                        # (nq, R, D)
    kx = kg.repeat_interleave(rep, dim=1).transpose(0, 1)              # (nq, S, D)
    vx = vg.repeat_interleave(rep, dim=1).transpose(0, 1)
    s = torch.matmul(qr, kx.transpose(1, 2)) * scale                   # (nq, R, S)
    vis = pos_global.unsqueeze(0) <= pos_local[rows].unsqueeze(1)      # (R, S) causal on global positions
    s = s.masked_fill(~vis.unsqueeze(0), float("-inf"))
    lse_ref = torch.logsumexp(s, dim=-1)                               # (nq, R)
    p = torch.exp(s - lse_ref.unsqueeze(-1))
    out_ref = torch.matmul(p, vx)                                      # (nq, R, D)
    note("out", out[rows][:, qh].transpose(0, 1), out_ref, 2e-2)
    if lse is not None:
        note("lse", lse[qh][:, rows], lse_ref, 2e-3)
    if with_grad:
        dor = dout[rows][:, qh].float().transpose(0, 1)
        delta = (dor * out_ref).sum(-1, keepdim=True)
        ds = p * (torch.matmul(dor, vx.transpose(1, 2)) - delta) * scale
        if dq is not None:
            note("dq", dq[rows][:, qh].transpose(0, 1), torch.matmul(ds, kx), 3e-2)
    del s, p

    # ---- one slice of local key rows: dk, dv ---------------------------------------------------------
    if with_grad and dk is not None and dv is not None:
        j0 = (s_l // 2 // 128) * 128 if s_l >= 256 else 0  # a slice in the middle of the shard (tile aligned)
        keys = torch.arange(j0, min(j0 + key_rows, s_l), device=dev)
        qg, dog, og = glob(q, qh), glob(dout, qh), glob(out, qh)                      # (S, nq, D)
        lse_g = layouts.unshard(scheme, _gather(lse[qh].transpose(0, 1), group, world), dim=0).float()  # (S, nq)
        ks = k[keys][:, kvh].float().repeat_interleave(rep, dim=1).transpose(0, 1)    # (nq, J, D)
        vs = v[keys][:, kvh].float().repeat_interleave(rep, dim=1).transpose(0, 1)
        qx, dox = qg.transpose(0, 1), dog.transpose(0, 1)                             # (nq, S, D)
        sc = torch.matmul(qx, ks.transpose(1, 2)) * scale                             # (nq, S, J)
        visk = pos_local[keys].unsqueeze(0) <= pos_global.unsqueeze(1)                # (S, J)
        lse_t = lse_g.transpose(0, 1).unsqueeze(-1)
        pk = torch.exp(sc - torch.where(torch.isinf(lse_t), torch.zeros_like(lse_t), lse_t))
        pk = pk.masked_fill(~visk.unsqueeze(0), 0.0)
        dv_ref = torch.matmul(pk.transpose(1, 2), dox)                                # (nq, J, D)
        delta_g = (dog * og).sum(-1).transpose(0, 1).unsqueeze(-1)                    # (nq, S, 1)
        dsk = pk * (torch.matmul(dox, vs.transpose(1, 2)) - delta_g) * scale
        dk_ref = torch.matmul(dsk.transpose(1, 2), qx)
        n = len(kvh)
        dk_ref = dk_ref.reshape(n, rep, keys.numel(), d).sum(1)                       # sum over the GQA group
        dv_ref = dv_ref.reshape(n, rep, keys.numel(), d).sum(1)
        note("dk", dk[keys][:, kvh].transpose(0, 1), dk_ref, 3e-2)
        note("dv", dv[keys][:, kvh].transpose(0, 1), dv_ref, 3e-2)
        res["key_rows"] = [int(keys[0]), int(keys[-1]) + 1]

    # worst case over the ranks
    names = sorted(err)
    t = torch.tensor([err[n] / tol[n] for n in names] + [err[n] for n in names], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    k_ = len(names)
    res["max_err"] = {n: float(t[k_ + i]) for i, n in enumerate(names)}
    res["tol"] = {n: tol[n] for n in names}
    res["worst_err_over_tol"] = float(t[:k_].max()) if k_ else 0.0
    res["ok"] = bool(res["worst_err_over_tol"] <= 1.0) and all(torch.isfinite(t[k_:]).tolist())
    return res
