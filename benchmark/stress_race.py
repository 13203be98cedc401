"""Run-to-run consistency stress: the forward and dK/dV are deterministic by construction, so any difference
between repeated runs on identical inputs is a race.  Reports which tensor and where."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ring_flash_attn_b200.ops import attn_cuda, plan as P  # noqa: E402
from ring_flash_attn_b200.ops.dense import block_bwd, block_fwd  # noqa: E402

torch.manual_seed(0)
bad = 0
for (sq, sk, hq, hkv, diag) in [(640, 640, 4, 4, 0), (500, 500, 8, 2, None), (1024, 1024, 2, 1, 0), (300, 700, 4, 2, 100)]:
    q = torch.randn(sq, hq, 128, device="cuda").to(torch.bfloat16)
    k = torch.randn(sk, hkv, 128, device="cuda").to(torch.bfloat16)
    v = torch.randn(sk, hkv, 128, device="cuda").to(torch.bfloat16)
    do = torch.randn(sq, hq, 128, device="cuda").to(torch.bfloat16)
    scale = 1 / math.sqrt(128)
    plan = P.CPPlan(1, 0, sq, sk, [P.QChunk(0, sq)], [P.Segment(0, 0, 0, sk, diag)])
    ref_out, ref_lse = block_fwd(q, k, v, scale, diag)
    delta = (ref_out * do.float()).sum(-1).transpose(0, 1).contiguous()
    ref_dq, ref_dk, ref_dv = block_bwd(do, q, k, v, ref_lse, delta, scale, diag)
    first = None
    for it in range(int(os.environ.get("STRESS_ITERS", "150"))):
        out, lse = attn_cuda.segments_forward(plan, plan.segments, q, k, v, scale)
        dq = torch.zeros(sq, hq, 128, device="cuda")
        dk = torch.zeros(sk, hkv, 128, device="cuda")
        dv = torch.zeros(sk, hkv, 128, device="cuda")
        attn_cuda.segments_backward(plan, plan.segments, do, q, k, v, ref_lse, delta, scale, dq, dk, dv)
        cur = {"out": out.float(), "lse": lse, "dq": dq, "dk": dk, "dv": dv}
        refs = {"out": ref_out, "lse": ref_lse, "dq": ref_dq, "dk": ref_dk, "dv": ref_dv}
        for name, t in cur.items():
            err = (t - refs[name]).abs()
            lim = 3e-2 * max(1.0, refs[name].abs().max().item())
            if err.max().item() > lim or not torch.isfinite(t).all():
                idx = torch.nonzero(err > lim)
                rows = sorted(set(idx[:, 0 if name != "lse" else 1].tolist()))
                heads = sorted(set(idx[:, 1 if name != "lse" else 0].tolist()))
                print(f"MISMATCH shape={(sq, sk, hq, hkv, diag)} iter={it} {name}: max err {err.max().item():.4f} "
                      f"n_bad={idx.shape[0]} rows[{rows[0]}..{rows[-1]}] (#{len(rows)}) heads={heads} "
                      f"cols={sorted(set((idx[:, 2] // 32 * 32).tolist())) if name != 'lse' else ''}")
                bad += 1
        if bad > 6:
            break
    print("shape", (sq, sk, hq, hkv, diag), "done, bad so far", bad, flush=True)
print("TOTAL BAD", bad)
