"""Concurrency stress of the fused multi-GPU launches (run under torchrun, N >= 2).

The fused forward / backward kernels spin on flags that OTHER ranks' kernels raise, so every rank's launch has to
become resident and make progress while unrelated work shares the GPU.  This script runs ITERS fused fwd+bwd steps
while

* a second stream on every rank issues NCCL all-reduces back to back (what FSDP / TP overlap does to a CP layer),
* ranks are skewed against each other by random host-side sleeps (late arrivals), and
* the shapes alternate between two plans (staging parity / epoch reuse across different layouts),

and checks every step against the first step's result: out / lse / dK / dV must be BITWISE identical run to run
(single writer per tile, fixed-order owner-side sum), dQ equal within fp32 reduction-order noise.  Any watchdog
trap, hang (the caller wraps this in `timeout`) or mismatch is a failure.  Prints one JSON line from rank 0.
"""
import json
import os
import random
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ring_flash_attn_b200 as rfa  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    iters = int(os.environ.get("STRESS_ITERS", "1000"))
    torch.manual_seed(100 + rank)
    rng = random.Random(7 + rank)
    shapes = [(1024, 8, 8), (768, 8, 2)]
    cases = []
    for s_l, hq, hkv in shapes:
        q = torch.randn(1, s_l, hq, 128, device=dev, dtype=torch.bfloat16)
        kv = torch.randn(1, s_l, 2, hkv, 128, device=dev, dtype=torch.bfloat16)
        do = torch.randn(1, s_l, hq, 128, device=dev, dtype=torch.bfloat16)
        cases.append([q, kv, do, None])
    side = torch.cuda.Stream(device=dev)
    noise = torch.randn(4 * 1024 * 1024, device=dev)
    stop = False
    bad = 0
    n_allreduce = 0
    t0 = time.time()
    for it in range(iters):
        q, kv, do, first = cases[it & 1]
        with torch.cuda.stream(side):  # unrelated collectives competing for SMs and NVLink
            for _ in range(2):
                dist.all_reduce(noise)
                n_allreduce += 1
        if rng.random() < 0.05:
            time.sleep(rng.random() * 0.003)  # this rank arrives late
        lq, lkv = q.detach().requires_grad_(True), kv.detach().requires_grad_(True)
        out, lse, _ = rfa.zigzag_ring_flash_attn_kvpacked_func(lq, lkv, causal=True, return_attn_probs=True)
        out.backward(do)
        cur = (out.detach(), lse, lkv.grad, lq.grad)
        if first is None:
            cases[it & 1][3] = tuple(t.clone() for t in cur)
        elif it % 10 == 0 or it > iters - 4:
            names = ("out", "lse", "dkv", "dq")
            for name, a, b in zip(names, cur, first):
                if name == "dq":
                    ok = (a.float() - b.float()).abs().max().item() <= 2e-2 * max(1.0, b.float().abs().max().item())
                else:
                    ok = torch.equal(a, b)
                if not ok:
                    bad += 1
                    print(f"rank {rank} iter {it}: {name} differs from the first run "
                          f"(max abs diff {(a.float() - b.float()).abs().max().item():.3e})", flush=True)
        if bad > 5:
            stop = True
        flag = torch.tensor([1 if stop else 0], device=dev)
        if it % 50 == 49:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag) > 0:
                break
    torch.cuda.synchronize()
    total_bad = torch.tensor([bad], device=dev)
    dist.all_reduce(total_bad)
    if rank == 0:
        print(json.dumps({"stress": "fused fwd+bwd under concurrent NCCL all-reduce + rank skew", "n_gpus": world,
                          "iters": it + 1, "nccl_allreduces_per_rank": n_allreduce, "mismatches": int(total_bad),
                          "seconds": round(time.time() - t0, 1), "ok": int(total_bad) == 0}))
    dist.destroy_process_group()
    if int(total_bad):
        sys.exit(1)


if __name__ == "__main__":
    main()
