"""Per-phase timing of the fused multi-GPU path (run under torchrun).  Writes gpurun_out/breakdown_N.json."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ring_flash_attn_b200 as rfa  # noqa: E402
from ring_flash_attn_b200.parallel import symm  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    res = {}
    for (tokens, hq, hkv) in [(4096, 32, 32), (8192, 32, 8)]:
        q = torch.randn(1, tokens, hq, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        k = torch.randn(1, tokens, hkv, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        v = torch.randn(1, tokens, hkv, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        do = torch.randn(1, tokens, hq, 128, device="cuda", dtype=torch.bfloat16)

        def fwd():
            with torch.no_grad():
                return rfa.zigzag_ring_flash_attn_func(q, k, v, causal=True)

        def fwdbwd():
            q.grad = k.grad = v.grad = None
            rfa.zigzag_ring_flash_attn_func(q, k, v, causal=True).backward(do)

        for ctas in [int(x) for x in os.environ.get("SWEEP", "8,16,32").split(",")]:
            ctx = symm.peer_context(None, torch.device("cuda", rank))
            ctx.n_push_ctas = ctas
            key = f"t{tokens}_hq{hq}_hkv{hkv}_push{ctas}"
            res[key] = {"fwd_ms": timeit(fwd), "fwdbwd_ms": timeit(fwdbwd)}
            if rank == 0:
                print(key, res[key], flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/breakdown_{world}.json", "w") as f:
            json.dump(res, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
