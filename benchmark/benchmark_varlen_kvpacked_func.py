"""Varlen benchmark: iter/s of ring / zigzag / llama3 packed attention (reference:
benchmark/benchmark_varlen_kvpacked_func.py - four packings of 8192 local tokens are cycled; llama3 uses
heads_k_stride=4).

    torchrun --nproc-per-node 8 benchmark/benchmark_varlen_kvpacked_func.py [--forward-only]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ring_flash_attn_b200 as rfa  # noqa: E402


def packings(tokens):
    u = tokens // 8
    return [[0, tokens], [0, 2 * u, 4 * u, tokens], [0, u, 2 * u, 3 * u, 4 * u, 5 * u, 6 * u, 7 * u, tokens],
            [0, 5 * u, tokens]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=8192, help="tokens per GPU")
    ap.add_argument("--nheads", type=int, default=32)
    ap.add_argument("--nheads-k", type=int, default=8)
    ap.add_argument("--forward-only", action="store_true")
    ap.add_argument("--num-iter", type=int, default=0)
    ap.add_argument("--heads-k-stride", type=int, default=4)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    T = args.tokens
    num_iter = args.num_iter or (500 if args.forward_only else 100)
    q = torch.randn(T, args.nheads, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
    kv = torch.randn(T, 2, args.nheads_k, 128, device=dev, dtype=torch.bfloat16, requires_grad=True)
    dout = torch.randn(T, args.nheads, 128, device=dev, dtype=torch.bfloat16)
    local_cus = [torch.tensor(p, dtype=torch.int32, device=dev) for p in packings(T)]
    global_cus = [c * world for c in local_cus]
    llama3_args = [rfa.llama3_flash_attn_prepare_cu_seqlens(c, True, rank, world) for c in global_cus]

    def make(name):
        def run(i):
            j = i % len(local_cus)
            if name == "llama3":
                cq, ck, mq, mk, ks = llama3_args[j]
                return rfa.llama3_flash_attn_varlen_kvpacked_func(q, kv, cq, ck, mq, mk, heads_k_stride=args.heads_k_stride,
                                                                 local_k_slice=ks, causal=True)
            if name == "zigzag_llama3":  # beyond the reference: flat zigzag, global cu_seqlens
                return rfa.zigzag_llama3_flash_attn_varlen_kvpacked_func(q, kv, global_cus[j], causal=True)
            cu = local_cus[j]
            mx = int((cu[1:] - cu[:-1]).max())
            fn = rfa.ring_flash_attn_varlen_kvpacked_func if name == "ring" else rfa.zigzag_ring_flash_attn_varlen_kvpacked_func
            return fn(q, kv, cu, mx, causal=True)
        return run

    res = {}
    for name in ("ring", "zigzag_ring", "llama3", "zigzag_llama3"):
        run = make(name)

        def step(i):
            if args.forward_only:
                with torch.no_grad():
                    run(i)
            else:
                q.grad = kv.grad = None
                run(i).backward(dout)

        for i in range(8):
            step(i)
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(num_iter):
            step(i)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = num_iter / float(t[0])
        if rank == 0:
            print(f"{name:12s} {res[name]:9.2f} iter/s", flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        mode = "fwd" if args.forward_only else "fwdbwd"
        with open(f"gpurun_out/bench_varlen_{mode}_{world}.json", "w") as f:
            json.dump({"world": world, "mode": mode, "tokens_per_gpu": T, "iter_per_s": res}, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
