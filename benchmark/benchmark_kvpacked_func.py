"""Batch-layout benchmark: iter/s of ring / zigzag / stripe kvpacked attention vs single-GPU flash attention.

Capability parity with the reference's benchmark/benchmark_kvpacked_func.py (Llama-3.1-8B attention shape:
batch 1, 8192 tokens per GPU, 32 query / 8 kv heads, head_dim 128, bf16, causal; CUDA-event timing), with
the measurement rules this repo uses everywhere: device time, max over ranks, warm-up first.

    torchrun --nproc-per-node 8 benchmark/benchmark_kvpacked_func.py [--forward-only] [--profile]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ring_flash_attn_b200 as rfa  # noqa: E402


def local_ours(q, kv, causal=True, **_):
    """Our own kernel at world size 1 on the local shard, i.e. plain causal flash attention."""
    return rfa.zigzag_ring_flash_attn_kvpacked_func(q, kv, causal=causal, group=LOCAL_GROUP)


def local_flash_attn(q, kv, causal=True, **_):
    """The reference's definition of "theoretic flash_attn" (/root/reference/README.md:103,
    benchmark/benchmark_kvpacked_func.py:140-147): the flash_attn LIBRARY's kvpacked function on the local shard,
    divided by the world size."""
    from flash_attn import flash_attn_kvpacked_func

    return flash_attn_kvpacked_func(q, kv, causal=causal)


LOCAL_GROUP = None


def benchmark(fn, args, num_iter, forward_only, profile_dir=None):
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    q = torch.randn(1, args.tokens, args.nheads, args.d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    kv = torch.randn(1, args.tokens, 2, args.nheads_k, args.d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    dout = torch.randn(1, args.tokens, args.nheads, args.d, device=dev, dtype=torch.bfloat16)

    def step():
        if forward_only:
            with torch.no_grad():
                fn(q, kv, causal=True)
        else:
            q.grad = kv.grad = None
            fn(q, kv, causal=True).backward(dout)

    prof = None
    if profile_dir:
        prof = torch.profiler.profile(
            activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
            schedule=torch.profiler.schedule(wait=5, warmup=5, active=5), record_shapes=True, with_stack=True,
            on_trace_ready=torch.profiler.tensorboard_trace_handler(os.path.join(profile_dir, f"rank_{rank}")))
        prof.start()
    for _ in range(max(3, num_iter // 10)):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(num_iter):
        step()
        if prof:
            prof.step()
    b.record()
    torch.cuda.synchronize()
    if prof:
        prof.stop()
    t = torch.tensor([a.elapsed_time(b) / 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return num_iter / float(t[0])


def main():
    global LOCAL_GROUP
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=8192, help="tokens per GPU")
    ap.add_argument("--nheads", type=int, default=32)
    ap.add_argument("--nheads-k", type=int, default=8)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--forward-only", action="store_true")
    ap.add_argument("--num-iter", type=int, default=0)
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank % torch.cuda.device_count()))
    world = dist.get_world_size()
    groups = [dist.new_group([r]) for r in range(world)] if world > 1 else []
    if world > 1:
        LOCAL_GROUP = groups[rank]
    num_iter = args.num_iter or (500 if args.forward_only else 100)
    res = {}
    rows = [("ours(local)", local_ours), ("ring", rfa.ring_flash_attn_kvpacked_func),
            ("zigzag_ring", rfa.zigzag_ring_flash_attn_kvpacked_func), ("stripe", rfa.stripe_flash_attn_kvpacked_func)]
    try:
        import flash_attn  # noqa: F401

        rows.insert(0, ("flash_attn(local)", local_flash_attn))
    except Exception as e:  # noqa: BLE001 - the library is optional; then only our own world-1 kernel is the yardstick
        if rank == 0:
            print(f"flash_attn unavailable ({type(e).__name__}): '% of theoretic' is relative to our own world-1 kernel")
    for name, fn in rows:
        prof_dir = os.path.join("benchmark", "logs", name) if args.profile else None
        its = benchmark(fn, args, num_iter, args.forward_only, prof_dir)
        res[name] = its
        if rank == 0:
            if name.endswith("(local)"):
                extra = f"  (/ {world} = {its / world:.1f} iter/s)"
            else:
                extra = "".join(f"  {100 * its / (res[b] / world):.1f}% of {b} / {world}"
                                for b in ("flash_attn(local)", "ours(local)") if b in res)
            print(f"{name:18s} {its:9.2f} iter/s{extra}", flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        mode = "fwd" if args.forward_only else "fwdbwd"
        with open(f"gpurun_out/bench_kvpacked_{mode}_{world}.json", "w") as f:
            json.dump({"world": world, "mode": mode, "tokens_per_gpu": args.tokens, "iter_per_s": res}, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
