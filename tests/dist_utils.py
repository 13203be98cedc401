"""Spawn a small torch.distributed world inside a test (gloo on CPU, nccl on GPUs)."""
import datetime
import os
import tempfile
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, backend, init_file, fn, args, err_q):
    try:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(rank)
        # a rank that fails an assertion leaves its peers inside a collective: bound that wait (default: 10 min)
        dist.init_process_group(backend, init_method=f"file://{init_file}", rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=120))
        fn(rank, world, *args)
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        err_q.put((rank, traceback.format_exc()))
        raise


def run_distributed(fn, world, *args, backend="gloo"):
    """Run ``fn(rank, world, *args)`` on ``world`` processes; re-raise the first failure."""
    ctx = mp.get_context("spawn")
    err_q = ctx.SimpleQueue()
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rendezvous")
        procs = [ctx.Process(target=_worker, args=(r, world, backend, init_file, fn, args, err_q))
                 for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(600)
        failed = [p for p in procs if p.exitcode != 0]
        for p in procs:
            if p.is_alive():
                p.kill()
        errs = []
        while not err_q.empty():
            errs.append(err_q.get())
        if errs:
            errs.sort()
            raise AssertionError("\n".join(f"rank {r} failed:\n{tb}" for r, tb in errs))
        assert not failed, f"{len(failed)} worker(s) exited abnormally"
