"""Tiny driver for ncu captures: a few fwd+bwd iterations of causal attention at the README shape (1 GPU)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ring_flash_attn_b200 as rfa  # noqa: E402

S, HQ, HKV = int(os.environ.get("S", 8192)), int(os.environ.get("HQ", 32)), int(os.environ.get("HKV", 8))
q = torch.randn(1, S, HQ, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
k = torch.randn(1, S, HKV, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
v = torch.randn(1, S, HKV, 128, device="cuda", dtype=torch.bfloat16, requires_grad=True)
do = torch.randn(1, S, HQ, 128, device="cuda", dtype=torch.bfloat16)
for _ in range(int(os.environ.get("ITERS", 3))):
    q.grad = k.grad = v.grad = None
    rfa.zigzag_ring_flash_attn_func(q, k, v, causal=True).backward(do)
torch.cuda.synchronize()
print("done")
