"""Tiny coverage driver for compute-sanitizer (1 GPU): one forward + backward of every kernel instantiation family -
bf16 d=128, fp16 d=64, sliding window, packed varlen with ragged tiles, block-scaled fp8 forward.

    compute-sanitizer --tool memcheck python benchmark/sanitize_target.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ring_flash_attn_b200 as rfa  # noqa: E402
from ring_flash_attn_b200.utils import fp8  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def fb(fn, *tensors, dout):
    out = fn(*tensors)
    out.backward(dout)
    torch.cuda.synchronize()
    return out


q = torch.randn(1, 640, 3, 4, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
fb(lambda x: rfa.zigzag_ring_flash_attn_qkvpacked_func(x, causal=True), q, dout=torch.randn(1, 640, 4, 128, device=dev).to(torch.bfloat16))
print("bf16 d=128 ok", flush=True)
q64 = torch.randn(1, 500, 3, 4, 64, device=dev).to(torch.float16).requires_grad_(True)
fb(lambda x: rfa.ring_flash_attn_qkvpacked_func(x, causal=True), q64, dout=torch.randn(1, 500, 4, 64, device=dev).to(torch.float16))
print("fp16 d=64 ok", flush=True)
fb(lambda x: rfa.ring_flash_attn_qkvpacked_func(x, causal=True, window_size=(100, 0)), q,
   dout=torch.randn(1, 640, 4, 128, device=dev).to(torch.bfloat16))
print("window ok", flush=True)
cu = torch.tensor([0, 1, 130, 131, 700], dtype=torch.int32, device=dev)
qv = torch.randn(700, 8, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
kv = torch.randn(700, 2, 2, 128, device=dev).to(torch.bfloat16).requires_grad_(True)
fb(lambda a, b: rfa.zigzag_llama3_flash_attn_varlen_kvpacked_func(a, b, cu, causal=True), qv, kv,
   dout=torch.randn(700, 8, 128, device=dev).to(torch.bfloat16))
print("varlen GQA ok", flush=True)
x = torch.randn(1, 512, 3, 4, 128, device=dev)
x8, scale = fp8.quantize_blockwise(x, [1, 128, 1, 1, 0])
rfa.stripe_flash_attn_qkvpacked_func(x8, causal=True, descale=scale)
torch.cuda.synchronize()
print("fp8 block-scaled ok", flush=True)
