"""torch.compile around the public functions (the reference's tests have a `compile` mode, test/test.sh:23-25)."""
import torch

import ring_flash_attn_b200 as rfa
from ring_flash_attn_b200.ops.dense import attention_oracle


def test_compiled_wrapper_runs_and_matches():
    torch.manual_seed(0)
    qkv = torch.randn(1, 32, 3, 2, 16, requires_grad=True)

    def f(x):
        return rfa.zigzag_ring_flash_attn_qkvpacked_func(x * 1.0, causal=True) * 2.0

    ref, _ = attention_oracle(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], True)
    out = torch.compile(f, backend="aot_eager")(qkv)  # inductor needs libgomp, absent on the CPU box
    torch.testing.assert_close(out, 2.0 * ref, atol=2e-5, rtol=2e-4)
    out.sum().backward()
    assert qkv.grad is not None and torch.isfinite(qkv.grad).all()
