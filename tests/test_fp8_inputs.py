"""Block-scaled fp8 inputs (extension, utils/fp8.py): quantise -> attention == attention on the dequantised values."""
import pytest
import torch

import ring_flash_attn_b200 as rfa
from ring_flash_attn_b200.ops.dense import attention_oracle
from ring_flash_attn_b200.parallel import layouts
from ring_flash_attn_b200.utils import fp8
from dist_utils import run_distributed

pytestmark = pytest.mark.skipif(not fp8.FP8_DTYPES, reason="this torch build has no float8 dtypes")


def test_quantize_dequantize_roundtrip_and_layouts():
    torch.manual_seed(0)
    x = torch.randn(2, 64, 3, 4, 32)
    for block in [(0, 0, 0, 0, 0), (1, 16, 1, 1, 0), (1, 1, 1, 1, 32), (1, 64, 3, 1, 0)]:
        q, d = fp8.quantize_blockwise(x, block)
        assert q.dtype == torch.float8_e4m3fn
        back = fp8.dequantize(q, d, torch.float32)
        assert (back - x).abs().max() < 0.07 * x.abs().max()
    with pytest.raises(ValueError):
        fp8.dequantize(q, None)
    with pytest.raises(ValueError):
        fp8.dequantize(q, torch.ones(2, 5, 1, 1, 1))


def _stripe_case(rank, world):
    # BASELINE.json config 5 in miniature: stripe qkvpacked, fp8 with one scale per (16-token block, head)
    torch.manual_seed(0)
    b, s, h, d = 2, 32 * world, 4, 16
    qkv = torch.randn(b, s, 3, h, d)
    import torch.distributed as dist

    dist.broadcast(qkv, src=0)
    local = layouts.shard_stripe(qkv, rank, world)
    q8, descale = fp8.quantize_blockwise(local, (1, 16, 1, 1, 0))
    out = rfa.stripe_flash_attn_qkvpacked_func(q8, causal=True, descale=descale)
    # oracle on the dequantised values of EVERY rank (what the other ranks contribute is their fp8 data too)
    shards = []
    for r in range(world):
        lr = layouts.shard_stripe(qkv, r, world)
        qr, dr = fp8.quantize_blockwise(lr, (1, 16, 1, 1, 0))
        shards.append(fp8.dequantize(qr, dr, torch.float32))
    full = layouts.unshard("stripe", shards, dim=1)
    ref, _ = attention_oracle(full[:, :, 0].bfloat16(), full[:, :, 1].bfloat16(), full[:, :, 2].bfloat16(), True)
    torch.testing.assert_close(out.float(), layouts.shard_stripe(ref, rank, world), atol=3e-2, rtol=3e-2)
    assert out.dtype == torch.bfloat16


@pytest.mark.parametrize("world", [1, 2])
def test_stripe_fp8_blockscaled(world):
    run_distributed(_stripe_case, world)


def test_kvpacked_and_unpacked_fp8_share_semantics():
    torch.manual_seed(1)
    q = torch.randn(1, 32, 4, 16)
    kv = torch.randn(1, 32, 2, 2, 16)
    q8, dq = fp8.quantize_blockwise(q, (1, 8, 1, 0))
    kv8, dkv = fp8.quantize_blockwise(kv, (1, 8, 1, 1, 0))
    a = rfa.ring_flash_attn_kvpacked_func(q8, kv8, causal=True, descale=(dq, dkv))
    b = rfa.ring_flash_attn_func(q8, kv8[:, :, 0], kv8[:, :, 1], causal=True, descale=(dq, dkv[:, :, 0], dkv[:, :, 1]))
    torch.testing.assert_close(a, b)
    with pytest.raises(ValueError):
        rfa.ring_flash_attn_kvpacked_func(q8, kv8, causal=True)


def test_fp8_kernel_opt_in_is_ignored_off_gpu(monkeypatch):
    """RFA_B200_FP8_KERNEL=1 only changes calls that can run on the kernels; CPU tensors keep the dequantise path."""
    import ring_flash_attn_b200 as rfa
    from ring_flash_attn_b200.ops.dense import attention_oracle
    from ring_flash_attn_b200.parallel import api
    from ring_flash_attn_b200.utils import fp8

    monkeypatch.setenv("RFA_B200_FP8_KERNEL", "1")
    torch.manual_seed(0)
    qkv = torch.randn(1, 64, 3, 2, 128)
    q8, d = fp8.quantize_blockwise(qkv, [0, 0, 1, 1, 0])
    deq = fp8.dequantize(q8, d, torch.float32)
    ref, _ = attention_oracle(deq[:, :, 0], deq[:, :, 1], deq[:, :, 2], True)
    out = rfa.ring_flash_attn_qkvpacked_func(q8, causal=True, descale=d)
    torch.testing.assert_close(out.float(), ref, atol=3e-2, rtol=3e-2)
    # canonical (blocks, heads) descale tables (what the kernel path consumes)
    x = torch.zeros(2, 8, 4, 128)
    assert api._scale_table(None, x).tolist() == [[1.0] * 4]
    assert api._scale_table(0.5, x).tolist() == [[0.5] * 4]
    assert api._scale_table(torch.tensor(2.0), x).tolist() == [[2.0] * 4]
    assert api._scale_table(torch.arange(4.0).view(1, 1, 4, 1), x).tolist() == [[0.0, 1.0, 2.0, 3.0]]
    blk = api._scale_table(torch.arange(16.0).view(2, 2, 4, 1), x)  # 4-token blocks x head, per batch element
    assert blk.shape == (4, 4) and blk[1].tolist() == [4.0, 5.0, 6.0, 7.0] and blk[2].tolist() == [8.0, 9.0, 10.0, 11.0]
    shared = api._scale_table(torch.arange(2.0).view(1, 2, 1, 1), x)  # the same token-block scales for every batch / head
    assert shared.shape == (4, 4) and shared[:, 0].tolist() == [0.0, 1.0, 0.0, 1.0]
    assert api._scale_table(torch.ones(2, 8, 4, 4), x) is None  # scales along head_dim (MX style): dequantise path
    assert api._scale_table(torch.ones(1, 3, 4, 1), x) is None  # does not tile the tokens
    packed = torch.zeros(12, 4, 128)
    assert api._scale_table(torch.arange(12.0).view(3, 4, 1), packed).shape == (3, 4)


def test_quantize_per_head_shapes_and_roundtrip():
    from ring_flash_attn_b200.parallel import api
    from ring_flash_attn_b200.utils import fp8

    torch.manual_seed(0)
    q = torch.randn(2, 16, 4, 128)
    kv = torch.randn(2, 16, 2, 2, 128)
    q8, dq = fp8.quantize_per_head(q)
    kv8, dkv = fp8.quantize_per_head(kv, keep_dims=(2,))
    assert dq.shape == (1, 1, 4, 1) and dkv.shape == (1, 1, 2, 2, 1)
    assert (fp8.dequantize(q8, dq, torch.float32) - q).abs().max() < 0.07 * q.abs().max()
    # exactly the granularity the kernel path accepts
    assert api._scale_table(dq, q8).shape == (1, 4)
    dk, dv = api._split_descale(dkv, 2, 2)
    assert api._scale_table(dk, kv8[:, :, 0]).shape == (1, 2) and api._scale_table(dv, kv8[:, :, 1]).shape == (1, 2)

