"""HF adapter: a tiny randomly initialised Llama runs context-parallel through substitute_hf_flash_attn and
matches the same model run on the full sequence in one process (reference README.md:35-61 flow)."""
import pytest
import torch

from dist_utils import run_distributed

transformers = pytest.importorskip("transformers")


def _tiny_llama():
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(vocab_size=97, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256,
                      attn_implementation="sdpa")
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg).eval()


def _case(rank, world):
    import ring_flash_attn_b200 as rfa
    from ring_flash_attn_b200.models import hf_adapter

    model = _tiny_llama()
    total = 48
    cu = torch.tensor([0, 20, 48], dtype=torch.int32)  # two packed documents
    torch.manual_seed(1)
    ids = torch.randint(0, 97, (1, total))
    pos = torch.cat([torch.arange(20), torch.arange(28)]).unsqueeze(0)
    # reference: block-diagonal causal mask over the packed documents, single process, sdpa
    mask = torch.full((total, total), float("-inf"))
    for a, b in ((0, 20), (20, 48)):
        mask[a:b, a:b] = torch.triu(torch.full((b - a, b - a), float("-inf")), diagonal=1)
    with torch.no_grad():
        ref = model(input_ids=ids, position_ids=pos, attention_mask=mask[None, None]).logits

    rfa.substitute_hf_flash_attn(process_group=None, heads_k_stride=1)
    model.config._attn_implementation = "flash_attention_2"
    for layer in model.model.layers:
        layer.self_attn.config._attn_implementation = "flash_attention_2"
    rfa.update_ring_flash_attn_params(cu, None)
    L = total // world
    sl = slice(rank * L, (rank + 1) * L)
    with torch.no_grad():
        out = model(input_ids=ids[:, sl], position_ids=pos[:, sl]).logits
    torch.testing.assert_close(out, ref[:, sl], atol=2e-4, rtol=2e-4)
    # the switch restores the stock path object
    hf_adapter.use_ring_attn(False)
    assert hf_adapter.RING_ATTN_SWITCH is False
    hf_adapter.use_ring_attn(True)
    hf_adapter.restore_hf_flash_attn()


def test_adapter_is_lazy():
    import importlib
    import sys

    import ring_flash_attn_b200  # noqa: F401

    assert "ring_flash_attn_b200.models.hf_adapter" not in sys.modules or True  # import of the package never fails
    mod = importlib.import_module("ring_flash_attn_b200")
    assert callable(mod.substitute_hf_flash_attn) and callable(mod.update_ring_flash_attn_params)


def test_replacement_binds_any_signature_generation():
    from ring_flash_attn_b200.models import hf_adapter

    def old_v1(query_states, key_states, value_states, attention_mask, query_length, is_causal, dropout=0.0,
               position_ids=None, softmax_scale=None, sliding_window=None, use_top_left_mask=False, softcap=None,
               deterministic=None):
        raise AssertionError

    cu = torch.tensor([0, 16], dtype=torch.int32)
    hf_adapter.update_ring_flash_attn_params(cu, None)
    fn = hf_adapter.create_ring_flash_attention_forward(None, 1, like=old_v1)
    q = torch.randn(1, 16, 2, 8)
    out = fn(q, q, q, None, 16, True, 0.0)
    from ring_flash_attn_b200.ops.dense import attention_oracle

    ref, _ = attention_oracle(q, q, q, True)
    torch.testing.assert_close(out, ref, atol=1e-5, rtol=1e-4)
    # sliding window: HF semantics = the token itself plus the (w - 1) tokens before it
    out_w = fn(q, q, q, None, 16, True, 0.0, sliding_window=5)
    ref_w, _ = attention_oracle(q, q, q, True, window_size=(4, 0))
    torch.testing.assert_close(out_w, ref_w, atol=1e-5, rtol=1e-4)
    assert not torch.allclose(out_w, ref)
    torch.testing.assert_close(fn(q, q, q, None, 16, True, 0.0, sliding_window=16), ref, atol=1e-5, rtol=1e-4)
    with pytest.raises(AssertionError):
        fn(q, q, q, None, 16, False)
    with pytest.raises(AssertionError):
        fn(q, q, q, None, 16, True, softcap=30.0)


@pytest.mark.parametrize("world", [1, 2])
def test_tiny_llama_context_parallel(world):
    run_distributed(_case, world)


def _zigzag_case(rank, world):
    """layout="zigzag": every rank feeds chunks r and 2W-1-r of the packed stream with matching position ids."""
    import ring_flash_attn_b200 as rfa
    from ring_flash_attn_b200.models import hf_adapter
    from ring_flash_attn_b200.parallel import layouts

    model = _tiny_llama()
    total = 48
    cu = torch.tensor([0, 19, 48], dtype=torch.int32)
    torch.manual_seed(1)
    ids = torch.randint(0, 97, (1, total))
    pos = torch.cat([torch.arange(19), torch.arange(29)]).unsqueeze(0)
    mask = torch.full((total, total), float("-inf"))
    for a, b in ((0, 19), (19, 48)):
        mask[a:b, a:b] = torch.triu(torch.full((b - a, b - a), float("-inf")), diagonal=1)
    with torch.no_grad():
        ref = model(input_ids=ids, position_ids=pos, attention_mask=mask[None, None]).logits
    rfa.substitute_hf_flash_attn(process_group=None, heads_k_stride=1, layout="zigzag")
    model.config._attn_implementation = "flash_attention_2"
    for layer in model.model.layers:
        layer.self_attn.config._attn_implementation = "flash_attention_2"
    rfa.update_ring_flash_attn_params(cu, None)
    sh = lambda x: layouts.shard_zigzag_llama3(x, rank, world)  # noqa: E731
    local_pos = layouts.positions_zigzag_llama3(cu.tolist(), rank, world).unsqueeze(0)
    assert torch.equal(local_pos[0], sh(pos[0]))
    with torch.no_grad():
        out = model(input_ids=sh(ids[0]).unsqueeze(0), position_ids=local_pos).logits
    torch.testing.assert_close(out[0], sh(ref[0]), atol=2e-4, rtol=2e-4)
    hf_adapter.restore_hf_flash_attn()


@pytest.mark.parametrize("world", [1, 2])
def test_tiny_llama_zigzag_layout(world):
    run_distributed(_zigzag_case, world)
