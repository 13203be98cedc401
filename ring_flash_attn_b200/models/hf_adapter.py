"""Hugging Face ``transformers`` integration.

Same public surface as the reference adapter (/root/reference/ring_flash_attn/adapters/hf_adapter.py):

* ``substitute_hf_flash_attn(process_group, heads_k_stride)`` - route every flash-attention call of an HF
  model through ``llama3_flash_attn_varlen_func`` over ``process_group`` (hf_adapter.py:361-393);
* ``update_ring_flash_attn_params(cu_seqlens, process_group)`` - once per batch: derive this rank's
  ``cu_seqlens_q/k``, ``max_seqlen_q/k`` and ``local_k_slice`` from the global ``cu_seqlens`` (:42-62);
* ``use_ring_attn(flag)`` - switch back to the stock implementation at run time (:65-67).

Design differences: the module is imported lazily (``import ring_flash_attn_b200`` never needs
``transformers``); instead of four hand-written closures for four generations of
``_flash_attention_forward`` signatures (:70-287) the replacement binds its arguments against whatever
signature the installed ``transformers`` exposes, so new keyword arguments do not break it; and every
module that imported ``_flash_attention_forward`` by name is patched as well (transformers >= 4.48 calls it
from ``transformers.integrations.flash_attention``).
"""
from __future__ import annotations

import inspect
import os
import sys
from typing import Optional

import torch
import torch.distributed as dist

from ..parallel.api import (llama3_flash_attn_prepare_cu_seqlens, llama3_flash_attn_varlen_func,
                            zigzag_llama3_flash_attn_varlen_func)

DATA_PARAMS = {}
RING_ATTN_SWITCH = True
# "llama3" (the reference's layout: every rank feeds a contiguous slice of the packed stream) or "zigzag"
# (extension: rank r feeds chunks r and 2W-1-r, see parallel.layouts.shard_zigzag_llama3 /
# positions_zigzag_llama3 - balanced causal work); chosen in substitute_hf_flash_attn(..., layout=...)
LAYOUT = "llama3"
_ORIGINALS = {}


def update_ring_flash_attn_params(cu_seqlens: torch.Tensor, process_group: Optional[dist.ProcessGroup]):
    """Call once per batch with the *global* cu_seqlens of the packed batch (hf_adapter.py:42-62)."""
    world = dist.get_world_size(group=process_group) if dist.is_initialized() else 1
    rank = dist.get_rank(group=process_group) if dist.is_initialized() else 0
    cu_q, cu_k, max_q, max_k, k_slice = llama3_flash_attn_prepare_cu_seqlens(cu_seqlens, True, rank, world)
    DATA_PARAMS.update(cu_seqlens_q=cu_q, cu_seqlens_k=cu_k, max_seqlen_q=max_q, max_seqlen_k=max_k,
                       local_k_slice=k_slice, cu_seqlens_global=cu_seqlens)


def use_ring_attn(flag: bool) -> None:
    global RING_ATTN_SWITCH
    RING_ATTN_SWITCH = bool(flag)


def _ring_core(query_states, key_states, value_states, *, is_causal, dropout, softmax_scale, sliding_window,
               softcap, deterministic, process_group, heads_k_stride):
    """(1, S_local, H, D) in, (1, S_local, H, D) out - the part shared by every entry point."""
    if softcap is not None:
        raise AssertionError("llama3_flash_attn_varlen_func does not support softcap yet.")
    if not is_causal:
        raise AssertionError("only causal attention is supported for now.")
    if query_states.size(0) != 1:
        raise AssertionError("varlen data should be processed in advance (batch size must be 1).")
    if not DATA_PARAMS:
        raise RuntimeError("call update_ring_flash_attn_params(cu_seqlens, group) before the model forward")
    window = (-1, -1)
    if sliding_window is not None and DATA_PARAMS["max_seqlen_k"] > sliding_window:
        # transformers' convention since 4.4x and in the installed 5.x (modeling_flash_attention_utils.py:
        # window_size = (w - 1, w - 1)): a token sees itself and the w - 1 tokens before it.  The reference adapter
        # passes (w, w) (/root/reference/ring_flash_attn/adapters/hf_adapter.py:121-130), the convention of the
        # transformers releases it was written against - one more key than today's HF models attend to.  The
        # window is applied to global positions in the document, whichever rank holds the keys.
        window = (int(sliding_window) - 1, 0)
    if deterministic is None:
        deterministic = os.environ.get("FLASH_ATTENTION_DETERMINISTIC", "0") == "1"
    if LAYOUT == "zigzag":
        out = zigzag_llama3_flash_attn_varlen_func(
            query_states.squeeze(0), key_states.squeeze(0), value_states.squeeze(0),
            DATA_PARAMS["cu_seqlens_global"], dropout_p=dropout, softmax_scale=softmax_scale, causal=True,
            window_size=window, deterministic=deterministic, group=process_group)
        return out.unsqueeze(0)
    out = llama3_flash_attn_varlen_func(
        query_states.squeeze(0), key_states.squeeze(0), value_states.squeeze(0),
        cu_seqlens_q=DATA_PARAMS["cu_seqlens_q"], cu_seqlens_k=DATA_PARAMS["cu_seqlens_k"],
        max_seqlen_q=DATA_PARAMS["max_seqlen_q"], max_seqlen_k=DATA_PARAMS["max_seqlen_k"],
        heads_k_stride=heads_k_stride, local_k_slice=DATA_PARAMS["local_k_slice"], dropout_p=dropout,
        softmax_scale=softmax_scale, causal=True, window_size=window, deterministic=deterministic,
        group=process_group)
    return out.unsqueeze(0)


def create_ring_flash_attention_forward(process_group, heads_k_stride: int, like=None):
    """A drop-in for ``transformers.modeling_flash_attention_utils._flash_attention_forward``.

    ``like`` is the function being replaced; positional/keyword arguments are bound against *its*
    signature, which covers every signature generation the reference special-cases (hf_adapter.py:74-287)."""
    sig = inspect.signature(like) if like is not None else None

    def _flash_attention_forward(*args, **kwargs):
        if sig is not None:
            try:
                bound = sig.bind_partial(*args, **kwargs).arguments
            except TypeError:
                bound = dict(kwargs)
            extra = bound.pop("kwargs", {}) if "kwargs" in bound else {}
            bound.update(extra)
        else:
            names = ["query_states", "key_states", "value_states", "attention_mask", "query_length", "is_causal",
                     "dropout", "position_ids", "softmax_scale", "sliding_window", "use_top_left_mask", "softcap",
                     "deterministic"]
            bound = dict(zip(names, args))
            bound.update(kwargs)
        q, k, v = bound["query_states"], bound["key_states"], bound["value_states"]
        target_dtype = bound.get("target_dtype")
        if target_dtype is not None and q.dtype == torch.float32:
            q, k, v = q.to(target_dtype), k.to(target_dtype), v.to(target_dtype)
        causal = bound.get("is_causal", True)
        if bound.get("use_top_left_mask", False):
            causal = causal and bound.get("query_length", 2) != 1
        return _ring_core(q, k, v, is_causal=causal, dropout=bound.get("dropout", 0.0) or 0.0,
                          softmax_scale=bound.get("softmax_scale"), sliding_window=bound.get("sliding_window"),
                          softcap=bound.get("softcap"), deterministic=bound.get("deterministic"),
                          process_group=process_group, heads_k_stride=heads_k_stride)

    return _flash_attention_forward


def _make_interface_forward(process_group, heads_k_stride, stock):
    """Entry for ``ALL_ATTENTION_FUNCTIONS["flash_attention_2"]`` (hf_adapter.py:293-358): HF hands us
    (B, H, S, D) tensors; FA-style kernels want (B, S, H, D)."""

    def flash_attention_forward(module, query, key, value, attention_mask, dropout: float = 0.0,
                                scaling: Optional[float] = None, sliding_window: Optional[int] = None,
                                softcap: Optional[float] = None, **kwargs):
        if not RING_ATTN_SWITCH and stock is not None:
            return stock(module, query, key, value, attention_mask, dropout=dropout, scaling=scaling,
                         sliding_window=sliding_window, softcap=softcap, **kwargs)
        query, key, value = query.transpose(1, 2), key.transpose(1, 2), value.transpose(1, 2)
        original_dtype = query.dtype
        if query.dtype == torch.float32 and query.is_cuda:
            # layer norms kept in fp32 (PEFT) silently upcast the hidden states; attention runs in the
            # model's compute dtype
            if torch.is_autocast_enabled():
                target = torch.get_autocast_gpu_dtype()
            elif hasattr(module.config, "_pre_quantization_dtype"):
                target = module.config._pre_quantization_dtype
            else:
                target = next(m for m in module.modules() if isinstance(m, torch.nn.Linear)).weight.dtype
            query, key, value = query.to(target), key.to(target), value.to(target)
        is_causal = kwargs.pop("is_causal", None)
        is_causal = getattr(module, "is_causal", True) if is_causal is None else is_causal
        out = _ring_core(query, key, value, is_causal=is_causal, dropout=dropout, softmax_scale=scaling,
                         sliding_window=sliding_window, softcap=softcap, deterministic=None,
                         process_group=process_group, heads_k_stride=heads_k_stride)
        return out.to(original_dtype), None

    return flash_attention_forward


def substitute_hf_flash_attn(process_group: Optional[dist.ProcessGroup], heads_k_stride: int, layout: str = "llama3"):
    """Patch ``transformers`` so that flash-attention layers run llama3-style context parallelism.

    ``layout="zigzag"`` (extension) selects the balanced flat zigzag layout: feed every rank
    ``layouts.shard_zigzag_llama3(input_ids)`` and ``layouts.positions_zigzag_llama3(cu_seqlens, rank, world)``."""
    global LAYOUT
    if layout not in ("llama3", "zigzag"):
        raise ValueError("layout must be 'llama3' or 'zigzag'")
    LAYOUT = layout
    try:
        import transformers
        import transformers.modeling_flash_attention_utils as fau
    except ImportError as e:  # pragma: no cover
        raise ImportError("substitute_hf_flash_attn needs the `transformers` package") from e

    old = _ORIGINALS.setdefault("_flash_attention_forward", fau._flash_attention_forward)
    new = create_ring_flash_attention_forward(process_group, heads_k_stride, like=old)

    def switchable(*args, **kwargs):
        return new(*args, **kwargs) if RING_ATTN_SWITCH else old(*args, **kwargs)

    switchable.__wrapped__ = old
    fau._flash_attention_forward = switchable
    # modules that did `from ..modeling_flash_attention_utils import _flash_attention_forward`
    # (also modules still bound to the closure of an EARLIER substitute call - recognisable by __wrapped__ - so that
    # a second call with another process group / heads_k_stride / layout takes effect everywhere)
    for name, mod in list(sys.modules.items()):
        if not name.startswith("transformers.") or mod is fau:
            continue
        cur = getattr(mod, "_flash_attention_forward", None)
        if cur is old or (cur is not None and getattr(cur, "__wrapped__", None) is old):
            setattr(mod, "_flash_attention_forward", switchable)

    try:
        from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    except Exception:  # noqa: BLE001 - transformers < 4.48 has no attention interface
        ALL_ATTENTION_FUNCTIONS = None
    if ALL_ATTENTION_FUNCTIONS is not None:
        try:
            stock = _ORIGINALS.setdefault("interface_fa2", ALL_ATTENTION_FUNCTIONS["flash_attention_2"])
        except KeyError:
            stock = None
        ALL_ATTENTION_FUNCTIONS["flash_attention_2"] = _make_interface_forward(process_group, heads_k_stride, stock)
    return transformers.__version__


def restore_hf_flash_attn() -> None:
    """Undo :func:`substitute_hf_flash_attn` (not in the reference; handy for tests)."""
    import transformers.modeling_flash_attention_utils as fau

    global LAYOUT
    LAYOUT = "llama3"

    old = _ORIGINALS.get("_flash_attention_forward")
    if old is not None:
        cur = fau._flash_attention_forward
        fau._flash_attention_forward = old
        for name, mod in list(sys.modules.items()):
            have = getattr(mod, "_flash_attention_forward", None) if name.startswith("transformers.") else None
            if have is not None and (have is cur or getattr(have, "__wrapped__", None) is old):
                setattr(mod, "_flash_attention_forward", old)
    stock = _ORIGINALS.get("interface_fa2")
    if stock is not None:
        from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS

        ALL_ATTENTION_FUNCTIONS["flash_attention_2"] = stock
