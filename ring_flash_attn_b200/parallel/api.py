"""Public attention entry points - the 18 functions of the reference plus ``prepare``.

Signatures, defaults and return values follow the reference exactly
(/root/reference/ring_flash_attn/__init__.py:1-35; batch: ring_flash_attn.py:223-301,
zigzag_ring_flash_attn.py:268-346, stripe_flash_attn.py:300-378; varlen:
ring_flash_attn_varlen.py:268-358, zigzag_ring_flash_attn_varlen.py:415-505,
llama3_flash_attn_varlen.py:390-504) so a user can switch by changing the import.

Unlike the reference there is a single autograd bridge: each scheme only contributes a *plan*
(``ops/plan.py``); the engine (``parallel/engine.py``) runs it on the fused sm_100a path or on the
torch.distributed fallback.
"""
from __future__ import annotations

import functools
import os
from typing import Optional, Tuple

import torch

from ..ops import plan as P
from . import engine
from .comm import group_info

__all__ = [
    "ring_flash_attn_func", "ring_flash_attn_kvpacked_func", "ring_flash_attn_qkvpacked_func",
    "zigzag_ring_flash_attn_func", "zigzag_ring_flash_attn_kvpacked_func", "zigzag_ring_flash_attn_qkvpacked_func",
    "stripe_flash_attn_func", "stripe_flash_attn_kvpacked_func", "stripe_flash_attn_qkvpacked_func",
    "ring_flash_attn_varlen_func", "ring_flash_attn_varlen_kvpacked_func", "ring_flash_attn_varlen_qkvpacked_func",
    "zigzag_ring_flash_attn_varlen_func", "zigzag_ring_flash_attn_varlen_kvpacked_func",
    "zigzag_ring_flash_attn_varlen_qkvpacked_func",
    "llama3_flash_attn_varlen_func", "llama3_flash_attn_varlen_kvpacked_func",
    "llama3_flash_attn_varlen_qkvpacked_func", "llama3_flash_attn_prepare_cu_seqlens",
    # beyond the reference (its README lists "zigzag llama3" as a TODO)
    "zigzag_llama3_flash_attn_varlen_func", "zigzag_llama3_flash_attn_varlen_kvpacked_func",
    "zigzag_llama3_flash_attn_varlen_qkvpacked_func",
]


# ----------------------------------------------------------------------------------------------
# autograd bridge
# ----------------------------------------------------------------------------------------------

class CPAttention(torch.autograd.Function):
    """One bridge for every scheme.  Inputs are token-major: q (T,Hq,D), k/v (T,Hkv,D)."""

    @staticmethod
    def forward(ctx, q, k, v, plan, scale, group, transport, heads_k_stride, deterministic):
        out, lse = engine.cp_forward(plan, q, k, v, scale, group, transport, heads_k_stride)
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.plan, ctx.scale, ctx.group = plan, scale, group
        ctx.transport, ctx.heads_k_stride, ctx.deterministic = transport, heads_k_stride, deterministic
        ctx.mark_non_differentiable(lse)
        return out, lse

    @staticmethod
    def backward(ctx, dout, _dlse):
        q, k, v, out, lse = ctx.saved_tensors
        dq, dk, dv = engine.cp_backward(ctx.plan, dout, q, k, v, out, lse, ctx.scale, ctx.group,
                                        ctx.transport, ctx.heads_k_stride, ctx.deterministic)
        return dq, dk, dv, None, None, None, None, None, None


KERNEL_HEAD_DIM = 128  # the sm_100a kernels are specialised for this head size


def _pad_head_dim(q) -> int:
    """Columns of zero padding that bring the head size to the kernels' 128.

    Zero columns change neither Q.K^T nor the first ``d`` columns of P.V, and their gradients are exactly zero,
    so a head size below 128 runs on the tcgen05 kernels (at the cost of the padded FLOPs) instead of dropping
    to the dense torch blocks.  ``RFA_B200_PAD_HEAD_DIM=0`` disables it, ``=force`` applies it on any device
    (used by the CPU tests)."""
    mode = os.environ.get("RFA_B200_PAD_HEAD_DIM", "1")
    d = q.shape[-1]
    if mode == "0" or d >= KERNEL_HEAD_DIM:
        return 0
    if mode == "force":
        return KERNEL_HEAD_DIM - d
    if q.dtype in (torch.bfloat16, torch.float16):
        from ..ops import cuda_ext

        if cuda_ext.available_for(q):
            return KERNEL_HEAD_DIM - d
    return 0


def _cp_apply(q, k, v, plan, scale, group, transport, heads_k_stride, deterministic, fp8=None):
    if fp8 is not None:
        # experimental fp8 forward: e4m3 q / k / v straight into the kernels (and over the wire), bf16 out.
        # Forward only - fp8 tensors carry no gradient.
        from ..ops import attn_cuda

        with attn_cuda.fp8_scales(*fp8), torch.no_grad():
            return engine.cp_forward(plan, q, k, v, scale, group, transport, heads_k_stride)
    pad = _pad_head_dim(q)
    if pad:
        d = q.shape[-1]
        q, k, v = (torch.nn.functional.pad(t, (0, pad)) for t in (q, k, v))
        out, lse = CPAttention.apply(q, k, v, plan, scale, group, transport, heads_k_stride, deterministic)
        return out[..., :d].contiguous(), lse
    return CPAttention.apply(q, k, v, plan, scale, group, transport, heads_k_stride, deterministic)


try:  # torch.compile: treat the op as an opaque eager call (plans, peer memory and launches are host code)
    _cp_apply = torch.compiler.disable(_cp_apply)
except AttributeError:  # pragma: no cover - very old torch
    pass


def _per_head(descale, x) -> Optional[torch.Tensor]:
    """(H,) fp32 vector if ``descale`` is a per-tensor or per-head scale of ``x`` (heads at dim -2), else None."""
    heads = x.shape[-2]
    if descale is None:
        return torch.ones(heads, dtype=torch.float32, device=x.device)
    if not isinstance(descale, torch.Tensor):
        return torch.full((heads,), float(descale), dtype=torch.float32, device=x.device)
    if descale.numel() == 1:
        return descale.reshape(1).to(device=x.device, dtype=torch.float32).expand(heads).contiguous()
    if descale.dim() == x.dim() and descale.shape[-2] == heads and descale.numel() == heads:
        return descale.reshape(heads).to(device=x.device, dtype=torch.float32).contiguous()
    return None


def _fp8_kernel_scales(q, k, v, dq, dk, dv, window_size):
    """Descales for the experimental fp8 forward kernel, or None when the call does not qualify: opt-in
    (``RFA_B200_FP8_KERNEL=1``), e4m3 q/k/v with head_dim 128 on a Blackwell GPU, per-tensor or per-head
    descales, no sliding window.  Finer block scales take the dequantise-to-bf16 path."""
    if os.environ.get("RFA_B200_FP8_KERNEL", "2") == "0":
        return None
    if not (q.dtype == k.dtype == v.dtype == torch.float8_e4m3fn) or q.shape[-1] != 128:
        return None
    if tuple(window_size) != (-1, -1):
        return None
    from ..ops import cuda_ext

    if not cuda_ext.available_for(q):
        return None
    sq, sk, sv = _per_head(dq, q), _per_head(dk, k), _per_head(dv, v)
    if sq is None or sk is None or sv is None:
        return None
    rep = q.shape[-2] // k.shape[-2]
    return (sq * sk.repeat_interleave(rep)).contiguous(), sv


def _maybe_dequant(q, k, v, descale, window_size=(-1, -1)):
    """fp8 extension (utils/fp8.py): ``descale`` is a tensor for a packed input or a (q, k, v) tuple.

    Returns (q, k, v, fp8_scales): either dequantised tensors and None, or - for calls that qualify for the
    experimental fp8 kernel - the untouched e4m3 tensors and their per-head descales."""
    from ..utils import fp8

    if not (fp8.is_fp8(q) or fp8.is_fp8(k) or fp8.is_fp8(v)):
        return q, k, v, None
    if descale is None:
        raise ValueError("fp8 q/k/v need descale=(q_descale, k_descale, v_descale)")
    dq, dk, dv = descale if isinstance(descale, (tuple, list)) else (descale, descale, descale)
    scales = _fp8_kernel_scales(q, k, v, dq, dk, dv, window_size)
    if scales is not None:
        return q, k, v, scales
    return fp8.dequantize(q, dq), fp8.dequantize(k, dk), fp8.dequantize(v, dv), None


def _split_descale(descale, pack_dim: int, n: int):
    """descale of a packed (kv / qkv) tensor -> per-tensor descales.  A tensor that carries the pack dimension
    (size n) is sliced like the data, anything else (None, a scalar tensor, per-head scales, ...) is shared."""
    if descale is None or isinstance(descale, (tuple, list)):
        return descale
    if isinstance(descale, torch.Tensor) and descale.dim() > pack_dim and descale.shape[pack_dim] == n:
        return tuple(descale.select(pack_dim, i) for i in range(n))
    return tuple(descale for _ in range(n))


def _check_common(q, dropout_p, window_size, alibi_slopes):
    if alibi_slopes is not None:
        raise NotImplementedError("alibi_slopes is not supported (same as the reference)")
    if dropout_p != 0.0:
        raise NotImplementedError("dropout_p > 0 is not supported by context-parallel attention")
    if len(tuple(window_size)) != 2:
        raise ValueError("window_size must be (left, right)")


def _scale(q, softmax_scale):
    return q.shape[-1] ** (-0.5) if softmax_scale is None else float(softmax_scale)


# The llama3 entry point only receives this rank's slice description; different global layouts can produce the
# SAME local description on some rank, so prepare() attaches the global cu_seqlens to the tensor objects it
# returns (attribute `_rfa_llama3`) and the entry point reads it back from the tensors the caller passes in.
# It lets a rank derive its peers' plans locally - the fused path must know which K/V rows each peer wants.


def cu_seqlens_to_host(cu: torch.Tensor) -> Tuple[int, ...]:
    """Host copy of a cu_seqlens tensor.  A device tensor is read back once; the result is remembered ON the
    tensor object together with its version counter (never keyed by address: the caching allocator hands the
    same address to the next batch's cu_seqlens)."""
    if not isinstance(cu, torch.Tensor):
        return tuple(int(x) for x in cu)
    if cu.device.type == "cpu":
        return tuple(int(x) for x in cu.tolist())
    cached = getattr(cu, "_rfa_host", None)
    if cached is not None and cached[0] == cu._version:
        return cached[1]
    vals = tuple(int(x) for x in cu.tolist())
    try:
        cu._rfa_host = (cu._version, vals)
    except AttributeError:  # pragma: no cover - exotic tensor subclasses
        pass
    return vals


def _window(window_size) -> Tuple[int, int]:
    left, right = (int(x) for x in window_size)
    return (-1 if left < 0 else left, -1 if right < 0 else right)


# Plans are cached per (scheme, rank, shapes, ...).  Each plan also knows how to produce its peers' plans
# (``plan.peer(r)``): the fused path derives from them which K/V rows every peer needs and which dK/dV rows every
# peer will send back, without any exchange at run time.

@functools.lru_cache(maxsize=512)
def _batch_plan(scheme, rank, world, batch, seqlen, causal, window=(-1, -1)):
    if scheme == "ring":
        plan = P.plan_ring(rank, world, batch, seqlen, causal, window)
    elif scheme == "zigzag":
        plan = P.plan_zigzag(rank, world, batch, seqlen, window)
    elif scheme == "stripe":
        plan = P.plan_stripe(rank, world, batch, seqlen, window)
    else:
        raise ValueError(scheme)
    plan.peer = lambda r: _batch_plan(scheme, r, world, batch, seqlen, causal, window)
    return plan


@functools.lru_cache(maxsize=512)
def _varlen_plan(scheme, rank, world, cu, causal, window=(-1, -1)):
    if scheme == "ring":
        plan = P.plan_ring_varlen(rank, world, cu, causal, window)
    elif scheme == "zigzag":
        plan = P.plan_zigzag_varlen(rank, world, cu, window)
    else:
        raise ValueError(scheme)
    plan.peer = lambda r: _varlen_plan(scheme, r, world, cu, causal, window)
    return plan


@functools.lru_cache(maxsize=512)
def _zigzag_llama3_plan(rank, world, global_cu, causal, window=(-1, -1)):
    plan = P.plan_zigzag_llama3(rank, world, global_cu, causal, window)
    plan.peer = lambda r: _zigzag_llama3_plan(r, world, global_cu, causal, window)
    return plan


@functools.lru_cache(maxsize=512)
def _llama3_plan(rank, world, tokens, cu_q, cu_k, k_start, causal, global_cu=None, window=(-1, -1)):
    # global_cu is part of the cache key on purpose: plans with equal local content but different global
    # layouts must not share their per-plan caches (peers' needs, push tables)
    plan = P.plan_llama3(rank, world, tokens, cu_q, cu_k, k_start, causal, window)
    if global_cu is not None:
        plan.peer = lambda r: _llama3_peer_plan(global_cu, causal, r, world, tokens, window)
    elif world > 1:
        # global layout unknown (cu tensors not produced by prepare()): the peers' needs cannot be derived
        # locally, so such calls use the torch.distributed all-gather transport around the same kernels
        plan.fused_ok = False
    return plan


def _run_batch(scheme, q, k, v, dropout_p, softmax_scale, causal, window_size, alibi_slopes,
               deterministic, return_attn_probs, group, descale=None):
    _check_common(q, dropout_p, window_size, alibi_slopes)
    q, k, v, fp8 = _maybe_dequant(q, k, v, descale, window_size)
    if scheme in ("zigzag", "stripe") and not causal:
        raise AssertionError(f"{scheme} attention only supports causal=True (as in the reference)")
    rank, world = group_info(group)
    b, s, hq, d = q.shape
    win = _window(window_size)
    plan = _batch_plan(scheme, rank, world, b, s, bool(causal), win)
    out, lse = _cp_apply(q.reshape(b * s, hq, d), k.reshape(b * s, k.shape[2], d),
                                 v.reshape(b * s, v.shape[2], d), plan, _scale(q, softmax_scale), group,
                                 "ring", 1, deterministic, fp8)
    out = out.view(b, s, hq, d)
    if not return_attn_probs:
        return out
    return out, lse.view(hq, b, s).permute(1, 0, 2).contiguous(), None


def _run_varlen(scheme, q, k, v, cu_seqlens, max_seqlen, dropout_p, softmax_scale, causal, window_size,
                alibi_slopes, deterministic, return_attn_probs, group, descale=None):
    _check_common(q, dropout_p, window_size, alibi_slopes)
    q, k, v, fp8 = _maybe_dequant(q, k, v, descale, window_size)
    if scheme == "zigzag" and not causal:
        raise AssertionError("zigzag attention only supports causal=True (as in the reference)")
    rank, world = group_info(group)
    cu_host = cu_seqlens_to_host(cu_seqlens)
    win = _window(window_size)
    plan = _varlen_plan(scheme, rank, world, cu_host, bool(causal), win)
    if plan.q_rows != q.shape[0]:
        raise ValueError(f"cu_seqlens[-1]={plan.q_rows} does not match the {q.shape[0]} local tokens")
    out, lse = _cp_apply(q, k, v, plan, _scale(q, softmax_scale), group, "ring", 1, deterministic, fp8)
    return (out, lse, None) if return_attn_probs else out


# ----------------------------------------------------------------------------------------------
# batch layout: ring / zigzag / stripe
# ----------------------------------------------------------------------------------------------

def _define_batch(scheme, prefix):
    def func(q, k, v, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1),
             alibi_slopes=None, deterministic=False, return_attn_probs=False, group=None, *, descale=None):
        return _run_batch(scheme, q, k, v, dropout_p, softmax_scale, causal, window_size, alibi_slopes,
                          deterministic, return_attn_probs, group, descale)

    def kvpacked_func(q, kv, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1),
                      alibi_slopes=None, deterministic=False, return_attn_probs=False, group=None, *,
                      descale=None):
        if descale is not None:  # (q_descale, kv_descale)
            dq, dkv = descale
            dk, dv = _split_descale(dkv, 2, 2)
            descale = (dq, dk, dv)
        return _run_batch(scheme, q, kv[:, :, 0], kv[:, :, 1], dropout_p, softmax_scale, causal,
                          window_size, alibi_slopes, deterministic, return_attn_probs, group, descale)

    def qkvpacked_func(qkv, dropout_p=0.0, softmax_scale=None, causal=False, window_size=(-1, -1),
                       alibi_slopes=None, deterministic=False, return_attn_probs=False, group=None, *,
                       descale=None):
        return _run_batch(scheme, qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], dropout_p, softmax_scale,
                          causal, window_size, alibi_slopes, deterministic, return_attn_probs, group,
                          _split_descale(descale, 2, 3))

    doc = ("{name}: q (B, S_local, Hq, D), k/v (B, S_local, Hkv, D) [kv (B,S_local,2,Hkv,D); "
           "qkv (B,S_local,3,H,D)] sharded with the '" + scheme + "' layout over `group`.  Returns out "
           "(and (out, softmax_lse (B,Hq,S_local), None) when return_attn_probs).")
    for f, suffix in ((func, "func"), (kvpacked_func, "kvpacked_func"), (qkvpacked_func, "qkvpacked_func")):
        f.__name__ = f.__qualname__ = f"{prefix}_flash_attn_{suffix}"
        f.__doc__ = doc.format(name=f.__name__)
    return func, kvpacked_func, qkvpacked_func


ring_flash_attn_func, ring_flash_attn_kvpacked_func, ring_flash_attn_qkvpacked_func = _define_batch("ring", "ring")
(zigzag_ring_flash_attn_func, zigzag_ring_flash_attn_kvpacked_func,
 zigzag_ring_flash_attn_qkvpacked_func) = _define_batch("zigzag", "zigzag_ring")
stripe_flash_attn_func, stripe_flash_attn_kvpacked_func, stripe_flash_attn_qkvpacked_func = _define_batch(
    "stripe", "stripe")


# ----------------------------------------------------------------------------------------------
# varlen layout: ring / zigzag
# ----------------------------------------------------------------------------------------------

def _define_varlen(scheme, prefix):
    def func(q, k, v, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None, causal=False,
             window_size=(-1, -1), alibi_slopes=None, deterministic=False, return_attn_probs=False,
             group=None, *, descale=None):
        return _run_varlen(scheme, q, k, v, cu_seqlens, max_seqlen, dropout_p, softmax_scale, causal,
                           window_size, alibi_slopes, deterministic, return_attn_probs, group, descale)

    def kvpacked_func(q, kv, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None, causal=False,
                      window_size=(-1, -1), alibi_slopes=None, deterministic=False,
                      return_attn_probs=False, group=None, *, descale=None):
        if descale is not None:
            dq, dkv = descale
            dk, dv = _split_descale(dkv, 1, 2)
            descale = (dq, dk, dv)
        return _run_varlen(scheme, q, kv[:, 0], kv[:, 1], cu_seqlens, max_seqlen, dropout_p, softmax_scale,
                           causal, window_size, alibi_slopes, deterministic, return_attn_probs, group, descale)

    def qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None, causal=False,
                       window_size=(-1, -1), alibi_slopes=None, deterministic=False,
                       return_attn_probs=False, group=None, *, descale=None):
        return _run_varlen(scheme, qkv[:, 0], qkv[:, 1], qkv[:, 2], cu_seqlens, max_seqlen, dropout_p,
                           softmax_scale, causal, window_size, alibi_slopes, deterministic,
                           return_attn_probs, group, _split_descale(descale, 1, 3))

    doc = ("{name}: packed q (T_local, Hq, D), k/v (T_local, Hkv, D) with ONE local cu_seqlens "
           "(global cu_seqlens // world_size); every document is sharded with the '" + scheme +
           "' layout.  Returns out (and (out, softmax_lse (Hq,T_local), None) when return_attn_probs).")
    for f, suffix in ((func, "func"), (kvpacked_func, "kvpacked_func"), (qkvpacked_func, "qkvpacked_func")):
        f.__name__ = f.__qualname__ = f"{prefix}_flash_attn_varlen_{suffix}"
        f.__doc__ = doc.format(name=f.__name__)
    return func, kvpacked_func, qkvpacked_func


(ring_flash_attn_varlen_func, ring_flash_attn_varlen_kvpacked_func,
 ring_flash_attn_varlen_qkvpacked_func) = _define_varlen("ring", "ring")
(zigzag_ring_flash_attn_varlen_func, zigzag_ring_flash_attn_varlen_kvpacked_func,
 zigzag_ring_flash_attn_varlen_qkvpacked_func) = _define_varlen("zigzag", "zigzag_ring")


# ----------------------------------------------------------------------------------------------
# llama3 (all-gather style context parallelism, varlen only)
# ----------------------------------------------------------------------------------------------

def llama3_flash_attn_prepare_cu_seqlens(cu_seqlens: torch.Tensor, causal: bool, rank: int, world_size: int):
    """Per-rank cu_seqlens for the llama3 layout (flat token stream split contiguously).

    Same outputs as the reference (/root/reference/ring_flash_attn/llama3_flash_attn_varlen.py:10-60):
    ``(cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k, local_k_slice)``.  The arithmetic is
    done once on the host (one device read instead of the reference's seven ``.item()`` syncs)."""
    cu = list(cu_seqlens_to_host(cu_seqlens))
    total = cu[-1]
    if total % world_size:
        raise AssertionError("total length must be divisible by world_size")
    L = total // world_size
    lo, hi = rank * L, (rank + 1) * L
    # documents overlapping [lo, hi): first doc whose end is > lo ... last doc whose start is < hi
    left = max(i for i in range(len(cu) - 1) if cu[i] <= lo)
    right = min(i for i in range(1, len(cu)) if cu[i] >= hi)
    window = cu[left:right + 1]
    cu_q = [min(max(x - lo, 0), L) for x in window]
    cu_q[0], cu_q[-1] = 0, L
    slice_left = cu[left]
    slice_right = hi if causal else cu[right]
    cu_k = list(window)
    cu_k[-1] = slice_right
    cu_k = [x - slice_left for x in cu_k]
    dev = cu_seqlens.device if isinstance(cu_seqlens, torch.Tensor) else None
    dt = cu_seqlens.dtype if isinstance(cu_seqlens, torch.Tensor) else torch.int32
    cu_q_t = torch.tensor(cu_q, dtype=dt, device=dev)
    cu_k_t = torch.tensor(cu_k, dtype=dt, device=dev)
    cu_q_t._rfa_host = (cu_q_t._version, tuple(cu_q))
    cu_k_t._rfa_host = (cu_k_t._version, tuple(cu_k))
    # remember the global layout ON the returned tensor objects (see _LLAMA3_GLOBAL note above)
    cu_q_t._rfa_llama3 = cu_k_t._rfa_llama3 = (tuple(cu), tuple(cu_q), tuple(cu_k), rank, world_size, bool(causal))
    max_q = max(b - a for a, b in zip(cu_q[:-1], cu_q[1:]))
    max_k = max(b - a for a, b in zip(cu_k[:-1], cu_k[1:]))
    return cu_q_t, cu_k_t, max_q, max_k, slice(slice_left, slice_right)


@functools.lru_cache(maxsize=512)
def _llama3_peer_plan(global_cu, causal, rank, world, tokens, window=(-1, -1)):
    cq, ck, _mq, _mk, ks = llama3_flash_attn_prepare_cu_seqlens(torch.tensor(global_cu, dtype=torch.int32), causal,
                                                               rank, world)
    return _llama3_plan(rank, world, tokens, tuple(cq.tolist()), tuple(ck.tolist()), int(ks.start), causal, global_cu,
                        window)


def llama3_flash_attn_varlen_func(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                                  heads_k_stride, local_k_slice, dropout_p=0.0, softmax_scale=None,
                                  causal=False, window_size=(-1, -1), alibi_slopes=None,
                                  deterministic=False, return_attn_probs=False, group=None, *, descale=None):
    """llama3-style CP: q/k/v (T_local, H*, D) are a contiguous slice of the flat token stream; the
    cu_seqlens / local_k_slice come from :func:`llama3_flash_attn_prepare_cu_seqlens`."""
    _check_common(q, dropout_p, window_size, alibi_slopes)
    q, k, v, fp8 = _maybe_dequant(q, k, v, descale, window_size)
    rank, world = group_info(group)
    k_start = local_k_slice.start or 0
    cu_q_host, cu_k_host = cu_seqlens_to_host(cu_seqlens_q), cu_seqlens_to_host(cu_seqlens_k)
    glob = None
    hit = getattr(cu_seqlens_q, "_rfa_llama3", None)
    if hit is not None and hit is getattr(cu_seqlens_k, "_rfa_llama3", None) and \
            hit[1:] == (cu_q_host, cu_k_host, rank, world, bool(causal)):
        glob = hit[0]
    plan = _llama3_plan(rank, world, q.shape[0], cu_q_host, cu_k_host, int(k_start), bool(causal), glob,
                        _window(window_size))
    out, lse = _cp_apply(q, k, v, plan, _scale(q, softmax_scale), group, "allgather",
                                 int(heads_k_stride), deterministic, fp8)
    return (out, lse, None) if return_attn_probs else out


def llama3_flash_attn_varlen_kvpacked_func(q, kv, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                                           heads_k_stride, local_k_slice, dropout_p=0.0, softmax_scale=None,
                                           causal=False, window_size=(-1, -1), alibi_slopes=None,
                                           deterministic=False, return_attn_probs=False, group=None, *,
                                           descale=None):
    """kv (T_local, 2, Hkv, D) variant of :func:`llama3_flash_attn_varlen_func`."""
    if descale is not None:
        dq, dkv = descale
        dk, dv = _split_descale(dkv, 1, 2)
        descale = (dq, dk, dv)
    return llama3_flash_attn_varlen_func(q, kv[:, 0], kv[:, 1], cu_seqlens_q, cu_seqlens_k, max_seqlen_q,
                                         max_seqlen_k, heads_k_stride, local_k_slice, dropout_p, softmax_scale,
                                         causal, window_size, alibi_slopes, deterministic, return_attn_probs,
                                         group, descale=descale)


def llama3_flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                                            heads_k_stride, local_k_slice, dropout_p=0.0, softmax_scale=None,
                                            causal=False, window_size=(-1, -1), alibi_slopes=None,
                                            deterministic=False, return_attn_probs=False, group=None, *,
                                            descale=None):
    """qkv (T_local, 3, H, D) variant of :func:`llama3_flash_attn_varlen_func`."""
    return llama3_flash_attn_varlen_func(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu_seqlens_q, cu_seqlens_k,
                                         max_seqlen_q, max_seqlen_k, heads_k_stride, local_k_slice, dropout_p,
                                         softmax_scale, causal, window_size, alibi_slopes, deterministic,
                                         return_attn_probs, group, descale=_split_descale(descale, 1, 3))


# ----------------------------------------------------------------------------------------------
# zigzag llama3: flat packed stream, rank r holds chunks r and 2W-1-r of 2W (beyond the reference)
# ----------------------------------------------------------------------------------------------

def zigzag_llama3_flash_attn_varlen_func(q, k, v, cu_seqlens, dropout_p=0.0, softmax_scale=None, causal=True,
                                         window_size=(-1, -1), alibi_slopes=None, deterministic=False,
                                         return_attn_probs=False, group=None, *, descale=None):
    """Load-balanced context parallelism for packed documents of ARBITRARY lengths.

    The flat token stream (all documents back to back, ``cu_seqlens`` = the GLOBAL cumulative lengths, the same on
    every rank) is cut into ``2 * world_size`` equal chunks and rank r holds chunks ``r`` and ``2W-1-r``
    (``parallel.layouts.shard_zigzag_llama3``).  Unlike the llama3 layout every rank does the same amount of causal
    work, and unlike the zigzag varlen layout no document length has to be divisible by ``2 * world_size``.
    q (T_local, Hq, D), k / v (T_local, Hkv, D); returns out (and (out, lse (Hq, T_local), None))."""
    _check_common(q, dropout_p, window_size, alibi_slopes)
    q, k, v, fp8 = _maybe_dequant(q, k, v, descale, window_size)
    rank, world = group_info(group)
    cu_host = cu_seqlens_to_host(cu_seqlens)
    if cu_host[-1] != q.shape[0] * world:
        raise ValueError(f"cu_seqlens[-1]={cu_host[-1]} must equal local tokens ({q.shape[0]}) x world size ({world}): "
                         "this entry point takes the GLOBAL cu_seqlens")
    plan = _zigzag_llama3_plan(rank, world, cu_host, bool(causal), _window(window_size))
    out, lse = _cp_apply(q, k, v, plan, _scale(q, softmax_scale), group, "ring", 1, deterministic, fp8)
    return (out, lse, None) if return_attn_probs else out


def zigzag_llama3_flash_attn_varlen_kvpacked_func(q, kv, cu_seqlens, dropout_p=0.0, softmax_scale=None, causal=True,
                                                  window_size=(-1, -1), alibi_slopes=None, deterministic=False,
                                                  return_attn_probs=False, group=None, *, descale=None):
    """kv (T_local, 2, Hkv, D) variant of :func:`zigzag_llama3_flash_attn_varlen_func`."""
    if descale is not None:
        dq, dkv = descale
        dk, dv = _split_descale(dkv, 1, 2)
        descale = (dq, dk, dv)
    return zigzag_llama3_flash_attn_varlen_func(q, kv[:, 0], kv[:, 1], cu_seqlens, dropout_p, softmax_scale, causal,
                                                window_size, alibi_slopes, deterministic, return_attn_probs, group,
                                                descale=descale)


def zigzag_llama3_flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens, dropout_p=0.0, softmax_scale=None, causal=True,
                                                   window_size=(-1, -1), alibi_slopes=None, deterministic=False,
                                                   return_attn_probs=False, group=None, *, descale=None):
    """qkv (T_local, 3, H, D) variant of :func:`zigzag_llama3_flash_attn_varlen_func`."""
    return zigzag_llama3_flash_attn_varlen_func(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu_seqlens, dropout_p, softmax_scale,
                                                causal, window_size, alibi_slopes, deterministic, return_attn_probs,
                                                group, descale=_split_descale(descale, 1, 3))

