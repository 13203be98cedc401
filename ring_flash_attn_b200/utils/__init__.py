"""Reference-compatible utility surface (/root/reference/ring_flash_attn/utils.py:10 ``__all__``)."""
from ..ops.merge import update_out_and_lse  # noqa: F401
from ..ops.lse_layout import flatten_varlen_lse, unflatten_varlen_lse  # noqa: F401
from ..parallel.comm import AllGatherComm, RingComm  # noqa: F401

import inspect as _inspect
from functools import lru_cache as _lru_cache


@_lru_cache(maxsize=None)
def _defaults(func):
    target = getattr(func, "_init_fn", func)  # unwrap torch.library CustomOpDef like the reference does
    sig = _inspect.signature(target)
    return {k: v.default for k, v in sig.parameters.items() if v.default is not _inspect.Parameter.empty}


def get_default_args(func) -> dict:
    """Keyword defaults of ``func`` as a fresh dict (reference: utils.py:13-29, used there to stay
    compatible with several flash_attn signatures; this library has no such dependency but keeps the
    utility for drop-in compatibility)."""
    d = dict(_defaults(func))
    if "softcap" in d:
        d["softcap"] = 0.0
    return d


__all__ = ["update_out_and_lse", "RingComm", "AllGatherComm", "flatten_varlen_lse", "unflatten_varlen_lse",
           "get_default_args"]
