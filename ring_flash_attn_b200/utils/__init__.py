"""Reference-compatible utility surface (/root/reference/ring_flash_attn/utils.py:10 ``__all__``)."""
from ..ops.merge import update_out_and_lse  # noqa: F401
from ..ops.lse_layout import flatten_varlen_lse, unflatten_varlen_lse  # noqa: F401
from ..parallel.comm import AllGatherComm, RingComm  # noqa: F401

__all__ = ["update_out_and_lse", "RingComm", "AllGatherComm", "flatten_varlen_lse", "unflatten_varlen_lse"]
