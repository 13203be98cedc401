"""Context-parallel runtime: transports, layouts, plan engine, public API."""
from .comm import AllGatherComm, RingComm  # noqa: F401
