"""Peer-memory context (placeholder until the NVLink path lands): reports "unavailable" so multi-GPU calls
use the torch.distributed transport around the sm_100a block kernels."""


def peer_context(group, device):
    return None
