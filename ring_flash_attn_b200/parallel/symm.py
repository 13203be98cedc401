"""Peer-memory runtime and the fused multi-GPU forward / backward drivers.

One :class:`PeerContext` exists per (process group, device).  It owns three CUDA-IPC buffers that every
rank of the group maps (``csrc/peer_mem.cpp``):

* the *signal pad*   - uint32 epochs written by peers (layout in ``csrc/attn_common.h``: kPad*),
* the *K/V staging*  - ``[parity][K|V][slot = source rank][rows][hkv][128]``: the rows of every other rank's
                       shard that this rank's plan needs, stored there by the *source's* attention kernel,
* the *dK/dV inbox*  - ``[slot = source rank][dK|dV][rows][hkv][128]`` in the model dtype: partial gradients for
                       this rank's shard, stored there by the peers' backward kernels (summed in fp32 by the owner).

A forward is ONE launch of ``attn_fwd_kernel``: its first CTAs push this rank's K/V rows to the peers that
need them (ring order; TMA bulk copies HBM -> shared memory -> peer HBM), the rest run the math and only wait on
a source's "landed" flag when they reach that source's segments.  A backward is ``delta`` + ONE launch of ``attn_bwd_kernel``
(same push CTAs; dK/dV tiles are stored straight into the owners' inboxes) + the owner-side reduction.
No NCCL call is on these paths; NCCL is used once, to exchange the IPC handles.

Replaces: RingComm.send_recv_kv / batch_isend_irecv, all_gather_into_tensor and reduce_scatter_tensor in
/root/reference/ring_flash_attn/utils.py:98-168 and llama3_flash_attn_varlen.py:97-115,292-293.
"""
from __future__ import annotations

import os
import socket
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from ..ops import attn_cuda, cuda_ext
from ..ops.plan import CPPlan
from .comm import group_info

PAD_WORDS = 1024
MASK32 = 0xFFFFFFFF
_CONTEXTS: Dict[Tuple[str, int], Optional["PeerContext"]] = {}


class PeerBuffer:
    """A cudaMalloc'd buffer mapped by every rank of the group."""

    def __init__(self, nbytes: int, device: torch.device, group):
        C = cuda_ext.load()
        self.nbytes = nbytes
        self.device = device
        self.local_ptr, handle = C.peer_alloc(int(nbytes), device.index)
        rank, world = group_info(group)
        handles: List[Optional[bytes]] = [None] * world
        dist.all_gather_object(handles, (device.index, handle), group=group)
        self.ptrs: List[int] = []
        for r, (_dev, h) in enumerate(handles):
            self.ptrs.append(self.local_ptr if r == rank else C.peer_open(h, device.index))
        self.rank = rank

    def tensor(self, byte_offset: int, shape, dtype) -> torch.Tensor:
        return cuda_ext.load().tensor_from_ptr(self.local_ptr + byte_offset, list(shape), dtype, self.device.index)

    def close(self):
        C = cuda_ext.load()
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                C.peer_close(p)
        C.peer_free(self.local_ptr)
        self.ptrs = []


class PeerContext:
    def __init__(self, group, device: torch.device):
        self.group = group
        self.device = device
        self.rank, self.world = group_info(group)
        self.pad = PeerBuffer(PAD_WORDS * 4, device, group)
        self.pad_tensor = self.pad.tensor(0, (PAD_WORDS,), torch.int32)
        self.counters = torch.zeros(64, dtype=torch.int32, device=device)
        self.epoch = 0
        self.last_bwd_epoch = 0
        self.sent_cum = [0] * self.world
        self.dkv_cum = [0] * self.world
        self.done_cum = 0
        self.ticket_cum = 0
        self.stage: Optional[PeerBuffer] = None
        self.stage_half = 0
        self.inbox: Optional[PeerBuffer] = None
        self.n_push_ctas = int(os.environ.get("RFA_B200_PUSH_CTAS", "24"))
        # copy-engine K/V transport (csrc/peer_mem.cpp:kv_push_dma): a side stream + a ring of epoch words
        self.side: Optional[torch.cuda.Stream] = None
        self.flag_host = self.flag_dev = None

    # -- buffers ------------------------------------------------------------------------------------
    def _quiesce(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    def ensure_stage(self, rows: int, hkv: int, dtype, head_dim: int = 128) -> None:
        """Grow-only staging capacity (two call-parity halves).  The per-call layout is derived from the
        current shapes, so batches of different length reuse the same mapping; a (collective) reallocation
        only happens when a call needs more bytes than any call before it."""
        esize = torch.empty((), dtype=dtype).element_size()
        need_half = 2 * self.world * rows * hkv * head_dim * esize  # [K|V][slot][rows] for one parity
        if self.stage is not None and need_half <= self.stage_half:
            return
        self._quiesce()
        if self.stage is not None:
            self.stage.close()
            # growing is collective (barrier + cudaMalloc + IPC exchange): over-allocate by a quarter when it has
            # to happen again, so that a stream of slowly growing packed batches does not stall every few steps
            need_half = max(need_half, self.stage_half + self.stage_half // 4)
        self.stage_half = (need_half + 4095) // 4096 * 4096
        self.stage = PeerBuffer(2 * self.stage_half, self.device, self.group)
        self._quiesce()

    def ensure_inbox(self, rows: int, hkv: int, dtype, head_dim: int = 128) -> None:
        esize = torch.empty((), dtype=dtype).element_size()
        need = self.world * 2 * rows * hkv * head_dim * esize
        self.inbox_kv_stride = rows * hkv * head_dim  # elements
        self.inbox_slot_stride = 2 * self.inbox_kv_stride
        if self.inbox is not None and need <= self.inbox.nbytes:
            return
        self._quiesce()
        if self.inbox is not None:
            self.inbox.close()
            need = max(need, self.inbox.nbytes + self.inbox.nbytes // 4)  # see ensure_stage
        self.inbox = PeerBuffer(need, self.device, self.group)
        self._quiesce()

    # -- host-side protocol state --------------------------------------------------------------------
    # The cumulative targets below are advanced BEFORE a launch (the launch needs them); if the launch then fails
    # on the host (a TORCH_CHECK, an invalid configuration) the device counters never moved, and the next call of
    # every rank would wait for epochs that are never reached.  Callers snapshot / restore around the launch.
    _STATE = ("epoch", "last_bwd_epoch", "done_cum", "ticket_cum")

    def snapshot(self):
        return ({n: getattr(self, n) for n in self._STATE}, list(self.sent_cum), list(self.dkv_cum))

    def restore(self, snap) -> None:
        scalars, sent, dkv = snap
        for n, v in scalars.items():
            setattr(self, n, v)
        self.sent_cum, self.dkv_cum = list(sent), list(dkv)

    # -- per-call context object --------------------------------------------------------------------
    def kv_transport(self, plan: CPPlan) -> str:
        """``push``: communication CTAs inside the attention launch (TMA bulk copies over NVLink).  ``dma``: the copy
        engines move the rows on a side stream while the launch (without push CTAs) waits on the same flags.
        ``RFA_B200_KV_TRANSPORT`` selects; llama3 plans (needs only known on the device) always push."""
        mode = os.environ.get("RFA_B200_KV_TRANSPORT", "push")
        if mode != "dma" or is_dynamic(plan) or not cuda_ext.load().dma_transport_available():
            return "push"
        return "dma"

    def push_dma(self, plan: CPPlan, k: torch.Tensor, v: torch.Tensor, fc) -> torch.cuda.Event:
        """Queue this call's K/V rows for every peer on the side stream; returns the event that marks them sent."""
        C = cuda_ext.load()
        if self.side is None:
            self.side = torch.cuda.Stream(device=self.device)
            self.flag_host = torch.zeros(1024, dtype=torch.int32).pin_memory()
            self.flag_dev = torch.zeros(1024, dtype=torch.int32, device=self.device)
        tasks = push_tasks_host(plan, self, int(fc.row_bytes))
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)  # K / V are produced on the main stream
        self.side.wait_event(ready)
        C.kv_push_dma(k, v, tasks, list(self.stage.ptrs), list(self.pad.ptrs), int(self.pad.local_ptr),
                      int(fc.parity_off), int(fc.row_bytes), self.rank, int(fc.epoch), self.flag_host, self.flag_dev,
                      self.side.cuda_stream)
        sent = torch.cuda.Event()
        sent.record(self.side)
        return sent

    def fused_ctx(self, plan: CPPlan, k: torch.Tensor, n_compute_ctas: int, dyn_needs: Optional[torch.Tensor] = None):
        C = cuda_ext.load()
        rows, hkv, d = plan.kv_rows, k.shape[1], k.shape[2]
        esize = k.element_size()
        row_bytes = hkv * d * esize
        self.epoch += 1
        parity = self.epoch & 1
        fc = C.FusedCtx()
        half = parity * self.stage_half
        region = self.world * rows * row_bytes
        fc.k_stage = self.stage.tensor(half, (self.world * rows, hkv, d), k.dtype)
        fc.v_stage = self.stage.tensor(half + region, (self.world * rows, hkv, d), k.dtype)
        fc.my_pad = self.pad_tensor
        fc.rows_cap, fc.region_bytes = rows, region
        if dyn_needs is None and self.kv_transport(plan) == "dma":
            # the copy engines push (push_dma); the launch has no communication CTAs
            tasks, per_dst, n_tasks = torch.zeros((0, 4), dtype=torch.int64, device=k.device), [0] * self.world, 0
        elif dyn_needs is None:
            tasks, per_dst = push_tasks(plan, self, row_bytes, k.device)
            n_tasks = int(tasks.shape[0])
        else:
            # the peers' needs only exist on the device (needs_gathered): the push CTAs cut their tasks out of the
            # table themselves; every (destination, K|V, range, chunk) counts, copied rows or not
            tasks = torch.zeros((0, 4), dtype=torch.int64, device=k.device)
            chunk = push_chunk_rows(row_bytes)
            chunks = -(-rows // chunk)
            fc.dyn_needs, fc.dyn_chunk_rows, fc.dyn_chunks = dyn_needs, chunk, chunks
            per_dst = [0 if d == self.rank else 2 * NEED_RANGES * chunks for d in range(self.world)]
            n_tasks = sum(per_dst)
        fc.push_tasks = tasks
        fc.counters = self.counters
        fc.stage_ptrs = list(self.stage.ptrs)
        fc.pad_ptrs = list(self.pad.ptrs)
        for d in range(self.world):
            self.sent_cum[d] = (self.sent_cum[d] + per_dst[d]) & MASK32
        fc.sent_targets = list(self.sent_cum)
        fc.n_push_ctas = min(self.n_push_ctas, n_tasks) if n_tasks else 0
        fc.row_bytes = row_bytes
        fc.my_rank = self.rank
        fc.world = self.world
        fc.epoch = self.epoch
        fc.parity_off = half
        self.done_cum = (self.done_cum + n_compute_ctas) & MASK32
        fc.done_target = self.done_cum
        return fc


def destroy_peer_contexts() -> None:
    """Unmap and free every peer buffer (collective: call on all ranks, e.g. before destroy_process_group)."""
    for ctx in list(_CONTEXTS.values()):
        if ctx is None:
            continue
        ctx._quiesce()
        for buf in (ctx.stage, ctx.inbox, ctx.pad):
            if buf is not None:
                buf.close()
        ctx.stage = ctx.inbox = None
    _CONTEXTS.clear()


def peer_context(group, device: torch.device) -> Optional[PeerContext]:
    """The context for (group, device), created collectively on first use; None if P2P is unavailable."""
    # keyed by the group's c10d NAME (unique for the life of the process; id(group) can be reused by a new group
    # after the old one is destroyed, which would hand it stale IPC mappings and epochs)
    key = (str(group.group_name) if group is not None else "", device.index)
    if key in _CONTEXTS:
        return _CONTEXTS[key]
    rank, world = group_info(group)
    C = cuda_ext.load()
    info: List[Optional[tuple]] = [None] * world
    dist.all_gather_object(info, (socket.gethostname(), device.index), group=group)
    ok = len({h for h, _ in info}) == 1 and len({d for _, d in info}) == world and world <= 16
    if ok:
        ok = all(d == device.index or C.can_access_peer(device.index, d) for _, d in info)
    flags: List[Optional[bool]] = [None] * world
    dist.all_gather_object(flags, bool(ok), group=group)
    ctx = PeerContext(group, device) if all(flags) else None
    _CONTEXTS[key] = ctx
    return ctx


# ----------------------------------------------------------------------------------------------
# who needs what
# ----------------------------------------------------------------------------------------------
NEED_RANGES = 4  # csrc/attn_common.h: kNeedRanges


def needs_matrix(plan: CPPlan, group) -> List[List[List[Tuple[int, int]]]]:
    """needs[dst][src] = list of [lo, hi) row ranges of src's shard that dst's plan reads, derived locally from the
    plan family (position-based schemes: every rank can build every peer's plan)."""
    cached = getattr(plan, "_needs", None)
    if cached is not None:
        return cached
    world = plan.world
    fam = [plan if r == plan.rank else plan.peer(r) for r in range(world)]
    needs = [[_ranges(fam[d], s) for s in range(world)] for d in range(world)]
    plan._needs = needs
    return needs


def is_dynamic(plan: CPPlan) -> bool:
    """Plans that cannot derive their peers' plans (llama3: a rank only sees its own slice of the global
    cu_seqlens).  Which rows every peer needs is then exchanged on the DEVICE right before each launch."""
    return plan.world > 1 and getattr(plan, "peer", None) is None


def dynamic_ok(plan: CPPlan) -> bool:
    """The device-side needs table holds NEED_RANGES row ranges per source."""
    return all(len(_ranges(plan, s)) <= NEED_RANGES for s in range(plan.world))


def local_needs_table(plan: CPPlan) -> torch.Tensor:
    """int32 (world src, NEED_RANGES, 2): the [lo, hi) row ranges of every source's shard this rank's plan reads."""
    table = torch.zeros((plan.world, NEED_RANGES, 2), dtype=torch.int32)
    for s in range(plan.world):
        for j, (lo, hi) in enumerate(_ranges(plan, s)):
            table[s, j, 0], table[s, j, 1] = lo, hi
    return table


def dynamic_push_task(needs_all, me: int, world: int, ti: int, chunk_rows: int, chunks: int, rows_cap: int,
                      row_bytes: int, region: int):
    """Host mirror of ``push_task_at`` (csrc/comm_device.cuh) for tests: (src_row, dst_off, rows, dst, which)."""
    per_dst = 2 * NEED_RANGES * chunks
    step = ti // per_dst + 1
    rem = ti - (step - 1) * per_dst
    dst = (me + step) % world
    which = rem // (NEED_RANGES * chunks)
    rem -= which * NEED_RANGES * chunks
    rng, c = rem // chunks, rem % chunks
    lo, hi = int(needs_all[dst][me][rng][0]), int(needs_all[dst][me][rng][1])
    r0 = lo + c * chunk_rows
    rows = min(chunk_rows, hi - r0) if r0 < hi else 0
    return r0, which * region + (me * rows_cap + r0) * row_bytes, rows, dst, which


def dynamic_row_mask(needs_all, me: int, world: int, row: int) -> int:
    """Host mirror of the reduce kernel's ``dyn_mask``: bit s set iff rank s returns dK/dV for ``row`` of my shard."""
    m = 0
    for s in range(world):
        if any(int(lo) <= row < int(hi) for lo, hi in needs_all[s][me]):
            m |= 1 << s
    return m


def needs_gathered(plan: CPPlan, ctx: "PeerContext", device) -> torch.Tensor:
    """int32 (world dst, world src, NEED_RANGES, 2) on the device: one small all-gather on the current stream per
    launch, no host synchronisation (the result is only ever read by the kernels that follow in stream order).

    It runs on EVERY call on purpose: two global layouts can give one rank the same local slice description (so
    that rank could reuse a cached answer) while its peers see different ones - a "gather only when my plan is
    new" rule would make the ranks disagree about whether a collective happens."""
    cache = attn_cuda._cache(plan)
    key = ("needs_local", device.index)
    if key not in cache:
        cache[key] = local_needs_table(plan).to(device)
    out = torch.empty((plan.world, plan.world, NEED_RANGES, 2), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(out, cache[key], group=ctx.group)
    return out


def push_chunk_rows(row_bytes: int) -> int:
    return max(16, (1 << 19) // row_bytes)  # ~512 KB per push task


def _ranges(plan: CPPlan, src: int) -> List[Tuple[int, int]]:
    iv = sorted((s.kv_row0, s.kv_row0 + s.kv_len) for s in plan.segments if s.src == src)
    out: List[Tuple[int, int]] = []
    for lo, hi in iv:
        if out and lo <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], hi))
        else:
            out.append((lo, hi))
    return out


def push_tasks_host(plan: CPPlan, ctx: PeerContext, row_bytes: int) -> torch.Tensor:
    """The push table as a CPU tensor with ONE task per (destination, range, K|V) - the copy engines take whole
    ranges, chunking only matters for spreading work over push CTAs."""
    key = ("push_host", row_bytes)
    cache = attn_cuda._cache(plan)
    if key not in cache:
        needs = needs_matrix(plan, ctx.group)
        me, world, rows = plan.rank, plan.world, plan.kv_rows
        region = world * rows * row_bytes
        table = []
        for step in range(1, world):
            dst = (me + step) % world
            for lo, hi in needs[dst][me]:
                for which in (0, 1):
                    table.append([lo, which * region + (me * rows + lo) * row_bytes, (hi - lo) | (dst << 32), which])
        cache[key] = torch.tensor(table, dtype=torch.int64) if table else torch.zeros((0, 4), dtype=torch.int64)
    return cache[key]


def push_tasks(plan: CPPlan, ctx: PeerContext, row_bytes: int, device):
    """(int64 task table on the device, tasks per destination).  Cached on the plan."""
    key = ("push", row_bytes, device.index)
    cache = attn_cuda._cache(plan)
    if key in cache:
        return cache[key]
    needs = needs_matrix(plan, ctx.group)
    me, world, rows = plan.rank, plan.world, plan.kv_rows
    chunk = push_chunk_rows(row_bytes)
    region = world * rows * row_bytes
    table, per_dst = [], [0] * world
    for step in range(1, world):
        dst = (me + step) % world  # ring order: the next neighbour consumes our shard first
        for lo, hi in needs[dst][me]:
            for r0 in range(lo, hi, chunk):
                n = min(chunk, hi - r0)
                for which in (0, 1):
                    dst_off = which * region + (me * rows + r0) * row_bytes
                    table.append([r0, dst_off, n | (dst << 32), which])
                    per_dst[dst] += 1
    t = torch.tensor(table, dtype=torch.int64).to(device) if table else torch.zeros((0, 4), dtype=torch.int64,
                                                                                     device=device)
    cache[key] = (t, per_dst)
    return cache[key]


# ----------------------------------------------------------------------------------------------
# fused forward / backward
# ----------------------------------------------------------------------------------------------

def _kv_ok(t: torch.Tensor) -> torch.Tensor:
    t = attn_cuda._rows3(t)
    return t if t.stride(1) == t.shape[2] else t.contiguous()


def fused_forward(plan: CPPlan, q, k, v, scale, group):
    ctx = peer_context(group, q.device)
    C = cuda_ext.load()
    k, v = _kv_ok(k), _kv_ok(v)
    rows, hq = plan.kv_rows, q.shape[1]
    ctx.ensure_stage(rows, k.shape[1], k.dtype, k.shape[2])
    offsets = {s: (0 if s == plan.rank else s * rows) for s in range(plan.world)}
    flags = {s: s for s in range(plan.world) if s != plan.rank}
    window = attn_cuda.has_window(plan.segments)
    if window:
        cache = attn_cuda._cache(plan)
        key = ("fwd_fused_window", q.device.index)
        if key not in cache:
            it_h, seg_h, lo_h, cov = attn_cuda.fwd_tables_window_host(plan, plan.segments, offsets, flags)
            cache[key] = (attn_cuda._to_dev(it_h, 8, q.device),
                          attn_cuda._to_dev(seg_h if seg_h else [[0, 0, 0, -1]], 4, q.device),
                          torch.tensor(lo_h if lo_h else [attn_cuda.LO_NONE], dtype=torch.int32).to(q.device), cov)
        items, segs, seg_lo, covered = cache[key]
    else:
        items, segs, covered = attn_cuda.fwd_tables(plan, plan.segments, offsets, q.device, ("fused",), flags)
    dyn = needs_gathered(plan, ctx, q.device) if is_dynamic(plan) else None
    tq = q.shape[0]
    out = (torch.empty if covered else torch.zeros)((tq, hq, q.shape[2]), dtype=attn_cuda.out_dtype(q),
                                                    device=q.device)
    lse = torch.empty((hq, tq), dtype=torch.float32, device=q.device)
    if not covered:
        lse.fill_(float("-inf"))
    snap = ctx.snapshot()
    sent = None
    try:
        fc = ctx.fused_ctx(plan, k, int(items.shape[0]) * hq, dyn)
        if dyn is None and ctx.kv_transport(plan) == "dma":
            sent = ctx.push_dma(plan, k, v, fc)
        if attn_cuda.is_fp8_kernel_input(q, k):
            sc = attn_cuda.current_fp8_scales()  # tables of every rank's shard (engine.cp_forward gathered them)
            C.attn_fwd_fused_fp8(attn_cuda._rows3(q), k, v, items, segs, *sc.args(), out, lse, tq, float(scale), fc)
        elif window:
            C.attn_fwd_fused_window(attn_cuda._rows3(q), k, v, items, segs, seg_lo, out, lse, tq, float(scale), fc)
        else:
            C.attn_fwd_fused(attn_cuda._rows3(q), k, v, items, segs, out, lse, tq, float(scale), fc)
    except Exception:
        ctx.restore(snap)  # nothing reached the device: keep host and device counters in step
        raise
    if sent is not None:
        torch.cuda.current_stream(q.device).wait_event(sent)  # k / v stay alive until the copy engines are done
    cuda_ext.note_launch()
    return out, lse


def reduce_tasks(plan: CPPlan, ctx: PeerContext, device):
    key = ("reduce", device.index)
    cache = attn_cuda._cache(plan)
    if key in cache:
        return cache[key]
    if is_dynamic(plan):
        # who contributes which rows is only known on the device: fixed row blocks, masks computed by the kernel
        rows, blk = plan.kv_rows, 512
        cache[key] = torch.tensor([[r0, min(blk, rows - r0), 0, 0] for r0 in range(0, rows, blk)],
                                  dtype=torch.int32).to(device)
        return cache[key]
    needs = needs_matrix(plan, ctx.group)
    me, world, rows = plan.rank, plan.world, plan.kv_rows
    cuts = {0, rows}
    for s in range(world):
        for lo, hi in needs[s][me]:
            cuts.update((lo, hi))
    cuts = sorted(cuts)
    table = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        mask = 0
        for s in range(world):
            if any(a <= lo and hi <= b for a, b in needs[s][me]):
                mask |= 1 << s
        table.append([lo, hi - lo, mask, 0])
    t = torch.tensor(table, dtype=torch.int32).to(device)
    cache[key] = t
    return t


def fused_backward(plan: CPPlan, dout, q, k, v, out, lse, scale, group, deterministic=False):
    ctx = peer_context(group, q.device)
    C = cuda_ext.load()
    k, v = _kv_ok(k), _kv_ok(v)
    rows, hkv = plan.kv_rows, k.shape[1]
    ctx.ensure_stage(rows, hkv, k.dtype, k.shape[2])
    ctx.ensure_inbox(rows, hkv, k.dtype, k.shape[2])
    offsets = {s: (0 if s == plan.rank else s * rows) for s in range(plan.world)}
    flags = {s: s for s in range(plan.world) if s != plan.rank}
    window = attn_cuda.has_window(plan.segments)
    if window:
        cache = attn_cuda._cache(plan)
        key = ("bwd_fused_window", q.device.index)
        if key not in cache:
            it_h, qs_h, per = attn_cuda.bwd_tables_window_host(plan, plan.segments, offsets, flags, fused=True)
            cache[key] = (attn_cuda._to_dev(it_h, 8, q.device),
                          attn_cuda._to_dev(qs_h if qs_h else [[0, 0, 0, 0]], 4, q.device), per)
        items, qsegs, per_owner = cache[key]
    else:
        items, qsegs, per_owner = attn_cuda.bwd_tables_fused(plan, offsets, q.device, flags)
    delta = attn_cuda.compute_delta(out, dout)
    dq = attn_cuda.dq_workspace.acquire(q)  # zeroed fp32 accumulator, re-zeroed by dq_finalize
    dyn = needs_gathered(plan, ctx, q.device) if is_dynamic(plan) else None
    snap = ctx.snapshot()
    sent = None
    try:
        fc = ctx.fused_ctx(plan, k, int(items.shape[0]) * hkv, dyn)
        if dyn is None and ctx.kv_transport(plan) == "dma":
            sent = ctx.push_dma(plan, k, v, fc)
        me, esize = plan.rank, k.element_size()
        fc.dk_ptrs = [p + me * ctx.inbox_slot_stride * esize for p in ctx.inbox.ptrs]
        fc.dv_ptrs = [p + (me * ctx.inbox_slot_stride + ctx.inbox_kv_stride) * esize for p in ctx.inbox.ptrs]
        for o in range(plan.world):
            ctx.dkv_cum[o] = (ctx.dkv_cum[o] + per_owner[o] * hkv) & MASK32
        fc.dkv_targets = list(ctx.dkv_cum)
        fc.dkv_wait_epoch = ctx.last_bwd_epoch
        launch = C.attn_bwd_fused_window if window else C.attn_bwd_fused
        launch(attn_cuda._rows3(q), attn_cuda._rows3(dout), k, v, dq, items, qsegs, lse.contiguous(), delta,
               q.shape[0], float(scale), fc)
    except Exception:
        ctx.restore(snap)  # nothing reached the device: keep host and device counters in step
        raise
    if sent is not None:
        torch.cuda.current_stream(q.device).wait_event(sent)
    cuda_ext.note_launch()
    # owner-side reduction of the inbox (waits for the peers' "gradients landed" epochs on the device)
    tasks = reduce_tasks(plan, ctx, q.device)
    dk = torch.empty((rows, hkv, k.shape[2]), dtype=k.dtype, device=q.device)
    dv = torch.empty((rows, hkv, k.shape[2]), dtype=k.dtype, device=q.device)
    inbox = ctx.inbox.tensor(0, (ctx.world * ctx.inbox_slot_stride,), k.dtype)
    ctx.ticket_cum = (ctx.ticket_cum + 128 * int(tasks.shape[0])) & MASK32  # kReduceBlocksPerTask
    C.reduce_dkv(inbox, ctx.inbox_slot_stride, ctx.inbox_kv_stride, dk, dv, tasks, fc, ctx.ticket_cum)
    cuda_ext.note_launch()
    ctx.last_bwd_epoch = ctx.epoch
    return attn_cuda.dq_workspace.finalize(dq, q), dk, dv
