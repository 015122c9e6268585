"""Host-side input pipeline helper: overlap the host->device copy of batch i+1 with the compute of batch i.

The reference leaves batch transfer to Lightning (synchronous ``batch.to(device)`` before every ``training_step``).  At ~50 ms
per 256-pair step the 259 MB of fp32 waveforms + images take ~5 ms over PCIe, so the copy runs on its own stream into one of
two device-side slots while the previous step computes.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, Iterator, Optional

import torch


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Restrict the calling thread to the CPUs of the NUMA node the GPU hangs off (sysfs; no libnuma needed) and return the node.

    Pinned staging buffers are then allocated (first touch) in memory local to the GPU's PCIe root: on a two-socket 8-GPU box a
    batch pinned on the far socket crosses the inter-socket link on every host->device copy, which was measured as an
    intermittent 1.2-1.5x slowdown of the end-to-end step (the 259 MB copy no longer hides behind a 38 ms step).  Returns None
    and changes nothing when the topology is not visible (containers without sysfs, single-node hosts)."""
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None


class DevicePrefetcher:
    """Iterates device-resident copies of pinned host batches (dicts of tensors), staged ahead of use on a copy stream.

    ``lookahead`` = how many batches the consumer holds BEYOND the one it is working on: 0 for a plain training loop (two device
    slots), 1 under ``TowerPipeline`` (which has already launched the towers of batch i + 1 when the step on batch i runs; three
    slots).  A slot is handed back to the copy stream when the consumer asks for the batch ``lookahead + 1`` positions later, by
    an event on the consumer's stream at that moment — everything it enqueued for the released batch precedes it."""

    _streams: Dict[str, "torch.cuda.Stream"] = {}  # one copy stream per device for the life of the process: the caching
    # allocator keeps freed blocks per stream, so a fresh stream per epoch would cudaMalloc (and synchronise) its slots again

    def __init__(self, batches: Optional[Iterable[Dict[str, torch.Tensor]]], device: torch.device, lookahead: int = 0):
        self.batches = batches
        self.device = device
        self.lookahead = int(lookahead)
        key = str(device)
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=device)
        self.copy_stream = DevicePrefetcher._streams[key]
        n = 2 + self.lookahead
        self.slots = [None] * n
        self.ready = [torch.cuda.Event() for _ in range(n)]
        self.released = [None] * n

    def iterate(self, batches: Iterable[Dict[str, torch.Tensor]]) -> "DevicePrefetcher":
        """Another pass (epoch) over new host batches through the same device slots."""
        self.batches = batches
        return self

    def _stage(self, slot: int, host: Dict[str, torch.Tensor]):
        with torch.cuda.stream(self.copy_stream):
            if self.released[slot] is not None:
                self.copy_stream.wait_event(self.released[slot])  # the step that read this slot has finished with it
            if self.slots[slot] is None or any(self.slots[slot][k].shape != v.shape for k, v in host.items()):
                self.slots[slot] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            for k, v in host.items():
                self.slots[slot][k].copy_(v, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        it = iter(self.batches)
        n = len(self.slots)
        i = 0
        try:
            self._stage(0, next(it))
        except StopIteration:
            return
        while True:
            slot = i % n
            torch.cuda.current_stream().wait_event(self.ready[slot])
            try:
                nxt = next(it)
                self._stage((i + 1) % n, nxt)  # last used by batch i + 1 - n, released when the consumer asked for batch i
                more = True
            except StopIteration:
                more = False
            yield self.slots[slot]
            # the consumer is back for the next batch: everything it enqueued for batch i - lookahead is on its stream by now
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            if i >= self.lookahead:
                self.released[(i - self.lookahead) % n] = ev
            if not more:
                return
            i += 1


class TowerPipeline:
    """Iterate device-resident batches with the FROZEN towers of batch i + 1 already running when the caller starts on batch i.

    A SpeechCLIP training step is two frozen towers (97 % of the FLOPs, compute-bound) followed by a tail of ~150 small, dependent
    kernels — weighted sum, branch forward, all-gather, InfoNCE, backward, gradient all-reduce, Adam — that leave most of the GPU
    idle (1.5 ms of a 6 ms step at 32 pairs per GPU).  The towers are frozen in every shipped configuration, so the towers of the
    NEXT batch do not depend on this batch's optimizer step: the pipeline launches them (``model.precompute_towers``) on the tower
    streams before it yields the current batch, and the tail of batch i overlaps the towers of batch i + 1.  Each yielded dict
    carries the handle as ``batch["_scb_towers"]``; ``KWClip_GeneralTransformer.forward`` picks it up.  Results are identical to
    the unpipelined loop (same kernels, same inputs); two slots of tower buffers alternate so that batch i + 1 never overwrites
    activations the backward pass of batch i still reads.

        for batch in TowerPipeline(model).iterate(DevicePrefetcher(loader, device, lookahead=1)):
            loss = model.training_step_end(model.training_step(batch))["loss"]; loss.backward(); optimizer.step()
    """

    def __init__(self, model, batches: Optional[Iterable[Dict[str, torch.Tensor]]] = None, head_priority: Optional[bool] = None):
        self.model = model
        self.batches = batches
        self._n = 0
        if head_priority is None:
            head_priority = os.environ.get("SCB_PIPELINE_PRIORITY", "1") != "0"
        self.head_priority = head_priority
        self._head_stream = None

    def head_stream(self, device) -> "torch.cuda.Stream":
        """The high-priority stream the loop body runs on while the pipeline is iterated (``head_priority``): the tail's small
        kernels are dependent on each other, so each one that queues behind a tower GEMM of the next batch adds its wait to the
        step; with priority they take the first SMs a tower kernel frees."""
        if self._head_stream is None:
            self._head_stream = torch.cuda.Stream(device=device, priority=-1)
        return self._head_stream

    def iterate(self, batches: Iterable[Dict[str, torch.Tensor]]) -> Iterator[Dict[str, torch.Tensor]]:
        self.batches = batches
        return iter(self)

    def _launch(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        out = dict(batch)
        out["_scb_towers"] = self.model.precompute_towers(batch, slot=1 + (self._n & 1))
        self._n += 1
        return out

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        if getattr(self.batches, "lookahead", 1) < 1:
            raise ValueError("TowerPipeline holds one batch beyond the current one: build the DevicePrefetcher with lookahead=1")
        if not self.head_priority:
            yield from self._iterate()
            return
        # The loop body (the caller's step) runs with the high-priority head stream current: the generator is suspended inside
        # the ``with`` block, and the stream context is per thread.  On exit (exhaustion, break or exception) the caller's own
        # stream is current again and ordered after everything the loop enqueued.
        dev = next(self.model.parameters()).device
        # the parameters' AccumulateGrad nodes live on the stream of their first backward; autograd orders the streams itself and
        # warns about the mismatch on every step otherwise
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        outer = torch.cuda.current_stream(dev)
        hs = self.head_stream(dev)
        hs.wait_stream(outer)
        try:
            with torch.cuda.stream(hs):
                yield from self._iterate()
        finally:
            outer.wait_stream(hs)

    def _iterate(self) -> Iterator[Dict[str, torch.Tensor]]:
        it = iter(self.batches)
        try:
            nxt = self._launch(next(it))
        except StopIteration:
            return
        for following in it:
            cur, nxt = nxt, self._launch(following)   # towers of the following batch are enqueued BEFORE the caller's step on `cur`
            yield cur
        yield nxt
