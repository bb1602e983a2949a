"""Clip sharding across the GPUs of one box.

Denoising is independent per clip (eval-mode norms use running statistics, diffusion_trainer.py:862), so the loop
runs with no collective: rank r owns a contiguous block of ceil(N / world) clips and the predicted maps are gathered
once after the loop (344 KB per clip as fp32, 86 KB as uint8).  The reference shards its test loaders with
``DistributedSampler`` (datasets/prepare_data.py:87-103), which deals clips out round-robin (rank r gets r, r + world,
...) after padding to a multiple of world; contiguous blocks are used here instead because they make the gathered
tensor come back in clip order with one equal-size all_gather -- which clip a rank processes does not change its map
(tests/test_sharding_gloo.py, the batch-invariance tests).  torch.distributed is the plumbing (NCCL over NVLink on the
GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_clips, rank, world):
    """[start, stop) of the clips owned by ``rank``; blocks of ceil(n/world), the tail ranks may be short/empty."""
    per = (n_clips + world - 1) // world
    start = min(n_clips, rank * per)
    return start, min(n_clips, start + per)


def micro_batches(start, stop, size):
    """[(lo, hi)] covering [start, stop) in micro-batches of at most ``size`` clips (BASELINE config 4: 32 per GPU)."""
    return [(lo, min(stop, lo + size)) for lo in range(start, stop, size)]


class MapGatherer:
    """Equal-size all_gather of the per-rank map blocks into preallocated buffers (nothing is allocated per call).

    ``local_maps`` is this rank's [n_local, 1, H, W] block (n_local may be smaller than ceil(n/world) on the last ranks:
    the padded tail keeps whatever it held and is trimmed from the result).

    Two ways to use it: ``g(local_maps)`` gathers in line; ``g.start(local_maps, slot)`` / ``g.finish(slot)`` put the
    collective on a side stream so that it overlaps whatever the caller enqueues next (the loop of the next batch) --
    two slots, i.e. the caller may have the gathers of two consecutive batches in flight and must not overwrite
    ``local_maps`` of a slot before ``finish(slot)`` (or the next ``start`` on that slot) has been enqueued."""

    def __init__(self, n_clips, map_shape, dtype, device, group=None):
        self.group = group
        self.n_clips = int(n_clips)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.per = (self.n_clips + self.world - 1) // self.world
        self.pad = torch.zeros((self.per,) + tuple(map_shape), dtype=dtype, device=device)
        self.out = torch.empty((self.world * self.per,) + tuple(map_shape), dtype=dtype, device=device)
        self.cuda = torch.device(device).type == "cuda"
        self._slots = None

    def __call__(self, local_maps):
        if self.world == 1:
            return local_maps[: self.n_clips]
        n = local_maps.shape[0]
        src = local_maps if (n == self.per and local_maps.is_contiguous()) else self.pad
        if src is self.pad:
            self.pad[:n].copy_(local_maps)
        dist.all_gather_into_tensor(self.out, src, group=self.group)
        return self.out[: self.n_clips]

    # ------------------------------------------------------------------ overlapped form
    def _make_slots(self):
        dev = self.out.device
        self._slots = []
        for k in range(2):
            self._slots.append({
                "out": self.out if k == 0 else torch.empty_like(self.out),
                "pad": self.pad if k == 0 else torch.zeros_like(self.pad),
                "done": torch.cuda.Event() if self.cuda else None,
                "busy": False,
            })
        self._stream = torch.cuda.Stream(device=dev) if self.cuda else None

    def start(self, local_maps, slot):
        """Enqueues the gather of ``local_maps`` behind the work already on the current stream, on the side stream."""
        if self._slots is None:
            self._make_slots()
        sl = self._slots[slot & 1]
        sl["local"] = local_maps
        if self.world == 1:
            sl["busy"] = True
            return
        n = local_maps.shape[0]
        direct = n == self.per and local_maps.is_contiguous()
        if not self.cuda:                                   # CPU tensors (gloo tests): nothing to overlap
            src = local_maps if direct else sl["pad"]
            if not direct:
                sl["pad"][:n].copy_(local_maps)
            dist.all_gather_into_tensor(sl["out"], src, group=self.group)
            sl["busy"] = True
            return
        cur = torch.cuda.current_stream(self.out.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            src = local_maps if direct else sl["pad"]
            if not direct:
                sl["pad"][:n].copy_(local_maps)
            dist.all_gather_into_tensor(sl["out"], src, group=self.group)
            sl["done"].record(self._stream)
        sl["busy"] = True

    def finish(self, slot):
        """Makes the current stream wait for the gather started on ``slot``; returns the [n_clips, ...] maps."""
        sl = self._slots[slot & 1]
        if not sl["busy"]:
            raise RuntimeError("MapGatherer.finish(%d) without a start" % slot)
        sl["busy"] = False
        if self.world == 1:
            return sl["local"][: self.n_clips]
        if self.cuda:
            torch.cuda.current_stream(self.out.device).wait_event(sl["done"])
        return sl["out"][: self.n_clips]

    def release(self, slot):
        """Current stream waits until ``slot``'s gather no longer reads its source (before the source is overwritten)."""
        if self._slots is not None and self._slots[slot & 1]["busy"]:
            self.finish(slot)


def gather_maps(local_maps, n_clips, group=None):
    """All ranks receive the [n_clips, 1, H, W] maps in clip order (one-shot convenience around MapGatherer)."""
    if not dist.is_available() or not dist.is_initialized():
        return local_maps[:n_clips]
    g = MapGatherer(n_clips, local_maps.shape[1:], local_maps.dtype, local_maps.device, group)
    return g(local_maps)
