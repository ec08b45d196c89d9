"""Batched localize-and-crop: the hot loop of ``process_spider`` / ``process_phenikaa``
(``datasets/classification/spider.py:90-152``, ``phenikaa.py:142-208``) run over many series
at once instead of one series per iteration with two device round trips each.

    slices (host, float32, ragged)  --H2D-->  K1 normalise+resize  -->  ConvNeXt localizer
        -->  K3 crop+letterbox (+ classifier-size resample)  --D2H-->  coords, crops

Series are independent, so multi-GPU is a plain shard by series index
(``shard_series``) plus one gather of crops and coordinates at the end
(``gather_results``); there is no per-op collective.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import ops
from .cropping import LocalizationModel, get_center_fallback_locations, get_rotation_angles, inverse_rotation, mm_to_pixels

NUM_LEVELS = 5


@dataclass
class CropBatch:
    coords: torch.Tensor  # float32 [B,5,2] (device)
    crops: torch.Tensor  # uint8 [B,5,ch,cw] (device)
    crops2: torch.Tensor | None  # uint8 [B,5,oh2,ow2] (device) -- classifier-size resample
    planes: torch.Tensor | None = None  # uint8 [B,H,W] K1 output (kept for inspection)
    times: dict = field(default_factory=dict)

    def to_host(self):
        return (self.coords.cpu().numpy(), self.crops.cpu().numpy(), None if self.crops2 is None else self.crops2.cpu().numpy())


def rotation_rows(c, shapes, last_disc_angle_boost: float = 1.0) -> torch.Tensor:
    """``rotation_table`` from coordinates already on the HOST (``c``: NumPy ``[B,L,2]``): float64 ``[B*L, 6]`` host tensor.
    Callers that pipeline chunks (the dataset driver) fetch the coordinates asynchronously and run this while the GPU works
    on the next chunk, so the exact host arithmetic (``np.polyfit``, as the reference) costs the GPU nothing."""
    rows = []
    for b, (h, w) in enumerate(shapes):
        locs = {i: (float(c[b, i, 0]), float(c[b, i, 1])) for i in range(c.shape[1])}
        ang = get_rotation_angles(locs, (h, w), last_disc_angle_boost)
        for i in range(c.shape[1]):
            rows.append(inverse_rotation(int(locs[i][0] * w), int(locs[i][1] * h), ang[i]))
    return torch.tensor(rows, dtype=torch.float64).reshape(-1, 6)


def rotation_table(coords: torch.Tensor, shapes, last_disc_angle_boost: float = 1.0) -> torch.Tensor:
    """Rotated crop mode (cropping.py:172-313): per (series, level) the inverse rotation OpenCV would apply, as
    float64 ``[B*L, 6]`` on the device.  The five-point angle fit is host arithmetic with the reference's own NumPy
    calls (``np.polyfit``), so this costs one small device->host read of the coordinates."""
    return rotation_rows(coords.detach().cpu().numpy(), shapes, last_disc_angle_boost).to(coords.device)


def crop_levels(pool: ops.SlicePool, coords: torch.Tensor, crop_delta_mm, spacings=None, crop_size=(128, 128),
                second_size=(256, 256), return_geom: bool = False, mode: str = "horizontal", last_disc_angle_boost: float = 1.0,
                inv_affine: torch.Tensor | None = None):
    """K3 over every (series, level): coords float32 [B,L,2] on the device.  ``spacings`` is a
    list of per-series (row, col) mm/px (``get_slice_spacing``, cropping.py:82-101); the
    reference always crops the 0.3 mm isotropic slice, so the default is (0.3, 0.3)."""
    B, L = int(coords.shape[0]), int(coords.shape[1])
    dev = coords.device
    if spacings is None:
        spacings = [(0.3, 0.3)] * B
    deltas = [mm_to_pixels(crop_delta_mm, sp) for sp in spacings]  # cropping.py:149-169, host ints
    delta = torch.tensor(deltas, dtype=torch.int32).repeat_interleave(L, dim=0).contiguous()
    max_box = (max(1, max(d[2] + d[3] for d in deltas)), max(1, max(d[0] + d[1] for d in deltas)))
    mh, mw = pool.max_hw
    max_box = (min(max_box[0], mh), min(max_box[1], mw))
    idx = torch.arange(B, dtype=torch.int32).repeat_interleave(L).contiguous()
    if mode not in ("horizontal", "rotated"):
        raise ValueError(f"unknown crop mode {mode!r}")
    inv = None
    if mode == "rotated":  # a table the caller computed ahead (rotation_rows on prefetched coordinates), or one D2H + host fit here
        inv = inv_affine.to(dev, non_blocking=True) if inv_affine is not None else rotation_table(coords, pool.shapes, last_disc_angle_boost)
    crops, crops2, geom = ops.crop_resample(pool, idx.to(dev, non_blocking=True), coords.reshape(B * L, 2).contiguous(),
                                            delta.to(dev, non_blocking=True), max_box, crop_size, second_size, return_geom,
                                            inv_affine=inv)
    crops = crops.view(B, L, *crops.shape[1:])
    if crops2 is not None:
        crops2 = crops2.view(B, L, *crops2.shape[1:])
    if geom is not None:
        geom = geom.view(B, L, 8)
    return crops, crops2, geom


def localize_and_crop(pool: ops.SlicePool, model: LocalizationModel | None, crop_delta_mm=(55, 15, 17.5, 20),
                      crop_size=(256, 256), image_size=(512, 512), second_size=(256, 256), spacings=None,
                      keep_planes: bool = False, times: dict | None = None, crop_mode: str = "horizontal",
                      last_disc_angle_boost: float = 1.0) -> CropBatch:
    """One batched pass of the hot path over a pool of middle slices already in HBM.
    ``model=None`` reproduces the reference's centre-crop fallback (__init__.py:194-197)."""
    dev = pool.data.device
    B = pool.n
    planes = None
    if model is not None:
        planes = ops.normalize_resize(pool, image_size)
        coords = model.predict_u8(planes, times)
    else:
        fb = get_center_fallback_locations()
        one = torch.tensor([fb[i] for i in range(NUM_LEVELS)], dtype=torch.float64)  # Python floats in the reference, not model outputs
        coords = one.unsqueeze(0).repeat(B, 1, 1).to(dev)
    crops, crops2, _ = crop_levels(pool, coords, crop_delta_mm, spacings, crop_size, second_size, mode=crop_mode,
                                   last_disc_angle_boost=last_disc_angle_boost)
    return CropBatch(coords, crops, crops2, planes if keep_planes else None, times or {})


def localize_and_crop_volumes(volumes, spacings, directions, model: LocalizationModel | None, device="cuda:0", **kw) -> CropBatch:
    """The per-series body of ``process_spider`` from the decoded volume on (spider.py:114-152): K0 (middle isotropic
    sagittal plane + its spacing) -> K1 -> localizer -> K3.  ``volumes``: ``[z, y, x]`` arrays
    (``sitk.GetArrayFromImage``), ``spacings`` / ``directions``: ``image.GetSpacing()`` / ``GetDirection()``."""
    from . import volumes as _vol

    pool, slice_spacings = _vol.midplane_resample(volumes, spacings, directions, device)
    return localize_and_crop(pool, model, spacings=slice_spacings, **kw)


# ------------------------------------------------------------------------------------------ streamed host path
class PinnedSeries:
    """A ragged batch of float32 middle slices staged once in pinned host memory (the layout of
    ``ops.SlicePool``): what the per-series loop of ``process_spider`` hands over after decoding."""

    def __init__(self, slices):
        self.host, self.offs, self.shapes = ops.SlicePool.pin(slices)
        self.offs = list(self.offs)
        self.ends = [o + (h * w + 3) // 4 * 4 for o, (h, w) in zip(self.offs, self.shapes)]

    @property
    def n(self) -> int:
        return len(self.shapes)

    @property
    def nbytes(self) -> int:
        return (self.ends[-1] if self.ends else 0) * 4


class PendingCrops:
    """Handle of one ``StreamedLocalizer.run_async`` call: the pinned host tensors are complete after ``result()``."""

    def __init__(self, out: dict, finished: torch.cuda.Event | None):
        self._out, self.finished = out, finished

    @property
    def device_coords(self) -> torch.Tensor:
        return self._out["coords"]

    @property
    def device_crops(self) -> torch.Tensor:
        return self._out["crops"]

    def result(self):
        if self.finished is not None:
            self.finished.synchronize()
        o = self._out
        return o["h_coords"], o["h_crops"], o.get("h_crops2")


class StreamedLocalizer:
    """End-to-end driver for host-resident series: the batch is cut into chunks of ``chunk`` series and
    the host->device copy of chunk i+1 (copy stream, pinned memory) runs under the kernels of chunk i
    (K1 -> localizer -> K3 on the compute stream); results go back on a third stream.  Same arithmetic
    as ``localize_and_crop``; only the schedule differs.

    ``run`` returns completed host tensors.  ``run_async`` returns a ``PendingCrops`` handle instead, so that a caller
    with more batches to do (a dataset is many batches) starts the next one before collecting the previous one: the first
    upload of batch k+1 then runs under the last kernels of batch k instead of in front of an idle GPU.  Two output slots
    alternate; a slot is reused only after its previous results have left the device."""

    def __init__(self, model: LocalizationModel | None, device="cuda:0", crop_delta_mm=(55, 15, 17.5, 20), crop_size=(256, 256),
                 image_size=(512, 512), second_size=(256, 256), chunk: int = 128):
        self.model, self.device = model, torch.device(device)
        self.crop_delta_mm, self.crop_size, self.image_size, self.second_size = crop_delta_mm, tuple(crop_size), tuple(image_size), second_size
        self.chunk = int(chunk)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.d2h_stream = torch.cuda.Stream(self.device)
        self._stage = [None, None]
        self._iso = None  # source-plane mode: the isotropic planes K0 writes (one buffer: K0 .. K3 of a chunk run in stream order)
        self._stage_free: list[torch.cuda.Event | None] = [None, None]  # last compute use of each staging buffer
        self._slots: list[dict] = [{}, {}]
        self._slot_busy: list[torch.cuda.Event | None] = [None, None]   # last D2H out of each output slot
        self._tables: dict = {}
        self._next_stage = 0

    @property
    def _out(self) -> dict:  # the slot `run` uses (kept for callers that read the device tensors after run())
        return self._slots[0]

    def _staging(self, j: int, nelem: int) -> torch.Tensor:
        buf = self._stage[j]
        if buf is None or buf.numel() < nelem:
            buf = torch.empty(nelem, dtype=torch.float32, device=self.device)
            self._stage[j] = buf
        return buf

    def _outputs(self, B: int, slot: int = 0):
        key = (B, self.crop_size, self.second_size)
        if self._slots[slot].get("key") != key:
            dev = self.device
            o = {"key": key,
                 "coords": torch.empty((B, NUM_LEVELS, 2), dtype=torch.float32, device=dev),
                 "crops": torch.empty((B, NUM_LEVELS, *self.crop_size), dtype=torch.uint8, device=dev),
                 "h_coords": torch.empty((B, NUM_LEVELS, 2), dtype=torch.float32).pin_memory(),
                 "h_crops": torch.empty((B, NUM_LEVELS, *self.crop_size), dtype=torch.uint8).pin_memory()}
            if self.second_size is not None:
                o["crops2"] = torch.empty((B, NUM_LEVELS, *self.second_size), dtype=torch.uint8, device=dev)
                o["h_crops2"] = torch.empty((B, NUM_LEVELS, *self.second_size), dtype=torch.uint8).pin_memory()
            self._slots[slot] = o
            self._slot_busy[slot] = None
        return self._slots[slot]

    def _index_tables(self, series: PinnedSeries, spacings, bounds):
        """Per-batch index tables (offsets inside each chunk, shapes, crop deltas, slice ids): built and uploaded once per
        (series, spacings) and kept -- a pinned allocation per call costs more than the kernels it feeds."""
        B, L, dev = series.n, NUM_LEVELS, self.device
        key = (id(series), B, self.chunk, None if spacings is None else tuple(map(tuple, spacings)), tuple(self.crop_delta_mm))
        t = self._tables.get("key") == key and self._tables
        if t:
            return t
        from_source = hasattr(series, "chunk_descs")  # volumes.PinnedVolumes: K0 runs on the device
        if spacings is None and from_source:
            spacings = series.spacings
        sp = spacings if spacings is not None else [(0.3, 0.3)] * B
        deltas = [mm_to_pixels(self.crop_delta_mm, s) for s in sp]
        iso_offs = series.out_offs if from_source else series.offs
        rel_offs = []
        for i0, i1 in bounds:
            rel_offs += [iso_offs[i] - iso_offs[i0] for i in range(i0, i1)]
        extra = {}
        if from_source:
            extra["descs"] = [series.chunk_descs(i0, i1) for i0, i1 in bounds]
        t = {"key": key, "series": series, "deltas": deltas, **extra,
             "offs": torch.tensor(rel_offs, dtype=torch.int64).pin_memory().to(dev, non_blocking=True),
             "hw": torch.tensor(series.shapes, dtype=torch.int32).reshape(-1, 2).pin_memory().to(dev, non_blocking=True),
             "delta": torch.tensor(deltas, dtype=torch.int32).repeat_interleave(L, dim=0).contiguous().pin_memory().to(dev, non_blocking=True),
             "idx": torch.arange(self.chunk, dtype=torch.int32).repeat_interleave(L).contiguous().pin_memory().to(dev, non_blocking=True)}
        self._tables = t
        return t

    def run(self, series: PinnedSeries, spacings=None):
        """Returns pinned host tensors ``(coords [B,5,2] f32, crops [B,5,ch,cw] u8, crops2 | None)``; they are
        complete when this call returns (it synchronises the result stream)."""
        return self.run_async(series, spacings, slot=0).result()

    def run_async(self, series, spacings=None, slot: int = 0) -> PendingCrops:
        """``series``: a ``PinnedSeries`` (isotropic middle slices, 5.7 MB per 1195^2 series over PCIe) or a
        ``volumes.PinnedVolumes`` (the two SOURCE planes per series, 2.1 MB at 512^2: K0 then runs on the device, fused
        with K1 through ``svb_k01_midplane_normalize_resize`` -- what the dataset driver's decode stage hands over)."""
        B, L, dev = series.n, NUM_LEVELS, self.device
        from_source = hasattr(series, "chunk_descs")
        o = self._outputs(B, slot)
        if B == 0:
            return PendingCrops(o, None)
        compute = torch.cuda.current_stream(dev)
        if self._slot_busy[slot] is not None:
            compute.wait_event(self._slot_busy[slot])  # the slot's previous results have left the device
        bounds = [(i0, min(i0 + self.chunk, B)) for i0 in range(0, B, self.chunk)]
        t = self._index_tables(series, spacings, bounds)
        deltas, offs_d, hw_d, delta_d, idx_d = t["deltas"], t["offs"], t["hw"], t["delta"], t["idx"]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        src_offs, src_ends = (series.vol_offs, series.vol_ends) if from_source else (series.offs, series.ends)
        max_elems = max(src_ends[i1 - 1] - src_offs[i0] for i0, i1 in bounds)
        j0 = self._next_stage  # staging buffers keep alternating across calls
        desc_d = [None, None]
        if from_source:
            iso_elems = max(series.out_offs[i1 - 1] + (series.shapes[i1 - 1][0] * series.shapes[i1 - 1][1] + 3) // 4 * 4 - series.out_offs[i0]
                            for i0, i1 in bounds)
            if self._iso is None or self._iso.numel() < iso_elems:
                self._iso = torch.empty(iso_elems, dtype=torch.float32, device=dev)

        def upload(c):
            i0, i1 = bounds[c]
            j = (j0 + c) & 1
            lo, hi = src_offs[i0], src_ends[i1 - 1]
            with torch.cuda.stream(self.copy_stream):
                if self._stage_free[j] is not None:
                    self.copy_stream.wait_event(self._stage_free[j])  # the chunk that last used this buffer is done with it
                buf = self._staging(j, max_elems)
                buf[: hi - lo].copy_(series.host[lo:hi], non_blocking=True)
                if from_source:
                    desc_d[c & 1] = t["descs"][c].to(dev, non_blocking=True)
                ready[c & 1].record(self.copy_stream)

        upload(0)
        for c, (i0, i1) in enumerate(bounds):
            j = (j0 + c) & 1
            if c + 1 < len(bounds):
                upload(c + 1)
            compute.wait_event(ready[c & 1])
            ready[c & 1] = torch.cuda.Event()
            n = i1 - i0
            shapes = series.shapes[i0:i1]
            coords = o["coords"][i0:i1]
            if from_source:
                # K0 (+ K1 when there is a model) on the device: the chunk's source planes -> isotropic planes -> uint8 planes
                pool = ops.SlicePool(self._iso, offs_d[i0:i1], hw_d[i0:i1], list(shapes))
                if self.model is not None:
                    planes = ops.midplane_normalize_resize(self._stage[j], desc_d[c & 1], pool, self.image_size)
                    self.model.predict_u8(planes, out=coords)
                else:
                    ops.midplane_resample_into(self._stage[j], desc_d[c & 1], pool)
                desc_d[c & 1].record_stream(compute)
            else:
                pool = ops.SlicePool(self._stage[j], offs_d[i0:i1], hw_d[i0:i1], list(shapes))
                if self.model is not None:
                    planes = ops.normalize_resize(pool, self.image_size)
                    self.model.predict_u8(planes, out=coords)
            xy = coords.view(n * L, 2)
            if self.model is None:
                # the fallback centres are Python floats in the reference: K3 gets them as float64 (int(x * w) sees the same double)
                fb = get_center_fallback_locations()
                fb64 = torch.tensor([fb[i] for i in range(L)], dtype=torch.float64, device=dev).unsqueeze(0).expand(n, L, 2)
                coords.copy_(fb64)
                xy = fb64.reshape(n * L, 2).contiguous()
            dl = deltas[i0:i1]
            mh, mw = pool.max_hw
            max_box = (min(max(1, max(d[2] + d[3] for d in dl)), mh), min(max(1, max(d[0] + d[1] for d in dl)), mw))
            ops.crop_resample(pool, idx_d[: n * L], xy, delta_d[i0 * L : i1 * L], max_box, self.crop_size,
                              self.second_size, out=o["crops"][i0:i1].view(n * L, *self.crop_size),
                              out2=None if self.second_size is None else o["crops2"][i0:i1].view(n * L, *self.second_size))
            done = torch.cuda.Event()
            done.record(compute)
            self._stage_free[j] = done
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(done)
                o["h_coords"][i0:i1].copy_(o["coords"][i0:i1], non_blocking=True)
                o["h_crops"][i0:i1].copy_(o["crops"][i0:i1], non_blocking=True)
                if self.second_size is not None:
                    o["h_crops2"][i0:i1].copy_(o["crops2"][i0:i1], non_blocking=True)
        self._next_stage = (j0 + len(bounds)) & 1
        finished = torch.cuda.Event()
        finished.record(self.d2h_stream)
        self._slot_busy[slot] = finished
        return PendingCrops(o, finished)


# ------------------------------------------------------------------------------------------ multi-GPU
def shard_series(sizes, world_size: int) -> list[list[int]]:
    """Greedy size-balanced partition of series indices (cost = H'*W' pixels) over ranks.
    Deterministic: every rank computes the same assignment with no communication."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    loads = [0] * world_size
    shards: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(sizes[i])
    for s in shards:
        s.sort()
    return shards


def gather_results(local_index: list[int], coords: torch.Tensor, crops: torch.Tensor, n_total: int, group=None):
    """The one exchange of the path: all ranks' (index, coords, crops) gathered on every rank
    with a padded ``all_gather`` (NCCL over NVLink on GPUs, gloo in the CPU tests) and
    scattered back into series order.  Returns ``(coords [n_total,5,2], crops [n_total,5,h,w])``."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out_c = torch.zeros((n_total,) + tuple(coords.shape[1:]), dtype=coords.dtype, device=coords.device)
        out_k = torch.zeros((n_total,) + tuple(crops.shape[1:]), dtype=crops.dtype, device=crops.device)
        if len(local_index):
            ix = torch.as_tensor(local_index, dtype=torch.long, device=coords.device)
            out_c[ix] = coords
            out_k[ix] = crops
        return out_c, out_k
    world = dist.get_world_size(group)
    dev = coords.device
    n_local = torch.tensor([len(local_index)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    n_max = max(int(c.item()) for c in counts)
    pad_idx = torch.full((n_max,), -1, dtype=torch.int64, device=dev)
    pad_c = torch.zeros((n_max,) + tuple(coords.shape[1:]), dtype=coords.dtype, device=dev)
    pad_k = torch.zeros((n_max,) + tuple(crops.shape[1:]), dtype=crops.dtype, device=dev)
    k = len(local_index)
    if k:
        pad_idx[:k] = torch.as_tensor(local_index, dtype=torch.int64, device=dev)
        pad_c[:k] = coords
        pad_k[:k] = crops
    all_idx = [torch.empty_like(pad_idx) for _ in range(world)]
    all_c = [torch.empty_like(pad_c) for _ in range(world)]
    all_k = [torch.empty_like(pad_k) for _ in range(world)]
    dist.all_gather(all_idx, pad_idx, group=group)
    dist.all_gather(all_c, pad_c, group=group)
    dist.all_gather(all_k, pad_k, group=group)
    out_c = torch.zeros((n_total,) + tuple(coords.shape[1:]), dtype=coords.dtype, device=dev)
    out_k = torch.zeros((n_total,) + tuple(crops.shape[1:]), dtype=crops.dtype, device=dev)
    for ix, c, kk in zip(all_idx, all_c, all_k):
        valid = ix >= 0
        out_c[ix[valid]] = c[valid]
        out_k[ix[valid]] = kk[valid]
    return out_c, out_k
