"""Data side of the hot path: grid-mode patching + per-patch normalisation on the device.

Reference: dataset.py `select_crop` (:220-228), `grid_crop` (:230-247), `grid_collate` (:265-274),
`prepare_patches` (:249-262) and utils.normalize_image (utils.py:4-23) - a 256x256 HR tile (and its 128x128 LR
twin) becomes 16 patches of 64x64 (32x32), patch index = row*4 + col, each patch min-max normalised per channel,
batches ordered tile-major.  Here that is ONE kernel launch per tensor (svrs_grid_patch_normalize, one CTA per
patch, bit-exact fp32) instead of CPU slicing in DataLoader workers.

The GeoTIFF / CSV reading of the reference's Sen2VenDataset / FloodDataset (:49-218, tifffile + polars) is host-side IO
in front of the hot path: `load_sen2venus_tiles` / `load_flood_patches` decode the files ONCE into an in-memory tile pool
(tiff_reader.imread: `tifffile` when installed, else a numpy restatement of the baseline TIFF decode; the index is a plain
tab-separated file), `TileDataset` wraps tiles that are already in memory and `synthetic_tiles` produces the multispectral
test tiles used by the benchmark.
"""
from __future__ import annotations

import csv
import os
from typing import Iterator, Optional, Tuple

import numpy as np
import torch


def grid_patch_normalize(tiles: torch.Tensor, patch_size: int, out_dtype=torch.float32, nhwc: bool = False) -> torch.Tensor:
    """tiles [T,C,S,S] (fp32 or int16, CUDA) -> patches [T*(S/P)^2, C, P, P] (or [.., P, P, C] if nhwc)."""
    from svrs_native.lib import BF16, F32, lib

    if not tiles.is_cuda:
        raise RuntimeError("grid_patch_normalize: tiles must be on a CUDA device (no CPU fallback)")
    if tiles.dtype not in (torch.float32, torch.int16):
        tiles = tiles.float()
    tiles = tiles.contiguous()
    T, C, S, S2 = tiles.shape
    assert S == S2 and S % patch_size == 0
    n = T * (S // patch_size) ** 2
    shape = (n, patch_size, patch_size, C) if nhwc else (n, C, patch_size, patch_size)
    out = torch.empty(shape, device=tiles.device, dtype=out_dtype)
    lib.grid_patch_normalize(tiles.data_ptr(), int(tiles.dtype == torch.int16), out.data_ptr(),
                             F32 if out_dtype == torch.float32 else BF16, int(nhwc), T, C, S, patch_size,
                             torch.cuda.current_stream().cuda_stream)
    return out


def _gather(tiles: torch.Tensor, patch_size: int, origins, npatch: int, compute_dtype, want_nchw: bool, rt=None):
    """svrs_patch_gather_normalize (TMA gather): tiles [T,C,S,S] -> (nchw fp32 | None, PatchBatch)."""
    from svrs_native.engine import PatchBatch
    from svrs_native.lib import lib

    if not tiles.is_cuda:
        raise RuntimeError("patch gather: tiles must be on a CUDA device (no CPU fallback)")
    if tiles.dtype not in (torch.float32, torch.int16):
        tiles = tiles.float()
    tiles = tiles.contiguous()
    T, C, S, S2 = tiles.shape
    assert S == S2
    dev = tiles.device
    P = patch_size
    nchw = torch.empty((npatch, C, P, P), device=dev, dtype=torch.float32) if want_nchw else None
    f32 = torch.empty((npatch, P, P, C), device=dev, dtype=torch.float32)
    bf = torch.empty((npatch, P, P, C), device=dev, dtype=torch.bfloat16) if compute_dtype == torch.bfloat16 else None
    lib.patch_gather_normalize(tiles.data_ptr(), int(tiles.dtype == torch.int16), T, C, S, P,
                               None if origins is None else origins.data_ptr(), npatch,
                               None if nchw is None else nchw.data_ptr(), f32.data_ptr(),
                               None if bf is None else bf.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rt is not None:
        rt.launches += 1
    return nchw, PatchBatch(f32, f32 if bf is None else bf)


def grid_patch_pair(tiles: torch.Tensor, patch_size: int, compute_dtype=torch.float32, rt=None):
    """Grid-mode patches of `tiles` as the fused step consumes them (one launch, tiles read once): a PatchBatch holding the
    NHWC fp32 patches (NLL target) and their NHWC copy in the compute dtype (operand of the first conv layer)."""
    T, _, S, _ = tiles.shape
    assert S % patch_size == 0
    return _gather(tiles, patch_size, None, T * (S // patch_size) ** 2, compute_dtype, False, rt)[1]


def random_crop_origins(n_tiles: int, lr_size: int, patch_size: int, generator: Optional[torch.Generator] = None,
                        device="cpu") -> torch.Tensor:
    """The reference's random-crop draws (dataset.py:205-208): one crop per tile, top / left uniform in
    [0, lr_size - patch_size//2) on the LR tile.  Returns int32 [n_tiles, 3] = (tile, top, left) for the LR tiles; the
    HR origins are (tile, 2*top, 2*left) (dataset.py:212-216), see hr_origins()."""
    half = patch_size // 2
    top = torch.randint(0, lr_size - half, (n_tiles,), generator=generator)
    left = torch.randint(0, lr_size - half, (n_tiles,), generator=generator)
    o = torch.stack((torch.arange(n_tiles), top, left), dim=1).to(torch.int32)
    return o.to(device)


def hr_origins(lr_origins: torch.Tensor) -> torch.Tensor:
    o = lr_origins.clone()
    o[:, 1:] *= 2
    return o


def random_crop_batch(lr_tiles: torch.Tensor, hr_tiles: torch.Tensor, patch_size: int, lr_origins: torch.Tensor,
                      compute_dtype=torch.float32, want_nchw: bool = True):
    """On-device equivalent of Sen2VenDataset(crop="random") (dataset.py:140-218): the same-origin LR / HR crops of every
    tile at the given origins, min-max normalised per patch and channel (utils.py:4-23), gathered by TMA from the
    device-resident tile pool.  Returns ((y_nchw, y_batch), (x_nchw, x_batch))."""
    n = lr_origins.shape[0]
    lo = lr_origins.to(lr_tiles.device, torch.int32).contiguous()
    ho = hr_origins(lo).contiguous()
    y = _gather(lr_tiles, patch_size // 2, lo, n, compute_dtype, want_nchw)
    x = _gather(hr_tiles, patch_size, ho, n, compute_dtype, want_nchw)
    return y, x


def grid_batch(lr_tiles: torch.Tensor, hr_tiles: torch.Tensor, patch_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(LR tiles [T,4,S/2,S/2], HR tiles [T,4,S,S]) -> (y, x) patch batch in the (y, x) order of
    Cond_SRVAE.train_step (cond_vae.py:327)."""
    return grid_patch_normalize(lr_tiles, patch_size // 2), grid_patch_normalize(hr_tiles, patch_size)


def synthetic_tiles(n_tiles: int, size: int = 256, seed: int = 1, device="cpu", as_int16: bool = False):
    """Synthetic multispectral tile pairs of the Sen2Venus shape: HR [T,4,S,S] reflectance-like values and the
    2x box-downsampled LR twin [T,4,S/2,S/2] (keeps the SR pairing meaningful).  Deterministic in `seed`."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(n_tiles, 4, size // 8, size // 8, generator=g)
    hr = torch.nn.functional.interpolate(base, size=(size, size), mode="bilinear", align_corners=False)
    hr = (hr + 0.15 * torch.rand(n_tiles, 4, size, size, generator=g)) * 3000.0
    lr = torch.nn.functional.avg_pool2d(hr, 2)
    if as_int16:
        hr, lr = hr.round().to(torch.int16), lr.round().to(torch.int16)
    return lr.to(device), hr.to(device)


class TileDataset(torch.utils.data.Dataset):
    """In-memory (LR tile, HR tile) pairs; item i -> (lr [4,S/2,S/2], hr [4,S,S])."""

    def __init__(self, lr_tiles: torch.Tensor, hr_tiles: torch.Tensor):
        assert lr_tiles.shape[0] == hr_tiles.shape[0]
        self.lr, self.hr = lr_tiles, hr_tiles

    def __len__(self):
        return self.lr.shape[0]

    def __getitem__(self, i):
        return self.lr[i], self.hr[i]


def _as_chw(img: np.ndarray, path: str) -> np.ndarray:
    """tifffile returns [S, H, W] for band-sequential files and [H, W, S] for pixel-interleaved ones; the reference indexes
    tiles as [C, H, W] (dataset.py:70, 205)."""
    if img.ndim == 2:
        return img[None]
    if img.ndim != 3:
        raise ValueError(f"{path}: expected a 2-D or 3-D image, got shape {img.shape}")
    if img.shape[0] > 16 and img.shape[2] <= 16:          # [H, W, S] -> [S, H, W]
        return np.ascontiguousarray(img.transpose(2, 0, 1))
    return img


def load_sen2venus_tiles(dataset: str = "ARM", root: Optional[str] = None, limit: Optional[int] = None) -> TileDataset:
    """The reference's Sen2VenDataset file handling (dataset.py:99-113, 165-174) as a one-off decode into a tile pool:
    `<root or cwd>/<dataset>/index.csv` (tab-separated, header row) lists one tile pair per row in the columns
    `b2b3b4b8_10m` (LR, [4, S/2, S/2]) and `b2b3b4b8_05m` (HR, [4, S, S]), paths relative to the dataset directory
    (bands="visu", the only implemented choice, :107-110).  Row order is kept (the 80 / 20 split of init_dataloader is by
    index, :31-33).  int16 files stay int16 in the pool (the patch gather converts on the fly); anything else becomes fp32
    as in the reference (:154-155)."""
    import tiff_reader
    base = os.path.join(root if root is not None else os.getcwd(), dataset)
    index = os.path.join(base, "index.csv")
    if not os.path.isfile(index):
        raise FileNotFoundError(f"Sen2Venus index not found: {index} (the reference reads <cwd>/{dataset}/index.csv, dataset.py:99-101)")
    with open(index, newline="") as f:
        rows = list(csv.DictReader(f, delimiter="\t"))
    p0, p1 = "b2b3b4b8_10m", "b2b3b4b8_05m"
    if rows and (p0 not in rows[0] or p1 not in rows[0]):
        raise KeyError(f"{index}: columns {p0!r} and {p1!r} are required (bands='visu', dataset.py:107-110)")
    if limit is not None:
        rows = rows[:limit]
    if not rows:
        raise ValueError(f"{index}: no tile pairs listed")
    lr, hr = [], []
    for r in rows:
        a = _as_chw(tiff_reader.imread(os.path.join(base, r[p0])), r[p0])
        b = _as_chw(tiff_reader.imread(os.path.join(base, r[p1])), r[p1])
        if b.shape[0] != a.shape[0] or b.shape[1] != 2 * a.shape[1] or b.shape[2] != 2 * a.shape[2]:
            raise ValueError(f"{r[p0]} {a.shape} / {r[p1]} {b.shape}: the HR tile must be twice the LR tile (same-origin crops, dataset.py:205-216)")
        lr.append(a)
        hr.append(b)
    keep_i16 = all(t.dtype == np.int16 for t in lr + hr)

    def stack(ts):
        t = torch.from_numpy(np.stack(ts))
        return t if keep_i16 else t.to(torch.float32)

    return TileDataset(stack(lr), stack(hr))


def load_flood_patches(root: str = "/scratch/disc/e.bardet/Simple-VAE-RS/floods", patch_size: int = 256) -> torch.Tensor:
    """The reference's FloodDataset.precompute_patches (dataset.py:56-93) on the CPU, once: every `<root>/<event>/S2/*.tif`
    (all bands) is cut into non-overlapping patch_size^2 patches, each scaled per band to its own [1 %, 99 %] quantile range
    `(p - q01) / (q99 - q01 + 1e-5)`, clipped to [0, 1]; patches containing NaN are dropped.  -> fp32 [N, C, P, P]."""
    import tiff_reader
    if not os.path.isdir(root):
        raise FileNotFoundError(f"flood tiles not found: {root} (dataset.py:57)")
    patches = []
    for event in os.listdir(root):
        s2 = os.path.join(root, event, "S2")
        if not os.path.isdir(s2):
            continue
        for name in os.listdir(s2):
            if not name.endswith(".tif"):
                continue
            img = _as_chw(tiff_reader.imread(os.path.join(s2, name)), name)
            height, width = img.shape[1], img.shape[2]
            for row in range(0, height - patch_size + 1, patch_size):
                for col in range(0, width - patch_size + 1, patch_size):
                    patch = img[:, row:row + patch_size, col:col + patch_size]
                    q = np.quantile(patch, [0.01, 0.99], axis=(1, 2), keepdims=True)
                    patch = np.clip((patch - q[0]) / (q[1] - q[0] + 1e-5), 0, 1)
                    patch = torch.tensor(patch, dtype=torch.float32)
                    if not torch.isnan(patch).any():
                        patches.append(patch)
    if not patches:
        raise ValueError(f"{root}: no <event>/S2/*.tif tiles of at least {patch_size} x {patch_size} pixels")
    return torch.stack(patches)


class TilePrefetcher:
    """Double-buffered host->device feed: the pinned tile batch of step i+1 is copied on a dedicated copy stream while
    step i computes (what DataLoader(pin_memory=True) + .to(non_blocking=True) gives the reference).  `batches` yields
    (lr_host, hr_host) tensors of constant shape; iteration yields device tensors that stay valid until the next but one
    iteration.  Synchronisation is by CUDA events only (no host sync)."""

    def __init__(self, batches, device="cuda", depth: int = 2):
        self.batches, self.device, self.depth = batches, torch.device(device), depth

    def __iter__(self):
        it = iter(self.batches)
        copy_stream = torch.cuda.Stream(device=self.device)
        bufs, ready, free = [], [], []
        pending = []          # slots whose copy has been issued, in order

        def issue(slot):
            try:
                lr_h, hr_h = next(it)
            except StopIteration:
                return False
            if slot == len(bufs):
                bufs.append((torch.empty(lr_h.shape, dtype=lr_h.dtype, device=self.device),
                             torch.empty(hr_h.shape, dtype=hr_h.dtype, device=self.device)))
                ready.append(torch.cuda.Event())
                free.append(None)
            with torch.cuda.stream(copy_stream):
                if free[slot] is not None:
                    copy_stream.wait_event(free[slot])       # the step that last read this slot has finished
                bufs[slot][0].copy_(lr_h, non_blocking=True)
                bufs[slot][1].copy_(hr_h, non_blocking=True)
                ready[slot].record(copy_stream)
            pending.append(slot)
            return True

        nxt = 0
        for _ in range(self.depth - 1):
            if issue(nxt % self.depth):
                nxt += 1
        while True:
            if issue(nxt % self.depth):
                nxt += 1
            if not pending:
                return
            slot = pending.pop(0)
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[slot])
            yield bufs[slot]
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            free[slot] = ev


def _dist_info():
    """(rank, world) of the default process group, (0, 1) outside torch.distributed."""
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


class GridPatchLoader:
    """Iterates tile batches and emits (y, x) patch batches produced on the device - the on-device equivalent
    of DataLoader(Sen2VenDataset(crop="grid"), collate_fn=grid_collate).

    Data parallel (rank, world): every rank walks the SAME shuffled tile order (same seed) and takes its contiguous share
    (svrs_native.parallel.shard_bounds) of each GLOBAL batch of `tiles_per_batch` tiles; a ragged last global batch that
    cannot give every rank a tile is dropped, so all ranks run the same number of steps.  `sample_offset` of the batch most
    recently yielded = global index of this rank's first patch inside the global batch (Philox eps keyed by global sample)."""

    def __init__(self, tiles: TileDataset, tiles_per_batch: int, patch_size: int, device="cuda", shuffle: bool = False,
                 seed: int = 0, rank: int = 0, world: int = 1):
        self.ds, self.tpb, self.P, self.device, self.shuffle = tiles, tiles_per_batch, patch_size, torch.device(device), shuffle
        self._gen = torch.Generator().manual_seed(seed)
        self.rank, self.world = rank, world
        self.sample_offset = 0
        self.global_batch = 0
        if tiles_per_batch % world:
            raise ValueError(f"GridPatchLoader: {tiles_per_batch} tiles per global batch do not split evenly over {world} ranks")

    def _global_batches(self):
        n = len(self.ds)
        order = torch.randperm(n, generator=self._gen) if self.shuffle else torch.arange(n)
        out = [order[i:i + self.tpb] for i in range(0, n, self.tpb)]
        # equal shards on every rank (the KL terms are batch MEANS scaled by 1/world): trim each global batch
        # to a multiple of the world size
        out = [b[:len(b) // self.world * self.world] for b in out]
        return [b for b in out if len(b)]

    def __len__(self):
        n = len(self.ds)
        full, tail = divmod(n, self.tpb)
        return full + (1 if tail >= self.world else 0)

    def _shard(self, idx):
        from svrs_native.parallel import shard_bounds
        lo, hi = shard_bounds(len(idx), self.rank, self.world)
        per_tile = (self.ds.hr.shape[-1] // self.P) ** 2
        return idx[lo:hi], lo * per_tile, len(idx) * per_tile

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        batches = [self._shard(idx) for idx in self._global_batches()]
        if self.ds.lr.is_cuda:
            for idx, off, gb in batches:
                self.sample_offset, self.global_batch = off, gb
                yield grid_batch(self.ds.lr[idx.to(self.ds.lr.device)], self.ds.hr[idx.to(self.ds.hr.device)], self.P)
            return
        if not batches:
            return
        full_len = len(batches[0][0])

        def host_batches(items):
            for idx, _, _ in items:
                yield self.ds.lr[idx].pin_memory(), self.ds.hr[idx].pin_memory()

        # constant-shape batches go through the pinned, double-buffered prefetcher; a ragged last batch is a plain copy
        ragged = len(batches[-1][0]) != full_len
        head = batches[:-1] if ragged else batches
        for (idx, off, gb), (lr_d, hr_d) in zip(head, TilePrefetcher(host_batches(head), self.device)):
            self.sample_offset, self.global_batch = off, gb
            yield grid_batch(lr_d, hr_d, self.P)
        if ragged:
            idx, off, gb = batches[-1]
            self.sample_offset, self.global_batch = off, gb
            yield grid_batch(self.ds.lr[idx].to(self.device), self.ds.hr[idx].to(self.device), self.P)


class RandomCropLoader:
    """On-device equivalent of the reference's DEFAULT loader, DataLoader(Sen2VenDataset(crop="random"), batch_size, shuffle)
    (dataset.py:13-47, 140-218): every epoch visits each tile once, one random same-origin LR / HR crop per tile
    (LR at (top, left), HR at (2*top, 2*left), dataset.py:205-216), `batch_size` crops per batch, each crop min-max
    normalised per channel (utils.py:4-23).  The tile pool is resident in HBM (one upload); crops are gathered by TMA
    (svrs_patch_gather_normalize), no CPU slicing and no H2D copy per step.

    Data parallel: all ranks draw the SAME tile order and crop origins (same seed) for the global batch and keep their
    contiguous share, so the union over ranks is exactly the single-process batch (`drop_last` semantics for a global
    batch smaller than the world size)."""

    def __init__(self, tiles: TileDataset, batch_size: int, patch_size: int, device="cuda", shuffle: bool = True, seed: int = 0,
                 rank: int = 0, world: int = 1):
        self.P, self.bs, self.device, self.shuffle = patch_size, batch_size, torch.device(device), shuffle
        self.lr = tiles.lr.to(self.device).contiguous()
        self.hr = tiles.hr.to(self.device).contiguous()
        self._gen = torch.Generator().manual_seed(seed)
        self.rank, self.world = rank, world
        self.sample_offset = 0
        self.global_batch = 0
        if batch_size % world:
            raise ValueError(f"RandomCropLoader: a global batch of {batch_size} crops does not split evenly over {world} ranks")

    def _global_batches(self):
        n = self.lr.shape[0]
        order = torch.randperm(n, generator=self._gen) if self.shuffle else torch.arange(n)
        return [order[i:i + self.bs] for i in range(0, n, self.bs)]

    def __len__(self):
        full, tail = divmod(self.lr.shape[0], self.bs)
        return full + (1 if tail >= self.world else 0)

    def plan(self):
        """One epoch as a list of (origins int32 [n_local, 3] = (tile, top, left) on the LR grid, sample_offset, global batch
        size) - pure host logic (the draws), shared by __iter__ and the CPU tests."""
        from svrs_native.parallel import shard_bounds
        lr_size = self.lr.shape[-1]
        out = []
        for idx in self._global_batches():
            o = random_crop_origins(len(idx), lr_size, self.P, self._gen)      # draws for the WHOLE global batch
            o[:, 0] = idx.to(torch.int32)
            n = len(idx) // self.world * self.world      # equal shards on every rank (see GridPatchLoader): trim AFTER the
            if n == 0:                                   # draws, so the draws do not depend on the world size
                continue
            lo, hi = shard_bounds(n, self.rank, self.world)
            out.append((o[lo:hi].contiguous(), lo, n))
        return out

    def __iter__(self):
        for o, lo, n in self.plan():
            self.sample_offset, self.global_batch = lo, n
            (y, _), (x, _) = random_crop_batch(self.lr, self.hr, self.P, o, torch.float32)
            yield y, x


class PatchTensorLoader:
    """DataLoader(FloodDataset, batch_size, shuffle) of the reference (dataset.py:27-46): batches of the precomputed,
    already normalised patches [b, C, P, P], as single tensors (the reference's FloodDataset yields no LR / HR pair).  The
    patch pool lives on `device`; under torchrun every rank keeps its contiguous share of each global batch."""

    def __init__(self, patches: torch.Tensor, batch_size: int, device="cuda", shuffle: bool = False, seed: int = 0,
                 rank: int = 0, world: int = 1):
        self.x, self.bs, self.shuffle = patches.to(device), batch_size, shuffle
        self._gen = torch.Generator().manual_seed(seed)
        self.rank, self.world = rank, world

    def __len__(self):
        return -(-self.x.shape[0] // self.bs)

    def __iter__(self):
        from svrs_native.parallel import shard_bounds
        n = self.x.shape[0]
        order = torch.randperm(n, generator=self._gen) if self.shuffle else torch.arange(n)
        for i in range(0, n, self.bs):
            idx = order[i:i + self.bs]
            m = len(idx) // self.world * self.world
            if m == 0:
                continue
            lo, hi = shard_bounds(m, self.rank, self.world)
            yield self.x[idx[lo:hi].to(self.x.device)]


def init_dataloader(dataset: str, batch_size: int = 16, patch_size: int = 64, device="cuda", n_tiles: int = 64,
                    crop: str = "random", seed: int = 0):
    """Same entry point as the reference (dataset.py:13-47): returns (train_loader, val_loader), 80 / 20 split by index,
    batches are (y, x) = (LR, HR) patch tensors.  `crop="random"` (the reference's only reachable mode, dataset.py:24):
    batch_size counts crops, exactly as in the reference; `crop="grid"`: every tile yields its (256/P)^2 grid patches and
    batch_size (in patches) is rounded DOWN to whole tiles (at least one).  Under torchrun the loaders shard every global
    batch by rank (see the loader classes).  Datasets, named as in the reference (:23-29): "Sen2Venus" / "sen2venus" / "s2v"
    = the tile pairs listed in <cwd>/ARM/index.csv, decoded once into a device-resident tile pool (load_sen2venus_tiles);
    "Floods" / "floods" = the reference's quantile-normalised 256^2 flood patches (load_flood_patches, single tensors);
    "synthetic" (not in the reference) = in-memory multispectral test tiles."""
    rank, world = _dist_info()
    if dataset in ("Floods", "floods"):
        patches = load_flood_patches(patch_size=256)                  # dataset.py:27: always 256
        split = int(0.8 * patches.shape[0])
        return (PatchTensorLoader(patches[:split], batch_size, device, shuffle=True, seed=seed, rank=rank, world=world),
                PatchTensorLoader(patches[split:], batch_size, device, shuffle=False))
    if dataset in ("Sen2Venus", "sen2venus", "s2v"):
        tiles = load_sen2venus_tiles()
        lr, hr = tiles.lr, tiles.hr
        n_tiles = lr.shape[0]
        split = int(0.8 * n_tiles)                                    # dataset.py:31
    elif dataset == "synthetic":
        lr, hr = synthetic_tiles(n_tiles)
        split = max(1, int(0.8 * n_tiles))
    else:
        raise ValueError(f"Unknown dataset: {dataset}")               # dataset.py:29-30
    tr_ds, va_ds = TileDataset(lr[:split], hr[:split]), TileDataset(lr[split:], hr[split:])
    if crop == "random":
        train = RandomCropLoader(tr_ds, batch_size, patch_size, device, shuffle=True, seed=seed, rank=rank, world=world)
        val = RandomCropLoader(va_ds, batch_size, patch_size, device, shuffle=False, seed=seed + 1)   # validation is replicated
        return train, val
    if crop != "grid":
        raise ValueError(f"crop must be 'random' or 'grid', got {crop!r}")
    per_tile = (256 // patch_size) ** 2
    tpb = max(1, batch_size // per_tile)
    tpb = (tpb + world - 1) // world * world          # whole tiles on every rank
    train = GridPatchLoader(tr_ds, tpb, patch_size, device, shuffle=True, seed=seed, rank=rank, world=world)
    val = GridPatchLoader(va_ds, max(1, batch_size // per_tile), patch_size, device)
    return train, val
