"""CPU tests (-m "not gpu") of the host-side data-parallel loader logic: every rank walks the same shuffled order and
crop draws and keeps a contiguous, equal share of each GLOBAL batch, so the union over ranks is the single-process batch
(VERDICT r1 weak #4 / ADVICE high: `torchrun train.py` used to train every rank on identical data)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vae-rs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import dataset as D  # noqa: E402


def _tiles(n):
    lr, hr = D.synthetic_tiles(n, 64, seed=3)
    return D.TileDataset(lr, hr)


def test_random_crop_loader_shards_are_a_partition_of_the_global_batch():
    ds = _tiles(22)
    single = D.RandomCropLoader(ds, 8, 32, device="cpu", shuffle=True, seed=5).plan()
    for world in (2, 4):
        ranks = [D.RandomCropLoader(ds, 8, 32, device="cpu", shuffle=True, seed=5, rank=r, world=world).plan() for r in range(world)]
        assert all(len(p) == len(ranks[0]) for p in ranks)              # same number of steps on every rank
        for step in range(len(ranks[0])):
            parts = [ranks[r][step] for r in range(world)]
            glob = torch.cat([o for o, _, _ in parts])
            n = parts[0][2]
            assert all(p[2] == n for p in parts) and glob.shape[0] == n
            assert all(p[0].shape[0] == n // world for p in parts)      # equal shards (KL terms are means scaled by 1/world)
            assert [p[1] for p in parts] == [r * (n // world) for r in range(world)]     # sample_offset = global index
            # the union over ranks is the single-process batch of the same step, trimmed to a multiple of the world size
            assert torch.equal(glob, single[step][0][:n])
    # the reference's draws (dataset.py:205-208): top, left in [0, lr_size - P/2)
    o = torch.cat([o for o, _, _ in single])
    assert int(o[:, 1:].min()) >= 0 and int(o[:, 1:].max()) < 32 - 16
    assert sorted(o[:, 0].tolist()) == list(range(22))                   # every tile once per epoch


def test_grid_loader_shards_tiles_by_rank():
    ds = _tiles(10)
    for world in (1, 2):
        ls = [D.GridPatchLoader(ds, 4, 32, "cpu", True, 0, r, world) for r in range(world)]
        plans = [[l._shard(b) for b in l._global_batches()] for l in ls]
        assert len({len(p) for p in plans}) == 1 and len(plans[0]) == len(ls[0])
        for step in range(len(plans[0])):
            tiles = torch.cat([plans[r][step][0] for r in range(world)])
            assert len(set(tiles.tolist())) == len(tiles)               # disjoint shards
            assert len({len(plans[r][step][0]) for r in range(world)}) == 1
            per_tile = (64 // 32) ** 2
            assert [plans[r][step][1] for r in range(world)] == [r * len(plans[0][step][0]) * per_tile for r in range(world)]
    try:
        D.GridPatchLoader(ds, 3, 32, "cpu", rank=0, world=2)
        raise AssertionError("uneven split must be refused")
    except ValueError:
        pass


def test_patch_tensor_loader_shards_and_order():
    """PatchTensorLoader (the reference's DataLoader over FloodDataset, dataset.py:27-46): unshuffled batches in index order
    incl. the ragged last one; under `world` ranks the shards of every global batch are equal, contiguous, and their union
    is the single-process batch trimmed to a multiple of the world size."""
    x = torch.arange(23, dtype=torch.float32).reshape(23, 1, 1, 1).expand(23, 2, 4, 4).contiguous()
    plain = list(D.PatchTensorLoader(x, 5, device="cpu"))
    assert len(plain) == len(D.PatchTensorLoader(x, 5, device="cpu")) == 5
    assert torch.equal(torch.cat(plain), x) and plain[-1].shape[0] == 3
    single = list(D.PatchTensorLoader(x, 6, device="cpu", shuffle=True, seed=4))
    assert sorted(torch.cat(single)[:, 0, 0, 0].tolist()) == list(range(23))          # a permutation: every patch once
    for world in (2, 3):
        ranks = [list(D.PatchTensorLoader(x, 6, device="cpu", shuffle=True, seed=4, rank=r, world=world)) for r in range(world)]
        assert len({len(b) for b in ranks}) == 1
        for step in range(len(ranks[0])):
            parts = [ranks[r][step] for r in range(world)]
            assert len({p.shape[0] for p in parts}) == 1
            n = sum(p.shape[0] for p in parts)
            assert torch.equal(torch.cat(parts), single[step][:n])
