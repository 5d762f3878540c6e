"""CPU tests of the host-side file handling in front of the hot path: the TIFF decode that replaces `tifffile.imread`
(dataset.py:69,151-152,170-171 of the reference) and the Sen2Venus / flood tile pools built from it
(dataset.py:56-93, 99-113, 165-174).  Independent checks: files written by Pillow (libtiff) and decoded by Pillow, and files
written by a small writer below for the layouts Pillow cannot produce (4-band int16, band-sequential, tiles, predictor,
big-endian, BigTIFF)."""
import io
import os
import struct
import sys
import zlib

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vae-rs_b200")]
import tiff_reader  # noqa: E402


# ------------------------------------------------------------------------------------------------ a tiny TIFF writer (test side)
def write_tiff(a: np.ndarray, planar: bool, bo: str = "<", big: bool = False, tile=None, rows_per_strip=None,
               deflate: bool = False, predictor: bool = False) -> bytes:
    """a: [S, H, W] (written band-sequential if planar else pixel-interleaved) or [H, W]."""
    if a.ndim == 2:
        a = a[None]
    S, H, W = a.shape
    kind = {"u": 1, "i": 2, "f": 3}[a.dtype.kind]
    bits = a.dtype.itemsize * 8
    src = [a[s][:, :, None] for s in range(S)] if planar else [a.transpose(1, 2, 0)]       # per plane [H, W, spp]
    chunks = []
    for pl in src:
        if tile:
            th, tw = tile
            for y in range(0, H, th):
                for x in range(0, W, tw):
                    c = np.zeros((th, tw, pl.shape[2]), a.dtype)
                    blk = pl[y:y + th, x:x + tw]
                    c[:blk.shape[0], :blk.shape[1]] = blk
                    chunks.append(c)
        else:
            rps = rows_per_strip or H
            for y in range(0, H, rps):
                chunks.append(pl[y:y + rps])
    raw = []
    for c in chunks:
        c = np.ascontiguousarray(c)
        if predictor:
            d = c.copy()
            d[:, 1:] = c[:, 1:] - c[:, :-1]              # wraps modulo 2^bits
            c = d
        b = c.astype(a.dtype.newbyteorder(bo)).tobytes()
        raw.append(zlib.compress(b) if deflate else b)
    head = 16 if big else 8
    offs, pos = [], head
    for r in raw:
        offs.append(pos)
        pos += len(r)
    data = b"".join(raw)
    ifd_pos = head + len(data)
    entries = [(256, 4, [W]), (257, 4, [H]), (258, 3, [bits] * S), (259, 3, [8 if deflate else 1]), (277, 3, [S]),
               (284, 3, [2 if planar else 1]), (339, 3, [kind] * S), (317, 3, [2 if predictor else 1])]
    if tile:
        entries += [(322, 4, [tile[1]]), (323, 4, [tile[0]]), (324, 16 if big else 4, offs), (325, 16 if big else 4, [len(r) for r in raw])]
    else:
        entries += [(278, 4, [rows_per_strip or H]), (273, 16 if big else 4, offs), (279, 16 if big else 4, [len(r) for r in raw])]
    entries.sort()
    fmt = {3: "H", 4: "I", 16: "Q"}
    n_fmt, cnt_fmt, inline, ent = ("Q", "Q", 8, 20) if big else ("H", "I", 4, 12)
    extra_pos = ifd_pos + struct.calcsize(n_fmt) + ent * len(entries) + (8 if big else 4)
    body, extra = b"", b""
    for tag, typ, vals in entries:
        v = struct.pack(bo + fmt[typ] * len(vals), *vals)
        if len(v) <= inline:
            field = v.ljust(inline, b"\0")
        else:
            field = struct.pack(bo + ("Q" if big else "I"), extra_pos + len(extra))
            extra += v
        body += struct.pack(bo + "HH" + cnt_fmt, tag, typ, len(vals)) + field
    ifd = struct.pack(bo + n_fmt, len(entries)) + body + struct.pack(bo + ("Q" if big else "I"), 0) + extra
    mark = b"II" if bo == "<" else b"MM"
    header = mark + (struct.pack(bo + "HHHQ", 43, 8, 0, ifd_pos) if big else struct.pack(bo + "HI", 42, ifd_pos))
    return header + data + ifd


# ------------------------------------------------------------------------------------------------ decode
@pytest.mark.parametrize("compression", [None, "tiff_lzw", "tiff_adobe_deflate", "packbits"])
@pytest.mark.parametrize("mode", ["I;16", "RGBA", "L", "F"])
def test_decode_matches_pillow(mode, compression):
    """Files written by Pillow / libtiff decode to exactly what Pillow itself reads back."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    H, W = 77, 130                                       # several strips, none of them aligned to anything
    if mode == "I;16":
        a = (rng.integers(0, 3000, (H, W)) + 40 * np.arange(W)[None, :]).astype(np.uint16)     # smooth-ish: LZW finds repeats
    elif mode == "RGBA":
        a = rng.integers(0, 8, (H, W, 4)).astype(np.uint8) * 30
    elif mode == "L":
        a = (np.arange(H * W).reshape(H, W) // 7 % 251).astype(np.uint8)
    else:
        a = rng.standard_normal((H, W)).astype(np.float32)
    buf = io.BytesIO()
    Image.fromarray(a).save(buf, format="TIFF", **({"compression": compression} if compression else {}))
    got = tiff_reader.decode(buf.getvalue())
    ref = np.array(Image.open(io.BytesIO(buf.getvalue())))
    assert got.shape == ref.shape == a.shape and got.dtype == a.dtype
    assert np.array_equal(got, ref) and np.array_equal(got, a)


@pytest.mark.parametrize("bo", ["<", ">"])
@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("layout", ["strips", "strips_ragged", "tiles"])
@pytest.mark.parametrize("planar", [True, False])
def test_decode_multiband_int16(planar, layout, big, bo):
    """The Sen2Venus tile format family: 4-band int16, band-sequential or pixel-interleaved, strips (incl. a short last strip)
    or tiles that do not divide the image, Deflate + horizontal differencing, either byte order, classic and BigTIFF."""
    rng = np.random.default_rng(11)
    a = rng.integers(-3000, 12000, (4, 100, 72)).astype(np.int16)
    kw = dict(strips={}, strips_ragged=dict(rows_per_strip=33), tiles=dict(tile=(48, 32)))[layout]
    for deflate, predictor in ((False, False), (True, False), (True, True)):
        f = write_tiff(a, planar, bo=bo, big=big, deflate=deflate, predictor=predictor, **kw)
        got = tiff_reader.decode(f)
        want = a if planar else a.transpose(1, 2, 0)     # tifffile's shapes: [S, H, W] planar, [H, W, S] chunky
        assert got.dtype == np.int16 and got.shape == want.shape
        assert np.array_equal(got, want), (planar, layout, big, bo, deflate, predictor)


def test_decode_without_strip_byte_counts():
    """StripByteCounts is omitted by some old writers; uncompressed strips then run up to the next strip (or the end)."""
    a = np.arange(4 * 20 * 12, dtype=np.int16).reshape(4, 20, 12)
    f = bytearray(write_tiff(a, True, rows_per_strip=7))
    i = bytes(f).index(struct.pack("<HH", 279, 4))       # StripByteCounts entry -> an unknown private tag
    f[i:i + 2] = struct.pack("<H", 65000)
    assert np.array_equal(tiff_reader.decode(bytes(f)), a)


def test_decode_rejects_what_it_does_not_implement():
    a = np.zeros((8, 8), np.uint8)
    f = bytearray(write_tiff(a, False))
    with pytest.raises(tiff_reader.TiffError):
        tiff_reader.decode(b"PK" + bytes(f[2:]))
    i = bytes(f).index(struct.pack("<HHI", 259, 3, 1))   # Compression entry -> JPEG (7)
    f[i + 8:i + 10] = struct.pack("<H", 7)
    with pytest.raises(tiff_reader.TiffError, match="compression"):
        tiff_reader.decode(bytes(f))


# ------------------------------------------------------------------------------------------------ tile pools
def _make_sen2venus(tmp_path, n=5, dtype=np.int16, chunky_hr=False):
    from dataset import synthetic_tiles
    lr, hr = synthetic_tiles(n, 64, seed=9, as_int16=True)
    lr, hr = lr.numpy().astype(dtype), hr.numpy().astype(dtype)
    base = tmp_path / "ARM"
    (base / "tiles").mkdir(parents=True)
    lines = ["start_x\tb2b3b4b8_10m\tb2b3b4b8_05m\tb5b6b7b8a_20m"]
    for i in range(n):
        (base / "tiles" / f"t{i}_10m.tif").write_bytes(write_tiff(lr[i], True, deflate=True, predictor=dtype != np.float32))
        (base / "tiles" / f"t{i}_05m.tif").write_bytes(write_tiff(hr[i], not chunky_hr, deflate=True, rows_per_strip=16))
        lines.append(f"{i}\ttiles/t{i}_10m.tif\ttiles/t{i}_05m.tif\tunused.tif")
    (base / "index.csv").write_text("\n".join(lines) + "\n")
    return lr, hr


def test_sen2venus_tile_pool(tmp_path, monkeypatch):
    """dataset.py:99-113 / 165-174: <cwd>/ARM/index.csv, tab separated, the two 'visu' columns, row order kept; int16 tiles
    stay int16 (the patch gather reads them as such), pixel-interleaved files are brought to [C, H, W]."""
    from dataset import load_sen2venus_tiles
    lr, hr = _make_sen2venus(tmp_path, chunky_hr=True)
    monkeypatch.chdir(tmp_path)
    ds = load_sen2venus_tiles()
    assert ds.lr.dtype == ds.hr.dtype == torch.int16
    assert torch.equal(ds.lr, torch.from_numpy(lr)) and torch.equal(ds.hr, torch.from_numpy(hr))
    assert len(ds) == 5 and tuple(ds[3][0].shape) == (4, 32, 32) and tuple(ds[3][1].shape) == (4, 64, 64)
    assert len(load_sen2venus_tiles(root=str(tmp_path), limit=2)) == 2
    # float files: fp32 pool, as the reference's torch.tensor(img, dtype=float32)
    sub = tmp_path / "f"
    sub.mkdir()
    lr32, hr32 = _make_sen2venus(sub, n=2, dtype=np.float32)
    ds32 = load_sen2venus_tiles(root=str(sub))
    assert ds32.lr.dtype == torch.float32 and torch.equal(ds32.hr, torch.from_numpy(hr32))


def test_sen2venus_errors(tmp_path, monkeypatch):
    from dataset import init_dataloader, load_sen2venus_tiles
    monkeypatch.chdir(tmp_path)
    with pytest.raises(FileNotFoundError, match="index.csv"):
        load_sen2venus_tiles()
    with pytest.raises(FileNotFoundError):
        init_dataloader("s2v", 4, 64, device="cpu")      # the reference's default --dataset: fails on the missing index, nothing else
    with pytest.raises(ValueError, match="Unknown dataset"):
        init_dataloader("imagenet", 4, 64, device="cpu")
    base = tmp_path / "ARM"
    base.mkdir()
    a = np.zeros((4, 32, 32), np.int16)
    (base / "a.tif").write_bytes(write_tiff(a, True))
    (base / "index.csv").write_text("b2b3b4b8_10m\tb2b3b4b8_05m\na.tif\ta.tif\n")
    with pytest.raises(ValueError, match="twice"):
        load_sen2venus_tiles()
    (base / "index.csv").write_text("x\ty\na.tif\ta.tif\n")
    with pytest.raises(KeyError):
        load_sen2venus_tiles()


def test_flood_patches_follow_the_reference_arithmetic(tmp_path):
    """dataset.py:56-93: non-overlapping patches, per-band [1 %, 99 %] quantile scaling with the 1e-5 guard, clip, NaN patches
    dropped, partial patches at the right / bottom edge skipped."""
    from dataset import load_flood_patches
    rng = np.random.default_rng(3)
    img = (rng.random((3, 70, 100)) * 4000).astype(np.float32)
    img[1, 40, 10] = np.nan                              # poisons the patch at (32, 0)
    s2 = tmp_path / "event_a" / "S2"
    s2.mkdir(parents=True)
    (s2 / "x.tif").write_bytes(write_tiff(img, True, deflate=True))
    (s2 / "notes.txt").write_text("ignored")
    (tmp_path / "stray_file").write_text("ignored")
    got = load_flood_patches(str(tmp_path), patch_size=32)
    want = []
    for row in range(0, 70, 32):
        for col in range(0, 100, 32):
            if row + 32 <= 70 and col + 32 <= 100:
                p = img[:, row:row + 32, col:col + 32]
                q = np.quantile(p, [0.01, 0.99], axis=(1, 2), keepdims=True)
                p = torch.tensor(np.clip((p - q[0]) / (q[1] - q[0] + 1e-5), 0, 1), dtype=torch.float32)
                if not torch.isnan(p).any():
                    want.append(p)
    assert len(want) == 5 and tuple(got.shape) == (5, 3, 32, 32)
    assert torch.equal(got, torch.stack(want))
    with pytest.raises(FileNotFoundError):
        load_flood_patches(str(tmp_path / "missing"))
