"""Minimal TIFF decoder for the tile files the reference reads with `tifffile.imread` (dataset.py:69, 151-152, 170-171).

`tifffile` / `imagecodecs` are not installed in this image, and the training tiles (Sen2Venus: 4-band int16 GeoTIFFs of
256x256 / 128x128 pixels) only need the baseline of the format, so the decode is restated here on numpy + zlib:

  * classic TIFF and BigTIFF, either byte order; the first image file directory only (GeoTIFF tags are ignored)
  * strips or tiles; PlanarConfiguration chunky (1) or separate (2)
  * Compression none (1), LZW (5), Deflate (8 / 32946), PackBits (32773); Predictor none (1) or horizontal differencing (2)
  * 8 / 16 / 32 / 64-bit unsigned, signed and IEEE samples

`imread(path)` returns what `tifffile.imread` returns for such files: [H, W] for one sample per pixel, [H, W, S] for
chunky multi-sample images, [S, H, W] for planar ones.  If `tifffile` is importable it is used instead.  This is host-side
IO before the hot path ("TIFF decode stays on CPU"); nothing here runs per training step.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Tuple

import numpy as np

_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
             16: "Q", 17: "q", 18: "Q"}

(T_WIDTH, T_LENGTH, T_BITS, T_COMPRESSION, T_STRIP_OFFSETS, T_SPP, T_ROWS_PER_STRIP, T_STRIP_COUNTS, T_PLANAR, T_PREDICTOR,
 T_TILE_W, T_TILE_L, T_TILE_OFFSETS, T_TILE_COUNTS, T_SAMPLE_FORMAT) = (256, 257, 258, 259, 273, 277, 278, 279, 284, 317,
                                                                        322, 323, 324, 325, 339)


class TiffError(ValueError):
    pass


def _read_ifd(buf: bytes) -> Tuple[str, Dict[int, tuple]]:
    """-> (byte-order prefix for struct, {tag: values}) of the first image file directory."""
    if len(buf) < 8 or buf[:2] not in (b"II", b"MM"):
        raise TiffError("not a TIFF file (bad byte-order mark)")
    bo = "<" if buf[:2] == b"II" else ">"
    magic = struct.unpack_from(bo + "H", buf, 2)[0]
    if magic == 42:
        big, (ifd,) = False, struct.unpack_from(bo + "I", buf, 4)
        n_fmt, ent_size, cnt_fmt, inline = "H", 12, "I", 4
    elif magic == 43:
        big, (ifd,) = True, struct.unpack_from(bo + "Q", buf, 8)
        n_fmt, ent_size, cnt_fmt, inline = "Q", 20, "Q", 8
    else:
        raise TiffError(f"not a TIFF file (magic {magic})")
    (n,) = struct.unpack_from(bo + n_fmt, buf, ifd)
    pos = ifd + struct.calcsize(n_fmt)
    tags: Dict[int, tuple] = {}
    for i in range(n):
        e = pos + i * ent_size
        tag, typ = struct.unpack_from(bo + "HH", buf, e)
        (count,) = struct.unpack_from(bo + cnt_fmt, buf, e + 4)
        fmt = _TYPE_FMT.get(typ)
        if fmt is None:
            continue                                     # unknown field type: not one of ours
        size = struct.calcsize("=" + fmt) * count
        val_pos = e + 4 + struct.calcsize(cnt_fmt)
        if size > inline:
            (val_pos,) = struct.unpack_from(bo + ("Q" if big else "I"), buf, val_pos)
        if val_pos + size > len(buf):
            raise TiffError(f"tag {tag}: value outside the file")
        if fmt == "c":
            tags[tag] = (buf[val_pos:val_pos + count],)
        else:
            tags[tag] = struct.unpack_from(bo + fmt * count, buf, val_pos)
    return bo, tags


def _lzw_decode(data: bytes) -> bytes:
    """TIFF LZW (MSB-first codes, 9..12 bits, ClearCode 256, EndOfInformation 257, 'early change')."""
    out = bytearray()
    table: List[bytes] = [bytes((i,)) for i in range(256)] + [b"", b""]
    bits, nbits, width = 0, 0, 9
    prev = None
    for byte in data:
        bits = (bits << 8) | byte
        nbits += 8
        while nbits >= width:
            code = (bits >> (nbits - width)) & ((1 << width) - 1)
            nbits -= width
            if code == 256:
                del table[258:]
                width, prev = 9, None
                continue
            if code == 257:
                return bytes(out)
            if prev is None:
                entry = table[code]
            elif code < len(table):
                entry = table[code]
                table.append(prev + entry[:1])
            elif code == len(table):
                entry = prev + prev[:1]
                table.append(entry)
            else:
                raise TiffError("corrupt LZW stream")
            out += entry
            prev = entry
            if len(table) >= (1 << width) - 1 and width < 12:       # early change: widen one code before the table is full
                width += 1
    return bytes(out)


def _packbits_decode(data: bytes) -> bytes:
    out = bytearray()
    i, n = 0, len(data)
    while i < n:
        h = data[i]
        i += 1
        if h < 128:
            out += data[i:i + h + 1]
            i += h + 1
        elif h > 128:
            out += data[i:i + 1] * (257 - h)
            i += 1
    return bytes(out)


def _decompress(chunk: bytes, compression: int) -> bytes:
    if compression == 1:
        return chunk
    if compression in (8, 32946):
        return zlib.decompress(chunk)
    if compression == 5:
        return _lzw_decode(chunk)
    if compression == 32773:
        return _packbits_decode(chunk)
    raise TiffError(f"unsupported TIFF compression {compression}")


def _dtype(bo: str, bits: int, sample_format: int) -> np.dtype:
    kind = {1: "u", 2: "i", 3: "f"}.get(sample_format)
    if kind is None or bits not in (8, 16, 32, 64) or (kind == "f" and bits < 32):
        raise TiffError(f"unsupported sample type (bits {bits}, format {sample_format})")
    return np.dtype(f"{bo}{kind}{bits // 8}")


def decode(buf: bytes) -> np.ndarray:
    bo, t = _read_ifd(buf)

    def one(tag, default=None):
        v = t.get(tag)
        if v is None:
            if default is None:
                raise TiffError(f"missing TIFF tag {tag}")
            return default
        return v[0]

    W, H = int(one(T_WIDTH)), int(one(T_LENGTH))
    spp = int(one(T_SPP, 1))
    bits_all = t.get(T_BITS, (1,))
    fmt_all = t.get(T_SAMPLE_FORMAT, (1,))
    if len(set(bits_all)) != 1 or len(set(fmt_all)) != 1:
        raise TiffError("samples of different types in one pixel are not supported")
    dt = _dtype(bo, int(bits_all[0]), int(fmt_all[0]))
    compression = int(one(T_COMPRESSION, 1))
    planar = int(one(T_PLANAR, 1)) if spp > 1 else 1
    predictor = int(one(T_PREDICTOR, 1))
    if predictor not in (1, 2):
        raise TiffError(f"unsupported TIFF predictor {predictor}")
    if predictor == 2 and dt.kind == "f":
        raise TiffError("horizontal differencing on floating-point samples is not supported")
    planes = spp if planar == 2 else 1                  # separately stored sample planes
    chunk_spp = 1 if planar == 2 else spp               # samples per pixel inside one chunk
    if T_TILE_OFFSETS in t:
        cw, ch = int(one(T_TILE_W)), int(one(T_TILE_L))
        offsets, counts = t[T_TILE_OFFSETS], t[T_TILE_COUNTS]
    else:
        cw, ch = W, min(int(one(T_ROWS_PER_STRIP, H)), H)
        if T_STRIP_OFFSETS not in t:
            raise TiffError("neither strips nor tiles")
        offsets = t[T_STRIP_OFFSETS]
        counts = t.get(T_STRIP_COUNTS)
        if counts is None:                              # some writers omit StripByteCounts: every strip runs to the next one
            ends = sorted(offsets) + [len(buf)]
            counts = tuple(ends[ends.index(o) + 1] - o for o in offsets)
    across, down = -(-W // cw), -(-H // ch)
    if len(offsets) != across * down * planes or len(counts) != len(offsets):
        raise TiffError(f"expected {across * down * planes} chunks, the file lists {len(offsets)}")
    tiled = T_TILE_OFFSETS in t
    out = np.zeros((planes, H, W, chunk_spp), dtype=dt.newbyteorder("="))
    k = 0
    for pl in range(planes):
        for cy in range(down):
            for cx in range(across):
                off, cnt = int(offsets[k]), int(counts[k])
                k += 1
                if off + cnt > len(buf):
                    raise TiffError("chunk outside the file")
                raw = _decompress(buf[off:off + cnt], compression)
                rows = ch if tiled else min(ch, H - cy * ch)           # the last strip holds only the remaining rows
                need = rows * cw * chunk_spp * dt.itemsize
                if len(raw) < need:
                    raise TiffError(f"chunk {k - 1}: {len(raw)} bytes decoded, {need} expected")
                a = np.frombuffer(raw, dtype=dt, count=rows * cw * chunk_spp).reshape(rows, cw, chunk_spp)
                if predictor == 2:
                    a = np.cumsum(a, axis=1, dtype=dt.newbyteorder("="))      # wraps modulo 2^bits, as the encoder's differences do
                y0, x0 = cy * ch, cx * cw
                h, w = min(rows, H - y0), min(cw, W - x0)
                out[pl, y0:y0 + h, x0:x0 + w, :] = a[:h, :w, :]
    if spp == 1:
        return out[0, :, :, 0]
    if planar == 2:
        return out[:, :, :, 0]
    return out[0]


def imread(path: str) -> np.ndarray:
    """`tifffile.imread(path)` for the tile files of the training sets; falls back to the decoder above when `tifffile` is
    not installed."""
    try:
        import tifffile                                   # the reference's reader, when present
    except ImportError:
        tifffile = None
    if tifffile is not None:
        return tifffile.imread(path)
    with open(path, "rb") as f:
        return decode(f.read())
