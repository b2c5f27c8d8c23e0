"""Host-side image I/O (no third-party dependency: zlib from the standard library).

  read_png    what `Texture::from_png_file` (texture.rs:26-45) gets from the `png` crate 0.16.8 with its default
              transformations: 8-bit samples, palette / low bit depths expanded; grey is replicated so the result
              is always RGB8 or RGBA8 (the two layouts `Texture::read_texel` handles, texture.rs:47-63).
  write_png / write_ppm   the resolved 0xAARRGGBB framebuffer (rasterizer/buffers.rs:121-124) as a file -- the
              headless replacement of `Renderer::display` (render.rs:116-127), which only feeds a minifb window.

The C++ host mirror has the same three functions in rusterizer_b200/host/rz_image.hpp.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_SIG = b"\x89PNG\r\n\x1a\n"


def _chunk(kind: bytes, body: bytes) -> bytes:
    return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)


def framebuffer_to_rgb(fb: np.ndarray) -> np.ndarray:
    fb = np.ascontiguousarray(fb, dtype=np.uint32)
    rgb = np.empty(fb.shape + (3,), np.uint8)
    rgb[..., 0] = (fb >> 16) & 0xFF
    rgb[..., 1] = (fb >> 8) & 0xFF
    rgb[..., 2] = fb & 0xFF
    return rgb


def encode_png(pixels: np.ndarray, level: int = 6) -> bytes:
    """u8[h][w][3|4] -> PNG bytes (filter type 0)."""
    px = np.ascontiguousarray(pixels, dtype=np.uint8)
    h, w, c = px.shape
    assert c in (3, 4)
    raw = np.empty((h, 1 + w * c), np.uint8)
    raw[:, 0] = 0
    raw[:, 1:] = px.reshape(h, w * c)
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 6, 0, 0, 0)
    return _SIG + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b"")


def write_png(path, fb: np.ndarray):
    """Resolved framebuffer u32[H][W] (0xAARRGGBB) -> RGB8 PNG."""
    with open(path, "wb") as f:
        f.write(encode_png(framebuffer_to_rgb(fb)))


def write_ppm(path, fb: np.ndarray):
    rgb = framebuffer_to_rgb(fb)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (rgb.shape[1], rgb.shape[0]))
        f.write(rgb.tobytes())


def decode_png(data: bytes) -> np.ndarray:
    """PNG bytes -> u8[h][w][3|4].  Non-interlaced; bit depths 1, 2, 4, 8, 16; all five colour types."""
    if data[:8] != _SIG:
        raise ValueError("png: bad signature")
    pos, idat, plte, trns, hdr = 8, [], b"", b"", None
    while pos + 12 <= len(data):
        (n,) = struct.unpack(">I", data[pos:pos + 4])
        kind, body = data[pos + 4:pos + 8], data[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        if zlib.crc32(kind + body) & 0xFFFFFFFF != crc:
            raise ValueError("png: chunk CRC mismatch")
        if kind == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"PLTE":
            plte = body
        elif kind == b"tRNS":
            trns = body
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
        pos += 12 + n
    if hdr is None:
        raise ValueError("png: missing IHDR")
    w, h, depth, ctype, comp, filt, interlace = hdr
    if comp or filt:
        raise ValueError("png: unknown compression/filter method")
    if interlace:
        raise ValueError("png: interlaced images are not supported")
    samples = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}.get(ctype)
    if samples is None:
        raise ValueError("png: bad colour type")
    if not (depth == 8 or (depth == 16 and ctype != 3) or (depth in (1, 2, 4) and ctype in (0, 3))):
        raise ValueError("png: bad bit depth for the colour type")
    if w == 0 or h == 0:
        raise ValueError("png: missing IHDR")
    if w > 65535 or h > 65535:  # untrusted header: bound the allocation (host/rz_image.hpp does the same)
        raise ValueError("png: image larger than 65535 pixels on a side")
    bpp = (samples * depth + 7) // 8
    stride = (w * samples * depth + 7) // 8
    d = zlib.decompressobj()
    blob = d.decompress(b"".join(idat), (stride + 1) * h)  # output capped at what the header announces (zip bomb)
    if d.unconsumed_tail:
        raise ValueError("png: image data longer than the header says")
    raw = np.frombuffer(blob, np.uint8)
    if raw.size < (stride + 1) * h:
        raise ValueError("png: image data too short")
    raw = raw[: (stride + 1) * h].reshape(h, stride + 1)
    lines = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    for y in range(h):  # un-filter; rows of type 0/2 are vectorised, the recursive types run per byte
        ft, cur = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            out = cur
        elif ft == 2:
            out = (cur + prev) & 0xFF
        elif ft in (1, 3, 4):
            out = cur.copy()
            for i in range(stride):
                a = out[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if ft == 1:
                    add = a
                elif ft == 3:
                    add = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    add = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                out[i] = (out[i] + add) & 0xFF
        else:
            raise ValueError("png: bad filter type")
        lines[y] = out
        prev = out
    # samples at native depth -> [h][w*samples]
    if depth == 8:
        smp = lines.astype(np.uint32)
    elif depth == 16:
        smp = (lines[:, 0::2].astype(np.uint32) << 8) | lines[:, 1::2]
    else:
        bits = np.unpackbits(lines, axis=1)[:, : w * samples * depth].reshape(h, w * samples, depth)
        smp = np.zeros((h, w * samples), np.uint32)
        for k in range(depth):
            smp = (smp << 1) | bits[:, :, k]
    maxv = (1 << depth) - 1

    def to8(v):
        if depth == 8:
            return v.astype(np.uint8)
        if depth == 16:
            return (v >> 8).astype(np.uint8)
        return (v * 255 // maxv).astype(np.uint8)

    smp = smp.reshape(h, w, samples)
    alpha = ctype in (4, 6) or len(trns) > 0
    out = np.empty((h, w, 4 if alpha else 3), np.uint8)
    if alpha:
        out[..., 3] = 255
    if ctype == 3:
        pal = np.frombuffer(plte, np.uint8).reshape(-1, 3)
        idx = smp[..., 0]
        if idx.max(initial=0) >= len(pal):
            raise ValueError("png: palette index out of range")
        out[..., :3] = pal[idx]
        if trns:
            ta = np.full(len(pal), 255, np.uint8)
            ta[: len(trns)] = np.frombuffer(trns, np.uint8)[: len(pal)]
            out[..., 3] = ta[idx]
    elif ctype in (0, 4):
        out[..., 0] = out[..., 1] = out[..., 2] = to8(smp[..., 0])
        if ctype == 4:
            out[..., 3] = to8(smp[..., 1])
        elif len(trns) >= 2:
            out[..., 3] = np.where(smp[..., 0] == struct.unpack(">H", trns[:2])[0], 0, 255)
    else:
        out[..., :3] = to8(smp[..., :3])
        if ctype == 6:
            out[..., 3] = to8(smp[..., 3])
        elif len(trns) >= 6:
            key = struct.unpack(">HHH", trns[:6])
            hit = (smp[..., 0] == key[0]) & (smp[..., 1] == key[1]) & (smp[..., 2] == key[2])
            out[..., 3] = np.where(hit, 0, 255)
    return out


def read_png(path) -> np.ndarray:
    with open(path, "rb") as f:
        return decode_png(f.read())
