"""Host image I/O: PNG decode for `Texture::from_png_file` (texture.rs:26-45) and the
headless replacement of `Renderer::display` (render.rs:116-127), in the Python mirror (rusterizer_b200/image.py)
and in the C++ mirror (rusterizer_b200/host/rz_image.hpp).  CPU only."""
import struct
import subprocess
import zlib
from pathlib import Path

import numpy as np
import pytest

from rusterizer_b200 import image
from rusterizer_b200.texture import CHECKERBOARD_SHA256, Texture

ROOT = Path(__file__).resolve().parents[1]


def _png(w, h, depth, ctype, rows: bytes, filters=None, plte=None, trns=None, split_idat=False, level=6) -> bytes:
    """Hand-assembled PNG: `rows` are the packed scanlines (no filter bytes); filters[y] picks the filter type."""
    samples = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    stride = (w * samples * depth + 7) // 8
    bpp = (samples * depth + 7) // 8
    raw = bytearray()
    prev = bytes(stride)
    for y in range(h):
        line = rows[y * stride:(y + 1) * stride]
        ft = (filters or [0] * h)[y]
        out = bytearray(stride)
        for i in range(stride):
            a = line[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            if ft == 0:
                pred = 0
            elif ft == 1:
                pred = a
            elif ft == 2:
                pred = b
            elif ft == 3:
                pred = (a + b) >> 1
            else:
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            out[i] = (line[i] - pred) & 0xFF
        raw.append(ft)
        raw += out
        prev = line
    z = zlib.compress(bytes(raw), level)
    chunks = [image._chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))]
    if plte:
        chunks.append(image._chunk(b"PLTE", plte))
    if trns:
        chunks.append(image._chunk(b"tRNS", trns))
    if split_idat:
        chunks += [image._chunk(b"IDAT", z[: len(z) // 3]), image._chunk(b"IDAT", z[len(z) // 3:])]
    else:
        chunks.append(image._chunk(b"IDAT", z))
    chunks.append(image._chunk(b"IEND", b""))
    return image._SIG + b"".join(chunks)


@pytest.fixture(scope="module")
def image_tool(tmp_path_factory):
    exe = tmp_path_factory.mktemp("imgtool") / "image_tool"
    host = ROOT / "rusterizer_b200" / "host"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", str(host / "image_tool.cpp"), "-o", str(exe)], check=True)
    return exe


def _cpp_decode(image_tool, png_bytes, tmp_path):
    src, dst = tmp_path / "in.png", tmp_path / "out.raw"
    src.write_bytes(png_bytes)
    out = subprocess.run([str(image_tool), "decode", str(src), str(dst)], check=True, capture_output=True, text=True).stdout
    w, h, c = map(int, out.split())
    return np.frombuffer(dst.read_bytes(), np.uint8).reshape(h, w, c)


def _cases():
    rng = np.random.default_rng(7)
    cases = {}
    rgba = rng.integers(0, 256, (13, 17, 4), dtype=np.uint8)
    cases["rgba8_all_filters"] = (_png(17, 13, 8, 6, rgba.tobytes(), filters=[y % 5 for y in range(13)], split_idat=True), rgba)
    rgb = rng.integers(0, 256, (9, 31, 3), dtype=np.uint8)
    cases["rgb8_paeth"] = (_png(31, 9, 8, 2, rgb.tobytes(), filters=[4] * 9), rgb)
    cases["rgb8_stored_deflate"] = (_png(31, 9, 8, 2, rgb.tobytes(), level=0), rgb)
    g8 = rng.integers(0, 256, (5, 7), dtype=np.uint8)
    cases["grey8"] = (_png(7, 5, 8, 0, g8.tobytes(), filters=[3] * 5), np.repeat(g8[..., None], 3, 2))
    ga = rng.integers(0, 256, (5, 7, 2), dtype=np.uint8)
    cases["grey_alpha8"] = (_png(7, 5, 8, 4, ga.tobytes(), filters=[1] * 5),
                            np.concatenate([np.repeat(ga[..., :1], 3, 2), ga[..., 1:]], 2))
    # 16-bit RGB: the high byte survives (png crate STRIP_16)
    rgb16 = rng.integers(0, 65536, (4, 6, 3), dtype=np.uint16)
    cases["rgb16"] = (_png(6, 4, 16, 2, rgb16.astype(">u2").tobytes(), filters=[2] * 4), (rgb16 >> 8).astype(np.uint8))
    # 4-bit palette with transparency for the first two entries
    pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    idx = rng.integers(0, 16, (6, 9), dtype=np.uint8)
    packed = bytearray()
    for y in range(6):
        row = list(idx[y]) + [0]
        packed += bytes((row[i] << 4) | row[i + 1] for i in range(0, 10, 2))
    want = np.concatenate([pal[idx], np.full((6, 9, 1), 255, np.uint8)], 2)
    want[..., 3] = np.where(idx == 0, 10, np.where(idx == 1, 200, 255))
    cases["palette4_trns"] = (_png(9, 6, 4, 3, bytes(packed), plte=pal.tobytes(), trns=bytes([10, 200])), want)
    # 1-bit grey (0 -> 0, 1 -> 255)
    bits = rng.integers(0, 2, (3, 11), dtype=np.uint8)
    cases["grey1"] = (_png(11, 3, 1, 0, np.packbits(bits, axis=1).tobytes()), np.repeat((bits * 255)[..., None], 3, 2).astype(np.uint8))
    return cases


CASES = _cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_png_decode_python_and_cpp(name, image_tool, tmp_path):
    png, want = CASES[name]
    got = image.decode_png(png)
    assert got.shape == want.shape and np.array_equal(got, want), "python decoder"
    cpp = _cpp_decode(image_tool, png, tmp_path)
    assert cpp.shape == want.shape and np.array_equal(cpp, want), "C++ decoder"
    try:  # an independent decoder, when the image has one
        import io

        from PIL import Image
    except ImportError:
        return
    ref = np.asarray(Image.open(io.BytesIO(png)).convert("RGBA" if want.shape[2] == 4 else "RGB"))
    if name != "rgb16":  # PIL has no 16-bit RGB mode
        assert np.array_equal(ref, want), "fixture disagrees with PIL"


def test_checkerboard_png_round_trip_is_the_reference_fixture(image_tool, tmp_path):
    """The decoded bytes of images/checkerboard.png are pinned by sha256: a PNG of the
    procedural checkerboard decodes, through Texture.from_png_file and through the C++ loader, to those bytes."""
    tex = Texture.checkerboard()
    p = tmp_path / "checkerboard.png"
    p.write_bytes(image.encode_png(tex.texels))
    loaded = Texture.from_png_file(p)
    assert loaded.texel_width == 4 and loaded.sha256() == CHECKERBOARD_SHA256
    cpp = _cpp_decode(image_tool, p.read_bytes(), tmp_path)
    assert Texture(cpp).sha256() == CHECKERBOARD_SHA256


def test_framebuffer_writers(image_tool, tmp_path):
    """write_png / write_ppm (Python and C++) reproduce the 0xAARRGGBB framebuffer exactly (alpha dropped)."""
    rng = np.random.default_rng(3)
    fb = rng.integers(0, 2 ** 32, (37, 53), dtype=np.uint64).astype(np.uint32) | np.uint32(0xFF000000)
    want = image.framebuffer_to_rgb(fb)
    image.write_png(tmp_path / "a.png", fb)
    assert np.array_equal(image.read_png(tmp_path / "a.png"), want)
    image.write_ppm(tmp_path / "a.ppm", fb)
    ppm = (tmp_path / "a.ppm").read_bytes()
    assert ppm.startswith(b"P6\n53 37\n255\n") and ppm[len(b"P6\n53 37\n255\n"):] == want.tobytes()
    (tmp_path / "fb.u32").write_bytes(fb.tobytes())
    subprocess.run([str(image_tool), "encode", "53", "37", str(tmp_path / "fb.u32"), str(tmp_path / "b.png")], check=True)
    assert np.array_equal(image.read_png(tmp_path / "b.png"), want)  # stored-deflate PNG from the C++ writer
    subprocess.run([str(image_tool), "ppm", "53", "37", str(tmp_path / "fb.u32"), str(tmp_path / "b.ppm")], check=True)
    assert (tmp_path / "b.ppm").read_bytes() == ppm
    # a large image crosses the 65535-byte stored-block limit of the C++ encoder
    big = rng.integers(0, 2 ** 32, (300, 301), dtype=np.uint64).astype(np.uint32)
    (tmp_path / "big.u32").write_bytes(big.tobytes())
    subprocess.run([str(image_tool), "encode", "301", "300", str(tmp_path / "big.u32"), str(tmp_path / "big.png")], check=True)
    assert np.array_equal(image.read_png(tmp_path / "big.png"), image.framebuffer_to_rgb(big))


def test_png_errors(image_tool, tmp_path):
    good, _ = CASES["rgb8_paeth"]
    bad_crc = bytearray(good)
    bad_crc[40] ^= 0xFF
    with pytest.raises(ValueError):
        image.decode_png(bytes(bad_crc))
    with pytest.raises(ValueError):
        image.decode_png(b"not a png at all")
    (tmp_path / "bad.png").write_bytes(bytes(bad_crc))
    r = subprocess.run([str(image_tool), "decode", str(tmp_path / "bad.png"), str(tmp_path / "o.raw")], capture_output=True, text=True)
    assert r.returncode == 1 and "png:" in r.stderr


def _raw_png(w, h, idat_payload: bytes, depth=8, ctype=6) -> bytes:
    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) + chunk(b"IDAT", idat_payload)
            + chunk(b"IEND", b""))


def test_png_untrusted_header_and_zip_bomb(image_tool, tmp_path):
    """A crafted file must be rejected, not trusted: IHDR dimensions whose scanline arithmetic would wrap size_t
    (W = 2^30 RGBA16, H = 2^31) and image data that inflates far beyond what the header announces."""
    huge = _raw_png(1 << 30, 1 << 31, zlib.compress(b"\0" * 64), depth=16)
    bomb = _raw_png(4, 4, zlib.compress(b"\0" * (64 << 20), 9))  # 64 MiB of zeros behind a 4x4 header
    for name, blob in (("huge", huge), ("bomb", bomb)):
        with pytest.raises(ValueError):
            image.decode_png(blob)
        (tmp_path / f"{name}.png").write_bytes(blob)
        r = subprocess.run([str(image_tool), "decode", str(tmp_path / f"{name}.png"), str(tmp_path / "o.raw")], capture_output=True, text=True)
        assert r.returncode != 0 and "png:" in (r.stderr + r.stdout)
