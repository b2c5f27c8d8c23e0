"""Second, independently written restatement of the reference raster path -- numpy, vectorised over the pixels of
one triangle's bounding box -- used ONLY to cross-check oracle/rz_oracle.c (tests/test_oracle_crosscheck.py).

TEST INFRASTRUCTURE, like everything under oracle/: nothing under rusterizer_b200/ may import it.

Why it exists: the reference is Rust and cannot be built in this image, so whole-frame behaviour (submission order,
post-depth shading position, texture sampling, resolve) is pinned by source only ("parity unpinned" in
DESIGN.md section 2).  Two restatements written separately from the same source -- a scalar C loop nest and this
array program -- that agree bit for bit on every depth sample, every packed colour sample and every resolved pixel
make a transcription slip in either one unlikely.  Every function cites the reference lines it follows
(paths relative to /root/reference/src).

Arithmetic: every + - * / below is one numpy float32 operation (IEEE binary32, round to nearest even, never
fused: numpy evaluates each ufunc separately), in the reference's source order.
"""
from __future__ import annotations

import numpy as np

F = np.float32
ZERO, ONE, HALF = F(0.0), F(1.0), F(0.5)
CLEAR_COLOR = np.uint32(0xFF191919)  # rasterizer/buffers.rs:5
CLEAR_DEPTH = np.finfo(np.float32).max  # rasterizer/buffers.rs:6
# rasterizer/mod.rs:109-114
RGSS = [(F(5.0) / F(8.0), F(1.0) / F(8.0)), (F(7.0) / F(8.0), F(5.0) / F(8.0)),
        (F(3.0) / F(8.0), F(7.0) / F(8.0)), (F(1.0) / F(8.0), F(3.0) / F(8.0))]
# runtime sample counts (extension; N_MSAA_SAMPLES is the constant 4 in the reference, mod.rs:23):
# 1 = pixel centre, 2 and 8 = the D3D11 standard patterns in sixteenths of a pixel
PATTERNS = {
    1: [(F(0.5), F(0.5))],
    2: [(F(0.75), F(0.75)), (F(0.25), F(0.25))],
    4: RGSS,
    8: [(F(k[0]) / F(16.0), F(k[1]) / F(16.0)) for k in ((9, 5), (7, 11), (13, 9), (5, 3), (3, 13), (1, 7), (11, 15), (15, 1))],
}
FS_TEXTURE, FS_COLOR, FS_DEBUG = 0, 1, 2  # enum FS, main.rs:23-27


def _dot(a, b):
    """Vector::dot, math/vector.rs:17-23: sum starts at 0.0 and accumulates in order."""
    s = ZERO
    for x, y in zip(a, b):
        s = s + x * y
    return s


def mat_mul(A, B):
    """Matrix * Matrix, math/matrix.rs:63-79: result[i][j] = row_i(A) . col_j(B)."""
    A = np.asarray(A, np.float32).reshape(4, 4)
    B = np.asarray(B, np.float32).reshape(4, 4)
    R = np.zeros((4, 4), np.float32)
    for i in range(4):
        for j in range(4):
            R[i, j] = _dot([A[i, k] for k in range(4)], [B[k, j] for k in range(4)])
    return R


def _clamp01(x):
    """f32::clamp(0.0, 1.0) as used by clamp_bary, rasterizer/mod.rs:102-106 (NaN stays NaN, -0.0 stays -0.0)."""
    x = np.asarray(x, np.float32)
    return np.where(x < ZERO, ZERO, np.where(x > ONE, ONE, x)).astype(np.float32)


def _as_usize(x):
    """Rust `f32 as usize`: truncation, saturating at 0 and usize::MAX, NaN -> 0."""
    x = np.asarray(x, np.float32)
    t = np.trunc(x)
    big = t >= F(1.8446744e19)
    safe = np.where(np.isnan(t) | (t <= ZERO) | big, ZERO, t)
    out = safe.astype(np.uint64)
    return np.where(big, np.uint64(0xFFFFFFFFFFFFFFFF), out)


def _as_u32(x):
    """Rust `f32 as u32` (saturating, NaN -> 0), color.rs:15-20."""
    x = np.asarray(x, np.float32)
    t = np.trunc(x)
    big = t >= F(4294967296.0)
    safe = np.where(np.isnan(t) | (t <= ZERO) | big, ZERO, t)
    return np.where(big, np.uint32(0xFFFFFFFF), safe.astype(np.uint32)).astype(np.uint32)


def to_argb(r, g, b, a):
    """Color::to_argb, color.rs:15-20: fields shifted and OR-ed without masking (u32 shifts drop the high bits)."""
    return ((_as_u32(a * F(255.0)) << np.uint32(24)) | (_as_u32(r * F(255.0)) << np.uint32(16))
            | (_as_u32(g * F(255.0)) << np.uint32(8)) | _as_u32(b * F(255.0))).astype(np.uint32)


class Texture:
    """texture.rs:8-13: bytes, width, height, texel_width (3 or 4), (0, 0) is the upper left corner."""

    def __init__(self, texels):
        t = np.ascontiguousarray(texels, np.uint8)
        self.height, self.width, self.texel_width = t.shape
        self.buf = t.reshape(-1)

    def read_texel(self, x, y):
        """Texture::read_texel + Color::from_rgba, texture.rs:47-63, color.rs:22-29 (arrays of usize coordinates).
        The reference panics on a read past the buffer; the cross-check scenes never get there (asserted)."""
        tw, w = np.uint64(self.texel_width), np.uint64(self.width)
        start = x * tw + y * tw * w
        assert int(start.max(initial=0)) + self.texel_width <= self.buf.size, "texel read past the buffer (reference panics)"
        s = start.astype(np.int64)
        ch = [self.buf[s + k].astype(np.float32) / F(255.0) for k in range(3)]
        a = self.buf[s + 3].astype(np.float32) / F(255.0) if self.texel_width == 4 else np.full(s.shape, F(255.0) / F(255.0), np.float32)
        return ch + [a]

    def sample(self, u, v):
        """Texture::sample, texture.rs:65-83: bilinear over floor/ceil texels, weights from fract()."""
        x = u * F(self.width - 1)
        y = v * F(self.height - 1)
        x0, x1 = _as_usize(np.floor(x)), _as_usize(np.ceil(x))
        y0, y1 = _as_usize(np.floor(y)), _as_usize(np.ceil(y))
        tl, tr = self.read_texel(x0, y0), self.read_texel(x1, y0)
        bl, br = self.read_texel(x0, y1), self.read_texel(x1, y1)
        xf = x - np.trunc(x)  # f32::fract
        yf = y - np.trunc(y)
        out = []
        for k in range(4):
            r0 = tl[k] * (ONE - xf) + tr[k] * xf
            r1 = bl[k] * (ONE - xf) + br[k] * xf
            out.append(r0 * (ONE - yf) + r1 * yf)
        return out


def _distance_measure(plane, p, guard=ONE):
    """clipping.rs:29-38; planes in CLIP_PLANES order LEFT, RIGHT, BOTTOM, TOP, NEAR, FAR (53-60).  guard: the four side
    planes sit at |x|, |y| <= guard * w (guard band, mod.rs:417-419; 1 = the reference)."""
    c = p[plane >> 1]
    w = guard * p[3] if plane < 4 else p[3]
    return w - c if (plane & 1) else w + c


def try_clip(verts, attrs, guard=ONE):
    """clipping::try_clip, clipping.rs:62-195.  verts: 3 x [x, y, z, w], attrs: 3 x [r, g, b, a, u, v] (np.float32
    scalars).  Returns None (Outside), "inside", or a list of (verts, attrs) fan triangles."""
    v10 = (verts[1][0] - verts[0][0], verts[1][1] - verts[0][1])
    v20 = (verts[2][0] - verts[0][0], verts[2][1] - verts[0][1])
    area2 = v10[0] * v20[1] - v20[0] * v10[1]  # triangle_2x_area on clip-space xy, mod.rs:15-21
    if abs(area2) < F(0.000001):
        return None
    inside = [[True, True] for _ in range(3)]
    outside = [[True, True] for _ in range(3)]
    for v in verts:
        for a in range(3):
            gw = guard * v[3] if a < 2 else v[3]  # "inside" against the guard band, "outside" against the frustum
            inside[a][0] &= bool(v[a] >= -gw)
            inside[a][1] &= bool(v[a] <= gw)
            outside[a][0] &= bool(v[a] < -v[3])
            outside[a][1] &= bool(v[a] > v[3])
    if any(any(x) for x in outside):
        return None
    if all(all(x) for x in inside):
        return "inside"
    out_v, out_a = [list(v) for v in verts], [list(a) for a in attrs]
    for plane in range(6):
        in_v, in_a = out_v, out_a
        out_v, out_a = [], []
        n = len(in_v)
        for i in range(n):
            pv, pa = in_v[(i + n - 1) % n], in_a[(i + n - 1) % n]
            cv, ca = in_v[i], in_a[i]
            pd, cd = _distance_measure(plane, pv, guard), _distance_measure(plane, cv, guard)
            pin, cin = bool(pd >= ZERO), bool(cd >= ZERO)
            if pin != cin:
                alpha = pd / (pd - cd)  # compute_intersection, clipping.rs:43-51
                out_v.append([(ONE - alpha) * pv[k] + alpha * cv[k] for k in range(4)])
                out_a.append([(ca[k] - pa[k]) * alpha + pa[k] for k in range(6)])
            if cin:
                out_v.append(cv)
                out_a.append(ca)
    if not out_v:
        return None
    return [([out_v[0], out_v[i + 1], out_v[i + 2]], [out_a[0], out_a[i + 1], out_a[i + 2]]) for i in range(len(out_v) - 2)]


class PyRasterizer:
    """Renderer + Rasterizer of the reference (render.rs:38-114, rasterizer/mod.rs:264-522) for the cross-check."""

    def __init__(self, width, height, msaa=4, guard_band=1.0):
        self.W, self.H = int(width), int(height)
        self.ns, self.pat, self.guard = int(msaa), PATTERNS[int(msaa)], F(guard_band)
        self.depth = np.full((self.H, self.W, self.ns), CLEAR_DEPTH, np.float32)  # DepthBuffer, buffers.rs:129-157
        self.color = np.full((self.H, self.W, self.ns), CLEAR_COLOR, np.uint32)   # ColorBuffer, buffers.rs:83-109
        self.textures = []
        self.world = self.view = self.projection = np.eye(4, dtype=np.float32)

    def bind_texture(self, index, texels):
        assert index == len(self.textures)  # uniform.rs:31
        self.textures.append(Texture(texels))

    def render(self, positions, attributes, indices, fs):
        """Renderer::render, render.rs:98-114, with the vertex shader of main.rs:147-152."""
        pos = np.asarray(positions, np.float32).reshape(-1, 3)
        att = np.asarray(attributes, np.float32).reshape(-1, 6)
        idx = np.asarray(indices).reshape(-1, 3)
        M = mat_mul(mat_mul(self.projection, self.view), self.world)  # projection * view * world, left to right
        X, Y, Z = pos[:, 0], pos[:, 1], pos[:, 2]
        clip = np.empty((pos.shape[0], 4), np.float32)
        with np.errstate(all="ignore"):
            for r in range(4):  # Matrix * Vector, math/vector.rs:219-240: row . (x, y, z, 1)
                s = ZERO + M[r, 0] * X
                s = s + M[r, 1] * Y
                s = s + M[r, 2] * Z
                clip[:, r] = s + M[r, 3] * ONE
            for tri in idx:  # primitive_assembly, render.rs:75-96; Rasterizer::rasterize, mod.rs:399-476
                verts = [[F(clip[i, k]) for k in range(4)] for i in tri]
                attrs = [[F(att[i, k]) for k in range(6)] for i in tri]
                res = try_clip(verts, attrs, self.guard)
                if res is None:
                    continue
                for v, a in ([(verts, attrs)] if res == "inside" else res):
                    self._raster_triangle(v, a, fs)

    def _raster_triangle(self, verts, attrs, fs):
        Wf, Hf = F(self.W), F(self.H)
        scr, wcam = [], []
        for v in verts:  # perspective_divide + viewport_transform, mod.rs:284-345
            nx, ny, nz = v[0] / v[3], v[1] / v[3], v[2] / v[3]
            x = Wf * (nx + ONE) / F(2.0)
            y = Hf * (ONE - (ny + ONE) / F(2.0))
            z = (nz + ONE) * HALF * (ONE - ZERO) + ZERO
            scr.append((x, y, z))
            wcam.append(v[3])
        p = [(s[0], s[1]) for s in scr]
        zs = [s[2] for s in scr]
        # RasterizerTriangle::new, mod.rs:187-222
        e = [(p[1][0] - p[0][0], p[1][1] - p[0][1]), (p[2][0] - p[1][0], p[2][1] - p[1][1]), (p[0][0] - p[2][0], p[0][1] - p[2][1])]
        n = [(-ev[1], ev[0]) for ev in e]
        v20 = (p[2][0] - p[0][0], p[2][1] - p[0][1])
        inv = ONE / (e[0][0] * v20[1] - v20[0] * e[0][1])
        # PixelBoundingBox::from + Rasterizer::bounding_box, bounding_box.rs:13-42, mod.rs:347-361
        px, py = np.array([q[0] for q in p], np.float32), np.array([q[1] for q in p], np.float32)
        mnx, mxx = np.fmin.reduce(np.concatenate(([CLEAR_DEPTH], px))), np.fmax.reduce(np.concatenate(([-CLEAR_DEPTH], px)))
        mny, mxy = np.fmin.reduce(np.concatenate(([CLEAR_DEPTH], py))), np.fmax.reduce(np.concatenate(([-CLEAR_DEPTH], py)))
        x0, x1 = max(int(_as_usize(np.floor(mnx))), 0), min(int(_as_usize(np.ceil(mxx))), self.W)
        y0, y1 = max(int(_as_usize(np.floor(mny))), 0), min(int(_as_usize(np.ceil(mxy))), self.H)
        if x0 >= x1 or y0 >= y1:
            return
        Xp = np.arange(x0, x1).astype(np.float32)[None, :]
        Yp = np.arange(y0, y1).astype(np.float32)[:, None]

        def eval_single(xs, ys):  # EdgeFunctions::eval_single, mod.rs:125-132
            return [(ZERO + n[k][0] * (xs - p[k][0])) + n[k][1] * (ys - p[k][1]) for k in range(3)]

        tie = [bool(n[k][0] > ZERO) or (not bool(n[k][0] < ZERO) and bool(n[k][1] < ZERO)) for k in range(3)]
        shape = (y1 - y0, x1 - x0)
        cov, sampled = [], []
        NS, PAT = self.ns, self.pat
        for i in range(NS):  # EdgeFunctions::eval + inside, mod.rs:134-170; RasterizerTriangle::fragment, 225-253
            ef = eval_single(Xp + PAT[i][0], Yp + PAT[i][1])
            ins = np.ones(shape, bool)
            for k in range(3):
                ins &= (ef[k] > ZERO) | (~(ef[k] < ZERO) & ~(ef[k] > ZERO) & tie[k])
            b0 = _clamp01(ef[1] * inv)
            b1 = _clamp01(ef[2] * inv)
            b2 = _clamp01(ONE - b0 - b1)
            z = b0 * zs[0] + b1 * zs[1] + b2 * zs[2]
            cov.append(ins)
            sampled.append(np.where(ins, z, ZERO).astype(np.float32))
        dview = self.depth[y0:y1, x0:x1]
        cview = self.color[y0:y1, x0:x1]
        dcov = [cov[i] & (sampled[i] < dview[..., i]) for i in range(NS)]  # depth_coverage, mod.rs:363-378 (strict <)
        shade = np.logical_or.reduce(dcov)
        if not shade.any():
            return
        # Fragment::interpolate, mod.rs:69-100: pixel centre if all samples passed, else the first passing sample
        allc = np.logical_and.reduce(dcov)
        xs = np.broadcast_to(Xp + HALF, shape).copy()
        ys = np.broadcast_to(Yp + HALF, shape).copy()
        taken = allc.copy()
        for i in range(NS):
            sel = dcov[i] & ~taken
            xs = np.where(sel, Xp + PAT[i][0], xs)
            ys = np.where(sel, Yp + PAT[i][1], ys)
            taken |= sel
        ef = eval_single(xs.astype(np.float32), ys.astype(np.float32))
        fu, fv, fw = ef[1] / wcam[0], ef[2] / wcam[1], ef[0] / wcam[2]
        s = fu + fv + fw
        u = _clamp01(fu / s)
        v = _clamp01(fv / s)
        w = _clamp01(ONE - u - v)
        att = [attrs[0][k] * u + attrs[1][k] * v + attrs[2][k] * w for k in range(6)]
        if fs == FS_TEXTURE:  # main.rs:67-77
            m = shade
            tu, tv = np.where(m, att[4], ZERO).astype(np.float32), np.where(m, att[5], ZERO).astype(np.float32)
            col = self.textures[0].sample(tu, tv)
        elif fs == FS_COLOR:
            col = att[:4]
        else:
            col = [sampled[0], sampled[0], sampled[0], np.full(shape, ONE, np.float32)]  # grayscale(frag_coords.depths[0])
        argb = to_argb(col[0], col[1], col[2], col[3])
        for i in range(NS):  # write_pixel, mod.rs:380-397
            cview[..., i] = np.where(dcov[i], argb, cview[..., i])
            dview[..., i] = np.where(dcov[i], sampled[i], dview[..., i])

    def framebuffer(self):
        """resolve_and_clear, mod.rs:478-518 + box_filter_color, buffers.rs:111-125.  Every pixel is resolved: the box
        filter of four clear samples is the clear colour, which is what untouched tiles hold."""
        c = self.color
        r = ((c & np.uint32(0x00FF0000)) >> np.uint32(16)).sum(-1, dtype=np.uint32)
        g = ((c & np.uint32(0x0000FF00)) >> np.uint32(8)).sum(-1, dtype=np.uint32)
        b = (c & np.uint32(0x000000FF)).sum(-1, dtype=np.uint32)
        n = np.uint32(self.ns)
        out = (np.uint32(0xFF) << np.uint32(24)) | ((r // n) << np.uint32(16)) | ((g // n) << np.uint32(8)) | (b // n)
        self.depth[...] = CLEAR_DEPTH
        self.color[...] = CLEAR_COLOR
        return out.astype(np.uint32)
