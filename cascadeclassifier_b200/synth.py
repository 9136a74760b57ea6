"""Synthetic positives (.vec layout) and background images for benchmarks and parity tests.

Stands in for `opencv_createsamples` on a generated pattern: the reference tool
needs OpenCV C++, which is not in this image, so the positives are rendered here and stored in
the identical .vec layout (tools/createsamples/utility.cpp:128-153, reader
traincascade/lib/src/imagestorage.cpp:138-182).  Everything is seeded (numpy PCG64) so that the
CPU oracle and the CUDA path see byte-identical inputs on any box.
"""
import os
import struct

import numpy as np


def _pattern(size=96):
    """High-contrast glyph on a flat background (a crude face: brow band, eyes, nose, mouth)."""
    p = np.full((size, size), 170.0, np.float32)
    s = size / 24.0
    def box(x0, y0, x1, y1, v):
        p[int(y0 * s):int(y1 * s), int(x0 * s):int(x1 * s)] = v
    box(3, 5, 21, 8, 60)      # brow band
    box(5, 8, 9, 11, 30)      # left eye
    box(15, 8, 19, 11, 30)    # right eye
    box(10, 8, 14, 16, 215)   # nose ridge
    box(7, 17, 17, 20, 70)    # mouth
    box(3, 11, 6, 16, 200)    # cheeks
    box(18, 11, 21, 16, 200)
    return p


def _warp(pat, W, H, rng):
    """Random affine jitter + contrast/brightness + noise; bilinear sampling with numpy only."""
    size = pat.shape[0]
    ang = rng.uniform(-0.18, 0.18)
    sc = rng.uniform(0.85, 1.15)
    tx, ty = rng.uniform(-0.08, 0.08, 2) * size
    ca, sa = np.cos(ang) / sc, np.sin(ang) / sc
    ys, xs = np.meshgrid((np.arange(H) + 0.5) / H * size, (np.arange(W) + 0.5) / W * size, indexing="ij")
    cx = cy = size / 2.0
    u = ca * (xs - cx) - sa * (ys - cy) + cx + tx
    v = sa * (xs - cx) + ca * (ys - cy) + cy + ty
    u = np.clip(u, 0, size - 1.001)
    v = np.clip(v, 0, size - 1.001)
    u0, v0 = u.astype(np.int32), v.astype(np.int32)
    fu, fv = u - u0, v - v0
    img = (pat[v0, u0] * (1 - fu) * (1 - fv) + pat[v0, u0 + 1] * fu * (1 - fv) +
           pat[v0 + 1, u0] * (1 - fu) * fv + pat[v0 + 1, u0 + 1] * fu * fv)
    gain = rng.uniform(0.3, 1.3)
    bias = rng.uniform(-50, 50)
    img = (img - 128.0) * gain + 128.0 + bias + rng.normal(0, rng.uniform(3, 40), img.shape)
    if rng.random() < 0.5:  # illumination gradient
        img += np.linspace(-1, 1, W)[None, :] * rng.uniform(-35, 35)
    if rng.random() < 0.35:  # occluding patch
        ph, pw = int(rng.integers(H // 5, H // 2)), int(rng.integers(W // 5, W // 2))
        py, px = int(rng.integers(0, H - ph)), int(rng.integers(0, W - pw))
        img[py:py + ph, px:px + pw] = rng.uniform(40, 220)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_positives(count, W, H, seed=12345):
    rng = np.random.default_rng(seed)
    pat = _pattern(96)
    if W != H:  # stretch the glyph to the window aspect (e.g. 48x96 pedestrian-shaped windows)
        pat = np.repeat(pat, 1, axis=0)
    out = np.empty((count, H, W), np.uint8)
    for i in range(count):
        out[i] = _warp(pat, W, H, rng)
    return out


def _distractor(size, rng):
    """A glyph-like patch that is NOT the positive pattern: scrambled / flipped / inverted variants."""
    pat = _pattern(48)
    kind = int(rng.integers(0, 8))
    if kind >= 5:                                # near miss: the glyph with half of it flattened
        pat = pat.copy()
        half = int(rng.integers(0, 4))
        if half == 0:
            pat[:24, :] = pat[:24, :].mean()
        elif half == 1:
            pat[24:, :] = pat[24:, :].mean()
        elif half == 2:
            pat[:, :24] = pat[:, :24].mean()
        else:
            pat[:, 24:] = pat[:, 24:].mean()
    elif kind == 0:
        pat = pat[::-1, :]                       # upside down
    elif kind == 1:
        pat = 255.0 - pat                        # inverted contrast
    elif kind == 2:
        q = 24
        blocks = [pat[i:i + q, j:j + q].copy() for i in (0, q) for j in (0, q)]
        order = rng.permutation(4)
        pat = np.block([[blocks[order[0]], blocks[order[1]]], [blocks[order[2]], blocks[order[3]]]])
    elif kind == 3:
        pat = np.roll(pat, int(rng.integers(8, 40)), axis=int(rng.integers(0, 2)))
    else:
        pat = pat.T.copy()                       # transposed
    idx = (np.arange(size) * (48.0 / size)).astype(np.int32)
    patch = pat[idx][:, idx]
    gain = rng.uniform(0.5, 1.3)
    return (patch - 128.0) * gain + 128.0 + rng.uniform(-40, 40)


def make_background(w, h, seed, clutter=0.0):
    """8-bit gray: multi-octave value noise + random rectangles / gradients (non-zero window variance).
    clutter > 0 pastes that many glyph-like distractors per 10^4 pixels (hard negatives)."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float32)
    amp = 1.0
    for octave in range(6):
        gh, gw = max(2, h >> (7 - octave)), max(2, w >> (7 - octave))
        g = rng.random((gh + 1, gw + 1)).astype(np.float32)
        ys = np.linspace(0, gh, h, endpoint=False)
        xs = np.linspace(0, gw, w, endpoint=False)
        y0, x0 = ys.astype(np.int32), xs.astype(np.int32)
        fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
        a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
        img += amp * (a * (1 - fx) * (1 - fy) + b * fx * (1 - fy) + c * (1 - fx) * fy + d * fx * fy)
        amp *= 0.6
    img = (img - img.min()) / max(1e-6, img.max() - img.min()) * 255.0
    for _ in range(int(rng.integers(8, 24))):
        x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8))
        x1, y1 = min(w, x0 + int(rng.integers(8, max(9, w // 3)))), min(h, y0 + int(rng.integers(8, max(9, h // 3))))
        kind = rng.random()
        if kind < 0.5:
            img[y0:y1, x0:x1] = img[y0:y1, x0:x1] * 0.3 + rng.uniform(0, 255) * 0.7
        else:
            img[y0:y1, x0:x1] += np.linspace(-60, 60, x1 - x0)[None, :] * rng.uniform(-1, 1)
    if clutter > 0:
        for _ in range(int(clutter * w * h / 1e4)):
            size = int(rng.integers(20, max(21, min(w, h) // 2)))
            if size >= min(w, h):
                continue
            x0, y0 = int(rng.integers(0, w - size)), int(rng.integers(0, h - size))
            img[y0:y0 + size, x0:x0 + size] = _distractor(size, rng)
    img += rng.normal(0, 3.0, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_backgrounds(count, w, h, seed0=1000, clutter=0.0):
    return [make_background(w, h, seed0 + i, clutter) for i in range(count)]


def make_backgrounds_device(count, w, h, seed, device, chunk=64):
    """`count` distinct 8-bit gray backgrounds generated ON the GPU (plumbing for benchmarks at BASELINE configs[4] size: 4096
    images of 1920x1080 would take the numpy generator above the better part of an hour).  Multi-octave value noise (bilinear
    upsampling of seeded random grids), random flat / ramp rectangles and per-pixel noise, so that window variance is never zero.
    Returns a uint8 torch tensor [count, h, w] on `device`; deterministic for (count, w, h, seed) on a given GPU model."""
    import torch
    import torch.nn.functional as Fn
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    out = torch.empty((count, h, w), dtype=torch.uint8, device=device)
    ys = torch.arange(h, device=device, dtype=torch.float32)[None, :, None]
    xs = torch.arange(w, device=device, dtype=torch.float32)[None, None, :]
    for c0 in range(0, count, chunk):
        n = min(chunk, count - c0)
        img = torch.zeros((n, 1, h, w), device=device)
        amp = 1.0
        for octave in range(5):
            gh, gw = max(2, h >> (7 - octave)), max(2, w >> (7 - octave))
            grid = torch.rand((n, 1, gh + 1, gw + 1), generator=g, device=device)
            img += amp * Fn.interpolate(grid, size=(h, w), mode="bilinear", align_corners=True)
            amp *= 0.6
        img = img[:, 0]
        lo = img.amin(dim=(1, 2), keepdim=True)
        hi = img.amax(dim=(1, 2), keepdim=True)
        img = (img - lo) / (hi - lo).clamp_min(1e-6) * 255.0
        for _ in range(12):
            r = torch.rand((n, 6), generator=g, device=device)
            x0 = (r[:, 0] * (w - 8))[:, None, None]
            y0 = (r[:, 1] * (h - 8))[:, None, None]
            x1 = x0 + 8 + r[:, 2][:, None, None] * (w / 3)
            y1 = y0 + 8 + r[:, 3][:, None, None] * (h / 3)
            inside = (xs >= x0) & (xs < x1) & (ys >= y0) & (ys < y1)
            level = (r[:, 4] * 255.0)[:, None, None]
            ramp = (xs - x0) / (x1 - x0) * 120.0 - 60.0
            flat = img * 0.3 + level * 0.7
            grad = img + ramp * (r[:, 5][:, None, None] * 2 - 1)
            img = torch.where(inside, torch.where((r[:, 5] < 0.5)[:, None, None], flat, grad), img)
        img += torch.randn((n, h, w), generator=g, device=device) * 3.0
        out[c0:c0 + n] = img.round().clamp(0, 255).to(torch.uint8)
    return out


# ---- on-disk formats -------------------------------------------------------------------------
def write_vec(path, samples):
    """.vec: int32 count, int32 vecSize, int16 0, int16 0; per sample 1 pad byte + vecSize int16."""
    n, H, W = samples.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<iihh", n, W * H, 0, 0))
        for i in range(n):
            f.write(b"\x00")
            f.write(samples[i].astype("<i2").tobytes())


def read_vec(path, W, H):
    with open(path, "rb") as f:
        count, vec_size, _, _ = struct.unpack("<iihh", f.read(12))
        if vec_size != W * H:
            raise ValueError("vec sample size %d does not match %dx%d" % (vec_size, W, H))
        raw = np.frombuffer(f.read(), np.uint8)
    rec = 1 + 2 * vec_size
    n = min(count, raw.size // rec)
    body = raw[: n * rec].reshape(n, rec)[:, 1:]
    vals = body.view("<i2").reshape(n, vec_size)
    return vals.astype(np.uint8).reshape(n, H, W)


def write_pgm(path, img):
    h, w = img.shape
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def read_pgm(path):
    with open(path, "rb") as f:
        data = f.read()
    tokens, pos = [], 0
    while len(tokens) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while data[pos:pos + 1] != b"\n":
                pos += 1
            continue
        start = pos
        while not data[pos:pos + 1].isspace():
            pos += 1
        tokens.append(data[start:pos])
    pos += 1
    if tokens[0] != b"P5":
        raise ValueError("only binary PGM (P5) is supported: %s" % path)
    w, h, _ = int(tokens[1]), int(tokens[2]), int(tokens[3])
    return np.frombuffer(data, np.uint8, w * h, pos).reshape(h, w).copy()


def write_dataset(directory, positives, backgrounds):
    os.makedirs(directory, exist_ok=True)
    write_vec(os.path.join(directory, "pos.vec"), positives)
    names = []
    for i, bg in enumerate(backgrounds):
        name = os.path.join(directory, "bg_%05d.pgm" % i)
        write_pgm(name, bg)
        names.append(name)
    with open(os.path.join(directory, "bg.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    return os.path.join(directory, "pos.vec"), os.path.join(directory, "bg.txt")
