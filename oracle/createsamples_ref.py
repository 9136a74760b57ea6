"""CPU restatement of `opencv_createsamples -img <image> -num <n> -vec <out>` (training samples from a single image by random
perspective distortions), the input-preparation step in front of the hot path.

TEST INFRASTRUCTURE ONLY (checker of cc_createsamples, tests/test_gpu_createsamples.py); never imported by the product.

Follows /root/reference/tools/createsamples/utility.cpp statement by statement:
  icvStartSampleDistortion   :516-578     mask from bgcolor / bgthreshold, border extension by erode / dilate
  icvRandomQuad              :419-466     random rotation (Rodrigues) + perspective of the image rectangle
  cvGetPerspectiveTransform  :180-223     8x8 linear system
  cvWarpPerspective          :226-417     the reference's OWN scan-line warp (not cv::warpPerspective)
  icvPlaceDistortedSample    :580-672     warp image + mask, blur the mask, crop, INTER_LINEAR_EXACT resize, alpha blend
  cvCreateTrainingSamples    :952-1029    the per-sample loop (no -bg: every sample starts as a bgcolor-filled window)
  icvWriteVecHeader / Sample :128-153     the .vec layout
and tools/createsamples/createsamples.cpp:60-166 for the flag defaults and cv::setRNGSeed(rngseed).

Arithmetic that lives in OpenCV (not under /root/reference) is taken from the real library through python cv2 4.13 wherever it is
exposed -- cv2.Rodrigues, cv2.gemm, cv2.solve, cv2.erode, cv2.dilate, cv2.GaussianBlur, cv2.resize(INTER_LINEAR_EXACT) -- so that
this oracle is pinned to OpenCV itself for those steps.  cv::RNG (theRNG().uniform) is not exposed to python; it is restated from
the published inline definitions (core/operations.hpp: multiply-with-carry, state * 4164903690 + carry) and its next() stream is
pinned against cv2.randShuffle, which draws from the same theRNG() (tests/test_oracle_createsamples.py)."""
import struct

import numpy as np


class CvRNG:
    """cv::RNG (opencv2/core/operations.hpp)"""

    def __init__(self, seed):
        self.state = int(seed) & 0xFFFFFFFFFFFFFFFF if seed else 0xFFFFFFFF

    def next(self):
        s = self.state
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
        self.state = s
        return s & 0xFFFFFFFF

    def real(self):                                   # RNG::operator double()
        t = self.next()
        return float(((t << 32) | self.next())) * 5.4210108624275221700372640043497e-20

    def uniform_d(self, a, b):                        # RNG::uniform(double, double)
        return self.real() * (b - a) + a

    def uniform_i(self, a, b):                        # RNG::uniform(int, int)
        return a if a == b else int(self.next() % (b - a) + a)


def start_sample_distortion(src, bgcolor, bgthreshold):
    """icvStartSampleDistortion: returns (src with extended borders, mask, dx, dy)"""
    import cv2
    src = np.ascontiguousarray(src, np.uint8).copy()
    m = src.astype(np.int32)
    mask = np.where((bgcolor - bgthreshold <= m) & (m <= bgcolor + bgthreshold), 0, 255).astype(np.uint8)
    er = cv2.erode(src, None)
    di = cv2.dilate(src, None)
    de = ((bgcolor - er.astype(np.int32)) & 0xFF)     # (uchar)(bgcolor - perode)
    dd = ((di.astype(np.int32) - bgcolor) & 0xFF)
    out = src.copy()
    sel = mask == 0
    a = sel & (de >= dd) & (de > bgthreshold)
    out[a] = er[a]
    b = sel & (dd > de) & (dd > bgthreshold)          # evaluated after the first assignment in the reference, on the same de / dd
    out[b] = di[b]
    return out, mask, src.shape[1] // 2, src.shape[0] // 2


def random_quad(rng, width, height, maxxangle, maxyangle, maxzangle):
    import cv2
    distfactor, distfactor2 = 3.0, 1.0
    rv = np.zeros((3, 1), np.float64)
    rv[0, 0] = rng.uniform_d(-maxxangle, maxxangle)
    rv[1, 0] = (maxyangle - abs(rv[0, 0])) * rng.uniform_d(-1.0, 1.0)
    rv[2, 0] = rng.uniform_d(-maxzangle, maxzangle)
    d = (distfactor + distfactor2 * rng.uniform_d(-1.0, 1.0)) * width
    rot, _ = cv2.Rodrigues(rv)
    halfw, halfh = 0.5 * width, 0.5 * height
    quad = np.array([[-halfw, -halfh], [halfw, -halfh], [halfw, halfh], [-halfw, halfh]], np.float64)
    for i in range(4):
        vec = cv2.gemm(rot, np.array([[quad[i, 0]], [quad[i, 1]], [0.0]], np.float64), 1.0, None, 1.0)
        quad[i, 0] = vec[0, 0] * d / (d + vec[2, 0]) + halfw
        quad[i, 1] = vec[1, 0] * d / (d + vec[2, 0]) + halfh
    return quad, rot, d


def perspective_transform(src_w, src_h, quad):
    import cv2
    a = np.zeros((8, 8), np.float64)
    b = np.zeros((8, 1), np.float64)
    for i in range(4):
        a[i, 0], a[i, 1], a[i, 2] = quad[i, 0], quad[i, 1], 1
    for i in range(4, 8):
        a[i, 3], a[i, 4], a[i, 5] = quad[i - 4, 0], quad[i - 4, 1], 1
    u, v = src_w - 1, src_h - 1
    a[1, 6], a[1, 7] = -quad[1, 0] * u, -quad[1, 1] * u
    a[2, 6], a[2, 7] = -quad[2, 0] * u, -quad[2, 1] * u
    b[1, 0] = b[2, 0] = u
    a[6, 6], a[6, 7] = -quad[2, 0] * v, -quad[2, 1] * v
    a[7, 6], a[7, 7] = -quad[3, 0] * v, -quad[3, 1] * v
    b[6, 0] = b[7, 0] = v
    ok, x = cv2.solve(a, b)
    c = np.ones(9, np.float64)
    c[:8] = x[:, 0]
    return c.reshape(3, 3)


def warp_rows(quad, dst_w, dst_h):
    """The scan-line walk of cvWarpPerspective: [(y, ix_min, ix_max)] for every destination row it writes, in its order."""
    direction = 0
    for i in range(4):
        ni, pi = (i + 1) % 4, (i - 1) % 4
        d = (quad[i][0] - quad[pi][0]) * (quad[ni][1] - quad[i][1]) - (quad[i][1] - quad[pi][1]) * (quad[ni][0] - quad[i][0])
        cur = 1 if d > 0 else (-1 if d < 0 else 0)
        if direction == 0:
            direction = cur
        elif direction * cur < 0:
            direction = 0
            break
    if direction == 0:
        raise ValueError("Quadrangle is nonconvex or degenerated.")
    left = 0
    for i in range(1, 4):
        if quad[i][1] < quad[left][1] or (quad[i][1] == quad[left][1] and quad[i][0] < quad[left][0]):
            left = i
    q = [None] * 4
    if direction > 0:
        for i in range(left, 4):
            q[i - left] = (quad[i][0], quad[i][1])
        for i in range(left):
            q[4 - left + i] = (quad[i][0], quad[i][1])
    else:
        for i in range(left, -1, -1):
            q[left - i] = (quad[i][0], quad[i][1])
        for i in range(3, left, -1):
            q[4 + left - i] = (quad[i][0], quad[i][1])
    left = right = 0
    if q[left][1] == q[left + 1][1]:
        right = 1
    next_left, next_right = 3, right + 1
    y_min = q[left][1] - 1

    def edge(p, n):
        k = (q[p][0] - q[n][0]) / (q[p][1] - q[n][1])
        bb = (q[p][1] * q[n][0] - q[p][0] * q[n][1]) / (q[p][1] - q[n][1])
        return k, bb
    k_left, b_left = edge(left, next_left)
    k_right, b_right = edge(right, next_right)
    rows = []
    while True:
        y_max = min(q[next_left][1], q[next_right][1])
        iy_min = max(int(np.rint(y_min)), 0) + 1
        iy_max = min(int(np.rint(y_max)), dst_h - 1)
        x_min = k_left * iy_min + b_left
        x_max = k_right * iy_min + b_right
        for y in range(iy_min, iy_max + 1):
            rows.append((y, max(int(np.rint(x_min)), 0), min(int(np.rint(x_max)), dst_w - 1)))
            x_min += k_left
            x_max += k_right
        if next_left == next_right or (next_left + 1 == next_right and q[next_left][1] == q[next_right][1]):
            break
        if y_max == q[next_left][1]:
            left = next_left
            next_left = left - 1
            k_left, b_left = edge(left, next_left)
        if y_max == q[next_right][1]:
            right = next_right
            next_right = right + 1
            k_right, b_right = edge(right, next_right)
        y_min = y_max
    return rows


def warp_perspective(src, dst, quad):
    """cvWarpPerspective(src, dst, quad): writes the pixels of dst inside the quadrangle (bilinear, fill value 0 outside src)"""
    c = perspective_transform(src.shape[1], src.shape[0], quad)
    sh, sw = src.shape
    s = src.astype(np.float64)
    for y, x0, x1 in warp_rows(quad, dst.shape[1], dst.shape[0]):
        if x1 < x0:
            continue
        x = np.arange(x0, x1 + 1, dtype=np.float64)
        div = c[2, 0] * x + c[2, 1] * y + c[2, 2]
        sx = (c[0, 0] * x + c[0, 1] * y + c[0, 2]) / div
        sy = (c[1, 0] * x + c[1, 1] * y + c[1, 2]) / div
        ix = np.floor(sx).astype(np.int64)
        iy = np.floor(sy).astype(np.int64)
        fx, fy = sx - ix, sy - iy

        def px(yy, xx, cond):
            v = np.zeros(len(x), np.float64)
            ok = cond & (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
            v[ok] = s[yy[ok], xx[ok]]
            return v
        i00 = px(iy, ix, (ix >= 0) & (ix < sw) & (iy >= 0) & (iy < sh))
        i10 = px(iy, ix + 1, (ix >= -1) & (ix + 1 < sw) & (iy >= 0) & (iy < sh))
        i01 = px(iy + 1, ix, (ix >= 0) & (ix < sw) & (iy >= -1) & (iy + 1 < sh))
        i11 = px(iy + 1, ix + 1, (ix >= -1) & (ix + 1 < sw) & (iy >= -1) & (iy + 1 < sh))
        i0 = i00 + (i10 - i00) * fx
        i1 = i01 + (i11 - i01) * fx
        dst[y, x0:x1 + 1] = (i0 + (i1 - i0) * fy).astype(np.uint8)      # (uchar)double: truncation; the value is in [0, 255]
    return dst


def place_distorted_sample(background, inverse, maxintensitydev, maxxangle, maxyangle, maxzangle, data, rng):
    """icvPlaceDistortedSample with inscribe = 0, maxshiftf = 0, maxscalef = 0 (what cvCreateTrainingSamples passes)"""
    import cv2
    src, mask, dx, dy, bgcolor = data
    quad, _, _ = random_quad(rng, src.shape[1], src.shape[0], maxxangle, maxyangle, maxzangle)
    quad = quad + np.array([float(dx), float(dy)])
    img = np.full((src.shape[0] + 2 * dy, src.shape[1] + 2 * dx), bgcolor, np.uint8)
    maskimg = np.zeros_like(img)
    warp_perspective(src, img, quad)
    warp_perspective(mask, maskimg, quad)
    maskimg = cv2.GaussianBlur(maskimg, (3, 3), 0, 0)
    crx, cry, crw, crh = dx, dy, src.shape[1], src.shape[0]
    maxshiftf = maxscalef = 0.0
    xshift = rng.uniform_d(0., maxshiftf)
    yshift = rng.uniform_d(0., maxshiftf)
    crx -= int(xshift * crw)
    cry -= int(yshift * crh)
    crw = int((1.0 + maxshiftf) * crw)
    crh = int((1.0 + maxshiftf) * crh)
    randscale = rng.uniform_d(0., maxscalef)
    crx -= int(0.5 * randscale * crw)
    cry -= int(0.5 * randscale * crh)
    crw = int((1.0 + randscale) * crw)
    crh = int((1.0 + randscale) * crh)
    f = np.float32
    bh, bw = background.shape
    scale = max(f(crw) / f(bw), f(crh) / f(bh))
    rx = int(f(-0.5) * (scale * f(bw) - f(crw)) + f(crx))
    ry = int(f(-0.5) * (scale * f(bh) - f(crh)) + f(cry))
    rw = int(scale * f(bw))
    rh = int(scale * f(bh))
    x0, y0 = max(rx, 0), max(ry, 0)
    x1, y1 = min(rx + rw, img.shape[1]), min(ry + rh, img.shape[0])          # roi & Rect(0, 0, img.size())
    rimg = cv2.resize(np.ascontiguousarray(img[y0:y1, x0:x1]), (bw, bh), interpolation=cv2.INTER_LINEAR_EXACT)
    rmask = cv2.resize(np.ascontiguousarray(maskimg[y0:y1, x0:x1]), (bw, bh), interpolation=cv2.INTER_LINEAR_EXACT)
    forecolordev = rng.uniform_i(-maxintensitydev, maxintensitydev)
    chart = np.clip(forecolordev + rimg.astype(np.int32), 0, 255)
    if inverse:
        chart = chart ^ 0xFF
    alpha = rmask.astype(np.int32)
    out = (chart * alpha + (255 - alpha) * background.astype(np.int32)) // 255
    return out.astype(np.uint8), dict(quad=quad, roi=(x0, y0, x1 - x0, y1 - y0), forecolordev=forecolordev, img=img, maskimg=maskimg)


CV_RANDOM_INVERT = 0x7FFFFFFF


def create_training_samples(src_gray, count, winw=24, winh=24, bgcolor=0, bgthreshold=80, invert=0, maxintensitydev=40,
                            maxxangle=1.1, maxyangle=1.1, maxzangle=0.5, rngseed=12345, debug=None):
    """cvCreateTrainingSamples without -bg; returns uint8 [count][winh][winw] (the payload of the .vec file)"""
    rng = CvRNG(rngseed)
    src, mask, dx, dy = start_sample_distortion(src_gray, bgcolor, bgthreshold)
    data = (src, mask, dx, dy, bgcolor)
    out = np.zeros((count, winh, winw), np.uint8)
    inverse = invert
    for i in range(count):
        sample = np.full((winh, winw), bgcolor, np.uint8)
        if invert == CV_RANDOM_INVERT:
            inverse = rng.uniform_i(0, 2)
        out[i], dbg = place_distorted_sample(sample, inverse, maxintensitydev, maxxangle, maxyangle, maxzangle, data, rng)
        if debug is not None:
            debug.append(dbg)
    return out


def write_vec(path, samples):
    """icvWriteVecHeader + icvWriteVecSample"""
    n, h, w = samples.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<iihh", n, w * h, 0, 0))
        for i in range(n):
            f.write(b"\x00")
            f.write(samples[i].astype("<i2").tobytes())
