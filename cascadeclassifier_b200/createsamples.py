"""Host-side mirror of `opencv_createsamples -img <image> -num <n> -vec <out>` (tools/createsamples/createsamples.cpp:60-221,
tools/createsamples/utility.cpp:952-1029 cvCreateTrainingSamples) over the C ABI: the per-pixel work runs on the GPU
(cc_createsamples), the .vec file is written with the reference's layout (utility.cpp:128-153).
`-info <collection> -vec <out>` (cvCreateTrainingSamplesFromInfo, utility.cpp:1125-1230) is file glue around cv::imread / cv::resize
-- one small resize per annotated object, no arithmetic of the tool's own -- and is mirrored on the host with OpenCV's python
binding doing exactly those two calls."""
import ctypes as C
import os
import struct

import numpy as np

from . import synth
from ._lib import CascadeCudaError, check

CV_RANDOM_INVERT = 0x7FFFFFFF


def cvCreateTrainingSamples(ctx, filename, image, bgcolor=0, bgthreshold=80, bgfilename=None, count=1000, invert=0,
                            maxintensitydev=40, maxxangle=1.1, maxyangle=1.1, maxzangle=0.5, showsamples=0, winwidth=24, winheight=24,
                            rngseed=12345):
    """Same argument order and defaults as the reference (createsamples.cpp:74-88).  `image` = path of the source image or a gray
    uint8 array.  Returns the samples [count][winheight][winwidth]; writes `filename` unless it is None."""
    if bgfilename is not None:
        raise CascadeCudaError("cvCreateTrainingSamples: -bg (samples pasted onto background images) is not implemented on the GPU path")
    if isinstance(image, str):
        import cv2
        img = cv2.imread(image, cv2.IMREAD_GRAYSCALE)
        if img is None:
            return None                                  # icvStartSampleDistortion fails silently, like the reference
    else:
        img = image
    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim != 2:
        raise CascadeCudaError("Source must be two-dimensional array of CV_8UC1 type.")
    out = np.zeros((int(count), int(winheight), int(winwidth)), np.uint8)
    check(ctx.lib.cc_createsamples(ctx.h, img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], int(bgcolor), int(bgthreshold), int(count),
                                   int(invert), int(maxintensitydev), float(maxxangle), float(maxyangle), float(maxzangle), int(winwidth),
                                   int(winheight), int(rngseed), out.ctypes.data_as(C.c_void_p)), "cc_createsamples")
    if filename is not None:
        synth.write_vec(filename, out)
    return out


def cvCreateTrainingSamplesFromInfo(infoname, vecfilename, num, showsamples=0, winwidth=24, winheight=24):
    """`opencv_createsamples -info <infoname> -vec <vecfilename> -num <num> -w -h` (utility.cpp:1125-1230): every line of the
    collection file is `<image> <count> x y w h [x y w h ...]` with the image path relative to the collection file; each rectangle
    is cut out of the gray image and resized to the window -- INTER_AREA when it shrinks in both directions, INTER_LINEAR_EXACT
    otherwise (:1205-1208) -- and appended to the .vec.  The header carries `num` as the reference writes it up front (:1172), also
    when the collection holds fewer objects.  Returns the number of samples written.  Differences: a parse error ends the file
    like the reference (:1217-1223); at the end of the collection the reference's loop would spin without reading anything
    (`count` stays 0, `total < num` stays true), here it ends."""
    import cv2
    if infoname is None or vecfilename is None:
        raise ValueError("infoname / vecfilename is NULL")
    d = os.path.dirname(vecfilename)
    try:
        if d:
            os.makedirs(d, exist_ok=True)
        with open(infoname, "r") as f:
            tokens = f.read().split()
        vec = open(vecfilename, "wb")
    except OSError:
        return 0
    base = os.path.dirname(infoname)
    total, pos = 0, 0
    with vec:
        vec.write(struct.pack("<iihh", int(num), int(winwidth) * int(winheight), 0, 0))
        while total < num and pos + 2 <= len(tokens):
            name = tokens[pos]
            try:
                count = int(tokens[pos + 1])
            except ValueError:
                break                                     # fscanf("%s %d") != 2: nothing is read from this line
            pos += 2
            src = cv2.imread(os.path.join(base, name), cv2.IMREAD_GRAYSCALE)
            error = False
            for _ in range(count):
                if total >= num:
                    break
                try:
                    x, y, w, h = (int(t) for t in tokens[pos:pos + 4])
                    if pos + 4 > len(tokens):
                        raise ValueError
                except ValueError:
                    error = True
                    break
                pos += 4
                if src is None or x < 0 or y < 0 or w <= 0 or h <= 0 or x + w > src.shape[1] or y + h > src.shape[0]:
                    raise ValueError("%s: rectangle (%d, %d, %d, %d) outside the image" % (name, x, y, w, h))      # cv::Mat::operator() asserts
                interp = cv2.INTER_AREA if (w >= winwidth and h >= winheight) else cv2.INTER_LINEAR_EXACT
                sample = cv2.resize(src[y:y + h, x:x + w], (int(winwidth), int(winheight)), interpolation=interp)
                vec.write(b"\x00")
                vec.write(np.ascontiguousarray(sample, np.uint8).astype("<i2").tobytes())
                total += 1
            if error:
                break
    return total
