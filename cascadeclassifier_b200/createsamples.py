"""Host-side mirror of `opencv_createsamples -img <image> -num <n> -vec <out>` (tools/createsamples/createsamples.cpp:60-221,
tools/createsamples/utility.cpp:952-1029 cvCreateTrainingSamples) over the C ABI: the per-pixel work runs on the GPU
(cc_createsamples), the .vec file is written with the reference's layout (utility.cpp:128-153)."""
import ctypes as C

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
