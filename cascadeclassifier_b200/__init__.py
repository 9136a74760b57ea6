"""cascadeclassifier_b200 — B200-native stage training and hard-negative mining for opencv_traincascade.

The compute path is libcascade_cuda.so (hand-written sm_100a CUDA behind the C ABI in
include/cascade_cuda.h).  This package only holds the host-side mirror of the reference's
CvFeatureEvaluator / CvCascadeBoost / CvCascadeClassifier interfaces and data-format helpers.
"""
from ._lib import CascadeCudaError, LIB_PATH, exported_symbols  # noqa: F401
from .evaluator import (Context, CvFeatureEvaluator, CvFeatureParams, CvHaarFeatureParams,  # noqa: F401
                        CvHOGFeatureParams, CvLBPFeatureParams)
from .boost import CvCascadeBoost, CvCascadeBoostParams  # noqa: F401

__version__ = "0.1.0"
