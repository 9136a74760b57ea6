import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def make_samples(n_pos, n_neg, W=24, H=24, seed=7):
    """Seeded training set: distorted glyph positives + windows cut from generated backgrounds."""
    from cascadeclassifier_b200 import synth
    pos = synth.make_positives(n_pos, W, H, seed=seed)
    rng = np.random.default_rng(seed + 1)
    bgs = synth.make_backgrounds(4, 320, 240, seed0=seed * 100)
    neg = np.empty((n_neg, H, W), np.uint8)
    for i in range(n_neg):
        b = bgs[i % len(bgs)]
        y = int(rng.integers(0, b.shape[0] - H))
        x = int(rng.integers(0, b.shape[1] - W))
        neg[i] = b[y:y + H, x:x + W]
    imgs = np.concatenate([pos, neg])
    labels = np.concatenate([np.ones(n_pos, np.uint8), np.zeros(n_neg, np.uint8)])
    return imgs, labels


def pytest_collection_modifyitems(config, items):
    """Everything under tests/test_gpu_*.py needs a CUDA device: make sure it carries the `gpu` marker (so `-m "not gpu"` never
    collects it), and when no device is visible skip those tests instead of reporting errors."""
    gpu_mark = pytest.mark.gpu
    have_gpu = None
    for item in items:
        if os.path.basename(str(item.fspath)).startswith("test_gpu_") and "gpu" not in item.keywords:
            item.add_marker(gpu_mark)
        if "gpu" in item.keywords:
            if have_gpu is None:
                have_gpu = _cuda_device_present()
            if not have_gpu:
                item.add_marker(pytest.mark.skip(reason="no CUDA device (run on the B200 box with -m gpu)"))


def _cuda_device_present():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return os.path.exists("/dev/nvidia0")


@pytest.fixture(scope="session")
def gpu_ctx():
    import cascadeclassifier_b200 as ccb
    try:
        return ccb.Context(0)
    except ccb.CascadeCudaError as e:      # the library itself has no CPU fallback; only the TEST is skipped
        pytest.skip("no CUDA device: %s" % e)
