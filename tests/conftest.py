import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PKG_NAME = "lightweight-face-detection-centernet_b200"
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    m = importlib.import_module(PKG_NAME)
    m.build()
    return m


@pytest.fixture(scope="session")
def oracle():
    import centerface_oracle
    return centerface_oracle


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(GOLD, "golden_v1.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def weights_path():
    return os.path.join(GOLD, "weights_e100.npz")


@pytest.fixture(scope="session")
def sd(oracle, weights_path):
    return oracle.load_weights(weights_path)


@pytest.fixture(scope="session")
def images():
    """The five bundled JPEGs (BGR u8), decoded from the committed byte fixtures."""
    import cv2
    z = np.load(os.path.join(GOLD, "images_jpeg.npz"))
    return {k[4:]: cv2.imdecode(z[k], cv2.IMREAD_COLOR) for k in z.files}


IMGS = ["1", "17", "2", "27", "8"]


@pytest.fixture(scope="session")
def f5_640(images):
    """SURVEY.md 8d parity inputs: each JPEG stretched to 640x640 (u8 BGR HWC)."""
    import cv2
    return {n: cv2.resize(images[n], (640, 640)) for n in IMGS}


@pytest.fixture(scope="session")
def oracle_heads_640(oracle, sd, f5_640):
    """Oracle forward on the F5 set (CPU, ~0.2 s per image), cached for the session."""
    import torch
    out = {}
    for n in IMGS:
        x = torch.from_numpy(oracle.normalize_u8(f5_640[n])).unsqueeze(0)
        out[n] = oracle.forward(sd, x)
    return out
