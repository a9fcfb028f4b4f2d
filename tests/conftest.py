import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if "hist_f32" not in p and "label_lut" not in p and not os.path.basename(p).startswith(("focal_", "crit_")))


def focal_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "focal_*.npz")))


def criterion_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "crit_*.npz")))


def load_criterion(name):
    d = load_focal(name)
    d["class_w"] = d["class_w"] if d["class_w"].size else None
    return d


def load_focal(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["logits_dtype"] = str(d["logits_dtype"])
    return d


def load_golden(name):
    """Golden case -> dict; logits come back as (array, oracle dtype code)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["name"] = name
    d["logits_dtype"] = str(d["logits_dtype"])
    d["loss_kind"] = str(d["loss_kind"])
    d["class_w"] = d["class_w"] if d["class_w"].size else None
    d["grad"] = d["grad"] if d["grad"].size else None
    return d


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(autouse=True)
def _fresh_branch_hint():
    """The OHEM branch hint (functional._BranchHint) is process-wide state: without this a test's route would depend on
    which test ran before it."""
    try:
        from segmentation_factory_b200 import functional as fn
        fn._BranchHint._by_device.clear()
    except Exception:
        pass
    yield
