import gzip
import json
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "MANIFEST.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def tiny_dir(tmp_path_factory):
    """The committed tiny indexes (gzip'ed *.cfr) unpacked next to their reads."""
    d = tmp_path_factory.mktemp("tiny")
    src = os.path.join(GOLDEN, "tiny")
    for f in os.listdir(src):
        p = os.path.join(src, f)
        if f.endswith(".cfr.gz"):
            with gzip.open(p, "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
        elif f.endswith(".fq"):
            shutil.copyfile(p, os.path.join(d, f))
    return str(d)


@pytest.fixture(scope="session")
def example_idx():
    import make_data
    d = make_data.ensure("example", log=lambda *a: None)
    if d is None:
        pytest.skip("data/example index not available (needs oracle/_ref + /root/reference/example)")
    return os.path.join(d, "cfr_ref_idx")


@pytest.fixture(scope="session")
def small_dir():
    import make_data
    d = make_data.ensure("small", log=lambda *a: None)
    if d is None:
        pytest.skip("data/small not available (needs oracle/_ref/centrifuger-build)")
    return d


def golden_path(*parts):
    return os.path.join(GOLDEN, *parts)
