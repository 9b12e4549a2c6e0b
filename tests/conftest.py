import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

PKG_NAME = "sequence-alignment-tools_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (the hyphen in its name rules out a plain import)."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def built_lib(pkg):
    """Builds lib/libpm_b200.so when it is missing or stale (nvcc, no GPU needed)."""
    build = importlib.import_module(PKG_NAME + ".build")
    build.build()
    return pkg.lib()


@pytest.fixture(scope="session")
def oracle():
    import pm_oracle
    pm_oracle.build()
    return pm_oracle


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "primer_match_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def halves_golden():
    with open(os.path.join(ROOT, "tests", "golden", "exact_halves_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def constraints_golden():
    with open(os.path.join(ROOT, "tests", "golden", "constraints_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def pcr_golden():
    with open(os.path.join(ROOT, "tests", "golden", "pcr_match_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def pcr_ties_golden():
    with open(os.path.join(ROOT, "tests", "golden", "pcr_ties_golden.json")) as f:
        return json.load(f)


def records_of(case):
    return [tuple(r) for r in case["records"]]


def as_tuples(recs):
    return [(r.seq_index, r.seq_start, r.seq_end, r.strand, r.pattern_index, r.editdist,
             r.stream_start, r.stream_end, r.alignment) for r in recs]
