import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("rag-project-icd10_b200")


@pytest.fixture(scope="session", autouse=True)
def _encoder_pdl_from_env():
    """ICD_TEST_ENC_PDL=0|1 runs the GPU suite with programmatic dependent launch forced off / on (default: the library's)."""
    v = os.environ.get("ICD_TEST_ENC_PDL")
    if v is not None and os.path.exists(os.path.join(ROOT, "rag-project-icd10_b200", "csrc", "libicdrag.so")):
        native = importlib.import_module("rag-project-icd10_b200._native")
        try:
            native.tune(enc_pdl=int(v))
        except Exception:
            pass
    yield


@pytest.fixture(scope="session")
def native(pkg):
    return importlib.import_module("rag-project-icd10_b200._native")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
