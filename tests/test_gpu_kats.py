"""The reference's known-answer tests (tests/golden/reference_kats.json) against the CUDA
product path, through the C ABI.  Needs a B200."""
import pytest

import kat_checks
from adapters import CudaImpl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def impl():
    return CudaImpl()


@pytest.mark.parametrize("check", kat_checks.ALL_CHECKS, ids=lambda f: f.__name__)
def test_reference_kats_on_gpu(impl, check):
    check(impl)
