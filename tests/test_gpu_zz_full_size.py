"""BASELINE config 2 at full size on the GPU, through size-independent properties (runs last: it is the longest GPU test)."""
import pytest

import cases
import mgtest as T
from minigraph_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return capi.load_product()


def test_config2_full_size_properties(lib, workdir):
    cases.case_full_size(lib, workdir, n_reads=10000, n_sub=300, n_ref=100)
