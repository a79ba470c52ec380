"""A slice of the randomised parity sweep (tests/fuzz_parity.py; 640 cases recorded in profiles/r2_fuzz_parity.txt)
as a regular GPU test: random sizes / densities / views against the C oracle, integer stages bit-exact."""
import pytest

pytestmark = pytest.mark.gpu


def test_fuzz_slice(cuda_device):
    from tests.fuzz_parity import run
    worst = run(n_cases=30, seed=7, verbose=False)
    assert worst["fwd"] <= 1e-5 and worst["grad"] <= 1e-4


def test_fuzz_variants_slice(cuda_device):
    """SH colours of every degree, precomputed covariance, toast composition (tests/fuzz_variants.py)."""
    from tests.fuzz_variants import run
    worst = run(n_cases=20, seed=3, verbose=False)
    assert worst["fwd"] <= 1e-5 and worst["grad"] <= 1e-4
