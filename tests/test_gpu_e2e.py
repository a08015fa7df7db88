"""GPU end-to-end parity through the public API against golden vectors produced by the unmodified reference:
pred_tracks within 1e-3 px (north-star tolerance), visibility thresholding exact."""
import pytest
import torch

from cases import CASES, compare, load_golden, run_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_matches_reference_golden(name):
    got = run_cuda(name)
    rep = compare(got, load_golden(name), tol_px=1e-3, tol_logit=1e-3)
    print(name, rep)


def test_library_is_the_path_that_ran():
    """The .so must be loaded in this process (no silent eager fallback)."""
    from cotracker_b200 import engine
    assert engine._lib is not None
    with open("/proc/self/maps") as f:
        assert "libct3_b200.so" in f.read()


def test_cpu_input_is_rejected():
    from cotracker_b200 import engine
    from cotracker_b200.build import build_cotracker
    m = build_cotracker(None, offline=True, window_len=8)
    with pytest.raises(engine.EngineError):
        m(torch.zeros(1, 2, 3, 64, 64), torch.zeros(1, 1, 3))
