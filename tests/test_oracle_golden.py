"""CPU: the oracle restatement reproduces the golden vectors produced by the unmodified reference."""
import pytest
import torch

from cases import CASES, compare, load_golden, run_oracle

ORACLE_CASES = [n for n, c in CASES.items() if c["kind"] != "predictor_online"]


@pytest.mark.parametrize("name", ORACLE_CASES)
def test_oracle_matches_reference_golden(name):
    torch.manual_seed(0)
    got = run_oracle(name)
    # oracle and reference run the same fp32 PyTorch ops; 1e-4 px leaves room for thread-order noise
    rep = compare(got, load_golden(name), tol_px=1e-4, tol_logit=1e-4)
    print(name, rep)


def test_golden_cases_move():
    """The fixtures are only useful if the tracks actually move (stress cases: several pixels)."""
    g = load_golden("offline_stress")["coords"]
    assert float((g[0, -1] - g[0, 0]).abs().max()) > 2.0
