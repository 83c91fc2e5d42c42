"""GPU: rolling covariance / turbulence precompute vs the reference's pandas expressions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
pd = pytest.importorskip("pandas")


def _frame(close):
    from finrl_b200 import synthetic as syn

    T, D = close.shape
    return syn.make_frame(close, np.zeros((0, T, D)), np.zeros(T), tech_names=[])


@pytest.mark.parametrize("T,D,lookback", [(300, 30, 252), (90, 7, 40), (150, 100, 60)])
def test_rolling_covariance_matches_pandas(T, D, lookback):
    from finrl_b200 import synthetic as syn
    from finrl_b200.preprocess import rolling_covariance

    close, _, _ = syn.make_tables(T, D, 0, seed=3)
    cov = rolling_covariance(close, lookback).cpu().numpy()
    df = _frame(close)
    assert cov.shape == (T - lookback, D, D)
    for i in (lookback, lookback + 1, (lookback + T) // 2, T - 1):
        # tutorials/2-Advance/FinRL_PortfolioAllocation_Explainable_DRL.py:157-174
        data_lookback = df.loc[i - lookback : i, :]
        price_lookback = data_lookback.pivot_table(index="date", columns="tic", values="close")
        want = price_lookback.pct_change().dropna().cov().values
        np.testing.assert_allclose(cov[i - lookback], want, rtol=1e-9, atol=1e-18)


@pytest.mark.parametrize("name", ["turbulence_d30", "turbulence_d7_dup", "turbulence_d100"])
def test_turbulence_matches_the_reference_function(name):
    """frl_turbulence (batched Jacobi eigen-solve + pinv cut-off) vs goldens produced by the UNMODIFIED
    FeatureEngineer.calculate_turbulence (preprocessors.py:215-267; tests/golden/make_golden.py turbulence):
    a DOW-30 shape, a rank-deficient universe (one ticker duplicated: the rcond cut-off decides) and 100 stocks."""
    import os

    from conftest import GOLDEN
    from finrl_b200.preprocess import turbulence_index

    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    got = turbulence_index(g["close"]).cpu().numpy()
    want = g["turbulence"]
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-12)
    assert (got[:252] == 0).all() and np.array_equal(got > 0, want > 0)  # incl. the two suppressed first positives


def test_preprocess_uses_no_library_linear_algebra():
    import inspect

    from finrl_b200 import preprocess

    src = inspect.getsource(preprocess)
    assert "torch.linalg" not in src and "einsum" not in src and "cumsum" not in src
