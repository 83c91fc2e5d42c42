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


def test_turbulence_matches_reference_formula():
    from finrl_b200 import synthetic as syn
    from finrl_b200.preprocess import turbulence_index

    T, D = 300, 30
    close, _, _ = syn.make_tables(T, D, 0, seed=4)
    got = turbulence_index(close).cpu().numpy()
    # finrl/meta/preprocessor/preprocessors.py:215-267, restated on the pivoted frame
    piv = _frame(close).pivot(index="date", columns="tic", values="close").pct_change()
    dates = piv.index
    want = [0.0] * 252
    count = 0
    for i in range(252, T):
        cur = piv[piv.index == dates[i]]
        hist = piv[(piv.index < dates[i]) & (piv.index >= dates[i - 252])]
        hist = hist.iloc[hist.isna().sum().min() :].dropna(axis=1)
        d = cur[[x for x in hist]] - np.mean(hist, axis=0)
        temp = d.values.dot(np.linalg.pinv(hist.cov())).dot(d.values.T)
        if temp > 0:
            count += 1
            want.append(float(temp[0][0]) if count > 2 else 0.0)
        else:
            want.append(0.0)
    np.testing.assert_allclose(got, np.asarray(want), rtol=1e-7, atol=1e-12)
    assert (got[:252] == 0).all() and (got[254:] > 0).all()
