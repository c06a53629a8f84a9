"""Pins the oracle's restatement of R's random-number arithmetic (oracle/r_rng.hpp) against widely published
R outputs, and the Philox generator against the Random123 known-answer vectors."""
import numpy as np


def test_unif_rand_known_answers(oracle):
    np.testing.assert_allclose(oracle.rng_probe(1, "unif", 3), [0.2655087, 0.3721239, 0.5728534], atol=5e-8)
    np.testing.assert_allclose(oracle.rng_probe(42, "unif", 2), [0.9148060, 0.9370754], atol=5e-8)
    np.testing.assert_allclose(oracle.rng_probe(123, "unif", 3), [0.2875775, 0.7883051, 0.4089769], atol=5e-8)


def test_exp_rand_known_answers(oracle):
    np.testing.assert_allclose(oracle.rng_probe(1, "exp", 3), [0.7551818, 1.1816428, 0.1457067], atol=5e-8)


def test_norm_rand_known_answers(oracle):
    np.testing.assert_allclose(oracle.rng_probe(1, "norm", 3), [-0.6264538, 0.1836433, -0.8356286], atol=5e-8)
    np.testing.assert_allclose(oracle.rng_probe(123, "norm", 3), [-0.56047565, -0.23017749, 1.55870831], atol=5e-9)


def test_rgamma_moments(oracle):
    for shape, scale in [(0.3, 2.0), (1.0, 1.0), (2.5, 0.5), (20.0, 0.1)]:
        x = oracle.rng_probe(7, "gamma", 200000, shape, scale)
        assert abs(x.mean() / (shape * scale) - 1) < 0.01
        assert abs(x.var() / (shape * scale * scale) - 1) < 0.03


def test_rgamma_distribution(oracle):
    """Kolmogorov-Smirnov against the Gamma cdf in every regime of R's rgamma.c: GS (a < 1), and the three
    (b, si, c) parameter ranges of GD (a <= 3.686, <= 13.022, above)."""
    from scipy import stats
    for shape, scale in [(0.3, 2.0), (0.9, 1.0), (1.0, 1.0), (2.5, 0.5), (3.686, 1.0), (7.0, 0.25), (20.0, 0.1), (400.0, 0.01)]:
        x = oracle.rng_probe(11, "gamma", 100000, shape, scale)
        p = stats.kstest(x, stats.gamma(shape, scale=scale).cdf).pvalue
        assert p > 1e-3, (shape, scale, p)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors: philox4x32-10
    assert [hex(v) for v in oracle.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(v) for v in oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_sample_rule(oracle):
    """RcppArmadillo::sample(x, 1, TRUE, p): descending sort, cumsum, first u <= c (Appendix A.2 of SURVEY.md)."""
    w = [0.2, 0.5, 0.3]
    assert oracle.sample(w, 0.49) == 1
    assert oracle.sample(w, 0.5) == 1
    assert oracle.sample(w, 0.51) == 2
    assert oracle.sample(w, 0.81) == 0
    assert oracle.sample([1.0, 1.0], 0.5) == 0     # ties keep index order
    assert oracle.sample([0.0, 3.0], 0.999) == 1
    import pytest
    with pytest.raises(oracle.OracleError):
        oracle.sample([0.0, 0.0], 0.3)
    with pytest.raises(oracle.OracleError):
        oracle.sample([float("nan"), 1.0], 0.3)
