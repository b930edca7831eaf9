"""Distribution tests of the oracle's samplers (the CUDA samplers are then checked bit-for-bit against
these, tests/test_gpu_parity.py).  Replaces Boost.Random's distributions (core/common.h:542-580) and the
vendored inverse-Gaussian / GIG samplers (common.h:99-111, 234-288); Kolmogorov-Smirnov against SciPy."""
import numpy as np
import pytest
from scipy import stats

import oracle_binding as ob

N = 40000
ALPHA = 1e-4  # fixed seeds: a failure is a bug, not bad luck


def draw(kind, p0=0.0, p1=0.0, p2=0.0, rng_mode=0, seed=123, it=5, n=N):
    out = np.empty(n)
    ob.lib().oracle_sample(kind, rng_mode, seed, it, p0, p1, p2, n, ob.dp(out))
    return out


@pytest.mark.parametrize("rng_mode", [0, 1])
def test_normal_uniform(rng_mode):
    z = draw(0, rng_mode=rng_mode)
    u = draw(1, rng_mode=rng_mode)
    assert stats.kstest(z, "norm").pvalue > ALPHA
    assert stats.kstest(u, "uniform").pvalue > ALPHA
    assert 0 < u.min() and u.max() < 1


@pytest.mark.parametrize("shape,scale", [(0.3, 2.0), (1.0, 1.0), (3.5, 0.01), (503.0, 0.2)])
@pytest.mark.parametrize("rng_mode", [0, 1])
def test_gamma(shape, scale, rng_mode):
    """Boost gamma_distribution(shape, scale) semantics (common.h:552-555)."""
    g = draw(2, shape, scale, rng_mode=rng_mode)
    assert stats.kstest(g, "gamma", args=(shape, 0, scale)).pvalue > ALPHA


@pytest.mark.parametrize("a,b", [(1.0, 1.0), (0.5, 3.0), (12.0, 40.0)])
def test_beta(a, b):
    x = draw(3, a, b)
    assert stats.kstest(x, "beta", args=(a, b)).pvalue > ALPHA


@pytest.mark.parametrize("mean,shape", [(1.0, 1.0), (0.05, 1.0), (7.0, 0.5), (300.0, 1.0)])
def test_inverse_gaussian(mean, shape):
    """InvGauss(mean, shape) (math/random.h:227-230, common.h:99-111); SciPy: invgauss(mu/lambda, scale=lambda)."""
    x = draw(4, mean, shape)
    assert stats.kstest(x, "invgauss", args=(mean / shape, 0, shape)).pvalue > ALPHA


@pytest.mark.parametrize("p,a,b", [(-0.5, 1.0, 2.0), (0.49, 3.0, 0.01), (-0.99, 1.0, 1e-6), (2.5, 0.4, 6.0), (-600.0, 0.02, 40.0),
                                   (-24.0, 1.0, 3.0)])
def test_gig(p, a, b):
    """GIG(p, a, b) density ~ x^(p-1) exp(-(a x + b / x) / 2) (math/random.h:220-223); the parameter
    regimes are those of the DL local scale (p ~ -0.5), DL global (p = N(a-1)), NG local (a_g - 0.5) and
    hierarchical Minnesota (p ~ -N/2).  SciPy: geninvgauss(p, sqrt(ab), scale=sqrt(b/a))."""
    x = draw(5, p, a, b)
    assert np.all(np.isfinite(x)) and x.min() > 0
    if p < -100:  # SciPy's GIG cdf breaks down; here a*x is ~1e-4, so GIG(p, a, b) is InvGamma(-p, b/2) to ~1e-5
        assert stats.kstest(x, "invgamma", args=(-p, 0, b / 2)).pvalue > ALPHA
    else:
        assert stats.kstest(x, "geninvgauss", args=(p, np.sqrt(a * b), 0, np.sqrt(b / a))).pvalue > ALPHA


def test_addressed_stream_is_order_free():
    """Element i of iteration `it` does not depend on how many elements are drawn or in which order."""
    a = draw(2, 2.0, 1.0, n=100)
    b = draw(2, 2.0, 1.0, n=1000)
    assert np.array_equal(a, b[:100])
    c = draw(2, 2.0, 1.0, n=100, it=6)
    assert not np.array_equal(a, c)
