"""Accuracy of the constant-bank exp / log / division of the fast arithmetic variant (xb_fastmath.h),
measured in ulps against correctly rounded host results over the argument ranges BSIM4 uses and beyond."""
import numpy as np
import pytest

import xyce_b200

pytestmark = pytest.mark.gpu


def ulps(got, want):
    want = np.asarray(want, dtype=np.float64)
    return np.abs(got - want) / np.spacing(np.abs(want))


@pytest.fixture(scope="module")
def eng():
    e = xyce_b200.Engine(0)
    yield e
    e.close()


def test_exp(eng):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-100, 100, 200000), rng.uniform(-1, 1, 100000), rng.uniform(-708, 709, 100000),
                        [0.0, -0.0, 1e-300, 34.0, -34.0, 709.78, -745.0]])
    got = eng.selftest_fastmath(0, x)
    want = np.exp(x.astype(np.longdouble)).astype(np.float64)
    assert np.max(ulps(got, want)) <= 2.0
    # saturating ends: finite, monotone, no NaN; NaN propagates
    ends = eng.selftest_fastmath(0, np.array([-1e4, -800.0, 800.0, 1e4, np.nan]))
    assert ends[0] >= 0.0 and ends[0] < 1e-320 and ends[1] < 1e-320 and ends[2] > 1e308 and ends[3] > 1e308 and np.isnan(ends[4])


def test_log(eng):
    rng = np.random.default_rng(2)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 200000)), rng.uniform(0.5, 2.0, 200000), 1.0 + rng.uniform(-1e-6, 1e-6, 50000),
                        [1.0, 2.0, 0.5, np.sqrt(2.0), 1e-308, 1e308]])
    got = eng.selftest_fastmath(1, x)
    want = np.log(x.astype(np.longdouble)).astype(np.float64)
    ok = want != 0.0
    assert np.max(ulps(got[ok], want[ok])) <= 3.0
    assert np.all(got[~ok] == 0.0)
    # special operands take the library path
    sp = eng.selftest_fastmath(1, np.array([0.0, -1.0, np.inf, 5e-324, np.nan]))
    assert sp[0] == -np.inf and np.isnan(sp[1]) and sp[2] == np.inf and abs(sp[3] - np.log(5e-324)) < 1e-12 and np.isnan(sp[4])


def test_div(eng):
    rng = np.random.default_rng(3)
    a = rng.normal(0, 1, 400000) * 10.0 ** rng.uniform(-100, 100, 400000)
    b = rng.normal(0, 1, 400000) * 10.0 ** rng.uniform(-100, 100, 400000)
    got = eng.selftest_fastmath(2, a, b)
    want = (a.astype(np.longdouble) / b.astype(np.longdouble)).astype(np.float64)
    assert np.max(ulps(got, want)) <= 2.0       # a * (1/b) with the reciprocal at <= 1 ulp


def test_sqrt(eng):
    rng = np.random.default_rng(4)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 300000), rng.uniform(0.0, 4.0, 200000), [1.0, 2.0, 4.0, 1e-300, 1e300]])
    got = eng.selftest_fastmath(3, x)
    want = np.sqrt(x.astype(np.longdouble)).astype(np.float64)
    assert np.max(ulps(got, want)) <= 1.0
    sp = eng.selftest_fastmath(3, np.array([0.0, np.inf, -1.0, np.nan]))
    # 0 passes through; negative, NaN -- and +inf, which the model code never takes a root of -- give NaN
    assert sp[0] == 0.0 and np.isnan(sp[1]) and np.isnan(sp[2]) and np.isnan(sp[3])
