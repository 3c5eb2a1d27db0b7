"""The far kernel's own FP64 elementary functions (csrc/fastmath.cuh: MUFU-seeded rsqrt with FP64 Newton steps, sincos with
FP64 Cody-Waite range reduction, exp for non-positive arguments) against numpy in extended precision: full double accuracy
(a few ulp), no SFU-grade approximations."""
import ctypes as C
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run(ps, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros((len(x), 6))
    rc = ps.cuda_lib().scuffcuda_selftest_math(ps.context(), len(x), x.ctypes.data, out.ctypes.data)
    assert rc == 0, ps.cuda_lib().scuffcuda_last_error(ps.context())
    return out


def test_rsqrt(ps):
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-12, 12, 200000), rng.uniform(0.5, 4.0, 200000)])
    out = run(ps, x)
    want = 1.0 / np.sqrt(x.astype(np.longdouble))
    e2 = np.abs((out[:, 0] - want) / want).astype(np.float64)
    e1 = np.abs((out[:, 1] - want) / want).astype(np.float64)
    print("rsqrt max rel err: polished %.3e, one-step %.3e (eps = 2.2e-16)" % (e2.max(), e1.max()))
    assert e2.max() <= 2.3e-16


def test_sincos(ps):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-50, 50, 300000), rng.uniform(-9.9e4, 9.9e4, 300000), rng.uniform(-1e-3, 1e-3, 1000),
                        np.arange(-40, 41) * (np.pi / 4)])
    out = run(ps, x)
    xl = x.astype(np.longdouble)
    es = np.abs(out[:, 2] - np.sin(xl)).astype(np.float64)
    ec = np.abs(out[:, 3] - np.cos(xl)).astype(np.float64)
    print("sincos max abs err: sin %.3e cos %.3e" % (es.max(), ec.max()))
    assert es.max() <= 3e-16 and ec.max() <= 3e-16


def test_exp_neg(ps):
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(0, 40, 300000), rng.uniform(0, 699, 100000), [0.0, 1e-300, 700.0, 701.0, 5000.0]])
    out = run(ps, x)
    want = np.exp(-x.astype(np.longdouble))
    ok = x <= 699.0
    rel = np.abs((out[ok, 4] - want[ok]) / want[ok]).astype(np.float64)
    print("exp(-x) max rel err %.3e" % rel.max())
    assert rel.max() <= 4e-16
    assert np.all(out[x > 700.0, 4] == 0.0)


def test_small_integer_division_is_exact(ps):
    """td_math.cuh td_div_int (tabulated reciprocal + one Markstein correction step, used by the Taylor-Duffy ExpRel series in
    place of an IEEE division) returns the correctly rounded quotient: bit-identical to x / d"""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-1, 1, 400000), 10.0 ** rng.uniform(-30, 30, 200000) * rng.choice([-1.0, 1.0], 200000),
                        rng.uniform(-0.1, 0.1, 400000), np.arange(1, 2001, dtype=float), [0.0]])
    out = run(ps, x)
    d = (np.arange(len(x)) % 127) + 1.0
    want = x / d
    bad = np.nonzero(out[:, 5] != want)[0]
    print("x / d: %d of %d quotients differ from the IEEE division" % (len(bad), len(x)))
    assert len(bad) == 0
