"""The restated Newton / step-control driver (xyce_b200/csrc/tran_driver.h) pinned on tables worked BY HAND from the
reference source, through a scripted backend (tests/host_mirror/driver_host.cpp) -- the GPU-vs-oracle .TRAN tests run the
same driver on both sides, so without these the "identical Newton iteration counts" claim would be self-referential.

Reference read for the expected values:
  DampedNewton::converged_           src/NonlinearSolverPKG/N_NLS_DampedNewton.C:1191-1397 (return codes N_NLS_ReturnCodes.h:
                                     normTooSmall 1, normalConvergence 2, nearConvergence 3, smallUpdate 4, tooManySteps -1,
                                     updateTooBig -2, stalled -3, nanFail -6, linearSolverFailed -9)
  transient / DC_OP defaults         N_NLS_NLParams.C:101-112 (maxNewtonStep 20, deltaXTol 0.33, RHSTol 1e-2, smallUpdateTol 1e-6),
                                     N_NLS_NLParams.h:462-663 (200, 1.0, 1e-6)
  OneStep::completeStep / rejectStep src/TimeIntegrationPKG/N_TIA_OneStep.C (rr = (tolAimFac / (estOverTol + 1e-4))^(1 / (order + 1)),
                                     tolAimFac 0.5, r_min 0.25, r_max 0.9, r_hincr_test = r_hincr = 2, step / 8 on a Newton failure)
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "host_mirror")])
LIB = C.CDLL(os.path.join(HERE, "host_mirror", "libxb_driver.so"))
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
NAN, BIG = float("nan"), 10.0


def newton(dc, evals, lin=(), max_newton=0):
    """evals: list of (||rhs||_2, ||rhs||_inf, max |dx / w|, all devices converged); [0] is the initial residual"""
    a = [np.ascontiguousarray([e[k] for e in evals], dtype=np.float64) for k in range(3)]
    dv = np.ascontiguousarray([e[3] for e in evals], dtype=np.int32)
    ln = np.ascontiguousarray(list(lin) + [0] * 64, dtype=np.int32)
    it = C.c_int()
    st = LIB.xbh_newton_script(int(dc), len(evals), dp(a[0]), dp(a[1]), dp(a[2]), ip(dv), len(ln), ip(ln), int(max_newton), C.byref(it))
    return st, it.value


INIT = (1.0, 1.0, 0.0, 1)
GOOD = (1e-3, 5e-3, 0.2, 1)           # ||rhs||_inf <= RHSTol = 1e-2 and update <= deltaXTol = 0.33
SLOW = (0.5, 0.5, BIG, 1)             # nothing satisfied, residual halves (rate 0.5: no stall)

NEWTON_TABLE = [
    # name, dc, script, linear rc, maxNewtonStep, expected (status, iterations)
    ("normal convergence at the first step", 0, [INIT, GOOD], (), 0, (2, 1)),
    ("three steps", 0, [INIT, SLOW, (0.25, 0.25, BIG, 1), GOOD], (), 0, (2, 3)),
    ("update test alone is not enough", 0, [INIT, (0.5, 0.5, 0.2, 1), GOOD], (), 0, (2, 2)),
    ("residual test alone is not enough", 0, [INIT, (1e-3, 5e-3, 0.5, 1), GOOD], (), 0, (2, 2)),
    ("norm too small", 0, [INIT, (1e-17, 1e-17, BIG, 1)], (), 0, (1, 1)),
    ("small update", 0, [INIT, (0.5, 0.5, 1e-7, 1)], (), 0, (4, 1)),
    ("devices not converged keep it going", 0, [INIT, (1e-3, 5e-3, 0.2, 0), (1e-3, 5e-3, 0.2, 0), GOOD], (), 0, (2, 3)),
    ("devices never converge", 0, [INIT, (1e-3, 5e-3, 0.2, 0)], (), 4, (-1, 4)),
    ("near convergence at the step limit, transient only", 0, [INIT, SLOW, (0.4, 0.4, BIG, 1), (0.3, 0.3, BIG, 1)], (), 3, (3, 3)),
    ("the same script in DC_OP mode: too many steps", 1, [INIT, SLOW, (0.4, 0.4, BIG, 1), (0.3, 0.3, BIG, 1)], (), 3, (-1, 3)),
    ("step limit without 10 % reduction", 0, [INIT, (0.99, 0.99, BIG, 1), (0.97, 0.97, BIG, 1), (0.95, 0.95, BIG, 1)], (), 3, (-1, 3)),
    ("step limit with the residual growing at the end", 0, [INIT, SLOW, (0.2, 0.2, BIG, 1), (0.3, 0.3, BIG, 1)], (), 3, (-1, 3)),
    ("update too big", 0, [(1e-300, 1e-300, 0.0, 1), (1e8, 1e8, BIG, 1)], (), 0, (-2, 1)),
    ("NaN residual", 0, [INIT, (NAN, NAN, BIG, 1)], (), 0, (-6, 1)),
    ("linear solver failure", 0, [INIT, SLOW], (5,), 0, (-9, 1)),
    ("stagnation below 90 %: near convergence", 0, [INIT, (0.8, 0.8, BIG, 1)] + [(0.8 * 0.9995 ** k, 0.8, BIG, 1) for k in range(1, 6)], (), 0, (3, 6)),
    ("stagnation while growing: stalled", 0, [INIT, (0.8, 0.8, BIG, 1)] + [(0.8 * 1.0005 ** k, 0.8, BIG, 1) for k in range(1, 6)], (), 0, (-3, 6)),
    ("stagnation above 90 %: stalled", 0, [INIT, (0.95, 0.95, BIG, 1)] + [(0.95 * 0.9995 ** k, 0.95, BIG, 1) for k in range(1, 6)], (), 0, (-3, 6)),
    ("no stagnation test in DC_OP mode", 1, [INIT, (0.8, 0.8, BIG, 1)] + [(0.8 * 0.9995 ** k, 0.8, BIG, 1) for k in range(1, 8)] + [(1e-9, 1e-9, 0.5, 1)], (), 0, (2, 9)),
    ("DC_OP tolerances: RHSTol 1e-6, deltaXTol 1", 1, [INIT, (1e-3, 5e-3, 0.2, 1), (1e-8, 1e-8, 0.9, 1)], (), 0, (2, 2)),
]


@pytest.mark.parametrize("name,dc,script,lin,maxn,want", NEWTON_TABLE, ids=[t[0] for t in NEWTON_TABLE])
def test_newton_return_codes(name, dc, script, lin, maxn, want):
    assert newton(dc, script, lin, maxn) == want


def step(accept, t, h, last_h, order, nsteps, est, stop=1.0, hmin=1e-15, hmax=1.0, nef=0, newton_status=2, max_order=2):
    out = np.zeros(6)
    ok = LIB.xbh_step_control(int(accept), C.c_double(t), C.c_double(h), C.c_double(last_h), int(order), int(nsteps), C.c_double(est),
                              C.c_double(stop), C.c_double(hmin), C.c_double(hmax), int(nef), int(newton_status), int(max_order), dp(out))
    return dict(ok=bool(ok), h=out[0], order=int(out[1]), next_time=out[2], time=out[3], saved=out[4], nef=int(out[5]))


H = 1e-3


def test_complete_step_ratio_table():
    rr = lambda est, order: (0.5 / (est + 1e-4)) ** (1.0 / (order + 1.0))
    # order 1, fewer than 2 steps taken: the order stays
    r = step(1, 0.1, H, H, 1, 0, 0.1);  assert rr(0.1, 1) >= 2 and r["h"] == 2 * H and r["order"] == 1            # doubling
    r = step(1, 0.1, H, H, 1, 0, 0.3);  assert 1 < rr(0.3, 1) < 2 and r["h"] == H                                 # dead band: step kept
    r = step(1, 0.1, H, H, 1, 0, 0.5);  assert rr(0.5, 1) <= 1 and r["h"] == pytest.approx(0.9 * H, rel=1e-15)    # clamped to r_max
    r = step(1, 0.1, H, H, 1, 0, 0.9);  assert r["h"] == pytest.approx(rr(0.9, 1) * H, rel=1e-15)                 # inside [r_min, r_max]
    r = step(1, 0.1, H, H, 1, 0, 100.); assert r["h"] == pytest.approx(0.25 * H, rel=1e-15)                       # clamped to r_min
    assert r["time"] == pytest.approx(0.1 + H) and r["next_time"] == pytest.approx(0.1 + H + 0.25 * H)
    # second accepted step: order 2 is tried with ITS ratio; kept only when that ratio exceeds 1.05
    r = step(1, 0.1, H, H, 1, 1, 0.01); assert rr(0.01, 2) > 1.05 and r["order"] == 2 and r["h"] == 2 * H
    r = step(1, 0.1, H, H, 1, 1, 0.45); assert rr(0.45, 2) <= 1.05 and r["order"] == 1 and r["h"] == H
    r = step(1, 0.1, H, H, 1, 1, 0.01, max_order=1); assert r["order"] == 1 and r["h"] == 2 * H                   # MAXORD = 1
    # already at order 2: cube-root ratio
    r = step(1, 0.1, H, H, 2, 5, 0.9);  assert r["order"] == 2 and r["h"] == pytest.approx(rr(0.9, 2) * H, rel=1e-15)
    r = step(1, 0.1, H, H, 2, 5, 0.05); assert rr(0.05, 2) >= 2 and r["h"] == 2 * H


def test_complete_step_limits_and_stop_time():
    r = step(1, 0.1, H, H, 1, 0, 0.01, hmax=1.5 * H); assert r["h"] == 1.5 * H                                     # maximum step
    r = step(1, 0.1, H, H, 1, 0, 100., hmin=0.5 * H); assert r["h"] == 0.5 * H                                     # minimum step
    # the doubled step would pass the stop time: clipped onto it, the unclipped step is remembered (savedTimeStep)
    r = step(1, 0.8, 0.1, 0.1, 1, 0, 0.01, stop=1.0, hmax=1.0)
    assert r["time"] == pytest.approx(0.9) and r["next_time"] == 1.0 and r["h"] == pytest.approx(0.1) and r["saved"] == pytest.approx(0.2)


def test_reject_step_table():
    r = step(0, 0.1, H, H, 2, 5, 0.0, newton_status=-1)            # Newton failure: an eighth of the step, minimum order
    assert r["ok"] and r["h"] == H / 8 and r["order"] == 1 and r["nef"] == 1 and r["next_time"] == pytest.approx(0.1 + H / 8)
    r = step(0, 0.1, H, H, 2, 5, 4.0)                              # first error-test failure: order 1 and ITS ratio, clamped
    assert r["order"] == 1 and r["h"] == pytest.approx(math.sqrt(0.5 / 4.0001) * H, rel=1e-15)
    r = step(0, 0.1, H, H, 1, 5, 1.2)
    assert r["h"] == pytest.approx(math.sqrt(0.5 / 1.2001) * H, rel=1e-15)
    r = step(0, 0.1, H, H, 1, 5, 1.0001)                           # ratio above r_max: clamped to 0.9
    assert r["h"] == pytest.approx(0.9 * H, rel=1e-15) or r["h"] == pytest.approx(math.sqrt(0.5 / 1.0002) * H, rel=1e-15)
    r = step(0, 0.1, H, H, 2, 5, 4.0, nef=1)                       # second failure in a row: r_min
    assert r["h"] == pytest.approx(0.25 * H, rel=1e-15) and r["order"] == 1 and r["nef"] == 2
    r = step(0, 0.1, H, H, 1, 5, 4.0, nef=14)                      # maxNumfail = 15 reached
    assert not r["ok"]
    r = step(0, 0.1, H, H, 1, 5, 4.0, nef=1, hmin=0.5 * H)         # cannot go below the minimum step
    assert not r["ok"]
