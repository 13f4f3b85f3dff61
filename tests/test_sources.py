"""Independent-source waveforms, break points and step caps of the transient driver (tran_driver.h), pinned on values
worked by hand from the reference's formulas (src/DeviceModelPKG/Core/N_DEV_SourceData.C: PulseData::updateSource
:1168-1248 / getBreakPoints :1442-1500 / getMaxTimeStepSize :1518, ExpData :811-835, SFFMData :2908-2922, PWLinData
:1770-1886 / :2044-2110), and the break-point behaviour of the time loop on the oracle backend (reference devices)."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_ref
import xyce_b200

LIB = xyce_b200.load_library()
LIB.xgpu_source_value.restype = C.c_double
LIB.xgpu_source_max_step.restype = C.c_double
dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def value(stype, params, t, pwl=None, bptol=0.0):
    p = np.zeros(7); p[:len(params)] = params
    tv = np.ascontiguousarray(pwl, dtype=np.float64).reshape(-1) if pwl is not None else None
    return LIB.xgpu_source_value(stype, dp(p), dp(tv) if tv is not None else None, C.c_double(t), C.c_double(bptol))


def breakpoints(stype, params, t, pwl=None):
    p = np.zeros(7); p[:len(params)] = params
    tv = np.ascontiguousarray(pwl, dtype=np.float64).reshape(-1) if pwl is not None else None
    out = np.zeros(64)
    n = LIB.xgpu_source_breakpoints(stype, dp(p), dp(tv) if tv is not None else None, C.c_double(t), 64, dp(out))
    return out[:n]


PULSE = [0.0, 2.0, 1e-9, 2e-9, 4e-9, 5e-9, 20e-9]          # v1 v2 td tr tf pw per


@pytest.mark.parametrize("t,want", [
    (0.0, 0.0), (1e-9, 0.0),                 # in the delay: V1
    (2e-9, 1.0),                             # half way up the rise: V1 + (V2 - V1) (t - TD) / TR
    (3e-9, 2.0), (5e-9, 2.0), (8e-9, 2.0),   # end of the rise (tolerant corner), flat top, end of the top
    (10e-9, 1.0),                            # half way down the fall
    (12e-9, 0.0), (15e-9, 0.0),              # end of the fall, rest of the period
    (22e-9, 1.0),                            # second period (t - TD > PER is folded back): half way up again
    (43e-9, 2.0),                            # third period, end of the rise (2 x PER + TD + TR)
])
def test_pulse_values(t, want):
    assert value(1, PULSE, t, bptol=1e-18) == pytest.approx(want, abs=1e-12)


def test_pulse_corner_within_bptol_counts_as_the_corner():
    # time - TR within bpTol of zero takes the ramp branch (value V2 at its end), just beyond it the flat top
    assert value(1, PULSE, 3e-9 + 5e-19, bptol=1e-18) == pytest.approx(2.0, abs=1e-9)
    assert value(1, [0, 2, 0, 0.0, 0.0, 5e-9, 0], 1e-9) == 2.0          # TR = 0: no division by zero, V2 on the top
    assert value(1, [0, 2, 0, 0.0, 0.0, 5e-9, 0], 6e-9) == 0.0


def test_pulse_break_points_cover_this_period_and_the_next():
    bp = breakpoints(1, PULSE, 0.0)
    assert np.allclose(bp, [1e-9, 3e-9, 8e-9, 12e-9, 21e-9, 23e-9, 28e-9, 32e-9, 41e-9], rtol=1e-12)
    bp = breakpoints(1, PULSE, 47e-9)                        # (47 - 1) / 20 -> period index 2
    assert np.allclose(bp[:4], [41e-9, 43e-9, 48e-9, 52e-9], rtol=1e-12) and np.isclose(bp[-1], 81e-9)
    assert len(breakpoints(1, [0, 1, 1e-9, 1e-9, 1e-9, 1e-9, 0.0], 0.0)) == 4          # PER = 0: one pulse only
    assert LIB.xgpu_source_max_step(1, dp(np.array(PULSE)), C.c_double(0.0)) == pytest.approx(0.1e-9)       # in the delay: TD / 10
    assert LIB.xgpu_source_max_step(1, dp(np.array(PULSE)), C.c_double(5e-9)) == pytest.approx(2e-9)        # then PER / 10
    assert LIB.xgpu_source_max_step(2, dp(np.zeros(7)), C.c_double(0.0)) == 1e99


def test_exp_and_sffm_values():
    EXP = [0.5, 2.5, 1e-9, 2e-9, 6e-9, 3e-9]               # v1 v2 td1 tau1 td2 tau2
    assert value(3, EXP, 0.5e-9) == 0.5
    assert value(3, EXP, 3e-9) == pytest.approx(0.5 + 2.0 * (1 - math.exp(-1.0)), rel=1e-15)
    assert value(3, EXP, 9e-9) == pytest.approx(0.5 + 2.0 * (1 - math.exp(-4.0)) - 2.0 * (1 - math.exp(-1.0)), rel=1e-14)
    SFFM = [1.0, 0.5, 1e6, 2.0, 1e5]                       # v0 va fc mdi fs
    t = 3.3e-7
    assert value(4, SFFM, t) == pytest.approx(1.0 + 0.5 * math.sin(2 * math.pi * 1e6 * t + 2.0 * math.sin(2 * math.pi * 1e5 * t)), rel=1e-15)


def test_pwl_values_and_break_points():
    tv = [(1e-9, 0.0), (2e-9, 3.0), (5e-9, 3.0), (6e-9, -1.0)]
    P = [0.5e-9, 0, 4, 0, 0]                               # td, offset, count, repeat, repeattime
    assert value(5, P, 0.2e-9, tv) == 0.0                  # before TD
    assert value(5, P, 1.0e-9, tv) == 0.0                  # first segment runs from (0, 0) to the first point
    assert value(5, P, 2.0e-9, tv) == pytest.approx(1.5)   # half way between the first two points (t - TD = 1.5 ns)
    assert value(5, P, 4.0e-9, tv) == pytest.approx(3.0)
    assert value(5, P, 6.0e-9, tv) == pytest.approx(1.0)   # t - TD = 5.5 ns: half way down 3 -> -1
    assert value(5, P, 9.0e-9, tv) == -1.0                 # past the last point: holds
    assert np.allclose(breakpoints(5, P, 0.0, tv), [1.5e-9, 2.5e-9, 5.5e-9, 6.5e-9])
    R = [0.0, 0, 4, 1, 2e-9]                               # repeat from t = 2 ns: loop length 4 ns
    assert value(5, R, 7.0e-9, tv) == pytest.approx(value(5, R, 3.0e-9, tv))          # 7 ns = 6 + 1 -> loop time 3 ns
    assert value(5, R, 6.5e-9, tv) == pytest.approx(value(5, R, 2.5e-9, tv))
    bp = breakpoints(5, R, 7.0e-9, tv)                     # points at or after REPEATTIME, shifted by one loop
    assert np.allclose(bp, [6e-9, 9e-9, 10e-9])
    off = [(0.0, 9.0)] * 3 + tv                            # a second source's table in front: offset in points
    assert value(5, [0.5e-9, 3, 4, 0, 0], 2.0e-9, off) == pytest.approx(1.5)


@pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind", ["pulse", "pwl"])
def test_time_loop_lands_on_break_points_and_restarts_there(kind):
    """Mixed-device netlist driven by a PULSE / PWL input on the oracle backend: every source corner inside the run is an
    accepted time point (StepErrorControl::updateStopTime), the step after it is a first-order restart, and no step
    exceeds a tenth of the pulse period (PulseData::getMaxTimeStepSize)."""
    from dev_common import mixed_netlist
    ref, lin, src, x0, probes = mixed_netlist()
    pwl = None
    if kind == "pulse":
        src["type"][0] = 1; src["params"][0] = [0.0, 2.0, 0.2e-6, 0.05e-6, 0.1e-6, 0.4e-6, 1.0e-6]
        corners = [0.2e-6, 0.25e-6, 0.65e-6, 0.75e-6, 1.2e-6, 1.25e-6, 1.65e-6, 1.75e-6]
    else:
        pwl = [(0.1e-6, 0.0), (0.3e-6, 2.0), (0.9e-6, 2.0), (1.0e-6, -1.0), (1.6e-6, 0.5)]
        src["type"][0] = 5; src["params"][0] = [0.0, 0, len(pwl), 0, 0, 0, 0]
        corners = [t for t, _ in pwl]
    ref.set_flags(transient=1)
    r = ref.tran_run(x0, 2e-6, 1e-9, probes, lin, src, pwl=pwl)
    assert r["rc"] == 0
    t = r["t"]
    for c in corners:
        assert np.min(np.abs(t - c)) <= 1e-18 + 1e-12 * c, c
    acc = r["steps"][r["steps"][:, 4] > 0]
    stops = sorted(corners) + [2e-6]
    for c, nxt in zip(stops[:-1], stops[1:]):
        k = int(np.argmin(np.abs(acc[:, 0] - c)))
        # restart (OneStep::initialize away from t = 0): order 1, first step at most a tenth of the way to the next stop
        assert acc[k + 1, 3] == 1 and acc[k + 1, 1] <= 0.1 * (nxt - c) * (1 + 1e-9)
    if kind == "pulse":
        assert np.max(acc[:, 1]) <= 0.1e-6 * (1 + 1e-9)
    # the input node follows the source exactly (ideal voltage source)
    vin = r["wave"][:, 0]
    for ti, vi in zip(t[1:], vin[1:]):
        p = np.zeros(7); p[:] = src["params"][0]
        assert abs(vi - value(int(src["type"][0]), p, ti, pwl, 1e-20)) < 1e-9
