"""The driver twin of timestep.f90 (cubep3m_b200_timestep / cubep3m_b200_expansion, also the body of the device kernel) against an independent
numpy restatement of SURVEY Appendix C, and against the analytic Einstein-de Sitter solution that the reference's units are built on
(tau = -3 / sqrt(a), variable_initialization.f90:15-34, so a = 9 / tau^2 when Omega_m = 1). CPU only."""
import ctypes as C

import numpy as np
import pytest

from cubep3m_b200.lib import clock_init, load_library, timestep

f32 = np.float32


def _expansion_np(a0, dt0, om, ol, wde=-1.0):
    """timestep.f90:241-293 in numpy: real(8) Taylor expansion evaluated twice with h = dt/2, real(4) results."""
    dt_x = f32(dt0) / f32(2)
    a_x = float(f32(a0))
    das = []
    for half in range(2):
        om_h = 4.0 / 9.0
        a3rlm = a_x ** (-3.0 * wde) * float(f32(ol)) / float(f32(om))
        arkm = a_x * (1.0 - float(f32(om)) - float(f32(ol))) / float(f32(om))
        adot = np.sqrt(om_h * a_x ** 3 * (1.0 + arkm + a3rlm))
        addot = a_x ** 2 * om_h * (1.5 + 2.0 * arkm + 1.5 * (1.0 - wde) * a3rlm)
        atdot = a_x * adot * om_h * (3.0 + 6.0 * arkm + 1.5 * (2.0 - 3.0 * wde) * (1.0 - wde) * a3rlm)
        h = float(dt_x)
        da = f32(adot * h + addot * h * h / 2.0 + atdot * h ** 3 / 6.0)
        das.append(da)
        if half == 0:
            a_x = float(f32(a0) + da)
    return das


@pytest.mark.parametrize("a0,dt0", [(0.0099, 1.0), (0.05, 0.37), (0.5, 0.02), (1.0 / 101, 0.004)])
def test_expansion_matches_numpy_restatement(a0, dt0):
    L = load_library()
    d1, d2 = C.c_float(), C.c_float()
    L.cubep3m_b200_expansion(a0, dt0, 0.27, 0.73, -1.0, C.byref(d1), C.byref(d2))
    r1, r2 = _expansion_np(a0, dt0, 0.27, 0.73)
    assert d1.value == pytest.approx(float(r1), rel=2e-7) and d2.value == pytest.approx(float(r2), rel=2e-7)


def test_einstein_de_sitter_track():
    """Omega_m = 1, Omega_L = 0: the clock must stay on a = 9 / tau^2 while it steps from z = 100 to z = 10 under the ra_max = 0.01 limiter
    (no force limiter active: dt_*_acc = 1000), and every step must respect da / (a + da) <= ra_max (timestep.f90:58-74)."""
    c = clock_init(100.0, omega_m=1.0, omega_l=0.0, ppint=1, pp_ext=1, a_target=1.0 / 11.0)
    n = 0
    while not c.checkpoint_step:
        a_before = c.a
        timestep(c)
        n += 1
        assert c.da / (a_before + c.da) <= 0.01 * (1 + 1e-5)
        assert c.a_mid == pytest.approx(a_before + c.da / 2, rel=1e-6)
        assert c.a == pytest.approx(9.0 / c.tau ** 2, rel=2e-4), n      # third-order Taylor steps of 1 % in a: well inside 2e-4 over ~230 steps
        assert n < 1000
    assert 200 < n < 300 and c.nts == n
    assert c.a == pytest.approx(1.0 / 11.0, rel=1e-4)


def test_limiters_bind_and_dt_old_follows():
    c = clock_init(100.0, ppint=1, pp_ext=0)
    timestep(c)
    first = c.dt
    assert c.dt_old == 0.0 and first > 0
    c.dt_f_acc, c.dt_pp_acc, c.dt_pp_ext_acc, c.dt_c_acc = 0.5 * first, 0.25 * first, 1e-3 * first, 0.75 * first
    timestep(c)
    assert c.dt_old == first
    assert c.dt == pytest.approx(0.25 * first)       # dt_pp_ext_acc only binds with -DPP_EXT (timestep.f90:104-111)
    c.pp_ext = 1
    timestep(c)
    assert c.dt == pytest.approx(1e-3 * first)
    c.ppint = 0
    timestep(c)
    assert c.dt == pytest.approx(0.5 * first)         # without PPINT neither PP limiter is read
