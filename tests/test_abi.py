"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/cubep3m_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "cubep3m_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(cubep3m_b200_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported(built):
    from cubep3m_b200.lib import load_library, SYMBOLS
    L = load_library()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in the header but not exported"
    assert set(SYMBOLS) == set(syms), set(SYMBOLS) ^ set(syms)


def test_struct_layout_matches_default_config(built):
    """ctypes mirror == C struct: default_config() of the library fills the same values as the Python mirror."""
    from cubep3m_b200.lib import load_library
    from cubep3m_b200.abi import Config, default_config
    c = Config()
    load_library().cubep3m_b200_default_config(C.byref(c))
    d = default_config()
    for name, _ in Config._fields_:
        a, b = getattr(c, name), getattr(d, name)
        assert a == pytest.approx(b, rel=1e-7), name
    assert abs(c.G - 1.0 / (6.0 * 3.141592654)) < 1e-8       # cubepm.par:148-149


def test_sass_is_sm100a(built):
    from cubep3m_b200.lib import SO_PATH
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    assert not re.search(r"sm_[89]\d", out), "only sm_100a code may be present"


def test_no_gpu_fails_loudly(built):
    """Without a CUDA device init must fail with ECUDA (there is no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cubep3m_b200.lib import ParticleMesh, Cubep3mError
    from cubep3m_b200 import default_config
    with pytest.raises(Cubep3mError) as e:
        ParticleMesh(default_config(nf_tile=112))
    assert e.value.status == 2


def test_product_does_not_import_oracle():
    """The product package must never route through the oracle."""
    pkg = os.path.join(ROOT, "cubep3m_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt.replace("Oracle", "").lower(), f


def test_expansion_matches_formula(built):
    """timestep.f90:241-293 restated in double precision by hand."""
    from cubep3m_b200.lib import load_library
    L = load_library()
    om, ol, w = 0.24, 0.76, -1.0
    for a0, dt in ((0.0099, 1.0), (0.05, 0.3), (0.5, 0.01)):
        da1, da2 = C.c_float(), C.c_float()
        L.cubep3m_b200_expansion(a0, dt, om, ol, w, C.byref(da1), C.byref(da2))

        def half(a, h):
            a3 = a ** (-3 * w) * ol / om
            ak = a * (1 - om - ol) / om
            ad = np.sqrt(4 / 9 * a ** 3 * (1 + ak + a3))
            add = a * a * 4 / 9 * (1.5 + 2 * ak + 1.5 * (1 - w) * a3)
            atd = a * ad * 4 / 9 * (3 + 6 * ak + 1.5 * (2 - 3 * w) * (1 - w) * a3)
            return ad * h + add * h * h / 2 + atd * h ** 3 / 6
        e1 = half(a0, dt / 2)
        e2 = half(a0 + e1, dt / 2)
        assert da1.value == pytest.approx(e1, rel=1e-5)
        assert da2.value == pytest.approx(e2, rel=1e-5)


def test_timestep_twin_limits_ra(built):
    """timestep.f90:58-74: the expansion limiter keeps da/(a+da) <= ra_max = 0.01; dt = min over the limiters."""
    from cubep3m_b200.lib import clock_init, timestep
    c = clock_init(100.0)
    timestep(c)
    assert c.nts == 1 and c.dt_old == 0.0
    assert c.da / c.a <= 0.0101
    c.dt_f_acc = 1e-3
    dt_prev = c.dt
    timestep(c)
    assert c.dt == pytest.approx(1e-3) and c.dt_old == pytest.approx(dt_prev)
    assert c.a_mid < c.a
