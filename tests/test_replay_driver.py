"""The driver's call order replayed in plain C through the C ABI (tests/replay_driver.c): the stand-in for the Fortran shim
fortran/particle_mesh_b200.f90, which cannot be compiled here. CPU: it must compile and link against the library; GPU: strict and
resident mode with a checkpoint step, both against the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "replay_driver")


def _build():
    import nvidia
    nccl_lib = os.path.join(list(nvidia.__path__)[0], "nccl", "lib")
    lib_dir, orc_dir = os.path.join(ROOT, "cubep3m_b200"), os.path.join(ROOT, "oracle")
    cmd = ["gcc", "-O2", "-std=c11", "-Wall", "-o", EXE, os.path.join(ROOT, "tests", "replay_driver.c"),
           f"-L{lib_dir}", "-lcubep3m_b200", f"-L{orc_dir}", "-lcubep3m_oracle", "-lm",
           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{orc_dir}", f"-Wl,-rpath-link,{nccl_lib}", "-Wl,--allow-shlib-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_replay_driver_compiles_against_the_c_abi(built):
    exe = _build()
    assert os.path.exists(exe)
    # without a GPU the very first library call must fail loudly (no CPU fallback): status 2 = CUDA failure
    import torch
    if not torch.cuda.is_available():
        data = os.path.join(ROOT, "cubep3m_b200", "data")
        r = subprocess.run([exe, os.path.join(data, "wfxyzf3.npy"), os.path.join(data, "wfxyzc2.npy"), "1", "1"], capture_output=True, text=True)
        assert r.returncode == 2 and "status 2" in r.stderr, (r.returncode, r.stderr[-500:])


@pytest.mark.gpu
def test_replay_driver_strict_and_resident_against_oracle(built):
    exe = _build()
    data = os.path.join(ROOT, "cubep3m_b200", "data")
    r = subprocess.run([exe, os.path.join(data, "wfxyzf3.npy"), os.path.join(data, "wfxyzc2.npy"), "5", "3"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:]); print(r.stderr[-3000:])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "REPLAY OK" in r.stdout
