"""_cublas_emulation.py: the preload is an optimisation of library GEMMs and must never get in the way."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, env_extra):
    env = dict(os.environ, **env_extra)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()[-1]


def test_inactive_when_torch_came_first_or_when_turned_off():
    import torch  # noqa: F401 -- the test process has torch loaded
    import _cublas_emulation as ce
    ce = importlib.reload(ce)
    assert ce.ACTIVE is False and "torch was imported first" in ce.WHY
    out = _run("import _cublas_emulation as c; print(c.ACTIVE, '|', c.WHY)", {"TTG_CUBLAS_EMULATION": "0"})
    assert out.startswith("False") and "TTG_CUBLAS_EMULATION=0" in out


def test_missing_toolkit_is_not_an_error_and_torch_still_works():
    out = _run("import _cublas_emulation as c; import torch; "
               "print(c.ACTIVE, '|', c.WHY, '|', float((torch.ones(3, 3) @ torch.ones(3, 3)).sum()))",
               {"CUDA_HOME": "/nonexistent", "TTG_CUBLAS_EMULATION": "1"})
    assert out.startswith("False") and "no cuBLAS under" in out and out.endswith("27.0")


def test_preload_before_torch_keeps_torch_working():
    out = _run("import _cublas_emulation as c; import torch; "
               "print(c.ACTIVE, '|', c.WHY, '|', float((torch.ones(3, 3) @ torch.ones(3, 3)).sum()))",
               {"TTG_CUBLAS_EMULATION": "1"})
    assert out.endswith("27.0")
    if out.startswith("True"):
        assert "BF16x9" in out
