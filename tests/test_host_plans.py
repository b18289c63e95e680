"""Launch planners are host code: check on the CPU that they terminate and keep their
invariants (a planner bug once hung a GPU run).  Compiles tests/host/plan_check.cu."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc missing")
def test_planners_terminate_and_hold_invariants(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = tmp_path / "plan_check"
    subprocess.run([nvcc, "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                    "-o", str(exe), os.path.join(ROOT, "tests", "host", "plan_check.cu")], check=True, timeout=600)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 bad" in out.stdout
