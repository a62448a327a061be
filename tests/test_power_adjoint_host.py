"""The reverse sweep of the K3 point function (csrc/jc_power_point.cuh, used by jc_power_adj.cu for multi-direction Jacobians) on
the CPU: the header compiles as plain C++, the harness tests/host/power_adjoint_check.cpp evaluates the sweep with libm on random
points of the physical ranges (linear, takahashi2012, smith2003) and compares every one of the 28 input gradients with a
complex-step derivative of the value function."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_point_adjoint_vs_complex_step(tmp_path):
    exe = str(tmp_path / "padj")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "power_adjoint_check.cpp")], check=True)
    out = subprocess.run([exe, "1500"], capture_output=True, text=True, check=True).stdout
    m = re.search(r"max_rel_err (\S+) max_val_diff (\S+) points (\d+)", out)
    assert m, out
    err, val, n = float(m.group(1)), float(m.group(2)), int(m.group(3))
    assert n == 1500
    assert val == 0.0, out     # the sweep's forward half is the value function, operation for operation
    assert err < 1e-9, out     # observed 1.3e-11 (cancellation in the sound-horizon gradient)


def test_field_numbers_match_the_header():
    """The host build of jc_power_point.cuh carries its own copy of the workspace field numbers: hold it to include/jc_b200.h."""
    hdr = open(os.path.join(ROOT, "include", "jc_b200.h")).read()

    def enum_values(first):
        body = hdr[hdr.index(first):]
        body = body[:body.index("}")]
        names = re.findall(r"^\s*(JC_[A-Z0-9_]+)", body, flags=re.M)
        return {n: i for i, n in enumerate(names)}

    vals = {}
    vals.update(enum_values("JC_NODE_CHI = 0"))
    vals.update(enum_values("JC_SCAL_LN13KEQ = 0"))
    src = open(os.path.join(ROOT, "jax_cosmo_b200", "csrc", "jc_power_point.cuh")).read()
    pairs = re.findall(r"JCP_((?:NODE|SCAL)_[A-Z0-9_]+) = (\d+)", src)
    assert len(pairs) == 27
    for name, num in pairs:
        assert vals["JC_" + name] == int(num), name
