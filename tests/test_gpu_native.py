"""GPU (-m gpu): the native C++ kernel tests (tests/native/test_kernels, test_bwd) run from pytest, so the driver's
GPU tier counts them.  They call the C ABI straight from C++ (no Python, no torch) and compare every forward /
data-gradient / weight-gradient kernel variant — row-streaming (shared- and tensor-memory operand, ELU on load),
persistent tile, CTA pair, slab, per-tap, stride 2, DepthToSpace, many-split wgrad crossing images — with CPU
loops on the same inputs; the binaries exit non-zero on any mismatch."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

NATIVE = os.path.join(ROOT, "tests", "native")


def _run(binary, env_extra=None, timeout=900):
    path = os.path.join(NATIVE, binary)
    if not os.path.exists(path):  # built by __graft_entry__.build(); build here if the snapshot lacks it
        subprocess.run(["make", "-C", NATIVE, binary], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    env = dict(os.environ)
    env.update(env_extra or {})
    proc = subprocess.run([path], cwd=NATIVE, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=timeout)
    tail = proc.stdout[-4000:]
    assert proc.returncode == 0, f"{binary} failed (rc {proc.returncode}):\n{tail}"
    assert "ALL PASSED" in proc.stdout, tail
    return proc.stdout


def test_native_forward_kernels():
    out = _run("test_kernels")
    assert "FAIL" not in out


def test_native_backward_kernels():
    out = _run("test_bwd")
    assert "FAIL" not in out
