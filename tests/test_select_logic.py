"""The exact-quantile window logic (pybmc_b200/csrc/select_logic.h, shared by host and device) on
adversarial columns: atoms, gaps between adjacent order statistics, heavy tails, outliers, biased
arrival order, tiny buffers.  Pure C++ on the CPU: compiled here with g++."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_selection_state_machine(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "select_harness")
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "select_harness.cpp")],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "select logic ok" in r.stdout
