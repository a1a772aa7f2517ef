"""Builds and runs the C++ host-mirror test (include/mp2gpu_plonky2.hpp over the C ABI) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, name="test_host_mirror"):
    exe = os.path.join(str(tmp_path), name)
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    pkg = os.path.join(ROOT, "mapreduce_plonky2_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
           "-L" + pkg, "-lmp2gpu", "-L" + os.path.join(ROOT, "oracle"), "-lmp2oracle",
           "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"]
    subprocess.run(cmd, check=True, env=env)
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path, oracle):
    """CPU: the header compiles against mp2gpu.h and links against both shared libraries."""
    exe = _build(tmp_path)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path, oracle):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp host mirror OK" in r.stdout


def test_cpp_host_logic_on_cpu(tmp_path, oracle):
    """Challenger and the FRI reduction schedule of the C++ mirror, run on the CPU over the oracle's permutation."""
    exe = _build(tmp_path, "test_host_logic")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpp host logic OK" in r.stdout
