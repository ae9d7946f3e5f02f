"""C++ host front-end (include/ocs2_ddp_cuda/BatchedRiccatiSolver.h) driven with stand-ins of the reference's Eigen-backed types
(tests/cpp/test_frontend.cpp). CPU: the header compiles warning-free as C++14 and the constructor refuses to run without a CUDA
device. GPU: every kernel family through the front-end against the CPU oracle."""
import os
import subprocess

import pytest
import torch

import ocs2_b200.lib as o2lib
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_frontend.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build", "test_frontend")


def build_frontend_test():
    o2lib.load_library()  # builds / locates ocs2_b200/libocs2_ddp_cuda.so
    orc.lib()         # builds / locates oracle/liblq_oracle.so
    deps = [SRC, os.path.join(ROOT, "include", "ocs2_ddp_cuda.h"), os.path.join(ROOT, "include", "ocs2_ddp_cuda", "BatchedRiccatiSolver.h"),
            os.path.join(ROOT, "oracle", "lq_oracle.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    libdir, orcdir = os.path.join(ROOT, "ocs2_b200"), os.path.join(ROOT, "oracle")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", f"-I{orcdir}", SRC, "-o", OUT, f"-L{libdir}",
           "-locs2_ddp_cuda", f"-L{orcdir}", "-llq_oracle", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{orcdir}"]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr
    return OUT


def test_frontend_compiles_and_refuses_cpu():
    exe = build_frontend_test()
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the no-device behaviour cannot be observed here")
    done = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert done.returncode == 0, done.stdout + done.stderr
    assert "no CPU fallback" in done.stdout


@pytest.mark.gpu
def test_frontend_matches_oracle():
    exe = build_frontend_test()
    done = subprocess.run([exe, "--gpu"], capture_output=True, text=True, timeout=600)
    print(done.stdout)
    assert done.returncode == 0, done.stdout + done.stderr
    assert done.stdout.count("ok  ") == 11
