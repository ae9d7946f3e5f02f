"""C++ host front-ends above the C ABI, driven with stand-ins of the reference's Eigen-backed types (Eigen is absent here):

* include/ocs2_ddp_cuda/BatchedRiccatiSolver.h   — tests/cpp/test_frontend.cpp
* include/ocs2_ddp_cuda/ShardedRiccatiSolver.h   — tests/cpp/test_sharded.cpp: N handles, one host thread each, in ONE process
* include/ocs2_ddp_cuda/GaussNewtonDDP_CUDA.h    — tests/cpp/test_ddp_cuda.cpp: ILQR_CUDA / SLQ_CUDA compiled against stand-ins of
  ocs2_ddp/ILQR.h, SLQ.h (tests/cpp/stubs) and driven through the seam of GaussNewtonDDP.h:167,176

CPU: the headers compile warning-free as C++14 and refuse to run without a CUDA device. GPU: against the CPU oracle."""
import os
import subprocess

import pytest
import torch

import ocs2_b200.lib as o2lib
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
HEADERS = [os.path.join(ROOT, "include", "ocs2_ddp_cuda.h"), os.path.join(ROOT, "oracle", "lq_oracle.h"),
           os.path.join(CPP, "problem_fixture.h"), os.path.join(CPP, "stubs", "ocs2_standins.h")] + \
          [os.path.join(ROOT, "include", "ocs2_ddp_cuda", h) for h in ("BatchedRiccatiSolver.h", "ShardedRiccatiSolver.h", "GaussNewtonDDP_CUDA.h")]


def build_cpp_test(name):
    o2lib.load_library()  # builds / locates ocs2_b200/libocs2_ddp_cuda.so
    orc.lib()             # builds / locates oracle/liblq_oracle.so
    src, out = os.path.join(CPP, name + ".cpp"), os.path.join(CPP, "_build", name)
    libs = [os.path.join(ROOT, "ocs2_b200", "libocs2_ddp_cuda.so"), os.path.join(ROOT, "oracle", "liblq_oracle.so")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in [src] + HEADERS + libs):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    libdir, orcdir = os.path.join(ROOT, "ocs2_b200"), os.path.join(ROOT, "oracle")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", f"-I{orcdir}", f"-I{CPP}/stubs", f"-I{CPP}", src, "-o", out,
           f"-L{libdir}", "-locs2_ddp_cuda", f"-L{orcdir}", "-llq_oracle", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{orcdir}", "-pthread"]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr
    return out


@pytest.mark.parametrize("name", ["test_frontend", "test_sharded", "test_ddp_cuda"])
def test_frontend_compiles_and_refuses_cpu(name):
    exe = build_cpp_test(name)
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the no-device behaviour cannot be observed here")
    done = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert done.returncode == 0, done.stdout + done.stderr
    assert "no CPU fallback" in done.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name,cases", [("test_frontend", 11), ("test_sharded", 7), ("test_ddp_cuda", 10)])
def test_frontend_matches_oracle(name, cases):
    exe = build_cpp_test(name)
    done = subprocess.run([exe, "--gpu"], capture_output=True, text=True, timeout=600)
    print(done.stdout)
    assert done.returncode == 0, done.stdout + done.stderr
    assert done.stdout.count("ok  ") == cases
