"""CPU checks of the drop-in boundary: the built library exports exactly what include/ocs2_ddp_cuda.h declares, the ctypes mirror
matches the header's struct layouts, and nothing computes without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import ocs2_b200 as o2
from ocs2_b200 import build as o2build
from ocs2_b200 import lib as _l

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ocs2_ddp_cuda.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(o2c_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    return o2build.build()  # nvcc cross-compiles sm_100a without a GPU


def test_header_symbols_match_the_binding_table():
    assert _declared_symbols() == sorted(_l.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol_and_nothing_else(libpath):
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    exported = sorted(line.split()[-1] for line in out.splitlines() if " T " in line)
    assert exported == _declared_symbols(), "the library must export exactly the o2c_* ABI (built with -fvisibility=hidden)"
    lib = C.CDLL(libpath)
    for name in _l.EXPORTED_SYMBOLS:
        assert getattr(lib, name) is not None


def test_struct_layouts_match_the_header(tmp_path):
    # o2c_config: 12 int32 + 3 double; o2c_field: ptr + 2 int64; views are arrays of fields plus a few scalars
    assert C.sizeof(_l.Config) == 12 * 4 + 3 * 8
    assert C.sizeof(_l.Field) == 24
    assert C.sizeof(_l.LqView) == 12 * 24 + 8 + 16 + 6 * 24 + 8 + 8 + 16 + 5 * 24 + 8  # + flags (int32, padded)
    assert C.sizeof(_l.SolutionView) == 8 * 24 + 16 + 8
    assert [f[0] for f in _l.Config._fields_][:5] == ["nx", "nu", "nc_max", "num_stages", "batch"]
    # sizes and the offsets of the scalar members as the C compiler lays the header's structs out
    src = tmp_path / "layout.c"
    src.write_text('''#include <stddef.h>
#include <stdio.h>
#include "ocs2_ddp_cuda.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(o2c_config), sizeof(o2c_field), sizeof(o2c_lq_view), sizeof(o2c_solution_view),
         sizeof(o2c_discretization_view), offsetof(o2c_lq_view, nc), offsetof(o2c_lq_view, time), offsetof(o2c_lq_view, event),
         offsetof(o2c_lq_view, jump_A), offsetof(o2c_lq_view, flags), offsetof(o2c_solution_view, status));
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_l.Config), C.sizeof(_l.Field), C.sizeof(_l.LqView), C.sizeof(_l.SolutionView), C.sizeof(_l.DiscretizationView),
            _l.LqView.nc.offset, _l.LqView.time.offset, _l.LqView.event.offset, _l.LqView.jump_A.offset, _l.LqView.flags.offset,
            _l.SolutionView.status.offset]
    assert got == want


def test_abi_version_and_error_strings_without_a_device(libpath):
    lib = _l.load_library()
    assert lib.o2c_abi_version() == 4
    # argument validation happens before any CUDA call
    h = C.c_void_p()
    assert lib.o2c_create(None, C.byref(h)) == 1  # O2C_ERR_INVALID_ARGUMENT
    assert b"" != lib.o2c_last_error()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    with pytest.raises(o2.O2cError) as e:
        o2.BatchedLqSolver(o2.Settings(), 4, 1, 10, 2)
    assert e.value.code == 3  # O2C_ERR_CUDA: the product path fails loudly, it never computes on the host


def test_unsupported_settings_are_rejected_in_the_host_mirror():
    with pytest.raises(o2.O2cError):
        o2.BatchedLqSolver(o2.Settings(algorithm=o2.ALG_SLQ, backwardPassIntegratorType="ODE45"), 4, 1, 10, 2)


def test_headers_compile_as_c99_and_cxx14():
    """The ABI header is plain C (a cgo / JNI / ctypes binding can include it); the front-end header is dependency-free C++14."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    c = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(root, "include", "ocs2_ddp_cuda.h")], capture_output=True, text=True)
    assert c.returncode == 0, c.stderr
    cxx = subprocess.run(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++",
                          os.path.join(root, "include", "ocs2_ddp_cuda", "BatchedRiccatiSolver.h")], capture_output=True, text=True)
    assert cxx.returncode == 0, cxx.stderr


def test_discretization_view_layout():
    """o2c_discretization_view: 8 fields of 24 bytes, the dt pointer, the stage count (padded to 8)."""
    import ctypes as C
    assert C.sizeof(_l.DiscretizationView) == 8 * 24 + 8 + 8
