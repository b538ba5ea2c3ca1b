"""C-ABI library: loads, exports every declared symbol, fails loudly without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = []
    for hdr in ("h2o.h", "h2o_dlpack.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        names += re.findall(r"H2O_API\s+[\w\s\*]+?\b(h2o_\w+)\s*\(", text)
    return sorted(set(names))


def _check_exports(built_lib):
    declared = _declared_symbols()
    assert len(declared) >= 38
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in include/*.h but not exported"
    from silver2_isaacsim_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared  # the ctypes table binds exactly the declared ABI


def test_library_exports_every_declared_symbol(built_lib):
    _check_exports(built_lib)


@pytest.mark.gpu
def test_library_exports_every_declared_symbol_on_the_gpu_box(built_lib):
    """Same check, run by the driver's `-m gpu` pass on the B200 box (the shipped .so, not a rebuild)."""
    _check_exports(built_lib)


def test_no_torch_or_cxx_types_in_the_abi():
    for hdr in ("h2o.h", "h2o_dlpack.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        assert "torch" not in text.replace("torch.cuda.current_stream", "").replace("torch.utils.dlpack", "") \
            .replace("wp.from_torch", "").replace("(Python/PyTorch)", "").replace("wraps torch memory", "") \
            .replace("C++/torch types", "")
        assert "std::" not in text and "at::" not in text


def test_bad_handle_and_messages(built_lib):
    assert built_lib.h2o_destroy(None) == 1  # H2O_ERR_BAD_HANDLE
    assert b"invalid" in built_lib.h2o_last_error()
    assert built_lib.h2o_version().startswith(b"h2o_b200")
    assert built_lib.h2o_n_bodies(None) == -1


def test_create_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = built_lib.h2o_create(ctypes.byref(h), 16, 0, 0)
    assert rc == 10 and not h.value  # H2O_ERR_NO_DEVICE
    assert b"no CPU path" in built_lib.h2o_last_error()
    from silver2_isaacsim_b200 import HydroEngine, H2OError
    with pytest.raises(H2OError):
        HydroEngine(16)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for top in ("silver2_isaacsim_b200", "tools", "examples", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "hydro_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def _check_c_host(built_lib, tmp_path):
    import subprocess
    from silver2_isaacsim_b200 import _lib
    inc = os.path.join(ROOT, "include")
    for hdr in ("h2o.h", "h2o_dlpack.h"):
        src = tmp_path / (hdr + ".c")
        src.write_text(f'#include "{hdr}"\nint main(void) {{ return (int)sizeof(h2o_handle) == 0; }}\n')
        subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, "-fsyntax-only", str(src)],
                       check=True)
    exe = tmp_path / "c_host"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-I", inc, os.path.join(ROOT, "examples", "c_host.c"), "-L", libdir,
                    "-lh2o_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir,
                    "-Wl,-rpath,/usr/local/cuda/lib64", "-o", str(exe)], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert res.returncode == 0 and "7038.67" in res.stdout and "F[0] = (" in res.stdout
        val = float(res.stdout.split("F[0] = (")[1].split(",")[2].split(")")[0])
        assert abs(val - 7038.675) < 0.01
    else:
        assert res.returncode == 1 and "no CPU path" in res.stderr  # fails loudly without a GPU


def test_headers_are_plain_c_and_link(built_lib, tmp_path):
    """include/*.h compile as C99 and the example C host links against the library."""
    _check_c_host(built_lib, tmp_path)


@pytest.mark.gpu
def test_c_host_runs_on_the_gpu_box(built_lib, tmp_path):
    """The plain-C host (examples/c_host.c) drives the engine on the B200 and gets GV1's buoyancy."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    _check_c_host(built_lib, tmp_path)
