"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/dist_b200.h declares."""
import ctypes
import importlib
import os
import re

import pytest

import cases

ROOT = cases.ROOT


@pytest.fixture(scope="module")
def libpath():
    return importlib.import_module("dist-renderer_b200.build").build()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "dist_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dist_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared_functions()
    for n in ("dist_render_depth_fwd", "dist_render_normal_fwd", "dist_render_depth_bwd", "dist_decoder_forward",
              "dist_decoder_input_grad", "dist_decoder_backward", "dist_fold_latent", "dist_last_error"):
        assert n in names


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for n in _declared_functions():
        assert hasattr(lib, n), "missing symbol " + n
    assert lib.dist_abi_version() == importlib.import_module("dist-renderer_b200._abi").ABI_VERSION == 3


def test_python_binding_matches_header(libpath):
    abi = importlib.import_module("dist-renderer_b200._abi")
    assert sorted(abi.PROTOTYPES) == _declared_functions()
    abi.lib()


def test_struct_sizes_match_header(libpath, tmp_path):
    """sizeof() of the ctypes mirrors equals the C compiler's view of the header structs."""
    import subprocess
    abi = importlib.import_module("dist-renderer_b200._abi")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "dist_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(dist_net_t),sizeof(dist_camera_t),sizeof(dist_march_t),sizeof(dist_workspace_t));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(abi.Net), ctypes.sizeof(abi.Camera), ctypes.sizeof(abi.March),
                     ctypes.sizeof(abi.Workspace)]


def test_no_oracle_import_in_product():
    """The product package must not reach into oracle/ (test infrastructure)."""
    pk = os.path.join(ROOT, "dist-renderer_b200")
    for f in os.listdir(pk):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pk, f)).read().replace("the CPU oracle", ""), f
