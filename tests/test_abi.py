"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every entry point that
``include/fastgen_b200.h`` declares; the ctypes structs match the header's field order; the product path refuses to
run without CUDA (there is no CPU fallback).  No compute call is made here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fastgen_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fg_[a-z0-9_]+)\s*\(", src)))


def _struct_fields(name):
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct \{([^{}]*)\}\s*" + name + r"\s*;", src, flags=re.S).group(1)
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        # "const void* a" | "int32_t n, h, w" | "fg_conv_seg seg[FG_MAX_SEG]"
        names = stmt.split(",")
        first = names[0].split()[-1]
        for n in [first] + [x.strip() for x in names[1:]]:
            fields.append(n.replace("*", "").split("[")[0].strip())
    return fields


def test_library_loads_and_exports_every_declared_symbol():
    from fastgen_b200 import _lib

    lib = _lib.lib()
    declared = _declared_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_lib.SYMBOLS) == declared, (sorted(set(declared) ^ set(_lib.SYMBOLS)))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/fastgen_b200.h but not exported"
    assert lib.fg_version() >= 100
    assert isinstance(lib.fg_last_error(), bytes)


@pytest.mark.parametrize("cname,pyname", [("fg_conv_seg", "ConvSeg"), ("fg_conv_params", "ConvParams"),
                                          ("fg_gn_apply_params", "GnApplyParams"),
                                          ("fg_edm_step_params", "EdmStepParams"),
                                          ("fg_ln_mod_params", "LnModParams"),
                                          ("fg_qk_norm_rope_params", "QkNormRopeParams"),
                                          ("fg_flash_attn_params", "FlashAttnParams")])
def test_ctypes_structs_follow_the_header(cname, pyname):
    from fastgen_b200 import _lib

    cls = getattr(_lib, pyname)
    assert [f[0] for f in cls._fields_] == _struct_fields(cname)


def test_no_cpu_fallback():
    from fastgen_b200 import _lib, ops
    from fastgen_b200.edm import EDMPrecond

    a = torch.zeros(8, 8)
    with pytest.raises(_lib.FastGenB200Error):
        ops.axpby(a, torch.ones(8, dtype=torch.float64))
    net = EDMPrecond(img_resolution=8, img_channels=3, model_type="DhariwalUNet", model_channels=32,
                     channel_mult=[1], num_blocks=1, attn_resolutions=[])
    with pytest.raises(_lib.FastGenB200Error):
        net(torch.zeros(1, 3, 8, 8), torch.ones(1))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from fastgen_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libfastgen_b200.so"))
    with pytest.raises(_lib.FastGenB200Error, match="not built"):
        _lib.lib()


def test_invalid_arguments_map_to_value_error_without_a_gpu():
    """argument validation happens on the host before any CUDA call, so it can be exercised on CPU"""
    from fastgen_b200 import _lib

    lib = _lib.lib()
    p = _lib.ConvParams()  # all zero: empty problem
    st = lib.fg_conv_gemm(ctypes.byref(p), None)
    assert st == -1
    with pytest.raises(ValueError, match="conv_gemm"):
        _lib.check(st, "fg_conv_gemm")


def test_header_is_self_contained_c(tmp_path):
    """include/fastgen_b200.h is the drop-in boundary: plain C99 (no CUDA / torch headers), usable from C++ as well"""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "fastgen_b200.h"\n'
                   "int main(void) { fg_conv_params p; fg_flash_attn_params a; (void)p; (void)a; return FG_OK; }\n")
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                   check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)], check=True)


def test_c_program_links_against_the_library(tmp_path):
    """a C caller (what a cgo / JNI / FFI binding would be) links libfastgen_b200.so and gets FG_ERR_INVALID with a message
    for an empty problem -- no GPU needed, argument validation is host-side"""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from fastgen_b200 import _lib

    _lib.lib()  # built?
    libdir = os.path.dirname(_lib.LIB_PATH)
    src = tmp_path / "caller.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "fastgen_b200.h"\n'
                   "int main(void) {\n"
                   "  fg_conv_params p; memset(&p, 0, sizeof p);\n"
                   "  int st = fg_conv_gemm(&p, (fg_stream_t)0);\n"
                   '  printf("%d %d %s\\n", fg_version() >= 100, st, fg_last_error());\n'
                   "  return st == FG_ERR_INVALID ? 0 : 1;\n}\n")
    exe = tmp_path / "caller"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-lfastgen_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert out.startswith("1 -1 ") and "conv_gemm" in out
