"""C-ABI checks that need no GPU: the library loads, exports every symbol the header
declares, the ctypes table covers the header, and argument validation fails loudly."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from splat_one_b200 import _lib


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200splat.h")).read()
    return sorted(set(re.findall(r"B200SPLAT_API[^;(]*?(b200splat_\w+)\s*\(", src)))


def test_library_loads_and_exports_every_header_symbol():
    lib = _lib.get_lib()
    syms = _header_symbols()
    assert len(syms) == 61
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200splat.h but not exported"
    assert lib.b200splat_abi_version() == _lib.ABI_VERSION
    assert lib.b200splat_arch() == b"sm_100a"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    src = open(os.path.join(ROOT, "include", "b200splat.h")).read()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"B200SPLAT_API[^;(]*?" + name + r"\s*\((.*?)\);", src, re.S)
        assert m, name
        body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S).strip()
        n = 0 if body in ("", "void") else body.count(",") + 1
        assert n == len(argtypes), (name, n, len(argtypes))


def test_validation_errors_are_reported_without_a_gpu():
    lib = _lib.get_lib()
    # degree 5 is rejected before any CUDA call
    rc = lib.b200splat_sh_fwd(1, 1, 36, 5, None, None, None, None, None)
    assert rc != 0
    assert b"degrees_to_use" in lib.b200splat_last_error()
    with pytest.raises(_lib.B200SplatError):
        _lib.check(rc, lib)
    rc = lib.b200splat_rasterize_fwd(1, 1, 1, 700, None, None, None, None, None, None, 16, 16, 16, 1, 1, None, None,
                                     None, None, None, None, None, None)
    assert rc != 0 and b"channels" in lib.b200splat_last_error()
    rc = lib.b200splat_projection_fwd(1, 1, None, None, None, None, None, None, 8, 8, 0.3, 0.01, 1e10, 0.0, 9,
                                      None, None, None, None, None, None)
    assert rc != 0 and b"camera model" in lib.b200splat_last_error()
    # zero-sized problems are a no-op and succeed
    assert lib.b200splat_sh_fwd(0, 0, 16, 3, None, None, None, None, None) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.B200SplatError, match="no CPU or PyTorch fallback"):
        _lib.get_lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "splat_one_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/b200splat.h must compile as C99 (no C++-isms, no torch
    types) and a plain C program must link against libb200splat.so and call it."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text(
        '#include "b200splat.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  if (b200splat_abi_version() <= 0) return 1;\n"
        "  /* validation error path, no GPU needed */\n"
        "  if (b200splat_sh_fwd(1, 1, 36, 5, 0, 0, 0, 0, 0) == 0) return 2;\n"
        '  printf("%s|%s\\n", b200splat_arch(), b200splat_last_error());\n'
        "  return 0;\n}\n")
    exe = tmp_path / "t"
    lib_dir = os.path.dirname(str(_lib.LIB_PATH))
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                        "-o", str(exe), "-L", lib_dir, "-lb200splat", f"-Wl,-rpath,{lib_dir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith("sm_100a|") and "degrees_to_use" in out.stdout
