"""The C-ABI library loads and exports every symbol include/srwn.h declares (no compute calls:
this runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
import sr_wavenet_b200 as srwn
from sr_wavenet_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "srwn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srwn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libsrwn.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding misses %s" % n
    assert lib.srwn_abi_version() == _lib.ABI_VERSION == 7


def test_binding_has_no_undeclared_symbols(lib):
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_config_struct_matches_header():
    text = open(os.path.join(ROOT, "include", "srwn.h")).read()
    body = text[text.index("typedef struct srwn_config {"):text.index("} srwn_config_t;")]
    fields = re.findall(r"^\s*(?:const\s+)?int32_t\*?\s+(\w+);", body, flags=re.M)
    assert fields == [f[0] for f in _lib.Config._fields_]


def test_argument_errors_need_no_gpu(lib):
    h = ctypes.c_void_p()
    assert lib.srwn_create(None, ctypes.byref(h)) == _lib.ERR_INVALID
    assert b"null" in lib.srwn_last_error()
    dil = (ctypes.c_int32 * 2)(1, 2)
    bad = _lib.Config(7, 2, dil, 2, 32, 128, 32, 128, 5, 0)
    assert lib.srwn_create(ctypes.byref(bad), ctypes.byref(h)) == _lib.ERR_INVALID
    unsupported = _lib.Config(_lib.TEACHER, 2, dil, 3, 32, 128, 32, 128, 5, 0)
    assert lib.srwn_create(ctypes.byref(unsupported), ctypes.byref(h)) == _lib.ERR_UNSUPPORTED
    assert lib.srwn_destroy(None) == _lib.OK
    assert lib.srwn_mol_loss(None, None, None, None, 1, 1, 5, None) == _lib.ERR_INVALID
    n = ctypes.c_size_t()
    assert lib.srwn_stft_workspace_bytes(4, 64000, 512, 256, ctypes.byref(n)) == _lib.OK and n.value > 0
    assert lib.srwn_stft_workspace_bytes(4, 64000, 500, 256, ctypes.byref(n)) == _lib.ERR_INVALID     # not a power of two
    assert lib.srwn_stft_workspace_bytes(4, 256, 512, 256, ctypes.byref(n)) == _lib.ERR_INVALID       # shorter than a frame


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    dil = (ctypes.c_int32 * 2)(1, 2)
    cfg = _lib.Config(_lib.TEACHER, 2, dil, 2, 32, 128, 32, 128, 5, 0)
    h = ctypes.c_void_p()
    assert lib.srwn_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.ERR_CUDA
    with pytest.raises(RuntimeError):
        srwn.WaveNetAutoEncoder(4096, 0, 5, [1, 2], skip_channels=128, latent_channels=32, pool_stride=128)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sr-wavenet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/srwn.h must compile as C99 (no C++ or torch types), and a C program that
    links only against libsrwn.so resolves every declared entry point."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    calls = "\n".join("  p[%d] = (fn_t)%s;" % (i, n) for i, n in enumerate(_declared_symbols()))
    src.write_text('#include "srwn.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t p[%d];\n%s\n'
                   '  printf("%%d %%d\\n", srwn_abi_version(), (int)(sizeof(p) / sizeof(p[0])));\n  return p[0] == 0;\n}\n'
                   % (len(_declared_symbols()), calls))
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-l:libsrwn.so", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == _lib.ABI_VERSION and int(out[1]) == len(_declared_symbols())


def test_product_library_exports_only_the_header(lib):
    """No tuning / debug entry point ships in the product library (those build only with -DSRWN_TUNING)."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (srwn_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == _declared_symbols()
