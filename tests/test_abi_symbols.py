"""The C-ABI library loads and exports every symbol include/idsp_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = os.path.join(ROOT, "include", "idsp_b200.h")
    pre = subprocess.run(["gcc", "-E", "-P", hdr], capture_output=True, text=True, check=True).stdout
    return sorted(set(re.findall(r"\b(idsp_[a-z0-9_]+)\s*\(", pre)))


def test_header_parses_as_c():
    hdr = os.path.join(ROOT, "include", "idsp_b200.h")
    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", hdr], check=True)


def test_library_exports_every_declared_symbol():
    from idsp_b200 import _lib
    from idsp_b200.build import build

    build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) > 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    # the python binding declares a signature for each of them too
    assert sorted(_lib.SIGNATURES) == syms
    assert L.idsp_b200_version() == 100


def test_no_device_is_a_clean_error_not_a_fallback():
    """Without a usable GPU init must fail loudly (there is no CPU path)."""
    import torch

    if torch.cuda.is_available():
        return
    from idsp_b200 import _lib

    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.idsp_b200_init(0, ctypes.byref(h))
    assert rc < 0 and not h.value
    assert L.idsp_b200_last_error()


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under idsp_b200/ may reference it."""
    pkg = os.path.join(ROOT, "idsp_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, fn), errors="replace").read()
                for pat in ("import oracle", "from oracle", "oracle/", "idsp_oracle", "libidsp_oracle", "orc_"):
                    assert pat not in src, (os.path.join(dp, fn), pat)
