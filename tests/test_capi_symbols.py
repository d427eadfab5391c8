"""CPU: the C-ABI library loads without a GPU and exports every symbol the headers declare; the
ctypes table covers them; compute entry points fail loudly (no CPU fallback) when no device exists."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ngsb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    from ngsolve_b200 import _capi
    lib = _capi.lib()
    names = declared("ngsb200.h") + declared("ngsb200_workloads.h")
    assert len(names) > 60
    for n in names:
        assert hasattr(lib, n), "libngsb200.so does not export %s" % n
    covered = set(_capi.PROTOTYPES) | set(_capi._SPECIAL)
    missing = [n for n in declared("ngsb200.h") if n not in covered]
    assert not missing, "ctypes table lacks %s" % missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ngsolve_b200 import _capi
    h = ctypes.c_void_p()
    assert _capi.lib().ngsb_ctx_create(-1, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in _capi.lib().ngsb_last_error()
    import ngsolve_b200.la as la
    with pytest.raises(la.NgsbError):
        la.default_context()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ngsolve_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("ngs_oracle", "oracle") or f == "__init__.py" and False, os.path.join(dirpath, f)


def test_every_option_is_documented_in_the_header():
    """ngsb_ctx_set_option's accepted names (csrc/vec.cu) all appear in the option list of include/ngsb200.h"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = re.findall(r'strcmp\(name, "([a-z0-9_]+)"\)', open(os.path.join(root, "ngsolve_b200", "csrc", "vec.cu")).read())
    header = open(os.path.join(root, "include", "ngsb200.h")).read()
    assert len(names) >= 15
    assert [n for n in names if '"%s"' % n not in header] == []
