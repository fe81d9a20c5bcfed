"""The C-ABI library loads and exports every symbol include/vpho_b200.h declares (no compute calls: CPU-only box)."""
import ctypes
import os
import re

import pytest

from vpho_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "vpho_b200.h")).read()
    return sorted(set(re.findall(r"\b(vpho_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from vpho_b200 import build
    lib = ctypes.CDLL(build.build())
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(capi.EXPORTED_SYMBOLS) == names, "ctypes signatures and header disagree"


def test_product_loader_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    capi._default = None
    with pytest.raises(capi.VphoError):
        capi.lib()


def test_hoi_args_struct_matches_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "vpho_b200.h")).read()
    end = hdr.index("} vpho_hoi_args;")
    body = hdr[hdr.rindex("typedef struct {", 0, end):end]
    fields = re.findall(r"\b([a-zA-Z_0-9]+);", body)
    assert fields == [f[0] for f in capi.HoiArgs._fields_]


def test_sample_args_struct_matches_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "vpho_b200.h")).read()
    end = hdr.index("} vpho_sample_args;")
    body = re.sub(r"/\*.*?\*/", "", hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"):end], flags=re.S)
    fields = []
    for decl in body.split(";"):
        if decl.strip():
            names = decl.split(",")
            fields.append(names[0].split()[-1].lstrip("*"))          # "const float* feat" -> feat
            fields += [n.strip().lstrip("*") for n in names[1:]]      # "int n_rows, rows_per_feat"
    assert fields == [f[0] for f in capi.SampleArgs._fields_]
