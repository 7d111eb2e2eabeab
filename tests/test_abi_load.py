"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from stinet_b200 import _abi

HEADER = os.path.join(ROOT, "include", "stinet_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stinet_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _abi.load()
    syms = declared_symbols()
    assert len(syms) >= 25
    raw = ctypes.CDLL(_abi.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/stinet_b200.h but not exported"
    assert set(syms) == set(_abi.SIGNATURES), "ctypes binding and header disagree"
    assert lib.stinet_abi_version() == 1


def test_ctypes_arity_matches_the_header_prototypes():
    """Every prototype in include/stinet_b200.h has as many parameters as its ctypes signature lists argument types
    (a silent mismatch would only surface as stack garbage on the GPU box)."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|long long|const char\*)\s+(stinet_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    assert len(protos) == len(_abi.SIGNATURES)
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_abi.SIGNATURES[name][1]), f"{name}: header has {n} parameters, binding {len(_abi.SIGNATURES[name][1])}"


def test_workspace_queries_are_pure_host_calls():
    assert _abi.query("stinet_csr_workspace_bytes", 1000, 6000) > 0
    assert _abi.query("stinet_gemm_workspace_bytes", 4096, 256, 64, 0) > 0
    assert _abi.query("stinet_segnorm_workspace_bytes", 4096, 64, 4) > 0


def test_argument_errors_come_back_as_codes_not_crashes():
    lib = _abi.load()
    rc = lib.stinet_linear_fwd(None, 0, None, 0, None, None, None, 0, 4, 4, 4, 0, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.stinet_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from stinet_b200 import ops, synthetic
    from stinet_b200.models import surfacetextureinpaintingnet as S
    with pytest.raises(_abi.StinetError):
        ops.linear(torch.zeros(4, 4), torch.zeros(4, 4))
    net = S.define_G(input_nc=4, output_nc=3, ngf=8, filter_type="edgeconv", norm="instance", n_blocks=1, n_levels=1,
                     pooling_type="max")
    batch = synthetic.make_batch("grid", 1, 1, size=8)
    with pytest.raises(_abi.StinetError):
        net(batch)
