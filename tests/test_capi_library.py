"""CPU suite: the C-ABI libraries load and export every symbol the headers declare
(no compute calls without a GPU)."""
import os
import re

import pytest

from conftest import ROOT
from schicexplorer_b200 import _capi


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(schic_[a-z0-9_]+)\s*\(", src)))


def test_headers_and_binding_tables_agree():
    assert sorted(_capi.BASE_SYMBOLS) == declared_functions("schic_mh.h")
    assert sorted(_capi.CUDA_SYMBOLS) == declared_functions("schic_mh_cuda.h")
    assert sorted(_capi.PCA_SYMBOLS) == declared_functions("schic_pca.h")


def test_cuda_library_exports_every_declared_symbol():
    api = _capi.CApi(_capi.CUDA_LIB_PATH)     # loads without a GPU (cudart is linked statically)
    assert api.backend == "cuda"
    assert api.has_symbols(_capi.BASE_SYMBOLS + _capi.CUDA_SYMBOLS + _capi.PCA_SYMBOLS) == []


def test_oracle_library_exports_the_base_abi(oracle):
    assert oracle.backend == "oracle"
    assert oracle.has_symbols(_capi.BASE_SYMBOLS) == []


def test_cuda_library_is_blackwell_native():
    """The shipped .so holds sm_100a SASS only (no PTX-JIT fallback to another arch)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _capi.CUDA_LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    api = _capi.CApi(_capi.CUDA_LIB_PATH)
    if api.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_capi.SchicError) as e:
        api.create(n_neighbors=3)
    assert e.value.status == -3 and "no CPU fallback" in str(e.value)


def test_product_package_never_references_the_oracle():
    pkg = os.path.join(ROOT, "schicexplorer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f
