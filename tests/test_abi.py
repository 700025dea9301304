"""The C-ABI library loads and exports every symbol include/mcmc_gpu.h
declares (no compute calls: there is no GPU in the build container)."""
import ctypes as C

from mcmc_ocaml_b200 import _abi


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    declared = _abi.declared_symbols()
    assert len(declared) >= 50
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"missing symbols: {missing}"
    lib.mg_abi_version.restype = C.c_int
    assert lib.mg_abi_version() == 1


def test_struct_layouts_match_the_header():
    """sizes the C compiler gives the ABI structs (checked against gcc)"""
    import os
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(_abi.__file__)))
    src = '#include <stdio.h>\n#include "mcmc_gpu.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",sizeof(mg_logfn),sizeof(mg_proposal),sizeof(mg_mcmc_cfg),sizeof(mg_into),sizeof(mg_rj_model),sizeof(mg_rjmcmc_cfg),sizeof(mg_nested_cfg));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "t")]).split()))
    want = [C.sizeof(x) for x in (_abi.mg_logfn, _abi.mg_proposal, _abi.mg_mcmc_cfg, _abi.mg_into, _abi.mg_rj_model,
                                  _abi.mg_rjmcmc_cfg, _abi.mg_nested_cfg)]
    assert sizes == want


def test_no_cpu_fallback():
    """without a CUDA device the product path must fail loudly"""
    import pytest
    import torch

    from mcmc_ocaml_b200 import Context, Failure
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(Failure):
        Context(0, 1)


def test_product_does_not_import_the_oracle():
    import os
    import re
    root = os.path.dirname(os.path.abspath(_abi.__file__))
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle/", text, re.M), f
