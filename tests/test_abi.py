"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/hxb200.h declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hxlib():
    from dft_efe_b200 import build, capi
    if not os.path.exists(capi.LIB_PATH):
        build.build()
    return capi.lib()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "hxb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hx_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported(hxlib):
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(hxlib, n)]
    assert not missing, missing


def test_binding_covers_header(hxlib):
    from dft_efe_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_struct_layout_matches_header(hxlib):
    from dft_efe_b200 import capi
    # 3 x u32 + 7 pointers, with natural alignment
    assert C.sizeof(capi.HaloDesc) == 4 * 4 + 3 * 8 + 8 + 3 * 8
    assert C.sizeof(capi.MeshDesc) % 8 == 0


def test_no_cpu_fallback(hxlib):
    """Without a CUDA device every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dft_efe_b200 import capi, synth
    prob = synth.build_problem(synth.MeshSpec(ncell=(2, 2, 2), p=2))[0]
    with pytest.raises(capi.HxError):
        capi.Plan(prob, max_block=4)
    with pytest.raises(capi.HxError):
        capi.microbench()


def test_every_exported_symbol_is_documented_for_the_integrator(hxlib):
    """INTEGRATION.md tells a dft-efe maintainer what each entry point replaces: no exported symbol may be missing."""
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in declared_symbols() if n not in txt]
    assert not missing, missing
