"""CPU: the C-ABI library loads, exports every symbol include/gridllm_native.h declares, and the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gridllm_native.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(gl_[a-z0-9_]+)\s*\(", hdr))
    names.discard("gl_token_cb")
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from gridllm_b200 import native as N
    lib = N.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for s in declared:
        assert hasattr(lib, s), f"libgridllm_native.so does not export {s}"
    assert sorted(N.ABI_SYMBOLS) == declared
    assert lib.gl_abi_version() == 2


def test_header_cites_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "gridllm_native.h")).read()
    for cite in ("OllamaService.ts:97-184", "OllamaService.ts:601-665", "OllamaService.ts:65-83", "OllamaService.ts:85-95"):
        assert cite in hdr


def test_no_cpu_fallback(tiny_gguf):
    """Without a GPU the engine must refuse to exist (GL_ERR_NO_DEVICE), never compute on the CPU."""
    from gridllm_b200 import native as N
    if N.device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(N.NativeError) as ei:
        N.Engine(tiny_gguf)
    assert ei.value.code == N.GL_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under gridllm_b200/ may import or link it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gridllm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
