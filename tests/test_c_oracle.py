"""CPU: the C restatement (oracle/c/llama_cpu.c -- the CPU baseline bench.py times) agrees with the
numpy oracle in both activation modes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def oc():
    so = os.path.join(ROOT, "oracle", "_ref", "liboracle_cpu.so")
    src = os.path.join(ROOT, "oracle", "c", "llama_cpu.c")
    if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(so)
    lib.oc_load.restype = C.c_void_p
    lib.oc_load.argtypes = [C.c_char_p, C.c_int]
    lib.oc_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.oc_reset.argtypes = [C.c_void_p]
    lib.oc_free.argtypes = [C.c_void_p]
    lib.oc_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("fixture", ["tiny_gguf", "tiny128_gguf", "tiny_q8_gguf", "tiny_f16_gguf"])
def test_c_port_matches_numpy_oracle(oc, fixture, request):
    from oracle import llama_oracle as O
    path = request.getfixturevalue(fixture)
    om = O.load_gguf(path)
    h = oc.oc_load(path.encode(), 128)
    assert h
    toks = np.random.Generator(np.random.PCG64(5)).integers(0, om.n_vocab - 3, size=20)
    for mode, act in ((0, "exact"), (1, "ggml")):
        oc.oc_reset(h)
        orc = O.LlamaOracle(om, act=act)
        for t in toks:
            lg = np.zeros(om.n_vocab, np.float32)
            assert oc.oc_step(h, int(t), mode, lg.ctypes.data_as(C.c_void_p), None) == 0
            ref = orc.step(int(t))
        # exact mode is tight; the int8-activation mode is discontinuous (a rounding flip moves an
        # activation by 1/127 of its block max), so float-vs-double differences upstream show at ~1e-2
        tol = 1e-4 if (mode == 0 or act == "exact") else 3e-2
        assert np.abs(lg - ref).max() <= tol * np.abs(ref).max(), (fixture, mode)
    oc.oc_free(h)


def test_c_generate_matches_oracle_ids(oc, tiny_gguf):
    from oracle import llama_oracle as O
    om = O.load_gguf(tiny_gguf)
    h = oc.oc_load(tiny_gguf.encode(), 128)
    prompt = np.random.Generator(np.random.PCG64(1000)).integers(0, om.n_vocab - 3, size=16).astype(np.int32)
    ref = O.LlamaOracle(om, act="exact").generate(prompt, 8)
    ids = np.zeros(8, np.int32)
    lps = np.zeros(8, np.float32)
    buf = np.zeros(om.n_vocab, np.float32)
    n = oc.oc_generate(h, prompt.ctypes.data_as(C.c_void_p), 16, 8, 0, ids.ctypes.data_as(C.c_void_p),
                       lps.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p))
    assert n == 8
    for i in range(8):
        if ids[i] != ref["ids"][i]:
            assert ref["margins"][i] < 1e-3
            break
        assert abs(lps[i] - ref["logprobs"][i]) < 1e-3
    oc.oc_free(h)


def test_batched_prefill_equals_token_by_token(oc, tiny128_gguf):
    """oc_prefill (what bench.py's reference arm times for the prompt phase: one pass over the weights for T tokens) computes
    exactly what T calls of oc_step compute; oc_fill_kv moves the position without touching the arithmetic."""
    from oracle import llama_oracle as O
    om = O.load_gguf(tiny128_gguf)
    oc.oc_prefill.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    oc.oc_fill_kv.argtypes = [C.c_void_p, C.c_int, C.c_uint]
    h = oc.oc_load(tiny128_gguf.encode(), 128)
    toks = np.random.Generator(np.random.PCG64(8)).integers(0, om.n_vocab - 3, size=37).astype(np.int32)
    a = np.zeros(om.n_vocab, np.float32)
    b = np.zeros(om.n_vocab, np.float32)
    for t in toks[:5]:
        oc.oc_step(h, int(t), 1, None, None)
    assert oc.oc_prefill(h, toks[5:].ctypes.data_as(C.c_void_p), 32, 1, a.ctypes.data_as(C.c_void_p)) == 0
    assert oc.oc_step(h, 3, 1, a.ctypes.data_as(C.c_void_p), None) == 0          # decode continues on the KV the prefill wrote
    oc.oc_reset(h)
    for t in toks:
        oc.oc_step(h, int(t), 1, b.ctypes.data_as(C.c_void_p), None)
    assert oc.oc_step(h, 3, 1, b.ctypes.data_as(C.c_void_p), None) == 0
    assert np.array_equal(a, b)
    assert oc.oc_fill_kv(h, 64, 1) == 0 and oc.oc_step(h, 3, 1, a.ctypes.data_as(C.c_void_p), None) == 0 and np.isfinite(a).all()
    assert oc.oc_prefill(h, toks.ctypes.data_as(C.c_void_p), 37, 0, None) == -1     # exact mode is not batched
    oc.oc_free(h)
