"""CPU: pins the synthetic-GGUF writer and the block quantisers against the independent gguf-py
package (reader + dequantisers), and the engine's own C++ GGUF reader against both."""
import ctypes

import numpy as np
import pytest


def test_quantisers_roundtrip_through_gguf_py():
    from gguf import quants, GGMLQuantizationType as T
    from oracle import gguf_synth as S
    rng = np.random.Generator(np.random.PCG64(1))
    x = rng.standard_normal((32, 2048), dtype=np.float32) * 0.05
    for t, gt, tol in ((S.Q8_0, T.Q8_0, 0.01), (S.Q4_K, T.Q4_K, 0.12), (S.Q6_K, T.Q6_K, 0.03)):
        b = S.quantize(x, t)
        assert b.shape == (32, S.row_bytes(t, 2048))
        y = quants.dequantize(b, gt)
        assert np.linalg.norm(y - x) / np.linalg.norm(x) < tol
    # Q8_0 is bit-identical to gguf-py's own (ggml-exact) quantiser
    assert np.array_equal(S.quantize(x, S.Q8_0), quants.quantize(x, T.Q8_0))


def test_q4k_scale_packing_is_inverse_of_gguf_py():
    from gguf.quants import Q4_K
    from oracle import gguf_synth as S
    rng = np.random.Generator(np.random.PCG64(2))
    sc = rng.integers(0, 64, size=(100, 8), dtype=np.uint8)
    mn = rng.integers(0, 64, size=(100, 8), dtype=np.uint8)
    s2, m2 = Q4_K.get_scale_min(S.pack_q4k_scales(sc, mn))
    assert np.array_equal(s2, sc) and np.array_equal(m2, mn)


def test_random_blocks_are_finite_and_scaled():
    from oracle import gguf_synth as S, llama_oracle as O
    rng = np.random.Generator(np.random.PCG64(3))
    for t in (S.Q4_K, S.Q6_K, S.Q8_0):
        for cols in (4096, 14336):
            w = O.dequantize(S.random_blocks(rng, t, 16, cols), t, (16, cols))
            assert np.isfinite(w).all()
            assert 0.5 < w.std() * np.sqrt(cols) < 1.6
            assert abs(w.mean()) * np.sqrt(cols) < 0.2


def test_written_file_reads_back_with_gguf_py(tiny_gguf):
    from gguf import GGUFReader
    from oracle import gguf_synth as S
    r = GGUFReader(tiny_gguf)
    names = [t.name for t in r.tensors]
    plan = S.tensor_plan(S.TINY, "q4_k_m")
    assert names == [p[0] for p in plan]
    assert r.get_field("general.architecture").contents() == "llama"
    assert r.get_field("llama.block_count").contents() == 2
    assert abs(r.get_field("llama.rope.freq_base").contents() - 10000.0) < 1e-3
    for t, (name, ty, shape) in zip(r.tensors, plan):
        if not name.endswith("_norm.weight"):
            assert int(t.tensor_type) == ty and tuple(reversed(t.shape.tolist())) == shape


def test_q4_k_m_recipe_matches_survey_byte_count():
    """Llama-3-8B q4_K_M matrix payload and per-token algorithmic bytes."""
    from oracle import gguf_synth as S
    plan = S.tensor_plan(S.LLAMA3_8B, "q4_k_m")
    total = sum(r * S.row_bytes(t, c) for n, t, (r, c) in plan)
    assert total == 4912898048
    layers = sum(r * S.row_bytes(t, c) for n, t, (r, c) in plan if n.startswith("blk.") and "norm" not in n)
    head = 128256 * S.row_bytes(S.Q6_K, 4096)
    assert layers == 4185391104 and head == 430940160
    assert layers + head + 65 * 4096 * 4 + 2304 == 4617398528
    q6_layers = [i for i in range(32) if S.q4_k_m_uses_q6(i, 32)]
    assert q6_layers == [0, 1, 2, 3, 6, 9, 12, 15, 18, 21, 24, 27, 28, 29, 30, 31]


def test_engine_gguf_reader_agrees(tiny_gguf, hostcheck_lib):
    from oracle import llama_oracle as O
    arch = ctypes.create_string_buffer(256)
    nkv = ctypes.c_uint64()
    tb = ctypes.c_uint64()
    n = hostcheck_lib.hc_gguf_probe(tiny_gguf.encode(), arch, 256, ctypes.byref(nkv), ctypes.byref(tb))
    m = O.load_gguf(tiny_gguf)
    assert n == len(m.raw) and arch.value == b"llama"
    assert tb.value == sum(v[2].nbytes for v in m.raw.values())
    # corrupt / missing files are rejected with a message, not a crash
    assert hostcheck_lib.hc_gguf_probe(b"/nonexistent.gguf", arch, 256, ctypes.byref(nkv), ctypes.byref(tb)) == -1
    assert b"cannot open" in arch.value


def test_engine_gguf_reader_rejects_truncated_files_and_survives_corruption(tiny_gguf, hostcheck_lib, tmp_path):
    """the product's GGUF reader (gguf_file.cpp) on damaged files: every truncation is an error with a message, random byte
    damage in the metadata / tensor-info region is either parsed or rejected -- never a crash or an out-of-bounds read"""
    data = open(tiny_gguf, "rb").read()
    arch = ctypes.create_string_buffer(256)
    nkv, tb = ctypes.c_uint64(), ctypes.c_uint64()
    cut = str(tmp_path / "damaged.gguf").encode()

    def probe(b):
        open(cut, "wb").write(b)
        return hostcheck_lib.hc_gguf_probe(cut, arch, 256, ctypes.byref(nkv), ctypes.byref(tb))
    assert probe(data) > 0
    rng = np.random.Generator(np.random.PCG64(1))
    cuts = sorted(set([0, 1, 3, 4, 8, 12, 16, 20, 24, 31, 32, 64, 100, 1000, 5000, 20000, len(data) // 2, len(data) - 1] + [int(x) for x in rng.integers(0, 60000, 30)]))
    for c in cuts:
        assert probe(data[:c]) == -1 and arch.value, c
    outcomes = {True: 0, False: 0}
    for _ in range(150):
        b = bytearray(data)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, 40000))] = int(rng.integers(0, 256))
        outcomes[probe(bytes(b)) >= 0] += 1
    assert outcomes[True] + outcomes[False] == 150 and outcomes[False] > 0
