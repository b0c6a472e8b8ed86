"""CPU: the GEMV lane program (gridllm_b200/csrc/rowdot.h: engine row layouts, repackers, dp4a
block decode, activation fixed point) executed on the host against the oracle.  Bit-level layout
bugs are caught here, before any GPU time is spent.  (The product never runs this on the CPU.)"""
import ctypes

import numpy as np
import pytest

from conftest import rel_l2


def _run(lib, t, blocks, rows, cols, x, abits):
    y = np.zeros(rows, np.float32)
    rc = lib.hc_gemv(t, blocks.ctypes.data_as(ctypes.c_void_p), rows, cols, x.ctypes.data_as(ctypes.c_void_p),
                     y.ctypes.data_as(ctypes.c_void_p), abits)
    assert rc == 0
    return y


@pytest.mark.parametrize("tname", ["Q4_K", "Q6_K", "Q8_0"])
@pytest.mark.parametrize("cols", [256, 768, 4096, 14336])
def test_lane_program_matches_oracle(hostcheck_lib, tname, cols):
    from oracle import gguf_synth as S, llama_oracle as O
    t = {"Q4_K": S.Q4_K, "Q6_K": S.Q6_K, "Q8_0": S.Q8_0}[tname]
    rng = np.random.Generator(np.random.PCG64(cols))
    rows = 9
    for src in ("random", "quant"):
        blocks = S.random_blocks(rng, t, rows, cols) if src == "random" else S.quantize(
            rng.standard_normal((rows, cols), dtype=np.float32) / np.sqrt(cols), t)
        wd = O.dequantize(blocks, t, (rows, cols))
        x = rng.standard_normal(cols).astype(np.float32) * rng.uniform(0.1, 10)
        assert rel_l2(_run(hostcheck_lib, t, blocks, rows, cols, x, 16), O.gemv(wd, x, "i16")) < 2e-6
        assert rel_l2(_run(hostcheck_lib, t, blocks, rows, cols, x, 8), O.gemv(wd, x, "q8")) < 2e-6
        assert rel_l2(_run(hostcheck_lib, t, blocks, rows, cols, x, 16), O.gemv(wd, x, "exact")) < 2e-4


def test_edge_activations(hostcheck_lib):
    """all-zero blocks, a single spike, huge / tiny magnitudes"""
    from oracle import gguf_synth as S, llama_oracle as O
    rng = np.random.Generator(np.random.PCG64(9))
    blocks = S.random_blocks(rng, S.Q4_K, 4, 512)
    wd = O.dequantize(blocks, S.Q4_K, (4, 512))
    for x in (np.zeros(512, np.float32), np.eye(1, 512, 77, dtype=np.float32)[0] * 3e4,
              rng.standard_normal(512).astype(np.float32) * 1e-20, rng.standard_normal(512).astype(np.float32) * 1e15):
        y = _run(hostcheck_lib, S.Q4_K, blocks, 4, 512, x, 16)
        ref = O.gemv(wd, x, "i16")
        assert np.isfinite(y).all()
        assert np.allclose(y, ref, rtol=1e-5, atol=1e-30 + 1e-6 * np.abs(ref).max())


def _run_r(lib, t, blocks, rows, cols, x, abits, rpi):
    y = np.zeros(rows, np.float32)
    rc = lib.hc_gemv_r(t, blocks.ctypes.data_as(ctypes.c_void_p), rows, cols, x.ctypes.data_as(ctypes.c_void_p),
                       y.ctypes.data_as(ctypes.c_void_p), abits, rpi)
    assert rc == 0
    return y


@pytest.mark.parametrize("tname", ["Q4_K", "Q6_K", "Q8_0"])
def test_rows_per_item_variants_are_bit_identical(hostcheck_lib, tname):
    """item_dot_* with 4, 2 or 1 rows per item (the GPU consumer picks 4 or 2 by type) give the same bits, including
    ragged last items whose unused slot rows hold garbage"""
    from oracle import gguf_synth as S
    t = {"Q4_K": S.Q4_K, "Q6_K": S.Q6_K, "Q8_0": S.Q8_0}[tname]
    rng = np.random.Generator(np.random.PCG64(41))
    for cols in (256, 768, 4096, 14336):
        for rows in (8, 11):
            blocks = S.random_blocks(rng, t, rows, cols)
            x = rng.standard_normal(cols).astype(np.float32)
            for ab in (16, 8):
                y1 = _run_r(hostcheck_lib, t, blocks, rows, cols, x, ab, 1)
                for rpi in (2, 4):
                    y2 = _run_r(hostcheck_lib, t, blocks, rows, cols, x, ab, rpi)
                    assert np.array_equal(y1, y2), (tname, cols, rows, ab, rpi)


@pytest.mark.parametrize("cols,expect", [(256, (1, 1)), (768, (1, 3)), (4096, (1, 16)), (14336, (4, 14)), (8192, (2, 16)),
                                          (28672, (7, 16)), (5632, (2, 11)), (4352, (0, 0)), (32768, (8, 16)), (300, (0, 0))])
def test_ksplit_table(hostcheck_lib, cols, expect):
    nks, nb = ctypes.c_int(), ctypes.c_int()
    rc = hostcheck_lib.hc_ksplit(cols, ctypes.byref(nks), ctypes.byref(nb))
    assert (nks.value, nb.value) == expect and (rc == 0) == (expect[0] > 0)


@pytest.mark.parametrize("tname", ["Q4_K", "Q6_K", "Q8_0"])
@pytest.mark.parametrize("cols", [256, 4096, 14336])
def test_engine_layout_round_trip(hostcheck_lib, tname, cols):
    """GGUF rows -> engine matrix ([tile][K-segment][row][segment], transposed Q6_K / Q8_0 segments) -> element access ==
    gguf dequantisation, bit for bit, for every tile size"""
    from oracle import gguf_synth as S, llama_oracle as O
    t = {"Q4_K": S.Q4_K, "Q6_K": S.Q6_K, "Q8_0": S.Q8_0}[tname]
    rng = np.random.Generator(np.random.PCG64(cols + 5))
    rows = 7
    blocks = S.random_blocks(rng, t, rows, cols)
    ref = O.dequantize(blocks, t, (rows, cols)).astype(np.float32)
    for tile_rows in (1, 2, 4):
        out = np.zeros((rows, cols), np.float32)
        rc = hostcheck_lib.hc_dequant_engine(t, blocks.ctypes.data_as(ctypes.c_void_p), rows, cols, tile_rows, out.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        assert np.array_equal(out, ref), tile_rows


# ---- batched decode GEMM on quantised weights: qtile packing + the dequantisation program of the kernel (qgemm_layout.h) ----------
@pytest.mark.parametrize("tname", ["Q4_K", "Q6_K"])
def test_qgemm_tile_program_reproduces_the_gguf_values(hostcheck_lib, tname):
    """128 GGUF super-blocks -> QG qtile (a lossless permutation of the same bytes) -> the kernel's per-thread fp16
    dequantisation into four 128-byte-swizzled operand tiles -> un-swizzled: equals the GGUF dequantisation (gguf-py-pinned
    oracle) to fp16 arithmetic: |got - exact| <= 2^-10 * (|d*sc*q| + |dmin*mn|) + one fp16 ulp of the result."""
    import ctypes
    from oracle import gguf_synth as S, llama_oracle as O
    t = getattr(S, tname)
    bb = S.BLOCK[t][1]
    lib = hostcheck_lib
    lib.hc_qg_dequant.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    for seed, mode in ((1, "random"), (2, "quantize")):
        rng = np.random.Generator(np.random.PCG64(seed))
        if mode == "random":
            blocks = S.random_blocks(rng, t, 128, 256)
        else:
            blocks = S.quantize(rng.standard_normal((128, 256), dtype=np.float32) * np.float32(0.02), t)
        blocks = np.ascontiguousarray(blocks).view(np.uint8).reshape(128, bb)
        out = np.zeros((128, 256), np.uint16)
        nbytes = ctypes.c_int(0)
        rc = lib.hc_qg_dequant(t, blocks.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nbytes))
        assert rc == 0 and nbytes.value == 128 * bb                   # the qtile is exactly the GGUF bytes of its blocks
        got = out.view(np.float16).astype(np.float64)
        exact = O.dequantize(blocks, t, (128, 256)).astype(np.float64)
        # magnitude of the terms that are rounded to fp16 before they are combined
        if t == S.Q4_K:
            d = blocks[:, 0:2].copy().view(np.float16).astype(np.float64)
            dmin = blocks[:, 2:4].copy().view(np.float16).astype(np.float64)
            mag = np.abs(d) * 63 * 15 + np.abs(dmin) * 63
        else:
            d = blocks[:, 208:210].copy().view(np.float16).astype(np.float64)
            mag = np.abs(d) * 128 * 32
        tol = 2.0 ** -10 * mag + np.abs(exact) * 2.0 ** -10 + 1e-12
        assert np.all(np.abs(got - exact) <= tol), (tname, mode, float(np.abs(got - exact).max()))
        # and it is not a loose match: the typical error is a fraction of an fp16 ulp of the weight
        nz = np.abs(exact) > 0
        assert np.median(np.abs(got - exact)[nz] / np.abs(exact)[nz]) < 2.0 ** -10
