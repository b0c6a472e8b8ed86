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
