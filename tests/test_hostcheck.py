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


@pytest.mark.parametrize("tname", ["Q4_K", "Q6_K"])
def test_paired_row_functions_equal_single_row(hostcheck_lib, tname):
    """unit_dot2_* (two rows in lock-step, what the GPU consumer runs for row pairs) == unit_dot_* bit for bit"""
    from oracle import gguf_synth as S
    t = {"Q4_K": S.Q4_K, "Q6_K": S.Q6_K}[tname]
    rng = np.random.Generator(np.random.PCG64(31))
    for cols in (256, 4096, 14336):
        blocks = S.random_blocks(rng, t, 10, cols)
        x = rng.standard_normal(cols).astype(np.float32)
        for ab in (16, 8):
            y1 = _run(hostcheck_lib, t, blocks, 10, cols, x, ab)
            y2 = np.zeros(10, np.float32)
            rc = hostcheck_lib.hc_gemv_pairs(t, blocks.ctypes.data_as(ctypes.c_void_p), 10, cols, x.ctypes.data_as(ctypes.c_void_p),
                                             y2.ctypes.data_as(ctypes.c_void_p), ab)
            assert rc == 0 and np.array_equal(y1, y2)


@pytest.mark.parametrize("tname", ["Q4_K", "Q6_K", "Q8_0"])
def test_quad_row_functions_equal_single_row(hostcheck_lib, tname):
    """quad_dot_* (four rows per lane iteration, activations read from the swizzled shared-memory planes -- the
    production consumer loop) == unit_dot_* bit for bit, including ragged row counts"""
    from oracle import gguf_synth as S
    t = {"Q4_K": S.Q4_K, "Q6_K": S.Q6_K, "Q8_0": S.Q8_0}[tname]
    rng = np.random.Generator(np.random.PCG64(41))
    for cols in (256, 768, 4096, 14336):
        for rows in (8, 11):
            blocks = S.random_blocks(rng, t, rows, cols)
            x = rng.standard_normal(cols).astype(np.float32)
            for ab in (16, 8):
                y1 = _run(hostcheck_lib, t, blocks, rows, cols, x, ab)
                y2 = np.zeros(rows, np.float32)
                rc = hostcheck_lib.hc_gemv_quads(t, blocks.ctypes.data_as(ctypes.c_void_p), rows, cols, x.ctypes.data_as(ctypes.c_void_p),
                                                 y2.ctypes.data_as(ctypes.c_void_p), ab)
                assert rc == 0 and np.array_equal(y1, y2), (tname, cols, rows, ab)
