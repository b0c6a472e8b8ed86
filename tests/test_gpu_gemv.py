"""GPU parity of the decode GEMV (through the C ABI, gl_gemv) against the oracle.

Tolerances: the kernel consumes activations snapped to 15-bit fixed point
per 32-column block, so against the oracle fed the SAME snapped activations only fp32
accumulation order differs -> rel-L2 <= 2e-5; against the exact-activation oracle (mode A) the
stated tolerance is rel-L2 <= 2e-4."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

SHAPES = [  # (rows, cols)
    (64, 256), (130, 512), (48, 768), (256, 1024), (40, 2048), (1024, 4096), (300, 5632), (296, 14336), (2, 4096), (1, 256),
    (9001, 4096), (5003, 14336),      # several rounds of items per CTA; ragged last items; 4 K-segments
]


@pytest.fixture(scope="module")
def engine(tiny_gguf):
    from gridllm_b200 import native as N
    e = N.Engine(tiny_gguf)
    yield e
    e.close()


@pytest.mark.parametrize("tname", ["Q4_K", "Q6_K", "Q8_0"])
@pytest.mark.parametrize("rows,cols", SHAPES)
def test_gemv_matches_oracle(engine, tname, rows, cols):
    from oracle import gguf_synth as S, llama_oracle as O
    t = {"Q4_K": S.Q4_K, "Q6_K": S.Q6_K, "Q8_0": S.Q8_0}[tname]
    rng = np.random.Generator(np.random.PCG64(1234 + rows * 7 + cols))
    blocks = S.random_blocks(rng, t, rows, cols)
    x = np.random.Generator(np.random.PCG64(42)).standard_normal(cols).astype(np.float32)
    y, _ = engine.gemv(t, blocks, rows, cols, x)
    wd = O.dequantize(blocks, t, (rows, cols))
    ref_snap = O.gemv(wd, x, "i16")
    ref_exact = O.gemv(wd, x, "exact")
    assert np.isfinite(y).all()
    assert rel_l2(y, ref_snap) <= 2e-5, (tname, rows, cols)
    assert rel_l2(y, ref_exact) <= 2e-4, (tname, rows, cols)


def test_gemv_quantised_master_weights(engine):
    """valid-block quantiser output (not raw random blocks), Llama attn shape"""
    from oracle import gguf_synth as S, llama_oracle as O
    rng = np.random.Generator(np.random.PCG64(5))
    w = rng.standard_normal((512, 4096), dtype=np.float32) / 64
    x = rng.standard_normal(4096).astype(np.float32)
    for t in (S.Q4_K, S.Q6_K, S.Q8_0):
        blocks = S.quantize(w, t)
        y, _ = engine.gemv(t, blocks, 512, 4096, x)
        assert rel_l2(y, O.gemv(O.dequantize(blocks, t, (512, 4096)), x, "i16")) <= 2e-5


def test_gemv_linearity_and_zero(engine):
    """size-independent properties: W(a x1 + b x2) = a W x1 + b W x2 within activation rounding; W 0 = 0"""
    from oracle import gguf_synth as S
    rng = np.random.Generator(np.random.PCG64(8))
    blocks = S.random_blocks(rng, S.Q4_K, 2048, 4096)
    x1 = rng.standard_normal(4096).astype(np.float32)
    x2 = rng.standard_normal(4096).astype(np.float32)
    y1, _ = engine.gemv(S.Q4_K, blocks, 2048, 4096, x1)
    y2, _ = engine.gemv(S.Q4_K, blocks, 2048, 4096, x2)
    y3, _ = engine.gemv(S.Q4_K, blocks, 2048, 4096, 2.0 * x1 - 3.0 * x2)
    assert rel_l2(y3, 2.0 * y1 - 3.0 * y2) <= 5e-4
    y0, _ = engine.gemv(S.Q4_K, blocks, 2048, 4096, np.zeros(4096, np.float32))
    assert np.all(y0 == 0)
    # power-of-two scaling of x is exact in the fixed point (scale-only change)
    y4, _ = engine.gemv(S.Q4_K, blocks, 2048, 4096, 4.0 * x1)
    assert np.array_equal(y4, 4.0 * y1)


def test_rmsnorm(engine):
    from oracle import llama_oracle as O
    rng = np.random.Generator(np.random.PCG64(3))
    x = rng.standard_normal(4096).astype(np.float32) * 3
    w = (1 + 0.1 * rng.standard_normal(4096)).astype(np.float32)
    y = engine.rmsnorm(x, w, 1e-5)
    assert rel_l2(y, O.rmsnorm(x, w, 1e-5)) <= 1e-6


def test_gemv_rejects_bad_shapes(engine):
    from gridllm_b200 import native as N
    from oracle import gguf_synth as S
    with pytest.raises(N.NativeError):
        engine.gemv(S.Q8_0, np.zeros((4, 34 * 3), np.uint8), 4, 96, np.zeros(96, np.float32))   # cols % 128 != 0
    with pytest.raises(N.NativeError):
        engine.gemv(99, np.zeros((4, 64), np.uint8), 4, 256, np.zeros(256, np.float32))
