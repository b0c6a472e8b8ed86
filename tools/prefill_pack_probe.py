"""Profiling driver: packed prompt passes of the bench model -- `n_seq` prompts of `n_tok` tokens opened together
(gl_seq_open_many: <= 2048 rows per pass), the prefill half of BASELINE config 3 (32 x 512) and config 4 (2048-token prompts).
usage: prefill_pack_probe.py <n_seq> <n_tok> <iters>   -- prints ms per batch of prompts and prompt tokens/s (ignore under ncu)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from gridllm_b200 import native as N
    path = bench.build_model_once("llama3_8b_q4km", 0, lambda: None)
    n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    n_tok = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    e = N.Engine(path, max_ctx=n_tok + 64, max_batch=max(2, n_seq), batch_weights=2)
    rng = np.random.Generator(np.random.PCG64(5))
    prompts = [rng.integers(0, e.info.n_vocab - 3, size=n_tok) for _ in range(n_seq)]
    best = 1e9
    for it in range(iters + 1):
        e.batch_counters(reset=True)
        t0 = time.perf_counter()
        slots = e.seq_open_many(prompts, [dict(num_predict=2, ignore_eos=True)] * n_seq)
        wall = time.perf_counter() - t0
        c = e.batch_counters(reset=True)
        for s in slots:
            e.seq_close(s)
        if it:
            best = min(best, c["prefill_ns"] / 1e6)
    print({"n_seq": n_seq, "n_tok": n_tok, "prefill_ms_device": best, "prompt_tok_s": n_seq * n_tok / best * 1e3, "wall_ms_last": wall * 1e3}, flush=True)
    e.close()


if __name__ == "__main__":
    main()
