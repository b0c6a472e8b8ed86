"""Profiling driver (run under ncu on the GPU box): a few decode steps of the Llama-3-8B-shaped model."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gguf_synth as S  # noqa: E402


def main():
    from gridllm_b200 import native as N
    path = "/dev/shm/prof_llama3_8b.gguf"
    if not os.path.exists(path):
        S.build_model(path, S.LLAMA3_8B, "q4_k_m", seed=1234, mode="random", with_vocab=False)
    os.environ.setdefault("GL_PREFILL", "1")      # no 16 GB fp16 copy needed to look at decode
    e = N.Engine(path, max_ctx=1024)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    ms, nl = e.time_decode(int(sys.argv[2]) if len(sys.argv) > 2 else 576, steps)
    print("ms_per_token", ms, "launches", nl, flush=True)
    e.close()


if __name__ == "__main__":
    main()
