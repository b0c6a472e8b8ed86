"""Decode-step time vs context length and KV split count (GPU box)."""
import json
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
    os.environ["GL_PREFILL"] = "1"
    for splits in sys.argv[1:] or ("4", "8", "16", "32"):
        os.environ["GL_ATTN_SPLITS"] = splits
        e = N.Engine(path, max_ctx=4096)
        row = {}
        for ctx in (1, 33, 129, 577, 1025, 2049, 4000):
            ms, _ = e.time_decode(ctx, 32)
            row[ctx] = round(ms, 4)
        print(json.dumps({"splits": int(splits), "ms": row}), flush=True)
        e.close()


if __name__ == "__main__":
    main()
