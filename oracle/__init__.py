"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GridLLM native-worker hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import, link or execute it, and only as the checker / the CPU baseline.

PARITY UNPINNED: the reference (GridLLM, TypeScript) holds no golden vector, KAT or
fixture for this path (its one integration script,
tests/integration/integration.ts:6-35, compares JSON key sets and typeof only), and the
arithmetic lives in an un-vendored, un-pinned third party (Ollama -> llama.cpp/ggml,
docs/deployment/docker-compose.dependencies.yml:14 ``ollama/ollama:latest``).  The oracle
therefore restates the *published* GGUF/ggml block formats and the Llama architecture and
is pinned against the independent implementations that exist in this image:
``gguf.quants`` (gguf-py, the ggml project's own Python dequantisers) for the block
formats and ``transformers.LlamaForCausalLM`` for the forward pass
(tests/test_oracle_pin.py).
"""
