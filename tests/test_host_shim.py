"""CPU: the N-API shim a GridLLM maintainer would build (host/napi/addon.cc) type-checks against the C ABI header and a
declarations-only stand-in for <node_api.h> (node is absent from the build image; tests/hostcheck/napi_stub/)."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_napi_shim_type_checks():
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "hostcheck", "napi_stub"),
                        os.path.join(ROOT, "host", "napi", "addon.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_binds_the_hot_path_entry_points():
    src = open(os.path.join(ROOT, "host", "napi", "addon.cc")).read()
    for sym in ("gl_engine_create", "gl_engine_destroy", "gl_engine_info", "gl_tokenize", "gl_detokenize", "gl_generate", "gl_embed",
                "gl_device_count", "gl_last_error", "gl_chat_template", "gl_seq_open", "gl_batch_step", "gl_seq_close", "gl_seq_stats", "gl_token_piece", "gl_token_text"):
        assert sym + "(" in src, sym
    ts = open(os.path.join(ROOT, "host", "src", "NativeInferenceService.ts")).read()
    for method in ("checkHealth", "getAvailableModels", "validateModel", "generateResponse", "generateStreamResponse", "generateChatResponse",
                   "generateChatStreamResponse", "generateEmbedding"):        # the 8-method surface of OllamaService
        assert method + "(" in ts, method
    # the twin takes the message framing from the GGUF's template (not a hard-wired family) and round-trips the whole conversation
    assert "native.chatTemplate(" in ts and "<|im_start|>" in ts and "[INST]" in ts
    assert "...Array.from(promptIds ?? [])" in ts and "token_ids" in ts
