// extern "C" surface of libgridllm_native.so -- see include/gridllm_native.h for the reference
// call site each entry point stands in for.
#include <cuda_runtime.h>

#include <cstring>
#include <new>
#include <stdexcept>
#include <string>

#include "../../include/gridllm_native.h"
#include "engine.h"

using gl::Engine;
using gl::Status;

struct gl_engine {
    Engine* impl;
};

namespace {
int ret(const Status& s) {
    if (!s.ok()) gl::set_last_error(s.msg);
    return s.code;
}
int bad(const char* m) {
    gl::set_last_error(m);
    return GL_ERR_INVALID;
}
}  // namespace

extern "C" {

int gl_abi_version(void) { return GL_ABI_VERSION; }
const char* gl_last_error(void) { return gl::get_last_error(); }

int gl_device_count(int* n) {
    if (!n) return bad("gl_device_count: null");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        gl::set_last_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
        return GL_ERR_NO_DEVICE;
    }
    *n = c;
    return GL_OK;
}

int gl_engine_create(const char* gguf_path, int device, const gl_engine_opts* opts, gl_engine** out) {
    if (!gguf_path || !out) return bad("gl_engine_create: null argument");
    *out = nullptr;
    // nothing may unwind through the extern "C" boundary: a damaged file that makes the loader run out of memory or throw
    // is an error code, not std::terminate
    try {
        Engine* e = nullptr;
        Status s = Engine::create(gguf_path, device, opts, &e);
        if (!s.ok()) return ret(s);
        *out = new gl_engine{e};
        return GL_OK;
    } catch (const std::bad_alloc&) {
        gl::set_last_error("gl_engine_create: out of host memory while loading (damaged or oversized GGUF?)");
        return GL_ERR_NOMEM;
    } catch (const std::exception& ex) {
        gl::set_last_error(std::string("gl_engine_create: ") + ex.what());
        return GL_ERR_FORMAT;
    }
}

void gl_engine_destroy(gl_engine* e) {
    if (!e) return;
    delete e->impl;
    delete e;
}

int gl_engine_info(const gl_engine* e, gl_model_info* out) {
    if (!e || !out) return bad("gl_engine_info: null argument");
    return ret(e->impl->info(out));
}

int gl_tokenize(const gl_engine* e, const char* utf8, int32_t n_bytes, int add_bos, int parse_special, int32_t* ids, int32_t cap,
                int32_t* n_out) {
    if (!e || !utf8 || !n_out) return bad("gl_tokenize: null argument");
    const gl::Tokenizer& t = e->impl->tokenizer();
    if (!t.ok()) { gl::set_last_error("model carries no supported tokenizer (tokenizer.ggml.model != gpt2)"); return GL_ERR_UNSUPPORTED; }
    std::string text(utf8, n_bytes >= 0 ? (size_t)n_bytes : std::strlen(utf8));
    std::vector<int32_t> v = t.encode(text, add_bos != 0, parse_special != 0);
    *n_out = (int32_t)v.size();
    if ((int32_t)v.size() > cap || (!ids && !v.empty())) { gl::set_last_error("gl_tokenize: output buffer too small"); return ids ? GL_ERR_INVALID : GL_OK; }
    if (!v.empty()) std::memcpy(ids, v.data(), v.size() * sizeof(int32_t));
    return GL_OK;
}

int gl_detokenize(const gl_engine* e, const int32_t* ids, int32_t n, char* buf, int32_t cap, int32_t* len_out) {
    if (!e || (!ids && n > 0) || !len_out) return bad("gl_detokenize: null argument");
    const gl::Tokenizer& t = e->impl->tokenizer();
    if (!t.ok()) { gl::set_last_error("model carries no supported tokenizer"); return GL_ERR_UNSUPPORTED; }
    std::string s = t.decode(ids, n);
    *len_out = (int32_t)s.size();
    if ((int32_t)s.size() > cap || !buf) { gl::set_last_error("gl_detokenize: output buffer too small"); return buf ? GL_ERR_INVALID : GL_OK; }
    std::memcpy(buf, s.data(), s.size());
    return GL_OK;
}

int gl_chat_template(const gl_engine* e, char* buf, int32_t cap, int32_t* len_out) {
    if (!e || !len_out) return bad("gl_chat_template: null argument");
    const std::string& s = e->impl->tokenizer().chat_template;      // "" when the file carries none
    *len_out = (int32_t)s.size();
    if (!buf) return GL_OK;                                          // size query
    if ((int32_t)s.size() > cap) { gl::set_last_error("gl_chat_template: output buffer too small"); return GL_ERR_INVALID; }
    std::memcpy(buf, s.data(), s.size());
    return GL_OK;
}

int gl_generate(gl_engine* e, const int32_t* prompt, int32_t n_prompt, const gl_sample_opts* opts, gl_token_cb cb, void* user,
                int32_t* out_ids, float* out_logprobs, gl_gen_stats* stats) {
    if (!e || !prompt) return bad("gl_generate: null argument");
    gl_sample_opts so{};
    if (opts) so = *opts;
    else { so.num_predict = 128; so.top_p = 1.f; }
    return ret(e->impl->generate(prompt, n_prompt, so, cb, user, out_ids, out_logprobs, stats));
}

int gl_embed(gl_engine* e, const int32_t* ids, const int32_t* seq_offsets, int32_t n_seq, float* out, gl_gen_stats* stats) {
    if (!e || !ids || !seq_offsets || !out || n_seq <= 0) return bad("gl_embed: bad argument");
    return ret(e->impl->embed(ids, seq_offsets, n_seq, out, stats));
}

// ---- continuous batching -------------------------------------------------------------------------
int gl_seq_open(gl_engine* e, const int32_t* prompt, int32_t n_prompt, const gl_sample_opts* opts, int32_t* slot) {
    if (!e || !prompt || !slot) return bad("gl_seq_open: null argument");
    gl_sample_opts so{};
    if (opts) so = *opts;
    else { so.num_predict = 128; so.top_p = 1.f; }
    int s = -1;
    const int rc = ret(e->impl->seq_open(prompt, n_prompt, so, &s));
    if (rc == GL_OK) *slot = s;
    return rc;
}

int gl_seq_open_many(gl_engine* e, const int32_t* ids, const int32_t* offsets, int32_t n_seq, const gl_sample_opts* opts, int32_t* slots,
                     int32_t* n_opened) {
    if (!e || !ids || !offsets || !opts || !slots || !n_opened) return bad("gl_seq_open_many: null argument");
    int k = 0;
    const int rc = ret(e->impl->seq_open_many(ids, offsets, n_seq, opts, slots, &k));
    *n_opened = k;
    return rc;
}

int gl_batch_step(gl_engine* e, int32_t* slots, int32_t* ids, float* logprobs, int32_t* done, int32_t cap, int32_t* n) {
    if (!e || !n) return bad("gl_batch_step: null argument");
    int k = 0;
    const int rc = ret(e->impl->batch_step(slots, ids, logprobs, done, cap, &k));
    *n = k;
    return rc;
}

int gl_seq_close(gl_engine* e, int32_t slot) {
    if (!e) return bad("gl_seq_close: null engine");
    return ret(e->impl->seq_close(slot));
}

int gl_seq_logits(gl_engine* e, int32_t slot, float* out, int32_t n_vocab) {
    if (!e || !out) return bad("gl_seq_logits: null argument");
    return ret(e->impl->seq_logits(slot, out, n_vocab));
}

int gl_seq_stats(gl_engine* e, int32_t slot, gl_gen_stats* stats) {
    if (!e || !stats) return bad("gl_seq_stats: null argument");
    return ret(e->impl->seq_stats(slot, stats));
}

int gl_token_piece(const gl_engine* e, int32_t id, char* buf, int32_t cap, int32_t* len_out) {
    if (!e || !len_out) return bad("gl_token_piece: null argument");
    const gl::Tokenizer& t = e->impl->tokenizer();
    if (!t.ok()) { *len_out = 0; return GL_OK; }                    // no tokenizer: no piece, like the gl_generate callback
    const std::string pc = t.piece(id);
    *len_out = (int32_t)pc.size();
    if (!buf) return GL_OK;
    if ((int32_t)pc.size() > cap) { gl::set_last_error("gl_token_piece: output buffer too small"); return GL_ERR_INVALID; }
    std::memcpy(buf, pc.data(), pc.size());
    return GL_OK;
}

int gl_token_text(const gl_engine* e, int32_t id, char* buf, int32_t cap, int32_t* len_out) {
    if (!e || !len_out) return bad("gl_token_text: null argument");
    const gl::Tokenizer& t = e->impl->tokenizer();
    if (!t.ok()) { *len_out = 0; return GL_OK; }
    const std::string tx = t.text(id);
    *len_out = (int32_t)tx.size();
    if (!buf) return GL_OK;
    if ((int32_t)tx.size() > cap) { gl::set_last_error("gl_token_text: output buffer too small"); return GL_ERR_INVALID; }
    std::memcpy(buf, tx.data(), tx.size());
    return GL_OK;
}

int gl_batch_counters(gl_engine* e, uint64_t out[8], int32_t reset) {
    if (!e || !out) return bad("gl_batch_counters: null argument");
    e->impl->batch_counters(out, reset != 0);
    return GL_OK;
}

int gl_time_batch_step(gl_engine* e, int32_t batch, int32_t ctx_len, int32_t iters, float* ms_per_step, int32_t* launches_per_step,
                       uint64_t* weight_bytes) {
    if (!e) return bad("gl_time_batch_step: null engine");
    return ret(e->impl->time_batch_step(batch, ctx_len, iters, ms_per_step, launches_per_step, weight_bytes));
}

int gl_last_logits(gl_engine* e, int32_t step, float* out, int32_t n_vocab) {
    if (!e || !out) return bad("gl_last_logits: null argument");
    return ret(e->impl->last_logits(step, out, n_vocab));
}

int gl_sample_logits(gl_engine* e, const float* logits, int32_t n_vocab, const gl_sample_opts* opts, int32_t out_index, int32_t* id,
                     float* logprob) {
    if (!e || !logits || !opts) return bad("gl_sample_logits: null argument");
    int tid = 0;
    const int rc = ret(e->impl->sample_logits(logits, n_vocab, *opts, out_index, &tid, logprob));
    if (rc == GL_OK && id) *id = tid;
    return rc;
}

int gl_gemv(gl_engine* e, int ggml_type, const void* w_host, int32_t rows, int32_t cols, const float* x, float* y, int32_t iters,
            float* kernel_ms) {
    if (!e) return bad("gl_gemv: null engine");
    return ret(e->impl->gemv_host(ggml_type, w_host, rows, cols, x, y, iters, kernel_ms));
}

int gl_gemv_model_tensor(gl_engine* e, const char* tensor_name, const float* x, float* y, int32_t iters, int32_t flush_l2,
                         float* kernel_ms, uint64_t* weight_bytes) {
    if (!e || !tensor_name || !x || !y) return bad("gl_gemv_model_tensor: null argument");
    return ret(e->impl->gemv_tensor(tensor_name, x, y, iters, flush_l2, kernel_ms, weight_bytes));
}

int gl_rmsnorm(gl_engine* e, const float* x, const float* w, int32_t n, float eps, float* y) {
    if (!e || !x || !w || !y || n <= 0) return bad("gl_rmsnorm: bad argument");
    return ret(e->impl->rmsnorm(x, w, n, eps, y));
}

int gl_decode_step(gl_engine* e, int32_t token, float* logits, int32_t* argmax, float* logprob) {
    if (!e) return bad("gl_decode_step: null engine");
    return ret(e->impl->decode_step(token, logits, argmax, logprob));
}

int gl_kv_reset(gl_engine* e) {
    if (!e) return bad("gl_kv_reset: null engine");
    return ret(e->impl->kv_reset());
}

int gl_position(const gl_engine* e, int32_t* pos) {
    if (!e || !pos) return bad("gl_position: null argument");
    *pos = e->impl->position();
    return GL_OK;
}

int gl_prefill(gl_engine* e, const int32_t* ids, int32_t n, float* last_logits) {
    if (!e || !ids) return bad("gl_prefill: null argument");
    return ret(e->impl->prefill(ids, n, last_logits));
}

int gl_time_decode(gl_engine* e, int32_t ctx_len, int32_t iters, float* ms_per_step, int32_t* launches_per_step) {
    if (!e) return bad("gl_time_decode: null engine");
    return ret(e->impl->time_decode(ctx_len, iters, ms_per_step, launches_per_step));
}

// profiling aid (not part of the reference-facing ABI): per-phase globaltimer stamps of the persistent kernel
int gl_debug_mega_trace(gl_engine* e, unsigned long long* out, int32_t cap, int32_t* n_ctas, int32_t* n_phases) {
    if (!e || !out) return bad("gl_debug_mega_trace: null argument");
    return ret(e->impl->mega_trace(out, cap, n_ctas, n_phases));
}

// profiling aid: %globaltimer stamps of every GEMV / attention launch of the last decode step (GL_TRACE=1)
int gl_debug_perop_trace(gl_engine* e, unsigned long long* out, int32_t cap, int32_t* n_launches) {
    if (!e || !out || !n_launches) return bad("gl_debug_perop_trace: null argument");
    return ret(e->impl->perop_trace(out, cap, n_launches));
}

}  // extern "C"
