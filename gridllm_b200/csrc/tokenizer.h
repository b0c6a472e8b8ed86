// Byte-level BPE tokenizer driven by GGUF metadata (tokenizer.ggml.{model,tokens,merges,...}).
// In the reference the tokenizer lives inside Ollama; OllamaService only ever ships prompt TEXT
// (/root/reference/client/src/services/OllamaService.ts:101-104, 190-195), so a native worker needs
// its own.  Supports tokenizer.ggml.model == "gpt2" (Llama-3 family: byte-level BPE; the pre-tokeniser is the
// llama-bpe split over Unicode code points -- general categories L* / N* from generated tables, tools/gen_unicode_ranges.py,
// and White_Space) and tokenizer.ggml.model == "llama" (Llama-2 / Mistral family, BASELINE config 5's model: SentencePiece BPE
// -- pieces with scores, U+2581 for spaces, <0xXX> byte fallback).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "gguf_file.h"

namespace gl {

// the pre-tokeniser alone (pieces in text order; their concatenation is the text)
std::vector<std::string> llama3_pretokenize(const std::string& text);

class Tokenizer {
public:
    bool load(const GGUFFile& f);
    bool ok() const { return ok_; }
    std::vector<int32_t> encode(const std::string& text, bool add_bos, bool parse_special) const;
    std::string decode(const int32_t* ids, int n) const;
    std::string piece(int32_t id) const;     // raw bytes of one token ("" for control tokens)
    std::string text(int32_t id) const;      // the vocabulary's spelling of a token, control tokens included (chat templates)
    int n_vocab() const { return (int)tokens_.size(); }
    int bos = -1, eos = -1, eot = -1;
    bool add_bos_default = true;
    std::string chat_template;

private:
    void bpe_word(const std::string& word_u, std::vector<int32_t>& out) const;
    void spm_text(const std::string& text, std::vector<int32_t>& out) const;      // one raw-text fragment, SentencePiece BPE
    bool ok_ = false;
    bool spm_ = false;                                // tokenizer.ggml.model == "llama"
    bool add_space_prefix_ = true;                    // SPM: a fragment that starts the text or follows a control token gets a leading space
    int unk_ = -1;
    std::vector<float> scores_;                       // SPM: merge priority of every piece
    int32_t byte_tok_[256];                           // SPM: <0xXX> byte-fallback pieces (-1: absent)
    bool ignore_merges_ = false;                      // pre-tokens found whole in the vocabulary skip the merge loop
    std::vector<std::string> tokens_;                 // in "unicode-escaped bytes" form (GPT-2)
    std::vector<int> types_;
    std::unordered_map<std::string, int32_t> tok2id_;
    std::unordered_map<std::string, int> merge_rank_;  // "a b" -> rank
    std::vector<std::pair<std::string, int32_t>> specials_;  // control tokens, longest first
    std::string byte2u_[256];                         // byte -> UTF-8 of its GPT-2 code point
    std::unordered_map<uint32_t, uint8_t> cp2byte_;
};

}  // namespace gl
