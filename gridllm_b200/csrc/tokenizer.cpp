// Byte-level BPE (GPT-2 style) from GGUF metadata.  See tokenizer.h.
#include "tokenizer.h"
#include "unicode_ranges.h"

#include <algorithm>
#include <climits>
#include <queue>

namespace gl {

namespace {

std::string cp_to_utf8(uint32_t cp) {
    std::string s;
    if (cp < 0x80) s += (char)cp;
    else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
    else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
    return s;
}

// split a UTF-8 string into code-point substrings
std::vector<std::string> utf8_chars(const std::string& s) {
    std::vector<std::string> out;
    for (size_t i = 0; i < s.size();) {
        unsigned char c = (unsigned char)s[i];
        size_t n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
        if (i + n > s.size()) n = 1;
        out.push_back(s.substr(i, n));
        i += n;
    }
    return out;
}

uint32_t utf8_decode1(const std::string& ch) {
    unsigned char c = (unsigned char)ch[0];
    if (c < 0x80 || ch.size() == 1) return c;
    if (ch.size() == 2) return ((c & 0x1F) << 6) | (ch[1] & 0x3F);
    if (ch.size() == 3) return ((c & 0x0F) << 12) | ((ch[1] & 0x3F) << 6) | (ch[2] & 0x3F);
    return ((c & 0x07) << 18) | ((ch[1] & 0x3F) << 12) | ((ch[2] & 0x3F) << 6) | (ch[3] & 0x3F);
}

struct Cp { uint32_t cp; uint32_t off; };      // code point and its byte offset in the text

bool in_ranges(const uint32_t (*r)[2], int n, uint32_t cp) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        if (cp < r[mid][0]) hi = mid - 1;
        else if (cp > r[mid][1]) lo = mid + 1;
        else return true;
    }
    return false;
}
inline bool is_letter(uint32_t c) { return c < 0x80 ? ((c | 0x20) >= 'a' && (c | 0x20) <= 'z') : in_ranges(UNI_LETTER, UNI_LETTER_N, c); }
inline bool is_digit(uint32_t c) { return c < 0x80 ? (c >= '0' && c <= '9') : in_ranges(UNI_NUMBER, UNI_NUMBER_N, c); }
// \s of the pre-tokeniser's regular expression: the Unicode White_Space property
inline bool is_space(uint32_t c) {
    return (c >= 0x9 && c <= 0xD) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 ||
           c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
inline bool is_nl(uint32_t c) { return c == '\n' || c == '\r'; }

}  // namespace

// llama-bpe pre-tokeniser (the "llama3" split of llama.cpp / the Llama-3 tokenizer.json [external]) restated over code points:
//  (?i:'s|'t|'re|'ve|'m|'ll|'d) | [^\r\n\p{L}\p{N}]?\p{L}+ | \p{N}{1,3} | ' '?[^\s\p{L}\p{N}]+[\r\n]* | \s*[\r\n]+ | \s+(?!\S) | \s+
// Alternatives are tried in this order at every position, each greedy with the backtracking the expression implies.  Pinned
// against the Hugging Face `tokenizers` regex engine in tests/test_tokenizer.py (ASCII, accents, CJK, other-script digits,
// Unicode spaces, emoji).  The contraction alternative folds case for ASCII and U+017F.
std::vector<std::string> llama3_pretokenize(const std::string& t) {
    std::vector<Cp> u;
    u.reserve(t.size());
    for (size_t i = 0; i < t.size();) {
        const unsigned char c = (unsigned char)t[i];
        size_t n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
        if (i + n > t.size()) n = 1;
        uint32_t cp = c;
        if (n == 2) cp = ((c & 0x1F) << 6) | (t[i + 1] & 0x3F);
        else if (n == 3) cp = ((c & 0x0F) << 12) | ((t[i + 1] & 0x3F) << 6) | (t[i + 2] & 0x3F);
        else if (n == 4) cp = ((c & 0x07) << 18) | ((t[i + 1] & 0x3F) << 12) | ((t[i + 2] & 0x3F) << 6) | (t[i + 3] & 0x3F);
        else if (c >= 0x80) cp = 0xFFFD;            // stray byte: neither letter, number nor space
        u.push_back(Cp{cp, (uint32_t)i});
        i += n;
    }
    const size_t n = u.size();
    u.push_back(Cp{0, (uint32_t)t.size()});          // sentinel: byte offset of the end
    std::vector<std::string> out;
    auto emit = [&](size_t a, size_t b) { out.push_back(t.substr(u[a].off, u[b].off - u[a].off)); };
    // case folding of the contraction alternative: ASCII, plus U+017F (long s), which folds to 's' in Unicode-aware engines
    auto lower = [](uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32 : (c == 0x17F ? (uint32_t)'s' : c); };
    size_t i = 0;
    while (i < n) {
        const uint32_t c = u[i].cp;
        // contractions
        if (c == '\'' && i + 1 < n) {
            const uint32_t a = lower(u[i + 1].cp);
            const uint32_t b = i + 2 < n ? lower(u[i + 2].cp) : 0;
            size_t len = 0;
            if (a == 's' || a == 't' || a == 'm' || a == 'd') len = 2;
            if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e') || (a == 'l' && b == 'l')) len = 3;
            if (len) { emit(i, i + len); i += len; continue; }
        }
        // [^\r\n L N]? L+
        {
            size_t j = i;
            if (!is_nl(c) && !is_letter(c) && !is_digit(c) && j + 1 < n && is_letter(u[j + 1].cp)) ++j;
            if (is_letter(u[j].cp) && j < n) {
                size_t k = j;
                while (k < n && is_letter(u[k].cp)) ++k;
                emit(i, k);
                i = k;
                continue;
            }
        }
        // N{1,3}
        if (is_digit(c)) {
            size_t k = i;
            while (k < n && k - i < 3 && is_digit(u[k].cp)) ++k;
            emit(i, k);
            i = k;
            continue;
        }
        // ' '? [^\s L N]+ [\r\n]*
        {
            size_t j = i;
            if (c == ' ' && j + 1 < n) ++j;
            auto punct = [&](size_t p) { const uint32_t d = u[p].cp; return !is_space(d) && !is_letter(d) && !is_digit(d); };
            if (j < n && punct(j)) {
                size_t k = j;
                while (k < n && punct(k)) ++k;
                while (k < n && is_nl(u[k].cp)) ++k;
                emit(i, k);
                i = k;
                continue;
            }
        }
        // whitespace runs
        if (is_space(c)) {
            size_t k = i;
            while (k < n && is_space(u[k].cp)) ++k;
            // \s*[\r\n]+ : through the last newline of the run
            size_t last_nl = std::string::npos;
            for (size_t p = i; p < k; ++p) if (is_nl(u[p].cp)) last_nl = p;
            if (last_nl != std::string::npos) { emit(i, last_nl + 1); i = last_nl + 1; continue; }
            // \s+(?!\S): leave the last space character to the next piece if something follows
            if (k < n && k - i > 1) { emit(i, k - 1); i = k - 1; continue; }
            // \s+ : a single space character before something that did not take it, or the run at the end of the text
            emit(i, k);
            i = k;
            continue;
        }
        emit(i, i + 1);
        ++i;
    }
    return out;
}

bool Tokenizer::load(const GGUFFile& f) {
    ok_ = false;
    const GGUFValue* toks = f.find("tokenizer.ggml.tokens");
    if (!toks || toks->arr_s.empty()) return false;
    const std::string model = f.get_s("tokenizer.ggml.model", "");
    if (model == "llama") {
        // SentencePiece BPE (Llama-2 / Mistral): pieces + scores; no pre-tokeniser, no merges list
        const GGUFValue* sc = f.find("tokenizer.ggml.scores");
        if (!sc || sc->arr_f.size() != toks->arr_s.size()) return false;
        tokens_ = toks->arr_s;
        scores_.assign(sc->arr_f.begin(), sc->arr_f.end());
        const GGUFValue* ty = f.find("tokenizer.ggml.token_type");
        types_.assign(tokens_.size(), 1);
        if (ty && ty->arr_i.size() == tokens_.size())
            for (size_t i = 0; i < tokens_.size(); ++i) types_[i] = (int)ty->arr_i[i];
        tok2id_.reserve(tokens_.size() * 2);
        for (size_t i = 0; i < tokens_.size(); ++i) tok2id_.emplace(tokens_[i], (int32_t)i);
        for (int b = 0; b < 256; ++b) byte_tok_[b] = -1;
        for (size_t i = 0; i < tokens_.size(); ++i) {
            const std::string& t = tokens_[i];
            if (types_[i] == 6 && t.size() == 6 && t.compare(0, 3, "<0x") == 0 && t[5] == '>') {
                auto hex = [](char c) { return c >= '0' && c <= '9' ? c - '0' : (c >= 'A' && c <= 'F' ? c - 'A' + 10 : (c >= 'a' && c <= 'f' ? c - 'a' + 10 : -1)); };
                const int hi = hex(t[3]), lo = hex(t[4]);
                if (hi >= 0 && lo >= 0) byte_tok_[hi * 16 + lo] = (int32_t)i;
            }
        }
        bos = (int)f.get_u("tokenizer.ggml.bos_token_id", 1);
        eos = (int)f.get_u("tokenizer.ggml.eos_token_id", 2);
        unk_ = (int)f.get_u("tokenizer.ggml.unknown_token_id", 0);
        eot = (int)f.get_u("tokenizer.ggml.eot_token_id", (uint64_t)-1);
        add_bos_default = f.get_u("tokenizer.ggml.add_bos_token", 1) != 0;
        add_space_prefix_ = f.get_u("tokenizer.ggml.add_space_prefix", 1) != 0;
        chat_template = f.get_s("tokenizer.chat_template", "");
        for (size_t i = 0; i < tokens_.size(); ++i)
            if (types_[i] == 3 || types_[i] == 4) specials_.emplace_back(tokens_[i], (int32_t)i);
        std::sort(specials_.begin(), specials_.end(), [](auto& a, auto& b) { return a.first.size() > b.first.size(); });
        spm_ = true;
        ok_ = true;
        return true;
    }
    if (model != "gpt2") return false;
    // Only the llama-bpe ("llama3") pre-tokeniser split is restated here.  Other byte-level BPE files that are also
    // general.architecture == llama (tekken, smollm, deepseek-llm, ... -- each has its own split regex in llama.cpp [external])
    // would tokenise silently wrong, so they load WITHOUT a tokenizer: text requests fail with a message, token-id requests work.
    {
        const std::string pre0 = f.get_s("tokenizer.ggml.pre", "");
        if (pre0 != "llama-bpe" && pre0 != "llama3" && pre0 != "llama-v3") return false;
    }
    tokens_ = toks->arr_s;
    const GGUFValue* ty = f.find("tokenizer.ggml.token_type");
    types_.assign(tokens_.size(), 1);
    if (ty && ty->arr_i.size() == tokens_.size())
        for (size_t i = 0; i < tokens_.size(); ++i) types_[i] = (int)ty->arr_i[i];
    tok2id_.reserve(tokens_.size() * 2);
    for (size_t i = 0; i < tokens_.size(); ++i) tok2id_.emplace(tokens_[i], (int32_t)i);
    const GGUFValue* mg = f.find("tokenizer.ggml.merges");
    if (mg)
        for (size_t i = 0; i < mg->arr_s.size(); ++i) merge_rank_.emplace(mg->arr_s[i], (int)i);
    bos = (int)f.get_u("tokenizer.ggml.bos_token_id", (uint64_t)-1);
    eos = (int)f.get_u("tokenizer.ggml.eos_token_id", (uint64_t)-1);
    eot = (int)f.get_u("tokenizer.ggml.eot_token_id", (uint64_t)-1);
    add_bos_default = f.get_u("tokenizer.ggml.add_bos_token", 1) != 0;
    // Llama-3's tokenizer.json sets ignore_merges; llama.cpp derives it from the pre-tokeniser name [external]
    const std::string pre = f.get_s("tokenizer.ggml.pre", "");
    ignore_merges_ = pre == "llama-bpe" || pre == "llama3" || pre == "llama-v3";
    chat_template = f.get_s("tokenizer.chat_template", "");
    // GPT-2 byte <-> unicode table
    std::vector<int> bs;
    for (int b = 33; b <= 126; ++b) bs.push_back(b);
    for (int b = 161; b <= 172; ++b) bs.push_back(b);
    for (int b = 174; b <= 255; ++b) bs.push_back(b);
    std::vector<int> cs(bs.begin(), bs.end());
    int extra = 0;
    for (int b = 0; b < 256; ++b)
        if (std::find(bs.begin(), bs.end(), b) == bs.end()) { bs.push_back(b); cs.push_back(256 + extra++); }
    for (size_t i = 0; i < bs.size(); ++i) {
        byte2u_[bs[i]] = cp_to_utf8((uint32_t)cs[i]);
        cp2byte_[(uint32_t)cs[i]] = (uint8_t)bs[i];
    }
    for (size_t i = 0; i < tokens_.size(); ++i)
        if (types_[i] == 3 || types_[i] == 4) specials_.emplace_back(tokens_[i], (int32_t)i);
    std::sort(specials_.begin(), specials_.end(), [](auto& a, auto& b) { return a.first.size() > b.first.size(); });
    ok_ = true;
    return true;
}

void Tokenizer::bpe_word(const std::string& word_u, std::vector<int32_t>& out) const {
    if (ignore_merges_) {      // a pre-token that is itself in the vocabulary is emitted whole (llama-bpe; see load())
        auto it = tok2id_.find(word_u);
        if (it != tok2id_.end()) { out.push_back(it->second); return; }
    }
    std::vector<std::string> sym = utf8_chars(word_u);
    while (sym.size() > 1) {
        int best = INT_MAX;
        size_t bi = 0;
        for (size_t i = 0; i + 1 < sym.size(); ++i) {
            auto it = merge_rank_.find(sym[i] + " " + sym[i + 1]);
            if (it != merge_rank_.end() && it->second < best) { best = it->second; bi = i; }
        }
        if (best == INT_MAX) break;
        sym[bi] += sym[bi + 1];
        sym.erase(sym.begin() + (long)bi + 1);
    }
    for (auto& s : sym) {
        auto it = tok2id_.find(s);
        if (it != tok2id_.end()) { out.push_back(it->second); continue; }
        for (auto& ch : utf8_chars(s)) {      // unknown merge result: fall back to byte tokens
            auto jt = tok2id_.find(ch);
            if (jt != tok2id_.end()) out.push_back(jt->second);
        }
    }
}

std::vector<int32_t> Tokenizer::encode(const std::string& text, bool add_bos, bool parse_special) const {
    std::vector<int32_t> out;
    if (!ok_) return out;
    if (add_bos && bos >= 0) out.push_back(bos);
    // split on control tokens first
    std::vector<std::pair<std::string, int32_t>> parts;   // (text, -1) or ("", special id)
    // control pieces (token type 3) are recognised in the text only with parse_special; user-defined pieces (type 4) always, as whole
    // pieces -- the rule of llama.cpp's special-token partition and of sentencepiece's user_defined_symbols [external]
    if (!specials_.empty()) {
        size_t i = 0, start = 0;
        while (i < text.size()) {
            bool hit = false;
            for (auto& sp : specials_) {
                if (!parse_special && types_[sp.second] == 3) continue;
                if (!sp.first.empty() && text.compare(i, sp.first.size(), sp.first) == 0) {
                    if (i > start) parts.emplace_back(text.substr(start, i - start), -1);
                    parts.emplace_back(std::string(), sp.second);
                    i += sp.first.size();
                    start = i;
                    hit = true;
                    break;
                }
            }
            if (!hit) ++i;
        }
        if (start < text.size()) parts.emplace_back(text.substr(start), -1);
    } else {
        parts.emplace_back(text, -1);
    }
    if (spm_) {
        // a fragment that starts the text or follows a control token is prefixed with a space (what llama.cpp's SPM path and
        // sentencepiece's add_dummy_prefix do [external]); spaces become U+2581 inside spm_text
        bool prev_special = true;
        for (auto& pr : parts) {
            if (pr.second >= 0) { out.push_back(pr.second); prev_special = true; continue; }
            if (pr.first.empty()) continue;                          // the empty text has no pieces (not even the prefix)
            spm_text((add_space_prefix_ && prev_special) ? " " + pr.first : pr.first, out);
            prev_special = false;
        }
        return out;
    }
    for (auto& pr : parts) {
        if (pr.second >= 0) { out.push_back(pr.second); continue; }
        for (auto& w : llama3_pretokenize(pr.first)) {
            std::string wu;
            for (unsigned char c : w) wu += byte2u_[c];
            bpe_word(wu, out);
        }
    }
    return out;
}

// SentencePiece BPE over one fragment (restating llm_tokenizer_spm of llama.cpp / the BPE model of sentencepiece [external]): the
// fragment's UTF-8 characters are the initial symbols; the adjacent pair whose concatenation is a piece with the HIGHEST score is merged
// first (ties: the leftmost pair), until no adjacent pair forms a piece; a symbol that is not a piece falls back to its bytes (<0xXX>).
void Tokenizer::spm_text(const std::string& raw, std::vector<int32_t>& out) const {
    std::string text;
    text.reserve(raw.size() + 8);
    for (char c : raw) {
        if (c == ' ') text += "\xE2\x96\x81";      // U+2581
        else text += c;
    }
    struct Sym { int prev, next; size_t off, n; };
    std::vector<Sym> sym;
    for (size_t i = 0; i < text.size();) {
        const unsigned char c = (unsigned char)text[i];
        size_t n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
        if (i + n > text.size()) n = text.size() - i;
        sym.push_back(Sym{(int)sym.size() - 1, (int)sym.size() + 1, i, n});
        i += n;
    }
    if (sym.empty()) return;
    sym.back().next = -1;
    struct Bigram { int left, right; float score; size_t size; };
    auto worse = [](const Bigram& a, const Bigram& b) { return a.score < b.score || (a.score == b.score && a.left > b.left); };
    std::priority_queue<Bigram, std::vector<Bigram>, decltype(worse)> queue(worse);
    auto try_add = [&](int l, int r) {
        if (l < 0 || r < 0) return;
        const std::string t = text.substr(sym[l].off, sym[l].n + sym[r].n);
        auto it = tok2id_.find(t);
        if (it == tok2id_.end() || (size_t)it->second >= scores_.size()) return;
        queue.push(Bigram{l, r, scores_[it->second], t.size()});
    };
    for (int i = 1; i < (int)sym.size(); ++i) try_add(i - 1, i);
    while (!queue.empty()) {
        const Bigram b = queue.top();
        queue.pop();
        Sym& l = sym[b.left];
        Sym& r = sym[b.right];
        if (l.n == 0 || r.n == 0 || l.n + r.n != b.size) continue;      // one of the two was merged away since this entry was queued
        l.n += r.n;
        r.n = 0;
        l.next = r.next;
        if (r.next >= 0) sym[r.next].prev = b.left;
        try_add(l.prev, b.left);
        try_add(b.left, l.next);
    }
    for (int i = 0; i >= 0; i = sym[i].next) {
        const std::string t = text.substr(sym[i].off, sym[i].n);
        auto it = tok2id_.find(t);
        if (it != tok2id_.end()) { out.push_back(it->second); continue; }
        for (unsigned char c : t) {
            if (byte_tok_[c] >= 0) out.push_back(byte_tok_[c]);
            else if (unk_ >= 0) out.push_back(unk_);
        }
    }
}

std::string Tokenizer::piece(int32_t id) const {
    if (!ok_ || id < 0 || id >= (int)tokens_.size()) return {};
    if (types_[id] == 3) return {};     // control tokens render as nothing
    if (spm_) {
        if (types_[id] == 6) {          // byte-fallback piece: the byte itself (maybe half a character: the host holds it back)
            for (int b = 0; b < 256; ++b)
                if (byte_tok_[b] == id) return std::string(1, (char)b);
            return {};
        }
        std::string out;
        const std::string& t = tokens_[id];
        for (size_t i = 0; i < t.size();) {
            if (t.compare(i, 3, "\xE2\x96\x81") == 0) { out += ' '; i += 3; }
            else out += t[i++];
        }
        return out;
    }
    std::string out;
    for (auto& ch : utf8_chars(tokens_[id])) {
        auto it = cp2byte_.find(utf8_decode1(ch));
        if (it != cp2byte_.end()) out += (char)it->second;
        else out += ch;
    }
    return out;
}

std::string Tokenizer::text(int32_t id) const {
    if (!ok_ || id < 0 || id >= (int)tokens_.size()) return {};
    if (types_[id] == 3 || types_[id] == 4) return tokens_[id];      // control / user-defined pieces are stored as they are written
    return piece(id);
}

std::string Tokenizer::decode(const int32_t* ids, int n) const {
    std::string out;
    for (int i = 0; i < n; ++i) out += piece(ids[i]);
    // SPM: the space the encoder put in front of a text that starts the sequence is not part of the text
    if (spm_ && add_space_prefix_ && !out.empty() && out[0] == ' ') out.erase(0, 1);
    return out;
}

}  // namespace gl
