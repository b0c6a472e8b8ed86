"""CPU: the engine's byte-level BPE tokenizer (gridllm_b200/csrc/tokenizer.cpp) on the synthetic GPT-2 style
vocabulary, cross-checked against a from-scratch Python BPE that applies the same merges."""
import ctypes

import numpy as np
import pytest


def _py_bpe(word_u, ranks):
    sym = list(word_u)
    while len(sym) > 1:
        best, bi = None, -1
        for i in range(len(sym) - 1):
            r = ranks.get(sym[i] + " " + sym[i + 1])
            if r is not None and (best is None or r < best):
                best, bi = r, i
        if best is None:
            break
        sym[bi:bi + 2] = [sym[bi] + sym[bi + 1]]
    return sym


@pytest.mark.parametrize("text", ["hello world", "the rain in spain", " hello  world\n\nthe end", "café naïve 中文", "x=42; y=1234567",
                                  "it's we'll they're", ""])
def test_encode_decode_roundtrip_and_merges(hostcheck_lib, tiny_gguf, text):
    from oracle import gguf_synth as S
    toks, merges, types = S.synth_vocab(S.TINY.n_vocab)
    ids = np.zeros(4096, np.int32)
    n = hostcheck_lib.hc_tokenize(tiny_gguf.encode(), text.encode(), 1, 0, ids.ctypes.data_as(ctypes.c_void_p), 4096)
    assert n >= 1 and ids[0] == S.TINY.n_vocab - 3                  # BOS
    buf = ctypes.create_string_buffer(8192)
    m = hostcheck_lib.hc_detokenize(tiny_gguf.encode(), ids[1:].ctypes.data_as(ctypes.c_void_p), n - 1, buf, 8192)
    assert buf.raw[:m].decode("utf-8") == text                      # lossless
    # every produced token is a legal BPE symbol of its word under the file's merge table
    b2u = S.gpt2_byte_to_unicode()
    ranks = {mg: i for i, mg in enumerate(merges)}
    produced = [toks[i] for i in ids[1:n]]
    assert "".join(produced) == "".join(b2u[b] for b in text.encode())
    if text == "hello world":
        # "hello" is in the synthetic vocabulary as a whole token that no merge sequence reaches: llama-bpe files ignore the
        # merges for such pre-tokens; " world" goes through the merge loop
        assert produced == ["hello"] + _py_bpe("".join(b2u[b] for b in b" world"), ranks)
        assert len(produced) < len(text)                            # merges were applied


def test_special_tokens(hostcheck_lib, tiny_gguf):
    from oracle import gguf_synth as S
    ids = np.zeros(64, np.int32)
    n = hostcheck_lib.hc_tokenize(tiny_gguf.encode(), b"hi<|eot_id|>", 0, 1, ids.ctypes.data_as(ctypes.c_void_p), 64)
    assert ids[n - 1] == S.TINY.n_vocab - 1
    n2 = hostcheck_lib.hc_tokenize(tiny_gguf.encode(), b"hi<|eot_id|>", 0, 0, ids.ctypes.data_as(ctypes.c_void_p), 64)
    assert n2 > n                                                    # not parsed: spelled out byte by byte


# ---- the pre-tokeniser against an independent regex engine -----------------------------------------------------------
LLAMA3_SPLIT = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")

PRETOK_TEXTS = [
    "Hello  world's café 中文。12345 x=42\n\nend  ", "it's We'LL they'Re I'M he'D 'tis o'clock", "a\tb\t\tc \t d", "tabs\t\tand\r\nCRLF\r\n\r\nlines \n x",
    "naïve façade Ünïcödé straße ǅ", "日本語のテキスト、句読点。「引用」", "Привет, мир! Ελληνικά; עברית ، العربية", "१२३४५ ٣٤٥ Ⅻ ½ ²³ 1234567890",
    "no break em　ideographic  spaces line", "emoji 😀😃 mixed👍🏽text 🇩🇪", "x=y+z*(a/b)-[c]{d}<e>|f&g^h%i$j#k@l!m~n`o", "   leading and trailing   ",
    "...!!!???\n\n\n", "CamelCaseWordsAndsnake_case_words and kebab-case", "price: $1,234.56 (≈€1.100,00) 50% off!!", "\n", " ", "", "a", "'", "''s", "１２３ｆｕｌｌｗｉｄｔｈ",
    "mixed١٢٣abc४५६def", "it'ſ ok x'ſt we'LL I'D", "end with space ", "end with spaces  \t", "\t\tword", " \n \n x", "a  \n  b", "def f(x):\n    return x**2  # comment\n",
]


def _hc_pretokenize(lib, text):
    import ctypes
    raw = text.encode("utf-8")
    ends = np.zeros(len(raw) + 8, np.int32)
    n = lib.hc_pretokenize(raw, len(raw), ends.ctypes.data_as(ctypes.c_void_p), len(ends))
    assert n >= 0
    out, a = [], 0
    for e in ends[:n]:
        out.append(raw[a:int(e)].decode("utf-8"))
        a = int(e)
    assert a == len(raw)                      # the pieces tile the text
    return out


def test_pretokeniser_matches_the_tokenizers_regex_engine(hostcheck_lib):
    tokenizers = pytest.importorskip("tokenizers")
    split = tokenizers.pre_tokenizers.Split(tokenizers.Regex(LLAMA3_SPLIT), behavior="isolated", invert=False)
    for text in PRETOK_TEXTS:
        ref = [p for p, _ in split.pre_tokenize_str(text)]
        assert _hc_pretokenize(hostcheck_lib, text) == ref, text


def test_pretokeniser_random_strings(hostcheck_lib):
    """random strings over an alphabet that mixes every character class the expression distinguishes"""
    tokenizers = pytest.importorskip("tokenizers")
    split = tokenizers.pre_tokenizers.Split(tokenizers.Regex(LLAMA3_SPLIT), behavior="isolated", invert=False)
    alphabet = list("ab Z9 0'\n\r\t.,!-_世界éß٣½ 　😀") + ["'s", "'ll", "  ", "\n\n"]
    rng = np.random.Generator(np.random.PCG64(11))
    for _ in range(400):
        text = "".join(rng.choice(alphabet, size=int(rng.integers(1, 24))))
        ref = [p for p, _ in split.pre_tokenize_str(text)]
        assert _hc_pretokenize(hostcheck_lib, text) == ref, repr(text)


def test_full_tokenizer_matches_the_tokenizers_library(hostcheck_lib, tiny_gguf):
    """pre-tokeniser + byte-level mapping + BPE merges + ignore_merges against Hugging Face `tokenizers` built from the same
    vocabulary and merge table (the configuration of Llama-3's tokenizer.json: Split(regex) -> ByteLevel, BPE(ignore_merges))"""
    tokenizers = pytest.importorskip("tokenizers")
    from oracle import gguf_synth as S
    toks, merges, _ = S.synth_vocab(S.TINY.n_vocab)
    tk = tokenizers.Tokenizer(tokenizers.models.BPE(vocab={t: i for i, t in enumerate(toks)}, merges=[tuple(m.split(" ")) for m in merges],
                                                    ignore_merges=True))
    tk.pre_tokenizer = tokenizers.pre_tokenizers.Sequence([
        tokenizers.pre_tokenizers.Split(tokenizers.Regex(LLAMA3_SPLIT), behavior="isolated", invert=False),
        tokenizers.pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    rng = np.random.Generator(np.random.PCG64(3))
    alphabet = list("abcdefghijklmnopqrstuvwxyz     ehtoa.,\n0123'sé世")
    texts = PRETOK_TEXTS + ["hello world", "the rain in spain stays mainly in the plain"]
    texts += ["".join(rng.choice(alphabet, size=int(rng.integers(1, 60)))) for _ in range(300)]
    for text in texts:
        ids = np.zeros(2048, np.int32)
        n = hostcheck_lib.hc_tokenize(tiny_gguf.encode(), text.encode(), 0, 0, ids.ctypes.data_as(ctypes.c_void_p), 2048)
        assert list(ids[:n]) == tk.encode(text).ids, repr(text)


def test_pretokeniser_unicode_fuzz(hostcheck_lib):
    """random strings over thousands of code points from every plane in use, exotic spaces, case-folding specials: the tables
    in unicode_ranges.h are derived from the same engine, so any difference is a difference in the split logic"""
    tokenizers = pytest.importorskip("tokenizers")
    split = tokenizers.pre_tokenizers.Split(tokenizers.Regex(LLAMA3_SPLIT), behavior="isolated", invert=False)
    rng = np.random.Generator(np.random.PCG64(321))
    cands = [chr(cp) for cp in list(range(0x20, 0x7F)) + [9, 10, 11, 12, 13, 0x1C, 0x1F, 0x85, 0xA0, 0x1680, 0x180E, 0x2003, 0x200A, 0x2028, 0x2029, 0x202F,
                                                            0x205F, 0x3000, 0x200B, 0x200D, 0xFEFF, 0xAD, 0x301, 0x20E3, 0x1F600, 0x1F1E9, 0x1F3FD, 0x2764,
                                                            0xFE0F, 0x17F, 0x212A, 0x130]]
    while len(cands) < 4000:
        cp = int(rng.integers(0x80, 0x3FFFF))
        if not 0xD800 <= cp <= 0xDFFF:
            cands.append(chr(cp))
    for _ in range(4000):
        text = "".join(rng.choice(cands, size=int(rng.integers(1, 12))))
        ref = [p for p, _ in split.pre_tokenize_str(text)]
        assert _hc_pretokenize(hostcheck_lib, text) == ref, repr(text)


# ---- SentencePiece BPE (tokenizer.ggml.model == "llama": Llama-2 / Mistral, the model family of BASELINE config 5) -----------------
@pytest.fixture(scope="module")
def spm_vocab(tmp_path_factory):
    """A SentencePiece BPE model trained here on synthetic text (byte fallback, dummy prefix, identity normalisation, [INST] / [/INST]
    as user-defined symbols -- the shape of the Llama-2 / Mistral tokenizers) and the vocabulary-only GGUF a converter would write for
    it: pieces, scores, token types, bos / eos / unk ids, add_space_prefix."""
    spm = pytest.importorskip("sentencepiece")
    import io
    import random
    from oracle import gguf_synth as S
    rnd = random.Random(1)
    words = ["the", "quick", "brown", "fox", "jumps", "over", "lazy", "dog", "hello", "world", "rain", "in", "spain", "stays", "mainly", "plain",
             "naïve", "café", "日本語", "テスト", "emoji", "data", "model", "token", "embedding", "worker", "scheduler", "gpu", "kernel", "Straße",
             "über", "zwölf", "1234", "42", "3.14", "(paren)", "semi;colon", "new\nline"]
    lines = [" ".join(rnd.choice(words) for _ in range(rnd.randint(3, 12))) for _ in range(4000)]
    model = io.BytesIO()
    spm.SentencePieceTrainer.train(sentence_iterator=iter(lines), model_writer=model, vocab_size=600, model_type="bpe", byte_fallback=True,
                                   character_coverage=0.995, bos_id=1, eos_id=2, unk_id=0, pad_id=-1, normalization_rule_name="identity",
                                   add_dummy_prefix=True, remove_extra_whitespaces=False, minloglevel=2, user_defined_symbols=["[INST]", "[/INST]"])
    sp = spm.SentencePieceProcessor(model_proto=model.getvalue())
    n = sp.get_piece_size()
    toks = [sp.id_to_piece(i) for i in range(n)]
    types = [2 if sp.is_unknown(i) else 3 if sp.is_control(i) else 6 if sp.is_byte(i) else 4 if toks[i] in ("[INST]", "[/INST]") else 1 for i in range(n)]
    g = S.GGUFFile()
    g.add_str("general.architecture", "llama")
    g.add_str("tokenizer.ggml.model", "llama")
    g.add_arr("tokenizer.ggml.tokens", S._STR, toks)
    g.add_arr("tokenizer.ggml.scores", S._F32, [sp.get_score(i) for i in range(n)])
    g.add_arr("tokenizer.ggml.token_type", S._I32, types)
    g.add_u32("tokenizer.ggml.bos_token_id", 1)
    g.add_u32("tokenizer.ggml.eos_token_id", 2)
    g.add_u32("tokenizer.ggml.unknown_token_id", 0)
    g.add_bool("tokenizer.ggml.add_bos_token", True)
    g.add_bool("tokenizer.ggml.add_space_prefix", True)
    path = str(tmp_path_factory.mktemp("spm") / "spm_vocab.gguf")
    g.write(path)
    return path, sp


def _spm_ids(lib, path, text, bos=0, special=0):
    ids = np.zeros(8192, dtype=np.int32)
    n = lib.hc_tokenize(path.encode(), text.encode(), bos, special, ids.ctypes.data_as(ctypes.c_void_p), 8192)
    assert n >= 0, n
    return ids[:n].tolist()


def test_sentencepiece_bpe_matches_the_sentencepiece_library(hostcheck_lib, spm_vocab):
    """tokenizer.cpp's SentencePiece path (highest-score adjacent pair first, U+2581 for spaces, the dummy prefix, <0xXX> byte fallback)
    against the `sentencepiece` library on the same model: curated strings (multiple / leading / trailing spaces, accents, CJK, an
    emoji and letters outside the model's alphabet -> byte pieces, digits, punctuation, newlines, the empty string) and 400 random
    ones; ids equal, and the text comes back from the ids."""
    import random
    path, sp = spm_vocab
    tests = ["hello world", "the quick brown fox", " leading space", "two  spaces", "trailing ", "naïve café über Straße", "日本語のテスト",
             "emoji 😀 zzz qqq", "1234 42 3.14", "new\nline\ttab", "", "a", "zwölfzwölf hello,world!(paren)", "   ", "\n"]
    rnd = random.Random(7)
    alphabet = list("abcdefghijklmnopqrstuvwxyz    \n.,;!()äöüß日本語テスト1234567890😀")
    tests += ["".join(rnd.choice(alphabet) for _ in range(rnd.randint(1, 48))) for _ in range(400)]
    out = ctypes.create_string_buffer(1 << 16)
    for t in tests:
        ours, ref = _spm_ids(hostcheck_lib, path, t), sp.encode(t)
        assert ours == ref, (t, ours, ref)
        if t:
            a = np.asarray(ours, dtype=np.int32)
            k = hostcheck_lib.hc_detokenize(path.encode(), a.ctypes.data_as(ctypes.c_void_p), len(ours), out, 1 << 16)
            assert k >= 0 and out.raw[:k].decode("utf-8") == sp.decode(ref) == t
    assert _spm_ids(hostcheck_lib, path, "hello", bos=1) == [1] + sp.encode("hello")          # BOS id, then the prefixed text


def test_sentencepiece_control_and_user_defined_pieces(hostcheck_lib, spm_vocab):
    """Mistral's [INST] framing: with parse_special the markers are single pieces and the text after a marker gets the space prefix
    again (what llama.cpp's SPM path does [external]); without it they are spelled out piece by piece."""
    path, sp = spm_vocab
    inst, inst_end = sp.piece_to_id("[INST]"), sp.piece_to_id("[/INST]")
    ids = _spm_ids(hostcheck_lib, path, "[INST] hello world [/INST]", bos=1, special=1)
    assert ids[0] == 1 and ids[1] == inst and ids[-1] == inst_end
    assert ids[2:-1] == sp.encode(" hello world ")             # the fragment " hello world " behind a marker: prefixed again
    # user-defined pieces are whole pieces with or without parse_special; the text behind one is a new fragment and gets the space
    # prefix again -- llama.cpp's rule (Ollama's tokenizer), where sentencepiece itself would not re-prefix; control pieces need
    # parse_special
    assert _spm_ids(hostcheck_lib, path, "a [INST] b", bos=0, special=0) == sp.encode("a ") + [inst] + sp.encode(" b")
    eos_text = sp.id_to_piece(2)
    plain = _spm_ids(hostcheck_lib, path, "a" + eos_text, bos=0, special=0)
    assert 2 not in plain and _spm_ids(hostcheck_lib, path, "a" + eos_text, bos=0, special=1)[-1] == 2
