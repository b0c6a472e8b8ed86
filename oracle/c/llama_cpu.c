/*
 * TEST INFRASTRUCTURE / CPU BASELINE -- plain-C restatement of the decode step the native worker
 * replaces.  Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * What it restates: the per-token forward that, for the reference, runs inside the Ollama daemon
 * reached from OllamaService.generateResponse / generateStreamResponse
 * (/root/reference/client/src/services/OllamaService.ts:142-145, 235-237).  Ollama (llama.cpp/ggml)
 * is not vendored in /root/reference and is un-pinned (docs/deployment/docker-compose.dependencies.yml:14,
 * image ollama/ollama:latest), so this follows the published GGUF block formats and Llama
 * architecture, and is pinned against oracle/llama_oracle.py (tests/test_c_oracle.py), which is in
 * turn pinned against gguf-py and transformers.  PARITY UNPINNED w.r.t. the reference itself.
 *
 * act_mode 0: exact -- dequantised fp32 weights x fp32 activations ("mode A", the specification).
 * act_mode 1: ggml-style ("mode B") -- activations quantised to int8 the way ggml's CPU backend does
 *             it ([external] ggml vec_dot pairing: K-quant weights x Q8_K = one scale per 256 columns
 *             + per-16 sums; Q8_0 weights x Q8_0 = one scale per 32), integer dot products (AVX2
 *             maddubs/madd when available).  This is the fair stand-in for llama.cpp's CPU kernels and
 *             is what bench.py times as the CPU baseline (label: "CPU restatement, not Ollama").
 * K/V are rounded to fp16 when cached, like the engine.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#ifdef __AVX2__
#include <immintrin.h>
#endif

enum { T_F32 = 0, T_F16 = 1, T_Q8_0 = 8, T_Q4_K = 12, T_Q6_K = 14, T_BF16 = 30 };

typedef struct { const uint8_t* data; int type; int rows, cols; size_t row_bytes; } mat_t;
typedef struct { const float* attn_norm; const float* ffn_norm; mat_t wq, wk, wv, wo, wg, wu, wd; } layer_t;

typedef struct {
    void* map; size_t map_len;
    int n_layer, n_embd, n_head, n_kv, hd, n_ff, n_vocab, n_ctx;
    float eps, rope_base;
    mat_t tok_embd, output; const float* output_norm;
    layer_t* layers;
    float *kc, *vc;            /* [layer][pos][n_kv*hd], values already fp16-rounded */
    float *cos_t, *sin_t;      /* [n_ctx][hd/2] */
    int pos;
    float *x, *xn, *q, *k, *v, *att, *g, *u, *h, *y, *sc;
    int8_t* xq; float* xs;     /* int8 activations + per-block scale (per 32, or per 256 for K-quants) */
    int* bs;                   /* per-16 sums of xq (Q8_K bsums) */
} model_t;

static float h2f(uint16_t h) { _Float16 f; memcpy(&f, &h, 2); return (float)f; }
static float f16_round(float v) { _Float16 f = (_Float16)v; return (float)f; }

static size_t type_row_bytes(int t, int cols) {
    switch (t) {
        case T_F32: return (size_t)cols * 4;
        case T_F16: case T_BF16: return (size_t)cols * 2;
        case T_Q8_0: return (size_t)cols / 32 * 34;
        case T_Q4_K: return (size_t)cols / 256 * 144;
        case T_Q6_K: return (size_t)cols / 256 * 210;
        default: return 0;
    }
}

/* ---- GGUF v3 container (public layout) ------------------------------------------------------- */
typedef struct { const uint8_t* p; const uint8_t* end; int ok; } cur_t;
static uint64_t rd(cur_t* c, int n) { uint64_t v = 0; if (c->end - c->p < n) { c->ok = 0; return 0; } memcpy(&v, c->p, n); c->p += n; return v; }
static void rd_str(cur_t* c, char* out, int cap) {
    uint64_t n = rd(c, 8);
    if (!c->ok || (uint64_t)(c->end - c->p) < n) { c->ok = 0; return; }
    if (out) { int m = (int)(n < (uint64_t)cap - 1 ? n : (uint64_t)cap - 1); memcpy(out, c->p, m); out[m] = 0; }
    c->p += n;
}
static const int scalar_size[13] = {1, 1, 2, 2, 4, 4, 4, 1, 0, 0, 8, 8, 8};
static double rd_num(cur_t* c, uint32_t t) {
    switch (t) {
        case 0: return (double)(uint8_t)rd(c, 1);
        case 1: return (double)(int8_t)rd(c, 1);
        case 2: return (double)(uint16_t)rd(c, 2);
        case 3: return (double)(int16_t)rd(c, 2);
        case 4: return (double)(uint32_t)rd(c, 4);
        case 5: return (double)(int32_t)rd(c, 4);
        case 6: { uint32_t b = (uint32_t)rd(c, 4); float f; memcpy(&f, &b, 4); return f; }
        case 7: return (double)(uint8_t)rd(c, 1);
        case 10: return (double)rd(c, 8);
        case 11: return (double)(int64_t)rd(c, 8);
        case 12: { uint64_t b = rd(c, 8); double f; memcpy(&f, &b, 8); return f; }
        default: c->ok = 0; return 0;
    }
}

typedef struct { char name[128]; uint32_t type; int64_t ne[4]; int nd; uint64_t off; } tinfo_t;

static const tinfo_t* find_t(const tinfo_t* ti, uint64_t n, const char* name) {
    for (uint64_t i = 0; i < n; ++i) if (!strcmp(ti[i].name, name)) return &ti[i];
    return NULL;
}

static int bind_mat(mat_t* m, const tinfo_t* t, const uint8_t* data_base) {
    if (!t) return -1;
    m->type = (int)t->type; m->cols = (int)t->ne[0]; m->rows = t->nd > 1 ? (int)t->ne[1] : 1;
    m->row_bytes = type_row_bytes(m->type, m->cols);
    if (!m->row_bytes) return -1;
    m->data = data_base + t->off;
    return 0;
}

void oc_free(void* vm);

void* oc_load(const char* path, int n_ctx) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return NULL;
    struct stat st; fstat(fd, &st);
    void* map = mmap(NULL, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return NULL;
    model_t* m = calloc(1, sizeof(model_t));
    m->map = map; m->map_len = st.st_size;
    cur_t c = {(const uint8_t*)map, (const uint8_t*)map + st.st_size, 1};
    if ((uint32_t)rd(&c, 4) != 0x46554747u) { oc_free(m); return NULL; }
    rd(&c, 4);
    uint64_t nt = rd(&c, 8), nkv = rd(&c, 8);
    uint64_t align = 32;
    m->eps = 1e-5f; m->rope_base = 10000.f;
    int n_ctx_train = 2048, rope_dim = 0;
    char rope_scaling[32] = "";          /* rope.scaling.type: "", none, linear are restated; anything else refuses to load */
    float rope_lin = 1.f;
    for (uint64_t i = 0; i < nkv && c.ok; ++i) {
        char key[256]; rd_str(&c, key, sizeof key);
        uint32_t t = (uint32_t)rd(&c, 4);
        if (t == 8) {
            const char* ks = strchr(key, '.'); ks = ks ? ks + 1 : key;
            if (!strcmp(ks, "rope.scaling.type")) rd_str(&c, rope_scaling, sizeof rope_scaling);
            else rd_str(&c, NULL, 0);
            continue;
        }
        if (t == 9) {
            uint32_t et = (uint32_t)rd(&c, 4); uint64_t n = rd(&c, 8);
            if (et == 8) { for (uint64_t j = 0; j < n && c.ok; ++j) rd_str(&c, NULL, 0); }
            else if (et < 13 && scalar_size[et]) { c.p += n * scalar_size[et]; if (c.p > c.end) c.ok = 0; }
            else c.ok = 0;
            continue;
        }
        double v = rd_num(&c, t);
        const char* k = strchr(key, '.'); k = k ? k + 1 : key;
        if (!strcmp(key, "general.alignment")) align = (uint64_t)v;
        else if (!strcmp(k, "block_count")) m->n_layer = (int)v;
        else if (!strcmp(k, "embedding_length")) m->n_embd = (int)v;
        else if (!strcmp(k, "feed_forward_length")) m->n_ff = (int)v;
        else if (!strcmp(k, "attention.head_count")) m->n_head = (int)v;
        else if (!strcmp(k, "attention.head_count_kv")) m->n_kv = (int)v;
        else if (!strcmp(k, "attention.layer_norm_rms_epsilon")) m->eps = (float)v;
        else if (!strcmp(k, "rope.freq_base")) m->rope_base = (float)v;
        else if (!strcmp(k, "rope.dimension_count")) rope_dim = (int)v;
        else if (!strcmp(k, "rope.scaling.factor")) rope_lin = (float)v;
        else if (!strcmp(k, "context_length")) n_ctx_train = (int)v;
    }
    tinfo_t* ti = calloc(nt, sizeof(tinfo_t));
    for (uint64_t i = 0; i < nt && c.ok; ++i) {
        rd_str(&c, ti[i].name, sizeof ti[i].name);
        ti[i].nd = (int)rd(&c, 4);
        for (int d = 0; d < ti[i].nd && d < 4; ++d) ti[i].ne[d] = (int64_t)rd(&c, 8);
        ti[i].type = (uint32_t)rd(&c, 4);
        ti[i].off = rd(&c, 8);
    }
    if (!c.ok || !m->n_layer || !m->n_embd || !m->n_head) { free(ti); oc_free(m); return NULL; }
    if (rope_scaling[0] && strcmp(rope_scaling, "none") && strcmp(rope_scaling, "linear")) { free(ti); oc_free(m); return NULL; }
    if (strcmp(rope_scaling, "linear") || !(rope_lin > 0.f)) rope_lin = 1.f;
    if (!m->n_kv) m->n_kv = m->n_head;
    m->hd = rope_dim ? rope_dim : m->n_embd / m->n_head;
    m->n_ctx = n_ctx > 0 ? n_ctx : n_ctx_train;
    size_t off = (size_t)(c.p - (const uint8_t*)map);
    off = (off + align - 1) / align * align;
    const uint8_t* base = (const uint8_t*)map + off;
    int bad = 0;
    bad |= bind_mat(&m->tok_embd, find_t(ti, nt, "token_embd.weight"), base);
    const tinfo_t* to = find_t(ti, nt, "output.weight");
    bad |= bind_mat(&m->output, to ? to : find_t(ti, nt, "token_embd.weight"), base);
    const tinfo_t* on = find_t(ti, nt, "output_norm.weight");
    if (!on) bad = 1; else m->output_norm = (const float*)(base + on->off);
    m->n_vocab = m->tok_embd.rows;
    const tinfo_t* rf = find_t(ti, nt, "rope_freqs.weight");      /* Llama-3.1+: per-pair frequency factors, F32[hd/2] */
    const float* rope_ff = (rf && rf->type == T_F32) ? (const float*)(base + rf->off) : NULL;
    m->layers = calloc(m->n_layer, sizeof(layer_t));
    for (int il = 0; il < m->n_layer && !bad; ++il) {
        char nm[160];
        layer_t* L = &m->layers[il];
        struct { const char* n; mat_t* mm; } it[] = {{"attn_q", &L->wq}, {"attn_k", &L->wk}, {"attn_v", &L->wv}, {"attn_output", &L->wo},
                                                    {"ffn_gate", &L->wg}, {"ffn_up", &L->wu}, {"ffn_down", &L->wd}};
        for (int j = 0; j < 7; ++j) { snprintf(nm, sizeof nm, "blk.%d.%s.weight", il, it[j].n); bad |= bind_mat(it[j].mm, find_t(ti, nt, nm), base); }
        snprintf(nm, sizeof nm, "blk.%d.attn_norm.weight", il);
        const tinfo_t* a = find_t(ti, nt, nm);
        snprintf(nm, sizeof nm, "blk.%d.ffn_norm.weight", il);
        const tinfo_t* f = find_t(ti, nt, nm);
        if (!a || !f) bad = 1; else { L->attn_norm = (const float*)(base + a->off); L->ffn_norm = (const float*)(base + f->off); }
    }
    free(ti);
    if (bad) { oc_free(m); return NULL; }
    const int kvd = m->n_kv * m->hd, qd = m->n_head * m->hd;
    m->kc = malloc((size_t)m->n_layer * m->n_ctx * kvd * 4);
    m->vc = malloc((size_t)m->n_layer * m->n_ctx * kvd * 4);
    m->cos_t = malloc((size_t)m->n_ctx * m->hd / 2 * 4);
    m->sin_t = malloc((size_t)m->n_ctx * m->hd / 2 * 4);
    for (int p = 0; p < m->n_ctx; ++p)
        for (int i = 0; i < m->hd / 2; ++i) {
            float inv = (float)pow((double)m->rope_base, -2.0 * i / m->hd);
            if (rope_ff) inv = inv / rope_ff[i];
            float ang = ((float)p / rope_lin) * inv;
            m->cos_t[(size_t)p * m->hd / 2 + i] = (float)cos((double)ang);
            m->sin_t[(size_t)p * m->hd / 2 + i] = (float)sin((double)ang);
        }
    int big = m->n_ff > m->n_embd ? m->n_ff : m->n_embd;
    if (qd > big) big = qd;
    m->x = malloc(m->n_embd * 4); m->xn = malloc(big * 4); m->q = malloc(qd * 4); m->k = malloc(kvd * 4); m->v = malloc(kvd * 4);
    m->att = malloc(qd * 4); m->g = malloc(m->n_ff * 4); m->u = malloc(m->n_ff * 4); m->h = malloc(m->n_ff * 4); m->y = malloc(big * 4);
    m->sc = malloc((size_t)m->n_ctx * 4 * 64);
    m->xq = malloc(big + 64); m->xs = malloc((big / 32 + 1) * 4); m->bs = malloc((big / 16 + 1) * 4);
    return m;
}

void oc_free(void* vm) {
    model_t* m = vm;
    if (!m) return;
    if (m->map) munmap(m->map, m->map_len);
    free(m->layers); free(m->kc); free(m->vc); free(m->cos_t); free(m->sin_t);
    free(m->x); free(m->xn); free(m->q); free(m->k); free(m->v); free(m->att); free(m->g); free(m->u); free(m->h); free(m->y); free(m->sc);
    free(m->xq); free(m->xs); free(m->bs);
    free(m);
}

void oc_reset(void* vm) { ((model_t*)vm)->pos = 0; }
int oc_pos(void* vm) { return ((model_t*)vm)->pos; }
void oc_info(void* vm, int* o) {
    model_t* m = vm;
    o[0] = m->n_layer; o[1] = m->n_embd; o[2] = m->n_head; o[3] = m->n_kv; o[4] = m->hd; o[5] = m->n_ff; o[6] = m->n_vocab; o[7] = m->n_ctx;
}
int oc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py picks the thread count that is fastest on the box it runs on (all logical CPUs is rarely it) */
void oc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- block decoders (public ggml formats) ------------------------------------------------------ */
static void q4k_scale_min(const uint8_t* s, int j, int* sc, int* mn) {
    if (j < 4) { *sc = s[j] & 63; *mn = s[j + 4] & 63; }
    else { *sc = (s[j + 4] & 0xF) | ((s[j - 4] >> 6) << 4); *mn = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
}

static void dequant_row(const uint8_t* row, int type, int cols, float* out) {
    if (type == T_F32) { memcpy(out, row, (size_t)cols * 4); return; }
    if (type == T_F16) { for (int i = 0; i < cols; ++i) out[i] = h2f(((const uint16_t*)row)[i]); return; }
    if (type == T_BF16) { for (int i = 0; i < cols; ++i) { uint32_t b = (uint32_t)((const uint16_t*)row)[i] << 16; memcpy(&out[i], &b, 4); } return; }
    if (type == T_Q8_0) {
        for (int b = 0; b < cols / 32; ++b) {
            const uint8_t* p = row + (size_t)b * 34; float d = h2f(*(const uint16_t*)p);
            for (int i = 0; i < 32; ++i) out[b * 32 + i] = d * (float)(int8_t)p[2 + i];
        }
        return;
    }
    if (type == T_Q4_K) {
        for (int b = 0; b < cols / 256; ++b) {
            const uint8_t* p = row + (size_t)b * 144; float d = h2f(*(const uint16_t*)p), dm = h2f(*(const uint16_t*)(p + 2));
            for (int j = 0; j < 8; ++j) {
                int sc, mn; q4k_scale_min(p + 4, j, &sc, &mn);
                const uint8_t* qs = p + 16 + (j >> 1) * 32;
                for (int l = 0; l < 32; ++l) { int q = (j & 1) ? (qs[l] >> 4) : (qs[l] & 0xF); out[b * 256 + j * 32 + l] = d * (float)sc * (float)q - dm * (float)mn; }
            }
        }
        return;
    }
    if (type == T_Q6_K) {
        for (int b = 0; b < cols / 256; ++b) {
            const uint8_t* p = row + (size_t)b * 210; float d = h2f(*(const uint16_t*)(p + 208));
            for (int e = 0; e < 256; ++e) {
                int h = e >> 7, r = e & 127;
                int ql = (p[h * 64 + (r & 63)] >> (4 * (r >> 6))) & 0xF;
                int qh = (p[128 + h * 32 + (r & 31)] >> (2 * (r >> 5))) & 3;
                out[b * 256 + e] = d * (float)(int8_t)p[192 + (e >> 4)] * (float)((ql | (qh << 4)) - 32);
            }
        }
    }
}

/* exact mode: fp32 dot against the dequantised row (double accumulate) */
static float row_dot_exact(const uint8_t* row, int type, int cols, const float* x, float* tmp) {
    dequant_row(row, type, cols, tmp);
    double a = 0;
    for (int i = 0; i < cols; ++i) a += (double)tmp[i] * x[i];
    return (float)a;
}

/* ggml-style mode: integer dots against int8 activations.
 * K-quants: xq quantised per 256 columns (scale xs[b], per-16 sums bs[]); Q8_0: per 32 columns. */
#ifdef __AVX2__
static inline int hsum_i32(__m256i v) {
    __m128i s = _mm_add_epi32(_mm256_castsi256_si128(v), _mm256_extracti128_si256(v, 1));
    s = _mm_add_epi32(s, _mm_shuffle_epi32(s, 0x4E));
    s = _mm_add_epi32(s, _mm_shuffle_epi32(s, 0xB1));
    return _mm_cvtsi128_si32(s);
}
#endif

static float row_dot_q8(const uint8_t* row, int type, int cols, const int8_t* xq, const float* xs, const int* bs, const float* x) {
    float acc = 0.f;
    if (type == T_Q8_0) {
        for (int b = 0; b < cols / 32; ++b) {
            const uint8_t* p = row + (size_t)b * 34; const int8_t* w = (const int8_t*)(p + 2); const int8_t* a = xq + b * 32;
            int s;
#ifdef __AVX2__
            const __m256i wv = _mm256_loadu_si256((const __m256i*)w), av = _mm256_loadu_si256((const __m256i*)a);
            const __m256i pr = _mm256_maddubs_epi16(_mm256_sign_epi8(wv, wv), _mm256_sign_epi8(av, wv));
            s = hsum_i32(_mm256_madd_epi16(pr, _mm256_set1_epi16(1)));
#else
            s = 0; for (int i = 0; i < 32; ++i) s += (int)w[i] * (int)a[i];
#endif
            acc += h2f(*(const uint16_t*)p) * xs[b] * (float)s;
        }
    } else if (type == T_Q4_K) {
        for (int b = 0; b < cols / 256; ++b) {
            const uint8_t* p = row + (size_t)b * 144; const float d = h2f(*(const uint16_t*)p), dm = h2f(*(const uint16_t*)(p + 2));
            const int8_t* a = xq + b * 256; const int* bsum = bs + b * 16;
            int sumi, summ = 0;
            int sc[8], mn[8];
            for (int j = 0; j < 8; ++j) { q4k_scale_min(p + 4, j, &sc[j], &mn[j]); summ += mn[j] * (bsum[2 * j] + bsum[2 * j + 1]); }
#ifdef __AVX2__
            __m256i accv = _mm256_setzero_si256();
            const __m256i m4 = _mm256_set1_epi8(0x0F);
            for (int c = 0; c < 4; ++c) {
                const __m256i q = _mm256_loadu_si256((const __m256i*)(p + 16 + c * 32));
                const __m256i lo = _mm256_and_si256(q, m4), hi = _mm256_and_si256(_mm256_srli_epi16(q, 4), m4);
                const __m256i a0 = _mm256_loadu_si256((const __m256i*)(a + 64 * c)), a1 = _mm256_loadu_si256((const __m256i*)(a + 64 * c + 32));
                accv = _mm256_add_epi32(accv, _mm256_madd_epi16(_mm256_maddubs_epi16(lo, a0), _mm256_set1_epi16((short)sc[2 * c])));
                accv = _mm256_add_epi32(accv, _mm256_madd_epi16(_mm256_maddubs_epi16(hi, a1), _mm256_set1_epi16((short)sc[2 * c + 1])));
            }
            sumi = hsum_i32(accv);
#else
            sumi = 0;
            for (int c = 0; c < 4; ++c) {
                const uint8_t* qs = p + 16 + c * 32; const int8_t* a0 = a + 64 * c; const int8_t* a1 = a0 + 32;
                int s0 = 0, s1 = 0;
                for (int l = 0; l < 32; ++l) { s0 += (int)(qs[l] & 0xF) * a0[l]; s1 += (int)(qs[l] >> 4) * a1[l]; }
                sumi += sc[2 * c] * s0 + sc[2 * c + 1] * s1;
            }
#endif
            acc += xs[b] * (d * (float)sumi - dm * (float)summ);
        }
    } else if (type == T_Q6_K) {
        for (int b = 0; b < cols / 256; ++b) {
            const uint8_t* p = row + (size_t)b * 210; const float d = h2f(*(const uint16_t*)(p + 208));
            const int8_t* sc = (const int8_t*)(p + 192); const int8_t* a = xq + b * 256; const int* bsum = bs + b * 16;
            int sumi, off = 0;
            for (int g = 0; g < 16; ++g) off += (int)sc[g] * bsum[g];
#ifdef __AVX2__
            __m256i accv = _mm256_setzero_si256();
            const __m256i m4 = _mm256_set1_epi8(0x0F), m2 = _mm256_set1_epi8(0x03);
            for (int h = 0; h < 2; ++h) {
                const __m256i ql0 = _mm256_loadu_si256((const __m256i*)(p + h * 64)), ql1 = _mm256_loadu_si256((const __m256i*)(p + h * 64 + 32));
                const __m256i qh = _mm256_loadu_si256((const __m256i*)(p + 128 + h * 32));
                __m256i qv[4];
                qv[0] = _mm256_or_si256(_mm256_and_si256(ql0, m4), _mm256_slli_epi16(_mm256_and_si256(qh, m2), 4));
                qv[1] = _mm256_or_si256(_mm256_and_si256(ql1, m4), _mm256_slli_epi16(_mm256_and_si256(_mm256_srli_epi16(qh, 2), m2), 4));
                qv[2] = _mm256_or_si256(_mm256_and_si256(_mm256_srli_epi16(ql0, 4), m4), _mm256_slli_epi16(_mm256_and_si256(_mm256_srli_epi16(qh, 4), m2), 4));
                qv[3] = _mm256_or_si256(_mm256_and_si256(_mm256_srli_epi16(ql1, 4), m4), _mm256_slli_epi16(_mm256_and_si256(_mm256_srli_epi16(qh, 6), m2), 4));
                for (int t = 0; t < 4; ++t) {
                    const __m256i av = _mm256_loadu_si256((const __m256i*)(a + h * 128 + t * 32));
                    const short s0 = sc[h * 8 + 2 * t], s1 = sc[h * 8 + 2 * t + 1];
                    const __m256i scv = _mm256_set_m128i(_mm_set1_epi16(s1), _mm_set1_epi16(s0));
                    accv = _mm256_add_epi32(accv, _mm256_madd_epi16(_mm256_maddubs_epi16(qv[t], av), scv));
                }
            }
            sumi = hsum_i32(accv);
#else
            sumi = 0;
            for (int g = 0; g < 16; ++g) {
                int s = 0;
                for (int i = 0; i < 16; ++i) {
                    int e = g * 16 + i, h = e >> 7, r = e & 127;
                    int ql = (p[h * 64 + (r & 63)] >> (4 * (r >> 6))) & 0xF;
                    int qh = (p[128 + h * 32 + (r & 31)] >> (2 * (r >> 5))) & 3;
                    s += (ql | (qh << 4)) * (int)a[e];
                }
                sumi += (int)sc[g] * s;
            }
#endif
            acc += d * xs[b] * (float)(sumi - 32 * off);
        }
    } else {                                     /* fp weights: plain float dot */
        if (type == T_F32) { const float* w = (const float*)row; for (int i = 0; i < cols; ++i) acc += w[i] * x[i]; }
        else if (type == T_F16) { const uint16_t* w = (const uint16_t*)row; for (int i = 0; i < cols; ++i) acc += h2f(w[i]) * x[i]; }
        else { const uint16_t* w = (const uint16_t*)row; for (int i = 0; i < cols; ++i) { uint32_t bb = (uint32_t)w[i] << 16; float f; memcpy(&f, &bb, 4); acc += f * x[i]; } }
    }
    return acc;
}

/* blk = 32 (Q8_0 pairing) or 256 (Q8_K pairing): d = amax/127, q = rint(x/d); per-16 sums */
static void quant_act(model_t* m, const float* x, int n, int blk) {
    for (int b = 0; b < n / blk; ++b) {
        float amax = 0.f;
        for (int i = 0; i < blk; ++i) { float a = fabsf(x[b * blk + i]); if (a > amax) amax = a; }
        const float inv = amax > 0.f ? 127.0f / amax : 0.f;
        m->xs[b] = amax / 127.0f;
        for (int i = 0; i < blk; ++i) m->xq[b * blk + i] = (int8_t)lrintf(x[b * blk + i] * inv);
    }
    for (int g = 0; g < n / 16; ++g) { int s = 0; for (int i = 0; i < 16; ++i) s += m->xq[g * 16 + i]; m->bs[g] = s; }
}

static void matvec(model_t* m, const mat_t* w, const float* x, float* y, int mode) {
    if (mode == 1 && (w->type == T_Q4_K || w->type == T_Q6_K)) quant_act(m, x, w->cols, 256);
    else if (mode == 1 && w->type == T_Q8_0) quant_act(m, x, w->cols, 32);
#pragma omp parallel
    {
        float* tmp = mode == 0 ? malloc((size_t)w->cols * 4) : NULL;
#pragma omp for schedule(static)
        for (int r = 0; r < w->rows; ++r) {
            const uint8_t* row = w->data + (size_t)r * w->row_bytes;
            y[r] = mode == 0 ? row_dot_exact(row, w->type, w->cols, x, tmp) : row_dot_q8(row, w->type, w->cols, m->xq, m->xs, m->bs, x);
        }
        free(tmp);
    }
}

static void rmsnorm(const float* x, const float* w, int n, float eps, float* y) {
    double ss = 0; for (int i = 0; i < n; ++i) ss += (double)x[i] * x[i];
    const float r = (float)(1.0 / sqrt(ss / n + eps));
    for (int i = 0; i < n; ++i) y[i] = (x[i] * r) * w[i];
}

static void rope(float* v, int n_heads, int hd, const float* c, const float* s) {
    for (int h = 0; h < n_heads; ++h)
        for (int i = 0; i < hd / 2; ++i) {
            float a = v[h * hd + 2 * i], b = v[h * hd + 2 * i + 1];
            v[h * hd + 2 * i] = a * c[i] - b * s[i];
            v[h * hd + 2 * i + 1] = a * s[i] + b * c[i];
        }
}

/* one token through all layers; returns 0; logits (may be NULL) get n_vocab floats; hidden (may be
 * NULL) gets the final residual stream */
int oc_step(void* vm, int token, int mode, float* logits, float* hidden) {
    model_t* m = vm;
    if (token < 0 || token >= m->n_vocab || m->pos >= m->n_ctx) return -1;
    const int E = m->n_embd, H = m->n_head, KV = m->n_kv, hd = m->hd, kvd = KV * hd, grp = H / KV, pos = m->pos;
    dequant_row(m->tok_embd.data + (size_t)token * m->tok_embd.row_bytes, m->tok_embd.type, E, m->x);
    const float* c = m->cos_t + (size_t)pos * hd / 2; const float* s = m->sin_t + (size_t)pos * hd / 2;
    for (int il = 0; il < m->n_layer; ++il) {
        layer_t* L = &m->layers[il];
        rmsnorm(m->x, L->attn_norm, E, m->eps, m->xn);
        matvec(m, &L->wq, m->xn, m->q, mode);
        matvec(m, &L->wk, m->xn, m->k, mode);
        matvec(m, &L->wv, m->xn, m->v, mode);
        rope(m->q, H, hd, c, s);
        rope(m->k, KV, hd, c, s);
        float* kc = m->kc + ((size_t)il * m->n_ctx) * kvd; float* vc = m->vc + ((size_t)il * m->n_ctx) * kvd;
        for (int i = 0; i < kvd; ++i) { kc[(size_t)pos * kvd + i] = f16_round(m->k[i]); vc[(size_t)pos * kvd + i] = f16_round(m->v[i]); }
        const float scale = 1.0f / sqrtf((float)hd);
#pragma omp parallel for schedule(static)
        for (int h = 0; h < H; ++h) {
            const int kvh = h / grp;
            float* sc = m->sc + (size_t)h * m->n_ctx;
            float mx = -INFINITY;
            for (int p = 0; p <= pos; ++p) {
                const float* kr = kc + (size_t)p * kvd + kvh * hd; double a = 0;
                for (int d = 0; d < hd; ++d) a += (double)m->q[h * hd + d] * kr[d];
                sc[p] = (float)a * scale; if (sc[p] > mx) mx = sc[p];
            }
            double den = 0; for (int p = 0; p <= pos; ++p) { sc[p] = expf(sc[p] - mx); den += sc[p]; }
            for (int d = 0; d < hd; ++d) {
                double a = 0; for (int p = 0; p <= pos; ++p) a += (double)sc[p] * vc[(size_t)p * kvd + kvh * hd + d];
                m->att[h * hd + d] = (float)(a / den);
            }
        }
        matvec(m, &L->wo, m->att, m->y, mode);
        for (int i = 0; i < E; ++i) m->x[i] += m->y[i];
        rmsnorm(m->x, L->ffn_norm, E, m->eps, m->xn);
        matvec(m, &L->wg, m->xn, m->g, mode);
        matvec(m, &L->wu, m->xn, m->u, mode);
        for (int i = 0; i < m->n_ff; ++i) m->h[i] = (m->g[i] / (1.0f + expf(-m->g[i]))) * m->u[i];
        matvec(m, &L->wd, m->h, m->y, mode);
        for (int i = 0; i < E; ++i) m->x[i] += m->y[i];
    }
    if (hidden) memcpy(hidden, m->x, (size_t)E * 4);
    if (logits) {
        rmsnorm(m->x, m->output_norm, E, m->eps, m->xn);
        matvec(m, &m->output, m->xn, logits, mode);
    }
    m->pos++;
    return 0;
}


/* ---- batched prompt evaluation ("prefill"): T tokens per weight pass ----------------------------------------------------
 * What llama.cpp's CPU backend does with a prompt [external]: every weight row is decoded once and used for all T tokens of
 * the batch, so the prompt costs one pass over the weights (compute-bound) instead of T passes (bandwidth-bound).  The
 * arithmetic per (row, token) is exactly oc_step's mode-1 arithmetic (int8 activations, integer dots), so oc_prefill and
 * T calls of oc_step give the same numbers (tests/test_c_oracle.py).  bench.py times this for the reference arm's prompt
 * phase -- charging the CPU a token-by-token prefill would flatter the GPU. */
static void quant_act_into(const float* x, int n, int blk, int8_t* xq, float* xs, int* bs) {
    for (int b = 0; b < n / blk; ++b) {
        float amax = 0.f;
        for (int i = 0; i < blk; ++i) { float a = fabsf(x[b * blk + i]); if (a > amax) amax = a; }
        const float inv = amax > 0.f ? 127.0f / amax : 0.f;
        xs[b] = amax / 127.0f;
        for (int i = 0; i < blk; ++i) xq[b * blk + i] = (int8_t)lrintf(x[b * blk + i] * inv);
    }
    for (int g = 0; g < n / 16; ++g) { int s = 0; for (int i = 0; i < 16; ++i) s += xq[g * 16 + i]; bs[g] = s; }
}

/* Y[t][r] = W[r] . X[t]  for t < T (X: [T][cols], Y: [T][rows]) */
static void matmat(const mat_t* w, const float* X, int T, float* Y, int8_t* xq, float* xs, int* bs) {
    const int cols = w->cols, quant = (w->type == T_Q4_K || w->type == T_Q6_K || w->type == T_Q8_0);
    const int blk = w->type == T_Q8_0 ? 32 : 256, nxs = cols / 32 + 1, nbs = cols / 16 + 1;
    if (quant) {
#pragma omp parallel for schedule(static)
        for (int t = 0; t < T; ++t) quant_act_into(X + (size_t)t * cols, cols, blk, xq + (size_t)t * cols, xs + (size_t)t * nxs, bs + (size_t)t * nbs);
    }
#pragma omp parallel for schedule(static)
    for (int r = 0; r < w->rows; ++r) {
        const uint8_t* row = w->data + (size_t)r * w->row_bytes;
        for (int t = 0; t < T; ++t)
            Y[(size_t)t * w->rows + r] = row_dot_q8(row, w->type, cols, xq + (size_t)t * cols, xs + (size_t)t * nxs, bs + (size_t)t * nbs, X + (size_t)t * cols);
    }
}

/* n prompt tokens from the model's current position; logits (may be NULL) of the LAST one.  mode must be 1. */
int oc_prefill(void* vm, const int* ids, int n, int mode, float* logits) {
    model_t* m = vm;
    if (mode != 1 || n <= 0 || m->pos + n > m->n_ctx) return -1;
    const int E = m->n_embd, H = m->n_head, KV = m->n_kv, hd = m->hd, kvd = KV * hd, qd = H * hd, grp = H / KV, FF = m->n_ff, pos0 = m->pos;
    for (int t = 0; t < n; ++t) if (ids[t] < 0 || ids[t] >= m->n_vocab) return -1;
    int big = FF > E ? FF : E; if (qd > big) big = qd;
    float* X = malloc((size_t)n * E * 4); float* XN = malloc((size_t)n * big * 4); float* Q = malloc((size_t)n * qd * 4);
    float* K = malloc((size_t)n * kvd * 4); float* V = malloc((size_t)n * kvd * 4); float* A = malloc((size_t)n * qd * 4);
    float* G = malloc((size_t)n * FF * 4); float* U = malloc((size_t)n * FF * 4); float* Y = malloc((size_t)n * big * 4);
    int8_t* xq = malloc((size_t)n * big + 64); float* xs = malloc((size_t)n * (big / 32 + 1) * 4); int* bs = malloc((size_t)n * (big / 16 + 1) * 4);
    float* SC = malloc((size_t)H * m->n_ctx * 4);
    for (int t = 0; t < n; ++t) dequant_row(m->tok_embd.data + (size_t)ids[t] * m->tok_embd.row_bytes, m->tok_embd.type, E, X + (size_t)t * E);
    const float scale = 1.0f / sqrtf((float)hd);
    for (int il = 0; il < m->n_layer; ++il) {
        layer_t* L = &m->layers[il];
        for (int t = 0; t < n; ++t) rmsnorm(X + (size_t)t * E, L->attn_norm, E, m->eps, XN + (size_t)t * E);
        matmat(&L->wq, XN, n, Q, xq, xs, bs);
        matmat(&L->wk, XN, n, K, xq, xs, bs);
        matmat(&L->wv, XN, n, V, xq, xs, bs);
        float* kc = m->kc + ((size_t)il * m->n_ctx) * kvd; float* vc = m->vc + ((size_t)il * m->n_ctx) * kvd;
        for (int t = 0; t < n; ++t) {
            const int pos = pos0 + t;
            const float* c = m->cos_t + (size_t)pos * hd / 2; const float* s = m->sin_t + (size_t)pos * hd / 2;
            rope(Q + (size_t)t * qd, H, hd, c, s);
            rope(K + (size_t)t * kvd, KV, hd, c, s);
            for (int i = 0; i < kvd; ++i) { kc[(size_t)pos * kvd + i] = f16_round(K[(size_t)t * kvd + i]); vc[(size_t)pos * kvd + i] = f16_round(V[(size_t)t * kvd + i]); }
        }
        for (int t = 0; t < n; ++t) {
            const int pos = pos0 + t;
#pragma omp parallel for schedule(static)
            for (int h = 0; h < H; ++h) {
                const int kvh = h / grp;
                float* sc = SC + (size_t)h * m->n_ctx;
                const float* q = Q + (size_t)t * qd + h * hd;
                float mx = -INFINITY;
                for (int p = 0; p <= pos; ++p) {
                    const float* kr = kc + (size_t)p * kvd + kvh * hd; double a = 0;
                    for (int d = 0; d < hd; ++d) a += (double)q[d] * kr[d];
                    sc[p] = (float)a * scale; if (sc[p] > mx) mx = sc[p];
                }
                double den = 0; for (int p = 0; p <= pos; ++p) { sc[p] = expf(sc[p] - mx); den += sc[p]; }
                for (int d = 0; d < hd; ++d) {
                    double a = 0; for (int p = 0; p <= pos; ++p) a += (double)sc[p] * vc[(size_t)p * kvd + kvh * hd + d];
                    A[(size_t)t * qd + h * hd + d] = (float)(a / den);
                }
            }
        }
        matmat(&L->wo, A, n, Y, xq, xs, bs);
        for (size_t i = 0; i < (size_t)n * E; ++i) X[i] += Y[i];
        for (int t = 0; t < n; ++t) rmsnorm(X + (size_t)t * E, L->ffn_norm, E, m->eps, XN + (size_t)t * E);
        matmat(&L->wg, XN, n, G, xq, xs, bs);
        matmat(&L->wu, XN, n, U, xq, xs, bs);
        for (size_t i = 0; i < (size_t)n * FF; ++i) G[i] = (G[i] / (1.0f + expf(-G[i]))) * U[i];
        matmat(&L->wd, G, n, Y, xq, xs, bs);
        for (size_t i = 0; i < (size_t)n * E; ++i) X[i] += Y[i];
    }
    if (logits) {
        rmsnorm(X + (size_t)(n - 1) * E, m->output_norm, E, m->eps, m->xn);
        matvec(m, &m->output, m->xn, logits, 1);
    }
    m->pos += n;
    free(X); free(XN); free(Q); free(K); free(V); free(A); free(G); free(U); free(Y); free(xq); free(xs); free(bs); free(SC);
    return 0;
}

/* Synthetic KV prefix: positions [0, n_pos) of every layer get small pseudo-random fp16-rounded values and the model's
 * position is set to n_pos.  bench.py uses it so that a bounded CPU sample decodes at the same context length as the tail
 * of the real 512-in / 128-out request without paying for the prompt first (timing only: the values mean nothing). */
int oc_fill_kv(void* vm, int n_pos, unsigned seed) {
    model_t* m = vm;
    if (n_pos < 0 || n_pos > m->n_ctx) return -1;
    const size_t kvd = (size_t)m->n_kv * m->hd;
    unsigned long long st = 0x9E3779B97F4A7C15ull ^ seed;
    for (int il = 0; il < m->n_layer; ++il) {
        float* kc = m->kc + ((size_t)il * m->n_ctx) * kvd; float* vc = m->vc + ((size_t)il * m->n_ctx) * kvd;
        for (size_t i = 0; i < (size_t)n_pos * kvd; ++i) {
            st = st * 6364136223846793005ull + 1442695040888963407ull;
            const float a = (float)((int)((st >> 33) & 0xFFFF) - 32768) * (1.0f / 32768.0f);
            st = st * 6364136223846793005ull + 1442695040888963407ull;
            const float b = (float)((int)((st >> 33) & 0xFFFF) - 32768) * (1.0f / 32768.0f);
            kc[i] = f16_round(a); vc[i] = f16_round(b);
        }
    }
    m->pos = n_pos;
    return 0;
}

/* greedy generate; returns number generated.  logits_buf must hold n_vocab floats. */
int oc_generate(void* vm, const int* prompt, int n_prompt, int n_predict, int mode, int* out_ids, float* out_lp, float* logits_buf) {
    model_t* m = vm;
    m->pos = 0;
    for (int i = 0; i < n_prompt; ++i)
        if (oc_step(m, prompt[i], mode, i == n_prompt - 1 ? logits_buf : NULL, NULL)) return -1;
    for (int g = 0; g < n_predict; ++g) {
        int best = 0; for (int i = 1; i < m->n_vocab; ++i) if (logits_buf[i] > logits_buf[best]) best = i;
        double se = 0; for (int i = 0; i < m->n_vocab; ++i) se += exp((double)logits_buf[i] - logits_buf[best]);
        out_ids[g] = best; if (out_lp) out_lp[g] = (float)(-log(se));
        if (g + 1 < n_predict && oc_step(m, best, mode, logits_buf, NULL)) return g + 1;
    }
    return n_predict;
}
