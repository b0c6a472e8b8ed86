// TEST INFRASTRUCTURE: runs the GEMV lane program (gridllm_b200/csrc/rowdot.h) on the CPU so the
// layout / bit-twiddling logic can be checked against the oracle without a GPU.  Never linked
// into libgridllm_native.so.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../gridllm_b200/csrc/rowdot.h"
#include "../../gridllm_b200/csrc/gguf_file.h"

using namespace gl;

template <int AB>
static void build_xunits(const float* x, int cols, std::vector<XUnit>& xs) {
    int nu = cols / UNIT_COLS;
    xs.resize(nu);
    for (int u = 0; u < nu; ++u) {
        XUnit& xu = xs[u];
        for (int b = 0; b < 4; ++b) {
            const float* xb = x + u * 128 + b * 32;
            float amax = 0.f;
            for (int i = 0; i < 32; ++i) amax = std::fmax(amax, std::fabs(xb[i]));
            int s0, s1;
            snap16<AB>(xb, amax, &xu.hi[8 * b], &xu.lo[8 * b], &s0);
            snap16<AB>(xb + 16, amax, &xu.hi[8 * b + 4], &xu.lo[8 * b + 4], &s1);
            xu.sx[b] = amax / (AB == 16 ? ACT16_RANGE : ACT8_RANGE);
            xu.sm[b] = xu.sx[b] * (float)(s0 + s1);
            xu.s16[2 * b] = s0;
            xu.s16[2 * b + 1] = s1;
        }
    }
}

// rows taken two at a time through the lock-step functions the GPU consumer uses for row pairs
template <int AB>
static int run_pairs(int type, const uint8_t* w, int rows, int cols, const float* x, float* y) {
    std::vector<XUnit> xs;
    build_xunits<AB>(x, cols, xs);
    int nu = cols / UNIT_COLS;
    size_t rb = row_bytes(type, cols), rs = align16(rb);
    std::vector<uint8_t> buf(2 * rs + 16);
    uint8_t* r0 = buf.data() + ((16 - ((uintptr_t)buf.data() & 15)) & 15);
    uint8_t* r1 = r0 + rs;
    for (int i = 0; i + 1 < rows; i += 2) {
        float acc0 = 0.f, acc1 = 0.f;
        if (type == T_Q4_K) {
            memcpy(r0, w + (size_t)i * rb, rb); memcpy(r1, w + (size_t)(i + 1) * rb, rb);
            for (int u = 0; u < nu; ++u) { float a, b; unit_dot2_q4k<AB>(r0 + (size_t)(u >> 1) * 144, r1 + (size_t)(u >> 1) * 144, u & 1, xs[u], a, b); acc0 += a; acc1 += b; }
        } else if (type == T_Q6_K) {
            repack_row_q6k(w + (size_t)i * rb, r0, cols / 256); repack_row_q6k(w + (size_t)(i + 1) * rb, r1, cols / 256);
            for (int u = 0; u < nu; ++u) { float a, b; unit_dot2_q6k<AB>(r0, r1, cols / 256, u, xs[u], a, b); acc0 += a; acc1 += b; }
        } else return -1;
        y[i] = acc0; y[i + 1] = acc1;
    }
    return 0;
}

template <int AB>
static int run(int type, const uint8_t* w, int rows, int cols, const float* x, float* y) {
    std::vector<XUnit> xs;
    build_xunits<AB>(x, cols, xs);
    int nu = cols / UNIT_COLS;
    size_t rb = row_bytes(type, cols), rs = align16(rb);
    std::vector<uint8_t> row(rs + 16);
    uint8_t* r = row.data() + ((16 - ((uintptr_t)row.data() & 15)) & 15);
    for (int i = 0; i < rows; ++i) {
        const uint8_t* src = w + (size_t)i * rb;
        float acc = 0.f;
        if (type == T_Q4_K) {
            memcpy(r, src, rb);
            for (int u = 0; u < nu; ++u) acc += unit_dot_q4k<AB>(r + (size_t)(u >> 1) * 144, u & 1, xs[u]);
        } else if (type == T_Q6_K) {
            repack_row_q6k(src, r, cols / 256);
            for (int u = 0; u < nu; ++u) acc += unit_dot_q6k<AB>(r, cols / 256, u, xs[u]);
        } else if (type == T_Q8_0) {
            repack_row_q80(src, r, cols);
            for (int u = 0; u < nu; ++u) acc += unit_dot_q80<AB>(r, cols, u, xs[u]);
        } else return -1;
        y[i] = acc;
    }
    return 0;
}

// rows taken four at a time through the quad functions, x read from swizzled planes exactly as the GPU prologue lays them out
template <int AB>
static int run_quads(int type, const uint8_t* w, int rows, int cols, const float* x, float* y) {
    std::vector<XUnit> xs;
    build_xunits<AB>(x, cols, xs);
    const int nu = cols / UNIT_COLS;
    // planes: hi [cols], lo [cols] (16-B chunk j of unit u at physical chunk j ^ (u & 7)), sx / sm [cols/32], s16 [cols/16]
    std::vector<uint8_t> raw(2 * (size_t)cols + 64);
    uint8_t* hi = raw.data() + ((16 - ((uintptr_t)raw.data() & 15)) & 15);
    uint8_t* lo = hi + cols;
    std::vector<float> sx(cols / 32 + 4), sm(cols / 32 + 4);
    std::vector<int> s16(cols / 16 + 4);
    float* sxp = sx.data() + ((16 - ((uintptr_t)sx.data() & 15)) & 15) / 4;
    float* smp = sm.data() + ((16 - ((uintptr_t)sm.data() & 15)) & 15) / 4;
    int* s16p = s16.data() + ((16 - ((uintptr_t)s16.data() & 15)) & 15) / 4;
    for (int u = 0; u < nu; ++u) {
        for (int j = 0; j < 8; ++j) {
            memcpy(hi + (size_t)u * 128 + ((j ^ (u & 7)) << 4), &xs[u].hi[4 * j], 16);
            memcpy(lo + (size_t)u * 128 + ((j ^ (u & 7)) << 4), &xs[u].lo[4 * j], 16);
        }
        for (int b = 0; b < 4; ++b) { sxp[4 * u + b] = xs[u].sx[b]; smp[4 * u + b] = xs[u].sm[b]; }
        for (int g = 0; g < 8; ++g) s16p[8 * u + g] = xs[u].s16[g];
    }
    size_t rb = row_bytes(type, cols), rs = align16(rb);
    std::vector<uint8_t> buf(4 * rs + 16);
    uint8_t* base = buf.data() + ((16 - ((uintptr_t)buf.data() & 15)) & 15);
    for (int i = 0; i < rows; i += 4) {
        const uint8_t* rp[4];
        for (int r = 0; r < 4; ++r) {
            const int ri = i + r < rows ? i + r : rows - 1;       // clamp like the GPU consumer does for ragged quads
            uint8_t* dst = base + (size_t)r * rs;
            if (type == T_Q4_K) memcpy(dst, w + (size_t)ri * rb, rb);
            else if (type == T_Q6_K) repack_row_q6k(w + (size_t)ri * rb, dst, cols / 256);
            else repack_row_q80(w + (size_t)ri * rb, dst, cols);
            rp[r] = dst;
        }
        float acc[4] = {0, 0, 0, 0};
        for (int u = 0; u < nu; ++u) {
            XPlanes xp{hi + (size_t)u * 128, lo + (size_t)u * 128, sxp + 4 * u, smp + 4 * u, s16p + 8 * u, u & 7};
            float o[4];
            if (type == T_Q4_K) {
                const uint8_t* bp[4];
                for (int r = 0; r < 4; ++r) bp[r] = rp[r] + (size_t)(u >> 1) * 144;
                quad_dot_q4k<AB>(bp, u & 1, xp, o);
            } else if (type == T_Q6_K) quad_dot_q6k<AB>(rp, cols / 256, u, xp, o);
            else quad_dot_q80<AB>(rp, cols, u, xp, o);
            for (int r = 0; r < 4; ++r) acc[r] += o[r];
        }
        for (int r = 0; r < 4 && i + r < rows; ++r) y[i + r] = acc[r];
    }
    return 0;
}

extern "C" int hc_gemv_quads(int type, const uint8_t* w, int rows, int cols, const float* x, float* y, int abits) {
    if (cols % 128) return -2;
    if ((type == T_Q4_K || type == T_Q6_K) && cols % 256) return -2;
    return abits == 16 ? run_quads<16>(type, w, rows, cols, x, y) : run_quads<8>(type, w, rows, cols, x, y);
}

extern "C" int hc_gemv_pairs(int type, const uint8_t* w, int rows, int cols, const float* x, float* y, int abits) {
    if (cols % 256 || (rows & 1)) return -2;
    return abits == 16 ? run_pairs<16>(type, w, rows, cols, x, y) : run_pairs<8>(type, w, rows, cols, x, y);
}

extern "C" int hc_gemv(int type, const uint8_t* w, int rows, int cols, const float* x, float* y, int abits) {
    if (cols % 128) return -2;
    if (type == T_Q4_K || type == T_Q6_K) { if (cols % 256) return -2; }
    return abits == 16 ? run<16>(type, w, rows, cols, x, y) : run<8>(type, w, rows, cols, x, y);
}

// GGUF reader check: returns number of tensors or -1; fills a few fields
extern "C" int hc_gguf_probe(const char* path, char* arch, int cap, uint64_t* n_kv, uint64_t* total_bytes) {
    GGUFFile f;
    std::string err = f.open(path);
    if (!err.empty()) { snprintf(arch, cap, "%s", err.c_str()); return -1; }
    snprintf(arch, cap, "%s", f.get_s("general.architecture", "").c_str());
    *n_kv = f.kv.size();
    uint64_t tot = 0;
    for (auto& t : f.tensors) tot += t.nbytes;
    *total_bytes = tot;
    return (int)f.tensors.size();
}

// tokenizer (product code, gridllm_b200/csrc/tokenizer.cpp) driven from the CPU tests
#include "../../gridllm_b200/csrc/tokenizer.h"
extern "C" int hc_tokenize(const char* gguf_path, const char* text, int add_bos, int parse_special, int32_t* ids, int cap) {
    GGUFFile f;
    if (!f.open(gguf_path).empty()) return -1;
    Tokenizer t;
    if (!t.load(f)) return -2;
    std::vector<int32_t> v = t.encode(text, add_bos != 0, parse_special != 0);
    if ((int)v.size() > cap) return -3;
    for (size_t i = 0; i < v.size(); ++i) ids[i] = v[i];
    return (int)v.size();
}
extern "C" int hc_detokenize(const char* gguf_path, const int32_t* ids, int n, char* buf, int cap) {
    GGUFFile f;
    if (!f.open(gguf_path).empty()) return -1;
    Tokenizer t;
    if (!t.load(f)) return -2;
    std::string s = t.decode(ids, n);
    if ((int)s.size() >= cap) return -3;
    memcpy(buf, s.data(), s.size());
    buf[s.size()] = 0;
    return (int)s.size();
}
