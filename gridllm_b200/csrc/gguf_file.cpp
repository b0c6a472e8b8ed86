// GGUF v2/v3 reader: mmap + bounds-checked cursor.  See gguf_file.h.
#include "gguf_file.h"

#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace gl {

BlockGeom block_geom(uint32_t type) {
    switch (type) {
        case T_F32: return {1, 4};
        case T_F16: return {1, 2};
        case T_BF16: return {1, 2};
        case T_Q8_0: return {32, 34};
        case T_Q4_K: return {256, 144};
        case T_Q6_K: return {256, 210};
        default: return {0, 0};
    }
}

const char* type_name(uint32_t type) {
    switch (type) {
        case T_F32: return "F32";
        case T_F16: return "F16";
        case T_BF16: return "BF16";
        case T_Q8_0: return "Q8_0";
        case T_Q4_K: return "Q4_K";
        case T_Q6_K: return "Q6_K";
        default: return "?";
    }
}

namespace {

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    template <typename T> T rd() {
        T v{};
        if ((size_t)(end - p) < sizeof(T)) { ok = false; return v; }
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        uint64_t n = rd<uint64_t>();
        if (!ok || (uint64_t)(end - p) < n) { ok = false; return {}; }
        std::string s((const char*)p, (size_t)n);
        p += n;
        return s;
    }
};

enum : uint32_t { V_U8, V_I8, V_U16, V_I16, V_U32, V_I32, V_F32, V_BOOL, V_STR, V_ARR, V_U64, V_I64, V_F64 };

bool read_scalar(Cursor& c, uint32_t t, GGUFValue& v) {
    switch (t) {
        case V_U8: v.u = c.rd<uint8_t>(); v.f = (double)v.u; break;
        case V_I8: { int8_t x = c.rd<int8_t>(); v.u = (uint64_t)(int64_t)x; v.f = x; } break;
        case V_U16: v.u = c.rd<uint16_t>(); v.f = (double)v.u; break;
        case V_I16: { int16_t x = c.rd<int16_t>(); v.u = (uint64_t)(int64_t)x; v.f = x; } break;
        case V_U32: v.u = c.rd<uint32_t>(); v.f = (double)v.u; break;
        case V_I32: { int32_t x = c.rd<int32_t>(); v.u = (uint64_t)(int64_t)x; v.f = x; } break;
        case V_F32: { float x = c.rd<float>(); v.f = x; v.u = (uint64_t)x; } break;
        case V_BOOL: v.u = c.rd<uint8_t>() ? 1 : 0; v.f = (double)v.u; break;
        case V_U64: v.u = c.rd<uint64_t>(); v.f = (double)v.u; break;
        case V_I64: { int64_t x = c.rd<int64_t>(); v.u = (uint64_t)x; v.f = (double)x; } break;
        case V_F64: v.f = c.rd<double>(); v.u = (uint64_t)v.f; break;
        default: return false;
    }
    return c.ok;
}

}  // namespace

GGUFFile::~GGUFFile() {
    if (map_) munmap(map_, map_len_);
}

std::string GGUFFile::open(const std::string& path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return "cannot open '" + path + "': " + std::strerror(errno);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 24) { ::close(fd); return "cannot stat / too small: " + path; }
    map_len_ = (size_t)st.st_size;
    map_ = mmap(nullptr, map_len_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map_ == MAP_FAILED) { map_ = nullptr; return "mmap failed: " + path; }
    file_bytes = map_len_;
    madvise(map_, map_len_, MADV_SEQUENTIAL);

    Cursor c{(const uint8_t*)map_, (const uint8_t*)map_ + map_len_};
    uint32_t magic = c.rd<uint32_t>();
    if (magic != 0x46554747u) return "not a GGUF file (bad magic)";
    version = c.rd<uint32_t>();
    if (version != 2 && version != 3) return "unsupported GGUF version " + std::to_string(version);
    uint64_t n_tensors = c.rd<uint64_t>();
    uint64_t n_kv = c.rd<uint64_t>();
    if (!c.ok || n_tensors > (1u << 20) || n_kv > (1u << 20)) return "corrupt GGUF header";

    for (uint64_t i = 0; i < n_kv; ++i) {
        std::string key = c.str();
        GGUFValue v;
        v.type = c.rd<uint32_t>();
        if (!c.ok) return "truncated GGUF metadata";
        if (v.type == V_STR) {
            v.s = c.str();
        } else if (v.type == V_ARR) {
            v.arr_type = c.rd<uint32_t>();
            uint64_t n = c.rd<uint64_t>();
            // every element takes at least one byte of the file: an element count beyond what is left is damage, and is
            // refused BEFORE any reserve() (2^28 strings would ask for 8 GB ahead of the first validated element)
            const uint64_t left = (uint64_t)(c.end - c.p);
            if (!c.ok || n > (1ull << 28) || n > left) return "corrupt GGUF array '" + key + "'";
            if (v.arr_type == V_STR) {
                v.arr_s.reserve((size_t)n);
                for (uint64_t j = 0; j < n && c.ok; ++j) v.arr_s.push_back(c.str());
            } else if (v.arr_type == V_F32 || v.arr_type == V_F64) {
                v.arr_f.reserve((size_t)n);
                for (uint64_t j = 0; j < n && c.ok; ++j) { GGUFValue e; if (!read_scalar(c, v.arr_type, e)) break; v.arr_f.push_back(e.f); }
            } else {
                v.arr_i.reserve((size_t)n);
                for (uint64_t j = 0; j < n && c.ok; ++j) { GGUFValue e; if (!read_scalar(c, v.arr_type, e)) { c.ok = false; break; } v.arr_i.push_back((int64_t)e.u); }
            }
        } else if (!read_scalar(c, v.type, v)) {
            return "bad GGUF value type for key '" + key + "'";
        }
        if (!c.ok) return "truncated GGUF metadata at key '" + key + "'";
        kv.emplace(std::move(key), std::move(v));
    }

    tensors.resize((size_t)n_tensors);
    for (auto& t : tensors) {
        t.name = c.str();
        uint32_t nd = c.rd<uint32_t>();
        if (!c.ok || nd > 4) return "corrupt GGUF tensor info";
        t.ne.resize(nd);
        for (uint32_t d = 0; d < nd; ++d) {
            t.ne[d] = (int64_t)c.rd<uint64_t>();
            if (t.ne[d] < 0 || t.ne[d] > (int64_t)1 << 40) return "corrupt GGUF tensor info (dimension out of range)";
        }
        {   // element count must not overflow (4 dims of up to 2^40 each could)
            unsigned __int128 prod = 1;
            for (uint32_t d = 0; d < nd; ++d) prod *= (unsigned __int128)(uint64_t)t.ne[d];
            if (prod > ((unsigned __int128)1 << 48)) return "corrupt GGUF tensor info (too many elements)";
        }
        t.type = c.rd<uint32_t>();
        t.offset = c.rd<uint64_t>();
        if (!c.ok) return "truncated GGUF tensor info";
    }
    uint64_t align = get_u("general.alignment", 32);
    if (align == 0 || (align & (align - 1))) return "bad general.alignment";
    size_t data_off = (size_t)(c.p - (const uint8_t*)map_);
    data_off = (data_off + align - 1) / align * align;
    if (data_off > map_len_) return "truncated GGUF (no tensor data section)";
    const size_t data_len = map_len_ - data_off;
    for (size_t i = 0; i < tensors.size(); ++i) {
        auto& t = tensors[i];
        BlockGeom g = block_geom(t.type);
        if (g.weights) {
            if (t.cols() % g.weights) return "tensor '" + t.name + "': cols not a multiple of the block size";
            t.nbytes = row_bytes(t.type, t.cols()) * (size_t)t.rows();
            // overflow-free: offset and size are compared against what is left, never added to anything first
            if (t.offset > data_len || t.nbytes > data_len - t.offset) return "tensor '" + t.name + "' runs past end of file";
        } else if (t.offset > data_len) {
            return "tensor '" + t.name + "' starts past end of file";
        }
        t.data = (const uint8_t*)map_ + data_off + t.offset;
        index_[t.name] = i;
    }
    return {};
}

const GGUFValue* GGUFFile::find(const std::string& key) const {
    auto it = kv.find(key);
    return it == kv.end() ? nullptr : &it->second;
}
uint64_t GGUFFile::get_u(const std::string& key, uint64_t dflt) const {
    auto* v = find(key);
    return v ? v->u : dflt;
}
double GGUFFile::get_f(const std::string& key, double dflt) const {
    auto* v = find(key);
    return v ? v->f : dflt;
}
std::string GGUFFile::get_s(const std::string& key, const std::string& dflt) const {
    auto* v = find(key);
    return (v && v->type == V_STR) ? v->s : dflt;
}
const GGUFTensor* GGUFFile::tensor(const std::string& name) const {
    auto it = index_.find(name);
    return it == index_.end() ? nullptr : &tensors[it->second];
}

}  // namespace gl
