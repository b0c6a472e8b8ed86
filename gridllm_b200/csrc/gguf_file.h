// GGUF v2/v3 container reader (host side of the engine's weight loader).
// From-scratch restatement of the public GGUF layout; replaces the model load that, in the
// reference, happens inside the Ollama daemon behind OllamaService (client/src/services/
// OllamaService.ts:17-25 constructs only an HTTP client; the GGUF never enters the tree).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace gl {

enum GGMLType : uint32_t { T_F32 = 0, T_F16 = 1, T_Q8_0 = 8, T_Q4_K = 12, T_Q6_K = 14, T_BF16 = 30 };

struct BlockGeom { int weights; int bytes; };
// returns {0,0} for types outside the hot path
BlockGeom block_geom(uint32_t type);
const char* type_name(uint32_t type);
inline size_t row_bytes(uint32_t type, int64_t cols) {
    BlockGeom g = block_geom(type);
    return g.weights ? (size_t)(cols / g.weights) * g.bytes : 0;
}

struct GGUFValue {
    uint32_t type = 0;                // GGUF metadata value type
    uint64_t u = 0;                   // any integer / bool
    double f = 0;                     // any float
    std::string s;                    // string
    uint32_t arr_type = 0;            // for arrays
    std::vector<std::string> arr_s;   // array of strings
    std::vector<int64_t> arr_i;       // array of ints
    std::vector<double> arr_f;        // array of floats
};

struct GGUFTensor {
    std::string name;
    uint32_t type = 0;
    std::vector<int64_t> ne;          // ne[0] = cols (contiguous), ne[1] = rows ...
    uint64_t offset = 0;              // from data section start
    const uint8_t* data = nullptr;    // into the mapping
    size_t nbytes = 0;
    int64_t cols() const { return ne.empty() ? 0 : ne[0]; }
    int64_t rows() const { int64_t r = 1; for (size_t i = 1; i < ne.size(); ++i) r *= ne[i]; return r; }
};

class GGUFFile {
public:
    GGUFFile() = default;
    ~GGUFFile();
    GGUFFile(const GGUFFile&) = delete;
    GGUFFile& operator=(const GGUFFile&) = delete;

    // returns empty string on success, else an error message
    std::string open(const std::string& path);

    const GGUFValue* find(const std::string& key) const;
    uint64_t get_u(const std::string& key, uint64_t dflt) const;
    double get_f(const std::string& key, double dflt) const;
    std::string get_s(const std::string& key, const std::string& dflt) const;
    const GGUFTensor* tensor(const std::string& name) const;

    uint32_t version = 0;
    uint64_t file_bytes = 0;
    std::map<std::string, GGUFValue> kv;
    std::vector<GGUFTensor> tensors;

private:
    void* map_ = nullptr;
    size_t map_len_ = 0;
    std::map<std::string, size_t> index_;
};

}  // namespace gl
