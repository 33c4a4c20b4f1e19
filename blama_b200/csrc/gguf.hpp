// gguf.hpp -- GGUF v2/v3 container reader (host side of blk_model_load).
//
// Stands in for llama.cpp's llama-model-loader / gguf.cpp as reached from reference
// inference/code/llama/Model.cpp:50-53.  Memory-maps the file, exposes metadata and per-tensor
// (type, shape, byte range); tensor bytes are handed to the uploader untouched.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace blk {

enum GgmlType : int { GT_F32 = 0, GT_F16 = 1, GT_Q8_0 = 8, GT_Q4_K = 12, GT_Q5_K = 13, GT_Q6_K = 14 };

inline bool ggml_type_geometry(int t, int& blck, int& bytes) {
    switch (t) {
        case GT_F32: blck = 1; bytes = 4; return true;
        case GT_F16: blck = 1; bytes = 2; return true;
        case GT_Q8_0: blck = 32; bytes = 34; return true;
        case GT_Q4_K: blck = 256; bytes = 144; return true;
        case GT_Q5_K: blck = 256; bytes = 176; return true;
        case GT_Q6_K: blck = 256; bytes = 210; return true;
        default: return false;
    }
}

struct GgufTensor {
    std::string name;
    int type = -1;
    int n_dims = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    const uint8_t* data = nullptr;
    size_t nbytes = 0;
    int64_t n_rows() const { return ne[1] * ne[2] * ne[3]; }
};

class GgufFile {
public:
    explicit GgufFile(const std::string& path) {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) throw std::runtime_error("cannot open " + path);
        struct stat st;
        if (fstat(fd_, &st) != 0) { ::close(fd_); throw std::runtime_error("cannot stat " + path); }
        size_ = (size_t)st.st_size;
        void* p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (p == MAP_FAILED) { ::close(fd_); throw std::runtime_error("mmap failed for " + path); }
        base_ = (const uint8_t*)p;
        try { parse(); } catch (...) { munmap((void*)base_, size_); ::close(fd_); throw; }
    }
    ~GgufFile() { if (base_) munmap((void*)base_, size_); if (fd_ >= 0) ::close(fd_); }
    GgufFile(const GgufFile&) = delete;
    GgufFile& operator=(const GgufFile&) = delete;

    const GgufTensor* find(const std::string& n) const { auto it = index_.find(n); return it == index_.end() ? nullptr : &tensors_[it->second]; }
    const std::vector<GgufTensor>& tensors() const { return tensors_; }
    bool has_num(const std::string& k) const { return nums_.count(k) != 0; }
    double num(const std::string& k, double def) const { auto it = nums_.find(k); return it == nums_.end() ? def : it->second; }
    double num_required(const std::string& k) const { auto it = nums_.find(k); if (it == nums_.end()) throw std::runtime_error("gguf: missing key " + k); return it->second; }
    const std::string* str(const std::string& k) const { auto it = strs_.find(k); return it == strs_.end() ? nullptr : &it->second; }
    const std::vector<std::string>* str_array(const std::string& k) const { auto it = str_arrays_.find(k); return it == str_arrays_.end() ? nullptr : &it->second; }
    // integer arrays (u8 .. i64 element types), e.g. tokenizer.ggml.token_type
    const std::vector<int64_t>* int_array(const std::string& k) const { auto it = int_arrays_.find(k); return it == int_arrays_.end() ? nullptr : &it->second; }
    size_t file_size() const { return size_; }

private:
    struct Cur {
        const uint8_t* p; const uint8_t* end;
        size_t left() const { return (size_t)(end - p); }
        template <class T> T get() { if (sizeof(T) > left()) throw std::runtime_error("gguf: truncated file"); T v; memcpy(&v, p, sizeof(T)); p += sizeof(T); return v; }
        std::string str() { uint64_t n = get<uint64_t>(); if (n > (uint64_t)(end - p)) throw std::runtime_error("gguf: truncated string"); std::string s((const char*)p, (size_t)n); p += n; return s; }
    };
    void read_value(Cur& c, uint32_t t, const std::string& key) {
        static const int width[13] = {1, 1, 2, 2, 4, 4, 4, 1, 0, 0, 8, 8, 8};
        switch (t) {
            case 0: nums_[key] = c.get<uint8_t>(); break;
            case 1: nums_[key] = c.get<int8_t>(); break;
            case 2: nums_[key] = c.get<uint16_t>(); break;
            case 3: nums_[key] = c.get<int16_t>(); break;
            case 4: nums_[key] = c.get<uint32_t>(); break;
            case 5: nums_[key] = c.get<int32_t>(); break;
            case 6: nums_[key] = c.get<float>(); break;
            case 7: nums_[key] = c.get<uint8_t>() ? 1 : 0; break;
            case 8: strs_[key] = c.str(); break;
            case 10: nums_[key] = (double)c.get<uint64_t>(); break;
            case 11: nums_[key] = (double)c.get<int64_t>(); break;
            case 12: nums_[key] = c.get<double>(); break;
            case 9: {
                uint32_t et = c.get<uint32_t>();
                uint64_t n = c.get<uint64_t>();
                nums_[key + ".count"] = (double)n;
                if (et == 8) {
                    // every string costs at least its 8-byte length: a count the remaining bytes cannot hold is a corrupt header
                    if (n > c.left() / 8) throw std::runtime_error("gguf: truncated array");
                    auto& v = str_arrays_[key];
                    v.reserve((size_t)n);
                    for (uint64_t i = 0; i < n; i++) v.push_back(c.str());
                } else if (et < 13 && width[et]) {
                    if (n > c.left() / (uint64_t)width[et]) throw std::runtime_error("gguf: truncated array");
                    if (et != 6 && et != 12 && n <= (1u << 24)) {      // integer / bool arrays are kept (token types); float arrays are skipped
                        auto& v = int_arrays_[key];
                        v.resize((size_t)n);
                        for (uint64_t i = 0; i < n; i++) {
                            const uint8_t* q = c.p + i * width[et];
                            int64_t x = 0;
                            switch (et) {
                                case 0: case 7: x = *q; break;
                                case 1: x = (int8_t)*q; break;
                                case 2: { uint16_t t2; memcpy(&t2, q, 2); x = t2; break; }
                                case 3: { int16_t t2; memcpy(&t2, q, 2); x = t2; break; }
                                case 4: { uint32_t t4; memcpy(&t4, q, 4); x = t4; break; }
                                case 5: { int32_t t4; memcpy(&t4, q, 4); x = t4; break; }
                                case 10: { uint64_t t8; memcpy(&t8, q, 8); x = (int64_t)t8; break; }
                                default: { int64_t t8; memcpy(&t8, q, 8); x = t8; break; }
                            }
                            v[(size_t)i] = x;
                        }
                    }
                    c.p += n * width[et];
                } else {
                    throw std::runtime_error("gguf: unsupported array element type");
                }
                break;
            }
            default: throw std::runtime_error("gguf: unknown metadata value type");
        }
    }
    void parse() {
        Cur c{base_, base_ + size_};
        if (c.get<uint32_t>() != 0x46554747u) throw std::runtime_error("gguf: bad magic");
        const uint32_t version = c.get<uint32_t>();
        if (version != 2 && version != 3) throw std::runtime_error("gguf: unsupported version " + std::to_string(version));
        const uint64_t n_tensors = c.get<uint64_t>();
        const uint64_t n_kv = c.get<uint64_t>();
        // header counts are bounded by what the file can hold: a key is >= 8 + 4 + 1 bytes, a tensor info >= 8 + 4 + 8 + 4 + 8
        if (n_kv > c.left() / 13 || n_tensors > c.left() / 32) throw std::runtime_error("gguf: header counts exceed the file size");
        for (uint64_t i = 0; i < n_kv; i++) {
            std::string key = c.str();
            uint32_t t = c.get<uint32_t>();
            read_value(c, t, key);
        }
        std::vector<uint64_t> offs(n_tensors);
        tensors_.resize(n_tensors);
        for (uint64_t i = 0; i < n_tensors; i++) {
            GgufTensor& t = tensors_[i];
            t.name = c.str();
            t.n_dims = (int)c.get<uint32_t>();
            if (t.n_dims < 1 || t.n_dims > 4) throw std::runtime_error("gguf: bad n_dims for " + t.name);
            for (int d = 0; d < t.n_dims; d++) {
                const uint64_t e = c.get<uint64_t>();
                if (e == 0 || e > (uint64_t)1 << 40) throw std::runtime_error("gguf: bad dimension for " + t.name);
                t.ne[d] = (int64_t)e;
            }
            t.type = (int)c.get<uint32_t>();
            offs[i] = c.get<uint64_t>();
        }
        const double align_d = num("general.alignment", 32);
        if (!(align_d >= 1 && align_d <= 65536)) throw std::runtime_error("gguf: bad general.alignment");
        const size_t align = (size_t)align_d;
        if (align & (align - 1)) throw std::runtime_error("gguf: general.alignment must be a power of two");
        const size_t data0 = ((size_t)(c.p - base_) + align - 1) / align * align;
        if (data0 > size_) throw std::runtime_error("gguf: truncated file");
        for (uint64_t i = 0; i < n_tensors; i++) {
            GgufTensor& t = tensors_[i];
            int blck = 0, bytes = 0;
            if (!ggml_type_geometry(t.type, blck, bytes)) throw std::runtime_error("gguf: unsupported tensor type " + std::to_string(t.type) + " for " + t.name);
            if (t.ne[0] % blck) throw std::runtime_error("gguf: row length not a multiple of the block size for " + t.name);
            // overflow-checked element count (every factor <= 2^40 was checked above)
            unsigned __int128 elems = 1;
            for (int d = 0; d < t.n_dims; d++) { elems *= (unsigned __int128)t.ne[d]; if (elems > ((unsigned __int128)1 << 62)) throw std::runtime_error("gguf: tensor too large: " + t.name); }
            const unsigned __int128 nb = elems / (unsigned)blck * (unsigned)bytes;
            const size_t room = size_ - data0;
            if (offs[i] > room || nb > (unsigned __int128)(room - offs[i])) throw std::runtime_error("gguf: tensor data out of range for " + t.name);
            t.nbytes = (size_t)nb;
            t.data = base_ + data0 + offs[i];
            index_[t.name] = (size_t)i;
        }
    }

    int fd_ = -1;
    const uint8_t* base_ = nullptr;
    size_t size_ = 0;
    std::vector<GgufTensor> tensors_;
    std::map<std::string, size_t> index_;
    std::map<std::string, double> nums_;
    std::map<std::string, std::string> strs_;
    std::map<std::string, std::vector<std::string>> str_arrays_;
    std::map<std::string, std::vector<int64_t>> int_arrays_;
};

} // namespace blk
