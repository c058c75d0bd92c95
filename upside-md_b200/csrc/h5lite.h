// Minimal self-contained HDF5 reader/writer for the engine (no libhdf5 in this image).
//
// Reads what libhdf5 1.8 / PyTables emit for Upside .up configuration files and parameter libraries:
// superblock v0/v1, v1 object headers (+continuations), symbol-table groups (B-tree v1 + local heap),
// contiguous / compact / chunked (B-tree v1) datasets with deflate, shuffle and fletcher32 filters,
// attribute messages v1-v3, fixed-point / IEEE float / fixed-length string datatypes.
// Writes the same subset (contiguous datasets only).  The whole file is held as an in-memory tree.
//
// Replaces, for this project, the libhdf5 calls wrapped by the reference's src/h5_support.{h,cpp}.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace h5l {

enum class Kind : uint8_t { Int, UInt, Float, String };

struct DType {
    Kind kind;
    uint32_t size;   // bytes per element (string: fixed width)
    DType(): kind(Kind::Float), size(4) {}
    DType(Kind k, uint32_t s): kind(k), size(s) {}
};

struct Array {
    DType dt;
    bool scalar = false;            // scalar dataspace (rank 0)
    std::vector<uint64_t> dims;     // empty for scalar
    std::vector<uint64_t> maxdims;  // writer/reader hint, may be empty
    std::vector<uint8_t> raw;       // little-endian, C order
    uint64_t count() const { uint64_t n = 1; for (auto d : dims) n *= d; return n; }
};

struct Node {
    bool is_group = true;
    std::map<std::string, std::unique_ptr<Node>> children;   // std::map => name-sorted, as H5_INDEX_NAME
    std::map<std::string, Array> attrs;
    Array data;                                              // datasets only
};

// throws std::string on malformed / unsupported input
std::unique_ptr<Node> load(const std::string& path);
void save(const Node& root, const std::string& path);

// path resolution relative to base: handles ".", leading "/" (relative to root passed in), nested "a/b/c"
Node* find(Node* base, const std::string& path);
inline const Node* find(const Node* base, const std::string& path) { return find(const_cast<Node*>(base), path); }
Node* ensure_group(Node* base, const std::string& path);

// numeric conversion of any Int/UInt/Float array to T (as H5Dread would do with a native memory type)
template <typename T> std::vector<T> as(const Array& a);
extern template std::vector<float> as<float>(const Array&);
extern template std::vector<double> as<double>(const Array&);
extern template std::vector<int> as<int>(const Array&);
extern template std::vector<long> as<long>(const Array&);
extern template std::vector<unsigned> as<unsigned>(const Array&);
std::vector<std::string> as_strings(const Array& a);

template <typename T> Array make_array(const std::vector<T>& v, const std::vector<uint64_t>& dims);
extern template Array make_array<float>(const std::vector<float>&, const std::vector<uint64_t>&);
extern template Array make_array<double>(const std::vector<double>&, const std::vector<uint64_t>&);
extern template Array make_array<int>(const std::vector<int>&, const std::vector<uint64_t>&);
extern template Array make_array<long>(const std::vector<long>&, const std::vector<uint64_t>&);
Array make_string_array(const std::vector<std::string>& v);
Array make_string_scalar(const std::string& s);

}  // namespace h5l
