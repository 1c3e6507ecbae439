// Minimal .npz reader / writer for model checkpoints.
//
// Wire format of the reference's checkpoints (src/graph/expression_graph.h:442-502 through
// 3rd_party/cnpy): a ZIP archive of STORED (uncompressed) members "<name>.npy", each a NumPy v1.0
// array file - magic "\x93NUMPY", version 1.0, little-endian uint16 header length, a Python-dict
// header "{'descr': '<f4', 'fortran_order': False, 'shape': (rows, cols), }" padded with spaces to
// a multiple of 16 bytes and terminated by '\n', then the raw little-endian data.  The model
// description travels in the same archive as the char array "special:model.yml"
// (src/common/config.cpp:54-67).  Written here from the format definitions (PKZIP APPNOTE local file
// header / central directory / end-of-central-directory records, NumPy NEP 1); files written by
// numpy.savez (stored, no ZIP64) and by the reference load, files written here load in numpy and in
// the reference.  Compressed members (numpy.savez_compressed) are rejected with a clear error.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common/definitions.h"

namespace marian {
namespace npz {

struct Array {
  std::vector<int> shape;
  char kind{'f'};       // 'f' float32, 'i' int8 (char arrays: special:model.yml)
  std::vector<char> bytes;

  size_t elements() const {
    size_t n = 1;
    for(int d : shape)
      n *= (size_t)d;
    return n;
  }
  const float* floats() const { return (const float*)bytes.data(); }
  std::string text() const { return std::string(bytes.data(), strnlen(bytes.data(), bytes.size())); }
};

inline uint32_t crc32(const char* data, size_t n, uint32_t crc = 0) {
  static uint32_t table[256];
  static bool init = false;
  if(!init) {
    for(uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for(int k = 0; k < 8; ++k)
        c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  crc = ~crc;
  for(size_t i = 0; i < n; ++i)
    crc = table[(crc ^ (uint8_t)data[i]) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

namespace detail {
template <class T>
inline void put(std::vector<char>& out, T v) {
  for(size_t i = 0; i < sizeof(T); ++i)
    out.push_back((char)((v >> (8 * i)) & 0xFF));
}
template <class T>
inline T get(const char* p) {
  T v = 0;
  for(size_t i = 0; i < sizeof(T); ++i)
    v |= (T)(uint8_t)p[i] << (8 * i);
  return v;
}

inline std::vector<char> npyHeader(const Array& a) {
  std::string dict = "{'descr': '";
  dict += a.kind == 'f' ? "<f4" : "|i1";
  dict += "', 'fortran_order': False, 'shape': (";
  for(size_t i = 0; i < a.shape.size(); ++i) {
    if(i)
      dict += ", ";
    dict += std::to_string(a.shape[i]);
  }
  if(a.shape.size() == 1)
    dict += ",";
  dict += "), }";
  size_t pad = 16 - (10 + dict.size()) % 16;  // preamble is 10 bytes; the dict ends with '\n'
  dict.append(pad, ' ');
  dict.back() = '\n';
  std::vector<char> h;
  h.push_back((char)0x93);
  for(char c : std::string("NUMPY"))
    h.push_back(c);
  h.push_back(1);
  h.push_back(0);
  put<uint16_t>(h, (uint16_t)dict.size());
  h.insert(h.end(), dict.begin(), dict.end());
  return h;
}

inline Array parseNpy(const char* p, size_t n, const std::string& member) {
  ABORT_IF(n < 10 || std::memcmp(p, "\x93NUMPY", 6) != 0, "npz: member is not a .npy array:", member);
  size_t hlen, off;
  if(p[6] == 1) {
    hlen = get<uint16_t>(p + 8);
    off = 10;
  } else {
    hlen = get<uint32_t>(p + 8);
    off = 12;
  }
  ABORT_IF(off + hlen > n, "npz: truncated .npy header in", member);
  std::string dict(p + off, hlen);
  Array a;
  auto d = dict.find("'descr':");
  ABORT_IF(d == std::string::npos, "npz: no descr in", member);
  auto q1 = dict.find('\'', d + 8), q2 = dict.find('\'', q1 + 1);
  std::string descr = dict.substr(q1 + 1, q2 - q1 - 1);
  size_t word;
  if(descr == "<f4" || descr == "=f4") {
    a.kind = 'f';
    word = 4;
  } else if(descr == "|i1" || descr == "<i1" || descr == "|S1" || descr == "|u1" || descr == "<u1") {
    a.kind = 'i';
    word = 1;
  } else {
    ABORT("npz: unsupported dtype", descr, "in", member, "(float32 parameters and char arrays only)");
  }
  ABORT_IF(dict.find("'fortran_order': False") == std::string::npos, "npz: fortran-ordered array in", member);
  auto s1 = dict.find('(', dict.find("'shape':")), s2 = dict.find(')', s1);
  std::string dims = dict.substr(s1 + 1, s2 - s1 - 1);
  size_t pos = 0;
  while(pos < dims.size()) {
    while(pos < dims.size() && (dims[pos] < '0' || dims[pos] > '9'))
      ++pos;
    if(pos >= dims.size())
      break;
    size_t e = pos;
    while(e < dims.size() && dims[e] >= '0' && dims[e] <= '9')
      ++e;
    a.shape.push_back(std::stoi(dims.substr(pos, e - pos)));
    pos = e;
  }
  size_t bytes = a.elements() * word;
  ABORT_IF(off + hlen + bytes > n, "npz: truncated data in", member);
  a.bytes.assign(p + off + hlen, p + off + hlen + bytes);
  return a;
}
}  // namespace detail

// Writes all arrays in the given order (the reference appends member by member in map order).
inline void save(const std::string& path, const std::vector<std::pair<std::string, Array>>& arrays) {
  FILE* fp = std::fopen(path.c_str(), "wb");
  ABORT_IF(!fp, "npz: cannot open for writing:", path);
  std::vector<char> central;
  uint32_t offset = 0;
  uint16_t count = 0;
  for(auto& kv : arrays) {
    std::string fname = kv.first + ".npy";
    std::vector<char> header = detail::npyHeader(kv.second);
    uint32_t size = (uint32_t)(header.size() + kv.second.bytes.size());
    ABORT_IF(header.size() + kv.second.bytes.size() >= 0xFFFFFFFFull, "npz: member too large for a non-ZIP64 archive:", kv.first);
    uint32_t crc = crc32(header.data(), header.size());
    crc = crc32(kv.second.bytes.data(), kv.second.bytes.size(), crc);

    std::vector<char> local;
    detail::put<uint32_t>(local, 0x04034b50u);
    detail::put<uint16_t>(local, 20);  // version needed
    detail::put<uint16_t>(local, 0);   // flags
    detail::put<uint16_t>(local, 0);   // method: stored
    detail::put<uint16_t>(local, 0);   // time
    detail::put<uint16_t>(local, 0x21);  // date (1980-01-01)
    detail::put<uint32_t>(local, crc);
    detail::put<uint32_t>(local, size);
    detail::put<uint32_t>(local, size);
    detail::put<uint16_t>(local, (uint16_t)fname.size());
    detail::put<uint16_t>(local, 0);
    local.insert(local.end(), fname.begin(), fname.end());

    detail::put<uint32_t>(central, 0x02014b50u);
    detail::put<uint16_t>(central, 20);  // version made by
    central.insert(central.end(), local.begin() + 4, local.begin() + 30);  // needed .. extra length
    detail::put<uint16_t>(central, 0);   // comment length
    detail::put<uint16_t>(central, 0);   // disk number
    detail::put<uint16_t>(central, 0);   // internal attributes
    detail::put<uint32_t>(central, 0);   // external attributes
    detail::put<uint32_t>(central, offset);
    central.insert(central.end(), fname.begin(), fname.end());

    std::fwrite(local.data(), 1, local.size(), fp);
    std::fwrite(header.data(), 1, header.size(), fp);
    std::fwrite(kv.second.bytes.data(), 1, kv.second.bytes.size(), fp);
    offset += (uint32_t)local.size() + size;
    ++count;
  }
  std::vector<char> eocd;
  detail::put<uint32_t>(eocd, 0x06054b50u);
  detail::put<uint16_t>(eocd, 0);
  detail::put<uint16_t>(eocd, 0);
  detail::put<uint16_t>(eocd, count);
  detail::put<uint16_t>(eocd, count);
  detail::put<uint32_t>(eocd, (uint32_t)central.size());
  detail::put<uint32_t>(eocd, offset);
  detail::put<uint16_t>(eocd, 0);
  std::fwrite(central.data(), 1, central.size(), fp);
  std::fwrite(eocd.data(), 1, eocd.size(), fp);
  std::fclose(fp);
}

// Members in archive order (walks the local file headers, as cnpy::npz_load does).
inline std::vector<std::pair<std::string, Array>> load(const std::string& path) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  ABORT_IF(!fp, "npz: cannot open:", path);
  std::vector<std::pair<std::string, Array>> out;
  while(true) {
    char h[30];
    if(std::fread(h, 1, 30, fp) != 30)
      break;
    uint32_t sig = detail::get<uint32_t>(h);
    if(sig != 0x04034b50u)
      break;  // central directory reached
    uint16_t flags = detail::get<uint16_t>(h + 6), method = detail::get<uint16_t>(h + 8);
    uint32_t csize = detail::get<uint32_t>(h + 18), usize = detail::get<uint32_t>(h + 22);  // 0xFFFFFFFF: see the ZIP64 record
    uint16_t nlen = detail::get<uint16_t>(h + 26), xlen = detail::get<uint16_t>(h + 28);
    std::string fname(nlen, '\0');
    ABORT_IF(std::fread(&fname[0], 1, nlen, fp) != nlen, "npz: truncated archive:", path);
    std::vector<char> extra(xlen);
    ABORT_IF(xlen && std::fread(extra.data(), 1, xlen, fp) != xlen, "npz: truncated archive:", path);
    uint64_t usize64 = usize, csize64 = csize;
    if(usize == 0xFFFFFFFFu || csize == 0xFFFFFFFFu) {
      // ZIP64 extended information (header id 1): numpy.savez writes every member this way
      bool found = false;
      for(size_t p = 0; p + 4 <= extra.size();) {
        uint16_t id = detail::get<uint16_t>(extra.data() + p), len = detail::get<uint16_t>(extra.data() + p + 2);
        if(id == 1 && len >= 16 && p + 4 + 16 <= extra.size()) {
          usize64 = detail::get<uint64_t>(extra.data() + p + 4);
          csize64 = detail::get<uint64_t>(extra.data() + p + 12);
          found = true;
          break;
        }
        p += 4 + (size_t)len;
      }
      ABORT_IF(!found, "npz: ZIP64 member without size record:", fname);
    }
    ABORT_IF(method != 0, "npz: compressed member", fname, "- only stored archives (numpy.savez, Marian) are supported");
    ABORT_IF((flags & 8) || csize64 != usize64, "npz: streamed member", fname, "is not supported");
    std::vector<char> data(usize64);
    ABORT_IF(std::fread(data.data(), 1, usize64, fp) != usize64, "npz: truncated member", fname);
    std::string name = fname.size() > 4 && fname.substr(fname.size() - 4) == ".npy" ? fname.substr(0, fname.size() - 4) : fname;
    out.push_back({name, detail::parseNpy(data.data(), data.size(), fname)});
  }
  std::fclose(fp);
  ABORT_IF(out.empty(), "npz: no arrays found in", path);
  return out;
}

}  // namespace npz
}  // namespace marian
