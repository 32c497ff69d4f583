// Reader for the flat P3W1 weight file (format: p3achygo_b200/weights.py; tag tree of
// python/export_weights.py:16-90).  Host-only, header-only.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace p3 {

struct WeightTensor {
  std::vector<uint32_t> dims;
  std::vector<float> data;
  size_t numel() const {
    size_t n = 1;
    for (uint32_t d : dims) n *= d;
    return n;
  }
};

struct WeightFile {
  std::map<std::string, int32_t> meta;
  std::map<std::string, WeightTensor> tensors;

  // returns empty string on success, else an error message
  std::string load(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return "cannot open " + path;
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    size_t off = 0;
    auto need = [&](size_t k) { return off + k <= buf.size(); };
    if (!need(8) || std::memcmp(buf.data(), "P3ACHYW1", 8) != 0) return path + ": not a P3W1 weight file";
    off = 8;
    auto rd_u32 = [&](uint32_t& v) { if (!need(4)) return false; std::memcpy(&v, &buf[off], 4); off += 4; return true; };
    auto rd_u16 = [&](uint16_t& v) { if (!need(2)) return false; std::memcpy(&v, &buf[off], 2); off += 2; return true; };
    auto rd_str = [&](std::string& s) {
      uint16_t len;
      if (!rd_u16(len) || !need(len)) return false;
      s.assign(reinterpret_cast<const char*>(&buf[off]), len);
      off += len;
      return true;
    };
    uint32_t n_meta;
    if (!rd_u32(n_meta)) return path + ": truncated";
    for (uint32_t i = 0; i < n_meta; ++i) {
      std::string key;
      uint32_t val;
      if (!rd_str(key) || !rd_u32(val)) return path + ": truncated meta";
      meta[key] = static_cast<int32_t>(val);
    }
    uint32_t n_t;
    if (!rd_u32(n_t)) return path + ": truncated";
    for (uint32_t i = 0; i < n_t; ++i) {
      std::string name;
      if (!rd_str(name) || !need(1)) return path + ": truncated tensor header";
      const uint8_t nd = buf[off++];
      if (nd > 8) return path + ": tensor " + name + " has " + std::to_string(nd) + " dimensions";
      WeightTensor t;
      t.dims.resize(nd);
      for (uint8_t d = 0; d < nd; ++d)
        if (!rd_u32(t.dims[d])) return path + ": truncated dims";
      off += (4 - off % 4) % 4;
      if (off > buf.size()) return path + ": truncated data for " + name;
      // the element count is bounded by what is left of the file BEFORE any multiplication can wrap
      const size_t max_cnt = (buf.size() - off) / 4;
      size_t cnt = 1;
      for (uint32_t dim : t.dims) {
        if (dim != 0 && cnt > max_cnt / dim) return path + ": truncated data for " + name;
        cnt *= dim;
      }
      if (cnt > max_cnt) return path + ": truncated data for " + name;
      t.data.resize(cnt);
      std::memcpy(t.data.data(), &buf[off], cnt * 4);
      off += cnt * 4;
      tensors[name] = std::move(t);
    }
    return "";
  }

  const WeightTensor* find(const std::string& name) const {
    auto it = tensors.find(name);
    return it == tensors.end() ? nullptr : &it->second;
  }
  int meta_or(const std::string& key, int dflt) const {
    auto it = meta.find(key);
    return it == meta.end() ? dflt : it->second;
  }
};

}  // namespace p3
