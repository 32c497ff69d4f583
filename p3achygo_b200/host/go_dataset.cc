#include "go_dataset.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>

namespace nn {

uint32_t Crc32c(const void* data, size_t n) {  // Castagnoli, reflected 0x82F63B78
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  uint32_t c = 0xFFFFFFFFu;
  const uint8_t* p = static_cast<const uint8_t*>(data);
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

uint32_t MaskedCrc32c(const void* data, size_t n) {  // TFRecord's mask: rotate right by 15, add a constant
  const uint32_t c = Crc32c(data, n);
  return ((c >> 15) | (c << 17)) + 0xA282EAD8u;
}

namespace {

bool ReadFile(const std::string& path, std::string* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  out->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  return true;
}

bool LooksLikeTfRecord(const std::string& s) {
  if (s.size() < 12) return s.empty();
  uint32_t crc;
  std::memcpy(&crc, s.data() + 8, 4);
  return crc == MaskedCrc32c(s.data(), 8);
}

bool Inflate(const std::string& in, std::string* out) {  // zlib or gzip wrapper, whole file
  z_stream zs{};
  if (inflateInit2(&zs, 15 + 32) != Z_OK) return false;
  zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in.data()));
  zs.avail_in = static_cast<uInt>(in.size());
  std::vector<char> buf(1 << 20);
  int rc = Z_OK;
  while (rc != Z_STREAM_END) {
    zs.next_out = reinterpret_cast<Bytef*>(buf.data());
    zs.avail_out = static_cast<uInt>(buf.size());
    rc = inflate(&zs, Z_NO_FLUSH);
    if (rc != Z_OK && rc != Z_STREAM_END) {
      inflateEnd(&zs);
      return false;
    }
    out->append(buf.data(), buf.size() - zs.avail_out);
    if (rc == Z_STREAM_END && zs.avail_in > 0) {  // concatenated streams
      if (inflateReset(&zs) != Z_OK) break;
      rc = Z_OK;
    }
  }
  inflateEnd(&zs);
  return true;
}

// ---- the part of the protobuf wire format a tf.train.Example uses ----------------------------------------------------
struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; p < end && shift < 64; shift += 7) {
      const uint8_t b = *p++;
      v |= static_cast<uint64_t>(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  Cursor sub() {  // length-delimited field
    const uint64_t n = varint();
    if (!ok || n > static_cast<uint64_t>(end - p)) {
      ok = false;
      return Cursor{end, end, false};
    }
    Cursor c{p, p + n};
    p += n;
    return c;
  }
  void skip(int wire) {
    if (wire == 0) varint();
    else if (wire == 1) p += 8;
    else if (wire == 2) sub();
    else if (wire == 5) p += 4;
    else ok = false;
    if (p > end) ok = false;
  }
};

struct FeatureValue {
  std::string bytes;           // first BytesList value
  std::vector<float> floats;   // FloatList
  bool has_bytes = false;
};

// Example { Features features = 1 }   Features { map<string, Feature> feature = 1 }
// Feature { BytesList bytes_list = 1 | FloatList float_list = 2 | Int64List int64_list = 3 }
bool ParseExample(const std::string& rec, std::map<std::string, FeatureValue>* out) {
  Cursor ex{reinterpret_cast<const uint8_t*>(rec.data()), reinterpret_cast<const uint8_t*>(rec.data()) + rec.size()};
  while (!ex.done() && ex.ok) {
    const uint64_t tag = ex.varint();
    if ((tag >> 3) != 1 || (tag & 7) != 2) {
      ex.skip(tag & 7);
      continue;
    }
    Cursor feats = ex.sub();
    while (!feats.done() && feats.ok) {
      const uint64_t t2 = feats.varint();
      if ((t2 >> 3) != 1 || (t2 & 7) != 2) {
        feats.skip(t2 & 7);
        continue;
      }
      Cursor entry = feats.sub();  // map entry: key = 1, value = 2
      std::string key;
      FeatureValue val;
      while (!entry.done() && entry.ok) {
        const uint64_t t3 = entry.varint();
        if ((t3 & 7) != 2) {
          entry.skip(t3 & 7);
          continue;
        }
        Cursor f = entry.sub();
        if ((t3 >> 3) == 1) {
          key.assign(reinterpret_cast<const char*>(f.p), f.end - f.p);
        } else if ((t3 >> 3) == 2) {
          while (!f.done() && f.ok) {  // Feature
            const uint64_t t4 = f.varint();
            if ((t4 & 7) != 2) {
              f.skip(t4 & 7);
              continue;
            }
            Cursor list = f.sub();
            const int kind = static_cast<int>(t4 >> 3);
            while (!list.done() && list.ok) {
              const uint64_t t5 = list.varint();
              if ((t5 >> 3) != 1) {
                list.skip(t5 & 7);
                continue;
              }
              if (kind == 1 && (t5 & 7) == 2) {  // BytesList.value
                Cursor b = list.sub();
                if (!val.has_bytes) val.bytes.assign(reinterpret_cast<const char*>(b.p), b.end - b.p);
                val.has_bytes = true;
              } else if (kind == 2 && (t5 & 7) == 2) {  // FloatList.value, packed
                Cursor b = list.sub();
                for (const uint8_t* q = b.p; q + 4 <= b.end; q += 4) {
                  float v;
                  std::memcpy(&v, q, 4);
                  val.floats.push_back(v);
                }
              } else if (kind == 2 && (t5 & 7) == 5) {  // FloatList.value, unpacked
                float v;
                if (list.p + 4 > list.end) { list.ok = false; break; }
                std::memcpy(&v, list.p, 4);
                list.p += 4;
                val.floats.push_back(v);
              } else {
                list.skip(t5 & 7);
              }
            }
            if (!list.ok) f.ok = false;
          }
          if (!f.ok) entry.ok = false;
        }
      }
      if (!entry.ok) return false;
      (*out)[key] = std::move(val);
    }
    if (!feats.ok) return false;
  }
  return ex.ok;
}

bool BytesInto(const std::map<std::string, FeatureValue>& m, const char* key, void* dst, size_t n) {
  auto it = m.find(key);
  if (it == m.end() || !it->second.has_bytes || it->second.bytes.size() < n) return false;
  std::memcpy(dst, it->second.bytes.data(), n);
  return true;
}

bool FloatInto(const std::map<std::string, FeatureValue>& m, const char* key, float* dst) {
  auto it = m.find(key);
  if (it == m.end() || it->second.floats.empty()) return false;
  *dst = it->second.floats[0];
  return true;
}

}  // namespace

GoDataset::GoDataset(size_t batch_size, std::string ds_path) : batch_size_(batch_size) {
  std::string raw, data;
  if (!ReadFile(ds_path, &raw)) {
    std::fprintf(stderr, "Failed to initialize reader for: %s\n", ds_path.c_str());  // CHECK, go_dataset.cc:36
    std::abort();
  }
  if (LooksLikeTfRecord(raw)) {
    data.swap(raw);
  } else if (!Inflate(raw, &data)) {  // RecordReaderOptions::Zlib(), go_dataset.cc:35
    std::fprintf(stderr, "Failed to initialize reader for: %s (neither TFRecord framing nor a zlib stream)\n", ds_path.c_str());
    std::abort();
  }
  std::vector<Row> batch;
  size_t pos = 0;
  int index = 0;
  while (pos + 12 <= data.size()) {
    uint64_t len;
    uint32_t len_crc;
    std::memcpy(&len, data.data() + pos, 8);
    std::memcpy(&len_crc, data.data() + pos + 8, 4);
    if (len_crc != MaskedCrc32c(data.data() + pos, 8) || len > data.size() - pos - 12 || data.size() - pos - 12 - len < 4) {
      std::fprintf(stderr, "Error reading TFRecord %d\n", index);  // go_dataset.cc:49-51
      break;
    }
    const std::string rec = data.substr(pos + 12, len);
    uint32_t data_crc;
    std::memcpy(&data_crc, data.data() + pos + 12 + len, 4);
    pos += 12 + len + 4;
    ++index;
    if (data_crc != MaskedCrc32c(rec.data(), rec.size())) {
      std::fprintf(stderr, "Error reading TFRecord %d\n", index - 1);
      continue;
    }
    std::map<std::string, FeatureValue> ex;
    Row row{};
    uint8_t bsize = 0;
    int16_t last_moves[P3_NUM_LAST_MOVES];
    bool ok = ParseExample(rec, &ex);
    ok = ok && BytesInto(ex, "bsize", &bsize, 1) && BytesInto(ex, "board", row.features.board, P3_NUM_BOARD_LOCS) &&
         BytesInto(ex, "last_moves", last_moves, sizeof(last_moves)) &&
         BytesInto(ex, "stones_atari", row.features.stones_atari, P3_NUM_BOARD_LOCS) &&
         BytesInto(ex, "stones_two_liberties", row.features.stones_two_liberties, P3_NUM_BOARD_LOCS) &&
         BytesInto(ex, "stones_three_liberties", row.features.stones_three_liberties, P3_NUM_BOARD_LOCS) &&
         BytesInto(ex, "stones_in_ladder", row.features.stones_laddered, P3_NUM_BOARD_LOCS) &&
         BytesInto(ex, "color", &row.features.color, 1) && BytesInto(ex, "pi", row.labels.policy.data(), sizeof(float) * P3_MAX_MOVES) &&
         FloatInto(ex, "score_margin", &row.labels.score_margin) && FloatInto(ex, "komi", &row.features.komi);
    if (!ok) {
      std::fprintf(stderr, "Error parsing TFRecord%d\n", index - 1);  // go_dataset.cc:54-57
      continue;
    }
    if (bsize != P3_BOARD_LEN) {  // CHECK, go_dataset.cc:82
      std::fprintf(stderr, "GoDataset: bsize %d != %d\n", bsize, P3_BOARD_LEN);
      std::abort();
    }
    row.features.bsize = bsize;
    for (int k = 0; k < P3_NUM_LAST_MOVES; ++k)  // game::AsLoc(encoding), cc/game/loc.h:29-31 (C++ division: -1 -> {0, -1})
      row.features.last_moves[k] = p3_loc{last_moves[k] / P3_BOARD_LEN, last_moves[k] % P3_BOARD_LEN};
    row.labels.did_win = row.labels.score_margin >= 0;  // go_dataset.cc:118
    batch.push_back(row);
    ++num_examples_;
    if (batch.size() == batch_size_) {
      batches_.push_back(std::move(batch));
      batch.clear();
    }
  }
  // the reference keeps the trailing partial batch padded to batch_size with default rows (go_dataset.cc:40-41, 125-126)
  if (!batch.empty() || batches_.empty()) {
    batch.resize(batch_size_, Row{});
    batches_.push_back(std::move(batch));
  }
}

}  // namespace nn

// ---- C entry points (tests, fixture writer) ------------------------------------------------------------------------------
extern "C" {

// Reads `path` with batch size `batch`; fills up to `cap` rows: features [cap] (1860 B each), policy [cap][362], score_margin [cap],
// did_win [cap].  Returns the number of examples parsed (rows beyond cap are counted, not stored).
long long p3_host_dataset_read(const char* path, int batch, int cap, p3_go_features* features, float* policy, float* score_margin,
                               unsigned char* did_win, long long* n_batches) {
  nn::GoDataset ds(static_cast<size_t>(batch), path);
  long long i = 0;
  const long long n = static_cast<long long>(ds.num_examples());
  for (auto& b : ds)
    for (auto& row : b) {
      if (i >= n || i >= cap) break;
      features[i] = row.features;
      std::memcpy(policy + i * P3_MAX_MOVES, row.labels.policy.data(), sizeof(float) * P3_MAX_MOVES);
      score_margin[i] = row.labels.score_margin;
      did_win[i] = row.labels.did_win ? 1 : 0;
      ++i;
    }
  if (n_batches) *n_batches = static_cast<long long>(ds.size());
  return n;
}

// Appends one TFRecord frame (length, masked crc, payload, masked crc) to `out` (capacity out_cap); returns the new size or -1.
long long p3_host_tfrecord_frame(const void* payload, long long n, unsigned char* out, long long at, long long out_cap) {
  if (at + 16 + n > out_cap) return -1;
  const uint64_t len = static_cast<uint64_t>(n);
  const uint32_t c1 = nn::MaskedCrc32c(&len, 8), c2 = nn::MaskedCrc32c(payload, static_cast<size_t>(n));
  std::memcpy(out + at, &len, 8);
  std::memcpy(out + at + 8, &c1, 4);
  std::memcpy(out + at + 12, payload, static_cast<size_t>(n));
  std::memcpy(out + at + 12 + n, &c2, 4);
  return at + 16 + n;
}

}  // extern "C"
