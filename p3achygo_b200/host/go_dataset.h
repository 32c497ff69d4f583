// Mirror of the reference's in-memory evaluation dataset nn::GoDataset (cc/nn/engine/go_dataset.{h,cc}): reads a TFRecord
// file of tf.train.Example protos - the format cc/recorder/make_tf_example.h:39-49 writes and python/test_data/generate.py
// restates - into fixed-size batches of {GoFeatures, GoLabels}.  The reference goes through TensorFlow's Example proto and its
// own record reader (cc/data/tfrecord/record_reader.h, zlib); neither is available here, so the TFRecord framing
// (length, masked CRC32C, payload, masked CRC32C), the optional zlib stream and the few proto fields needed are read directly.
#pragma once
#include <array>
#include <string>
#include <vector>

#include "engine_iface.h"

namespace nn {

struct GoLabels {  // cc/nn/engine/go_features.h:24-28
  std::array<float, P3_MAX_MOVES> policy;
  float score_margin;
  bool did_win;
};

class GoDataset final {
 public:
  struct Row {
    GoFeatures features;
    GoLabels labels;
  };
  // Aborts (the reference CHECKs) when the file cannot be read; records that fail to parse are skipped with a message, as
  // go_dataset.cc:47-57 does.
  GoDataset(size_t batch_size, std::string ds_path);
  size_t batch_size() const { return batch_size_; }
  size_t size() const { return batches_.size(); }
  size_t num_examples() const { return num_examples_; }
  std::vector<std::vector<Row>>::iterator begin() { return batches_.begin(); }
  std::vector<std::vector<Row>>::iterator end() { return batches_.end(); }

 private:
  std::vector<std::vector<Row>> batches_;
  size_t batch_size_;
  size_t num_examples_ = 0;
};

// TFRecord framing helpers (also used by the fixture writer behind p3_host_tfrecord_write)
uint32_t Crc32c(const void* data, size_t n);
uint32_t MaskedCrc32c(const void* data, size_t n);

}  // namespace nn
