#include "b200_engine.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>

namespace nn {
namespace {
namespace fs = std::filesystem;

// The reference aborts on engine errors (trt_engine.cc:27-35); so does the adapter.
#define P3_CHECK(call)                                                                       \
  do {                                                                                       \
    int _rc = (call);                                                                        \
    if (_rc != 0) {                                                                          \
      std::fprintf(stderr, "libp3b200 error %d at %s:%d: %s\n", _rc, __FILE__, __LINE__, p3_last_error()); \
      std::abort();                                                                          \
    }                                                                                        \
  } while (0)
}  // namespace

std::unique_ptr<B200Engine> B200Engine::Create(std::string path, int batch_size, int version, int device, int precision) {
  if (precision < 0) {
    const char* env = std::getenv("P3_PRECISION");
    precision = (env && std::strcmp(env, "fp32") == 0) ? P3_PRECISION_FP32
                : (env && std::strcmp(env, "fp16") == 0) ? P3_PRECISION_FP16 : P3_PRECISION_BF16;
  }
  p3_engine* e = nullptr;
  P3_CHECK(p3_engine_create(path.c_str(), device, batch_size, version, precision, &e));
  return std::unique_ptr<B200Engine>(new B200Engine(e, std::move(path), batch_size));
}

B200Engine::~B200Engine() { p3_engine_destroy(engine_); }
void B200Engine::LoadBatch(int batch_id, const GoFeatures& features) { P3_CHECK(p3_engine_load_batch(engine_, batch_id, AsC(features))); }
void B200Engine::LoadBatchSym(int batch_id, const GoFeatures& features, int sym) {
  P3_CHECK(p3_engine_load_batch_sym(engine_, batch_id, AsC(features), sym));
}
void B200Engine::RunInference() { P3_CHECK(p3_engine_run_inference(engine_)); }
void B200Engine::GetBatch(int batch_id, NNInferResult& result) { P3_CHECK(p3_engine_get_batch(engine_, batch_id, AsC(result))); }
void B200Engine::LoadBatchBank(int bank, int batch_id, const GoFeatures& features, int sym) {
  P3_CHECK(p3_engine_load_batch_bank(engine_, bank, batch_id, AsC(features), sym));
}
void B200Engine::LoadGameBank(int bank, int batch_id, const int16_t* moves, int num_moves, int color, float komi,
                              const int8_t* forbidden, int sym) {
  P3_CHECK(p3_engine_load_game_bank(engine_, bank, batch_id, moves, num_moves, color, komi, forbidden, sym));
}
bool B200Engine::LoadGameRecord(int batch_id, const int16_t* moves, int num_moves, int color_to_move, float komi, int sym) {
  return p3_engine_load_game_bank(engine_, 0, batch_id, moves, num_moves, color_to_move, komi, nullptr, sym) == P3_OK;
}
void B200Engine::Submit(int bank) { P3_CHECK(p3_engine_submit(engine_, bank)); }
void B200Engine::Wait(int bank) { P3_CHECK(p3_engine_wait(engine_, bank)); }
void B200Engine::GetBatchBank(int bank, int batch_id, NNInferResult& result) {
  P3_CHECK(p3_engine_get_batch_bank(engine_, bank, batch_id, AsC(result)));
}
void B200Engine::GetOwnershipBank(int bank, int batch_id, std::array<float, P3_NUM_BOARD_LOCS>& own) {
  P3_CHECK(p3_engine_get_ownership_bank(engine_, bank, batch_id, own.data()));
}
void B200Engine::SetLeafResults(bool enabled) { P3_CHECK(p3_engine_set_result_mode(engine_, enabled ? P3_RESULT_LEAF : P3_RESULT_FULL)); }
void B200Engine::GetLeaf(int batch_id, p3_leaf_result& leaf) { P3_CHECK(p3_engine_get_leaf(engine_, batch_id, &leaf)); }
void B200Engine::GetLeafBank(int bank, int batch_id, p3_leaf_result& leaf) { P3_CHECK(p3_engine_get_leaf_bank(engine_, bank, batch_id, &leaf)); }
void B200Engine::GetOwnership(int batch_id, std::array<float, P3_NUM_BOARD_LOCS>& own) {
  P3_CHECK(p3_engine_get_ownership(engine_, batch_id, own.data()));
}

#ifndef P3_REFERENCE_TREE  // the reference tree has its own KindToString (engine.h:45-57) and engine factory (engine_factory.cc)
std::string KindToString(Engine::Kind kind) {
  switch (kind) {
    case Engine::Kind::kTrt: return "TensorRT";
    case Engine::Kind::kTF: return "TF";
    case Engine::Kind::kTFTrt: return "TF-TRT";
    case Engine::Kind::kTFXla: return "TF-XLA";
    case Engine::Kind::kB200: return "B200";
    default: return "??";
  }
}

// engine_factory.cc:16-35 plus one rule: a regular file ending in ".p3w" is a flat P3W1 weight file.
Engine::Kind KindFromEnginePath(std::string path) {
  fs::path filepath(path);
  if (fs::is_regular_file(filepath)) {
    const auto ext = filepath.extension();
    if (ext == ".trt") return Engine::Kind::kTrt;
    if (ext == ".pb") return Engine::Kind::kTFXla;
    if (ext == ".p3w") return Engine::Kind::kB200;
    return Engine::Kind::kUnknown;
  }
  if (filepath.filename() == "_trt") return Engine::Kind::kTFTrt;
  return Engine::Kind::kTF;
}

int GetVersionFromModelPath(std::string path) {
  fs::path filepath(path);
  fs::path dir = fs::is_regular_file(filepath) ? filepath.parent_path() : filepath;
  fs::path version_file = dir / "VERSION";
  if (fs::exists(version_file) && fs::is_regular_file(version_file)) {
    std::ifstream ifs(version_file);
    int version;
    if (ifs >> version) return version;
    std::fprintf(stderr, "Failed to parse VERSION file at %s, defaulting to version 1\n", version_file.c_str());
  }
  return 1;
}

std::unique_ptr<Engine> CreateEngine(Engine::Kind kind, std::string path, int batch_size, int version) {
  switch (kind) {
    case Engine::Kind::kB200:
      return B200Engine::Create(path, batch_size, version);
    default:
      std::fprintf(stderr, "Unknown Engine Kind.\n");  // LOG(FATAL), engine_factory.cc:70
      std::abort();
  }
}
#endif  // !P3_REFERENCE_TREE

}  // namespace nn
