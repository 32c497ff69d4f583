// Host-side mirror of the reference's engine boundary, for builds where the reference tree (and
// abseil) is not present.  Same names, members and call contract as
//   nn::Engine          cc/nn/engine/engine.h:22-43
//   nn::NNInferResult   cc/nn/engine/engine.h:12-20     (layout == p3_infer_result, 7568 B)
//   nn::GoFeatures      cc/nn/engine/go_features.h:12-22 (layout == p3_go_features, 1860 B)
// In the reference tree the adapter derives from the real nn::Engine instead.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <string>

#include "p3_b200.h"

#ifdef P3_REFERENCE_TREE
// Inside the reference tree (oracle/Makefile builds this configuration against the unmodified reference
// sources): the real nn::Engine / GoFeatures / NNInferResult, with Engine::Kind::kB200 added by edit 1 of INTEGRATION.md.
#include "cc/nn/engine/engine.h"

namespace nn {
static_assert(sizeof(GoFeatures) == sizeof(p3_go_features), "nn::GoFeatures and p3_go_features must be the same bytes");
static_assert(sizeof(NNInferResult) == sizeof(p3_infer_result) && alignof(NNInferResult) == 16, "nn::NNInferResult layout");
inline const p3_go_features* AsC(const GoFeatures& f) { return reinterpret_cast<const p3_go_features*>(&f); }
inline p3_infer_result* AsC(NNInferResult& r) { return reinterpret_cast<p3_infer_result*>(&r); }
}  // namespace nn
#else

namespace nn {

using GoFeatures = ::p3_go_features;
using NNInferResult = ::p3_infer_result;
static_assert(sizeof(GoFeatures) == 1860, "GoFeatures mirror");
static_assert(sizeof(NNInferResult) == 7568 && alignof(NNInferResult) == 16, "NNInferResult mirror");
inline const p3_go_features* AsC(const GoFeatures& f) { return &f; }
inline p3_infer_result* AsC(NNInferResult& r) { return &r; }

class Engine {
 public:
  enum class Kind : uint8_t { kUnknown = 0, kTrt = 1, kTF = 2, kTFTrt = 3, kTFXla = 4, kB200 = 5 };
  virtual ~Engine() = default;
  virtual Kind kind() = 0;
  virtual std::string path() = 0;
  // worker threads, concurrent for distinct ids, no lock held; may overlap RunInference
  virtual void LoadBatch(int batch_id, const GoFeatures& features) = 0;
  // one caller at a time; always evaluates the full batch
  virtual void RunInference() = 0;
  // worker threads, concurrent for distinct ids; never overlaps RunInference
  virtual void GetBatch(int batch_id, NNInferResult& result) = 0;
  virtual void GetOwnership(int batch_id, std::array<float, P3_NUM_BOARD_LOCS>& own) = 0;

 protected:
  Engine() = default;
};

std::string KindToString(Engine::Kind kind);                                  // engine.h:45-57
Engine::Kind KindFromEnginePath(std::string path);                            // engine_factory.cc:16-35
int GetVersionFromModelPath(std::string path);                                // engine_factory.cc:37-54
std::unique_ptr<Engine> CreateEngine(Engine::Kind kind, std::string path, int batch_size, int version);  // :56-73

}  // namespace nn
#endif  // P3_REFERENCE_TREE
