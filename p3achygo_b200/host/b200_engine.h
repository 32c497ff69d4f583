// nn::B200Engine — the adapter a p3achygo maintainer drops next to trt_engine.{h,cc}: an nn::Engine
// whose four virtuals forward to the C ABI of libp3b200.so.  Errors are fatal, as in the reference
// (CUDA_OK / CHECK abort, cc/nn/engine/trt_engine.cc:27-35).
#pragma once
#include "engine_iface.h"

namespace nn {

class B200Engine final : public Engine {
 public:
  // precision: P3_PRECISION_BF16 (default; env P3_PRECISION=fp32 selects the parity path, fp16 the fp16-operand tensor path)
  static std::unique_ptr<B200Engine> Create(std::string path, int batch_size, int version, int device = 0,
                                            int precision = -1);
  ~B200Engine() override;

  Engine::Kind kind() override { return Engine::Kind::kB200; }
  std::string path() override { return path_; }
  void LoadBatch(int batch_id, const GoFeatures& features) override;
  void RunInference() override;
  void GetBatch(int batch_id, NNInferResult& result) override;
  void GetOwnership(int batch_id, std::array<float, P3_NUM_BOARD_LOCS>& own) override;
  // NNInterface::LoadBatch / GetBatch with the symmetry handled on the GPU (nn_interface.cc:245-277, nn_interface.h:263-287):
  // features in the game's own orientation + the symmetry; GetBatch then returns the un-rotated policies.
  void LoadBatchSym(int batch_id, const GoFeatures& features, int sym);

  // Pipelined form over the engine's two slot banks (include/p3_b200.h): fill one bank while the other is in flight.
  // RunInference() == Submit(0) + Wait(0) in effect; results are identical.
  static constexpr int kNumBanks = P3_NUM_BANKS;
  void LoadBatchBank(int bank, int batch_id, const GoFeatures& features, int sym = 0);
  // NNInterface::LoadBatch from the game record (nn_interface.cc:245-277): moves as p3_game_derive encodes them; the board, the
  // liberty grids, the laddered stones and the last moves are derived on the GPU by the next run of that bank.
  void LoadGameBank(int bank, int batch_id, const int16_t* moves, int num_moves, int color, float komi, const int8_t* forbidden, int sym);
  // The same for the serial RunInference cycle, in the shape of INTEGRATION.md's optional edit 5 (nn::Engine::LoadGameRecord, a
  // virtual with a `return false` default that NNInterface::LoadBatch tries before it computes GoFeatures on the host): overrides
  // it where the base class has it.  Pass-alive regions are derived from the record; results come back un-rotated.
  // false = not loaded (record too long, ...): the caller falls back to LoadBatch.
  bool LoadGameRecord(int batch_id, const int16_t* moves, int num_moves, int color_to_move, float komi, int sym);
  void Submit(int bank);
  void Wait(int bank);
  void GetBatchBank(int bank, int batch_id, NNInferResult& result);
  // GetOwnership for a slot of a waited bank (each bank keeps its own copy of the auxiliary outputs)
  void GetOwnershipBank(int bank, int batch_id, std::array<float, P3_NUM_BOARD_LOCS>& own);
  // Compact leaf records: SetLeafResults(true) makes runs copy back one p3_leaf_result (4360 B: the three policy
  // arrays + value / E[score] / Var[score] / err of mcts::LeafEvaluator's InitFields, cc/mcts/leaf_evaluator.cc:83-112) per
  // slot instead of the 7568-byte NNInferResult; read them with GetLeaf / GetLeafBank.
  void SetLeafResults(bool enabled);
  void GetLeaf(int batch_id, p3_leaf_result& leaf);
  void GetLeafBank(int bank, int batch_id, p3_leaf_result& leaf);

  p3_engine* handle() { return engine_; }
  int batch_size() const { return batch_size_; }

 private:
  B200Engine(p3_engine* e, std::string path, int batch_size) : engine_(e), path_(std::move(path)), batch_size_(batch_size) {}
  p3_engine* engine_;
  std::string path_;
  int batch_size_;
};

}  // namespace nn
