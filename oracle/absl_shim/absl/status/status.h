// Build scaffolding for oracle/_ref ONLY.
#pragma once
#include <string>
namespace absl {
class Status {
 public:
  Status() = default;
  bool ok() const { return true; }
};
inline Status OkStatus() { return Status(); }
}  // namespace absl
