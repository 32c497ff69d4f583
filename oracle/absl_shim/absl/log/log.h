// Build scaffolding for oracle/_ref ONLY.
#pragma once
#include "absl/log/check.h"
#define LOG(sev) ::absl_shim::LogStream(std::string(#sev) == "FATAL")
