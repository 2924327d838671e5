// Fused on-chip N = 128 multislice kernels (placeholder until the register-resident path lands).
#pragma once
#include "general_kernels.cuh"
#include "../../include/ptyrad_b200.h"
#include <string>

namespace ptyb { namespace fused128 {
inline size_t scratch_bytes(const ptyb200_cfg&, int) { return 0; }
inline bool covers(const ptyb200_cfg&) { return false; }
inline int forward(const ptyb200_cfg&, int, FwdArgs, unsigned char*, cudaStream_t, std::string& err) { err = "fused path not built"; return 3; }
inline int backward(const ptyb200_cfg&, int, BwdArgs, unsigned char*, float2*, float2*, cudaStream_t, std::string& err) { err = "fused path not built"; return 3; }
}}
