// k2_fused.cuh -- fused Gaborish + EPF + colour kernel (see DESIGN.md "K2").  Placeholder until the fused kernel lands:
// every frame takes the staged path in k2_restore.cuh.
#pragma once
#include "common.cuh"
struct jxlb200_ctx;
static inline int k2_fused_init(jxlb200_ctx *) { return 0; }
static inline bool k2_fused_supported(const K2Params &) { return false; }
static inline int k2_fused_launch(jxlb200_ctx *, const K2Params &, const float *) { return 0; }
