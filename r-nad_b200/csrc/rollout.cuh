// Argument block shared by the fused rollout kernels (K2).
#pragma once

#include "game.cuh"

namespace rnad {

struct RolloutArgs {
    const uint32_t* ev_tab;
    const uint32_t* tr_tab;
    int A, C;
    rnad_mlp_weights w;
    int64_t B;
    int T;
    uint64_t seed;
    int64_t game_offset;
    const float* uniforms;   // (T,B,2) or nullptr
    TrajPtrs out;
    int32_t* t_last;
};

int rollout_fp32(const RolloutArgs& g, cudaStream_t st);
int rollout_tc(const RolloutArgs& g, void* workspace, cudaStream_t st);
bool rollout_tc_supported(int A, int width);
int64_t rollout_tc_workspace_bytes(int A);
int rollout_tc2(const RolloutArgs& g, void* workspace, cudaStream_t st);
int64_t rollout_tc2_workspace_bytes(int A);
bool rollout_tc2_supported(int A, int width, int C);

}  // namespace rnad
