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
    const uint64_t* seed_dev;   // if non-null: the seed is read from device memory (a captured step varies it)
    int64_t game_offset;
    const float* uniforms;   // (T,B,2) or nullptr
    TrajPtrs out;
    int32_t* stats;          // device int32[4]: longest game in half-moves (t_eff + 1), valid slots of player 0, of player 1, -
};

// end of a rollout thread's work: the longest game it saw (last valid half-move, -1 = none) and its valid slots per player
__device__ __forceinline__ void publish_stats(int32_t* stats, int last_valid, int n0, int n1, int lane) {
    last_valid = warp_max(last_valid);
    n0 = warp_sum(n0);
    n1 = warp_sum(n1);
    if (lane == 0) {
        if (last_valid >= 0) atomicMax(stats, last_valid + 1);
        if (n0) atomicAdd(stats + 1, n0);
        if (n1) atomicAdd(stats + 2, n1);
    }
}

int rollout_fp32(const RolloutArgs& g, cudaStream_t st);
int rollout_tc(const RolloutArgs& g, void* workspace, cudaStream_t st);
bool rollout_tc_supported(int A, int width);
int64_t rollout_tc_workspace_bytes(int A);
int rollout_tc2(const RolloutArgs& g, void* workspace, cudaStream_t st, bool f16);
int64_t rollout_tc2_workspace_bytes(int A);
bool rollout_tc2_supported(int A, int width, int C);

}  // namespace rnad
