// C-ABI entry of the fused rollout (K2): argument checks and engine dispatch.
#include "rollout.cuh"

using namespace rnad;

extern "C" int rnad_rollout(const uint32_t* ev_tab, const uint32_t* tr_tab, int A, int C, const rnad_mlp_weights* w,
                            int64_t B, int T, uint64_t seed, const uint64_t* seed_dev, int64_t game_offset,
                            const float* uniforms, int precision, const rnad_trajectory* out, int32_t* stats,
                            void* workspace, void* stream) {
    RNAD_REQUIRE(ev_tab && tr_tab && w && out && stats, "rnad_rollout: null pointer");
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS, "rnad_rollout: max_actions %d outside [1,%d]", A, RNAD_MAX_ACTIONS);
    RNAD_REQUIRE(C >= 1 && C <= RNAD_MAX_TRANSITIONS, "rnad_rollout: max_transitions %d outside [1,%d]", C,
                 RNAD_MAX_TRANSITIONS);
    RNAD_REQUIRE(B >= 0 && T >= 0, "rnad_rollout: negative batch or horizon");
    RNAD_REQUIRE(w->width >= 1, "rnad_rollout: bad net width %d", w->width);
    RNAD_REQUIRE(w->value_fc0_w && w->value_fc0_b && w->value_fc1_w && w->value_fc1_b && w->policy_fc0_w &&
                     w->policy_fc0_b && w->policy_fc1_w && w->policy_fc1_b,
                 "rnad_rollout: null weight pointer");
    RNAD_REQUIRE(out->indices && out->turns && out->observations && out->policy && out->actions && out->rewards &&
                     out->values && out->masks,
                 "rnad_rollout: null trajectory pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // the kernels atomicMax / atomicAdd into stats: start from zero here, in stream order (also for an empty batch)
    int rc = check_cuda(cudaMemsetAsync(stats, 0, 4 * sizeof(int32_t), st), "cudaMemsetAsync(stats)");
    if (rc) return rc;
    if (B == 0 || T == 0) return RNAD_OK;

    RolloutArgs g;
    g.ev_tab = ev_tab;
    g.tr_tab = tr_tab;
    g.A = A;
    g.C = C;
    g.w = *w;
    g.B = B;
    g.T = T;
    g.seed = seed;
    g.seed_dev = seed_dev;
    g.game_offset = game_offset;
    g.uniforms = uniforms;
    g.out = TrajPtrs{out->indices, out->turns, out->observations, out->policy,
                     out->actions, out->rewards, out->values, out->masks, out->logits, out->returns};
    g.stats = stats;
    switch (precision) {
        case RNAD_PREC_FP32: return rollout_fp32(g, st);
        case RNAD_PREC_TF32: return rollout_tc(g, workspace, st);
        case RNAD_PREC_TF32X2: return rollout_tc2(g, workspace, st, false);
        case RNAD_PREC_F16X2: return rollout_tc2(g, workspace, st, true);
    }
    set_error("rnad_rollout: unknown precision %d", precision);
    return RNAD_EINVAL;
}

// 1 if the tensor-core engine serves this net shape, else 0
extern "C" int rnad_rollout_tc_supported(int A, int width) { return rollout_tc_supported(A, width) ? 1 : 0; }

// 1 if the RNAD_PREC_TF32X2 engine serves this net and tree shape, else 0
extern "C" int rnad_rollout_tc2_supported(int A, int width, int C) { return rollout_tc2_supported(A, width, C) ? 1 : 0; }

// bytes of device scratch rnad_rollout needs for this net shape and engine (0 = none)
extern "C" int64_t rnad_rollout_workspace_bytes(int A, int width, int precision) {
    if (precision == RNAD_PREC_TF32 && rollout_tc_supported(A, width)) return rollout_tc_workspace_bytes(A);
    if ((precision == RNAD_PREC_TF32X2 || precision == RNAD_PREC_F16X2) && rollout_tc2_supported(A, width, 1)) return rollout_tc2_workspace_bytes(A);
    return 0;
}
