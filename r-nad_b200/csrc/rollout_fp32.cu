// K2, fp32 validation build: Episodes.generate fused with MLP.forward, every
// multiply-add on the CUDA cores in fp32 (RNAD_PREC_FP32).  One thread per
// game, the whole net resident in shared memory, weights read as broadcast
// 16-byte loads and reused for four hidden units at a time.
// Reference: environment/episode.py:175-230, nn/net.py:37-51.
#include "game.cuh"
#include "rollout.cuh"

namespace rnad {

template <int A>
__global__ void __launch_bounds__(128) rollout_fp32_kernel(RolloutArgs g) {
    constexpr int KIN = 2 * A * A;
    constexpr int AP = round_up(A, 4);
    const int W = g.w.width;
    extern __shared__ __align__(16) float smem[];
    // transposed first layers [k][j] so that 4 consecutive hidden units are one float4
    float* w1v = smem;                    // KIN * W
    float* w1p = w1v + KIN * W;           // KIN * W
    float* b1v = w1p + KIN * W;           // W
    float* b1p = b1v + W;                 // W
    float* w2v = b1p + W;                 // W
    float* w2p = w2v + W;                 // W * AP   [j][a]
    float* b2 = w2p + W * AP;             // 1 + A (value bias, policy biases)

    for (int e = threadIdx.x; e < KIN * W; e += blockDim.x) {
        const int j = e / KIN, k = e % KIN;               // source is (W, KIN) row-major
        w1v[k * W + j] = g.w.value_fc0_w[e];
        w1p[k * W + j] = g.w.policy_fc0_w[e];
    }
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
        b1v[j] = g.w.value_fc0_b[j];
        b1p[j] = g.w.policy_fc0_b[j];
        w2v[j] = g.w.value_fc1_w[j];
        for (int a = 0; a < AP; ++a) w2p[j * AP + a] = a < A ? g.w.policy_fc1_w[a * W + j] : 0.f;
    }
    if (threadIdx.x == 0) b2[0] = g.w.value_fc1_b[0];
    if (threadIdx.x < A) b2[1 + threadIdx.x] = g.w.policy_fc1_b[threadIdx.x];
    __syncthreads();

    int last_valid = -1, n_valid0 = 0, n_valid1 = 0;
    const uint64_t seed = g.seed_dev != nullptr ? __ldg(g.seed_dev) : g.seed;
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < g.B; b += (int64_t)gridDim.x * blockDim.x) {
        int node = 1;   // every game starts at the root (episode.py:22)
        int row_action = 0;
        float game_return = 0.f;
        Node<A> n;
        for (int t = 0; t < g.T; ++t) {
            const int turn = t & 1;
            if (turn == 0) load_node<A>(g.ev_tab, node, n);
            if (node != 0) {
                last_valid = max(last_valid, t);
                n_valid0 += turn == 0;
                    n_valid1 += turn;
            }
            float x[KIN];
            build_obs<A>(n, turn, x);
            const int n_legal = turn == 0 ? n.rows : n.cols;
            const int64_t slot = (int64_t)t * g.B + b;
#pragma unroll
            for (int k = 0; k < KIN; ++k) st_stream(g.out.observations + slot * KIN + k, x[k]);

            float value = b2[0];
            float logit[A];
#pragma unroll
            for (int a = 0; a < A; ++a) logit[a] = b2[1 + a];
            for (int j = 0; j < W; j += 4) {
                float4 hv = *reinterpret_cast<const float4*>(b1v + j);
                float4 hp = *reinterpret_cast<const float4*>(b1p + j);
#pragma unroll
                for (int k = 0; k < KIN; ++k) {
                    const float4 wv = *reinterpret_cast<const float4*>(w1v + k * W + j);
                    const float4 wp = *reinterpret_cast<const float4*>(w1p + k * W + j);
                    hv.x = fmaf(x[k], wv.x, hv.x); hv.y = fmaf(x[k], wv.y, hv.y);
                    hv.z = fmaf(x[k], wv.z, hv.z); hv.w = fmaf(x[k], wv.w, hv.w);
                    hp.x = fmaf(x[k], wp.x, hp.x); hp.y = fmaf(x[k], wp.y, hp.y);
                    hp.z = fmaf(x[k], wp.z, hp.z); hp.w = fmaf(x[k], wp.w, hp.w);
                }
                const float4 ov = *reinterpret_cast<const float4*>(w2v + j);
                value = fmaf(fmaxf(hv.x, 0.f), ov.x, value);
                value = fmaf(fmaxf(hv.y, 0.f), ov.y, value);
                value = fmaf(fmaxf(hv.z, 0.f), ov.z, value);
                value = fmaf(fmaxf(hv.w, 0.f), ov.w, value);
                const float h[4] = {fmaxf(hp.x, 0.f), fmaxf(hp.y, 0.f), fmaxf(hp.z, 0.f), fmaxf(hp.w, 0.f)};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int a = 0; a < A; ++a) logit[a] = fmaf(h[u], w2p[(j + u) * AP + a], logit[a]);
            }

            float policy[A];
            masked_softmax<A>(logit, n_legal, policy);
            Uniforms2 u;
            if (g.uniforms != nullptr) {
                u.action = g.uniforms[slot * 2 + 0];
                u.chance = g.uniforms[slot * 2 + 1];
            } else {
                u = philox_uniforms(seed, (uint32_t)t, (uint64_t)(g.game_offset + b));
            }
            const int action = sample_icdf(policy, A, u.action);
            float reward = 0.f;
            const int node_now = node;
            if (turn == 0) {
                row_action = action;
            } else {
                int child;
                transition(g.tr_tab, A, g.C, node, row_action, action, u.chance, child, reward);
                node = child;
            }
            game_return += reward;
            write_record<A>(g.out, slot, node_now, turn, n_legal, policy, action, value, reward, logit);
        }
        if (g.out.returns != nullptr) g.out.returns[b] = game_return;
    }
    publish_stats(g.stats, last_valid, n_valid0, n_valid1, threadIdx.x & 31);
}

template <int A>
static int launch_fp32(const RolloutArgs& g, cudaStream_t st) {
    constexpr int KIN = 2 * A * A;
    constexpr int AP = round_up(A, 4);
    const int W = g.w.width;
    const size_t smem = sizeof(float) * ((size_t)2 * KIN * W + 3 * W + (size_t)W * AP + 1 + A + 3);
    if (smem > 227 * 1024) {
        set_error("rnad_rollout(fp32): net of width %d with %d actions needs %zu B of shared memory (max 232448)", W, A,
                  smem);
        return RNAD_EUNSUPPORTED;
    }
    int rc = check_cuda(cudaFuncSetAttribute(rollout_fp32_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem), "cudaFuncSetAttribute(rollout_fp32)");
    if (rc) return rc;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_fp32_kernel<A>, 128, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (g.B + 127) / 128;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    rollout_fp32_kernel<A><<<(int)blocks, 128, smem, st>>>(g);
    RNAD_CHECK_LAUNCH("rollout_fp32_kernel");
    return RNAD_OK;
}

int rollout_fp32(const RolloutArgs& g, cudaStream_t st) {
    if (g.w.width % 4 != 0) {
        set_error("rnad_rollout(fp32): width %d must be a multiple of 4", g.w.width);
        return RNAD_EUNSUPPORTED;
    }
    switch (g.A) {
        case 1: return launch_fp32<1>(g, st);
        case 2: return launch_fp32<2>(g, st);
        case 3: return launch_fp32<3>(g, st);
        case 4: return launch_fp32<4>(g, st);
        case 5: return launch_fp32<5>(g, st);
        case 6: return launch_fp32<6>(g, st);
        case 7: return launch_fp32<7>(g, st);
        case 8: return launch_fp32<8>(g, st);
    }
    set_error("rnad_rollout: max_actions %d unsupported", g.A);
    return RNAD_EINVAL;
}

}  // namespace rnad
