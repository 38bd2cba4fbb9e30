// K2, tensor-core build (RNAD_PREC_TF32): Episodes.generate fused with
// MLP.forward as a persistent kernel.
//
// A CTA owns a tile of 128 games for all T half-moves; game g of the tile is TMEM
// lane g and is served by two threads (g and g + 128) that split the hidden units.
// Per half-move:
//
//   threads 0..127   gather the node record (once per full move), build the
//                    observation row, store it (a) as tf32 into the A-operand tile
//                    [128 x KP] in shared memory (canonical K-major core-matrix
//                    layout; the bias rides along as a constant-1 input column when
//                    K has padding to spare) and (b) as fp32 into a staging tile
//   all 256 threads  copy the staging tile to the (T,B,2,A,A) trajectory as one
//                    contiguous, coalesced block while the tensor core works
//   one thread       tcgen05.mma kind::tf32, M=128, N=256, K=8 per instruction,
//                    against the SMEM-resident first-layer weights: value trunk,
//                    then policy trunk, fp32 accumulators in 256 TMEM columns
//   all 256 threads  tcgen05.ld epilogue (double-buffered 32-column chunks): relu and
//                    the second layers on the CUDA cores, 128 hidden units each
//   threads 0..127   masked softmax, Philox inverse-CDF action draw, trajectory
//                    record; on column half-moves the chance draw + child gather.
//
// The weights arrive in shared memory as ONE TMA bulk copy (cp.async.bulk) of an
// image that a small pre-kernel lays out in operand order (tf32-rounded) once per
// rollout.  No host synchronisation inside the rollout (the reference syncs twice
// per half-move, episode.py:96,124).  Two CTAs (16 warps) are resident per SM: 256
// TMEM columns and <= 107 KB shared memory each.
// Reference: environment/episode.py:175-230, nn/net.py:37-51.
#include "game.cuh"
#include "rollout.cuh"
#include "tc_common.cuh"

namespace rnad {

namespace tc {

constexpr int kThreads = 256;    // two threads per game
constexpr int kAccCols = 256;    // TMEM columns per CTA (one trunk at a time)
constexpr int kColsPerThread = kAccCols / 2;

// Shared-memory plan.  [0, kImageBytes) is the weight image, identical in the
// global workspace (written by pack_weights_kernel) and in shared memory.
template <int A>
struct Plan {
    static constexpr int KIN = 2 * A * A;
    static constexpr bool kBiasInK = (KIN % 8) != 0;                 // a padded K column is free to carry the bias
    static constexpr int KP = round_up(KIN + (kBiasInK ? 1 : 0), 8);
    static constexpr int kB = 0;                                     // first layers, both trunks: 512 x KP tf32
    static constexpr int kW2v = kB + 2 * kHidden * KP * 4;           // value_fc1.weight, 256 f32
    static constexpr int kW2p = kW2v + kHidden * 4;                  // policy_fc1.weight as [j][4] f32
    static constexpr int kB1 = kW2p + kHidden * 16;                  // first-layer biases (used when !kBiasInK), 512 f32
    static constexpr int kB2 = kB1 + 2 * kHidden * 4;                // value bias, policy biases (8 f32)
    static constexpr int kImageBytes = kB2 + 32;
    static constexpr int kA = kImageBytes;                           // A operand tile: 128 x KP tf32
    static constexpr int kObs = kA + kTileM * KP * 4;                // fp32 observation staging: 128 x KIN
    static constexpr int kPart = kObs + kTileM * KIN * 4;            // partial sums of the upper-half threads: 128 x 8 f32
    static constexpr int kBar = kPart + kTileM * 32;                 // two mbarriers
    static constexpr int kTmem = kBar + 16;                          // TMEM base address
    static constexpr int kBytes = kTmem + 16;
    static_assert(kImageBytes % 16 == 0 && kA % 16 == 0 && kObs % 16 == 0 && kPart % 16 == 0 && kBar % 8 == 0, "alignment");
};

template <int A>
__global__ void pack_weights_kernel(rnad_mlp_weights w, uint8_t* __restrict__ image) {
    using P = Plan<A>;
    constexpr int KIN = P::KIN, KP = P::KP;
    const int stride = gridDim.x * blockDim.x;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 2 * kHidden * KP; e += stride) {
        const int n = e / KP, k = e % KP;   // hidden unit n of [value trunk | policy trunk], input k
        const int j = n & (kHidden - 1);
        float v = 0.f;
        if (k < KIN) v = (n < kHidden ? w.value_fc0_w : w.policy_fc0_w)[j * KIN + k];
        else if (P::kBiasInK && k == KIN) v = (n < kHidden ? w.value_fc0_b : w.policy_fc0_b)[j];
        *reinterpret_cast<float*>(image + P::kB + operand_offset<KP>(n, k)) = to_tf32(v);
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < kHidden; j += stride) {
        reinterpret_cast<float*>(image + P::kW2v)[j] = w.value_fc1_w[j];
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        p.x = w.policy_fc1_w[j];
        if (A > 1) p.y = w.policy_fc1_w[1 * kHidden + j];
        if (A > 2) p.z = w.policy_fc1_w[2 * kHidden + j];
        if (A > 3) p.w = w.policy_fc1_w[3 * kHidden + j];
        reinterpret_cast<float4*>(image + P::kW2p)[j] = p;
        reinterpret_cast<float*>(image + P::kB1)[j] = w.value_fc0_b[j];
        reinterpret_cast<float*>(image + P::kB1)[kHidden + j] = w.policy_fc0_b[j];
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) {
        float v = 0.f;
        if (threadIdx.x == 0) v = w.value_fc1_b[0];
        else if ((int)threadIdx.x <= A) v = w.policy_fc1_b[threadIdx.x - 1];
        reinterpret_cast<float*>(image + P::kB2)[threadIdx.x] = v;
    }
}

template <int A>
__global__ void __launch_bounds__(kThreads, 2) rollout_tc_kernel(RolloutArgs g, const uint8_t* __restrict__ image) {
    using P = Plan<A>;
    constexpr int KIN = P::KIN;
    constexpr int KP = P::KP;
    static_assert(A <= 4, "policy_fc1 rows are staged as one float4 per hidden unit");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane_g = tid & (kTileM - 1);        // game of the tile == TMEM lane
    const int half = tid >> 7;                    // which 128 hidden units of each trunk this thread reduces
    const uint32_t mbar_mma = smem_u32(smem + P::kBar);
    const uint32_t mbar_img = smem_u32(smem + P::kBar + 8);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    // ---- one-time set-up: TMEM, mbarriers, weight image by TMA bulk copy
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kAccCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_mma) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_img) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_img), "r"(P::kImageBytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(smem)),
            "l"(image), "r"(P::kImageBytes), "r"(mbar_img)
            : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(mbar_img, 0);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_mine = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * kColsPerThread);
    const float* w2v = reinterpret_cast<const float*>(smem + P::kW2v);
    const float4* w2p = reinterpret_cast<const float4*>(smem + P::kW2p);
    const float* b1 = reinterpret_cast<const float*>(smem + P::kB1);
    const float* b2 = reinterpret_cast<const float*>(smem + P::kB2);
    float* s_obs = reinterpret_cast<float*>(smem + P::kObs);
    float* s_part = reinterpret_cast<float*>(smem + P::kPart);
    const uint32_t a_base = smem_u32(smem + P::kA);
    const uint32_t b_base = smem_u32(smem + P::kB);

    uint32_t phase = 0;
    int last_valid = -1, n_valid0 = 0, n_valid1 = 0;
    const uint64_t seed = g.seed_dev != nullptr ? __ldg(g.seed_dev) : g.seed;
    const int64_t num_tiles = (g.B + kTileM - 1) / kTileM;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t tile_base = tile * kTileM;
        const int64_t b = tile_base + lane_g;
        const bool active = b < g.B;
        const int tile_games = (int)min((int64_t)kTileM, g.B - tile_base);
        int node = active ? 1 : 0;
        int row_action = 0;
        float game_return = 0.f;
        Node<A> n;
#pragma unroll
        for (int i = 0; i < A * A; ++i) n.ev[i] = 0.f;
        n.rows = n.cols = 1;

        for (int t = 0; t < g.T; ++t) {
            const int turn = t & 1;
            const int64_t slot = (int64_t)t * g.B + b;
            int n_legal = 1;
            if (half == 0) {
                if (turn == 0 && active) load_node<A>(g.ev_tab, node, n);
                if (node != 0) {
                    last_valid = max(last_valid, t);
                    n_valid0 += turn == 0;
                    n_valid1 += turn;
                }
                n_legal = turn == 0 ? n.rows : n.cols;
                float x[KP];
                {
                    float obs[KIN];
                    build_obs<A>(n, turn, obs);
#pragma unroll
                    for (int k = 0; k < KIN; ++k) {
                        x[k] = obs[k];
                        s_obs[lane_g * KIN + k] = obs[k];
                    }
                }
#pragma unroll
                for (int k = KIN; k < KP; ++k) x[k] = (P::kBiasInK && k == KIN) ? 1.f : 0.f;
#pragma unroll
                for (int q = 0; q < KP / 4; ++q) {
                    const float4 v = make_float4(to_tf32_fast(x[4 * q]), to_tf32_fast(x[4 * q + 1]), to_tf32_fast(x[4 * q + 2]),
                                                 to_tf32_fast(x[4 * q + 3]));
                    *reinterpret_cast<float4*>(smem + P::kA + operand_offset<KP>(lane_g, 4 * q)) = v;
                }
                fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
            }
            tc_fence_before();
            __syncthreads();          // (1) operand tile and observation staging complete; TMEM free again

            float vacc[4] = {0.f, 0.f, 0.f, 0.f};
            float lacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};

#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int s = 0; s < KP / 8; ++s) {
                        const uint64_t ad = make_desc<KP>(a_base + s * 256);
                        const uint64_t bd = make_desc<KP>(b_base + pass * (kHidden / 8) * ((KP / 4) * 128) + s * 256);
                        mma_tf32(tmem_base, ad, bd, s > 0);
                    }
                    mma_commit(mbar_mma);
                }
                if (pass == 0) {
                    // the tile's observations are one contiguous block of the (T,B,2,A,A) tensor
                    float* dst = g.out.observations + ((int64_t)t * g.B + tile_base) * KIN;
                    const int n_float = tile_games * KIN;
                    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                        for (int i = tid * 4; i < n_float; i += kThreads * 4) {
                            if (i + 4 <= n_float) {
                                __stcs(reinterpret_cast<float4*>(dst + i), *reinterpret_cast<const float4*>(s_obs + i));
                            } else {
                                for (int k = i; k < n_float; ++k) __stcs(dst + k, s_obs[k]);
                            }
                        }
                    } else {
                        for (int i = tid; i < n_float; i += kThreads) __stcs(dst + i, s_obs[i]);
                    }
                }
                mbar_wait(mbar_mma, phase);
                phase ^= 1u;
                tc_fence_after();
                const int col0 = pass * kHidden + half * kColsPerThread;   // index into b1 (both trunks)
                const int wcol0 = half * kColsPerThread;                   // index into w2v / w2p
                uint32_t ra[32], rb[32];
                tmem_ld32(tmem_mine, ra);
#pragma unroll
                for (int c = 0; c < kColsPerThread / 32; ++c) {
                    tmem_ld_wait();
                    if (c + 1 < kColsPerThread / 32) {
                        if (c & 1) tmem_ld32(tmem_mine + (c + 1) * 32, ra);
                        else tmem_ld32(tmem_mine + (c + 1) * 32, rb);
                    }
                    if (pass == 0) {
                        if (c & 1) consume_chunk<A, true, !P::kBiasInK>(rb, 0, b1 + col0 + c * 32, w2v + wcol0 + c * 32, w2p, vacc, lacc);
                        else consume_chunk<A, true, !P::kBiasInK>(ra, 0, b1 + col0 + c * 32, w2v + wcol0 + c * 32, w2p, vacc, lacc);
                    } else {
                        if (c & 1) consume_chunk<A, false, !P::kBiasInK>(rb, 0, b1 + col0 + c * 32, w2v, w2p + wcol0 + c * 32, vacc, lacc);
                        else consume_chunk<A, false, !P::kBiasInK>(ra, 0, b1 + col0 + c * 32, w2v, w2p + wcol0 + c * 32, vacc, lacc);
                    }
                }
                tc_fence_before();   // TMEM reads done before the next MMA overwrites the accumulator
                if (pass == 0) __syncthreads();   // (2)
            }

            const float v_part = (vacc[0] + vacc[1]) + (vacc[2] + vacc[3]);
            float l_part[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) l_part[a] = lacc[0][a] + lacc[1][a];
            if (half == 1) {
                float4* dst = reinterpret_cast<float4*>(s_part + lane_g * 8);
                dst[0] = make_float4(v_part, 0.f, 0.f, 0.f);
                dst[1] = make_float4(l_part[0], l_part[1], l_part[2], l_part[3]);
            }
            __syncthreads();          // (3) upper-half partial sums visible

            if (half == 0) {
                const float4* src = reinterpret_cast<const float4*>(s_part + lane_g * 8);
                const float4 pv = src[0], pl = src[1];
                const float value = (v_part + pv.x) + b2[0];
                const float other[4] = {pl.x, pl.y, pl.z, pl.w};
                float logit[A];
#pragma unroll
                for (int a = 0; a < A; ++a) logit[a] = (l_part[a] + other[a]) + b2[1 + a];

                float policy[A];
                masked_softmax<A>(logit, n_legal, policy);
                Uniforms2 u;
                if (g.uniforms != nullptr) {
                    u.action = active ? g.uniforms[slot * 2 + 0] : 0.f;
                    u.chance = active ? g.uniforms[slot * 2 + 1] : 0.f;
                } else {
                    u = philox_uniforms(seed, (uint32_t)t, (uint64_t)(g.game_offset + b));
                }
                const int action = sample_icdf(policy, A, u.action);
                float reward = 0.f;
                const int node_now = node;
                if (turn == 0) {
                    row_action = action;
                } else if (active) {
                    int child;
                    transition(g.tr_tab, A, g.C, node, row_action, action, u.chance, child, reward);
                    node = child;
                }
                game_return += reward;
                if (active) write_record<A>(g.out, slot, node_now, turn, n_legal, policy, action, value, reward, logit);
            }
        }
        if (half == 0 && active && g.out.returns != nullptr) g.out.returns[b] = game_return;
    }

    publish_stats(g.stats, last_valid, n_valid0, n_valid1, tid & 31);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAccCols) : "memory");
    }
}

template <int A>
static int launch(const RolloutArgs& g, uint8_t* workspace, cudaStream_t st) {
    using P = Plan<A>;
    pack_weights_kernel<A><<<24, 256, 0, st>>>(g.w, workspace);
    RNAD_CHECK_LAUNCH("pack_weights_kernel");
    // Two CTAs per SM share the 512 TMEM columns; pad the shared-memory request so that a
    // third CTA can never become resident and spin inside tcgen05.alloc.
    size_t smem = P::kBytes;
    const size_t floor_two_per_sm = 227 * 1024 / 3 + 1024;
    if (smem < floor_two_per_sm) smem = floor_two_per_sm;
    if (smem > 227 * 1024) {
        set_error("rnad_rollout(tf32): %zu B of shared memory needed", smem);
        return RNAD_EUNSUPPORTED;
    }
    const int per_sm = 2 * (smem + 1024) <= 228 * 1024 ? 2 : 1;
    int rc = check_cuda(cudaFuncSetAttribute(rollout_tc_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(rollout_tc, smem)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(rollout_tc_kernel<A>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared),
                    "cudaFuncSetAttribute(rollout_tc, carveout)");
    if (rc) return rc;
    int64_t blocks = (g.B + kTileM - 1) / kTileM;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    rollout_tc_kernel<A><<<(int)blocks, kThreads, smem, st>>>(g, workspace);
    RNAD_CHECK_LAUNCH("rollout_tc_kernel");
    return RNAD_OK;
}

template <int A>
constexpr int64_t image_bytes() { return Plan<A>::kImageBytes; }

}  // namespace tc

bool rollout_tc_supported(int A, int width) { return width == tc::kHidden && A >= 2 && A <= 4; }

int64_t rollout_tc_workspace_bytes(int A) {
    switch (A) {
        case 2: return tc::image_bytes<2>();
        case 3: return tc::image_bytes<3>();
        case 4: return tc::image_bytes<4>();
    }
    return 0;
}

int rollout_tc(const RolloutArgs& g, void* workspace, cudaStream_t st) {
    if (!rollout_tc_supported(g.A, g.w.width)) {
        set_error("rnad_rollout(tf32): needs width == 256 and 2 <= max_actions <= 4 (got width %d, max_actions %d)",
                  g.w.width, g.A);
        return RNAD_EUNSUPPORTED;
    }
    if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0) {
        set_error("rnad_rollout(tf32): needs a 16-byte aligned workspace of rnad_rollout_workspace_bytes()");
        return RNAD_EINVAL;
    }
    switch (g.A) {
        case 2: return tc::launch<2>(g, (uint8_t*)workspace, st);
        case 3: return tc::launch<3>(g, (uint8_t*)workspace, st);
        case 4: return tc::launch<4>(g, (uint8_t*)workspace, st);
    }
    return RNAD_EINVAL;
}

}  // namespace rnad
