// Learner-side net passes of RNaD.__learn, fused (SURVEY.md section 8 f#3):
//
//   rnad_learner_forward   the four forward_batch calls of rnad.py:373-380 in ONE pass
//                          over the trajectory: learner (logit, pi, log_pi, v), target
//                          net (v only), regularisation nets (log_pi only) = five trunk
//                          evaluations per step instead of eight.  The engine in use is
//                          learner_fwd_tc2.cu (both layers on tcgen05, fp16 operands); the
//                          kernel in THIS file (first layers tcgen05 kind::tf32, second
//                          layers on the CUDA cores) is the first build, kept for A/B runs
//                          (RNAD_LEARNER_FWD_V1).
//   rnad_learner_backward  parameter gradients of the learner net given d loss/d logit
//   (_split)               and d loss/d v (from rnad_learner_targets).  Dispatch, newest
//                          first:  split mode, max_actions <= 3 -> learner_bwd_f16.cu (fp16
//                          operands, mask formulation, decoupled tensor-memory regions);
//                          flat mode, max_actions <= 3 -> learner_bwd_tc3.cu (tf32, mask
//                          formulation);  max_actions = 4 -> learner_bwd_tc2_kernel here:
//                          the two trunks recomputed TRANSPOSED on the tensor core,
//                              dW2 = g^T relu(h),  dh = (g W2) * [h > 0],  dW1 = dh^T x
//                          with MMAs whose A operands (relu^T, dh^T) sit in tensor memory
//                          and whose accumulators stay there across all tiles of a CTA.
//                          A second kernel adds the per-CTA partials in a fixed order
//                          (deterministic), per player in split mode.  learner_bwd_tc_kernel
//                          (two CTAs per SM) and learner_bwd_kernel (reduction on the CUDA
//                          cores) are earlier builds kept for A/B runs (RNAD_LEARNER_BWD_V1,
//                          RNAD_LEARNER_BWD_CUDA_CORES).
//
// Reference: nn/net.py:64-85 (forward_batch), learn/rnad.py:373-380, 424-425.
// In the reference these are 8 x T small GEMMs + ~40 elementwise launches forward and
// autograd's mirror image backward, with the (T*B x 256) activations written to and
// re-read from HBM about ten times per update.
#include <cstdlib>

#include "learner_bwd.cuh"
#include "tc_common.cuh"
#include "tc_pipe.cuh"

namespace rnad {

// pipelined forward with both layers on the tensor core (learner_fwd_tc2.cu)
bool learner_forward_tc2_supported(int A, int width);
int64_t learner_forward_tc2_image_bytes(int A);
int learner_forward_tc2(const float* obs, int64_t N, int A, const rnad_mlp_weights* net, const rnad_mlp_weights* target,
                        const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_, const rnad_learner_fwd_out* out,
                        void* workspace, cudaStream_t st, int mode, bool others_only);

namespace tc {

constexpr int kLearnThreads = 256;   // two threads per trajectory row
constexpr int kFwdTrunks = 5;        // learner value, learner policy, target value, reg policy, reg_ policy

// ------------------------------------------------------------------ forward

template <int A>
struct FwdPlan : Shape<A> {
    using S = Shape<A>;
    static constexpr int kB = 0;                                        // 5 first-layer operands
    static constexpr int kW2v = kB + kFwdTrunks * S::kTrunkBytes;       // value_fc1.weight of learner, target
    static constexpr int kW2p = kW2v + 2 * kHidden * 4;                 // policy_fc1.weight [j][4] of learner, reg, reg_
    static constexpr int kB1 = kW2p + 3 * kHidden * 16;                 // first-layer biases, 5 x 256
    static constexpr int kB2 = kB1 + kFwdTrunks * kHidden * 4;          // second-layer biases: 5 x 4 f32
    static constexpr int kImageBytes = kB2 + kFwdTrunks * 16;
    static constexpr int kA = kImageBytes;                              // A operand tile
    static constexpr int kPart = kA + kTileM * S::KP * 4;               // upper-half partial sums: 128 x 5 x float4
    static constexpr int kBar = kPart + kTileM * kFwdTrunks * 16;       // mbarriers: image, stage 0, stage 1
    static constexpr int kTmem = kBar + 32;
    static constexpr int kBytes = kTmem + 16;
    static_assert(kImageBytes % 16 == 0 && kA % 16 == 0 && kPart % 16 == 0 && kBar % 8 == 0, "alignment");
};

struct FwdNets {
    rnad_mlp_weights net, target, reg, reg_;
};

struct FwdOut {
    float* logit;
    float* pi;
    float* log_pi;
    float* v;
    float* v_target;
    float* log_pi_reg;
    float* log_pi_reg_;
};

template <int A>
__global__ void pack_fwd_image_kernel(FwdNets w, uint8_t* __restrict__ image) {
    using P = FwdPlan<A>;
    const int thread = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    const float* w1[kFwdTrunks] = {w.net.value_fc0_w, w.net.policy_fc0_w, w.target.value_fc0_w, w.reg.policy_fc0_w,
                                   w.reg_.policy_fc0_w};
    const float* b1[kFwdTrunks] = {w.net.value_fc0_b, w.net.policy_fc0_b, w.target.value_fc0_b, w.reg.policy_fc0_b,
                                   w.reg_.policy_fc0_b};
#pragma unroll
    for (int tr = 0; tr < kFwdTrunks; ++tr) {
        pack_trunk_operand<P::KIN, P::KP, P::kBiasInK>(w1[tr], b1[tr], image + P::kB + tr * P::kTrunkBytes, thread,
                                                       n_threads);
        for (int j = thread; j < kHidden; j += n_threads) reinterpret_cast<float*>(image + P::kB1)[tr * kHidden + j] = b1[tr][j];
    }
    const float* w2v[2] = {w.net.value_fc1_w, w.target.value_fc1_w};
    const float* w2p[3] = {w.net.policy_fc1_w, w.reg.policy_fc1_w, w.reg_.policy_fc1_w};
    for (int j = thread; j < kHidden; j += n_threads) {
#pragma unroll
        for (int i = 0; i < 2; ++i) reinterpret_cast<float*>(image + P::kW2v)[i * kHidden + j] = w2v[i][j];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            p.x = w2p[i][j];
            if (A > 1) p.y = w2p[i][1 * kHidden + j];
            if (A > 2) p.z = w2p[i][2 * kHidden + j];
            if (A > 3) p.w = w2p[i][3 * kHidden + j];
            reinterpret_cast<float4*>(image + P::kW2p)[i * kHidden + j] = p;
        }
    }
    if (thread < kFwdTrunks * 4) {
        const int tr = thread >> 2, a = thread & 3;
        const float* b2[kFwdTrunks] = {w.net.value_fc1_b, w.net.policy_fc1_b, w.target.value_fc1_b, w.reg.policy_fc1_b,
                                       w.reg_.policy_fc1_b};
        const int n = (tr == 0 || tr == 2) ? 1 : A;
        reinterpret_cast<float*>(image + P::kB2)[thread] = a < n ? b2[tr][a] : 0.f;
    }
}

template <int KP>
__device__ __forceinline__ void issue_trunk_mma(uint32_t a_base, uint32_t b_trunk, uint32_t d_tmem, uint32_t mbar) {
#pragma unroll
    for (int s = 0; s < KP / 8; ++s) mma_tf32(d_tmem, make_desc<KP>(a_base + s * 256), make_desc<KP>(b_trunk + s * 256), s > 0);
    mma_commit(mbar);
}

// net.py:76-80: masked softmax / log-softmax of one row
template <int A>
__device__ __forceinline__ void policy_heads(const float (&logit)[A], const bool (&mask)[A], float (&pi)[A],
                                             float (&log_pi)[A]) {
    float e[A];
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        e[a] = mask[a] ? expf(logit[a]) : 0.f;
        sum += e[a];
    }
    const float denom = fmaxf(sum, 1e-12f);
    const float log_sum = logf(sum);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        pi[a] = e[a] / denom;
        log_pi[a] = mask[a] ? logit[a] - log_sum : 0.f;
    }
}

template <int A>
__global__ void __launch_bounds__(kLearnThreads, 1) learner_fwd_kernel(const float* __restrict__ obs, int64_t N,
                                                                        const uint8_t* __restrict__ image, FwdOut out) {
    using P = FwdPlan<A>;
    constexpr int KIN = P::KIN, KP = P::KP;
    static_assert(A <= 4, "one float4 of second-layer weights per hidden unit");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & (kTileM - 1), half = tid >> 7;
    const uint32_t bar_img = smem_u32(smem + P::kBar), bar_stage[2] = {bar_img + 8, bar_img + 16};
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        mbar_init(bar_stage[0], 1);
        mbar_init(bar_stage[1], 1);
        mbar_fence_init();
        tma_bulk_load(smem, image, P::kImageBytes, bar_img);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(bar_img, 0);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_mine = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 128);
    const uint32_t a_base = smem_u32(smem + P::kA), b_base = smem_u32(smem + P::kB);
    const float* b1 = reinterpret_cast<const float*>(smem + P::kB1);
    const float* b2 = reinterpret_cast<const float*>(smem + P::kB2);
    const float* w2v = reinterpret_cast<const float*>(smem + P::kW2v);
    const float4* w2p = reinterpret_cast<const float4*>(smem + P::kW2p);
    float4* s_part = reinterpret_cast<float4*>(smem + P::kPart);

    uint32_t phase[2] = {0u, 0u};
    const int64_t num_tiles = (N + kTileM - 1) / kTileM;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t row = tile * kTileM + lane;
        const bool active = row < N;
        bool mask[A];
#pragma unroll
        for (int a = 0; a < A; ++a) mask[a] = false;
        if (half == 0) {
            float x[KIN];
            load_row<KIN>(obs, row, active, x);
#pragma unroll
            for (int a = 0; a < A; ++a) mask[a] = x[A * A + a * A] != 0.f;   // obs[:, 1, :, 0]
            store_operand_row<KIN, KP, P::kBiasInK>(smem + P::kA, lane, x);
            fence_async_smem();
        }
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {           // converged warp, one elected lane issues (see learner_bwd_tc_kernel)
            tc_fence_after();
            if (tcp::elect_one()) {
                issue_trunk_mma<KP>(a_base, b_base + 0 * P::kTrunkBytes, tmem_base + 0, bar_stage[0]);
                issue_trunk_mma<KP>(a_base, b_base + 1 * P::kTrunkBytes, tmem_base + 256, bar_stage[1]);
            }
            __syncwarp();
        }

        float4 part[kFwdTrunks];
#pragma unroll
        for (int pass = 0; pass < kFwdTrunks; ++pass) {
            constexpr int kNone = 0;
            (void)kNone;
            const int stage = pass & 1;
            const bool value_pass = pass == 0 || pass == 2;
            mbar_wait(bar_stage[stage], phase[stage]);
            phase[stage] ^= 1u;
            tc_fence_after();
            float vacc[4] = {0.f, 0.f, 0.f, 0.f};
            float lacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            const float* b1p = b1 + pass * kHidden + half * 128;
            const float* w2vp = w2v + (pass == 2 ? kHidden : 0) + half * 128;
            const float4* w2pp = w2p + (pass == 1 ? 0 : pass == 3 ? kHidden : 2 * kHidden) + half * 128;
            const uint32_t taddr = tmem_mine + stage * 256;
            uint32_t ra[32], rb[32];
            tmem_ld32(taddr, ra);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_ld_wait();
                if (c + 1 < 4) {
                    if (c & 1) tmem_ld32(taddr + (c + 1) * 32, ra);
                    else tmem_ld32(taddr + (c + 1) * 32, rb);
                }
                if (value_pass) {
                    if (c & 1) consume_chunk<A, true, !P::kBiasInK>(rb, 0, b1p + c * 32, w2vp + c * 32, w2pp, vacc, lacc);
                    else consume_chunk<A, true, !P::kBiasInK>(ra, 0, b1p + c * 32, w2vp + c * 32, w2pp, vacc, lacc);
                } else {
                    if (c & 1) consume_chunk<A, false, !P::kBiasInK>(rb, 0, b1p + c * 32, w2vp, w2pp + c * 32, vacc, lacc);
                    else consume_chunk<A, false, !P::kBiasInK>(ra, 0, b1p + c * 32, w2vp, w2pp + c * 32, vacc, lacc);
                }
            }
            part[pass] = value_pass ? make_float4((vacc[0] + vacc[1]) + (vacc[2] + vacc[3]), 0.f, 0.f, 0.f)
                                    : make_float4(lacc[0][0] + lacc[1][0], lacc[0][1] + lacc[1][1],
                                                  lacc[0][2] + lacc[1][2], lacc[0][3] + lacc[1][3]);
            tc_fence_before();
            if (pass + 2 < kFwdTrunks) {
                __syncthreads();   // every thread has drained this accumulator stage
                if (warp == 0) {
                    tc_fence_after();
                    if (tcp::elect_one())
                        issue_trunk_mma<KP>(a_base, b_base + (pass + 2) * P::kTrunkBytes, tmem_base + stage * 256,
                                            bar_stage[stage]);
                    __syncwarp();
                }
            }
        }
        if (half == 1) {
#pragma unroll
            for (int pass = 0; pass < kFwdTrunks; ++pass) s_part[lane * kFwdTrunks + pass] = part[pass];
        }
        __syncthreads();
        if (half == 0 && active) {
            float4 tot[kFwdTrunks];
#pragma unroll
            for (int pass = 0; pass < kFwdTrunks; ++pass) {
                const float4 o = s_part[lane * kFwdTrunks + pass];
                const float4 bias = *reinterpret_cast<const float4*>(b2 + pass * 4);
                tot[pass] = make_float4((part[pass].x + o.x) + bias.x, (part[pass].y + o.y) + bias.y,
                                        (part[pass].z + o.z) + bias.z, (part[pass].w + o.w) + bias.w);
            }
            out.v[row] = tot[0].x;
            out.v_target[row] = tot[2].x;
            const float heads[3][4] = {{tot[1].x, tot[1].y, tot[1].z, tot[1].w},
                                       {tot[3].x, tot[3].y, tot[3].z, tot[3].w},
                                       {tot[4].x, tot[4].y, tot[4].z, tot[4].w}};
            float* log_dst[3] = {out.log_pi, out.log_pi_reg, out.log_pi_reg_};
#pragma unroll
            for (int hnet = 0; hnet < 3; ++hnet) {
                float logit[A], pi[A], log_pi[A];
#pragma unroll
                for (int a = 0; a < A; ++a) logit[a] = heads[hnet][a];
                policy_heads<A>(logit, mask, pi, log_pi);
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    log_dst[hnet][row * A + a] = log_pi[a];
                    if (hnet == 0) {
                        out.logit[row * A + a] = logit[a];
                        out.pi[row * A + a] = pi[a];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// ----------------------------------------------------------------- backward

template <int A>
struct BwdPlan : Shape<A> {
    using S = Shape<A>;
    static constexpr int kHStride = kHidden + 1;                        // fp32 words per row of the activation tile
    static constexpr int kB = 0;                                        // learner value, policy first layers
    static constexpr int kB1 = kB + 2 * S::kTrunkBytes;                 // their biases, 2 x 256
    static constexpr int kImageBytes = kB1 + 2 * kHidden * 4;
    static constexpr int kA = kImageBytes;                              // A operand tile (tf32)
    static constexpr int kG = kA + kTileM * S::KP * 4;                  // [128][8]: d_v, d_logit[0..A)
    static constexpr int kH = kG + kTileM * 32;                         // relu(h) of one trunk, [128][257] f32
    static constexpr int kXraw = kH + kTileM * kHStride * 4;            // exact fp32 inputs [128][KIN] (if they fit)
    static constexpr bool kExactX = kXraw + kTileM * S::KIN * 4 + 64 <= 227 * 1024;
    static constexpr int kBar = round_up(kXraw + (kExactX ? kTileM * S::KIN * 4 : 0), 16);
    static constexpr int kTmem = kBar + 32;
    static constexpr int kBytes = kTmem + 16;
    static_assert(kBytes <= 227 * 1024, "backward tile does not fit in shared memory");
};

template <int A>
__global__ void pack_bwd_image_kernel(rnad_mlp_weights w, uint8_t* __restrict__ image) {
    using P = BwdPlan<A>;
    const int thread = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    pack_trunk_operand<P::KIN, P::KP, P::kBiasInK>(w.value_fc0_w, w.value_fc0_b, image + P::kB, thread, n_threads);
    pack_trunk_operand<P::KIN, P::KP, P::kBiasInK>(w.policy_fc0_w, w.policy_fc0_b, image + P::kB + P::kTrunkBytes,
                                                   thread, n_threads);
    for (int j = thread; j < kHidden; j += n_threads) {
        reinterpret_cast<float*>(image + P::kB1)[j] = w.value_fc0_b[j];
        reinterpret_cast<float*>(image + P::kB1)[kHidden + j] = w.policy_fc0_b[j];
    }
}

template <int A>
__global__ void __launch_bounds__(kLearnThreads, 1) learner_bwd_kernel(const float* __restrict__ obs, int64_t N,
                                                                        const uint8_t* __restrict__ image,
                                                                        rnad_mlp_weights w,
                                                                        const float* __restrict__ d_logit,
                                                                        const float* __restrict__ d_v,
                                                                        float* __restrict__ partials) {
    using P = BwdPlan<A>;
    constexpr int KIN = P::KIN, KP = P::KP, HS = P::kHStride;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & (kTileM - 1), half = tid >> 7;
    const uint32_t bar_img = smem_u32(smem + P::kBar), bar_stage[2] = {bar_img + 8, bar_img + 16};
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        mbar_init(bar_stage[0], 1);
        mbar_init(bar_stage[1], 1);
        mbar_fence_init();
        tma_bulk_load(smem, image, P::kImageBytes, bar_img);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(bar_img, 0);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_mine = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 128);
    const uint32_t a_base = smem_u32(smem + P::kA), b_base = smem_u32(smem + P::kB);
    const float* b1 = reinterpret_cast<const float*>(smem + P::kB1);
    float* s_g = reinterpret_cast<float*>(smem + P::kG);
    float* s_h = reinterpret_cast<float*>(smem + P::kH);
    float* s_x = reinterpret_cast<float*>(smem + P::kXraw);

    // thread j owns hidden unit j of both trunks in the reduction phase
    const int j = tid;
    const float w2v_j = w.value_fc1_w[j];
    float w2p_j[A];
#pragma unroll
    for (int a = 0; a < A; ++a) w2p_j[a] = w.policy_fc1_w[a * kHidden + j];
    float gw1[2][KIN], gb1[2] = {0.f, 0.f}, gw2v = 0.f, gw2p[A], gb2 = 0.f;
#pragma unroll
    for (int k = 0; k < KIN; ++k) gw1[0][k] = gw1[1][k] = 0.f;
#pragma unroll
    for (int a = 0; a < A; ++a) gw2p[a] = 0.f;

    uint32_t phase[2] = {0u, 0u};
    const int64_t num_tiles = (N + kTileM - 1) / kTileM;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t row = tile * kTileM + lane;
        const bool active = row < N;
        if (half == 0) {
            float x[KIN];
            load_row<KIN>(obs, row, active, x);
            store_operand_row<KIN, KP, P::kBiasInK>(smem + P::kA, lane, x);
            if (P::kExactX) {
#pragma unroll
                for (int k = 0; k < KIN; ++k) s_x[lane * KIN + k] = x[k];
            }
            fence_async_smem();
        } else {
            float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) {
                g0.x = d_v[row];
                g0.y = d_logit[row * A + 0];
                if (A > 1) g0.z = d_logit[row * A + 1];
                if (A > 2) g0.w = d_logit[row * A + 2];
                if (A > 3) g1.x = d_logit[row * A + 3];
            }
            reinterpret_cast<float4*>(s_g + lane * 8)[0] = g0;
            reinterpret_cast<float4*>(s_g + lane * 8)[1] = g1;
        }
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            issue_trunk_mma<KP>(a_base, b_base, tmem_base + 0, bar_stage[0]);
            issue_trunk_mma<KP>(a_base, b_base + P::kTrunkBytes, tmem_base + 256, bar_stage[1]);
        }
#pragma unroll
        for (int tr = 0; tr < 2; ++tr) {
            // ---- activations of this trunk: TMEM -> relu -> shared [row][hidden]
            mbar_wait(bar_stage[tr], phase[tr]);
            phase[tr] ^= 1u;
            tc_fence_after();
            const uint32_t taddr = tmem_mine + tr * 256;
            uint32_t ra[32], rb[32];
            tmem_ld32(taddr, ra);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_ld_wait();
                if (c + 1 < 4) {
                    if (c & 1) tmem_ld32(taddr + (c + 1) * 32, ra);
                    else tmem_ld32(taddr + (c + 1) * 32, rb);
                }
                const int col = half * 128 + c * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float h = __uint_as_float((c & 1) ? rb[i] : ra[i]);
                    if (!P::kBiasInK) h += b1[tr * kHidden + col + i];
                    s_h[lane * HS + col + i] = fmaxf(h, 0.f);
                }
            }
            tc_fence_before();
            __syncthreads();
            // ---- reduce over the 128 rows of the tile, hidden unit j per thread
            {
                float acc_w1[KIN], acc_b1 = 0.f, acc_w2[A];
#pragma unroll
                for (int k = 0; k < KIN; ++k) acc_w1[k] = 0.f;
#pragma unroll
                for (int a = 0; a < A; ++a) acc_w2[a] = 0.f;
                float acc_b2 = 0.f;
#pragma unroll 2
                for (int n = 0; n < kTileM; ++n) {
                    const float r = s_h[n * HS + j];
                    const float4 g0 = reinterpret_cast<const float4*>(s_g + n * 8)[0];
                    float s;
                    if (tr == 0) {
                        s = g0.x * w2v_j;
                        acc_w2[0] = fmaf(g0.x, r, acc_w2[0]);
                    } else {
                        const float gl[4] = {g0.y, g0.z, g0.w, A > 3 ? s_g[n * 8 + 4] : 0.f};
                        s = 0.f;
#pragma unroll
                        for (int a = 0; a < A; ++a) {
                            s = fmaf(gl[a], w2p_j[a], s);
                            acc_w2[a] = fmaf(gl[a], r, acc_w2[a]);
                        }
                    }
                    if (tr == 0 && j <= A) acc_b2 += s_g[n * 8 + j];   // threads 0..A also sum the output-bias gradients
                    const float dh = r > 0.f ? s : 0.f;
                    acc_b1 += dh;
                    if (P::kExactX) {
                        const float2* xr = reinterpret_cast<const float2*>(s_x + n * KIN);
#pragma unroll
                        for (int k2 = 0; k2 < KIN / 2; ++k2) {
                            const float2 xv = xr[k2];
                            acc_w1[2 * k2] = fmaf(dh, xv.x, acc_w1[2 * k2]);
                            acc_w1[2 * k2 + 1] = fmaf(dh, xv.y, acc_w1[2 * k2 + 1]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < KIN / 4; ++q) {   // tf32-rounded inputs straight from the operand tile
                            const float4 xv = *reinterpret_cast<const float4*>(smem + P::kA + operand_offset<KP>(n, 4 * q));
                            acc_w1[4 * q + 0] = fmaf(dh, xv.x, acc_w1[4 * q + 0]);
                            acc_w1[4 * q + 1] = fmaf(dh, xv.y, acc_w1[4 * q + 1]);
                            acc_w1[4 * q + 2] = fmaf(dh, xv.z, acc_w1[4 * q + 2]);
                            acc_w1[4 * q + 3] = fmaf(dh, xv.w, acc_w1[4 * q + 3]);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < KIN; ++k) gw1[tr][k] += acc_w1[k];
                gb1[tr] += acc_b1;
                if (tr == 0) {
                    gw2v += acc_w2[0];
                    gb2 += acc_b2;
                } else {
#pragma unroll
                    for (int a = 0; a < A; ++a) gw2p[a] += acc_w2[a];
                }
            }
            __syncthreads();   // s_h (and after the second trunk: the operand tile, s_g, s_x) may be overwritten
        }
    }

    // ---- this CTA's partial gradient, flat in state_dict order
    float* dst = partials + (int64_t)blockIdx.x * P::kParams;
#pragma unroll
    for (int k = 0; k < KIN; ++k) {
        dst[P::kOffV0w + j * KIN + k] = gw1[0][k];
        dst[P::kOffP0w + j * KIN + k] = gw1[1][k];
    }
    dst[P::kOffV0b + j] = gb1[0];
    dst[P::kOffP0b + j] = gb1[1];
    dst[P::kOffV1w + j] = gw2v;
#pragma unroll
    for (int a = 0; a < A; ++a) dst[P::kOffP1w + a * kHidden + j] = gw2p[a];
    if (j == 0) dst[P::kOffV1b] = gb2;
    else if (j <= A) dst[P::kOffP1b + j - 1] = gb2;

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// -------------------------------------------------------- backward on the tensor core
//
// dW1 = dh^T x, db1 = dh^T 1 and dW2 = relu(h)^T g are contractions over the ROWS of a tile, while tensor memory
// holds an accumulator with rows as lanes.  So the trunks are recomputed TRANSPOSED: H^T[j][n] = W1[j,:] . X^T[:,n]
// (A operand = the first-layer weights, B operand = the observation tile, both K-major in shared memory) puts hidden
// unit j on lane j and row n on column n.  The thread of lane j turns its columns into relu(h)^T and
// dh^T = ((g W2) * [h > 0])^T in place, and those are A operands *in tensor memory* (K = rows) for
//     D_w1[j][c] += sum_n dh^T[j][n]   BX[c][n]      BX rows: x^T (2A^2 rows), ones     -> dW1, db1
//     D_w2[j][c] += sum_n relu^T[j][n] BG[c][n]      BG rows: d_v, d_logit[0..A)        -> dW2
// whose accumulators stay in tensor memory over all tiles of the CTA.  No activation ever leaves the SM and the
// CUDA cores do ~9 instructions per (row, hidden unit) instead of ~40.  Deterministic (fixed order everywhere).

constexpr int kBwdTcThreads = 256;   // two threads per hidden unit (TMEM lane), 32 rows (columns) of a 64-row stage each

// Two CTAs share an SM, one per trunk (256 TMEM columns each): while one waits for its MMAs the other runs its
// elementwise stage.  A stage is (128-unit half of the trunk) x (64-row half of the tile).
template <int A>
struct BwdTcPlan : Shape<A> {
    using S = Shape<A>;
    static constexpr int kNX = round_up(S::KIN + 1, 16);               // rows of BX = N of the dW1 / db1 MMA
    static constexpr int kNG = 16;                                     // rows of BG = N of the dW2 MMA
    static constexpr int kLbo = 144;                                   // K-chunk stride of BX / BG: 128 + 16 bytes of padding make the
                                                                       // transposing stores (32 rows n of one operand row c) conflict-free
    static constexpr int kSboT = (kTileM / 4) * kLbo;                  // 8-row groups of BX / BG
    // global image (pack_bwd_tc_image_kernel): both trunks' first layers, their biases, the transposed second layers
    static constexpr int kB = 0;                                       // [2][256 x KP] tf32, K-major
    static constexpr int kB1 = kB + 2 * S::kTrunkBytes;                // [2][256] f32
    static constexpr int kW2T = kB1 + 2 * kHidden * 4;                 // [2][256 x 8] tf32 K-major: row j = (w2v[j],0..) or (0,w2p[0][j],..)
    static constexpr int kImageBytes = kW2T + 2 * kHidden * 8 * 4;
    // shared memory of a CTA: its trunk's slices of the image, then the tile's operands
    static constexpr int kSW1 = 0;
    static constexpr int kSB1 = kSW1 + S::kTrunkBytes;
    static constexpr int kSW2T = kSB1 + kHidden * 4;
    static constexpr int kX = kSW2T + kHidden * 8 * 4;                 // observation tile [128 x KP] tf32, K-major
    static constexpr int kBX = kX + kTileM * S::KP * 4;
    static constexpr int kBG = kBX + (kNX / 8) * kSboT;
    static constexpr int kG = kBG + (kNG / 8) * kSboT;                 // g tile [128 x 8] tf32, K-major: d_v, d_logit[0..A)
    static constexpr int kRed = kG + kTileM * 32;                      // [4 warps][8] partial sums of g (output-bias gradients)
    static constexpr int kBar = kRed + 4 * 32;
    static constexpr int kTmem = kBar + 32;
    static constexpr int kBytes = kTmem + 16;
    // tensor memory (256 columns): [0,64) H^T -> relu^T, [64,128) S^T -> dh^T, then 2 x kNX of D_w1 and 2 x 16 of D_w2
    static constexpr int kColW1 = 128, kColW2 = 128 + 2 * kNX;
    static_assert(kColW2 + 2 * kNG <= 256, "accumulators do not fit the CTA's share of tensor memory");
    static_assert(2 * (kBytes + 1024) <= 228 * 1024, "two CTAs per SM do not fit shared memory");
    static_assert(kX % 16 == 0 && kBX % 16 == 0 && kBG % 16 == 0 && kG % 16 == 0 && kBar % 8 == 0, "alignment");
};

template <int A>
__global__ void pack_bwd_tc_image_kernel(rnad_mlp_weights w, uint8_t* __restrict__ image) {
    using P = BwdTcPlan<A>;
    const int thread = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    pack_trunk_operand<P::KIN, P::KP, P::kBiasInK>(w.value_fc0_w, w.value_fc0_b, image + P::kB, thread, n_threads);
    pack_trunk_operand<P::KIN, P::KP, P::kBiasInK>(w.policy_fc0_w, w.policy_fc0_b, image + P::kB + P::kTrunkBytes,
                                                   thread, n_threads);
    for (int j = thread; j < kHidden; j += n_threads) {
        reinterpret_cast<float*>(image + P::kB1)[j] = w.value_fc0_b[j];
        reinterpret_cast<float*>(image + P::kB1)[kHidden + j] = w.policy_fc0_b[j];
    }
    for (int e = thread; e < 2 * kHidden * 8; e += n_threads) {
        const int row = e / 8, k = e % 8, trunk = row / kHidden, j = row % kHidden;
        float v = 0.f;
        if (trunk == 0 && k == 0) v = w.value_fc1_w[j];
        if (trunk == 1 && k >= 1 && k <= A) v = w.policy_fc1_w[(k - 1) * kHidden + j];
        *reinterpret_cast<float*>(image + P::kW2T + trunk * kHidden * 32 + operand_offset<8>(j, k)) = to_tf32(v);
    }
}

template <int A>
__global__ void __launch_bounds__(kBwdTcThreads, 2) learner_bwd_tc_kernel(const float* __restrict__ obs, int64_t N,
                                                                           int T_split, int64_t B_split,
                                                                           const uint8_t* __restrict__ image,
                                                                           const float* __restrict__ d_logit,
                                                                           const float* __restrict__ d_v,
                                                                           float* __restrict__ partials) {
    using P = BwdTcPlan<A>;
    constexpr int KIN = P::KIN, KP = P::KP;
    constexpr int kSbo1 = (KP / 4) * 128;
    static_assert(A <= 4, "g[n] is staged as 8 floats");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane32 = tid & 31;
    const int trunk = blockIdx.x & 1;                         // 0: value trunk, 1: policy trunk
    const int cta = blockIdx.x >> 1, n_ctas = gridDim.x >> 1; // the two CTAs of a pair walk the same tiles
    const int quad = warp & 3, cpart = warp >> 2;             // TMEM lane quadrant (hidden units), 32-column (row) part of a stage
    const int j_local = quad * 32 + lane32;                   // hidden unit of the current 128-unit half == TMEM lane
    const uint32_t bar_img = smem_u32(smem + P::kBar), bar_mma = bar_img + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    if (warp == 0) tmem_alloc<256>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        mbar_init(bar_mma, 1);
        mbar_fence_init();
        // this trunk's slices of the weight image: three bulk copies counted on one barrier
        constexpr uint32_t kBytesIn = P::kTrunkBytes + kHidden * 4 + kHidden * 32;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_img), "r"(kBytesIn) : "memory");
        auto bulk = [&](int dst, const uint8_t* src, uint32_t bytes) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(smem + dst)),
                         "l"(src), "r"(bytes), "r"(bar_img)
                         : "memory");
        };
        bulk(P::kSW1, image + P::kB + trunk * P::kTrunkBytes, P::kTrunkBytes);
        bulk(P::kSB1, image + P::kB1 + trunk * kHidden * 4, kHidden * 4);
        bulk(P::kSW2T, image + P::kW2T + trunk * kHidden * 32, kHidden * 32);
    }
    // operand rows that are never written stay zero
    for (int i = tid; i < ((P::kNX + P::kNG) / 8) * P::kSboT / 4; i += kBwdTcThreads) reinterpret_cast<uint32_t*>(smem + P::kBX)[i] = 0u;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(bar_img, 0);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    {   // clear the gradient accumulators: columns [128, 256), 64 per thread of a lane
        uint32_t zero[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) zero[i] = 0u;
#pragma unroll
        for (int q = 0; q < 2; ++q) tcp::tmem_st32(tmem_lane + 128 + cpart * 64 + q * 32, zero);
        tcp::tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const float* b1 = reinterpret_cast<const float*>(smem + P::kSB1);
    auto off_t = [](int c, int n) { return (c >> 3) * P::kSboT + (n >> 2) * P::kLbo + (c & 7) * 16 + (n & 3) * 4; };
    float gsum[1 + A];
#pragma unroll
    for (int a = 0; a <= A; ++a) gsum[a] = 0.f;
    // this lane's hidden units (one per 128-unit half): first-layer bias where it does not ride in K
    float bias_j[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) bias_j[half] = P::kBiasInK ? 0.f : b1[half * 128 + j_local];

    uint32_t phase = 0;
    // Which tiles this CTA pair walks.  Flat mode (T_split == 0): tile u = rows [128 u, 128 u + 128) of the N rows, pair
    // `cta` takes u = cta, cta + n_ctas, ...  Split mode (the rows are a (T_split, B_split) trajectory): tiles never
    // straddle a half-move - half-move t has ceil(B / 128) tiles, the last one short - and pairs with an even index walk
    // the tiles of even t (player 0's steps), pairs with an odd index those of odd t (player 1's): the per-pair partial
    // sums then add up to one UNNORMALISED gradient per player (reduce_partials_kernel), which the caller divides by the
    // global step counts after the exchange.
    const bool split = T_split > 0;
    const int player = split ? (cta & 1) : 0;
    const int64_t tiles_per_t = split ? (B_split + kTileM - 1) / kTileM : 0;
    const int64_t my_first = split ? (cta >> 1) : cta, my_stride = split ? (n_ctas >> 1) : n_ctas;
    const int64_t num_tiles = split ? (int64_t)((T_split - player + 1) / 2) * tiles_per_t : (N + kTileM - 1) / kTileM;
    // threads 0..127 own one row of every tile; its observation and output gradients, one tile ahead
    float x_next[KIN], g_next[1 + A];
    auto load_tile_row = [&](int64_t u) {
        int64_t row = u * kTileM + tid;
        bool active = u < num_tiles && row < N;
        if (split) {
            const int64_t tt = 2 * (u / tiles_per_t) + player, j = (u % tiles_per_t) * kTileM + tid;
            row = tt * B_split + j;
            active = u < num_tiles && j < B_split;
        }
        load_row<KIN>(obs, active ? row : 0, active, x_next);
        g_next[0] = active ? __ldg(d_v + row) : 0.f;
#pragma unroll
        for (int a = 0; a < A; ++a) g_next[1 + a] = active ? __ldg(d_logit + row * A + a) : 0.f;
    };
    if (tid < kTileM) load_tile_row(my_first);
    for (int64_t tile = my_first; tile < num_tiles; tile += my_stride) {
        // ---- the tile's operands: observation tile (B of the recompute), x^T | 1 and g^T (B of the gradient MMAs);
        //      the row's data was loaded one tile ahead, the next tile's loads are issued right after it is consumed
        if (tid < kTileM) {
            const int n = tid;
            float x[KIN], g[1 + A];
#pragma unroll
            for (int k = 0; k < KIN; ++k) x[k] = x_next[k];
#pragma unroll
            for (int a = 0; a <= A; ++a) g[a] = g_next[a];
            load_tile_row(tile + my_stride);
            store_operand_row<KIN, KP, P::kBiasInK>(smem + P::kX, n, x);
#pragma unroll
            for (int k = 0; k < KIN; ++k) *reinterpret_cast<float*>(smem + P::kBX + off_t(k, n)) = to_tf32_fast(x[k]);
            *reinterpret_cast<float*>(smem + P::kBX + off_t(KIN, n)) = 1.f;
            float g8[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) g8[a] = a <= A ? to_tf32_fast(g[a < 1 + A ? a : 0]) : 0.f;
#pragma unroll
            for (int a = 0; a <= A; ++a) {
                *reinterpret_cast<float*>(smem + P::kBG + off_t(a, n)) = g8[a];
                gsum[a] += g[a];
            }
            // the same values row-major, as the B operand of S^T = W2^T . G^T
            *reinterpret_cast<float4*>(smem + P::kG + operand_offset<8>(n, 0)) = make_float4(g8[0], g8[1], g8[2], g8[3]);
            *reinterpret_cast<float4*>(smem + P::kG + operand_offset<8>(n, 4)) = make_float4(g8[4], g8[5], g8[6], g8[7]);
            fence_async_smem();
        }
        tc_fence_before();
        __syncthreads();

        // stage st = (hidden half, row half).  H^T = W1[half] . X^T[rows]  (M = hidden units, N = 64 rows, K = inputs)
        // into columns [0, 64);  S^T = W2^T[half] . G^T[rows] = (g W2)^T (K = the 1 + A outputs) into columns [64, 128)
        auto recompute = [&](int st) {
            const int half = st >> 1, rh = st & 1;
            const uint32_t a_base = smem_u32(smem + P::kSW1) + half * (128 / 8) * kSbo1;
            const uint32_t b_base = smem_u32(smem + P::kX) + rh * (64 / 8) * kSbo1;
#pragma unroll
            for (int ks = 0; ks < KP / 8; ++ks)
                mma_ss_n(tmem_base, make_desc<KP>(a_base + ks * 256), make_desc<KP>(b_base + ks * 256), idesc_tf32(64), ks > 0);
            mma_ss_n(tmem_base + 64, make_desc<8>(smem_u32(smem + P::kSW2T) + half * 128 * 32),
                     make_desc<8>(smem_u32(smem + P::kG) + rh * 64 * 32), idesc_tf32(64), false);
        };
        // (the whole warp runs this converged and one ELECTED lane issues: behind an `if (tid == 0)` ptxas wraps every
        //  UTCHMMA in a waterfall loop and the issue rate drops to one MMA per ~75 cycles)
        if (warp == 0) {
            tc_fence_after();
            if (tcp::elect_one()) {
                recompute(0);
                mma_commit(bar_mma);
            }
            __syncwarp();
        }
#pragma unroll 1
        for (int st = 0; st < 4; ++st) {
            const int half = st >> 1, rh = st & 1;
            const float bias = half ? bias_j[1] : bias_j[0];
            mbar_wait(bar_mma, phase);                        // H^T, S^T of this stage (and the gradient MMAs of the previous one)
            phase ^= 1u;
            tc_fence_after();
            // ---- this thread's 32 rows of its hidden unit: relu^T over H^T, dh^T = S^T where h > 0, both in place
            //      (the tensor core truncates these fp32 A operands to tf32)
            uint32_t hr[32], dh[32];
            tmem_ld32(tmem_lane + cpart * 32, hr);
            tmem_ld32(tmem_lane + 64 + cpart * 32, dh);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float h = __uint_as_float(hr[i]) + bias;
                const bool on = h > 0.f;
                hr[i] = on ? __float_as_uint(h) : 0u;
                dh[i] = on ? dh[i] : 0u;
            }
            tcp::tmem_st32(tmem_lane + cpart * 32, hr);
            tcp::tmem_st32(tmem_lane + 64 + cpart * 32, dh);
            tcp::tmem_st_wait();
            tc_fence_before();
            __syncthreads();
            // ---- D_w2[half] += relu^T BG^T, D_w1[half] += dh^T BX^T (K = the stage's 64 rows), then the next stage's
            //      H^T / S^T: the tensor core executes one thread's MMAs in order, so the recompute may overwrite the
            //      columns right behind the MMAs that read them, and one commit covers the three groups
            if (warp == 0) {
                tc_fence_after();
                const uint64_t bx = desc_lbo_sbo(smem_u32(smem + P::kBX) + rh * 16 * P::kLbo, P::kLbo, P::kSboT);
                const uint64_t bg = desc_lbo_sbo(smem_u32(smem + P::kBG) + rh * 16 * P::kLbo, P::kLbo, P::kSboT);
                if (tcp::elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        tcp::mma_ts(tmem_base + P::kColW2 + half * P::kNG, tmem_base + ks * 8, bg + (uint64_t)((ks * 2 * P::kLbo) >> 4),
                                    idesc_tf32(P::kNG), true);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        tcp::mma_ts(tmem_base + P::kColW1 + half * P::kNX, tmem_base + 64 + ks * 8,
                                    bx + (uint64_t)((ks * 2 * P::kLbo) >> 4), idesc_tf32(P::kNX), true);
                    if (st < 3) recompute(st + 1);
                    mma_commit(bar_mma);
                }
                __syncwarp();
            }
        }
        mbar_wait(bar_mma, phase);                            // the last gradient MMAs have read the tile's operands
        phase ^= 1u;
        tc_fence_after();
        __syncthreads();
    }

    // ---- this CTA's share of the pair's partial gradient, flat in state_dict order
    float* dst = partials + (int64_t)cta * P::kParams;
    if (cpart == 0) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int j = half * 128 + j_local;
            uint32_t acc[P::kNX];
#pragma unroll
            for (int q = 0; q < P::kNX / 16; ++q) tmem_ld16(tmem_lane + P::kColW1 + half * P::kNX + q * 16, acc + q * 16);
            tmem_ld_wait();
            float* w1_dst = dst + (trunk == 0 ? P::kOffV0w : P::kOffP0w) + j * KIN;
#pragma unroll
            for (int k = 0; k < KIN; ++k) w1_dst[k] = __uint_as_float(acc[k]);
            dst[(trunk == 0 ? P::kOffV0b : P::kOffP0b) + j] = __uint_as_float(acc[KIN]);
            uint32_t acc2[16];
            tmem_ld16(tmem_lane + P::kColW2 + half * P::kNG, acc2);
            tmem_ld_wait();
            if (trunk == 0) {
                dst[P::kOffV1w + j] = __uint_as_float(acc2[0]);
            } else {
#pragma unroll
                for (int a = 0; a < A; ++a) dst[P::kOffP1w + a * kHidden + j] = __uint_as_float(acc2[1 + a]);
            }
        }
    }
    // output-bias gradients: sums of g over the CTA's rows (threads 0..127 each own one row of every tile)
    float* s_red = reinterpret_cast<float*>(smem + P::kRed);
    if (tid < kTileM) {
#pragma unroll
        for (int a = 0; a <= A; ++a) {
            const float v = warp_sum(gsum[a]);
            if (lane32 == 0) s_red[warp * 8 + a] = v;
        }
    }
    __syncthreads();
    if (tid <= A) {
        const float v = (s_red[0 * 8 + tid] + s_red[1 * 8 + tid]) + (s_red[2 * 8 + tid] + s_red[3 * 8 + tid]);
        if (tid == 0 && trunk == 0) dst[P::kOffV1b] = v;
        if (tid > 0 && trunk == 1) dst[P::kOffP1b + tid - 1] = v;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
}

// -------------------------------------------------------- backward, software-pipelined
//
// The same mathematics as learner_bwd_tc_kernel, re-cut so that the tensor core never waits for the CUDA cores and
// vice versa: ONE CTA per SM owns all 512 tensor-memory columns and both trunks, and a stage's H^T / S^T live in one
// of TWO 128-column buffers.  A tile (128 rows) is eight stages s = (trunk, 128-unit hidden half, 64-row half);
//     consumers (16 warps, thread = hidden unit x 16 rows):  wait recompute(s) -> relu^T / dh^T in place in buffer s & 1
//     issuer (one elected lane, after the consumers' barrier): grad(s) = 16 MMAs reading buffer s & 1, then
//                                                              recompute(s + 2) INTO that buffer, one commit
// so while the consumers work on stage s + 1 (the other buffer, whose recompute was issued a stage earlier) the tensor
// core runs grad(s) and recompute(s + 2).  MMAs of one thread execute in issue order, which is what lets
// recompute(s + 2) overwrite the buffer grad(s) reads.  Four producer warps (one thread per tile row) build the next
// tile's operands - the observation tile, x^T | 1, g and g^T - in the other half of a double-buffered shared-memory
// region behind full / empty mbarriers, so the stream of stages never drains at a tile boundary.
// [learner_bwd_tc_kernel: two CTAs per SM, MMA -> elementwise -> MMA serialised per stage, tensor pipe 31 % active,
//  147 us at cfg2; this kernel: 128 us, 35 %.  What bounds a stage (measured, scripts/microbench/tmem_ldst.cu and
//  mma_shapes.cu): ~350 cycles of MMAs + ~320 cycles in which the consumers read and rewrite the stage's 64 KB of
//  tensor memory (tcgen05.ld -> ALU -> tcgen05.st sustains ~200 B / cycle / SM each way, the load side being the slow
//  one) + ~300 issue cycles of relu / mask arithmetic per SM sub-partition, and the tensor core's own tensor-memory
//  traffic (A operands and accumulators) does not overlap with the consumers' - a variant with three buffers and two
//  stages of slack ran no faster (133 us), i.e. the stage time is the SUM of these, not their maximum.]
#ifdef RNAD_TRACE_BWD
// development aid: cycle stamps of CTA 0, stages 16..79: [role][stage - 16][event]; role 0 = consumer warp 0, 1 = consumer
// warp 15, 2 = issuer 0, 3 = issuer 1, 4 = producer warp 0 (per tile)
__device__ long long g_bwd_trace[5][64][8];
#define BTR(role, s, ev) do { if (blockIdx.x == 0 && (s) >= 16 && (s) < 80 && lane32 == 0) g_bwd_trace[role][(s) - 16][ev] = clock64(); } while (0)
#else
#define BTR(role, s, ev) do { } while (0)
#endif
constexpr int kBwd2Consumers = 512, kBwd2Producers = 128, kBwd2Issuers = 64;
constexpr int kBwd2Threads = kBwd2Consumers + kBwd2Producers + kBwd2Issuers;

template <int A>
struct BwdTc2Plan : Shape<A> {
    using S = Shape<A>;
    using PT = BwdTcPlan<A>;
    static constexpr int kNX = PT::kNX, kNG = PT::kNG, kLbo = PT::kLbo, kSboT = PT::kSboT;
    // shared memory: the whole weight image of pack_bwd_tc_image_kernel (both trunks), then two tile buffers
    static constexpr int kSW1 = PT::kB, kSB1 = PT::kB1, kSW2T = PT::kW2T;
    static constexpr int kTile = round_up(PT::kImageBytes, 128);
    static constexpr int kX = 0;                                        // within a tile buffer
    static constexpr int kBX = kX + kTileM * S::KP * 4;
    static constexpr int kBG = kBX + (kNX / 8) * kSboT;
    static constexpr int kG = kBG + (kNG / 8) * kSboT;
    static constexpr int kTileBytes = round_up(kG + kTileM * 32, 128);
    static constexpr int kRed = kTile + 2 * kTileBytes;                 // [4 producer warps][8] sums of g
    static constexpr int kBar = kRed + 4 * 32;                          // image, recompute[2], consumed[2], full[2], empty[2]
    static constexpr int kTmem = kBar + 80;
    static constexpr int kBytes = kTmem + 16;
    // tensor memory: two stage buffers [H^T 64 | S^T 64], then per trunk 2 x kNX of D_w1 and 2 x 16 of D_w2
    static constexpr int kAcc = 256, kAccTrunk = 2 * kNX + 2 * kNG;
    static_assert(kAcc + 2 * kAccTrunk <= 512, "accumulators do not fit tensor memory");
    static_assert(kBytes <= 227 * 1024, "shared memory plan does not fit");
    static_assert(kTile % 16 == 0 && kBX % 16 == 0 && kBG % 16 == 0 && kG % 16 == 0 && kBar % 8 == 0, "alignment");
};

// Stage s of a CTA's stream: tile k = s >> 3; within the tile  half = s & 1 (the 128-unit hidden half - also the
// tensor-memory buffer and the issuer warp of the stage), trunk = (s >> 1) & 1, row half = (s >> 2) & 1.  Consecutive
// stages therefore add into DIFFERENT accumulators and every accumulator is only ever touched by ONE issuing thread,
// whose MMAs execute in issue order: the sums are bit-reproducible although two warps issue.
template <int A>
__global__ void __launch_bounds__(kBwd2Threads, 1) learner_bwd_tc2_kernel(const float* __restrict__ obs, int64_t N,
                                                                           int T_split, int64_t B_split,
                                                                           const uint8_t* __restrict__ image,
                                                                           const float* __restrict__ d_logit,
                                                                           const float* __restrict__ d_v,
                                                                           float* __restrict__ partials) {
    using P = BwdTc2Plan<A>;
    constexpr int KIN = P::KIN, KP = P::KP;
    constexpr int kSbo1 = (KP / 4) * 128;
    static_assert(A <= 4, "g[n] is staged as 8 floats");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane32 = tid & 31;
    const uint32_t bar_img = smem_u32(smem + P::kBar);
    auto bar_r = [&](int b) { return bar_img + 8 + 8 * b; };        // recompute into buffer b complete (and every MMA its issuer issued before)
    auto bar_c = [&](int b) { return bar_img + 24 + 8 * b; };       // the consumers are done with buffer b (relu^T / dh^T in place)
    auto bar_full = [&](int b) { return bar_img + 40 + 8 * b; };    // tile operands of shared-memory buffer b written
    auto bar_empty = [&](int b) { return bar_img + 56 + 8 * b; };   // every MMA reading shared-memory buffer b complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_r(b), 1);
            mbar_init(bar_c(b), kBwd2Consumers / 32);
            mbar_init(bar_full(b), kBwd2Producers / 32);
            mbar_init(bar_empty(b), 2);                              // one commit per issuer
        }
        mbar_fence_init();
        tma_bulk_load(smem, image, BwdTcPlan<A>::kImageBytes, bar_img);
    }
    // operand rows that are never written stay zero (both tile buffers)
    for (int i = tid; i < 2 * P::kTileBytes / 4; i += kBwd2Threads) reinterpret_cast<uint32_t*>(smem + P::kTile)[i] = 0u;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid < kBwd2Consumers) {
        // clear the gradient accumulators: columns [256, 256 + 2 * kAccTrunk), 16 at a time, split over the four column parts
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int c = P::kAcc + (warp >> 2) * 16; c < P::kAcc + 2 * P::kAccTrunk; c += 64)
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_base + c),
                "r"(0u)
                : "memory");
        tcp::tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // which tiles this CTA walks: as in learner_bwd_tc_kernel, with the CTA in the role of the CTA pair
    const int cta = blockIdx.x, n_ctas = gridDim.x;
    const bool split = T_split > 0;
    const int player = split ? (cta & 1) : 0;
    const int64_t tiles_per_t = split ? (B_split + kTileM - 1) / kTileM : 0;
    const int64_t my_first = split ? (cta >> 1) : cta, my_stride = split ? (n_ctas >> 1) : n_ctas;
    const int64_t num_tiles = split ? (int64_t)((T_split - player + 1) / 2) * tiles_per_t : (N + kTileM - 1) / kTileM;
    const int64_t my_tiles = my_first < num_tiles ? (num_tiles - 1 - my_first) / my_stride + 1 : 0;
    const int64_t n_stages = my_tiles * 8;
    float* dst = partials + (int64_t)cta * P::kParams;

    if (tid >= kBwd2Consumers + kBwd2Producers) {
        // ------------------------------------------------------------ issuers: warp b issues the stages with s & 1 == b
        const int b = warp - (kBwd2Consumers + kBwd2Producers) / 32;
        tcp::mbar_wait_c(bar_img, 0);
        int64_t seen_full = -1;
        auto need_tile = [&](int64_t k) {       // (whole warp) the producers have written tile k's operands
            if (k != seen_full) {
                tcp::mbar_wait_c(bar_full((int)(k & 1)), (uint32_t)(k >> 1) & 1u);
                seen_full = k;
            }
        };
        auto recompute = [&](int64_t s) {      // H^T | S^T of stage s into tensor-memory buffer b (elected lane)
            const int64_t k = s >> 3;
            const int trunk = (int)(s >> 1) & 1, rh = (int)(s >> 2) & 1;
            const uint32_t tile = smem_u32(smem + P::kTile + (int)(k & 1) * P::kTileBytes);
            const uint32_t d = tmem_base + (uint32_t)b * 128;
            const uint32_t a_base = smem_u32(smem + P::kSW1) + trunk * P::kTrunkBytes + b * (128 / 8) * kSbo1;
            const uint32_t b_base = tile + P::kX + rh * (64 / 8) * kSbo1;
#pragma unroll
            for (int ks = 0; ks < KP / 8; ++ks)
                mma_ss_n(d, make_desc<KP>(a_base + ks * 256), make_desc<KP>(b_base + ks * 256), idesc_tf32(64), ks > 0);
            mma_ss_n(d + 64, make_desc<8>(smem_u32(smem + P::kSW2T) + trunk * kHidden * 32 + b * 128 * 32),
                     make_desc<8>(tile + P::kG + rh * 64 * 32), idesc_tf32(64), false);
        };
        if (b < n_stages) {                     // fill the pipeline: stage b
            need_tile(0);
            tc_fence_after();
            if (tcp::elect_one()) {
                recompute(b);
                mma_commit(bar_r(b));
            }
            __syncwarp();
        }
#pragma unroll 1
        for (int64_t s = b; s < n_stages; s += 2) {
            const int trunk = (int)(s >> 1) & 1, rh = (int)(s >> 2) & 1;
            const int64_t k = s >> 3;
            const bool more = s + 2 < n_stages;
            BTR(2 + b, s, 0);
            if (more) need_tile((s + 2) >> 3);
            BTR(2 + b, s, 1);
            tcp::mbar_wait_c(bar_c(b), (uint32_t)(s >> 1) & 1u);            // relu^T / dh^T of stage s are in buffer b
            tc_fence_after();
            BTR(2 + b, s, 2);
            // grad(s): D_w2 += relu^T BG^T, D_w1 += dh^T BX^T (K = the stage's 64 rows); then recompute(s + 2) into the buffer
            // just read; one commit covers both groups
            const uint32_t tile = smem_u32(smem + P::kTile + (int)(k & 1) * P::kTileBytes);
            const uint64_t bx = desc_lbo_sbo(tile + P::kBX + rh * 16 * P::kLbo, P::kLbo, P::kSboT);
            const uint64_t bg = desc_lbo_sbo(tile + P::kBG + rh * 16 * P::kLbo, P::kLbo, P::kSboT);
            const uint32_t acc = tmem_base + P::kAcc + trunk * P::kAccTrunk;
            const uint32_t buf = tmem_base + (uint32_t)b * 128;
            if (tcp::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    tcp::mma_ts(acc + 2 * P::kNX + b * P::kNG, buf + ks * 8, bg + (uint64_t)((ks * 2 * P::kLbo) >> 4),
                                idesc_tf32(P::kNG), true);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    tcp::mma_ts(acc + b * P::kNX, buf + 64 + ks * 8, bx + (uint64_t)((ks * 2 * P::kLbo) >> 4),
                                idesc_tf32(P::kNX), true);
                if ((s & 7) >= 6) mma_commit(bar_empty((int)(k & 1)));   // this issuer's last reads of the tile's shared-memory operands
                if (more) recompute(s + 2);
                mma_commit(bar_r(b));
            }
            __syncwarp();
            BTR(2 + b, s, 3);
        }
    } else if (tid >= kBwd2Consumers) {
        // ------------------------------------------------------------ producers: one thread per tile row
        const int n = tid - kBwd2Consumers, pw = n >> 5;
        auto off_t = [](int c, int nn) { return (c >> 3) * P::kSboT + (nn >> 2) * P::kLbo + (c & 7) * 16 + (nn & 3) * 4; };
        float gsum[1 + A];
#pragma unroll
        for (int a = 0; a <= A; ++a) gsum[a] = 0.f;
        for (int64_t k = 0; k < my_tiles; ++k) {
            const int64_t u = my_first + k * my_stride;
            int64_t row = u * kTileM + n;
            bool active = row < N;
            if (split) {
                const int64_t tt = 2 * (u / tiles_per_t) + player, j = (u % tiles_per_t) * kTileM + n;
                row = tt * B_split + j;
                active = j < B_split;
            }
            float x[KIN], g[1 + A];
            load_row<KIN>(obs, active ? row : 0, active, x);
            g[0] = active ? __ldg(d_v + row) : 0.f;
#pragma unroll
            for (int a = 0; a < A; ++a) g[1 + a] = active ? __ldg(d_logit + row * A + a) : 0.f;
            const int tb = (int)(k & 1);
            if (pw == 0) BTR(4, 8 * k, 0);
            if (k >= 2) tcp::mbar_wait_c(bar_empty(tb), (uint32_t)((k >> 1) - 1) & 1u);   // the MMAs of tile k - 2 are done with it
            if (pw == 0) BTR(4, 8 * k, 1);
            uint8_t* tile = smem + P::kTile + tb * P::kTileBytes;
            store_operand_row<KIN, KP, P::kBiasInK>(tile + P::kX, n, x);
#pragma unroll
            for (int kk = 0; kk < KIN; ++kk) *reinterpret_cast<float*>(tile + P::kBX + off_t(kk, n)) = to_tf32_fast(x[kk]);
            *reinterpret_cast<float*>(tile + P::kBX + off_t(KIN, n)) = 1.f;
            float g8[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) g8[a] = a <= A ? to_tf32_fast(g[a < 1 + A ? a : 0]) : 0.f;
#pragma unroll
            for (int a = 0; a <= A; ++a) {
                *reinterpret_cast<float*>(tile + P::kBG + off_t(a, n)) = g8[a];
                gsum[a] += g[a];
            }
            *reinterpret_cast<float4*>(tile + P::kG + operand_offset<8>(n, 0)) = make_float4(g8[0], g8[1], g8[2], g8[3]);
            *reinterpret_cast<float4*>(tile + P::kG + operand_offset<8>(n, 4)) = make_float4(g8[4], g8[5], g8[6], g8[7]);
            fence_async_smem();
            __syncwarp();
            if (lane32 == 0) tcp::mbar_arrive(bar_full(tb));
            if (pw == 0) BTR(4, 8 * k, 2);
        }
        // output-bias gradients: sums of g over the CTA's rows
        float* s_red = reinterpret_cast<float*>(smem + P::kRed);
#pragma unroll
        for (int a = 0; a <= A; ++a) {
            const float v = warp_sum(gsum[a]);
            if (lane32 == 0) s_red[pw * 8 + a] = v;
        }
    } else {
        // ------------------------------------------------------------ consumers: thread = hidden unit x 16 rows of a stage
        const int quad = warp & 3, cpart = warp >> 2;             // TMEM lane quadrant (hidden unit), 16-column (row) part of a stage
        const int j_local = quad * 32 + lane32;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        tcp::mbar_wait_c(bar_img, 0);
        const float* b1 = reinterpret_cast<const float*>(smem + P::kSB1);
        float bias_j[4];                                           // [trunk][half]
#pragma unroll
        for (int c = 0; c < 4; ++c) bias_j[c] = P::kBiasInK ? 0.f : b1[(c >> 1) * kHidden + (c & 1) * 128 + j_local];
#pragma unroll 1
        for (int64_t s = 0; s < n_stages; ++s) {
            const int b = (int)(s & 1), trunk = (int)(s >> 1) & 1;
            const float bias = bias_j[trunk * 2 + b];
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 0);
            tcp::mbar_wait_c(bar_r(b), (uint32_t)(s >> 1) & 1u);
            tc_fence_after();
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 1);
            // ---- this thread's 16 rows of its hidden unit: relu^T over H^T, dh^T = S^T where h > 0, both in place
            uint32_t hr[16], dh[16];
            const uint32_t th = tmem_lane + b * 128 + cpart * 16;
            tmem_ld16(th, hr);
            tmem_ld16(th + 64, dh);
            tmem_ld_wait();
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 2);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float h = __uint_as_float(hr[i]) + bias;
                const bool on = h > 0.f;
                hr[i] = on ? __float_as_uint(h) : 0u;
                dh[i] = on ? dh[i] : 0u;
            }
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(th),
                "r"(hr[0]), "r"(hr[1]), "r"(hr[2]), "r"(hr[3]), "r"(hr[4]), "r"(hr[5]), "r"(hr[6]), "r"(hr[7]), "r"(hr[8]),
                "r"(hr[9]), "r"(hr[10]), "r"(hr[11]), "r"(hr[12]), "r"(hr[13]), "r"(hr[14]), "r"(hr[15])
                : "memory");
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(th + 64),
                "r"(dh[0]), "r"(dh[1]), "r"(dh[2]), "r"(dh[3]), "r"(dh[4]), "r"(dh[5]), "r"(dh[6]), "r"(dh[7]), "r"(dh[8]),
                "r"(dh[9]), "r"(dh[10]), "r"(dh[11]), "r"(dh[12]), "r"(dh[13]), "r"(dh[14]), "r"(dh[15])
                : "memory");
            tcp::tmem_st_wait();
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 3);
            tc_fence_before();
            __syncwarp();
            if (lane32 == 0) tcp::mbar_arrive(bar_c(b));
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 4);
        }
        // every gradient MMA complete: the last commit of each issuer
        if (n_stages >= 2) {
            tcp::mbar_wait_c(bar_r(0), (uint32_t)(n_stages >> 1) & 1u);
            tcp::mbar_wait_c(bar_r(1), (uint32_t)(n_stages >> 1) & 1u);
        }
        tc_fence_after();

        // ---- this CTA's partial gradient, flat in state_dict order: column part c reads (trunk, half) = (c >> 1, c & 1)
        {
            const int trunk = cpart >> 1, half = cpart & 1;
            const int j = half * 128 + j_local;
            const uint32_t acc = tmem_lane + P::kAcc + trunk * P::kAccTrunk;
            uint32_t w[P::kNX];
#pragma unroll
            for (int q = 0; q < P::kNX / 16; ++q) tmem_ld16(acc + half * P::kNX + q * 16, w + q * 16);
            tmem_ld_wait();
            float* w1_dst = dst + (trunk == 0 ? P::kOffV0w : P::kOffP0w) + j * KIN;
#pragma unroll
            for (int kk = 0; kk < KIN; ++kk) w1_dst[kk] = __uint_as_float(w[kk]);
            dst[(trunk == 0 ? P::kOffV0b : P::kOffP0b) + j] = __uint_as_float(w[KIN]);
            uint32_t w2[16];
            tmem_ld16(acc + 2 * P::kNX + half * P::kNG, w2);
            tmem_ld_wait();
            if (trunk == 0) {
                dst[P::kOffV1w + j] = __uint_as_float(w2[0]);
            } else {
#pragma unroll
                for (int a = 0; a < A; ++a) dst[P::kOffP1w + a * kHidden + j] = __uint_as_float(w2[1 + a]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid <= A) {
        const float* s_red = reinterpret_cast<const float*>(smem + P::kRed);
        const float v = (s_red[0 * 8 + tid] + s_red[1 * 8 + tid]) + (s_red[2 * 8 + tid] + s_red[3 * 8 + tid]);
        if (tid == 0) dst[P::kOffV1b] = v;
        else dst[P::kOffP1b + tid - 1] = v;
    }
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

#ifdef RNAD_TRACE_BWD
}  // namespace tc
}  // namespace rnad
extern "C" __attribute__((visibility("default"))) int rnad_debug_bwd_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, rnad::tc::g_bwd_trace, sizeof(rnad::tc::g_bwd_trace));
}
namespace rnad {
namespace tc {
#endif

// Sum of the per-CTA partial gradients in a FIXED order (deterministic): eight lanes per parameter take the partials
// p = lane, lane + 8, ... (eight loads in flight per thread as well) and meet in a shuffle tree.
// Split mode (gridDim.y == 2): blockIdx.y = player; player p's partials are the rows p, p + 2, ... (pairs with index
// parity p) and its sum goes to flat_grad + p * n_params.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_parts, int n_params,
                                       float* __restrict__ flat_grad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 3, sub = t & 7;
    const int step = gridDim.y, first = blockIdx.y;      // 1, 0 in flat mode
    float acc = 0.f;
    if (i < n_params) {
        float a0 = 0.f, a1 = 0.f;
        int p = first + sub * step;
        for (; p + 8 * step < n_parts; p += 16 * step) {
            a0 += partials[(int64_t)p * n_params + i];
            a1 += partials[(int64_t)(p + 8 * step) * n_params + i];
        }
        if (p < n_parts) a0 += partials[(int64_t)p * n_params + i];
        acc = a0 + a1;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (i < n_params && sub == 0) flat_grad[(int64_t)blockIdx.y * n_params + i] = acc;
}


// the forward kernels' weight image sits at the start of the workspace: room for the larger of the two builds
template <int A>
int64_t fwd_image_reserve() {
    const int64_t v1 = FwdPlan<A>::kImageBytes, v2 = learner_forward_tc2_image_bytes(A);
    return round_up((int)(v1 > v2 ? v1 : v2), 256);
}

// then the backward's images: the tf32 one (pack_bwd_tc_image_kernel) and, where that engine exists, the fp16 one
template <int A>
int64_t bwd_image_reserve() {
    return round_up(BwdTcPlan<A>::kImageBytes, 256) + round_up((int)learner_backward_f16_image_bytes(A), 256);
}

template <int A>
int64_t workspace_bytes() {
    return fwd_image_reserve<A>() + bwd_image_reserve<A>() + (int64_t)kMaxBwdCtas * Shape<A>::kParams * 4;
}

template <int A, typename Kernel>
int prepare(Kernel kernel, size_t smem, const char* what) {
    int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), what);
    if (rc) return rc;
    return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared), what);
}

template <int A>
int launch_forward(const float* obs, int64_t N, const FwdNets& nets, const FwdOut& out, uint8_t* workspace,
                   cudaStream_t st) {
    using P = FwdPlan<A>;
    static_assert(P::kBytes <= 227 * 1024, "forward image does not fit in shared memory");
    pack_fwd_image_kernel<A><<<48, 256, 0, st>>>(nets, workspace);
    RNAD_CHECK_LAUNCH("pack_fwd_image_kernel");
    int rc = prepare<A>(learner_fwd_kernel<A>, P::kBytes, "cudaFuncSetAttribute(learner_fwd)");
    if (rc) return rc;
    int64_t blocks = (N + kTileM - 1) / kTileM;
    if (blocks > sm_count()) blocks = sm_count();
    learner_fwd_kernel<A><<<(int)blocks, kLearnThreads, P::kBytes, st>>>(obs, N, workspace, out);
    RNAD_CHECK_LAUNCH("learner_fwd_kernel");
    return RNAD_OK;
}

// T_split > 0: the rows are a (T_split, B_split) trajectory and flat_grad receives TWO unnormalised gradients,
// player 0's (rows of even t) then player 1's (see learner_bwd_tc_kernel).
template <int A>
int launch_backward(const float* obs, int64_t N, int T_split, int64_t B_split, const rnad_mlp_weights& w,
                    const float* d_logit, const float* d_v, float* flat_grad, uint8_t* workspace, cudaStream_t st,
                    int mode = 0) {
    using P = BwdPlan<A>;
    using PT = BwdTcPlan<A>;
    static_assert(PT::kImageBytes >= P::kImageBytes, "the workspace reserves the larger image");
    uint8_t* image = workspace + fwd_image_reserve<A>();
    uint8_t* image_f16 = image + round_up(PT::kImageBytes, 256);
    float* partials = reinterpret_cast<float*>(image + bwd_image_reserve<A>());
    int64_t blocks = (N + kTileM - 1) / kTileM;
    const int cap = sm_count() < kMaxBwdCtas ? sm_count() : kMaxBwdCtas;
    if (T_split > 0) {
        // an even number of CTA pairs, half of them per player; at least one pair each (a player without rows writes zeros)
        const int64_t per_player = (int64_t)((T_split + 1) / 2) * ((B_split + kTileM - 1) / kTileM);
        blocks = 2 * (per_player < cap / 2 ? per_player : cap / 2);
    } else if (blocks > cap) {
        blocks = cap;
    }
    static const bool cuda_core_reduction = getenv("RNAD_LEARNER_BWD_CUDA_CORES") != nullptr;   // the previous kernel, for A/B runs
    if (cuda_core_reduction && T_split == 0 && mode == 0) {
        pack_bwd_image_kernel<A><<<24, 256, 0, st>>>(w, image);
        RNAD_CHECK_LAUNCH("pack_bwd_image_kernel");
        int rc = prepare<A>(learner_bwd_kernel<A>, P::kBytes, "cudaFuncSetAttribute(learner_bwd)");
        if (rc) return rc;
        learner_bwd_kernel<A><<<(int)blocks, kLearnThreads, P::kBytes, st>>>(obs, N, image, w, d_logit, d_v, partials);
        RNAD_CHECK_LAUNCH("learner_bwd_kernel");
    } else {
        // split mode (one UNNORMALISED gradient per player: the learner step) runs on the fp16-operand engine
        // (learner_bwd_f16.cu) where it exists; RNAD_LEARNER_BWD_TF32 keeps the tf32 kernels there too, for A/B runs
        static const bool tf32_only = getenv("RNAD_LEARNER_BWD_TF32") != nullptr || getenv("RNAD_LEARNER_BWD_V1") != nullptr ||
                                      getenv("RNAD_LEARNER_BWD_V2") != nullptr;
        const bool f16_engine = !tf32_only && learner_backward_f16_supported(A);
        if (f16_engine && (T_split > 0 || mode == 2)) {
            int rc = learner_backward_f16(A, obs, N, T_split, B_split, w, d_logit, d_v, image_f16, partials, (int)blocks, st, mode);
            if (rc) return rc;
            if (mode != 2) {
                reduce_partials_kernel<<<dim3((P::kParams * 8 + 255) / 256, 2), 256, 0, st>>>(partials, (int)blocks, P::kParams, flat_grad);
                RNAD_CHECK_LAUNCH("reduce_partials_kernel");
            }
            return RNAD_OK;      // (pack only: the prepacked entry point is the split one, which this engine serves)
        }
        if (mode != 1) {      // (mode as in learner_fwd_tc2.cu: 0 pack + run, 1 prepacked, 2 pack only)
            pack_bwd_tc_image_kernel<A><<<32, 256, 0, st>>>(w, image);
            RNAD_CHECK_LAUNCH("pack_bwd_tc_image_kernel");
            if (mode == 2) return RNAD_OK;
        }
        static const bool two_ctas = getenv("RNAD_LEARNER_BWD_V1") != nullptr;   // the first kernel (two CTAs per SM), for A/B runs
        static const bool v2 = getenv("RNAD_LEARNER_BWD_V2") != nullptr;         // the S^T formulation where the mask one fits, for A/B runs
        static_assert(PT::kB == 0 && PT::kB1 == 2 * PT::kTrunkBytes, "learner_bwd_tc3.cu reads the head of this image");
        if (!two_ctas && !v2 && learner_backward_tc3_supported(A)) {
            int rc = learner_backward_tc3(A, obs, N, T_split, B_split, w, d_logit, d_v, image, partials, (int)blocks, st);
            if (rc) return rc;
            reduce_partials_kernel<<<dim3((P::kParams * 8 + 255) / 256, T_split > 0 ? 2 : 1), 256, 0, st>>>(
                partials, (int)blocks, P::kParams, flat_grad);
            RNAD_CHECK_LAUNCH("reduce_partials_kernel");
            return RNAD_OK;
        }
        if (!two_ctas) {
            using P2 = BwdTc2Plan<A>;
            // one CTA per SM (all 512 tensor-memory columns): more than half of the shared memory keeps a second one out
            const size_t smem2 = P2::kBytes > 116 * 1024 ? P2::kBytes : 116 * 1024;
            int rc = prepare<A>(learner_bwd_tc2_kernel<A>, smem2, "cudaFuncSetAttribute(learner_bwd_tc2)");
            if (rc) return rc;
            learner_bwd_tc2_kernel<A><<<(int)blocks, kBwd2Threads, smem2, st>>>(obs, N, T_split, B_split, image, d_logit, d_v,
                                                                               partials);
            RNAD_CHECK_LAUNCH("learner_bwd_tc2_kernel");
            reduce_partials_kernel<<<dim3((P::kParams * 8 + 255) / 256, T_split > 0 ? 2 : 1), 256, 0, st>>>(
                partials, (int)blocks, P::kParams, flat_grad);
            RNAD_CHECK_LAUNCH("reduce_partials_kernel");
            return RNAD_OK;
        }
        // two CTAs per SM (one per trunk, 256 TMEM columns each): pad the shared-memory request so that a third can
        // never become resident and spin inside tcgen05.alloc
        size_t smem = PT::kBytes;
        const size_t floor_two_per_sm = 227 * 1024 / 3 + 1024;
        if (smem < floor_two_per_sm) smem = floor_two_per_sm;
        int rc = prepare<A>(learner_bwd_tc_kernel<A>, smem, "cudaFuncSetAttribute(learner_bwd_tc)");
        if (rc) return rc;
        learner_bwd_tc_kernel<A><<<2 * (int)blocks, kBwdTcThreads, smem, st>>>(obs, N, T_split, B_split, image, d_logit,
                                                                               d_v, partials);
        RNAD_CHECK_LAUNCH("learner_bwd_tc_kernel");
    }
    reduce_partials_kernel<<<dim3((P::kParams * 8 + 255) / 256, T_split > 0 ? 2 : 1), 256, 0, st>>>(
        partials, (int)blocks, P::kParams, flat_grad);
    RNAD_CHECK_LAUNCH("reduce_partials_kernel");
    return RNAD_OK;
}

}  // namespace tc
}  // namespace rnad

using namespace rnad;

static bool weights_ok(const rnad_mlp_weights* w) {
    return w && w->value_fc0_w && w->value_fc0_b && w->value_fc1_w && w->value_fc1_b && w->policy_fc0_w &&
           w->policy_fc0_b && w->policy_fc1_w && w->policy_fc1_b;
}

extern "C" {

int rnad_learner_mlp_supported(int A, int width) { return (width == tc::kHidden && A >= 2 && A <= 4) ? 1 : 0; }

int64_t rnad_learner_mlp_workspace_bytes(int A, int width) {
    if (!rnad_learner_mlp_supported(A, width)) return 0;
    switch (A) {
        case 2: return tc::workspace_bytes<2>();
        case 3: return tc::workspace_bytes<3>();
        case 4: return tc::workspace_bytes<4>();
    }
    return 0;
}

int rnad_learner_param_count(int A, int width) {
    if (!rnad_learner_mlp_supported(A, width)) return 0;
    return 2 * width * (2 * A * A) + 2 * width + width + 1 + A * width + A;
}

static int learner_forward_impl(const float* observations, int64_t N, int A, const rnad_mlp_weights* net,
                                const rnad_mlp_weights* target, const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_,
                                const rnad_learner_fwd_out* out, void* workspace, void* stream, int prepacked,
                                int others_only = 0) {
    RNAD_REQUIRE(observations && out && workspace, "rnad_learner_forward: null pointer");
    RNAD_REQUIRE(weights_ok(net) && weights_ok(target) && weights_ok(reg) && weights_ok(reg_),
                 "rnad_learner_forward: null weight pointer");
    RNAD_REQUIRE(out->v_target && out->log_pi_reg && out->log_pi_reg_ &&
                     (others_only || (out->logit && out->pi && out->log_pi && out->v)),
                 "rnad_learner_forward: null output pointer");
    RNAD_REQUIRE(N >= 0, "rnad_learner_forward: negative row count");
    if (!rnad_learner_mlp_supported(A, net->width) || target->width != net->width || reg->width != net->width ||
        reg_->width != net->width) {
        set_error("rnad_learner_forward: needs four nets of width 256 and 2 <= max_actions <= 4");
        return RNAD_EUNSUPPORTED;
    }
    RNAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rnad_learner_forward: workspace must be 256-byte aligned");
    if (N == 0) return RNAD_OK;
    static const bool v1 = getenv("RNAD_LEARNER_FWD_V1") != nullptr;   // the previous kernel, for A/B runs
    if (!v1 && learner_forward_tc2_supported(A, net->width))
        return learner_forward_tc2(observations, N, A, net, target, reg, reg_, out, workspace, (cudaStream_t)stream,
                                   prepacked ? 1 : 0, others_only != 0);
    RNAD_REQUIRE(!prepacked && !others_only, "rnad_learner_forward_prepacked: not available with RNAD_LEARNER_FWD_V1");
    tc::FwdNets nets{*net, *target, *reg, *reg_};
    tc::FwdOut o{out->logit, out->pi, out->log_pi, out->v, out->v_target, out->log_pi_reg, out->log_pi_reg_};
    cudaStream_t st = (cudaStream_t)stream;
    switch (A) {
        case 2: return tc::launch_forward<2>(observations, N, nets, o, (uint8_t*)workspace, st);
        case 3: return tc::launch_forward<3>(observations, N, nets, o, (uint8_t*)workspace, st);
        case 4: return tc::launch_forward<4>(observations, N, nets, o, (uint8_t*)workspace, st);
    }
    return RNAD_EINVAL;
}

int rnad_learner_forward(const float* observations, int64_t N, int A, const rnad_mlp_weights* net,
                         const rnad_mlp_weights* target, const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_,
                         const rnad_learner_fwd_out* out, void* workspace, void* stream) {
    return learner_forward_impl(observations, N, A, net, target, reg, reg_, out, workspace, stream, 0);
}

int rnad_learner_forward_prepacked(const float* observations, int64_t N, int A, const rnad_mlp_weights* net,
                                   const rnad_mlp_weights* target, const rnad_mlp_weights* reg,
                                   const rnad_mlp_weights* reg_, const rnad_learner_fwd_out* out, int others_only,
                                   void* workspace, void* stream) {
    return learner_forward_impl(observations, N, A, net, target, reg, reg_, out, workspace, stream, 1, others_only);
}

int rnad_learner_pack(int A, const rnad_mlp_weights* net, const rnad_mlp_weights* target, const rnad_mlp_weights* reg,
                      const rnad_mlp_weights* reg_, int others_only, void* workspace, void* stream) {
    RNAD_REQUIRE(workspace, "rnad_learner_pack: null pointer");
    RNAD_REQUIRE(weights_ok(net) && weights_ok(target) && weights_ok(reg) && weights_ok(reg_), "rnad_learner_pack: null weight pointer");
    if (!rnad_learner_mlp_supported(A, net->width) || target->width != net->width || reg->width != net->width ||
        reg_->width != net->width) {
        set_error("rnad_learner_pack: needs four nets of width 256 and 2 <= max_actions <= 4");
        return RNAD_EUNSUPPORTED;
    }
    RNAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rnad_learner_pack: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = learner_forward_tc2(nullptr, 0, A, net, target, reg, reg_, nullptr, workspace, st, 2, others_only != 0);
    if (rc) return rc;
    switch (A) {
        case 2: return tc::launch_backward<2>(nullptr, 0, 0, 0, *net, nullptr, nullptr, nullptr, (uint8_t*)workspace, st, 2);
        case 3: return tc::launch_backward<3>(nullptr, 0, 0, 0, *net, nullptr, nullptr, nullptr, (uint8_t*)workspace, st, 2);
        case 4: return tc::launch_backward<4>(nullptr, 0, 0, 0, *net, nullptr, nullptr, nullptr, (uint8_t*)workspace, st, 2);
    }
    return RNAD_EINVAL;
}

int rnad_learner_backward(const float* observations, int64_t N, int A, const rnad_mlp_weights* net,
                          const float* d_logit, const float* d_v, float* flat_grad, void* workspace, void* stream) {
    RNAD_REQUIRE(observations && d_logit && d_v && flat_grad && workspace, "rnad_learner_backward: null pointer");
    RNAD_REQUIRE(weights_ok(net), "rnad_learner_backward: null weight pointer");
    RNAD_REQUIRE(N >= 1, "rnad_learner_backward: empty batch");
    if (!rnad_learner_mlp_supported(A, net->width)) {
        set_error("rnad_learner_backward: needs width 256 and 2 <= max_actions <= 4");
        return RNAD_EUNSUPPORTED;
    }
    RNAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rnad_learner_backward: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    switch (A) {
        case 2: return tc::launch_backward<2>(observations, N, 0, 0, *net, d_logit, d_v, flat_grad, (uint8_t*)workspace, st);
        case 3: return tc::launch_backward<3>(observations, N, 0, 0, *net, d_logit, d_v, flat_grad, (uint8_t*)workspace, st);
        case 4: return tc::launch_backward<4>(observations, N, 0, 0, *net, d_logit, d_v, flat_grad, (uint8_t*)workspace, st);
    }
    return RNAD_EINVAL;
}

static int learner_backward_split_impl(const float* observations, int T, int64_t B, int A, const rnad_mlp_weights* net,
                                       const float* d_logit, const float* d_v, float* player_grads, void* workspace,
                                       void* stream, int prepacked) {
    RNAD_REQUIRE(observations && d_logit && d_v && player_grads && workspace, "rnad_learner_backward_split: null pointer");
    RNAD_REQUIRE(weights_ok(net), "rnad_learner_backward_split: null weight pointer");
    RNAD_REQUIRE(T >= 1 && B >= 1, "rnad_learner_backward_split: empty trajectory");
    if (!rnad_learner_mlp_supported(A, net->width)) {
        set_error("rnad_learner_backward_split: needs width 256 and 2 <= max_actions <= 4");
        return RNAD_EUNSUPPORTED;
    }
    RNAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rnad_learner_backward_split: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = (int64_t)T * B;
    switch (A) {
        case 2: return tc::launch_backward<2>(observations, N, T, B, *net, d_logit, d_v, player_grads, (uint8_t*)workspace, st, prepacked);
        case 3: return tc::launch_backward<3>(observations, N, T, B, *net, d_logit, d_v, player_grads, (uint8_t*)workspace, st, prepacked);
        case 4: return tc::launch_backward<4>(observations, N, T, B, *net, d_logit, d_v, player_grads, (uint8_t*)workspace, st, prepacked);
    }
    return RNAD_EINVAL;
}

int rnad_learner_backward_split(const float* observations, int T, int64_t B, int A, const rnad_mlp_weights* net,
                                const float* d_logit, const float* d_v, float* player_grads, void* workspace,
                                void* stream) {
    return learner_backward_split_impl(observations, T, B, A, net, d_logit, d_v, player_grads, workspace, stream, 0);
}

int rnad_learner_backward_split_prepacked(const float* observations, int T, int64_t B, int A, const rnad_mlp_weights* net,
                                          const float* d_logit, const float* d_v, float* player_grads, void* workspace,
                                          void* stream) {
    return learner_backward_split_impl(observations, T, B, A, net, d_logit, d_v, player_grads, workspace, stream, 1);
}

}  // extern "C"
