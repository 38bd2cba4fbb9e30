// K3: policy post-processing, two-player v-trace, NeuRD / critic targets and the
// analytic loss gradients, as one reverse pass over trajectory time per game.
// Reference: learn/vtrace.py:24-55 (process_policy), 70-87, 141-204 (helpers),
// 207-352 (v_trace), 355-431 (losses); learn/rnad.py:365-425 (reward transform
// and glue).  Arithmetic is written with explicit round-to-nearest intrinsics
// in the reference's operation order (no FMA contraction), so that results
// agree with the reference's fp32 torch-CPU ops bit for bit wherever the
// reference's own summation order is reproducible (A <= 5; see sum_lanes).
#include <cstdlib>

#include "common.cuh"

namespace rnad {

namespace {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }

// torch's CPU sum over a short contiguous last dimension: four lane
// accumulators filled round-robin, combined left to right.  For A <= 4 this is
// the plain left-to-right sum.
template <int A>
__device__ __forceinline__ float sum_lanes(const float (&x)[A]) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < A; ++a) acc[a & 3] = a < 4 ? x[a] : add(acc[a & 3], x[a]);
    float s = acc[0];
#pragma unroll
    for (int j = 1; j < 4; ++j)
        if (j < A) s = add(s, acc[j]);
    return s;
}

// vtrace.py:24-55.  Ties in the descending order go to the lower action id.
template <int A>
__device__ __forceinline__ void process_policy_row(const float (&pi)[A], const float (&mask)[A], float n_disc,
                                                   float eps, float (&out)[A]) {
    float mx = pi[0];
#pragma unroll
    for (int a = 1; a < A; ++a) mx = fmaxf(mx, pi[a]);
    const bool all_small = mx < eps;
    float mp[A];
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const float keep = mul(mask[a], (pi[a] >= eps || all_small) ? 1.f : 0.f);
        mp[a] = mul(keep, pi[a]);
    }
    const float total = sum_lanes<A>(mp);
    float q[A], blocks[A];
#pragma unroll
    for (int a = 0; a < A; ++a) {
        q[a] = dvd(mp[a], total);
        blocks[a] = ceilf(mul(n_disc, q[a]));
    }
    int rank[A];
#pragma unroll
    for (int a = 0; a < A; ++a) {
        int r = 0;
#pragma unroll
        for (int o = 0; o < A; ++o) r += (q[o] > q[a]) || (q[o] == q[a] && o < a);
        rank[a] = r;
    }
    float left = n_disc;
#pragma unroll
    for (int a = 0; a < A; ++a) out[a] = 0.f;
#pragma unroll
    for (int r = 0; r < A; ++r) {
#pragma unroll
        for (int a = 0; a < A; ++a) {
            if (rank[a] == r) {
                const float x = fminf(left, blocks[a]);
                left = sub(left, x);
                out[a] = x;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < A; ++a) out[a] = dvd(out[a], n_disc);
}

template <int A>
__global__ void process_policy_kernel(const float* __restrict__ policy, const float* __restrict__ mask, int64_t n_rows,
                                      float n_disc, float eps, float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
        float pi[A], m[A], o[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            pi[a] = policy[i * A + a];
            m[a] = mask[i * A + a];
        }
        process_policy_row<A>(pi, m, n_disc, eps, o);
#pragma unroll
        for (int a = 0; a < A; ++a) out[i * A + a] = o[a];
    }
}

// The carry of the reverse scan for one player (vtrace.py:58-68).
struct Carry {
    float R, Ru, nv, nvt, IS;
    __device__ __forceinline__ void reset() {
        R = 0.f;
        Ru = 0.f;
        nv = 0.f;
        nvt = 0.f;
        IS = 1.f;
    }
};

struct Scalars {
    float neg_eta, lambda_, c, rho, gamma;
};

// One step of _loop_v_trace (vtrace.py:262-333) for one player, the part that touches the carry.  `own`/`opp`:
// the slot is valid and belongs to this player / to the other one.  Returns the v-trace target `vt` and
// `q_tail` = dR + gamma * IS * nvt - v, the carry-dependent factor of learning_output (vtrace.py:303-309).
__device__ __forceinline__ void vtrace_scan_step(Carry& k, const Scalars& s, bool own, bool opp, float v, float reward,
                                                 float cs, float ent, float& vt, float& q_tail) {
    const float Ru2 = add(add(reward, mul(s.gamma, k.Ru)), ent);
    const float dR = add(reward, mul(s.gamma, k.R));
    const float w = mul(cs, k.IS);
    const float t1 = mul(fminf(w, s.rho), sub(add(Ru2, mul(s.gamma, k.nv)), v));
    const float t2 = mul(mul(mul(s.lambda_, fminf(w, s.c)), s.gamma), sub(k.nvt, k.nv));
    vt = add(add(v, t1), t2);
    q_tail = sub(add(dR, mul(mul(s.gamma, k.IS), k.nvt)), v);
    if (own) {
        k.R = 0.f;
        k.Ru = 0.f;
        k.nv = v;
        k.nvt = vt;
        k.IS = 1.f;
    } else if (opp) {
        k.R = add(ent, mul(cs, dR));
        k.Ru = Ru2;
        k.nv = mul(s.gamma, k.nv);
        k.nvt = mul(s.gamma, k.nvt);
        k.IS = w;
    } else {
        k.reset();
    }
}

// learning_output of an own step (vtrace.py:303-309): v + elp[a] + a_oh[a] / mu(a) * q_tail
template <int A>
__device__ __forceinline__ void learning_output_row(float v, const float (&elp)[A], const float (&a_oh)[A], float inv_mu,
                                                    float q_tail, float (&lo)[A]) {
#pragma unroll
    for (int a = 0; a < A; ++a) lo[a] = add(add(v, elp[a]), mul(mul(a_oh[a], inv_mu), q_tail));
}

template <int A>
__device__ __forceinline__ void vtrace_step(Carry& k, const Scalars& s, bool own, bool opp, float v, float reward,
                                            float cs, float inv_mu, float ent, const float (&elp)[A],
                                            const float (&a_oh)[A], float& vt_out, float (&lo_out)[A]) {
    float vt, q_tail, lo[A];
    vtrace_scan_step(k, s, own, opp, v, reward, cs, ent, vt, q_tail);
    learning_output_row<A>(v, elp, a_oh, inv_mu, q_tail, lo);
#pragma unroll
    for (int a = 0; a < A; ++a) lo_out[a] = own ? lo[a] : 0.f;
    vt_out = own ? vt : 0.f;
}

// net.py:76-80 (forward_batch): masked softmax / log-softmax of one row of logits
template <int A>
__device__ __forceinline__ void policy_from_logits(const float (&logit)[A], const float (&mask)[A], float (&pi)[A],
                                                   float (&log_pi)[A]) {
    float e[A];
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        e[a] = mask[a] != 0.f ? expf(logit[a]) : 0.f;
        sum += e[a];
    }
    const float denom = fmaxf(sum, 1e-12f);
    const float log_sum = logf(sum);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        pi[a] = e[a] / denom;
        log_pi[a] = mask[a] != 0.f ? logit[a] - log_sum : 0.f;
    }
}

// vtrace.py:180-204: sum(a_oh * pi) * valid + (1 - valid)
template <int A>
__device__ __forceinline__ float select_prob(const float (&a_oh)[A], const float (&pi)[A], float valid) {
    float prod[A];
#pragma unroll
    for (int a = 0; a < A; ++a) prod[a] = mul(a_oh[a], pi[a]);
    return add(mul(sum_lanes<A>(prod), valid), sub(1.f, valid));
}

template <int A>
__global__ void vtrace_kernel(const float* __restrict__ v, const float* __restrict__ valid,
                              const int64_t* __restrict__ player_id, const float* __restrict__ mu,
                              const float* __restrict__ merged, const float* __restrict__ merged_log,
                              const float* __restrict__ player_others, const float* __restrict__ actions_oh,
                              const float* __restrict__ reward, int player, Scalars s, int T, int64_t B,
                              float* __restrict__ v_target, int64_t* __restrict__ has_played,
                              float* __restrict__ learning_output) {
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        Carry k;
        k.reset();
        for (int t = T - 1; t >= 0; --t) {
            const int64_t i = (int64_t)t * B + b;
            const float val = valid[i];
            const bool is_valid = val != 0.f;
            const bool mine = player_id[i] == player;
            float a_oh[A], pm[A], pmu[A], L[A], ones[A], prod[A], elp[A], lo[A];
#pragma unroll
            for (int a = 0; a < A; ++a) {
                a_oh[a] = actions_oh[i * A + a];
                pm[a] = merged[i * A + a];
                pmu[a] = mu[i * A + a];
                L[a] = merged_log[i * A + a];
                ones[a] = 1.f;
                prod[a] = mul(pm[a], L[a]);
            }
            const float po = player_others[i];
            const float mu_a = select_prob<A>(a_oh, pmu, val);
            const float cs = dvd(select_prob<A>(a_oh, pm, val), mu_a);
            const float inv_mu = dvd(select_prob<A>(a_oh, ones, val), mu_a);
            const float ent = mul(mul(s.neg_eta, sum_lanes<A>(prod)), po);
#pragma unroll
            for (int a = 0; a < A; ++a) elp[a] = mul(mul(s.neg_eta, L[a]), po);
            float vt;
            vtrace_step<A>(k, s, is_valid && mine, is_valid && !mine, v[i], reward[i], cs, inv_mu, ent, elp, a_oh, vt,
                           lo);
            v_target[i] = vt;
            has_played[i] = (is_valid && mine) ? 1 : 0;   // _has_played reduces to this (vtrace.py:141-177)
#pragma unroll
            for (int a = 0; a < A; ++a) learning_output[i * A + a] = lo[a];
        }
    }
}

__global__ void count_played_kernel(const int64_t* __restrict__ indices, const int64_t* __restrict__ turns, int64_t n,
                                    int32_t* counts) {
    int c0 = 0, c1 = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = indices[i] != 0;
        const int64_t p = turns[i];
        c0 += valid && p == 0;
        c1 += valid && p == 1;
    }
    c0 = warp_sum(c0);
    c1 = warp_sum(c1);
    if ((threadIdx.x & 31) == 0) {
        if (c0) atomicAdd(counts + 0, c0);
        if (c1) atomicAdd(counts + 1, c1);
    }
}

constexpr int kLearnerBlock = 128;
constexpr int kMaxLearnerBlocks = 8192;

template <int A>
__global__ void __launch_bounds__(kLearnerBlock) learner_targets_kernel(rnad_learner_io io, rnad_learner_params p,
                                                                         int T, int64_t B, float* partials) {
    const int32_t* cnt = io.global_counts != nullptr ? io.global_counts : io.counts;
    float N[2] = {1.f, 1.f};                       // unnormalised mode: the caller divides by the step counts later
    if (!io.unnormalised) {
        N[0] = fmaxf((float)cnt[0], 1.f);
        N[1] = fmaxf((float)cnt[1], 1.f);
    }
    if (p.alpha_dev != nullptr) p.alpha = *p.alpha_dev;
    Scalars s;
    s.neg_eta = -p.eta;
    s.lambda_ = p.lambda_;
    s.c = p.c;
    s.rho = p.rho;
    s.gamma = p.gamma;
    const float one_minus_alpha = 1.f - p.alpha;
    const float n_disc = (float)p.n_disc;
    float lv[2] = {0.f, 0.f}, ln[2] = {0.f, 0.f};

    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        Carry k[2];
        k[0].reset();
        k[1].reset();
        for (int t = T - 1; t >= 0; --t) {
            const int64_t i = (int64_t)t * B + b;
            const bool is_valid = io.indices[i] != 0;
            const float val = is_valid ? 1.f : 0.f;
            const int turn = (int)io.turns[i];
            float a_oh[A], mu[A], pi[A], mask[A], L[A], ones[A], prod[A], pt[A], logit[A];
#pragma unroll
            float log_pi[A];
            for (int a = 0; a < A; ++a) {
                a_oh[a] = io.actions_oh[i * A + a];
                mu[a] = io.mu[i * A + a];
                mask[a] = io.masks[i * A + a];
                logit[a] = io.logit[i * A + a];
                ones[a] = 1.f;
            }
            if (io.pi != nullptr) {
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    pi[a] = io.pi[i * A + a];
                    log_pi[a] = io.log_pi[i * A + a];
                }
            } else {
                policy_from_logits<A>(logit, mask, pi, log_pi);      // net.py:76-80
            }
#pragma unroll
            for (int a = 0; a < A; ++a)   // rnad.py:382  log_pi - (alpha*log_pi_reg + (1-alpha)*log_pi_reg_)
                L[a] = sub(log_pi[a], add(mul(p.alpha, io.log_pi_reg[i * A + a]), mul(one_minus_alpha, io.log_pi_reg_[i * A + a])));
            process_policy_row<A>(pi, mask, n_disc, p.eps_threshold, pt);
#pragma unroll
            for (int a = 0; a < A; ++a) prod[a] = mul(pt[a], L[a]);
            const float v_net = io.v_target_net[i];
            const float reward = io.rewards[i];
            const float mu_a = select_prob<A>(a_oh, mu, val);
            const float cs = dvd(select_prob<A>(a_oh, pt, val), mu_a);
            const float inv_mu = dvd(select_prob<A>(a_oh, ones, val), mu_a);
            const float ent_base = mul(s.neg_eta, sum_lanes<A>(prod));

            float d_v = 0.f;
            float d_logit[A];
#pragma unroll
            for (int a = 0; a < A; ++a) d_logit[a] = 0.f;

#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
                const bool own = is_valid && turn == pl;
                const bool opp = is_valid && turn != pl;
                const float po = own ? 1.f : (opp ? -1.f : 0.f);   // _player_others (vtrace.py:70-87)
                const float ent = mul(ent_base, po);
                float elp[A], lo[A];
#pragma unroll
                for (int a = 0; a < A; ++a) elp[a] = mul(mul(s.neg_eta, L[a]), po);
                float vt;
                vtrace_step<A>(k[pl], s, own, opp, v_net, pl == 0 ? reward : -reward, cs, inv_mu, ent, elp, a_oh, vt, lo);
                if (io.v_target[pl]) io.v_target[pl][i] = vt;
                if (io.has_played[pl]) io.has_played[pl][i] = own ? 1 : 0;
                if (io.learning_output[pl]) {
#pragma unroll
                    for (int a = 0; a < A; ++a) io.learning_output[pl][i * A + a] = lo[a];
                }
                if (own) {
                    // critic (vtrace.py:377-393)
                    const float dv = sub(io.v[i], vt);
                    lv[pl] = add(lv[pl], mul(dv, dv));
                    d_v = io.unnormalised ? p.value_weight * 2.f * dv : p.value_weight * 2.f * dv / N[pl];   // (x / 1 == x: no division)
                    // NeuRD (vtrace.py:355-367, 396-431) with importance_sampling_correction == 1
                    float pq[A], ll[A];
#pragma unroll
                    for (int a = 0; a < A; ++a) {
                        pq[a] = mul(pt[a], lo[a]);
                        ll[a] = mul(logit[a], mask[a]);
                    }
                    const float baseline = sum_lanes<A>(pq);
                    const float mean_logit = dvd(sum_lanes<A>(ll), (float)A);
                    float g[A], term[A];
#pragma unroll
                    for (int a = 0; a < A; ++a) {
                        float adv = sub(lo[a], baseline);
                        adv = fminf(fmaxf(adv, -p.neurd_clip), p.neurd_clip);
                        const float lc = sub(logit[a], mean_logit);
                        const float force = add(lc > -p.beta ? fminf(adv, 0.f) : 0.f, lc < p.beta ? fmaxf(adv, 0.f) : 0.f);
                        term[a] = mul(mask[a], mul(lc, force));
                        g[a] = -mask[a] * force / N[pl];
                    }
                    ln[pl] = add(ln[pl], sum_lanes<A>(term));
                    float gsum = 0.f;
#pragma unroll
                    for (int a = 0; a < A; ++a) gsum += g[a];
#pragma unroll
                    for (int a = 0; a < A; ++a) d_logit[a] = p.neurd_weight * (g[a] - mask[a] * gsum / (float)A);
                }
            }
            io.d_v[i] = d_v;
#pragma unroll
            for (int a = 0; a < A; ++a) {
                io.d_logit[i * A + a] = d_logit[a];
                if (io.pi_processed) io.pi_processed[i * A + a] = pt[a];
            }
        }
    }

    // per-block partial sums of the four loss numerators, fixed order -> deterministic totals
    __shared__ float red[4][kLearnerBlock / 32];
    float vals[4] = {lv[0], lv[1], ln[0], ln[1]};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float w = warp_sum(vals[q]);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = w;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float acc = 0.f;
        for (int w = 0; w < kLearnerBlock / 32; ++w) acc += red[threadIdx.x][w];
        partials[blockIdx.x * 4 + threadIdx.x] = acc;
    }
}

// The same computation cut the way the north star words it: everything that is independent per (t, b) slot -
// the reward-transform term, process_policy, the policy ratios, later the NeuRD force and both gradients - runs
// with ONE THREAD PER SLOT (a block is 32 games x T half-moves; a warp is 32 games of one half-move, so every
// (T, B, .) access stays coalesced), and only the 5-float carry of the two players' reverse scans is sequential: one
// thread per (game, player) walks the T slots of its game through shared memory (5 words in, 2 words out per
// slot).  Against one thread per game this puts T times as many loads in flight and fills the SMs at cfg2
// (65,536 games were 0.58 waves of 128-thread blocks).  Same operations in the same order: bit-identical outputs.
constexpr int kTbGames = 32;
constexpr int kTbMaxT = 32;

// kTCap: the block's capacity in half-moves (8, 16 or 32): sizes the shared-memory arrays, so that short trajectories
// are not limited to four blocks per SM by 29 KB of shared memory they do not use.
// `ticket` (device uint32, zero before the launch and zero again after it): the last block to finish adds the
// per-block partial sums in block order (deterministic) - no separate reduction launch.
template <int A, int kTCap>
#ifndef RNAD_K3_BLOCKS_X
#define RNAD_K3_BLOCKS_X 1024
#endif
__global__ void __launch_bounds__(kTbGames* kTCap, RNAD_K3_BLOCKS_X / (kTbGames * kTCap))
    learner_targets_tb_kernel(rnad_learner_io io, rnad_learner_params p, int T, int64_t B, float* partials,
                              unsigned int* ticket) {
    __shared__ float s_v[kTCap][kTbGames], s_reward[kTCap][kTbGames], s_cs[kTCap][kTbGames], s_ent[kTCap][kTbGames];
    __shared__ float s_vt[kTCap][kTbGames], s_q[kTCap][kTbGames];
    __shared__ int s_flags[kTCap][kTbGames];
    __shared__ float red[4][kTCap];
    __shared__ bool s_last;
    const int g = threadIdx.x, t = threadIdx.y;
    const int32_t* cnt = io.global_counts != nullptr ? io.global_counts : io.counts;
    float N[2] = {1.f, 1.f};
    if (!io.unnormalised) {
        N[0] = fmaxf((float)cnt[0], 1.f);
        N[1] = fmaxf((float)cnt[1], 1.f);
    }
    if (p.alpha_dev != nullptr) p.alpha = *p.alpha_dev;
    Scalars s;
    s.neg_eta = -p.eta;
    s.lambda_ = p.lambda_;
    s.c = p.c;
    s.rho = p.rho;
    s.gamma = p.gamma;
    const float one_minus_alpha = 1.f - p.alpha;
    const float n_disc = (float)p.n_disc;
    float lv[2] = {0.f, 0.f}, ln[2] = {0.f, 0.f};

    const int64_t n_tiles = (B + kTbGames - 1) / kTbGames;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b = tile * kTbGames + g;
        const bool active = b < B;
        const int64_t i = (int64_t)t * B + (active ? b : 0);
        // ---- per-slot part 1: everything the scans need
        const bool is_valid = active && io.indices[i] != 0;
        const float val = is_valid ? 1.f : 0.f;
        const int turn = active ? (int)io.turns[i] : 0;
        float a_oh[A], mu[A], pi[A], mask[A], L[A], ones[A], prod[A], pt[A], logit[A], log_pi[A], lreg[A], lreg_[A];
        // every load of the slot first, in one round trip to memory (the arithmetic below depends on the first of them)
#pragma unroll
        for (int a = 0; a < A; ++a) {
            a_oh[a] = io.actions_oh[i * A + a];
            mu[a] = io.mu[i * A + a];
            mask[a] = io.masks[i * A + a];
            logit[a] = io.logit[i * A + a];
            lreg[a] = io.log_pi_reg[i * A + a];
            lreg_[a] = io.log_pi_reg_[i * A + a];
            ones[a] = 1.f;
        }
        const float v_net = io.v_target_net[i];
        const float reward = io.rewards[i];
        const float v_learner = io.v[i];
        if (io.pi != nullptr) {
#pragma unroll
            for (int a = 0; a < A; ++a) {
                pi[a] = io.pi[i * A + a];
                log_pi[a] = io.log_pi[i * A + a];
            }
        } else {
            policy_from_logits<A>(logit, mask, pi, log_pi);      // net.py:76-80
        }
#pragma unroll
        for (int a = 0; a < A; ++a)   // rnad.py:382  log_pi - (alpha*log_pi_reg + (1-alpha)*log_pi_reg_)
            L[a] = sub(log_pi[a], add(mul(p.alpha, lreg[a]), mul(one_minus_alpha, lreg_[a])));
        process_policy_row<A>(pi, mask, n_disc, p.eps_threshold, pt);
#pragma unroll
        for (int a = 0; a < A; ++a) prod[a] = mul(pt[a], L[a]);
        const float mu_a = select_prob<A>(a_oh, mu, val);
        const float cs = dvd(select_prob<A>(a_oh, pt, val), mu_a);
        const float inv_mu = dvd(select_prob<A>(a_oh, ones, val), mu_a);
        const float ent_base = mul(s.neg_eta, sum_lanes<A>(prod));
        s_v[t][g] = v_net;
        s_reward[t][g] = reward;
        s_cs[t][g] = cs;
        s_ent[t][g] = ent_base;
        s_flags[t][g] = (is_valid ? 1 : 0) | (turn << 1);
        __syncthreads();
        // ---- the sequential part: one thread per (game, player) walks the game's slots backwards
        for (int pl = t; pl < 2; pl += blockDim.y) {
            Carry k;
            k.reset();
            for (int tt = T - 1; tt >= 0; --tt) {
                const int f = s_flags[tt][g];
                const bool own = (f & 1) && (f >> 1) == pl, opp = (f & 1) && (f >> 1) != pl;
                const float po = own ? 1.f : (opp ? -1.f : 0.f);   // _player_others (vtrace.py:70-87)
                const float r = s_reward[tt][g];
                float vt, q_tail;
                vtrace_scan_step(k, s, own, opp, s_v[tt][g], pl == 0 ? r : -r, s_cs[tt][g], mul(s_ent[tt][g], po), vt, q_tail);
                if (own) {
                    s_vt[tt][g] = vt;
                    s_q[tt][g] = q_tail;
                }
            }
        }
        __syncthreads();
        // ---- per-slot part 2: learning_output of the slot's owner, both losses and their gradients
        float d_v = 0.f;
        float d_logit[A];
#pragma unroll
        for (int a = 0; a < A; ++a) d_logit[a] = 0.f;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            const bool own = is_valid && turn == pl;
            float vt = 0.f, lo[A];
#pragma unroll
            for (int a = 0; a < A; ++a) lo[a] = 0.f;
            if (own) {
                float elp[A];
#pragma unroll
                for (int a = 0; a < A; ++a) elp[a] = mul(mul(s.neg_eta, L[a]), 1.f);
                vt = s_vt[t][g];
                learning_output_row<A>(v_net, elp, a_oh, inv_mu, s_q[t][g], lo);
                // critic (vtrace.py:377-393)
                const float dv = sub(v_learner, vt);
                lv[pl] = add(lv[pl], mul(dv, dv));
                d_v = p.value_weight * 2.f * dv / N[pl];
                // NeuRD (vtrace.py:355-367, 396-431) with importance_sampling_correction == 1
                float pq[A], ll[A];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    pq[a] = mul(pt[a], lo[a]);
                    ll[a] = mul(logit[a], mask[a]);
                }
                const float baseline = sum_lanes<A>(pq);
                const float mean_logit = dvd(sum_lanes<A>(ll), (float)A);
                float gg[A], term[A];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    float adv = sub(lo[a], baseline);
                    adv = fminf(fmaxf(adv, -p.neurd_clip), p.neurd_clip);
                    const float lc = sub(logit[a], mean_logit);
                    const float force = add(lc > -p.beta ? fminf(adv, 0.f) : 0.f, lc < p.beta ? fmaxf(adv, 0.f) : 0.f);
                    term[a] = mul(mask[a], mul(lc, force));
                    gg[a] = io.unnormalised ? -mask[a] * force : -mask[a] * force / N[pl];
                }
                ln[pl] = add(ln[pl], sum_lanes<A>(term));
                float gsum = 0.f;
#pragma unroll
                for (int a = 0; a < A; ++a) gsum += gg[a];
#pragma unroll
                const float gmean = gsum / (float)A;      // mask is 0 or 1: (mask * gsum) / A == mask * (gsum / A), one division
#pragma unroll
                for (int a = 0; a < A; ++a) d_logit[a] = p.neurd_weight * (gg[a] - mask[a] * gmean);
            }
            if (active) {
                if (io.v_target[pl]) io.v_target[pl][i] = vt;
                if (io.has_played[pl]) io.has_played[pl][i] = own ? 1 : 0;
                if (io.learning_output[pl]) {
#pragma unroll
                    for (int a = 0; a < A; ++a) io.learning_output[pl][i * A + a] = lo[a];
                }
            }
        }
        if (active) {
            io.d_v[i] = d_v;
#pragma unroll
            for (int a = 0; a < A; ++a) {
                io.d_logit[i * A + a] = d_logit[a];
                if (io.pi_processed) io.pi_processed[i * A + a] = pt[a];
            }
        }
        // (no third barrier: the next tile's part 1 writes the scan INPUTS, which nobody reads after the second barrier,
        //  and its scan writes s_vt / s_q only after the next barrier, which every thread reaches after this part 2)
    }

    // per-block partial sums of the four loss numerators, fixed order -> deterministic totals
    float vals[4] = {lv[0], lv[1], ln[0], ln[1]};
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float w = warp_sum(vals[q]);
        if (g == 0) red[q][t] = w;
    }
    __syncthreads();
    if (t == 0 && g < 4) {
        float acc = 0.f;
        for (int w = 0; w < (int)blockDim.y; ++w) acc += red[g][w];
        partials[blockIdx.x * 4 + g] = acc;
        __threadfence();                  // (only the writers: a fence by all 256 threads costs a membar stall each)
    }
    // ---- the last block to arrive sums the partials of all blocks, in block order
    __syncthreads();
    if (t == 0 && g == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int tid = t * kTbGames + g, n_thr = blockDim.x * blockDim.y;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < (int)gridDim.x; i += n_thr) {      // one round of independent 16-byte loads
        const float4 q = __ldcg(reinterpret_cast<const float4*>(partials) + i);
        acc[0] += q.x;
        acc[1] += q.y;
        acc[2] += q.z;
        acc[3] += q.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float w = warp_sum(acc[q]);
        if (g == 0) red[q][t] = w;
    }
    __syncthreads();
    if (tid == 0) {
        float tot[4];
        for (int q = 0; q < 4; ++q) {
            float x = 0.f;
            for (int w = 0; w < (int)blockDim.y; ++w) x += red[q][w];
            tot[q] = x;
        }
        if (io.unnormalised) {            // the four numerators (critic p0, p1, NeuRD p0, p1)
            for (int q = 0; q < 4; ++q) io.loss_sums[q] = tot[q];
        } else {
            io.losses[0] = tot[0] / N[0] + tot[1] / N[1];
            io.losses[1] = -(tot[2] / N[0] + tot[3] / N[1]);
        }
        *ticket = 0u;                     // ready for the next launch
    }
}

__global__ void reduce_losses_kernel(const float* __restrict__ partials, int n_blocks, const int32_t* counts,
                                     const int32_t* global_counts, float* losses, float* loss_sums) {
    __shared__ float red[4][32];
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 4 warps, one per numerator
    float acc = 0.f;
    for (int i = lane; i < n_blocks; i += 32) acc += partials[i * 4 + q];
    red[q][lane] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot[4];
        for (int k = 0; k < 4; ++k) {
            float a = 0.f;
            for (int l = 0; l < 32; ++l) a += red[k][l];
            tot[k] = a;
        }
        if (loss_sums != nullptr) {       // unnormalised mode: the four numerators (critic p0, p1, NeuRD p0, p1)
            for (int k = 0; k < 4; ++k) loss_sums[k] = tot[k];
            return;
        }
        const int32_t* cnt = global_counts != nullptr ? global_counts : counts;
        const float n0 = fmaxf((float)cnt[0], 1.f), n1 = fmaxf((float)cnt[1], 1.f);
        losses[0] = tot[0] / n0 + tot[1] / n1;
        losses[1] = -(tot[2] / n0 + tot[3] / n1);
    }
}

int blocks_for(int64_t n, int block, int cap) {
    int64_t g = (n + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

template <int A>
int launch_process_policy(const float* policy, const float* mask, int64_t n_rows, int n_disc, float eps, float* out,
                          cudaStream_t st) {
    process_policy_kernel<A><<<blocks_for(n_rows, 256, sm_count() * 16), 256, 0, st>>>(policy, mask, n_rows,
                                                                                         (float)n_disc, eps, out);
    RNAD_CHECK_LAUNCH("process_policy_kernel");
    return RNAD_OK;
}

template <int A>
int launch_vtrace(const float* v, const float* valid, const int64_t* player_id, const float* mu, const float* merged,
                  const float* merged_log, const float* player_others, const float* actions_oh, const float* reward,
                  int player, Scalars s, int T, int64_t B, float* v_target, int64_t* has_played, float* learning_output,
                  cudaStream_t st) {
    vtrace_kernel<A><<<blocks_for(B, 128, sm_count() * 16), 128, 0, st>>>(
        v, valid, player_id, mu, merged, merged_log, player_others, actions_oh, reward, player, s, T, B, v_target,
        has_played, learning_output);
    RNAD_CHECK_LAUNCH("vtrace_kernel");
    return RNAD_OK;
}

template <int A>
int launch_learner(const rnad_learner_io& io, const rnad_learner_params& p, int T, int64_t B, float* partials,
                   cudaStream_t st) {
    static const bool per_game = getenv("RNAD_TARGETS_PER_GAME") != nullptr;   // the previous kernel, for A/B runs
    int blocks;
    if (T <= kTbMaxT && !per_game) {
        // one thread per (t, b) slot, a block = one 32-game tile at a time: the blocks of one SM are in different phases
        // (loads / scan / stores) and hide each other's latencies
        const dim3 block(kTbGames, T);
        const int64_t n_tiles = (B + kTbGames - 1) / kTbGames;
        // four resident blocks per SM walk the tiles: the per-block work (parameters, the loss partials' reduction and
        // ticket) is paid once per ~3 tiles instead of once per tile (37.5 -> 33.8 us at cfg2; RNAD_K3_BLOCKS overrides)
        const int64_t resident = 4 * (int64_t)sm_count();
        blocks = (int)(n_tiles < resident ? n_tiles : resident);
        static const char* blocks_env = getenv("RNAD_K3_BLOCKS");
        if (blocks_env != nullptr && atoi(blocks_env) > 0)
            blocks = (int)(n_tiles < atoi(blocks_env) ? n_tiles : (atoi(blocks_env) < kMaxLearnerBlocks ? atoi(blocks_env) : kMaxLearnerBlocks));
        unsigned int* ticket = reinterpret_cast<unsigned int*>(partials + (size_t)kMaxLearnerBlocks * 4);
        if (T <= 8) learner_targets_tb_kernel<A, 8><<<blocks, block, 0, st>>>(io, p, T, B, partials, ticket);
        else if (T <= 16) learner_targets_tb_kernel<A, 16><<<blocks, block, 0, st>>>(io, p, T, B, partials, ticket);
        else learner_targets_tb_kernel<A, 32><<<blocks, block, 0, st>>>(io, p, T, B, partials, ticket);
        RNAD_CHECK_LAUNCH("learner_targets_tb_kernel");
        return RNAD_OK;
    } else {
        blocks = blocks_for(B, kLearnerBlock, kMaxLearnerBlocks);
        learner_targets_kernel<A><<<blocks, kLearnerBlock, 0, st>>>(io, p, T, B, partials);
        RNAD_CHECK_LAUNCH("learner_targets_kernel");
    }
    reduce_losses_kernel<<<1, 128, 0, st>>>(partials, blocks, io.counts, io.global_counts, io.losses,
                                            io.unnormalised ? io.loss_sums : nullptr);
    RNAD_CHECK_LAUNCH("reduce_losses_kernel");
    return RNAD_OK;
}

}  // namespace
}  // namespace rnad

using namespace rnad;

#define RNAD_DISPATCH_A(A_, CALL)                                              \
    switch (A_) {                                                              \
        case 1: { constexpr int kA = 1; return CALL; }                         \
        case 2: { constexpr int kA = 2; return CALL; }                         \
        case 3: { constexpr int kA = 3; return CALL; }                         \
        case 4: { constexpr int kA = 4; return CALL; }                         \
        case 5: { constexpr int kA = 5; return CALL; }                         \
        case 6: { constexpr int kA = 6; return CALL; }                         \
        case 7: { constexpr int kA = 7; return CALL; }                         \
        case 8: { constexpr int kA = 8; return CALL; }                         \
    }

extern "C" {

int rnad_process_policy(const float* policy, const float* mask, int64_t n_rows, int A, int n_disc, float eps_threshold,
                        float* out, void* stream) {
    RNAD_REQUIRE(policy && mask && out, "rnad_process_policy: null pointer");
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS, "rnad_process_policy: %d actions unsupported", A);
    RNAD_REQUIRE(n_rows >= 0 && n_disc >= 1, "rnad_process_policy: bad size");
    if (n_rows == 0) return RNAD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    RNAD_DISPATCH_A(A, launch_process_policy<kA>(policy, mask, n_rows, n_disc, eps_threshold, out, st));
    return RNAD_EINVAL;
}

int rnad_vtrace(const float* v, const float* valid, const int64_t* player_id, const float* acting_policy,
                const float* merged_policy, const float* merged_log_policy, const float* player_others,
                const float* actions_oh, const float* reward, int player, float eta, float lambda_, float c, float rho,
                float gamma, int T, int64_t B, int A, float* v_target, int64_t* has_played, float* learning_output,
                void* stream) {
    RNAD_REQUIRE(v && valid && player_id && acting_policy && merged_policy && merged_log_policy && player_others &&
                     actions_oh && reward && v_target && has_played && learning_output,
                 "rnad_vtrace: null pointer");
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS, "rnad_vtrace: %d actions unsupported", A);
    RNAD_REQUIRE(T >= 0 && B >= 0, "rnad_vtrace: bad shape");
    if (T == 0 || B == 0) return RNAD_OK;
    Scalars s;
    s.neg_eta = -eta;
    s.lambda_ = lambda_;
    s.c = c;
    s.rho = rho;
    s.gamma = gamma;
    cudaStream_t st = (cudaStream_t)stream;
    RNAD_DISPATCH_A(A, launch_vtrace<kA>(v, valid, player_id, acting_policy, merged_policy, merged_log_policy,
                                         player_others, actions_oh, reward, player, s, T, B, v_target, has_played,
                                         learning_output, st));
    return RNAD_EINVAL;
}

int64_t rnad_learner_targets_workspace(int T, int64_t B) {
    (void)T;
    (void)B;
    return (int64_t)kMaxLearnerBlocks * 4 * sizeof(float) + 256;   // per-block partial sums + the last-block ticket
}

int rnad_count_played(const int64_t* indices, const int64_t* turns, int T, int64_t B, int32_t* counts, void* stream) {
    RNAD_REQUIRE(indices && turns && counts, "rnad_count_played: null pointer");
    RNAD_REQUIRE(T >= 0 && B >= 0, "rnad_count_played: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_cuda(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st), "cudaMemsetAsync(counts)");
    if (rc) return rc;
    const int64_t n = (int64_t)T * B;
    if (n == 0) return RNAD_OK;
    count_played_kernel<<<blocks_for(n, 256, sm_count() * 8), 256, 0, st>>>(indices, turns, n, counts);
    RNAD_CHECK_LAUNCH("count_played_kernel");
    return RNAD_OK;
}

int rnad_learner_targets(const rnad_learner_io* io, const rnad_learner_params* p, int T, int64_t B, int A,
                         void* workspace, void* stream) {
    RNAD_REQUIRE(io && p && workspace, "rnad_learner_targets: null pointer");
    RNAD_REQUIRE((io->pi == nullptr) == (io->log_pi == nullptr), "rnad_learner_targets: pi and log_pi must be given together");
    RNAD_REQUIRE(io->indices && io->turns && io->mu && io->actions_oh && io->rewards && io->masks && io->logit &&
                     io->v && io->v_target_net && io->log_pi_reg && io->log_pi_reg_ &&
                     io->d_logit && io->d_v,
                 "rnad_learner_targets: null tensor pointer");
    RNAD_REQUIRE(io->unnormalised ? io->loss_sums != nullptr : (io->losses && io->counts),
                 "rnad_learner_targets: null loss / count pointer");
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS, "rnad_learner_targets: %d actions unsupported", A);
    RNAD_REQUIRE(T >= 1 && B >= 1, "rnad_learner_targets: empty trajectory");
    cudaStream_t st = (cudaStream_t)stream;
    if (!io->unnormalised) {
        int rc = rnad_count_played(io->indices, io->turns, T, B, io->counts, stream);
        if (rc) return rc;
    }
    RNAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "rnad_learner_targets: workspace must be 16-byte aligned");
    RNAD_DISPATCH_A(A, launch_learner<kA>(*io, *p, T, B, (float*)workspace, st));
    return RNAD_EINVAL;
}

}  // extern "C"
