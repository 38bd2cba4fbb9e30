// Per-game device logic shared by the fused rollout kernels (K2): node record
// loads, observation assembly, masked softmax, transition, trajectory record.
// Reference: environment/episode.py:46-125, 194-211; nn/net.py:37-51.
#pragma once

#include "common.cuh"

namespace rnad {

template <int A>
struct Node {
    float ev[A * A];   // expected payoff of the row player, row-major (r, c)
    int rows, cols;    // legal prefix rectangle
};

template <int A>
__device__ __forceinline__ void load_node(const uint32_t* __restrict__ ev_tab, int s, Node<A>& n) {
    constexpr int EVS = ev_stride_of(A);
    const uint4* rec = reinterpret_cast<const uint4*>(ev_tab + (int64_t)s * EVS);
    uint32_t w[EVS];
#pragma unroll
    for (int i = 0; i < EVS / 4; ++i) {
        const uint4 q = __ldg(rec + i);
        w[4 * i + 0] = q.x;
        w[4 * i + 1] = q.y;
        w[4 * i + 2] = q.z;
        w[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int i = 0; i < A * A; ++i) n.ev[i] = __uint_as_float(w[i]);
    n.rows = w[A * A] & 0xff;
    n.cols = (w[A * A] >> 8) & 0xff;
}

// obs flat order [ch][i][j] (episode.py:62-67): the row player sees (ev, legal),
// the column player sees (-ev^T, legal^T).
template <int A>
__device__ __forceinline__ void build_obs(const Node<A>& n, int turn, float (&x)[2 * A * A]) {
#pragma unroll
    for (int i = 0; i < A; ++i)
#pragma unroll
        for (int j = 0; j < A; ++j) {
            const float e_row = n.ev[i * A + j];
            const float e_col = -n.ev[j * A + i];
            x[i * A + j] = turn == 0 ? e_row : e_col;
            const bool lg = turn == 0 ? (i < n.rows && j < n.cols) : (j < n.rows && i < n.cols);
            x[A * A + i * A + j] = lg ? 1.f : 0.f;
        }
}

// net.py:45-46: e = where(mask, exp(logit), 0); policy = e / max(sum e, 1e-12)
template <int A>
__device__ __forceinline__ void masked_softmax(const float (&logit)[A], int n_legal, float (&policy)[A]) {
    float e[A];
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        e[a] = a < n_legal ? expf(logit[a]) : 0.f;
        sum = __fadd_rn(sum, e[a]);
    }
    const float denom = fmaxf(sum, 1e-12f);
#pragma unroll
    for (int a = 0; a < A; ++a) policy[a] = __fdiv_rn(e[a], denom);
}

// episode.py:106-121 with the inverse-CDF chance draw (same rule as sample_icdf).
// The C probabilities, child ids and payoffs of (s, r, c) share one or two
// 32-byte sectors, so the second pair of loads hits L1.
__device__ __forceinline__ void transition(const uint32_t* __restrict__ tr_tab, int A, int C, int s, int r, int c,
                                           float u, int& child, float& reward) {
    const int trs = tr_stride_of(C);
    const uint32_t* ent = tr_tab + ((int64_t)s * A * A + r * A + c) * trs;
    float acc = 0.f;
    int choice = 0;
    bool done = false;
#pragma unroll 4
    for (int k = 0; k < C; ++k) {
        const float pk = __uint_as_float(__ldg(ent + k));
        acc = __fadd_rn(acc, pk);
        const bool positive = pk > 0.f;
        if (!done && positive) choice = k;
        done = done || (positive && (u < acc));
    }
    child = (int)__ldg(ent + C + choice);
    const float v = __uint_as_float(__ldg(ent + 2 * C + choice));
    reward = child == 0 ? v : 0.f;
}

struct TrajPtrs {
    int64_t* indices;
    int64_t* turns;
    float* observations;
    float* policy;
    float* actions;
    float* rewards;
    float* values;
    float* masks;
    float* logits;   // optional (may be null): the policy head's logits, (T,B,A)
    float* returns;  // optional (may be null): every game's payoff for the row player = the sum of its rewards, (B)
};

// the per-(t, b) record except the observation (episode.py:196-211)
template <int A>
__device__ __forceinline__ void write_record(const TrajPtrs& o, int64_t slot, int node, int turn, int n_legal,
                                             const float (&policy)[A], int action, float value, float reward,
                                             const float (&logit)[A]) {
    if (o.logits != nullptr) {
#pragma unroll
        for (int a = 0; a < A; ++a) st_stream(o.logits + slot * A + a, logit[a]);
    }
    st_stream(o.indices + slot, (int64_t)node);
    st_stream(o.turns + slot, (int64_t)turn);
    st_stream(o.values + slot, value);
    st_stream(o.rewards + slot, reward);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        st_stream(o.policy + slot * A + a, policy[a]);
        st_stream(o.actions + slot * A + a, a == action ? 1.f : 0.f);
        st_stream(o.masks + slot * A + a, a < n_legal ? 1.f : 0.f);
    }
}

}  // namespace rnad
