// K1: packed node tables, States.observations, States.step, categorical sampling.
// Reference: environment/tree.py:125-140 (tables), environment/episode.py:46-125.
#include "common.cuh"

namespace rnad {

// ---------------------------------------------------------------- packing

__global__ void pack_ev_kernel(const float* __restrict__ ev, const float* __restrict__ legal, int64_t S, int A,
                               uint32_t* __restrict__ ev_tab, int32_t* bad_flag) {
    const int evs = ev_stride_of(A);
    const int64_t n = S * evs;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = e / evs;
        const int w = (int)(e % evs);
        uint32_t out = 0u;
        if (w < A * A) {
            out = __float_as_uint(ev[s * A * A + w]);
        } else if (w == A * A) {
            const float* lg = legal + s * A * A;
            int rows = 0, cols = 0;
            for (int r = 0; r < A; ++r) rows += lg[r * A] != 0.f;
            for (int c = 0; c < A; ++c) cols += lg[c] != 0.f;
            bool ok = rows >= 1 && cols >= 1;
            for (int r = 0; r < A; ++r)
                for (int c = 0; c < A; ++c) {
                    const float expect = (r < rows && c < cols) ? 1.f : 0.f;
                    ok = ok && (lg[r * A + c] == expect);
                }
            if (!ok) atomicMax(bad_flag, 1);
            out = (uint32_t)rows | ((uint32_t)cols << 8);
        }
        ev_tab[e] = out;
    }
}

__global__ void pack_tr_kernel(const int64_t* __restrict__ index, const float* __restrict__ value,
                               const float* __restrict__ chance, int64_t S, int C, int A,
                               uint32_t* __restrict__ tr_tab, int32_t* bad_flag) {
    const int trs = tr_stride_of(C);
    const int aa = A * A;
    const int64_t n = S * aa;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = e / aa;
        const int rc = (int)(e % aa);
        uint32_t* dst = tr_tab + e * trs;
        for (int k = 0; k < C; ++k) {
            const int64_t src = (s * C + k) * aa + rc;
            const int64_t child = index[src];
            if (child < 0 || child >= S || child > 0x7fffffffLL) atomicMax(bad_flag, 2);
            dst[k] = __float_as_uint(chance[src]);
            dst[C + k] = (uint32_t)(int32_t)child;
            dst[2 * C + k] = __float_as_uint(value[src]);
        }
        for (int k = 3 * C; k < trs; ++k) dst[k] = 0u;
    }
}

// ------------------------------------------------------------ observations

// One thread per output float: writes are fully coalesced, the 2*A*A threads
// of a game read the same one or two 32-byte sectors of its node record.
__global__ void observe_kernel(const uint32_t* __restrict__ ev_tab, int A, const int32_t* __restrict__ idx, int turn,
                               int64_t B, float* __restrict__ obs, float* __restrict__ mask) {
    const int aa = A * A;
    const int evs = ev_stride_of(A);
    const int64_t n = B * 2 * aa;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / (2 * aa);
        const int j = (int)(e % (2 * aa));
        const int ch = j / aa;
        const int i = (j % aa) / A;
        const int jj = j % A;
        const uint32_t* rec = ev_tab + (int64_t)idx[b] * evs;
        // row player sees (r=i, c=jj); the column player sees the transpose
        const int r = turn == 0 ? i : jj;
        const int c = turn == 0 ? jj : i;
        float out;
        if (ch == 0) {
            const float v = __uint_as_float(rec[r * A + c]);
            out = turn == 0 ? v : -v;
        } else {
            const uint32_t dims = rec[aa];
            const int rows = dims & 0xff, cols = (dims >> 8) & 0xff;
            out = (r < rows && c < cols) ? 1.f : 0.f;
            if (mask != nullptr && jj == 0) mask[b * A + i] = out;
        }
        obs[e] = out;
    }
}

// ------------------------------------------------------------------- step

__global__ void step_kernel(const uint32_t* __restrict__ tr_tab, int A, int C, int32_t* __restrict__ idx,
                            const int64_t* __restrict__ row_actions, const int64_t* __restrict__ col_actions,
                            const float* __restrict__ u_chance, uint64_t seed, int t, int64_t game_offset, int64_t B,
                            float* __restrict__ reward, int32_t* alive) {
    const int trs = tr_stride_of(C);
    int live = 0;
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const int s = idx[b];
        const int r = (int)row_actions[b], c = (int)col_actions[b];
        const uint32_t* ent = tr_tab + ((int64_t)s * A * A + r * A + c) * trs;
        const float u = u_chance != nullptr ? u_chance[b] : philox_uniforms(seed, (uint32_t)t, (uint64_t)(game_offset + b)).chance;
        float p[RNAD_MAX_TRANSITIONS];
#pragma unroll
        for (int k = 0; k < RNAD_MAX_TRANSITIONS; ++k) p[k] = k < C ? __uint_as_float(ent[k]) : 0.f;
        const int k = sample_icdf(p, C, u);
        const int child = (int)ent[C + k];
        const float v = __uint_as_float(ent[2 * C + k]);
        idx[b] = child;
        reward[b] = child == 0 ? v : 0.f;
        live += child != 0;
    }
    live = warp_sum(live);
    if ((threadIdx.x & 31) == 0 && live != 0 && alive != nullptr) atomicAdd(alive, live);
}

__global__ void sample_kernel(const float* __restrict__ p, int64_t B, int N, const float* __restrict__ u_in,
                              uint64_t seed, int t, int64_t game_offset, int64_t* __restrict__ out) {
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const float u = u_in != nullptr ? u_in[b] : philox_uniforms(seed, (uint32_t)t, (uint64_t)(game_offset + b)).action;
        float acc = 0.f;
        int choice = 0;
        bool done = false;
        for (int k = 0; k < N; ++k) {
            const float pk = p[b * N + k];
            acc = __fadd_rn(acc, pk);
            const bool positive = pk > 0.f;
            if (!done && positive) choice = k;
            done = done || (positive && (u < acc));
        }
        out[b] = choice;
    }
}

static int grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace rnad

using namespace rnad;

extern "C" {

int rnad_tree_pack(const int64_t* index, const float* value, const float* chance, const float* expected_value,
                   const float* legal, int64_t S, int C, int A, uint32_t* ev_tab, uint32_t* tr_tab,
                   int32_t* bad_flag, void* stream) {
    RNAD_REQUIRE(index && value && chance && expected_value && legal && ev_tab && tr_tab && bad_flag,
                 "rnad_tree_pack: null pointer");
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS, "rnad_tree_pack: max_actions %d outside [1,%d]", A, RNAD_MAX_ACTIONS);
    RNAD_REQUIRE(C >= 1 && C <= RNAD_MAX_TRANSITIONS, "rnad_tree_pack: max_transitions %d outside [1,%d]", C,
                 RNAD_MAX_TRANSITIONS);
    RNAD_REQUIRE(S >= 2 && S <= 0x7fffffffLL, "rnad_tree_pack: node count %lld outside [2, 2^31)", (long long)S);
    cudaStream_t st = (cudaStream_t)stream;
    pack_ev_kernel<<<grid_for(S * ev_stride_of(A), 256), 256, 0, st>>>(expected_value, legal, S, A, ev_tab, bad_flag);
    RNAD_CHECK_LAUNCH("pack_ev_kernel");
    pack_tr_kernel<<<grid_for(S * A * A, 256), 256, 0, st>>>(index, value, chance, S, C, A, tr_tab, bad_flag);
    RNAD_CHECK_LAUNCH("pack_tr_kernel");
    return RNAD_OK;
}

int rnad_observe(const uint32_t* ev_tab, int A, const int32_t* idx, int turn, int64_t B, float* obs, float* mask,
                 void* stream) {
    RNAD_REQUIRE(ev_tab && idx && obs, "rnad_observe: null pointer");
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS, "rnad_observe: max_actions %d unsupported", A);
    RNAD_REQUIRE(turn == 0 || turn == 1, "rnad_observe: turn must be 0 or 1");
    RNAD_REQUIRE(B >= 0, "rnad_observe: negative batch");
    if (B == 0) return RNAD_OK;
    observe_kernel<<<grid_for(B * 2 * A * A, 256), 256, 0, (cudaStream_t)stream>>>(ev_tab, A, idx, turn, B, obs, mask);
    RNAD_CHECK_LAUNCH("observe_kernel");
    return RNAD_OK;
}

int rnad_step(const uint32_t* tr_tab, int A, int C, int32_t* idx, const int64_t* row_actions,
              const int64_t* col_actions, const float* u_chance, uint64_t seed, int t, int64_t game_offset, int64_t B,
              float* reward, int32_t* alive, void* stream) {
    RNAD_REQUIRE(tr_tab && idx && row_actions && col_actions && reward, "rnad_step: null pointer");
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS && C >= 1 && C <= RNAD_MAX_TRANSITIONS, "rnad_step: bad A/C");
    RNAD_REQUIRE(B >= 0, "rnad_step: negative batch");
    if (B == 0) return RNAD_OK;
    step_kernel<<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(tr_tab, A, C, idx, row_actions, col_actions,
                                                                     u_chance, seed, t, game_offset, B, reward, alive);
    RNAD_CHECK_LAUNCH("step_kernel");
    return RNAD_OK;
}

int rnad_sample_categorical(const float* p, int64_t B, int N, const float* u_in, uint64_t seed, int t,
                            int64_t game_offset, int64_t* out, void* stream) {
    RNAD_REQUIRE(p && out, "rnad_sample_categorical: null pointer");
    RNAD_REQUIRE(N >= 1 && B >= 0, "rnad_sample_categorical: bad shape");
    if (B == 0) return RNAD_OK;
    sample_kernel<<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(p, B, N, u_in, seed, t, game_offset, out);
    RNAD_CHECK_LAUNCH("sample_kernel");
    return RNAD_OK;
}

}  // extern "C"
