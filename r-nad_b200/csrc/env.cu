// K1: packed node tables, States.observations, States.step, categorical sampling.
// Reference: environment/tree.py:125-140 (tables), environment/episode.py:46-125.
#include <cstdlib>

#include "game.cuh"

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

// One thread per game: three 16-byte loads fetch its 48-byte node record (A = 3; one or two sectors), the observation
// is assembled in registers, and a warp's 32 observations - one contiguous block of the (B, 2, A, A) tensor - are
// staged in shared memory and stored as coalesced 16-byte words (the same scheme as the fused rollout's store_obs).
// A compile-time: no integer divisions.  [The previous kernel - one thread per output float, runtime A - issued 78
// instructions per game for 18 floats and ran at 0.15 of the HBM roofline on the 14.9 M-node cfg3 tree.]
//
// kStageNodes > 0 is the north star's "node tables TMA-staged into shared memory", as far as it can go for a
// data-dependent gather: the records of nodes 0 .. kStageNodes-1 (the root and, in level order, the top levels) are
// fetched once per CTA with ONE bulk copy (cp.async.bulk, completion on an mbarrier) and games standing on those
// nodes read shared memory.  Measured against the plain gather in profiles/r02_k1_standalone.md.
constexpr int kObsWarps = 4;
__host__ __device__ constexpr int stage_nodes_of(int A) {   // up to 1024 records, at most 96 KB of shared memory
    return 96 * 1024 / (ev_stride_of(A) * 4) < 1024 ? 96 * 1024 / (ev_stride_of(A) * 4) : 1024;
}

template <int A, int kStageNodes>
__global__ void __launch_bounds__(kObsWarps * 32) observe_kernel(const uint32_t* __restrict__ ev_tab,
                                                                  const int32_t* __restrict__ idx, int turn, int64_t B,
                                                                  int64_t S, float* __restrict__ obs,
                                                                  float* __restrict__ mask) {
    constexpr int KIN = 2 * A * A, EVS = ev_stride_of(A);
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* s_stage = reinterpret_cast<float*>(smem_raw) + (threadIdx.x >> 5) * 32 * KIN;
    const uint32_t* s_tab = reinterpret_cast<const uint32_t*>(smem_raw + kObsWarps * 32 * KIN * 4);
    const int lane = threadIdx.x & 31;
    int64_t staged = 0;
    if (kStageNodes > 0) {
        staged = S < kStageNodes ? S : kStageNodes;
        uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kObsWarps * 32 * KIN * 4 + kStageNodes * EVS * 4);
        const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t bytes = (uint32_t)(staged * EVS * 4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(s_tab)),
                         "l"(ev_tab), "r"(bytes), "r"(bar_addr)
                         : "memory");
        }
        __syncthreads();
        uint32_t done = 0;
        while (!done)
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(bar_addr)
                : "memory");
    }
    const int64_t n_groups = (B + 31) / 32;
    for (int64_t grp = blockIdx.x * (int64_t)kObsWarps + (threadIdx.x >> 5); grp < n_groups; grp += (int64_t)gridDim.x * kObsWarps) {
        const int64_t base = grp * 32, b = base + lane;
        const int rows = (int)(B - base < 32 ? B - base : 32);
        Node<A> n;
        if (b < B) {
            const int s = idx[b];
            if (kStageNodes > 0 && s < staged) {
                const uint32_t* w = s_tab + s * EVS;
#pragma unroll
                for (int i = 0; i < A * A; ++i) n.ev[i] = __uint_as_float(w[i]);
                n.rows = w[A * A] & 0xff;
                n.cols = (w[A * A] >> 8) & 0xff;
            } else {
                load_node<A>(ev_tab, s, n);
            }
        } else {
#pragma unroll
            for (int i = 0; i < A * A; ++i) n.ev[i] = 0.f;
            n.rows = n.cols = 0;
        }
        float x[KIN];
        build_obs<A>(n, turn, x);
        __syncwarp();                       // the previous group's copy out of the staging rows is done
#pragma unroll
        for (int k = 0; k < KIN; k += 2) *reinterpret_cast<float2*>(s_stage + lane * KIN + k) = make_float2(x[k], x[k + 1]);
        __syncwarp();
        float* dst = obs + base * KIN;      // 32 * KIN * 4 bytes per group: 16-byte aligned whenever `obs` is
        const int n_float = rows * KIN;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            for (int i = lane * 4; i < n_float; i += 128)    // (rows * KIN is even; a ragged tail of 2 floats below)
                if (i + 4 <= n_float) __stcs(reinterpret_cast<float4*>(dst + i), *reinterpret_cast<const float4*>(s_stage + i));
                else *reinterpret_cast<float2*>(dst + i) = *reinterpret_cast<const float2*>(s_stage + i);
        } else {
            for (int i = lane; i < n_float; i += 32) dst[i] = s_stage[i];
        }
        if (mask != nullptr)                // the mover's legal actions: obs[:, 1, :, 0]
            for (int i = lane; i < rows * A; i += 32) mask[base * A + i] = s_stage[(i / A) * KIN + A * A + (i % A) * A];
    }
}

// ------------------------------------------------------------------- step

// One thread per game; the whole transition entry (chance row, child ids, payoffs: 32 bytes at C = 2, 48 at C = 3,
// 16-byte aligned) arrives with independent 16-byte loads, C compile-time.
template <int C>
__global__ void __launch_bounds__(128) step_kernel(const uint32_t* __restrict__ tr_tab, int A, int32_t* __restrict__ idx,
                                                   const int64_t* __restrict__ row_actions,
                                                   const int64_t* __restrict__ col_actions,
                                                   const float* __restrict__ u_chance, uint64_t seed, int t,
                                                   int64_t game_offset, int64_t B, float* __restrict__ reward,
                                                   int32_t* alive) {
    constexpr int TRS = tr_stride_of(C);
    int live = 0;
    for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const int s = idx[b];
        const int r = (int)row_actions[b], c = (int)col_actions[b];
        const uint4* ent = reinterpret_cast<const uint4*>(tr_tab + ((int64_t)s * A * A + r * A + c) * TRS);
        uint32_t w[TRS];
#pragma unroll
        for (int q = 0; q < TRS / 4; ++q) {
            const uint4 v = __ldg(ent + q);
            w[4 * q + 0] = v.x;
            w[4 * q + 1] = v.y;
            w[4 * q + 2] = v.z;
            w[4 * q + 3] = v.w;
        }
        const float u = u_chance != nullptr ? u_chance[b] : philox_uniforms(seed, (uint32_t)t, (uint64_t)(game_offset + b)).chance;
        // inverse CDF of the chance row (same rule as sample_icdf), carrying child and payoff along
        float acc = 0.f;
        uint32_t child = w[C], val = w[2 * C];
        bool done = false;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const float pk = __uint_as_float(w[k]);
            acc = __fadd_rn(acc, pk);
            const bool positive = pk > 0.f;
            if (!done && positive) {
                child = w[C + k];
                val = w[2 * C + k];
            }
            done = done || (positive && (u < acc));
        }
        idx[b] = (int)child;
        reward[b] = child == 0 ? __uint_as_float(val) : 0.f;
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

template <int A>
static int launch_observe(const uint32_t* ev_tab, const int32_t* idx, int turn, int64_t B, int64_t S, float* obs,
                          float* mask, cudaStream_t st) {
    constexpr int KIN = 2 * A * A, EVS = ev_stride_of(A);
    static const bool stage = getenv("RNAD_K1_STAGE") != nullptr;      // the staged variant, for the A/B measurement
    const int64_t groups = (B + 31) / 32;
    int64_t blocks = (groups + kObsWarps - 1) / kObsWarps;
    if (stage) {
        // persistent: every CTA pays for one bulk copy of the staged records
        constexpr int kStageNodesMax = stage_nodes_of(A);
        constexpr size_t smem = kObsWarps * 32 * KIN * 4 + kStageNodesMax * EVS * 4 + 16;
        static_assert(smem <= 227 * 1024, "staging does not fit");
        if (blocks > 2 * sm_count()) blocks = 2 * sm_count();
        int rc = check_cuda(cudaFuncSetAttribute(observe_kernel<A, kStageNodesMax>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(observe_kernel)");
        if (rc) return rc;
        observe_kernel<A, kStageNodesMax><<<(int)blocks, kObsWarps * 32, smem, st>>>(ev_tab, idx, turn, B, S, obs, mask);
    } else {
        constexpr size_t smem = kObsWarps * 32 * KIN * 4;
        if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
        if (smem > 48 * 1024) {
            int rc = check_cuda(cudaFuncSetAttribute(observe_kernel<A, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "cudaFuncSetAttribute(observe_kernel)");
            if (rc) return rc;
        }
        observe_kernel<A, 0><<<(int)blocks, kObsWarps * 32, smem, st>>>(ev_tab, idx, turn, B, S, obs, mask);
    }
    RNAD_CHECK_LAUNCH("observe_kernel");
    return RNAD_OK;
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

int rnad_observe(const uint32_t* ev_tab, int A, int64_t S, const int32_t* idx, int turn, int64_t B, float* obs,
                 float* mask, void* stream) {
    RNAD_REQUIRE(ev_tab && idx && obs, "rnad_observe: null pointer");
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS, "rnad_observe: max_actions %d unsupported", A);
    RNAD_REQUIRE(turn == 0 || turn == 1, "rnad_observe: turn must be 0 or 1");
    RNAD_REQUIRE(B >= 0, "rnad_observe: negative batch");
    if (B == 0) return RNAD_OK;
    RNAD_REQUIRE(S >= 2, "rnad_observe: the table has %lld nodes", (long long)S);
    switch (A) {
        case 1: return launch_observe<1>(ev_tab, idx, turn, B, S, obs, mask, (cudaStream_t)stream);
        case 2: return launch_observe<2>(ev_tab, idx, turn, B, S, obs, mask, (cudaStream_t)stream);
        case 3: return launch_observe<3>(ev_tab, idx, turn, B, S, obs, mask, (cudaStream_t)stream);
        case 4: return launch_observe<4>(ev_tab, idx, turn, B, S, obs, mask, (cudaStream_t)stream);
        case 5: return launch_observe<5>(ev_tab, idx, turn, B, S, obs, mask, (cudaStream_t)stream);
        case 6: return launch_observe<6>(ev_tab, idx, turn, B, S, obs, mask, (cudaStream_t)stream);
        case 7: return launch_observe<7>(ev_tab, idx, turn, B, S, obs, mask, (cudaStream_t)stream);
        case 8: return launch_observe<8>(ev_tab, idx, turn, B, S, obs, mask, (cudaStream_t)stream);
    }
    return RNAD_EINVAL;
}

int rnad_step(const uint32_t* tr_tab, int A, int C, int32_t* idx, const int64_t* row_actions,
              const int64_t* col_actions, const float* u_chance, uint64_t seed, int t, int64_t game_offset, int64_t B,
              float* reward, int32_t* alive, void* stream) {
    RNAD_REQUIRE(tr_tab && idx && row_actions && col_actions && reward, "rnad_step: null pointer");
    RNAD_REQUIRE(A >= 1 && A <= RNAD_MAX_ACTIONS && C >= 1 && C <= RNAD_MAX_TRANSITIONS, "rnad_step: bad A/C");
    RNAD_REQUIRE(B >= 0, "rnad_step: negative batch");
    if (B == 0) return RNAD_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define RNAD_STEP_CASE(c)                                                                                              \
    case c:                                                                                                            \
        step_kernel<c><<<grid_for(B, 128), 128, 0, st>>>(tr_tab, A, idx, row_actions, col_actions, u_chance, seed, t,  \
                                                         game_offset, B, reward, alive);                              \
        break;
    switch (C) {
        RNAD_STEP_CASE(1) RNAD_STEP_CASE(2) RNAD_STEP_CASE(3) RNAD_STEP_CASE(4)
        RNAD_STEP_CASE(5) RNAD_STEP_CASE(6) RNAD_STEP_CASE(7) RNAD_STEP_CASE(8)
    }
#undef RNAD_STEP_CASE
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
