// K2, tensor-core build (RNAD_PREC_TF32): Episodes.generate fused with
// MLP.forward as a persistent kernel.  A CTA owns a tile of 128 games (thread
// i = game i = TMEM lane i) for all T half-moves:
//
//   gather node record -> observation row (registers)
//     -> A operand tile [128 x KP] tf32 in shared memory (canonical K-major
//        core-matrix layout, bias folded in as a constant-1 input column)
//     -> tcgen05.mma kind::tf32, M=128, N=256, both first layers of the net
//        (value trunk, then policy trunk) against the SMEM-resident weights,
//        fp32 accumulators in TMEM
//     -> tcgen05.ld epilogue: relu + second layers on the CUDA cores
//     -> masked softmax, Philox inverse-CDF action draw, trajectory record,
//        chance draw + child gather on column half-moves.
//
// No host synchronisation inside the rollout (the reference syncs twice per
// half-move, episode.py:96,124).  Two CTAs are resident per SM (256 TMEM
// columns and ~66 KB shared memory each) so one tile's gathers / MMA latency
// hide under the other's epilogue.
// Reference: environment/episode.py:175-230, nn/net.py:37-51.
#include "game.cuh"
#include "rollout.cuh"

namespace rnad {

namespace tc {

constexpr int kTileM = 128;      // games per CTA tile == threads per CTA
constexpr int kHidden = 256;     // width of each trunk
constexpr int kAccCols = 256;    // TMEM columns per CTA (one trunk at a time)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// canonical K-major, no-swizzle operand layout: 8-row x 16-byte core matrices,
// K chunks of a row group adjacent (LBO = 128 B), row groups SBO bytes apart
template <int KP>
__device__ __forceinline__ uint32_t operand_offset(int row, int k) {
    constexpr int SBO = (KP / 4) * 128;
    return (uint32_t)((row >> 3) * SBO + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4);
}

template <int KP>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    constexpr uint64_t LBO = 128 >> 4;
    constexpr uint64_t SBO = ((KP / 4) * 128) >> 4;
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= LBO << 16;                                 // leading-dimension byte offset, bits [16,30)
    d |= SBO << 32;                                 // stride-dimension byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    return d;                                       // base offset 0, layout type 0 = no swizzle
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t kInstrDesc = (1u << 4)            // D format f32
                                | (2u << 7)          // A format tf32
                                | (2u << 10)         // B format tf32
                                | ((uint32_t)(kAccCols >> 3) << 17)
                                | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(kInstrDesc), "r"((uint32_t)accumulate)
        : "memory");
}

__device__ __forceinline__ void mma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    const long long start = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
        if (!done && clock64() - start > 4000000000LL) __trap();   // a lost MMA completion must not hang the device
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int A>
struct Smem {
    static constexpr int KIN = 2 * A * A;
    static constexpr int KP = round_up(KIN + 1, 8);        // + constant-1 column carrying the bias
    static constexpr int kB = 0;                           // weights, both trunks: 512 x KP tf32
    static constexpr int kA = kB + 2 * kHidden * KP * 4;   // observation tile: 128 x KP tf32
    static constexpr int kW2v = kA + kTileM * KP * 4;      // value_fc1.weight, 256 f32
    static constexpr int kW2p = kW2v + kHidden * 4;        // policy_fc1.weight as [j][4] f32
    static constexpr int kB2 = kW2p + kHidden * 16;        // value bias, policy biases
    static constexpr int kBar = kB2 + 32;                  // mbarrier
    static constexpr int kTmem = kBar + 8;                 // TMEM base address
    static constexpr int kBytes = kTmem + 8;
};

template <int A>
__global__ void __launch_bounds__(tc::kTileM) rollout_tc_kernel(RolloutArgs g) {
    using L = Smem<A>;
    constexpr int KIN = L::KIN;
    constexpr int KP = L::KP;
    static_assert(A <= 4, "policy_fc1 rows are staged as one float4 per hidden unit");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const uint32_t mbar = smem_u32(smem + L::kBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kTmem);

    // ---- one-time set-up: TMEM, mbarrier, weights in operand layout
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kAccCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < 2 * kHidden * KP; e += kTileM) {
        const int n = e / KP, k = e % KP;   // hidden unit n of [value trunk | policy trunk], input k
        const float* w = n < kHidden ? g.w.value_fc0_w : g.w.policy_fc0_w;
        const float* bias = n < kHidden ? g.w.value_fc0_b : g.w.policy_fc0_b;
        const int j = n & (kHidden - 1);
        float v = 0.f;
        if (k < KIN) v = w[j * KIN + k];
        else if (k == KIN) v = bias[j];
        *reinterpret_cast<float*>(smem + L::kB + operand_offset<KP>(n, k)) = to_tf32(v);
    }
    for (int j = tid; j < kHidden; j += kTileM) {
        reinterpret_cast<float*>(smem + L::kW2v)[j] = g.w.value_fc1_w[j];
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        p.x = g.w.policy_fc1_w[j];
        if (A > 1) p.y = g.w.policy_fc1_w[1 * kHidden + j];
        if (A > 2) p.z = g.w.policy_fc1_w[2 * kHidden + j];
        if (A > 3) p.w = g.w.policy_fc1_w[3 * kHidden + j];
        reinterpret_cast<float4*>(smem + L::kW2p)[j] = p;
    }
    if (tid == 0) reinterpret_cast<float*>(smem + L::kB2)[0] = g.w.value_fc1_b[0];
    if (tid < A) reinterpret_cast<float*>(smem + L::kB2)[1 + tid] = g.w.policy_fc1_b[tid];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's 32 TMEM lanes
    const float* w2v = reinterpret_cast<const float*>(smem + L::kW2v);
    const float4* w2p = reinterpret_cast<const float4*>(smem + L::kW2p);
    const float* b2 = reinterpret_cast<const float*>(smem + L::kB2);
    const uint32_t a_base = smem_u32(smem + L::kA);
    const uint32_t b_base = smem_u32(smem + L::kB);

    uint32_t phase = 0;
    int last_valid = -1;
    const int64_t num_tiles = (g.B + kTileM - 1) / kTileM;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t b = tile * kTileM + tid;
        const bool active = b < g.B;
        int node = active ? 1 : 0;
        int row_action = 0;
        Node<A> n;
#pragma unroll
        for (int i = 0; i < A * A; ++i) n.ev[i] = 0.f;
        n.rows = n.cols = 1;

        for (int t = 0; t < g.T; ++t) {
            const int turn = t & 1;
            if (turn == 0 && active) load_node<A>(g.ev_tab, node, n);
            if (node != 0) last_valid = max(last_valid, t);
            const int n_legal = turn == 0 ? n.rows : n.cols;
            const int64_t slot = (int64_t)t * g.B + b;
            {
                float x[KP];
                {
                    float obs[KIN];
                    build_obs<A>(n, turn, obs);
#pragma unroll
                    for (int k = 0; k < KIN; ++k) x[k] = obs[k];
                }
                x[KIN] = 1.f;
#pragma unroll
                for (int k = KIN + 1; k < KP; ++k) x[k] = 0.f;
                if (active) {
#pragma unroll
                    for (int k = 0; k < KIN; ++k) st_stream(g.out.observations + slot * KIN + k, x[k]);
                }
                // this game's row of the A operand
#pragma unroll
                for (int q = 0; q < KP / 4; ++q) {
                    const float4 v = make_float4(to_tf32(x[4 * q]), to_tf32(x[4 * q + 1]), to_tf32(x[4 * q + 2]),
                                                 to_tf32(x[4 * q + 3]));
                    *reinterpret_cast<float4*>(smem + L::kA + operand_offset<KP>(tid, 4 * q)) = v;
                }
            }
            fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncthreads();

            float value = b2[0];
            float logit[A];
#pragma unroll
            for (int a = 0; a < A; ++a) logit[a] = b2[1 + a];

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
                    mma_commit(mbar);
                }
                mbar_wait(mbar, phase);
                phase ^= 1u;
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < kAccCols; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tmem_lane + c0, r);
                    tmem_ld_wait();
                    if (pass == 0) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 w = *reinterpret_cast<const float4*>(w2v + c0 + i);
                            value = fmaf(fmaxf(__uint_as_float(r[i + 0]), 0.f), w.x, value);
                            value = fmaf(fmaxf(__uint_as_float(r[i + 1]), 0.f), w.y, value);
                            value = fmaf(fmaxf(__uint_as_float(r[i + 2]), 0.f), w.z, value);
                            value = fmaf(fmaxf(__uint_as_float(r[i + 3]), 0.f), w.w, value);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float h = fmaxf(__uint_as_float(r[i]), 0.f);
                            const float4 w = w2p[c0 + i];
                            logit[0] = fmaf(h, w.x, logit[0]);
                            if constexpr (A > 1) logit[1] = fmaf(h, w.y, logit[1]);
                            if constexpr (A > 2) logit[2] = fmaf(h, w.z, logit[2]);
                            if constexpr (A > 3) logit[3] = fmaf(h, w.w, logit[3]);
                        }
                    }
                }
                tc_fence_before();   // TMEM reads done before the next MMA overwrites the accumulator
                if (pass == 0) __syncthreads();
            }

            float policy[A];
            masked_softmax<A>(logit, n_legal, policy);
            Uniforms2 u;
            if (g.uniforms != nullptr) {
                u.action = active ? g.uniforms[slot * 2 + 0] : 0.f;
                u.chance = active ? g.uniforms[slot * 2 + 1] : 0.f;
            } else {
                u = philox_uniforms(g.seed, (uint32_t)t, (uint64_t)(g.game_offset + b));
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
            if (active) write_record<A>(g.out, slot, node_now, turn, n_legal, policy, action, value, reward);
        }
    }

    last_valid = warp_max(last_valid);
    if ((tid & 31) == 0 && last_valid >= 0) atomicMax(g.t_last, last_valid);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAccCols) : "memory");
    }
}

template <int A>
static int launch(const RolloutArgs& g, cudaStream_t st) {
    using L = Smem<A>;
    // Two CTAs per SM share the 512 TMEM columns; pad the shared-memory request so that a
    // third CTA can never become resident and spin inside tcgen05.alloc.
    size_t smem = L::kBytes;
    const size_t floor_two_per_sm = 227 * 1024 / 3 + 1024;
    if (smem < floor_two_per_sm) smem = floor_two_per_sm;
    if (smem > 227 * 1024) {
        set_error("rnad_rollout(tf32): %zu B of shared memory needed", smem);
        return RNAD_EUNSUPPORTED;
    }
    int rc = check_cuda(cudaFuncSetAttribute(rollout_tc_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(rollout_tc)");
    if (rc) return rc;
    int per_sm = 0;
    rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rollout_tc_kernel<A>, kTileM, smem),
                    "occupancy(rollout_tc)");
    if (rc) return rc;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 512 / kAccCols) per_sm = 512 / kAccCols;
    int64_t blocks = (g.B + kTileM - 1) / kTileM;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    rollout_tc_kernel<A><<<(int)blocks, kTileM, smem, st>>>(g);
    RNAD_CHECK_LAUNCH("rollout_tc_kernel");
    return RNAD_OK;
}

}  // namespace tc

bool rollout_tc_supported(int A, int width) { return width == tc::kHidden && A >= 2 && A <= 4; }

int rollout_tc(const RolloutArgs& g, cudaStream_t st) {
    if (!rollout_tc_supported(g.A, g.w.width)) {
        set_error("rnad_rollout(tf32): needs width == 256 and 2 <= max_actions <= 4 (got width %d, max_actions %d)",
                  g.w.width, g.A);
        return RNAD_EUNSUPPORTED;
    }
    switch (g.A) {
        case 2: return tc::launch<2>(g, st);
        case 3: return tc::launch<3>(g, st);
        case 4: return tc::launch<4>(g, st);
    }
    return RNAD_EINVAL;
}

}  // namespace rnad
