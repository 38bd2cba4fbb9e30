// K2, tensor-core build with BOTH layers of the net on tcgen05 (RNAD_PREC_TF32X2):
// Episodes.generate fused with MLP.forward as a persistent, warp-specialised kernel.
//
// A CTA owns a tile of 128 games for all T half-moves; game g of the tile is TMEM
// lane g.  Nine warps:
//
//   warp 8 (one lane)   issues every tcgen05.mma.  Per half-move the 512 hidden units
//                       (value trunk | policy trunk) are processed as 8 chunks of 64
//                       through a ring of three 64-column TMEM slots:
//                         MMA1(c)  D[128 x 64]  = obs[128 x KP] (smem) x W1_c^T (smem)   kind::tf32
//                         MMA2(c)  D2[128 x 16] += relu(D)[128 x 64] (TMEM) x W2_c^T (smem)
//                       i.e. the second layers read their A operand straight from tensor
//                       memory; column 0 of D2 is the value, columns 1..A the logits.
//   warps 0..7          epilogue of MMA1: tcgen05.ld 32 columns, relu (the bias rides in K
//                       as a constant-1 input column where K has padding, else one FADD),
//                       tcgen05.st back in place, arrive on the chunk's mbarrier.  This
//                       is all the CUDA cores do per hidden unit: one FMNMX.
//   warps 0..3          head: read D2, masked softmax, Philox inverse-CDF action draw,
//                       chance draw + child gather on column half-moves, trajectory
//                       record, next observation -> A-operand tile (tf32) + fp32 staging.
//   warps 4..7          copy the staged observations of the tile (one contiguous block of
//                       the (T,B,2,A,A) tensor) to HBM with coalesced 16-byte stores.
//
// Everything is ordered by mbarriers (no CTA-wide barrier inside the rollout); two CTAs
// per SM (256 TMEM columns each) overlap one tile's head with the other's MMAs.
// The weights arrive as ONE TMA bulk copy of an image a pre-kernel lays out in operand
// order.  Reference: environment/episode.py:175-230, nn/net.py:37-51.
#include "game.cuh"
#include "rollout.cuh"
#include "tc_common.cuh"

namespace rnad {
namespace tc2 {

using namespace rnad::tc;

constexpr int kHeadWarps = 4;              // warps 0..3: one thread per game
#ifndef RNAD_TC2_EPI_WARPS
#define RNAD_TC2_EPI_WARPS 8
#endif
constexpr int kEpiWarps = RNAD_TC2_EPI_WARPS;   // warps 4..: relu epilogue of the first layers (4 or 8)
constexpr int kMmaWarp = kHeadWarps + kEpiWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;
constexpr int kChunk = 64;                 // hidden units per pipeline stage
constexpr int kChunks = 2 * kHidden / kChunk;
constexpr int kSlots = 3;
constexpr int kD2Col = kSlots * kChunk;    // 16 columns of second-layer accumulators
constexpr int kObsCol = kD2Col + 16;       // up to 32 columns: the observations, A operand of the first layers
constexpr int kTmemCols = 256;
constexpr int kN2 = 16;                    // N of the second-layer MMA (smallest legal at M = 128)
constexpr int kK2 = 2 * kHidden;           // its K: value trunk | policy trunk

template <int A>
struct Plan {
    static constexpr int KIN = 2 * A * A;
    static constexpr bool kBiasInK = (KIN % 8) != 0;
    static constexpr int KP = round_up(KIN + (kBiasInK ? 1 : 0), 8);
    static constexpr bool kStage = A < 4;                            // A = 4: a row is one 128-byte line, stored directly
    static constexpr int kSbo1 = (KP / 4) * 128;                     // bytes between 8-row groups of a [rows x KP] operand
    static constexpr int kW1 = 0;                                    // [512 x KP] tf32
    static constexpr int kW2 = kW1 + 2 * kHidden * KP * 4;           // rows 0..7 of [16 x 512] tf32 (rows 8..15 alias what follows)
    static constexpr int kB1 = kW2 + 8 * kK2 * 4;                    // first-layer biases, 512 f32 (used when !kBiasInK)
    static constexpr int kB2 = kB1 + 2 * kHidden * 4;                // value bias, policy biases (8 f32)
    static constexpr int kImageBytes = kB2 + 32;
    static constexpr int kObs = kImageBytes;                         // fp32 observation staging [128 x KIN]
    static constexpr int kBar = kObs + (kStage ? kTileM * KIN * 4 : 0);
    static constexpr int kNumBars = 2 + 2 * kSlots + 1;              // image, A-ready, d1[3], relu[3], d2
    static constexpr int kTmem = kBar + 8 * kNumBars;
    static constexpr int kBytes = kTmem + 16;
    // the second 8-row group of the W2 operand is read 16 KB behind the first: it must stay inside the allocation
    static constexpr int kMinBytes = kW2 + 2 * 8 * kK2 * 4;
    static_assert(kImageBytes % 16 == 0 && kObs % 16 == 0 && kBar % 8 == 0 && (32 * KIN * 4) % 16 == 0, "alignment");
    static_assert(KP <= 32 && kObsCol + KP <= kTmemCols, "observation columns do not fit");
};

__host__ __device__ constexpr uint32_t instr_desc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}

// A operand from tensor memory (lane = row, one 32-bit column per K element)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}

__device__ __forceinline__ uint64_t desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

#ifdef RNAD_TRACE
// development aid: cycle stamps of CTA 0's first tile.  [role][half-move][event]
__device__ long long g_trace[3][16][24];
#define TR(role, t, ev) do { if (blockIdx.x == 0 && tile == blockIdx.x && (t) < 16) g_trace[role][t][ev] = clock64(); } while (0)
#else
#define TR(role, t, ev) do { } while (0)
#endif

template <int A>
__global__ void pack_weights_kernel(rnad_mlp_weights w, uint8_t* __restrict__ image) {
    using P = Plan<A>;
    constexpr int KIN = P::KIN, KP = P::KP;
    const int thread = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int e = thread; e < 2 * kHidden * KP; e += stride) {
        const int n = e / KP, k = e % KP;   // hidden unit n of [value trunk | policy trunk], input k
        const int j = n & (kHidden - 1);
        float v = 0.f;
        if (k < KIN) v = (n < kHidden ? w.value_fc0_w : w.policy_fc0_w)[j * KIN + k];
        else if (P::kBiasInK && k == KIN) v = (n < kHidden ? w.value_fc0_b : w.policy_fc0_b)[j];
        *reinterpret_cast<float*>(image + P::kW1 + operand_offset<KP>(n, k)) = to_tf32(v);
    }
    // second layers as ONE [8 x 512] operand: row 0 = value_fc1 over the value trunk's hidden units,
    // rows 1..A = policy_fc1 over the policy trunk's, zero elsewhere
    for (int e = thread; e < 8 * kK2; e += stride) {
        const int n = e / kK2, k = e % kK2;
        float v = 0.f;
        if (n == 0 && k < kHidden) v = w.value_fc1_w[k];
        else if (n >= 1 && n <= A && k >= kHidden) v = w.policy_fc1_w[(n - 1) * kHidden + (k - kHidden)];
        *reinterpret_cast<float*>(image + P::kW2 + operand_offset<kK2>(n, k)) = to_tf32(v);
    }
    for (int j = thread; j < kHidden; j += stride) {
        reinterpret_cast<float*>(image + P::kB1)[j] = w.value_fc0_b[j];
        reinterpret_cast<float*>(image + P::kB1)[kHidden + j] = w.policy_fc0_b[j];
    }
    if (thread < 8) {
        float v = 0.f;
        if (thread == 0) v = w.value_fc1_b[0];
        else if (thread <= A) v = w.policy_fc1_b[thread - 1];
        reinterpret_cast<float*>(image + P::kB2)[thread] = v;
    }
}

// masked softmax with the fast exp2 / reciprocal units (net.py:45-46: e = where(mask, exp(logit), 0); e / max(sum e, 1e-12)).
// A few ulp from the exact formula; the recorded policy is the one every later decision uses.
template <int A>
__device__ __forceinline__ void masked_softmax_fast(const float (&logit)[A], int n_legal, float (&policy)[A]) {
    float e[A];
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        e[a] = a < n_legal ? exp2f(logit[a] * 1.4426950408889634f) : 0.f;
        sum += e[a];
    }
    const float inv = __frcp_rn(fmaxf(sum, 1e-12f));
#pragma unroll
    for (int a = 0; a < A; ++a) policy[a] = e[a] * inv;
}

template <int A>
__global__ void __launch_bounds__(kThreads, 2) rollout_tc2_kernel(RolloutArgs g, const uint8_t* __restrict__ image) {
    using P = Plan<A>;
    constexpr int KIN = P::KIN, KP = P::KP;
    static_assert(A <= 4, "value + logits must fit the 8 useful columns of the second-layer accumulator");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const uint32_t bar0 = smem_u32(smem + P::kBar);
    const uint32_t bar_img = bar0, bar_a = bar0 + 8;
    const uint32_t bar_d2 = bar0 + 16 + 16 * kSlots;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);
    // bar_d1[s] = bar0 + 16 + 8 s, bar_relu[s] = bar0 + 16 + 8 (kSlots + s)

    if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        mbar_init(bar_a, kHeadWarps);                               // one arrival per head warp
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(bar0 + 16 + 8 * s, 1);                        // MMA1 of the slot complete (tcgen05.commit)
            mbar_init(bar0 + 16 + 8 * (kSlots + s), kEpiWarps);     // relu written back, one arrival per epilogue warp
        }
        mbar_init(bar_d2, 1);
        mbar_fence_init();
        tma_bulk_load(smem, image, P::kImageBytes, bar_img);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t num_tiles = (g.B + kTileM - 1) / kTileM;

    if (warp == kMmaWarp) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop converged and one elected lane issues: the descriptors then live in
        // uniform registers (an `if (lane == 0)` branch makes ptxas wrap every UTCHMMA in a waterfall loop).
        mbar_wait(bar_img, 0);
        const uint64_t w1_desc = desc_sbo(smem_u32(smem + P::kW1), P::kSbo1);
        const uint64_t w2_desc = desc_sbo(smem_u32(smem + P::kW2), 8 * kK2 * 4);
        constexpr uint32_t kIdesc1 = instr_desc(kChunk), kIdesc2 = instr_desc(kN2);
        uint32_t ph_a = 0, ph_relu = 0;   // bit s of ph_relu = parity of slot s
        auto mma1 = [&](int c) {          // first layers, hidden units [64c, 64c+64): A = observations in tensor memory
            const int slot = c % kSlots;
#pragma unroll
            for (int s = 0; s < KP / 8; ++s)
                mma_ts(tmem_base + slot * kChunk, tmem_base + kObsCol + s * 8,
                       w1_desc + (uint64_t)((c * (kChunk / 8) * P::kSbo1 + s * 256) >> 4), kIdesc1, s > 0);
            mma_commit(bar0 + 16 + 8 * slot);
        };
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int t = 0; t < g.T; ++t) {
                TR(0, t, 0);
                mbar_wait(bar_a, ph_a);
                ph_a ^= 1u;
                tc_fence_after();
                TR(0, t, 1);
                if (elect_one()) {
                    mma1(0);
                    mma1(1);
                    mma1(2);
                }
                __syncwarp();
                TR(0, t, 2);
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    const int slot = c % kSlots;
                    mbar_wait(bar0 + 16 + 8 * (kSlots + slot), (ph_relu >> slot) & 1u);
                    ph_relu ^= 1u << slot;
                    tc_fence_after();
                    TR(0, t, 3 + 2 * c);
                    if (elect_one()) {
#pragma unroll
                        for (int s = 0; s < kChunk / 8; ++s)   // second layers: A = relu(hidden) in tensor memory
                            mma_ts(tmem_base + kD2Col, tmem_base + slot * kChunk + s * 8,
                                   w2_desc + (uint64_t)(((c * (kChunk / 8) + s) * 256) >> 4), kIdesc2, (c | s) != 0);
                        if (c + kSlots < kChunks) mma1(c + kSlots);
                        if (c == kChunks - 1) mma_commit(bar_d2);
                    }
                    __syncwarp();
                    TR(0, t, 4 + 2 * c);
                }
            }
        }
    } else if (warp >= kHeadWarps) {
        // ------------------------------------------------------------ epilogue of the first layers
        const int e = warp - kHeadWarps;
        constexpr int kPart = kEpiWarps / 4;                 // warps sharing a lane quadrant split the chunk's columns
        constexpr int kCols = kChunk / kPart;
        const uint32_t tmem_mine = tmem_base + ((uint32_t)((e & 3) * 32) << 16) + (uint32_t)((e >> 2) * kCols);
        const float* b1 = reinterpret_cast<const float*>(smem + P::kB1) + (e >> 2) * kCols;
        if (!P::kBiasInK) mbar_wait(bar_img, 0);
        uint32_t ph_d1 = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int t = 0; t < g.T; ++t) {
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    const int slot = c % kSlots;
                    if (c == 0 && (tid & 31) == 0 && e == 0) TR(2, t, 0);
                    mbar_wait(bar0 + 16 + 8 * slot, (ph_d1 >> slot) & 1u);
                    ph_d1 ^= 1u << slot;
                    tc_fence_after();
                    if ((tid & 31) == 0 && e == 0) TR(2, t, 1 + 2 * c);
                    const uint32_t taddr = tmem_mine + slot * kChunk;
                    uint32_t r[kCols];
#pragma unroll
                    for (int q = 0; q < kCols / 32; ++q) tmem_ld32(taddr + q * 32, *reinterpret_cast<uint32_t(*)[32]>(r + q * 32));
                    tmem_ld_wait();
                    if (!P::kBiasInK) {
                        const float4* bias = reinterpret_cast<const float4*>(b1 + c * kChunk);
#pragma unroll
                        for (int i = 0; i < kCols / 4; ++i) {
                            const float4 bb = bias[i];
                            r[4 * i + 0] = __float_as_uint(__uint_as_float(r[4 * i + 0]) + bb.x);
                            r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + bb.y);
                            r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + bb.z);
                            r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + bb.w);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < kCols; ++i) r[i] = __float_as_uint(fmaxf(__uint_as_float(r[i]), 0.f));
#pragma unroll
                    for (int q = 0; q < kCols / 32; ++q) tmem_st32(taddr + q * 32, *reinterpret_cast<uint32_t(*)[32]>(r + q * 32));
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(bar0 + 16 + 8 * (kSlots + slot));
                    if ((tid & 31) == 0 && e == 0) TR(2, t, 2 + 2 * c);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ head: one thread per game
        const int lane_g = tid;                               // game of the tile == TMEM lane
        const int lane = tid & 31;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        float* s_obs = reinterpret_cast<float*>(smem + P::kObs) + warp * 32 * KIN;   // this warp's 32 rows
        mbar_wait(bar_img, 0);
        const float b2v = reinterpret_cast<const float*>(smem + P::kB2)[0];
        float b2p[A];
#pragma unroll
        for (int a = 0; a < A; ++a) b2p[a] = reinterpret_cast<const float*>(smem + P::kB2)[1 + a];
        constexpr int EVS = ev_stride_of(A);
        const int trs = tr_stride_of(g.C);

        uint32_t ph_d2 = 0;
        int last_valid = -1;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t tile_base = tile * kTileM;
            const int64_t b = tile_base + lane_g;
            const bool active = b < g.B;
            int node = active ? 1 : 0;
            int row_action = 0;
            Node<A> n;
#pragma unroll
            for (int i = 0; i < A * A; ++i) n.ev[i] = 0.f;
            n.rows = n.cols = 1;
            float x[KIN];

            // observation of half-move t -> tensor memory (tf32 A operand of the first layers); critical path
            auto publish_obs = [&](int t) {
                const int turn = t & 1;
                if (turn == 0 && active) load_node<A>(g.ev_tab, node, n);
                build_obs<A>(n, turn, x);
#pragma unroll
                for (int q = 0; q < KP / 8; ++q) {
                    uint32_t v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = 8 * q + u;
                        v[u] = __float_as_uint(k < KIN ? to_tf32(x[k < KIN ? k : 0]) : ((P::kBiasInK && k == KIN) ? 1.f : 0.f));
                    }
                    tmem_st8(tmem_lane + kObsCol + 8 * q, v);
                }
                tmem_st_wait();
                tc_fence_before();       // orders these stores and this thread's tcgen05.ld of D2 before the next MMAs
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_a);
            };
            // the same observation in fp32 -> trajectory; off the critical path.  A warp's 32 rows are one
            // contiguous block of the (T,B,2,A,A) tensor: staged in shared memory, stored as coalesced 16-byte words.
            auto store_obs = [&](int t) {
                const int64_t row0 = (int64_t)t * g.B + tile_base + warp * 32;
                float* dst = g.out.observations + row0 * KIN;
                const int rows = (int)max((int64_t)0, min((int64_t)32, g.B - (tile_base + warp * 32)));
                if (P::kStage && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    __syncwarp();        // the previous half-move's copy is done
#pragma unroll
                    for (int k = 0; k < KIN; k += 2)
                        *reinterpret_cast<float2*>(s_obs + lane * KIN + k) = make_float2(x[k], x[k + 1]);
                    __syncwarp();
                    const int n_float = rows * KIN;
                    for (int i = lane * 4; i < n_float; i += 128) {
                        if (i + 4 <= n_float) {
                            __stcs(reinterpret_cast<float4*>(dst + i), *reinterpret_cast<const float4*>(s_obs + i));
                        } else {
                            for (int k = i; k < n_float; ++k) __stcs(dst + k, s_obs[k]);
                        }
                    }
                } else if (active) {
                    float2* d2 = reinterpret_cast<float2*>(g.out.observations + ((int64_t)t * g.B + b) * KIN);
#pragma unroll
                    for (int k = 0; k < KIN / 2; ++k) __stcs(d2 + k, make_float2(x[2 * k], x[2 * k + 1]));
                }
            };
            auto draw = [&](int t) {
                Uniforms2 u;
                if (g.uniforms != nullptr) {
                    const int64_t slot_tb = (int64_t)t * g.B + b;
                    u.action = active ? __ldg(g.uniforms + slot_tb * 2 + 0) : 0.f;
                    u.chance = active ? __ldg(g.uniforms + slot_tb * 2 + 1) : 0.f;
                } else {
                    u = philox_uniforms(g.seed, (uint32_t)t, (uint64_t)(g.game_offset + b));
                }
                return u;
            };

            publish_obs(0);
            store_obs(0);
            Uniforms2 u = draw(0);

            for (int t = 0; t < g.T; ++t) {
                const int turn = t & 1;
                if (node != 0) last_valid = max(last_valid, t);
                if (tid == 0) TR(1, t, 0);
                mbar_wait(bar_d2, ph_d2);
                ph_d2 ^= 1u;
                tc_fence_after();
                if (tid == 0) TR(1, t, 1);
                uint32_t d2[8];
                tmem_ld8(tmem_lane + kD2Col, d2);
                tmem_ld_wait();
                const float value = __uint_as_float(d2[0]) + b2v;
                float logit[A];
#pragma unroll
                for (int a = 0; a < A; ++a) logit[a] = __uint_as_float(d2[1 + a]) + b2p[a];
                const int n_legal = turn == 0 ? n.rows : n.cols;
                float policy[A];
                masked_softmax_fast<A>(logit, n_legal, policy);
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
                if (tid == 0) TR(1, t, 2);
                if (t + 1 < g.T) publish_obs(t + 1);
                if (tid == 0) TR(1, t, 3);

                // ---- off the critical path: the tensor core and the epilogue warps are busy with half-move t + 1
                if (active) write_record<A>(g.out, (int64_t)t * g.B + b, node_now, turn, n_legal, policy, action, value, reward);
                if (t + 1 < g.T) {
                    store_obs(t + 1);
                    u = draw(t + 1);
                    if (turn == 0 && active) {
                        // warm L1 for the next half-move's gathers: the transition entries of (node, row_action, *)
                        // and the node records of their children
                        const uint32_t* ent = g.tr_tab + ((int64_t)node * A * A + row_action * A) * trs;
                        for (int c = 0; c < A; ++c)
                            for (int k = 0; k < g.C; ++k) {
                                const int child = (int)__ldg(ent + c * trs + g.C + k);
                                const uint32_t* rec = g.ev_tab + (int64_t)child * EVS;
                                prefetch_l1(rec);
                                if (EVS * 4 > 32) prefetch_l1(rec + EVS - 1);
                            }
                    }
                }
                if (tid == 0) TR(1, t, 4);
            }
        }
        last_valid = warp_max(last_valid);
        if (lane == 0 && last_valid >= 0) atomicMax(g.t_last, last_valid);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc<kTmemCols>(tmem_base);
}

template <int A>
static int launch(const RolloutArgs& g, uint8_t* workspace, cudaStream_t st) {
    using P = Plan<A>;
    pack_weights_kernel<A><<<32, 256, 0, st>>>(g.w, workspace);
    RNAD_CHECK_LAUNCH("pack_weights_kernel");
    // Two CTAs per SM share the 512 TMEM columns; pad the shared-memory request so that a
    // third CTA can never become resident and spin inside tcgen05.alloc.
    size_t smem = P::kBytes > P::kMinBytes ? P::kBytes : P::kMinBytes;
    const size_t floor_two_per_sm = 227 * 1024 / 3 + 1024;
    if (smem < floor_two_per_sm) smem = floor_two_per_sm;
    if (smem > 227 * 1024) {
        set_error("rnad_rollout(tf32x2): %zu B of shared memory needed", smem);
        return RNAD_EUNSUPPORTED;
    }
    const int per_sm = 2 * (smem + 1024) <= 228 * 1024 ? 2 : 1;
    int rc = check_cuda(cudaFuncSetAttribute(rollout_tc2_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(rollout_tc2, smem)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(rollout_tc2_kernel<A>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared),
                    "cudaFuncSetAttribute(rollout_tc2, carveout)");
    if (rc) return rc;
    int64_t blocks = (g.B + kTileM - 1) / kTileM;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    rollout_tc2_kernel<A><<<(int)blocks, kThreads, smem, st>>>(g, workspace);
    RNAD_CHECK_LAUNCH("rollout_tc2_kernel");
    return RNAD_OK;
}

}  // namespace tc2

#ifdef RNAD_TRACE
extern "C" __attribute__((visibility("default"))) int rnad_debug_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, tc2::g_trace, sizeof(tc2::g_trace));
}
#endif

int64_t rollout_tc2_workspace_bytes(int A) {
    switch (A) {
        case 2: return tc2::Plan<2>::kImageBytes;
        case 3: return tc2::Plan<3>::kImageBytes;
        case 4: return tc2::Plan<4>::kImageBytes;
    }
    return 0;
}

int rollout_tc2(const RolloutArgs& g, void* workspace, cudaStream_t st) {
    if (!rollout_tc_supported(g.A, g.w.width)) {
        set_error("rnad_rollout(tf32x2): needs width == 256 and 2 <= max_actions <= 4 (got width %d, max_actions %d)",
                  g.w.width, g.A);
        return RNAD_EUNSUPPORTED;
    }
    if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0) {
        set_error("rnad_rollout(tf32x2): needs a 16-byte aligned workspace of rnad_rollout_workspace_bytes()");
        return RNAD_EINVAL;
    }
    switch (g.A) {
        case 2: return tc2::launch<2>(g, (uint8_t*)workspace, st);
        case 3: return tc2::launch<3>(g, (uint8_t*)workspace, st);
        case 4: return tc2::launch<4>(g, (uint8_t*)workspace, st);
    }
    return RNAD_EINVAL;
}

}  // namespace rnad
