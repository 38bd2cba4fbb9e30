// rnad_learner_forward, pipelined build: the four forward_batch calls of rnad.py:373-380 (five trunk evaluations per
// trajectory row) with BOTH layers on tcgen05, as a persistent warp-specialised kernel - the structure of the
// rollout engine (rollout_tc2.cu) without the game logic.
//
// A tile is 128 trajectory rows (row = TMEM lane).  Its ten *chunks* - 128 hidden units of, in turn, the learner's
// value and policy trunks, the target net's value trunk and the two regularisation nets' policy trunks - stream
// through a ring of three 128-column TMEM slots:
//     MMA1  D[128 x 128]   = obs[128 x KP] (TMEM) x W1_chunk^T (smem)                     kind::f16, fp32 accumulate
//     MMA2  D2[128 x 16]  += relu(D)[128 x 128] (TMEM) x W2_chunk^T (smem)
// Operands are fp16 (11-bit significand like tf32, round to nearest even; observations, weights and activations of this
// net are far inside fp16's range): K = 16 per MMA, relu(hidden) goes back into tensor memory as packed pairs with one
// cvt.rn.relu.f16x2 per two hidden units - the SAME numerics, K order and partial-sum order as the rollout engine
// RNAD_PREC_F16X2 (rollout_tc2.cu), so an on-policy learner step may take the learner net's logits / values from the
// rollout (bit-identical to this kernel's).
// with two 16-column accumulators per tile: D2a = (v, logit[0..A)) of the learner, D2b = (v_target, logit_reg[0..A),
// logit_reg_[0..A)).  Observations and accumulators are double-buffered across tiles.
//   warps 12..15  issue the MMAs (stream item i by warp i % 4; MMA2s only ever ADD into accumulators the output
//                 warps cleared, each half of a trunk into its own columns, so their order across warps is free)
//   warps 4..11   relu epilogue in tensor memory (bias via the constant-1 input column, else one FADD)
//   warps 0..3    one thread per row: observation -> tensor memory (fp16 pairs); one tile later the heads: masked softmax /
//                 log-softmax (net.py:76-80) of the three logit sets, the two values, written time-major.
// Reference: nn/net.py:64-85.  A <= 3: one launch over all five trunks.  A = 4: the five first layers (160 KB) and
// second layers do not fit one CTA's shared memory and target value + two logit sets need nine accumulator columns, so
// the kernel is launched twice - the learner's two trunks, then target / reg / reg_ (v_target and logit_reg in the
// first accumulator, logit_reg_ in the second); the second pass re-reads the observations (128 of ~330 bytes per row).
#include "tc_common.cuh"
#include "tc_pipe.cuh"

namespace rnad {
namespace fwd2 {

using namespace rnad::tc;
using namespace rnad::tcp;

constexpr int kRowWarps = 4, kEpiWarps = 8, kMmaWarps = 4;   // epilogue: 4 lane quadrants x 2 halves of 64 columns
#ifndef RNAD_FWD_EPI_GROUPS
#define RNAD_FWD_EPI_GROUPS 2
#endif
// 1: all eight epilogue warps work on every stream item (a warp = a lane quadrant x 64 columns);
// 2: two groups of four warps take alternate items (a warp = a lane quadrant x all 128 columns, in two passes) - an
//    item's round trip (barrier wake-up, tcgen05.ld, pack, tcgen05.st, arrival) is mostly latency, and two items in
//    the epilogue at a time hide it
constexpr int kEpiGroups = RNAD_FWD_EPI_GROUPS;
static_assert(kEpiGroups == 1 || kEpiGroups == 2, "one or two epilogue groups");
constexpr int kKStep = 16, kEsz = 2;                         // K of one kind::f16 MMA, bytes per operand element
constexpr int kMmaWarp = kRowWarps + kEpiWarps;
constexpr int kThreads = (kMmaWarp + kMmaWarps) * 32;
constexpr int kChunk = 128, kSlots = 3, kAllTrunks = 5;
constexpr int kObsCol = kSlots * kChunk;                              // 2 x 32 columns of observations
constexpr int kD2Col = kObsCol + 64;                                  // 2 x (16 + 16) columns of accumulators

// A launch covers the trunks [T0, T0 + NT) of: 0 learner value, 1 learner policy, 2 target value, 3 reg policy, 4 reg_
// policy.  Which accumulator a trunk adds into, and at which row (= accumulator column) its outputs start:
//   (T0, NT) = (0, 5): D2a = (v, logit), D2b = (v_target, logit_reg, logit_reg_)          [A <= 3]
//   (0, 2):            D2a = (v, logit)                                                   [A = 4, first launch]
//   (2, 3):            D2a = (v_target, logit_reg), D2b = (logit_reg_ from column 1)      [A = 4, second launch]
template <int A, int T0, int NT>
struct Plan {
    static constexpr int kTrunks = NT;
    static constexpr int kItemsPerTile = NT * (kHidden / kChunk);
    __host__ __device__ static constexpr bool second_acc(int trunk) { return NT == 5 ? trunk >= 2 : (T0 == 2 && trunk == 4); }
    __host__ __device__ static constexpr int row0(int trunk) {
        return (trunk == 0 || trunk == 2) ? 0 : (trunk == 4 && NT == 5 ? 1 + A : 1);
    }
    static constexpr int KIN = 2 * A * A;
    static constexpr bool kBiasInK = (KIN % kKStep) != 0;
    static constexpr int KP = round_up(KIN + (kBiasInK ? 1 : 0), kKStep);
    static constexpr int kObsCols = KP * kEsz / 4;                    // tensor-memory columns of one observation row
    static constexpr int kSbo1 = KP * kEsz * 8;                       // bytes between 8-row groups of a [rows x KP] operand
    static constexpr int kSbo2 = kChunk * kEsz * 8;                   // ... of a [16 x 128] second-layer operand
    static constexpr int kTrunkBytes = kHidden * KP * kEsz;
    static constexpr int kW2ChunkBytes = 16 * kChunk * kEsz;          // a [16 x 128] operand: rows 0..7 used by a trunk's first half, 8..15 by its second
    static constexpr int kW1 = 0;                                     // 5 trunks [256 x KP] fp16
    static constexpr int kW2 = kW1 + kTrunks * kTrunkBytes;           // 10 chunks
    static constexpr int kB1 = kW2 + kItemsPerTile * kW2ChunkBytes;   // first-layer biases, 5 x 256 f32
    static constexpr int kB2 = kB1 + kTrunks * kHidden * 4;           // second-layer biases: 4 f32 per trunk
    static constexpr int kImageBytes = kB2 + kTrunks * 16;
    static constexpr int kBar = round_up(kImageBytes, 8);
    static constexpr int kNumBars = 1 + 2 + 2 + 2 + 4 * kSlots;       // image, obs[2], d2 done[2], d2 free[2], d1[6], relu[6]
    static constexpr int kTmem = kBar + 8 * kNumBars;
    static constexpr int kBytes = kTmem + 16;
    static_assert(NT == 5 ? 1 + 2 * A <= 8 : 1 + A <= 8, "the outputs sharing an accumulator must fit its 8 columns");
    static_assert(kBytes <= 227 * 1024, "shared memory plan does not fit");
    static_assert(kObsCols <= 32 && kObsCols % 8 == 0, "observation columns do not fit");
};

// element (row, k) of a K-major, no-swizzle [rows x KP] fp16 operand: 8-row x 16-byte core matrices, the K chunks of a
// row group adjacent (LBO = 128 B), row groups KP * 16 bytes apart
template <int KP>
__host__ __device__ __forceinline__ uint32_t op_off(int row, int k) {
    return (uint32_t)((row >> 3) * (KP * kEsz * 8) + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * kEsz);
}
// two fp32 -> one word of two fp16 (round to nearest even), `lo` in bits 0..15: a tensor-memory column of a 16-bit A
// operand holds two K-adjacent elements, the even one in the low half
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint16_t to_f16(float x) { return (uint16_t)(pack_f16x2(x, 0.f) & 0xffffu); }
__device__ __forceinline__ void mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}

struct Nets {
    rnad_mlp_weights net, target, reg, reg_;
};
struct Out {
    float *logit, *pi, *log_pi, *v, *v_target, *log_pi_reg, *log_pi_reg_;
};

template <int A, int T0, int NT>
__global__ void pack_image_kernel(Nets w, uint8_t* __restrict__ image) {
    using P = Plan<A, T0, NT>;
    constexpr int kTrunks = kAllTrunks;
    const int thread = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    const float* w1[kTrunks] = {w.net.value_fc0_w, w.net.policy_fc0_w, w.target.value_fc0_w, w.reg.policy_fc0_w,
                                w.reg_.policy_fc0_w};
    const float* b1[kTrunks] = {w.net.value_fc0_b, w.net.policy_fc0_b, w.target.value_fc0_b, w.reg.policy_fc0_b,
                                w.reg_.policy_fc0_b};
    const float* w2[kTrunks] = {w.net.value_fc1_w, w.net.policy_fc1_w, w.target.value_fc1_w, w.reg.policy_fc1_w,
                                w.reg_.policy_fc1_w};
    const float* b2[kTrunks] = {w.net.value_fc1_b, w.net.policy_fc1_b, w.target.value_fc1_b, w.reg.policy_fc1_b,
                                w.reg_.policy_fc1_b};
    const int n_out[kTrunks] = {1, A, 1, A, A};
#pragma unroll
    for (int lt = 0; lt < NT; ++lt) {
        const int t = lt, g = T0 + lt;      // local index (position in the image), global trunk
        const int row0_t = P::row0(g);
        for (int e = thread; e < kHidden * P::KP; e += n_threads) {       // first layer (+ bias in K): [256 x KP] fp16
            const int j = e / P::KP, k = e % P::KP;
            float v = 0.f;
            if (k < P::KIN) v = w1[g][j * P::KIN + k];
            else if (P::kBiasInK && k == P::KIN) v = b1[g][j];
            *reinterpret_cast<uint16_t*>(image + P::kW1 + t * P::kTrunkBytes + op_off<P::KP>(j, k)) = to_f16(v);
        }
        for (int j = thread; j < kHidden; j += n_threads) reinterpret_cast<float*>(image + P::kB1)[t * kHidden + j] = b1[g][j];
        // second layer: per 128-unit half a [16 x 128] K-major operand, rows = accumulator columns.  The first half of a
        // trunk uses rows row0 + o, the second rows 8 + row0 + o (zero elsewhere): the two halves leave their partial
        // sums in separate accumulator columns, added by the output warps in a fixed order, so the result does not
        // depend on the order in which the MMA warps got to issue (bit-reproducible).
        for (int e = thread; e < 2 * 16 * kChunk; e += n_threads) {
            const int half = e / (16 * kChunk), r = (e / kChunk) % 16, k = e % kChunk;
            const int o = r - (8 * half + row0_t);
            const float v = (o >= 0 && o < n_out[g]) ? w2[g][o * kHidden + half * kChunk + k] : 0.f;
            *reinterpret_cast<uint16_t*>(image + P::kW2 + (t * 2 + half) * P::kW2ChunkBytes + op_off<kChunk>(r, k)) = to_f16(v);
        }
        if (thread < 4) reinterpret_cast<float*>(image + P::kB2)[t * 4 + thread] = thread < n_out[g] ? b2[g][thread] : 0.f;
    }
}

// net.py:76-80: masked softmax / log-softmax of one row
template <int A>
__device__ __forceinline__ void heads(const float (&logit)[A], uint32_t mask_bits, float (&pi)[A], float (&log_pi)[A]) {
    float e[A];
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        e[a] = (mask_bits >> a) & 1u ? expf(logit[a]) : 0.f;
        sum += e[a];
    }
    const float denom = fmaxf(sum, 1e-12f);
    const float log_sum = logf(sum);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        pi[a] = e[a] / denom;
        log_pi[a] = (mask_bits >> a) & 1u ? logit[a] - log_sum : 0.f;
    }
}

template <int A, int T0, int NT>
__global__ void __launch_bounds__(kThreads, 1) learner_fwd_tc2_kernel(const float* __restrict__ obs, int64_t N,
                                                                       const uint8_t* __restrict__ image, Out out) {
    using P = Plan<A, T0, NT>;
    constexpr int KIN = P::KIN, KP = P::KP, kItemsPerTile = P::kItemsPerTile;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(smem + P::kBar);
    const uint32_t bar_img = bar0;
    auto bar_obs = [&](int b) { return bar0 + 8 + 8 * b; };                  // observations of a tile in tensor memory
    auto bar_d2 = [&](int b) { return bar0 + 24 + 8 * b; };                  // all ten MMA2 groups of a tile complete
    auto bar_free = [&](int b) { return bar0 + 40 + 8 * b; };                // accumulators read and cleared
    auto bar_d1 = [&](int k) { return bar0 + 56 + 8 * k; };                  // per stream item i: barrier i % 6
    auto bar_relu = [&](int k) { return bar0 + 56 + 8 * (2 * kSlots + k); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_obs(b), kRowWarps);
            mbar_init(bar_d2(b), kItemsPerTile);
            mbar_init(bar_free(b), kRowWarps);
        }
        for (int k = 0; k < 2 * kSlots; ++k) {
            mbar_init(bar_d1(k), 1);
            mbar_init(bar_relu(k), kEpiWarps / kEpiGroups);
        }
        mbar_fence_init();
        tma_bulk_load(smem, image, P::kImageBytes, bar_img);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t num_tiles = (N + kTileM - 1) / kTileM;
    int64_t my_tiles = 0;
    if ((int64_t)blockIdx.x < num_tiles) my_tiles = (num_tiles - 1 - blockIdx.x) / gridDim.x + 1;
    const uint32_t n_items = (uint32_t)(my_tiles * kItemsPerTile);

    if (warp >= kMmaWarp) {
        // ------------------------------------------------------------ MMA issuers
        const int w = warp - kMmaWarp;
        mbar_wait_c(bar_img, 0);
        const uint64_t w1_desc = desc_sbo(smem_u32(smem + P::kW1), P::kSbo1);
        const uint64_t w2_desc = desc_sbo(smem_u32(smem + P::kW2), P::kSbo2);
        // D = f32, A and B = f16 (format 0), both K-major
        constexpr uint32_t kIdesc1 = (1u << 4) | ((uint32_t)(kChunk >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        constexpr uint32_t kIdesc2 = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        int seen_obs = -1, seen_free = -1;       // tiles whose observations / cleared accumulators this warp has seen
        // first layers of chunk c of local tile k into `slot`; barrier index of the item
        auto mma1 = [&](int k, int c, int slot, int bar_index) {
#pragma unroll
            for (int s = 0; s < KP / kKStep; ++s)      // (8 tensor-memory columns and 256 operand bytes per K step)
                mma_ts_f16(tmem_base + slot * kChunk, tmem_base + kObsCol + (k & 1) * 32 + s * 8,
                       w1_desc + (uint64_t)(((c >> 1) * P::kTrunkBytes + (c & 1) * (kChunk / 8) * P::kSbo1 + s * 256) >> 4),
                       kIdesc1, s > 0);
            mma_commit(bar_d1(bar_index));
        };
        auto need_obs = [&](int k) {             // (each MMA warp checks for itself: the warps do not order each other)
            if (k != seen_obs) {
                mbar_wait_c(bar_obs(k & 1), (uint32_t)(k >> 1) & 1u);
                tc_fence_after();
                seen_obs = k;
            }
        };
        if (w == 0 && n_items > 0) {             // fill the ring: chunks 0..2 of the first tile
            need_obs(0);
            if (elect_one())
                for (int j = 0; j < kSlots; ++j) mma1(0, j, j, j);
            __syncwarp();
        }
        // item i = w, w + 4, ...; (tile k, chunk c) of item i and of item j = i + 3, all incremental
        int slot = w % kSlots, rb = w % (2 * kSlots);
        uint32_t par = 0;
        int k_i = 0, c_i = w, k_j = 0, c_j = w + kSlots;
        while (c_j >= kItemsPerTile) {           // (a launch over two or three trunks has only four or six items per tile)
            c_j -= kItemsPerTile;
            ++k_j;
        }
#pragma unroll 1
        for (uint32_t i = w; i < n_items; i += kMmaWarps) {
            const bool has_j = i + kSlots < n_items;
            if (has_j) need_obs(k_j);
            if (k_i != seen_free) {              // the accumulators of this tile's buffer were read and cleared (tile k_i - 2)
                mbar_wait_c(bar_free(k_i & 1), ((uint32_t)(k_i >> 1) & 1u) ^ 1u);
                tc_fence_after();
                seen_free = k_i;
            }
            mbar_wait_c(bar_relu(rb), par);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d2 = tmem_base + kD2Col + (k_i & 1) * 32 + (P::second_acc(T0 + (c_i >> 1)) ? 16 : 0);
#pragma unroll
                for (int s = 0; s < kChunk / kKStep; ++s)   // the epilogue warp of a 64-column half packs it into that half's first 32 columns
                    mma_ts_f16(d2, tmem_base + slot * kChunk + (s >> 2) * 64 + (s & 3) * 8,
                               w2_desc + (uint64_t)((c_i * P::kW2ChunkBytes + s * 256) >> 4), kIdesc2, true);
                mma_commit(bar_d2(k_i & 1));
                if (has_j) mma1(k_j, c_j, slot, rb >= kSlots ? rb - kSlots : rb + kSlots);
            }
            __syncwarp();
            slot = slot + 1 == kSlots ? 0 : slot + 1;
            rb += kMmaWarps;
            if (rb >= 2 * kSlots) {
                rb -= 2 * kSlots;
                par ^= 1u;
            }
            c_i += kMmaWarps;
            if (c_i >= kItemsPerTile) {
                c_i -= kItemsPerTile;
                ++k_i;
            }
            c_j += kMmaWarps;
            if (c_j >= kItemsPerTile) {
                c_j -= kItemsPerTile;
                ++k_j;
            }
        }
    } else if (warp >= kRowWarps) {
        // ------------------------------------------------------------ relu epilogue
        const int e = warp - kRowWarps;
        const int quad = e & 3;                                // lane quadrant
        const int group = kEpiGroups == 2 ? e >> 2 : 0;        // which items (kEpiGroups == 2) ...
        const int half0 = kEpiGroups == 2 ? 0 : e >> 2;        // ... or which 64-column half of every item
        constexpr int kCols = 64, kPasses = kEpiGroups;        // a pass = 64 columns = one packed 32-column store
        const uint32_t tmem_quad = tmem_base + ((uint32_t)(quad * 32) << 16);
        const float* b1 = reinterpret_cast<const float*>(smem + P::kB1);
        if (!P::kBiasInK) mbar_wait_c(bar_img, 0);
#pragma unroll 1
        for (uint32_t i = (uint32_t)group; i < n_items; i += kEpiGroups) {
            const uint32_t slot = i % kSlots, rb = i % (2 * kSlots), par = (i / (2 * kSlots)) & 1u, c = i % kItemsPerTile;
            mbar_wait_c(bar_d1(rb), par);
            tc_fence_after();
#pragma unroll
            for (int pass = 0; pass < kPasses; ++pass) {
                const int half = half0 + pass;
                const uint32_t taddr = tmem_quad + slot * kChunk + (uint32_t)(half * kCols);
                uint32_t r[kCols];
#pragma unroll
                for (int q = 0; q < kCols / 32; ++q) tmem_ld32p(taddr + q * 32, r + q * 32);
                tmem_ld_wait();
                if (!P::kBiasInK) {
                    const float4* bias = reinterpret_cast<const float4*>(b1 + c * kChunk + half * kCols);   // chunk c = trunk c/2, half c%2
#pragma unroll
                    for (int k = 0; k < kCols / 4; ++k) {
                        const float4 bb = bias[k];
                        r[4 * k + 0] = __float_as_uint(__uint_as_float(r[4 * k + 0]) + bb.x);
                        r[4 * k + 1] = __float_as_uint(__uint_as_float(r[4 * k + 1]) + bb.y);
                        r[4 * k + 2] = __float_as_uint(__uint_as_float(r[4 * k + 2]) + bb.z);
                        r[4 * k + 3] = __float_as_uint(__uint_as_float(r[4 * k + 3]) + bb.w);
                    }
                }
                uint32_t pk[kCols / 2];          // relu, round to fp16, pack: one instruction per two hidden units
#pragma unroll
                for (int k = 0; k < kCols / 2; ++k) pk[k] = pack_relu_f16x2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
                tmem_st32(taddr, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_relu(rb));
        }
    } else {
        // ------------------------------------------------------------ rows: observations in, heads out
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        mbar_wait_c(bar_img, 0);
        const float* b2 = reinterpret_cast<const float*>(smem + P::kB2);
        {   // both accumulator buffers start from zero
            const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
            for (int q = 0; q < 8; ++q) tmem_st8(tmem_lane + kD2Col + q * 8, zero);
            tmem_st_wait();
        }
        uint32_t mask_prev = 0;
        int64_t row_prev = -1;
        // The observation row of tile k + 1 is fetched while tile k's row is packed and tile k - 1's heads are computed:
        // the row warps' loop used to be [HBM round trip of tile k] -> [heads of tile k - 1], ~2,400 cycles per tile that
        // the six items of a three-trunk launch (~2,300 cycles of tensor-core work) could not hide.
        float xn[KIN];
        bool active_n = false;
        int64_t rown = 0;
        auto fetch = [&](int64_t k) {
            const int64_t tile = (int64_t)blockIdx.x + k * gridDim.x;
            rown = tile * kTileM + tid;
            active_n = rown < N;
            const float2* src = reinterpret_cast<const float2*>(obs + rown * KIN);
#pragma unroll
            for (int q = 0; q < KIN / 2; ++q) {
                const float2 v = active_n ? __ldg(src + q) : make_float2(0.f, 0.f);
                xn[2 * q] = v.x;
                xn[2 * q + 1] = v.y;
            }
        };
        if (my_tiles > 0) fetch(0);
#pragma unroll 1
        for (int64_t k = 0; k <= my_tiles; ++k) {
            uint32_t mask_now = 0;
            int64_t row_now = -1;
            if (k < my_tiles) {
                // ---- this tile's observation row -> tensor memory (its buffer is free: the heads of tile k - 2 are out)
                const int64_t row = rown;
                const bool active = active_n;
                float x[KIN];
#pragma unroll
                for (int q = 0; q < KIN; ++q) x[q] = xn[q];
                if (k + 1 < my_tiles) fetch(k + 1);
#pragma unroll
                for (int a = 0; a < A; ++a) mask_now |= (x[A * A + a * A] != 0.f ? 1u : 0u) << a;   // obs[:, 1, :, 0]
                row_now = active ? row : -1;
                auto obs_k = [&](int kk) {   // element kk of the padded observation row (the constant 1 carries the bias)
                    return kk < KIN ? x[kk < KIN ? kk : 0] : ((P::kBiasInK && kk == KIN) ? 1.f : 0.f);
                };
#pragma unroll
                for (int q = 0; q < P::kObsCols / 8; ++q) {
                    uint32_t v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = pack_f16x2(obs_k(2 * (8 * q + u)), obs_k(2 * (8 * q + u) + 1));
                    tmem_st8(tmem_lane + kObsCol + (k & 1) * 32 + 8 * q, v);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_obs(k & 1));
            }
            if (k > 0) {
                // ---- heads of the previous tile
                const int kp = (int)(k - 1);
                mbar_wait_c(bar_d2(kp & 1), (uint32_t)(kp >> 1) & 1u);
                tc_fence_after();
                uint32_t da[8], db[8];
                {
                    uint32_t da_hi[8], db_hi[8];     // partial sums of the second half of every trunk
                    tmem_ld8(tmem_lane + kD2Col + (kp & 1) * 32, da);
                    tmem_ld8(tmem_lane + kD2Col + (kp & 1) * 32 + 8, da_hi);
                    tmem_ld8(tmem_lane + kD2Col + (kp & 1) * 32 + 16, db);
                    tmem_ld8(tmem_lane + kD2Col + (kp & 1) * 32 + 24, db_hi);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        da[q] = __float_as_uint(__uint_as_float(da[q]) + __uint_as_float(da_hi[q]));
                        db[q] = __float_as_uint(__uint_as_float(db[q]) + __uint_as_float(db_hi[q]));
                    }
                }
                {
                    const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
                    for (int q = 0; q < 4; ++q) tmem_st8(tmem_lane + kD2Col + (kp & 1) * 32 + 8 * q, zero);
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_free(kp & 1));
                if (row_prev >= 0) {
                    const int64_t row = row_prev;
                    float lg[A], pi[A], lp[A];
                    if (T0 == 0) {               // the learner's trunks: local trunks 0 (value), 1 (policy), both in D2a
                        out.v[row] = __uint_as_float(da[0]) + b2[0];
#pragma unroll
                        for (int a = 0; a < A; ++a) lg[a] = __uint_as_float(da[1 + a]) + b2[1 * 4 + a];
                        heads<A>(lg, mask_prev, pi, lp);
#pragma unroll
                        for (int a = 0; a < A; ++a) {
                            out.logit[row * A + a] = lg[a];
                            out.pi[row * A + a] = pi[a];
                            out.log_pi[row * A + a] = lp[a];
                        }
                    }
                    if (T0 + NT == kAllTrunks) { // target value, reg and reg_ policies: local trunks (2, 3, 4) - T0
                        constexpr int lt = 2 - T0;
                        // (0, 5): all three in D2b at columns 0, 1.., 1 + A..;  (2, 3): v_target, logit_reg in D2a, logit_reg_ in D2b
                        const uint32_t* first = NT == 5 ? db : da;
                        out.v_target[row] = __uint_as_float(first[0]) + b2[lt * 4];
#pragma unroll
                        for (int a = 0; a < A; ++a) lg[a] = __uint_as_float(first[1 + a]) + b2[(lt + 1) * 4 + a];
                        heads<A>(lg, mask_prev, pi, lp);
#pragma unroll
                        for (int a = 0; a < A; ++a) out.log_pi_reg[row * A + a] = lp[a];
#pragma unroll
                        for (int a = 0; a < A; ++a) lg[a] = __uint_as_float(db[P::row0(4) + a]) + b2[(lt + 2) * 4 + a];
                        heads<A>(lg, mask_prev, pi, lp);
#pragma unroll
                        for (int a = 0; a < A; ++a) out.log_pi_reg_[row * A + a] = lp[a];
                    }
                }
            }
            mask_prev = mask_now;
            row_prev = row_now;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc<512>(tmem_base);
}

// mode: 0 = pack the weight image, then run; 1 = the image is already in the workspace (rnad_learner_pack); 2 = pack only
template <int A, int T0, int NT>
int launch_part(const float* obs, int64_t N, const Nets& nets, const Out& out, uint8_t* workspace, cudaStream_t st, int mode) {
    using P = Plan<A, T0, NT>;
    if (mode != 1) {
        pack_image_kernel<A, T0, NT><<<64, 256, 0, st>>>(nets, workspace);
        RNAD_CHECK_LAUNCH("learner fwd2 pack_image_kernel");
        if (mode == 2) return RNAD_OK;
    }
    const size_t smem = P::kBytes > 116 * 1024 ? P::kBytes : 116 * 1024;   // one CTA per SM (all 512 TMEM columns)
    int rc = check_cuda(cudaFuncSetAttribute(learner_fwd_tc2_kernel<A, T0, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(learner_fwd_tc2)");
    if (rc) return rc;
    int64_t blocks = (N + kTileM - 1) / kTileM;
    if (blocks > sm_count()) blocks = sm_count();
    learner_fwd_tc2_kernel<A, T0, NT><<<(int)blocks, kThreads, smem, st>>>(obs, N, workspace, out);
    RNAD_CHECK_LAUNCH("learner_fwd_tc2_kernel");
    return RNAD_OK;
}

template <int A>
int launch(const float* obs, int64_t N, const Nets& nets, const Out& out, uint8_t* workspace, cudaStream_t st, int mode,
           bool others_only) {
    if (others_only) return launch_part<A, 2, 3>(obs, N, nets, out, workspace, st, mode);
    if constexpr (A <= 3) {
        return launch_part<A, 0, 5>(obs, N, nets, out, workspace, st, mode);
    } else {
        // two launches, each with its own image (one behind the other in the workspace)
        int rc = launch_part<A, 0, 2>(obs, N, nets, out, workspace, st, mode);
        if (rc) return rc;
        return launch_part<A, 2, 3>(obs, N, nets, out, workspace + round_up(Plan<A, 0, 2>::kImageBytes, 256), st, mode);
    }
}

}  // namespace fwd2

bool learner_forward_tc2_supported(int A, int width) { return width == tc::kHidden && A >= 2 && A <= 4; }

int64_t learner_forward_tc2_image_bytes(int A) {
    switch (A) {
        case 2: return fwd2::Plan<2, 0, 5>::kImageBytes;
        case 3: return fwd2::Plan<3, 0, 5>::kImageBytes;
        case 4: return round_up(fwd2::Plan<4, 0, 2>::kImageBytes, 256) + fwd2::Plan<4, 2, 3>::kImageBytes;
    }
    return 0;
}

int learner_forward_tc2(const float* obs, int64_t N, int A, const rnad_mlp_weights* net, const rnad_mlp_weights* target,
                        const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_, const rnad_learner_fwd_out* out,
                        void* workspace, cudaStream_t st, int mode, bool others_only) {
    fwd2::Nets nets{*net, *target, *reg, *reg_};
    fwd2::Out o{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (out != nullptr) o = fwd2::Out{out->logit, out->pi, out->log_pi, out->v, out->v_target, out->log_pi_reg, out->log_pi_reg_};
    switch (A) {
        case 2: return fwd2::launch<2>(obs, N, nets, o, (uint8_t*)workspace, st, mode, others_only);
        case 3: return fwd2::launch<3>(obs, N, nets, o, (uint8_t*)workspace, st, mode, others_only);
        case 4: return fwd2::launch<4>(obs, N, nets, o, (uint8_t*)workspace, st, mode, others_only);
    }
    return RNAD_EUNSUPPORTED;
}

}  // namespace rnad
