// K2, tensor-core build with BOTH layers of the net on tcgen05 (RNAD_PREC_F16X2, the default: fp16 operands,
// kind::f16; RNAD_PREC_TF32X2: the same kernel with tf32 operands, kind::tf32 - see Plan<A, F16>):
// Episodes.generate fused with MLP.forward as a persistent, warp-specialised kernel.
//
// One CTA per SM owns all 512 TMEM columns and keeps TWO tiles of 128 games ("sides")
// in flight, half a step out of phase: while the tensor core and the epilogue warps
// work on one side's half-move, the other side's head warps sample actions, advance
// the games and publish the next observations.  Game g of a tile is TMEM lane g.
//
//   warps 16..19            issue the tcgen05.mma of ONE stream of chunks
//                           (side 0, t) c0..c3, (side 1, t) c0..c3, (side 0, t+1) ...
//                           (one elected lane each; warp 16 + c issues the chunks with index c: several issuers
//                           keep the tensor core's queue fed).  A chunk is 128 hidden units of
//                           [policy trunk | value trunk]:
//                             MMA1  D[128 x 128]  = obs[128 x KP] (TMEM) x W1_c^T (smem)      3 slots
//                             MMA2  D2[128 x 16] += relu(D)[128 x 128] (TMEM) x W2_c^T (smem)
//                           both with the A operand in tensor memory.  The policy trunk's chunks come
//                           first and have their own accumulator (logits in columns 1..A): the heads
//                           start on them while the value trunk (value in column 0 of its own
//                           accumulator) is still in the ring.
//   warps 8..15             epilogue of MMA1: tcgen05.ld, relu (the bias rides in K as a
//                           constant-1 input column where K has padding, else one FADD),
//                           tcgen05.st back in place, arrive on the slot's mbarrier.  One
//                           FMNMX per hidden unit is all the CUDA cores do for the net.
//   warps 0..3 / 4..7       heads of side 0 / 1, one thread per game: read D2, masked
//                           softmax, Philox inverse-CDF action draw, next observation ->
//                           tensor memory (operand format).  Off the critical path (the tensor core is
//                           busy with the other side): trajectory record, fp32 observations
//                           (staged per warp, stored as coalesced 16-byte words), the next
//                           uniforms, and on row half-moves the chance draw + child gather
//                           for EVERY column action the opponent may pick, parked in shared
//                           memory - so that no global load sits between a column action and
//                           the next observation.
//
// Everything is ordered by mbarriers; no CTA-wide barrier inside the rollout.  Every
// CTA lays the weights out in operand order in its own shared memory at start-up (the
// whole launch is this one kernel).  Reference: environment/episode.py:175-230, nn/net.py:37-51.
#include <atomic>
#include "game.cuh"
#include "rollout.cuh"
#include "tc_common.cuh"
#include "tc_pipe.cuh"

namespace rnad {
namespace tc2 {

using namespace rnad::tc;
using namespace rnad::tcp;

constexpr int kSides = 2;                        // tiles in flight per CTA
constexpr int kHeadWarps = 4 * kSides;           // warps 0..7: one thread per game
#ifndef RNAD_TC2_EPI_WARPS
#define RNAD_TC2_EPI_WARPS 8
#endif
// relu epilogue warps: 4 lane quadrants x 2 column halves, and with 16 warps x 2 item parities (each warp then takes
// every other stream item, and registers are re-allocated between the roles).  Measured on cfg2: 8 warps 0.066 ms,
// 16 warps 0.072 ms per rollout - more warps lengthen the heads' critical path more than they shorten the epilogue's.
constexpr int kEpiWarps = RNAD_TC2_EPI_WARPS;
#ifndef RNAD_TC2_EPI_WIDE
#define RNAD_TC2_EPI_WIDE 0
#endif
// RNAD_TC2_EPI_WIDE (8 warps only): a warp takes all 128 columns of its lane quadrant, of every other stream item
constexpr bool kEpiWide = RNAD_TC2_EPI_WIDE != 0 && kEpiWarps == 8;
constexpr int kEpiStride = kEpiWide ? 2 : kEpiWarps / 8;   // an epilogue warp takes every kEpiStride-th stream item
constexpr int kEpiPerItem = kEpiWide ? 4 : 8;              // warps working on one stream item
static_assert(kEpiWarps == 8 || kEpiWarps == 16, "8 or 16 epilogue warps");
constexpr int kMmaWarp = kHeadWarps + kEpiWarps;   // first of the MMA warps
constexpr int kMmaWarps = 4;                       // stream item i is issued by MMA warp i % 4, i.e. one warp per chunk index
constexpr int kThreads = (kMmaWarp + kMmaWarps) * 32;
constexpr int kChunk = 128;                      // hidden units per pipeline stage
constexpr int kChunks = 2 * kHidden / kChunk;    // per side and half-move
constexpr int kSlots = 3;
constexpr int kSideCol = kSlots * kChunk;        // per side 64 columns: 16 of second-layer accumulators, 32 of observations
constexpr int kTmemCols = 512;
constexpr int kN2 = 16;                          // N of the second-layer MMA (smallest legal at M = 128)
constexpr int kK2 = 2 * kHidden;                 // its K: value trunk | policy trunk

// per side: the policy trunk's second-layer accumulator (logits in columns 1..A), the value trunk's (value in column
// 0), then the observations
__host__ __device__ constexpr int d2p_col(int side) { return kSideCol + 64 * side; }
__host__ __device__ constexpr int d2v_col(int side) { return kSideCol + 64 * side + 16; }
__host__ __device__ constexpr int obs_col(int side) { return kSideCol + 64 * side + 32; }
// Chunk order within a half-move: the POLICY trunk first (c = 0, 1), the value trunk after it (c = 2, 3).  The next
// observation only needs the sampled action, so the heads work on it while the value chunks are still in the ring.
__host__ __device__ constexpr int chunk_hidden(int c) { return c < 2 ? kHidden + c * kChunk : (c - 2) * kChunk; }

// F16 = false: kind::tf32 (RNAD_PREC_TF32X2); F16 = true: kind::f16 with fp16 operands (RNAD_PREC_F16X2) - the same 11-bit
// significand as tf32 (payoffs, weights and activations of this net sit far inside fp16's range), but K = 16 per MMA
// instead of 8: half the tensor-core dispatches for both layers, and relu(hidden) goes back into tensor memory as packed
// pairs (half the tcgen05.st bytes, one cvt.rn.relu.f16x2 per two hidden units).
template <int A, bool F16>
struct Plan {
    static constexpr int KIN = 2 * A * A;
    static constexpr int kEsz = F16 ? 2 : 4;                         // bytes per operand element
    static constexpr int kKStep = 32 / kEsz;                         // K of one MMA
    static constexpr bool kBiasInK = (KIN % kKStep) != 0;
    static constexpr int KP = round_up(KIN + (kBiasInK ? 1 : 0), kKStep);
    static constexpr int kObsCols = KP * kEsz / 4;                   // tensor-memory columns of one observation row
    static constexpr int kCandWords = A * A + 3;                     // child, reward, ev[A*A], rows|cols
    static constexpr int kSbo1 = KP * kEsz * 8;                      // bytes between 8-row groups of a [rows x KP] operand
    static constexpr int kSbo2 = kK2 * kEsz * 8;                     // ... of the [16 x 512] second-layer operand
    static constexpr int kW1 = 0;                                    // [512 x KP] tf32 / fp16
    static constexpr int kW2 = kW1 + 2 * kHidden * KP * kEsz;        // [16 x 512]: rows 0..7 serve the first chunk of a trunk, 8..15 the second
    static constexpr int kB1 = kW2 + kN2 * kK2 * kEsz;               // first-layer biases, 512 f32 (used when !kBiasInK)
    static constexpr int kB2 = kB1 + 2 * kHidden * 4;                // value bias, policy biases (8 f32)
    static constexpr int kImageBytes = kB2 + 32;
    static constexpr int kObs = kImageBytes;                         // fp32 observation staging, [side][128 x KIN]
    static constexpr int kCand = kObs + kSides * kTileM * KIN * 4;   // transition candidates, [side][A][kCandWords][128]
    static constexpr int kBar = kCand + kSides * A * kCandWords * kTileM * 4;
    static constexpr int kNumBars = 1 + kSides + 4 * kSlots + 3 * kSides;   // (unused), obs-ready[2], d1[6], relu[6], logits[2], value[2], value-free[2]
    static constexpr int kTmem = kBar + 8 * kNumBars;
    static constexpr int kRoot = kTmem + 16;                         // the root's node record (every game starts there)
    static constexpr int kBytes = kRoot + round_up(ev_stride_of(A) * 4, 16);
    static_assert(kBytes <= 227 * 1024, "shared memory plan does not fit");
    static_assert(kImageBytes % 16 == 0 && kObs % 16 == 0 && kCand % 16 == 0 && kBar % 8 == 0 && (32 * KIN * 4) % 16 == 0,
                  "alignment");
    static_assert(kObsCols <= 32 && kObsCols % 8 == 0, "observation columns do not fit");
    static_assert(kMmaWarps == kChunks, "one MMA warp per chunk index");
};

// D = f32; A and B formats: 0 = f16, 2 = tf32; both K-major
__host__ __device__ constexpr uint32_t instr_desc(int n, bool f16) {
    return (1u << 4) | ((f16 ? 0u : 2u) << 7) | ((f16 ? 0u : 2u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

// operand element (row, k) of a K-major, no-swizzle [rows x KP] operand with ESZ-byte elements: 8-row x 16-byte core
// matrices, the K chunks of a row group adjacent (LBO = 128 B), row groups KP * ESZ * 8 bytes apart
template <int KP, int ESZ>
__host__ __device__ __forceinline__ uint32_t op_off(int row, int k) {
    constexpr int kPer = 16 / ESZ;
    return (uint32_t)((row >> 3) * (KP * ESZ * 8) + (k / kPer) * 128 + (row & 7) * 16 + (k % kPer) * ESZ);
}

// two fp32 -> one 32-bit word of two fp16 (round to nearest even), element `lo` in bits 0..15: a tensor-memory column of a
// 16-bit A operand holds two K-adjacent elements, the even one in the low half
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

template <bool F16>
__device__ __forceinline__ void mma_ts_k(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool acc) {
    if constexpr (F16) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
            "}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
            : "memory");
    } else {
        mma_ts(d_tmem, a_tmem, b_desc, idesc, acc);
    }
}

// register re-allocation between the warp roles (all four warps of a warpgroup execute the same one)
template <int N>
__device__ __forceinline__ void regs_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void regs_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
// launch: 896 threads x 72; after re-allocation 8 x 32 x 104 + 16 x 32 x 56 + 4 x 32 x 72 = 64,512 of 65,536
constexpr int kRegsHead = 104, kRegsEpi = 56;
static_assert(kHeadWarps % 4 == 0 && kEpiWarps % 4 == 0, "roles must fill whole warpgroups");

#ifdef RNAD_TRACE
// development aid: cycle stamps of CTA 0's first tile pair.  [role][half-move][event]
__device__ long long g_trace[5][16][24];
#define TR(role, t, ev) do { if (blockIdx.x == 0 && first_pair && (t) < 16) g_trace[role][t][ev] = clock64(); } while (0)
// stream items: flat [role 2 = MMA warp, role 3 = epilogue warp 0][2 * item + which]
#define TRI(role, item, which) do { if (blockIdx.x == 0 && (item) >= 16 && (item) < 16 + 40) (&g_trace[role][0][0])[8 * ((item) - 16) + (which)] = clock64(); } while (0)
#else
#define TR(role, t, ev) do { } while (0)
#define TRI(role, item, which) do { } while (0)
#endif

// relu in place of the first-layer accumulators, split over the two arithmetic pipes of a scheduler: even hidden units
// max(x, 0) (FMNMX, alu pipe), odd hidden units x + |x| = 2 relu(x) (FADD, fma pipe) - exactly twice the value for every
// finite x, and pack_weights halves the second-layer weights of the odd units (a power of two: bit-identical products)
__device__ __forceinline__ float relu_split(float x, int k) { return (k & 1) ? x + fabsf(x) : fmaxf(x, 0.f); }
__device__ __forceinline__ float relu_split_scale(int hidden) { return (hidden & 1) ? 0.5f : 1.f; }

// The weight image in MMA-operand order, written straight into the CTA's shared memory by all of its threads
// (43 KB of fp32 nn.Linear tensors from L2 -> 66 KB of tf32 operands): no pre-kernel, no workspace.
template <int A, bool F16>
__device__ __forceinline__ void pack_weights(const rnad_mlp_weights& w, uint8_t* __restrict__ image, int thread) {
    using P = Plan<A, F16>;
    constexpr int KIN = P::KIN, KP = P::KP;
    static_assert(kThreads >= 2 * kHidden, "one thread per first-layer row");
    if (thread < 2 * kHidden) {
        // first layers: thread n owns hidden unit n of [value trunk | policy trunk]: one row of 2A^2 inputs (scalar loads,
        // all in flight at once; the tensors are only guaranteed 4-byte aligned) -> sixteen-byte operand chunks
        const int n = thread, j = n & (kHidden - 1);
        const float* src = (n < kHidden ? w.value_fc0_w : w.policy_fc0_w) + j * KIN;
        float x[KP];
#pragma unroll
        for (int k = 0; k < KIN; ++k) x[k] = __ldg(src + k);
        const float bias = __ldg((n < kHidden ? w.value_fc0_b : w.policy_fc0_b) + j);
#pragma unroll
        for (int k = KIN; k < KP; ++k) x[k] = (P::kBiasInK && k == KIN) ? bias : 0.f;
        if constexpr (F16) {
#pragma unroll
            for (int q = 0; q < KP / 8; ++q)
                *reinterpret_cast<uint4*>(image + P::kW1 + op_off<KP, 2>(n, 8 * q)) =
                    make_uint4(pack_f16x2(x[8 * q], x[8 * q + 1]), pack_f16x2(x[8 * q + 2], x[8 * q + 3]),
                               pack_f16x2(x[8 * q + 4], x[8 * q + 5]), pack_f16x2(x[8 * q + 6], x[8 * q + 7]));
        } else {
#pragma unroll
            for (int q = 0; q < KP / 4; ++q)
                *reinterpret_cast<float4*>(image + P::kW1 + op_off<KP, 4>(n, 4 * q)) =
                    make_float4(to_tf32_fast(x[4 * q]), to_tf32_fast(x[4 * q + 1]), to_tf32_fast(x[4 * q + 2]), to_tf32_fast(x[4 * q + 3]));
        }
        reinterpret_cast<float*>(image + P::kB1)[n] = bias;
        // second layers as ONE [16 x 512] operand; thread k owns column k (hidden unit k of [value trunk | policy trunk]).
        // Row 0 = value_fc1, rows 1..A = policy_fc1 for the hidden units of a trunk's FIRST 128-chunk, rows 8 and 9..8+A
        // for those of its SECOND chunk, zero elsewhere: the two chunks of a trunk leave their partial sums in separate
        // columns of the accumulator (the other chunk's MMAs add exact zeros there) and the head adds them in a fixed
        // order - the result does not depend on the order in which the MMA warps got to issue.
        const int k = thread;
        float col[kN2];
#pragma unroll
        for (int r = 0; r < kN2; ++r) col[r] = 0.f;
        const int r0 = (k & kChunk) ? 8 : 0;
        if (k < kHidden) {
            const float wv = __ldg(w.value_fc1_w + k);
            col[0] = r0 == 0 ? wv : 0.f;
            col[8] = r0 == 8 ? wv : 0.f;
        } else {
#pragma unroll
            for (int a = 0; a < A; ++a) {
                const float wp = __ldg(w.policy_fc1_w + a * kHidden + (k - kHidden));
                col[1 + a] = r0 == 0 ? wp : 0.f;
                col[9 + a] = r0 == 8 ? wp : 0.f;
            }
        }
        if constexpr (F16) {
#pragma unroll
            for (int r = 0; r < kN2; ++r)
                *reinterpret_cast<uint16_t*>(image + P::kW2 + op_off<kK2, 2>(r, k)) = (uint16_t)(pack_f16x2(col[r], 0.f) & 0xffffu);
        } else {
#pragma unroll
            for (int r = 0; r < kN2; ++r)    // (relu_split: the epilogue hands odd hidden units over doubled)
                *reinterpret_cast<float*>(image + P::kW2 + op_off<kK2, 4>(r, k)) = to_tf32_fast(col[r]) * relu_split_scale(k);
        }
    }
    if (thread < 8) {
        float v = 0.f;
        if (thread == 0) v = __ldg(w.value_fc1_b);
        else if (thread <= A) v = __ldg(w.policy_fc1_b + thread - 1);
        reinterpret_cast<float*>(image + P::kB2)[thread] = v;
    }
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// masked softmax on the fast exp2 / reciprocal units (net.py:45-46: e = where(mask, exp(logit), 0); e / max(sum e, 1e-12)).
// A few ulp from the exact formula; the recorded policy is the one every later decision uses.
template <int A>
__device__ __forceinline__ void masked_softmax_fast(const float (&logit)[A], int n_legal, float (&policy)[A]) {
    float e[A];
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        e[a] = a < n_legal ? ex2_approx(logit[a] * 1.4426950408889634f) : 0.f;
        sum += e[a];
    }
    const float inv = rcp_approx(fmaxf(sum, 1e-12f));
#pragma unroll
    for (int a = 0; a < A; ++a) policy[a] = e[a] * inv;
}

// Row half-move, off the critical path: for EVERY column action c the opponent may answer with, the chance draw
// at uniform u (episode.py:106-121), the child id, the reward and the child's node record - two rounds of
// independent 16-byte gathers - parked in shared memory as [c][word][game]: child, reward, ev[A*A], rows|cols.
template <int A, int C>
__device__ __forceinline__ void gather_candidates(const uint32_t* __restrict__ tr_tab, const uint32_t* __restrict__ ev_tab,
                                                  int node, int row_action, float u, bool active, uint32_t* s_cand,
                                                  int cand_words) {
    constexpr int TRS = tr_stride_of(C);
    constexpr int EVS = ev_stride_of(A);
    uint32_t w[A][TRS];
    const uint4* ent = reinterpret_cast<const uint4*>(tr_tab + ((int64_t)node * A * A + row_action * A) * TRS);
#pragma unroll
    for (int c = 0; c < A; ++c)
#pragma unroll
        for (int q = 0; q < TRS / 4; ++q) {
            const uint4 v = active ? __ldg(ent + c * (TRS / 4) + q) : make_uint4(0u, 0u, 0u, 0u);
            w[c][4 * q + 0] = v.x;
            w[c][4 * q + 1] = v.y;
            w[c][4 * q + 2] = v.z;
            w[c][4 * q + 3] = v.w;
        }
    int child[A];
    float rew[A];
#pragma unroll
    for (int c = 0; c < A; ++c) {
        float acc = 0.f;
        uint32_t ch = w[c][C], val = w[c][2 * C];
        bool done = false;
#pragma unroll
        for (int k = 0; k < C; ++k) {       // same rule as transition() / sample_icdf
            const float pk = __uint_as_float(w[c][k]);
            acc = __fadd_rn(acc, pk);
            const bool positive = pk > 0.f;
            if (!done && positive) {
                ch = w[c][C + k];
                val = w[c][2 * C + k];
            }
            done = done || (positive && (u < acc));
        }
        child[c] = active ? (int)ch : 0;
        rew[c] = child[c] == 0 ? __uint_as_float(val) : 0.f;
    }
    uint4 rec[A][EVS / 4];
#pragma unroll
    for (int c = 0; c < A; ++c)
#pragma unroll
        for (int q = 0; q < EVS / 4; ++q)
            rec[c][q] = active ? __ldg(reinterpret_cast<const uint4*>(ev_tab + (int64_t)child[c] * EVS) + q)
                               : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int c = 0; c < A; ++c) {
        uint32_t* dst = s_cand + c * cand_words * kTileM;
        dst[0] = (uint32_t)child[c];
        dst[kTileM] = __float_as_uint(active ? rew[c] : 0.f);
        const uint32_t* r = reinterpret_cast<const uint32_t*>(rec[c]);
#pragma unroll
        for (int i = 0; i < A * A; ++i) dst[(2 + i) * kTileM] = r[i];
        dst[(2 + A * A) * kTileM] = active ? (r[A * A] & 0xffffu) : 0x0101u;
    }
}

template <int A, int C, bool F16>
__global__ void __launch_bounds__(kThreads, 1) rollout_tc2_kernel(RolloutArgs g) {
    using P = Plan<A, F16>;
    constexpr int KIN = P::KIN, KP = P::KP;
    static_assert(!F16 || kEpiWarps == 8, "the fp16 epilogue exists for eight epilogue warps (both layouts)");
    static_assert(A <= 4, "value + logits must fit the 8 useful columns of the second-layer accumulator");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const uint32_t bar0 = smem_u32(smem + P::kBar);
    const uint32_t bar_img = bar0;
    auto bar_a = [&](int side) { return bar0 + 8 + 8 * side; };                       // observations of `side` in TMEM
    // Per stream item i: first layers complete (tcgen05.commit) and relu written back (one arrival per epilogue warp
    // of the item), barrier i % (2 kSlots) each.  Two barriers per slot: a waiter that skips items (the epilogue warps
    // take every other item, an MMA warp every fourth) could otherwise see its parity wait for item i satisfied by the
    // completion of item i - 2 kSlots while item i - kSlots is still in flight.
    auto bar_d1 = [&](int k) { return bar0 + 8 + 8 * kSides + 8 * k; };
    auto bar_relu = [&](int k) { return bar0 + 8 + 8 * kSides + 8 * (2 * kSlots + k); };
    auto bar_d2p = [&](int side) { return bar0 + 8 + 8 * kSides + 32 * kSlots + 8 * side; };                // logits complete
    auto bar_d2v = [&](int side) { return bar0 + 8 + 8 * kSides + 32 * kSlots + 8 * (kSides + side); };     // value complete
    auto bar_vfree = [&](int side) { return bar0 + 8 + 8 * kSides + 32 * kSlots + 8 * (2 * kSides + side); };   // value read, cleared
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

#ifdef RNAD_TRACE
    if (blockIdx.x == 0 && tid == 0) g_trace[0][15][0] = clock64();
#endif
    if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
    if (tid == 0) {
        for (int s = 0; s < kSides; ++s) {
            mbar_init(bar_a(s), 4);              // one arrival per head warp of the side
            mbar_init(bar_d2p(s), 2);            // the two MMA warps of the policy chunks commit
            mbar_init(bar_d2v(s), 2);            // the two of the value chunks
            mbar_init(bar_vfree(s), 4);          // one arrival per head warp of the side
        }
        for (int s = 0; s < 2 * kSlots; ++s) {
            mbar_init(bar_d1(s), 1);
            mbar_init(bar_relu(s), kEpiPerItem);
        }
        mbar_fence_init();
    }
    // every game starts at the root (node 1): its record is fetched once, underneath the weight packing
    uint32_t root_word = 0;
    if (tid < ev_stride_of(A)) root_word = __ldg(g.ev_tab + ev_stride_of(A) + tid);
    pack_weights<A, F16>(g.w, smem, tid);
    if (tid < ev_stride_of(A)) reinterpret_cast<uint32_t*>(smem + P::kRoot)[tid] = root_word;
    fence_async_smem();          // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
#ifdef RNAD_TRACE
    if (blockIdx.x == 0 && tid == 0) g_trace[0][15][1] = clock64();
#endif
    const uint32_t tmem_base = *tmem_slot;
    const int64_t num_tiles = (g.B + kTileM - 1) / kTileM;
    const int64_t num_pairs = (num_tiles + kSides - 1) / kSides;
    int64_t my_pairs = 0;                        // pairs this CTA plays: blockIdx.x, blockIdx.x + gridDim.x, ...
    if ((int64_t)blockIdx.x < num_pairs) my_pairs = (num_pairs - 1 - blockIdx.x) / gridDim.x + 1;

    if (warp >= kMmaWarp) {
        // ------------------------------------------------------------ MMA issuers (keep the launch allocation)
        // The whole warp runs the loop converged and one elected lane issues: the descriptors then live in
        // uniform registers (an `if (lane == 0)` branch makes ptxas wrap every UTCHMMA in a waterfall loop).
        const uint64_t w1_desc = desc_sbo(smem_u32(smem + P::kW1), P::kSbo1);
        const uint64_t w2_desc = desc_sbo(smem_u32(smem + P::kW2), P::kSbo2);
        constexpr uint32_t kIdesc1 = instr_desc(kChunk, F16), kIdesc2 = instr_desc(kN2, F16);
        // Stream item i = ((pair * T + t) * kSides + side) * kChunks + c lives in slot i % kSlots.  Its step is
        //     wait relu(i) -> MMA2(i) -> MMA1(i + kSlots) into the slot MMA2(i) has just read -> commit,
        // and MMA warp w takes the items with c == w: issuing blocks while the tensor core's queue is full and every
        // step has barrier latency around it, so several issuers keep the queue fed.  MMAs of one thread execute in
        // issue order (that covers the slot reuse); MMA2s of different warps all ADD into the side's accumulator, each
        // chunk's non-zero weights into its own 8 columns (pack_weights), which the head adds in a fixed order and zeroes
        // after reading: their order does not matter, not even for the last bit.
        const int w = warp - kMmaWarp;           // == chunk index c of every item this warp issues
        const uint32_t n_items = (uint32_t)(my_pairs * g.T * kSides * kChunks);
        int seen_hm[kSides] = {-1, -1};
        // the observations of half-move `hm` of `side` must have been published before a first-layer MMA reads them
        // (each MMA warp checks for itself: the warps do not order each other)
        auto need_obs = [&](int side, int hm) {
            if (hm != seen_hm[side]) {
                mbar_wait_c(bar_a(side), (uint32_t)hm & 1u);
                tc_fence_after();
                seen_hm[side] = hm;
            }
        };
        // first layers of a chunk: A = observations in tensor memory (called by the elected lane)
        auto mma1 = [&](int c, int side, int slot, int bar_index) {
#pragma unroll
            for (int s = 0; s < KP / P::kKStep; ++s)   // (8 tensor-memory columns and 256 operand bytes per K step, either kind)
                mma_ts_k<F16>(tmem_base + slot * kChunk, tmem_base + obs_col(side) + s * 8,
                              w1_desc + (uint64_t)(((chunk_hidden(c) / 8) * P::kSbo1 + s * 256) >> 4), kIdesc1, s > 0);
            mma_commit(bar_d1(bar_index));
        };
        if (w == 0 && n_items > 0) {             // fill the ring: items 0 .. kSlots-1 are chunks 0 .. 2 of (side 0, half-move 0)
            need_obs(0, 0);
            if (elect_one())
                for (int j = 0; j < kSlots; ++j) mma1(j, 0, j, j);
            __syncwarp();
        }
        // item i = w, w + 4, ...: all index arithmetic incremental (no divisions in the loop)
        int slot = w % kSlots, rb = w % (2 * kSlots);   // i % 3, i % 6
        uint32_t par = 0;                                // (i / 6) & 1
        int side = 0, hm = 0;                            // (i / 4) & 1, i / 8
        const int cj = (w + kSlots) & (kChunks - 1);     // chunk index of item j = i + 3
        int side_j = (w + kSlots) >> 2, hm_j = 0;        // its side ((j / 4) & 1) and half-move (j / 8), j = w + 3 < 8
#pragma unroll 1
        for (uint32_t i = w; i < n_items; i += kMmaWarps) {
            const bool has_j = i + kSlots < n_items;
            if (has_j) need_obs(side_j, hm_j);
            if (w >= 2) {                // value chunks: the heads have read and cleared the previous half-move's value
                mbar_wait_c(bar_vfree(side), (uint32_t)hm & 1u);
                tc_fence_after();
            }
            mbar_wait_c(bar_relu(rb), par);
            tc_fence_after();
            TRI(2, i, 0);
            if (elect_one()) {
#pragma unroll
                for (int s = 0; s < kChunk / P::kKStep; ++s)   // second layers: A = relu(hidden) in tensor memory
                    // (fp16: the epilogue warps of a 64-column half pack it into the first 32 columns of that half)
                    mma_ts_k<F16>(tmem_base + (w < 2 ? d2p_col(side) : d2v_col(side)),
                                  tmem_base + slot * kChunk + (F16 ? (s >> 2) * 64 + (s & 3) * 8 : s * 8),
                                  w2_desc + (uint64_t)(((chunk_hidden(w) / P::kKStep + s) * 256) >> 4), kIdesc2, true);
                mma_commit(w < 2 ? bar_d2p(side) : bar_d2v(side));
                TRI(2, i, 1);
                if (has_j) mma1(cj, side_j, slot, rb >= kSlots ? rb - kSlots : rb + kSlots);   // item i + 3, into the slot just read
            }
            __syncwarp();
            TRI(2, i, 2);
#ifdef RNAD_TRACE
            if (has_j && blockIdx.x == 0 && i >= 16 && i < 56) {   // when do the first layers just issued complete?
                mbar_wait_c(bar_d1((i + kSlots) % (2 * kSlots)), ((i + kSlots) / (2 * kSlots)) & 1u);
                TRI(2, i, 5);
            }
#endif
            // i += 4
            slot = slot + 1 == kSlots ? 0 : slot + 1;
            rb += kMmaWarps;
            if (rb >= 2 * kSlots) {
                rb -= 2 * kSlots;
                par ^= 1u;
            }
            side ^= 1;
            if (side == 0) ++hm;
            side_j ^= 1;
            if (side_j == 0) ++hm_j;             // j / 8 advances whenever j / 4 becomes even
        }
    } else if (warp >= kHeadWarps) {
        // ------------------------------------------------------------ epilogue of the first layers
        if (kEpiWarps == 16) regs_dec<kRegsEpi>();
        const int e = warp - kHeadWarps;
        // lane quadrant, 64-column half, item parity
        const int quad = e & 3, half = kEpiWide ? 0 : (e >> 2) & 1, parity = kEpiWide ? (e >> 2) & 1 : e >> 3;
        constexpr int kCols = kChunk / 2;
        const uint32_t tmem_mine = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * kCols);
        const float* b1 = reinterpret_cast<const float*>(smem + P::kB1) + half * kCols;
        const uint32_t n_items = (uint32_t)(my_pairs * g.T * kSides * kChunks);
        // one stream item: wait for its first layers, relu this warp's 32 lanes x 64 columns in place, hand it to MMA2
        auto epi_item = [&](uint32_t i, int slot, int rb, int c, uint32_t par) {
            (void)i;
            mbar_wait_c(bar_d1(rb), par);
            tc_fence_after();
            if (e == 0) TRI(2, i, 3);
            TRI(4, i, e);
            const uint32_t taddr = tmem_mine + slot * kChunk;
            if (kEpiWarps == 8) {
#pragma unroll
              for (int hh = 0; hh < (kEpiWide ? 2 : 1); ++hh) {
                // both 32-column loads in flight, one wait
                const uint32_t taddr = tmem_mine + slot * kChunk + hh * kCols;
                uint32_t r[kCols];
#pragma unroll
                for (int q = 0; q < kCols / 32; ++q) tmem_ld32p(taddr + q * 32, r + q * 32);
                tmem_ld_wait();
                if (e == 0) TRI(2, i, 6);
                if (!P::kBiasInK) {
                    const float4* bias = reinterpret_cast<const float4*>(b1 + chunk_hidden(c) + hh * kCols);
#pragma unroll
                    for (int k = 0; k < kCols / 4; ++k) {
                        const float4 bb = bias[k];
                        r[4 * k + 0] = __float_as_uint(__uint_as_float(r[4 * k + 0]) + bb.x);
                        r[4 * k + 1] = __float_as_uint(__uint_as_float(r[4 * k + 1]) + bb.y);
                        r[4 * k + 2] = __float_as_uint(__uint_as_float(r[4 * k + 2]) + bb.z);
                        r[4 * k + 3] = __float_as_uint(__uint_as_float(r[4 * k + 3]) + bb.w);
                    }
                }
                if constexpr (F16) {
                    uint32_t pk[kCols / 2];        // relu, round to fp16, pack: one instruction per two hidden units
#pragma unroll
                    for (int k = 0; k < kCols / 2; ++k)
                        pk[k] = pack_relu_f16x2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
                    tmem_st32(taddr, pk);
                } else {
#pragma unroll
                    for (int k = 0; k < kCols; ++k) r[k] = __float_as_uint(relu_split(__uint_as_float(r[k]), k));
#pragma unroll
                    for (int q = 0; q < kCols / 32; ++q) tmem_st32(taddr + q * 32, r + q * 32);
                }
              }
            } else {
#pragma unroll
                for (int q = 0; q < kCols / 32; ++q) {   // 32 columns at a time: 32 data registers (the 16-warp budget)
                    uint32_t r[32];
                    tmem_ld32p(taddr + q * 32, r);
                    tmem_ld_wait();
                    if (!P::kBiasInK) {
                        const float4* bias = reinterpret_cast<const float4*>(b1 + chunk_hidden(c) + q * 32);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 bb = bias[k];
                            r[4 * k + 0] = __float_as_uint(__uint_as_float(r[4 * k + 0]) + bb.x);
                            r[4 * k + 1] = __float_as_uint(__uint_as_float(r[4 * k + 1]) + bb.y);
                            r[4 * k + 2] = __float_as_uint(__uint_as_float(r[4 * k + 2]) + bb.z);
                            r[4 * k + 3] = __float_as_uint(__uint_as_float(r[4 * k + 3]) + bb.w);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(relu_split(__uint_as_float(r[k]), k));
                    tmem_st32(taddr + q * 32, r);
                }
            }
            tmem_st_wait();
            if (e == 0) TRI(2, i, 7);
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar_relu(rb));
            if (e == 0) TRI(2, i, 4);
            TRI(3, i, e);
        };
        if constexpr (kEpiWarps == 8 && !kEpiWide) {
            // every warp takes every item: unrolled over the period of the slot ring (3), the barrier ring (6) and the
            // chunks of a half-move (4), so that tensor-memory addresses, barrier addresses and parities are immediates -
            // what an epilogue warp does between two items is on the critical path of the stream
            constexpr int kPeriod = 12;
#pragma unroll 1
            for (uint32_t i0 = 0; i0 < n_items; i0 += kPeriod) {
#pragma unroll
                for (int u = 0; u < kPeriod; ++u) {
                    if (u > 0 && (u & 3) == 0 && i0 + u >= n_items) break;    // (n_items is a multiple of 8)
                    epi_item(i0 + u, u % kSlots, u % (2 * kSlots), u % kChunks, (uint32_t)(u / (2 * kSlots)) & 1u);
                }
            }
        } else {
            int slot = parity, rb = parity, c = parity;      // i % 3, i % 6, i % 4 of item i = parity, parity + kEpiStride, ...
            uint32_t par = 0;                                // (i / 6) & 1
#pragma unroll 1
            for (uint32_t i = parity; i < n_items; i += kEpiStride) {
                epi_item(i, slot, rb, c, par);
                slot = slot + kEpiStride >= kSlots ? slot + kEpiStride - kSlots : slot + kEpiStride;
                rb += kEpiStride;
                if (rb >= 2 * kSlots) {
                    rb -= 2 * kSlots;
                    par ^= 1u;
                }
                c = (c + kEpiStride) & (kChunks - 1);
            }
        }
    } else {
        // ------------------------------------------------------------ heads: one thread per game
        if (kEpiWarps == 16) regs_inc<kRegsHead>();
        const int side = warp >> 2;
        const int lane_g = tid & (kTileM - 1);                // game of the tile == TMEM lane
        const int lane = tid & 31;
        const int hw = warp & 3;                              // head warp of the side == lane quadrant
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(hw * 32) << 16);
        const uint32_t my_d2p = tmem_lane + d2p_col(side), my_d2v = tmem_lane + d2v_col(side), my_obs = tmem_lane + obs_col(side);
        float* s_obs = reinterpret_cast<float*>(smem + P::kObs) + (side * kTileM + hw * 32) * KIN;   // this warp's 32 rows
        uint32_t* s_cand = reinterpret_cast<uint32_t*>(smem + P::kCand) + side * A * P::kCandWords * kTileM + lane_g;
        const float b2v = reinterpret_cast<const float*>(smem + P::kB2)[0];
        float b2p[A];
#pragma unroll
        for (int a = 0; a < A; ++a) b2p[a] = reinterpret_cast<const float*>(smem + P::kB2)[1 + a];
        const uint32_t my_bar_a = bar_a(side), my_bar_d2p = bar_d2p(side), my_bar_d2v = bar_d2v(side), my_bar_vfree = bar_vfree(side);

        uint32_t ph_d2 = 0;
        int last_valid = -1, n_valid0 = 0, n_valid1 = 0;
        const uint64_t seed = g.seed_dev != nullptr ? __ldg(g.seed_dev) : g.seed;
        for (int64_t k = 0; k < my_pairs; ++k) {
#ifdef RNAD_TRACE
            const bool first_pair = k == 0;
#endif
            const int64_t tile = ((int64_t)blockIdx.x + k * gridDim.x) * kSides + side;
            const int64_t tile_base = tile * kTileM;
            const int64_t b = tile_base + lane_g;
            const bool active = b < g.B;                      // (a side without a tile plays along with idle lanes)
            int node = active ? 1 : 0;
            int row_action = 0;
            float game_return = 0.f;
            Node<A> n;
#pragma unroll
            for (int i = 0; i < A * A; ++i) n.ev[i] = 0.f;
            n.rows = n.cols = 1;
            float x[KIN];

            // observation words of half-move t in tf32 (A operand of the first layers) -> tensor memory; critical path.
            // `prev_t` >= 0: the heads run ahead of the value trunk of half-move prev_t (they only waited for its logits),
            // whose first layers still READ the observation in place: wait for the first-layer commits of its two value
            // chunks (stream items i2, i2 + 1) before overwriting.  Normally complete long before - those MMAs are
            // issued ahead of the second layers of the policy chunks - but nothing else orders them.  (Parity: every
            // earlier item on these barriers is complete once the logits are, and the next one needs this publication.)
            auto publish_obs = [&](int prev_t) {
                tmem_ld_wait();          // this thread's read of the accumulators is complete before they are cleared
                if (prev_t >= 0) {
                    const uint32_t i2 = (((uint32_t)(k * g.T + prev_t)) * kSides + side) * kChunks + 2;
                    mbar_wait_c(bar_d1(i2 % (2 * kSlots)), (i2 / (2 * kSlots)) & 1u);
                    mbar_wait_c(bar_d1((i2 + 1) % (2 * kSlots)), ((i2 + 1) / (2 * kSlots)) & 1u);
                    tc_fence_after();
                }
                {
                    const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    tmem_st8(my_d2p, zero);  // the second-layer MMAs of the next half-move only ever accumulate
                    tmem_st8(my_d2p + 8, zero);
                }
                auto obs_k = [&](int kk) {   // element kk of the padded observation row (the constant 1 carries the bias)
                    return kk < KIN ? x[kk < KIN ? kk : 0] : ((P::kBiasInK && kk == KIN) ? 1.f : 0.f);
                };
#pragma unroll
                for (int q = 0; q < P::kObsCols / 8; ++q) {
                    uint32_t v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int col = 8 * q + u;
                        if constexpr (F16) v[u] = pack_f16x2(obs_k(2 * col), obs_k(2 * col + 1));
                        else v[u] = __float_as_uint(col < KIN ? to_tf32_fast(obs_k(col)) : obs_k(col));
                    }
                    tmem_st8(my_obs + 8 * q, v);
                }
                tmem_st_wait();
                tc_fence_before();       // orders the stores above before the MMAs the arrival releases
                __syncwarp();
                if (lane == 0) mbar_arrive(my_bar_a);
            };
            // the same observation in fp32 -> trajectory; off the critical path.  A warp's 32 rows are one
            // contiguous block of the (T,B,2,A,A) tensor: staged in shared memory, stored as coalesced 16-byte words.
            auto store_obs = [&](int t) {
                const int64_t row0 = (int64_t)t * g.B + tile_base + hw * 32;
                float* dst = g.out.observations + row0 * KIN;
                const int rows = (int)max((int64_t)0, min((int64_t)32, g.B - (tile_base + hw * 32)));
                if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    __syncwarp();        // the previous half-move's copy is done
#pragma unroll
                    for (int kk = 0; kk < KIN; kk += 2)
                        *reinterpret_cast<float2*>(s_obs + lane * KIN + kk) = make_float2(x[kk], x[kk + 1]);
                    __syncwarp();
                    const int n_float = rows * KIN;
                    for (int i = lane * 4; i < n_float; i += 128)
                        __stcs(reinterpret_cast<float4*>(dst + i), *reinterpret_cast<const float4*>(s_obs + i));
                } else if (active) {
                    float2* d = reinterpret_cast<float2*>(g.out.observations + ((int64_t)t * g.B + b) * KIN);
#pragma unroll
                    for (int kk = 0; kk < KIN / 2; ++kk) __stcs(d + kk, make_float2(x[2 * kk], x[2 * kk + 1]));
                }
            };
            auto draw = [&](int t) {
                Uniforms2 u;
                if (g.uniforms != nullptr) {
                    const int64_t slot_tb = (int64_t)t * g.B + b;
                    u.action = active ? __ldg(g.uniforms + slot_tb * 2 + 0) : 0.f;
                    u.chance = active ? __ldg(g.uniforms + slot_tb * 2 + 1) : 0.f;
                } else {
                    u = philox_uniforms(seed, (uint32_t)t, (uint64_t)(g.game_offset + b));
                }
                return u;
            };

            Uniforms2 u;
            u.action = u.chance = 0.f;
            // half-move -1 is the set-up of the tile: it only publishes the root observation
#pragma unroll 1
            for (int t = -1; t < g.T; ++t) {
                const int turn = t & 1;
                const bool more = t + 1 < g.T;
                // the value accumulator: read (t >= 0), cleared, handed back to the MMA warps of the value chunks
                auto take_value = [&]() {
                    float v = 0.f;
                    if (t >= 0) {
                        mbar_wait_c(my_bar_d2v, ph_d2);        // (same parity sequence as the logits barrier)
                        tc_fence_after();
                        uint32_t dv[8], dw[8];                 // partial sums of the trunk's two chunks
                        tmem_ld8(my_d2v, dv);
                        tmem_ld8(my_d2v + 8, dw);
                        tmem_ld_wait();
                        v = (__uint_as_float(dv[0]) + __uint_as_float(dw[0])) + b2v;
                    }
                    const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    tmem_st8(my_d2v, zero);
                    tmem_st8(my_d2v + 8, zero);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(my_bar_vfree);
                    return v;
                };
                if (t < 0) {
                    if (active) {
                        const uint32_t* rw = reinterpret_cast<const uint32_t*>(smem + P::kRoot);
#pragma unroll
                        for (int i = 0; i < A * A; ++i) n.ev[i] = __uint_as_float(rw[i]);
                        n.rows = rw[A * A] & 0xff;
                        n.cols = (rw[A * A] >> 8) & 0xff;
                    }
                    build_obs<A>(n, 0, x);
                    publish_obs(-1);
                    if (k == 0) take_value();   // (later pairs: cleared and handed back after the previous pair's last half-move)
                } else {
                    if (node != 0) {
                        last_valid = max(last_valid, t);
                        n_valid0 += turn == 0;
                    n_valid1 += turn;
                    }
                    const int n_legal = turn == 0 ? n.rows : n.cols;
                    // a row half-move does not move the game: the column player's observation is known beforehand
                    if (turn == 0) build_obs<A>(n, 1, x);
                    if (lane_g == 0) TR(side, t, 0);
                    mbar_wait_c(my_bar_d2p, ph_d2);                // the logits; the value trunk is still in the ring
                    tc_fence_after();
                    if (lane_g == 0) TR(side, t, 1);
                    uint32_t d2[8], d3[8];                         // partial sums of the policy trunk's two chunks
                    tmem_ld8(my_d2p, d2);
                    tmem_ld8(my_d2p + 8, d3);
                    if (turn == 0 && more) publish_obs(t);         // critical path of a row half-move ends here
                    tmem_ld_wait();
                    float logit[A];
#pragma unroll
                    for (int a = 0; a < A; ++a) logit[a] = (__uint_as_float(d2[1 + a]) + __uint_as_float(d3[1 + a])) + b2p[a];
                    float policy[A];
                    masked_softmax_fast<A>(logit, n_legal, policy);
                    const int action = sample_icdf(policy, A, u.action);
                    float reward = 0.f;
                    const int node_now = node;
                    if (turn == 0) {
                        row_action = action;
                    } else {
                        // the outcome of (node, row_action, action) was gathered during the row half-move
                        const uint32_t* cand = s_cand + action * P::kCandWords * kTileM;
#pragma unroll
                        for (int i = 0; i < A * A; ++i) n.ev[i] = __uint_as_float(cand[(2 + i) * kTileM]);
                        const uint32_t rc = cand[(2 + A * A) * kTileM];
                        n.rows = rc & 0xff;
                        n.cols = rc >> 8;
                        node = (int)cand[0];
                        reward = __uint_as_float(cand[kTileM]);
                        if (more) {
                            build_obs<A>(n, 0, x);
                            publish_obs(t);                        // critical path of a column half-move ends here
                        }
                    }
                    if (lane_g == 0) TR(side, t, 3);
                    // ---- from here on off the critical path: the tensor core and the epilogue warps are busy
                    const float value = take_value();
                    ph_d2 ^= 1u;
                    game_return += reward;
                    if (active)
                        write_record<A>(g.out, (int64_t)t * g.B + b, node_now, turn, n_legal, policy, action, value, reward, logit);
                }
                if (more) {
                    store_obs(t + 1);
                    u = draw(t + 1);
                    if (t >= 0 && turn == 0)
                        gather_candidates<A, C>(g.tr_tab, g.ev_tab, node, row_action, u.chance, active, s_cand, P::kCandWords);
                }
                if (t >= 0 && lane_g == 0) TR(side, t, 4);
            }
            if (active && g.out.returns != nullptr) g.out.returns[b] = game_return;
        }
        publish_stats(g.stats, last_valid, n_valid0, n_valid1, lane);
    }

    tc_fence_before();
    __syncthreads();
#ifdef RNAD_TRACE
    if (blockIdx.x == 0 && tid == 0) g_trace[0][15][2] = clock64();
#endif
    if (warp == kMmaWarp) tmem_dealloc<kTmemCols>(tmem_base);
}

template <int A, int C, bool F16>
static int launch(const RolloutArgs& g, cudaStream_t st) {
    using P = Plan<A, F16>;
    // One CTA per SM owns all 512 TMEM columns: ask for more than half of the shared memory so that a second
    // CTA can never become resident and spin inside tcgen05.alloc.
    size_t smem = P::kBytes;
    const size_t floor_one_per_sm = 116 * 1024;
    if (smem < floor_one_per_sm) smem = floor_one_per_sm;
    // function attributes are per device: set once per device and instantiation (a few microseconds of host time each)
    static std::atomic<int> configured_device{-1};
    int device = 0;
    int rc = check_cuda(cudaGetDevice(&device), "cudaGetDevice");
    if (rc) return rc;
    if (configured_device.load(std::memory_order_acquire) != device) {
        rc = check_cuda(cudaFuncSetAttribute(rollout_tc2_kernel<A, C, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "cudaFuncSetAttribute(rollout_tc2, smem)");
        if (rc) return rc;
        rc = check_cuda(cudaFuncSetAttribute(rollout_tc2_kernel<A, C, F16>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             cudaSharedmemCarveoutMaxShared),
                        "cudaFuncSetAttribute(rollout_tc2, carveout)");
        if (rc) return rc;
        configured_device.store(device, std::memory_order_release);
    }
    const int64_t tiles = (g.B + kTileM - 1) / kTileM;
    int64_t blocks = (tiles + kSides - 1) / kSides;
    if (blocks > sm_count()) blocks = sm_count();
    rollout_tc2_kernel<A, C, F16><<<(int)blocks, kThreads, smem, st>>>(g);
    RNAD_CHECK_LAUNCH("rollout_tc2_kernel");
    return RNAD_OK;
}

}  // namespace tc2

#ifdef RNAD_TRACE
extern "C" __attribute__((visibility("default"))) int rnad_debug_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, tc2::g_trace, sizeof(tc2::g_trace));
}
#endif

bool rollout_tc2_supported(int A, int width, int C) { return rollout_tc_supported(A, width) && C >= 1 && C <= 4; }

int64_t rollout_tc2_workspace_bytes(int A) {
    (void)A;
    return 0;   // the weight image is built in shared memory by the kernel itself
}

int rollout_tc2(const RolloutArgs& g, void* workspace, cudaStream_t st, bool f16) {
    if (!rollout_tc2_supported(g.A, g.w.width, g.C)) {
        set_error("rnad_rollout(tf32x2): needs width == 256, 2 <= max_actions <= 4 and max_transitions <= 4 "
                  "(got width %d, max_actions %d, max_transitions %d)", g.w.width, g.A, g.C);
        return RNAD_EUNSUPPORTED;
    }
    (void)workspace;
#define RNAD_TC2_CASE(a, c) \
    if (g.A == a && g.C == c) return f16 ? tc2::launch<a, c, true>(g, st) : tc2::launch<a, c, false>(g, st);
    RNAD_TC2_CASE(2, 1) RNAD_TC2_CASE(2, 2) RNAD_TC2_CASE(2, 3) RNAD_TC2_CASE(2, 4)
    RNAD_TC2_CASE(3, 1) RNAD_TC2_CASE(3, 2) RNAD_TC2_CASE(3, 3) RNAD_TC2_CASE(3, 4)
    RNAD_TC2_CASE(4, 1) RNAD_TC2_CASE(4, 2) RNAD_TC2_CASE(4, 3) RNAD_TC2_CASE(4, 4)
#undef RNAD_TC2_CASE
    return RNAD_EINVAL;
}

}  // namespace rnad
