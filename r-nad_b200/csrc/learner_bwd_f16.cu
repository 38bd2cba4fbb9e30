// Learner backward with fp16 operands (kind::f16) - the engine of the learner step (split mode, max_actions <= 3).
//
// The mask formulation of learner_bwd_tc3.cu,
//     dW1[j, k] = sum_a W2[a, j] * ( sum_n M[j, n] (g_a[n] x_k[n]) ),      dW2[a, j] = sum_n relu(h)[j, n] g_a[n],
// with every tensor-core operand in fp16 - the same 11-bit significand as tf32, like the rollout engine
// RNAD_PREC_F16X2 and the learner forward (whose first-layer arithmetic the recompute here repeats operand for
// operand).  The 0/1 mask M is exact in any format; relu(h) was already fp16 in the forward; g_a x_k and g_a are
// rounded once (fp32 product -> fp16, saturating).  That is safe for UNNORMALISED gradients - d_v, d_logit of
// rnad_learner_targets in unnormalised mode are O(1) and bounded by the NeuRD clip - and would lose precision for
// gradients already divided by ~10^5 step counts (fp16 is subnormal below 6e-5), which is why the flat entry point
// rnad_learner_backward keeps the tf32 kernels.
//
// What fp16 buys beyond K = 16 per MMA: relu^T and M^T of a 64-row stage are 32 + 32 packed columns, so tensor memory
// holds H^T (fp32, written by the recompute MMA, read by the consumers) and relu^T | M^T (written by the consumers, read
// by the gradient MMAs) in SEPARATE regions, two of each.  The issuer recomputes stage s + 2 as soon as the consumers
// have LOADED stage s - before they have packed and stored it and before grad(s) has run - so the chain
// "consumers -> gradient MMAs -> recompute -> consumers" of the tf32 kernels (one stage buffer serves all three) is cut:
//     consumers (group b):  wait H(s) -> tcgen05.ld -> [H free] -> relu / mask, pack -> wait grad(s - 2) -> tcgen05.st -> [RM ready]
//     issuer b:             wait [H free] -> recompute(s + 2) -> commit H;   wait [RM ready] -> grad(s) -> commit
// ONE CTA per SM, 512 tensor-memory columns: H 2 x 64, relu^T | M^T 2 x 64, accumulators 256 (resident over all tiles
// of the CTA; each touched by one issuing thread only: bit-reproducible).  Two consumer groups of eight warps (one per
// hidden half == stage parity), two issuer warps, four producer warps that build the next tile's operands in the
// other half of a double-buffered shared-memory region and prefetch the tile after it into registers.
// Reference: loss.backward() of rnad.py:425 through nn/net.py:37-51.
#include <cuda_fp16.h>

#include "learner_bwd.cuh"

namespace rnad {
namespace tc {

namespace {

#ifndef RNAD_BWDH_SELF_ISSUE
#define RNAD_BWDH_SELF_ISSUE 0
#endif
// 0: two issuer warps issue the MMAs (the default);
// 1: the first warp of each consumer group issues the group's MMAs itself, right behind a named barrier of the group's
//    256 threads, and the issuer warps are not launched: two mbarrier hops per stage fall off the chain, but the
//    leader warp's own stage work now waits for its issue sequences - measured 101 vs 87 us at cfg2, not adopted
constexpr bool kSelfIssue = RNAD_BWDH_SELF_ISSUE != 0;
#ifndef RNAD_BWDH_NAMED_BARRIERS
#define RNAD_BWDH_NAMED_BARRIERS 1
#endif
// consumers -> issuer hand-overs ("H^T loaded", "relu^T | M^T stored"): 1 = hardware named barriers shared by the group's
// eight warps and its issuer warp (288 threads; the issuer is released ~40 cycles after the last consumer warp arrives),
// 0 = mbarriers the issuer polls (an arrival reaches it after 150-250 cycles)
constexpr bool kNamedBarriers = RNAD_BWDH_NAMED_BARRIERS != 0 && !kSelfIssue;
constexpr int kConsumers = 512, kProducers = 128, kIssuers = kSelfIssue ? 0 : 64;
constexpr int kThreadsH = kConsumers + kProducers + kIssuers;

template <int A>
struct PlanH {
    using S = Shape<A>;                                                  // parameter offsets (state_dict order)
    static constexpr int KIN = 2 * A * A;
    static constexpr bool kBiasInK = (KIN % 16) != 0;                    // the bias rides in K as a constant-1 input column
    static constexpr int KP = round_up(KIN + 1, 16);
    static constexpr int kSbo1 = KP * 2 * 8;                             // bytes between 8-row groups of a [rows x KP] fp16 operand
    static constexpr int kTrunkBytes = kHidden * KP * 2;
    static constexpr int kSet = KIN + 1;                                 // columns of one g_a set: g_a x_0 .. g_a x_{KIN-1}, g_a
    static constexpr int kNXV = round_up(kSet, 16);                      // N of the value trunk's dW1 MMA
    static constexpr int kNXP = round_up(A * kSet, 16);                  // ... of the policy trunk's
    static constexpr int kNG = 16;                                       // N of the dW2 MMA (rows of BG: d_v, d_logit[0..A))
    // B operands of the gradient MMAs, K = tile row n, N = operand row c: MN-MAJOR (the tensor core transposes, not the
    // producers) - core matrix = 8 rows n x 16 bytes (8 consecutive c), the n-groups of a c-group adjacent (LBO = 128),
    // c-groups kSboN apart.  A producer thread owns row n and writes eight c per 16-byte store; the eight lanes of a
    // store phase fill one contiguous 128-byte core matrix (conflict-free).
    static constexpr int kLboK = 128;
    static constexpr int kSboN = (kTileM / 8) * kLboK;
    static constexpr int kImageBytes = 2 * kTrunkBytes;                  // both trunks' first layers [256 x KP] fp16, K-major
    static constexpr int kTile = round_up(kImageBytes, 128);
    static constexpr int kX = 0;                                         // within a tile buffer: observation tile [128 x KP] fp16
    static constexpr int kBXV = kX + kTileM * KP * 2;                    // [kNXV x 128]  d_v x | d_v
    static constexpr int kBXP = kBXV + (kNXV / 8) * kSboN;               // [kNXP x 128]  d_logit[a] x | d_logit[a], a = 0..A-1
    static constexpr int kBG = kBXP + (kNXP / 8) * kSboN;                // [16 x 128]    d_v, d_logit[0..A)
    static constexpr int kTileBytes = round_up(kBG + (kNG / 8) * kSboN, 128);
    static constexpr int kRed = kTile + 2 * kTileBytes;
    static constexpr int kBar = kRed + 4 * 32;                           // image, H[2], H free[2], RM ready[2], grad[2], full[2], empty[2]
    static constexpr int kTmem = kBar + 8 * 13 + 8;
    static constexpr int kBytes = kTmem + 16;
    // tensor memory
    static constexpr int kColH = 0, kColRM = 128, kAccV = 256, kAccP = kAccV + 2 * kNXV + 2 * kNG;
    static constexpr bool kFits = kBiasInK && kAccP + 2 * kNXP + 2 * kNG <= 512 && kBytes <= 227 * 1024;
    __host__ __device__ static constexpr int acc(int trunk) { return trunk == 0 ? kAccV : kAccP; }
    __host__ __device__ static constexpr int nx(int trunk) { return trunk == 0 ? kNXV : kNXP; }
};

// element (row, k) of a K-major, no-swizzle [rows x KP] fp16 operand
template <int KP>
__host__ __device__ __forceinline__ uint32_t op_off16(int row, int k) {
    return (uint32_t)((row >> 3) * (KP * 16) + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}
__device__ __forceinline__ uint16_t f16_bits(float x) {      // round to nearest even; beyond +-65504 -> +-65504, not Inf
    uint16_t h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
    return h;
}
// two fp32 -> one word of two fp16, `lo` in bits 0..15 (saturating like f16_bits)
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// D = f32, A and B = f16, M = 128; A K-major, B K-major or (b_mn) MN-major
__host__ __device__ constexpr uint32_t idesc_f16(int n, bool b_mn = false) {
    return (1u << 4) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

#ifdef RNAD_TRACE_BWD
// development aid (scripts/trace_bwd_f16.py): cycle stamps of CTA 0, stages 16..79: [role][stage - 16][event]; role 0 / 1 =
// first warp of consumer group 0 / 1, 2 / 3 = issuer 0 / 1, 4 = producer warp 0 (per tile, at its first stage)
__device__ long long g_bwdh_trace[5][64][8];
#define HTR(role, s, ev) do { if (blockIdx.x == 0 && (s) >= 16 && (s) < 80 && lane32 == 0) g_bwdh_trace[role][(s) - 16][ev] = clock64(); } while (0)
#else
#define HTR(role, s, ev) do { } while (0)
#endif

template <int A>
__global__ void pack_bwd_f16_image_kernel(rnad_mlp_weights w, uint8_t* __restrict__ image) {
    using P = PlanH<A>;
    const int thread = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    for (int e = thread; e < 2 * kHidden * P::KP; e += n_threads) {
        const int trunk = e / (kHidden * P::KP), j = (e / P::KP) % kHidden, k = e % P::KP;
        const float* w1 = trunk == 0 ? w.value_fc0_w : w.policy_fc0_w;
        const float* b1 = trunk == 0 ? w.value_fc0_b : w.policy_fc0_b;
        const float v = k < P::KIN ? w1[j * P::KIN + k] : (k == P::KIN ? b1[j] : 0.f);
        *reinterpret_cast<uint16_t*>(image + trunk * P::kTrunkBytes + op_off16<P::KP>(j, k)) = f16_bits(v);
    }
}

template <int A>
__global__ void __launch_bounds__(kThreadsH, 1) learner_bwd_f16_kernel(const float* __restrict__ obs, int64_t N, int T_split,
                                                                       int64_t B_split, const uint8_t* __restrict__ image,
                                                                       const float* __restrict__ w2v,
                                                                       const float* __restrict__ w2p,
                                                                       const float* __restrict__ d_logit,
                                                                       const float* __restrict__ d_v,
                                                                       float* __restrict__ partials) {
    using P = PlanH<A>;
    using S = Shape<A>;
    constexpr int KIN = P::KIN, KP = P::KP, kSet = P::kSet;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane32 = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // (tells the compiler that the warp index is warp-uniform)
    const uint32_t bar_img = smem_u32(smem + P::kBar);
    auto bar_h = [&](int b) { return bar_img + 8 + 8 * b; };         // recompute into H region b complete
    auto bar_hl = [&](int b) { return bar_img + 24 + 8 * b; };       // the consumers have loaded H region b
    auto bar_c = [&](int b) { return bar_img + 40 + 8 * b; };        // relu^T | M^T of the stage are in RM region b
    auto bar_g = [&](int b) { return bar_img + 56 + 8 * b; };        // the gradient MMAs reading RM region b complete
    auto bar_full = [&](int b) { return bar_img + 72 + 8 * b; };     // tile operands of shared-memory buffer b written
    auto bar_empty = [&](int b) { return bar_img + 88 + 8 * b; };    // every MMA reading shared-memory buffer b complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_h(b), 1);
            mbar_init(bar_hl(b), kConsumers / 64);
            mbar_init(bar_c(b), kConsumers / 64);
            mbar_init(bar_g(b), 1);
            mbar_init(bar_full(b), kProducers / 32);
            mbar_init(bar_empty(b), 2);                              // one commit per issuer
        }
        mbar_fence_init();
        tma_bulk_load(smem, image, P::kImageBytes, bar_img);
    }
    // operand rows that are never written stay zero (both tile buffers)
    for (int i = tid; i < 2 * P::kTileBytes / 4; i += kThreadsH) reinterpret_cast<uint32_t*>(smem + P::kTile)[i] = 0u;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    if (tid < kConsumers) {
        // clear the gradient accumulators: columns [256, 512), 16 at a time, split over the four column parts
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int c = P::kAccV + (warp >> 2) * 16; c < 512; c += 64)
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_base + c),
                "r"(0u)
                : "memory");
        tcp::tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // which tiles this CTA walks (learner_bwd_tc2_kernel's rule: split mode alternates CTAs between the players)
    const int cta = blockIdx.x, n_ctas = gridDim.x;
    const bool split = T_split > 0;
    const int player = split ? (cta & 1) : 0;
    const int64_t tiles_per_t = split ? (B_split + kTileM - 1) / kTileM : 0;
    const int64_t my_first = split ? (cta >> 1) : cta, my_stride = split ? (n_ctas >> 1) : n_ctas;
    const int64_t num_tiles = split ? (int64_t)((T_split - player + 1) / 2) * tiles_per_t : (N + kTileM - 1) / kTileM;
    const int64_t my_tiles = my_first < num_tiles ? (num_tiles - 1 - my_first) / my_stride + 1 : 0;
    const int64_t n_stages = my_tiles * 8;
    float* dst = partials + (int64_t)cta * S::kParams;

    // region / hidden half / stage parity this warp serves: issuer warp b, consumer group b (producers: unused)
    const int b = tid >= kConsumers + kProducers ? warp - (kConsumers + kProducers) / 32 : (warp >> 3) & 1;
    // MMA issue (the issuer warps, or with kSelfIssue the first warp of each consumer group), for region / hidden half b.
    // Everything an MMA takes is computed BEFORE the barrier wait it follows and pinned there (pin()): after the wait
    // only the register -> uniform-register moves and the tcgen05.mma themselves remain - the descriptor arithmetic
    // (a dependent scalar chain of ~60 instructions) used to sit between the arrival and the first MMA of both groups.
    const uint32_t n_st = (uint32_t)n_stages;
    const uint32_t tile0 = smem_u32(smem + P::kTile);
    const uint32_t d_h = tmem_base + P::kColH + (uint32_t)b * 64, rm = tmem_base + P::kColRM + (uint32_t)b * 64;
    const uint64_t w1_desc = tcp::desc_sbo(smem_u32(smem) + b * (128 / 8) * P::kSbo1, P::kSbo1);
    const uint64_t x_desc = tcp::desc_sbo(tile0 + P::kX, P::kSbo1);
    const uint64_t g_desc = desc_lbo_sbo(tile0, P::kLboK, P::kSboN);
    auto hand_over = [](int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kConsumers / 2 + 32) : "memory"); };   // a consumer group + its issuer warp
    auto pin32 = [](uint32_t& v) { asm volatile("" : "+r"(v)::"memory"); };
    auto pin64 = [](uint64_t& v) { asm volatile("" : "+l"(v)::"memory"); };
    uint32_t seen_full = 0xffffffffu;
    auto need_tile = [&](uint32_t k) {      // (whole warp) the producers have written tile k's operands
        if (k != seen_full) {
            tcp::mbar_wait_c(bar_full((int)(k & 1)), (k >> 1) & 1u);
            seen_full = k;
        }
    };
    // first-layer operands of stage s: A = W1 of (trunk, hidden half b), B = the stage's 64 observation rows
    auto recompute_ops = [&](uint32_t s, uint64_t& a, uint64_t& bb) {
        const uint32_t trunk = (s >> 1) & 1u, rh = (s >> 2) & 1u, kb = (s >> 3) & 1u;
        a = w1_desc + (uint64_t)((trunk * P::kTrunkBytes) >> 4);
        bb = x_desc + (uint64_t)((kb * P::kTileBytes + rh * (64 / 8) * P::kSbo1) >> 4);
        pin64(a);
        pin64(bb);
    };
    auto recompute = [&](uint64_t a, uint64_t bb) {     // H^T of a stage into H region b (elected lane)
#pragma unroll
        for (int ks = 0; ks < KP / 16; ++ks) mma_ss_f16(d_h, a + (uint64_t)(ks * 16), bb + (uint64_t)(ks * 16), idesc_f16(64), ks > 0);
        mma_commit(bar_h(b));
    };
    if (tid >= kConsumers + kProducers) {
        // ------------------------------------------------------------ issuers: warp b issues the stages with s & 1 == b
        // stage s: tile k = s >> 3, hidden half = s & 1 (== region == issuer == consumer group), trunk = (s >> 1) & 1,
        // row half = (s >> 2) & 1
        if (!kSelfIssue) tcp::mbar_wait_c(bar_img, 0);
        if (!kSelfIssue && (uint32_t)b < n_st) {               // fill the pipeline: stage b
            uint64_t a, bb;
            recompute_ops((uint32_t)b, a, bb);
            need_tile(0);
            tc_fence_after();
            if (tcp::elect_one()) recompute(a, bb);
            __syncwarp();
        }
#pragma unroll 1
        for (uint32_t s = (uint32_t)b; !kSelfIssue && s < n_st; s += 2) {
            const uint32_t trunk = (s >> 1) & 1u, rh = (s >> 2) & 1u, kb = (s >> 3) & 1u;
            const uint32_t par = (s >> 1) & 1u;
            const bool more = s + 2 < n_st;
            // operands of grad(s): D_w2 += relu^T BG^T, D_w1 += M^T BX^T (K = the stage's 64 rows, 16 per MMA)
            const uint32_t t_off = kb * P::kTileBytes + rh * 8 * P::kLboK;
            uint64_t bx = g_desc + (uint64_t)((t_off + (trunk == 0 ? P::kBXV : P::kBXP)) >> 4);
            uint64_t bg = g_desc + (uint64_t)((t_off + P::kBG) >> 4);
            const uint32_t nx = trunk == 0 ? P::kNXV : P::kNXP;
            uint32_t acc1 = tmem_base + (trunk == 0 ? P::kAccV : P::kAccP) + (uint32_t)b * nx;
            uint32_t acc2 = tmem_base + (trunk == 0 ? P::kAccV : P::kAccP) + 2 * nx + (uint32_t)b * P::kNG;
            uint32_t idx = trunk == 0 ? idesc_f16(P::kNXV, true) : idesc_f16(P::kNXP, true);
            const uint32_t empty_bar = bar_empty((int)kb);
            pin64(bx);
            pin64(bg);
            pin32(acc1);
            pin32(acc2);
            pin32(idx);
            uint64_t a2 = 0, b2 = 0;
            if (more) recompute_ops(s + 2, a2, b2);
            HTR(2 + b, s, 0);
            if (more) {
                need_tile((s + 2) >> 3);
                if (kNamedBarriers) hand_over(1 + b);
                else tcp::mbar_wait_c(bar_hl(b), par);                   // H^T of stage s is in the consumers' registers
                tc_fence_after();
                HTR(2 + b, s, 1);
                if (tcp::elect_one()) recompute(a2, b2);
                __syncwarp();
                HTR(2 + b, s, 2);
            }
            if (kNamedBarriers) {
                if (!more) hand_over(1 + b);                             // (every stage has both hand-overs, also the last ones)
                hand_over(3 + b);
            } else {
                tcp::mbar_wait_c(bar_c(b), par);                         // relu^T | M^T of stage s are in RM region b
            }
            tc_fence_after();
            HTR(2 + b, s, 3);
            if (tcp::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_ts_f16(acc2, rm + ks * 8, bg + (uint64_t)((ks * 2 * P::kLboK) >> 4), idesc_f16(P::kNG, true), true);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_ts_f16(acc1, rm + 32 + ks * 8, bx + (uint64_t)((ks * 2 * P::kLboK) >> 4), idx, true);
                if ((s & 7) >= 6) mma_commit(empty_bar);                 // this issuer's last reads of the tile's shared-memory operands
                mma_commit(bar_g(b));
            }
            __syncwarp();
            HTR(2 + b, s, 4);
        }
    } else if (tid >= kConsumers) {
        // ------------------------------------------------------------ producers: one thread per tile row
        const int n = tid - kConsumers, pw = n >> 5;
        float gsum[1 + A];
#pragma unroll
        for (int a = 0; a <= A; ++a) gsum[a] = 0.f;
        float x[KIN], g[1 + A];
        auto fetch = [&](int64_t k) {          // the row's observation and gradient signal -> registers
            const int64_t u = my_first + k * my_stride;
            int64_t row = u * kTileM + n;
            bool active = row < N;
            if (split) {
                const int64_t tt = 2 * (u / tiles_per_t) + player, j = (u % tiles_per_t) * kTileM + n;
                row = tt * B_split + j;
                active = j < B_split;
            }
            load_row<KIN>(obs, active ? row : 0, active, x);
            g[0] = active ? __ldg(d_v + row) : 0.f;
#pragma unroll
            for (int a = 0; a < A; ++a) g[1 + a] = active ? __ldg(d_logit + row * A + a) : 0.f;
        };
        if (my_tiles > 0) fetch(0);
        for (int64_t k = 0; k < my_tiles; ++k) {
            float xc[KIN], gc[1 + A];
#pragma unroll
            for (int i = 0; i < KIN; ++i) xc[i] = x[i];
#pragma unroll
            for (int a = 0; a <= A; ++a) gc[a] = g[a];
            if (k + 1 < my_tiles) fetch(k + 1);                         // in flight while this tile is written
            const int tb = (int)(k & 1);
            if (pw == 0) HTR(4, 8 * k, 0);
            if (k >= 2) tcp::mbar_wait_c(bar_empty(tb), (uint32_t)((k >> 1) - 1) & 1u);   // the MMAs of tile k - 2 are done with it
            if (pw == 0) HTR(4, 8 * k, 1);
            uint8_t* tile = smem + P::kTile + tb * P::kTileBytes;
            // observation row [x | 1 | 0 ..] as fp16, 8 K elements per 16-byte store
#pragma unroll
            for (int q = 0; q < KP / 8; ++q) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k0 = 8 * q + 2 * u, k1 = k0 + 1;
                    const float lo = k0 < KIN ? xc[k0 < KIN ? k0 : 0] : (k0 == KIN ? 1.f : 0.f);
                    const float hi = k1 < KIN ? xc[k1 < KIN ? k1 : 0] : (k1 == KIN ? 1.f : 0.f);
                    v[u] = pack2(lo, hi);
                }
                *reinterpret_cast<uint4*>(tile + P::kX + op_off16<KP>(n, 8 * q)) = make_uint4(v[0], v[1], v[2], v[3]);
            }
            // the masks' B operands: the row's gradient signal times its observation, and the signal itself (-> db1);
            // eight consecutive operand rows c per 16-byte store
            const int row_off = (n >> 3) * P::kLboK + (n & 7) * 16;
            auto val_v = [&](int c) { return c < KIN ? gc[0] * xc[c < KIN ? c : 0] : (c == KIN ? gc[0] : 0.f); };
            auto val_p = [&](int c) {
                const int a = c / kSet, kk = c % kSet;
                return a < A ? (kk < KIN ? gc[1 + (a < A ? a : 0)] * xc[kk < KIN ? kk : 0] : gc[1 + (a < A ? a : 0)]) : 0.f;
            };
#pragma unroll
            for (int cg = 0; cg < (kSet + 7) / 8; ++cg)
                *reinterpret_cast<uint4*>(tile + P::kBXV + cg * P::kSboN + row_off) =
                    make_uint4(pack2(val_v(8 * cg), val_v(8 * cg + 1)), pack2(val_v(8 * cg + 2), val_v(8 * cg + 3)),
                               pack2(val_v(8 * cg + 4), val_v(8 * cg + 5)), pack2(val_v(8 * cg + 6), val_v(8 * cg + 7)));
#pragma unroll
            for (int cg = 0; cg < (A * kSet + 7) / 8; ++cg)
                *reinterpret_cast<uint4*>(tile + P::kBXP + cg * P::kSboN + row_off) =
                    make_uint4(pack2(val_p(8 * cg), val_p(8 * cg + 1)), pack2(val_p(8 * cg + 2), val_p(8 * cg + 3)),
                               pack2(val_p(8 * cg + 4), val_p(8 * cg + 5)), pack2(val_p(8 * cg + 6), val_p(8 * cg + 7)));
            {
                float g8[8];
#pragma unroll
                for (int a = 0; a < 8; ++a) g8[a] = a <= A ? gc[a <= A ? a : 0] : 0.f;
                *reinterpret_cast<uint4*>(tile + P::kBG + row_off) =
                    make_uint4(pack2(g8[0], g8[1]), pack2(g8[2], g8[3]), pack2(g8[4], g8[5]), pack2(g8[6], g8[7]));
            }
#pragma unroll
            for (int a = 0; a <= A; ++a) gsum[a] += gc[a];
            fence_async_smem();
            __syncwarp();
            if (lane32 == 0) tcp::mbar_arrive(bar_full(tb));
            if (pw == 0) HTR(4, 8 * k, 2);
        }
        // output-bias gradients: sums of g over the CTA's rows
        float* s_red = reinterpret_cast<float*>(smem + P::kRed);
#pragma unroll
        for (int a = 0; a <= A; ++a) {
            const float v = warp_sum(gsum[a]);
            if (lane32 == 0) s_red[pw * 8 + a] = v;
        }
    } else {
        // ------------------------------------------------------------ consumers: thread = hidden unit x 32 rows of a stage
        const int quad = warp & 3, cpart = warp >> 2;             // TMEM lane quadrant (hidden unit); (trunk, half) of the final read-out
        // (b = warp >> 3: group == region == stage parity)
        const bool leader = kSelfIssue && (warp & 7) == 0;          // issues the group's MMAs
        auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + b), "n"(kConsumers / 2) : "memory"); };
        tcp::mbar_wait_c(bar_img, 0);
        if (leader && (uint32_t)b < n_st) {                        // fill the pipeline: stage b
            uint64_t a, bb;
            recompute_ops((uint32_t)b, a, bb);
            need_tile(0);
            tc_fence_after();
            if (tcp::elect_one()) recompute(a, bb);
            __syncwarp();
        }
        const int cw = cpart & 1;                                  // which 32 of the stage's 64 rows
        const int j_local = quad * 32 + lane32;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t th = tmem_lane + P::kColH + b * 64 + cw * 32;
        const uint32_t trm = tmem_lane + P::kColRM + b * 64 + cw * 16;
        const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll 1
        for (uint32_t s = (uint32_t)b; s < n_st; s += 2) {
            const uint32_t i = s >> 1;
            const bool tr = (warp & 7) == 0;
            const bool more = s + 2 < n_st;
            // (leader) operands of grad(s) and recompute(s + 2), ahead of the waits
            const uint32_t trunk = (s >> 1) & 1u, rh = (s >> 2) & 1u, kb = (s >> 3) & 1u;
            const uint32_t t_off = kb * P::kTileBytes + rh * 8 * P::kLboK;
            uint64_t bx = g_desc + (uint64_t)((t_off + (trunk == 0 ? P::kBXV : P::kBXP)) >> 4);
            uint64_t bg = g_desc + (uint64_t)((t_off + P::kBG) >> 4);
            const uint32_t nx = trunk == 0 ? P::kNXV : P::kNXP;
            uint32_t acc1 = tmem_base + (trunk == 0 ? P::kAccV : P::kAccP) + (uint32_t)b * nx;
            uint32_t acc2 = tmem_base + (trunk == 0 ? P::kAccV : P::kAccP) + 2 * nx + (uint32_t)b * P::kNG;
            uint32_t idx = trunk == 0 ? idesc_f16(P::kNXV, true) : idesc_f16(P::kNXP, true);
            const uint32_t empty_bar = bar_empty((int)kb);
            uint64_t a2 = 0, b2 = 0;
            if (leader) {
                pin64(bx);
                pin64(bg);
                pin32(acc1);
                pin32(acc2);
                pin32(idx);
                if (more) recompute_ops(s + 2, a2, b2);
            }
            if (tr) HTR(b, s, 0);
            tcp::mbar_wait_c(bar_h(b), i & 1u);
            tc_fence_after();
            if (tr) HTR(b, s, 1);
            uint32_t hr[32];
            tmem_ld16(th, hr);
            tmem_ld16(th + 16, hr + 16);
            tmem_ld_wait();
            if (tr) HTR(b, s, 2);
            tc_fence_before();
            if (kSelfIssue) {
                group_sync();                                       // H^T of stage s is in the group's registers
                if (leader && more) {
                    need_tile((s + 2) >> 3);
                    tc_fence_after();
                    if (tcp::elect_one()) recompute(a2, b2);
                    __syncwarp();
                }
            } else if (kNamedBarriers) {
                hand_over(1 + b);                                   // the issuer may recompute stage s + 2 into the region
            } else {
                __syncwarp();
                if (lane32 == 0) tcp::mbar_arrive(bar_hl(b));
            }
            // ---- rows (2c, 2c + 1) -> one packed column: relu^T, and next to it the 0/1 mask (of the ROUNDED value: a
            // hidden unit whose positive pre-activation rounds to zero in fp16 counts as switched off in both)
            uint32_t re[16], mk[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                re[c] = pack2_relu(__uint_as_float(hr[2 * c]), __uint_as_float(hr[2 * c + 1]));
                const __half2 m = __hgt2(*reinterpret_cast<const __half2*>(&re[c]), zero2);
                mk[c] = *reinterpret_cast<const uint32_t*>(&m);
            }
            if (i >= 1) {                                           // grad(s - 2) has read the region
                tcp::mbar_wait_c(bar_g(b), (i - 1) & 1u);
                tc_fence_after();
            }
            if (tr) HTR(b, s, 3);
            tmem_st16(trm, re);
            tmem_st16(trm + 32, mk);
            tcp::tmem_st_wait();
            if (tr) HTR(b, s, 4);
            tc_fence_before();
            if (kSelfIssue) {
                group_sync();                                       // relu^T | M^T of stage s are in RM region b
                if (leader) {
                    tc_fence_after();
                    if (tcp::elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_ts_f16(acc2, rm + ks * 8, bg + (uint64_t)((ks * 2 * P::kLboK) >> 4), idesc_f16(P::kNG, true), true);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_ts_f16(acc1, rm + 32 + ks * 8, bx + (uint64_t)((ks * 2 * P::kLboK) >> 4), idx, true);
                        if ((s & 7) >= 6) mma_commit(empty_bar);    // this group's last reads of the tile's shared-memory operands
                        mma_commit(bar_g(b));
                    }
                    __syncwarp();
                }
            } else if (kNamedBarriers) {
                hand_over(3 + b);                                   // relu^T | M^T of stage s are in RM region b
            } else {
                __syncwarp();
                if (lane32 == 0) tcp::mbar_arrive(bar_c(b));
            }
            if (tr) HTR(b, s, 5);
        }
        // every gradient MMA complete: the last commit of the group's issuer; then the groups meet
        if (n_stages >= 2) tcp::mbar_wait_c(bar_g(b), (uint32_t)((n_stages >> 1) - 1) & 1u);
        tc_fence_before();
        asm volatile("bar.sync 5, %0;" ::"n"(kConsumers) : "memory");
        tc_fence_after();

        // ---- this CTA's partial gradient, flat in state_dict order: column part c reads (trunk, half) = (c >> 1, c & 1);
        // the second-layer weights (fp32) enter here: dW1[j, k] = sum_a W2[a, j] D[j, a * kSet + k], a ascending
        const int half = cpart & 1;
        const int j = half * 128 + j_local;
        if ((cpart >> 1) == 0) {
            const uint32_t acc = tmem_lane + P::kAccV;
            uint32_t w[P::kNXV];
#pragma unroll
            for (int q = 0; q < P::kNXV / 16; ++q) tmem_ld16(acc + half * P::kNXV + q * 16, w + q * 16);
            tmem_ld_wait();
            const float w2 = __ldg(w2v + j);
            float* w1_dst = dst + S::kOffV0w + j * KIN;
#pragma unroll
            for (int kk = 0; kk < KIN; ++kk) w1_dst[kk] = w2 * __uint_as_float(w[kk]);
            dst[S::kOffV0b + j] = w2 * __uint_as_float(w[KIN]);
            uint32_t d2[16];
            tmem_ld16(acc + 2 * P::kNXV + half * P::kNG, d2);
            tmem_ld_wait();
            dst[S::kOffV1w + j] = __uint_as_float(d2[0]);
        } else {
            const uint32_t acc = tmem_lane + P::kAccP;
            uint32_t w[P::kNXP];
#pragma unroll
            for (int q = 0; q < P::kNXP / 16; ++q) tmem_ld16(acc + half * P::kNXP + q * 16, w + q * 16);
            tmem_ld_wait();
            float w2[A];
#pragma unroll
            for (int a = 0; a < A; ++a) w2[a] = __ldg(w2p + a * kHidden + j);
            float* w1_dst = dst + S::kOffP0w + j * KIN;
#pragma unroll
            for (int kk = 0; kk <= KIN; ++kk) {
                float v = w2[0] * __uint_as_float(w[kk]);
#pragma unroll
                for (int a = 1; a < A; ++a) v = fmaf(w2[a], __uint_as_float(w[a * kSet + kk]), v);
                if (kk < KIN) w1_dst[kk] = v;
                else dst[S::kOffP0b + j] = v;
            }
            uint32_t d2[16];
            tmem_ld16(acc + 2 * P::kNXP + half * P::kNG, d2);
            tmem_ld_wait();
#pragma unroll
            for (int a = 0; a < A; ++a) dst[S::kOffP1w + a * kHidden + j] = __uint_as_float(d2[1 + a]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid <= A) {
        const float* s_red = reinterpret_cast<const float*>(smem + P::kRed);
        const float v = (s_red[0 * 8 + tid] + s_red[1 * 8 + tid]) + (s_red[2 * 8 + tid] + s_red[3 * 8 + tid]);
        if (tid == 0) dst[S::kOffV1b] = v;
        else dst[S::kOffP1b + tid - 1] = v;
    }
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int A>
int launch_h(const float* obs, int64_t N, int T_split, int64_t B_split, const rnad_mlp_weights& w, const float* d_logit,
             const float* d_v, uint8_t* image, float* partials, int blocks, cudaStream_t st, int mode) {
    using P = PlanH<A>;
    if constexpr (!P::kFits) {
        set_error("learner_backward_f16: max_actions = %d is not served by the fp16 backward", A);
        return RNAD_EUNSUPPORTED;
    } else {
        if (mode != 1) {
            pack_bwd_f16_image_kernel<A><<<16, 256, 0, st>>>(w, image);
            RNAD_CHECK_LAUNCH("pack_bwd_f16_image_kernel");
            if (mode == 2) return RNAD_OK;
        }
        // one CTA per SM (all 512 tensor-memory columns): more than half of the shared memory keeps a second one out
        const size_t smem = P::kBytes > 116 * 1024 ? P::kBytes : 116 * 1024;
        int rc = check_cuda(cudaFuncSetAttribute(learner_bwd_f16_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(learner_bwd_f16)");
        if (rc) return rc;
        rc = check_cuda(cudaFuncSetAttribute(learner_bwd_f16_kernel<A>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             cudaSharedmemCarveoutMaxShared),
                        "cudaFuncSetAttribute(learner_bwd_f16)");
        if (rc) return rc;
        learner_bwd_f16_kernel<A><<<blocks, kThreadsH, smem, st>>>(obs, N, T_split, B_split, image, w.value_fc1_w,
                                                                  w.policy_fc1_w, d_logit, d_v, partials);
        RNAD_CHECK_LAUNCH("learner_bwd_f16_kernel");
        return RNAD_OK;
    }
}

}  // namespace

#ifdef RNAD_TRACE_BWD
}  // namespace tc
}  // namespace rnad
extern "C" __attribute__((visibility("default"))) int rnad_debug_bwdh_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, rnad::tc::g_bwdh_trace, sizeof(rnad::tc::g_bwdh_trace));
}
namespace rnad {
namespace tc {
#endif

bool learner_backward_f16_supported(int A) {
    switch (A) {
        case 2: return PlanH<2>::kFits;
        case 3: return PlanH<3>::kFits;
        case 4: return PlanH<4>::kFits;
    }
    return false;
}

int64_t learner_backward_f16_image_bytes(int A) {
    switch (A) {
        case 2: return PlanH<2>::kImageBytes;
        case 3: return PlanH<3>::kImageBytes;
        case 4: return PlanH<4>::kImageBytes;
    }
    return 0;
}

int learner_backward_f16(int A, const float* obs, int64_t N, int T_split, int64_t B_split, const rnad_mlp_weights& w,
                         const float* d_logit, const float* d_v, uint8_t* image, float* partials, int blocks,
                         cudaStream_t st, int mode) {
    switch (A) {
        case 2: return launch_h<2>(obs, N, T_split, B_split, w, d_logit, d_v, image, partials, blocks, st, mode);
        case 3: return launch_h<3>(obs, N, T_split, B_split, w, d_logit, d_v, image, partials, blocks, st, mode);
        case 4: return launch_h<4>(obs, N, T_split, B_split, w, d_logit, d_v, image, partials, blocks, st, mode);
    }
    return RNAD_EINVAL;
}

}  // namespace tc
}  // namespace rnad
