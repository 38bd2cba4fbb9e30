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
// What fp16 buys beyond K = 16 per MMA: a 64-row stage's relu^T and M^T are 32 + 32 packed columns - exactly the 64
// columns its fp32 H^T took, so a stage lives IN PLACE in one 64-column region and tensor memory holds FOUR of them
// next to 256 columns of accumulators.  The kernels are bound by the latency of the loop
//     recompute(s) -> [commit] -> consumers: tcgen05.ld, relu / mask, pack, tcgen05.st -> [named barrier] -> issuer:
//     grad(s), recompute(s + 4) -> ...
// (~1,200 cycles of barrier hops and tensor-memory round trips, little of it work), so what counts is how many rows
// are inside that loop at a time: four regions x 64 rows instead of the two stage buffers of the tf32 kernels.  Region
// r = s & 3 = (trunk, hidden half) serves exactly one accumulator pair and has its own consumer group (four warps, one
// per lane quadrant, two passes of 32 rows each: H^T columns [32p, 32p + 32) -> relu^T at [32p, 32p + 16), M^T at
// [32p + 16, 32p + 32), so a pass only overwrites what it has loaded) and its own issuer warp (every accumulator is
// touched by one issuing thread only, whose MMAs execute in issue order: bit-reproducible, and recompute(s + 4) may
// follow grad(s) without a barrier).  ONE CTA per SM, 768 threads: 16 consumer warps, 4 issuer warps, 4 producer
// warps that build the next tile's operands in the other half of a double-buffered shared-memory region and
// prefetch the tile after it into registers.
// [History at cfg2: tf32, S^T formulation 128 us -> tf32 mask formulation 106 -> fp16 with H^T and relu^T | M^T in
//  separate regions, two stages in flight (recompute issued as soon as a stage is loaded) 87 -> named-barrier
//  hand-overs 84.6 -> four in-place regions 64 -> consumers arrive / issuer syncs 60 (bench.py's per-call time, which
//  includes the 5 us reduction of the per-CTA partials; the kernel alone: 53 us under ncu, tensor pipe 53 % active).]
// Reference: loss.backward() of rnad.py:425 through nn/net.py:37-51.
#include <cuda_fp16.h>

#include "learner_bwd.cuh"

namespace rnad {
namespace tc {

namespace {

constexpr int kRegions = 4;                       // stages in flight = (trunk, hidden half) = accumulator pairs
constexpr int kConsumers = 512, kProducers = 128, kIssuers = 32 * kRegions;
constexpr int kThreadsH = kConsumers + kProducers + kIssuers;
constexpr int kGroupThreads = kConsumers / kRegions + 32;   // a consumer group and its issuer warp (named barrier)

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
    static constexpr int kBar = kRed + 4 * 32;                           // image, H[4], full[2], empty[2]
    static constexpr int kTmem = kBar + 8 * 9 + 8;
    static constexpr int kBytes = kTmem + 16;
    // tensor memory
    static constexpr int kColStage = 0, kAccV = 256, kAccP = kAccV + 2 * kNXV + 2 * kNG;       // four 64-column stage regions, then the accumulators
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
// development aid (scripts/trace_bwd_f16.py): cycle stamps of CTA 0, stages 16..79: [role][stage - 16][event]; role 0 =
// first warp of consumer group 0 (its stages only), 2 = the stage's issuer, 4 = producer warp 0 (per tile, at its first stage)
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
    auto bar_h = [&](int r) { return bar_img + 8 + 8 * r; };         // recompute into stage region r complete (and every MMA its issuer issued before)
    auto bar_full = [&](int b) { return bar_img + 40 + 8 * b; };     // tile operands of shared-memory buffer b written
    auto bar_empty = [&](int b) { return bar_img + 56 + 8 * b; };    // every MMA reading shared-memory buffer b complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        for (int r = 0; r < kRegions; ++r) mbar_init(bar_h(r), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_full(b), kProducers / 32);
            mbar_init(bar_empty(b), kRegions);                       // one commit per issuer
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

    // stage s of a CTA's stream: tile k = s >> 3; region r = s & 3 = (trunk << 1 | hidden half) - also the issuer warp, the
    // consumer group and the accumulator pair of the stage; row half = (s >> 2) & 1.  A region sees the stages r, r + 4, ...
    const uint32_t n_st = (uint32_t)n_stages;
    // consumer group r -> issuer r: a hardware named barrier of kGroupThreads threads; the consumers ARRIVE (they do not
    // wait for the issuer), the issuer SYNCs.  (.aligned: the warp must be converged - the compiler does not know that
    // about inline PTX, and a polling loop may have left it diverged, hence the __syncwarp.)  The barrier is free for the
    // group's next stage in time: that stage's H^T only exists once the issuer is past this sync.
    auto hand_over_arrive = [](int r) {
        __syncwarp();
        asm volatile("bar.arrive %0, %1;" ::"r"(1 + r), "n"(kGroupThreads) : "memory");
    };
    auto hand_over_sync = [](int r) {
        __syncwarp();
        asm volatile("bar.sync %0, %1;" ::"r"(1 + r), "n"(kGroupThreads) : "memory");
    };

    if (tid >= kConsumers + kProducers) {
        // ------------------------------------------------------------ issuers: warp r issues the stages with s & 3 == r
        // Everything an MMA takes is computed BEFORE the hand-over it follows and pinned there (pin()): behind it only the
        // register -> uniform-register moves and the tcgen05.mma themselves remain.
        const int r = warp - (kConsumers + kProducers) / 32;
        const uint32_t trunk = (uint32_t)r >> 1, half = (uint32_t)r & 1u;
        tcp::mbar_wait_c(bar_img, 0);
        const uint32_t tile0 = smem_u32(smem + P::kTile);
        const uint32_t region = tmem_base + P::kColStage + (uint32_t)r * 64;
        const uint64_t w1_desc = tcp::desc_sbo(smem_u32(smem) + trunk * P::kTrunkBytes + half * (128 / 8) * P::kSbo1, P::kSbo1);
        const uint64_t x_desc = tcp::desc_sbo(tile0 + P::kX, P::kSbo1);
        const uint64_t g_desc = desc_lbo_sbo(tile0, P::kLboK, P::kSboN);
        const uint32_t nx = trunk == 0 ? P::kNXV : P::kNXP;
        const uint32_t acc1 = tmem_base + (trunk == 0 ? P::kAccV : P::kAccP) + half * nx;                 // D_w1 of (trunk, half)
        const uint32_t acc2 = tmem_base + (trunk == 0 ? P::kAccV : P::kAccP) + 2 * nx + half * P::kNG;    // D_w2
        const uint32_t idx = trunk == 0 ? idesc_f16(P::kNXV, true) : idesc_f16(P::kNXP, true);
        auto pin64 = [](uint64_t& v) { asm volatile("" : "+l"(v)::"memory"); };
        uint32_t seen_full = 0xffffffffu;
        auto need_tile = [&](uint32_t k) {      // (whole warp) the producers have written tile k's operands
            if (k != seen_full) {
                tcp::mbar_wait_c(bar_full((int)(k & 1)), (k >> 1) & 1u);
                seen_full = k;
            }
        };
        // B operand of the first layers of stage s: the stage's 64 observation rows
        auto x_of = [&](uint32_t s) {
            const uint32_t rh = (s >> 2) & 1u, kb = (s >> 3) & 1u;
            uint64_t d = x_desc + (uint64_t)((kb * P::kTileBytes + rh * (64 / 8) * P::kSbo1) >> 4);
            pin64(d);
            return d;
        };
        auto recompute = [&](uint64_t xb) {     // H^T of a stage into the region (elected lane)
#pragma unroll
            for (int ks = 0; ks < KP / 16; ++ks) mma_ss_f16(region, w1_desc + (uint64_t)(ks * 16), xb + (uint64_t)(ks * 16), idesc_f16(64), ks > 0);
            mma_commit(bar_h(r));
        };
        if ((uint32_t)r < n_st) {               // fill the pipeline: stage r
            const uint64_t xb = x_of((uint32_t)r);
            need_tile(0);
            tc_fence_after();
            if (tcp::elect_one()) recompute(xb);
            __syncwarp();
        }
#pragma unroll 1
        for (uint32_t s = (uint32_t)r; s < n_st; s += kRegions) {
            const uint32_t rh = (s >> 2) & 1u, kb = (s >> 3) & 1u;
            const bool more = s + kRegions < n_st;
            // operands of grad(s): D_w2 += relu^T BG^T, D_w1 += M^T BX^T (K = the stage's 64 rows, 16 per MMA)
            const uint32_t t_off = kb * P::kTileBytes + rh * 8 * P::kLboK;
            uint64_t bx = g_desc + (uint64_t)((t_off + (trunk == 0 ? P::kBXV : P::kBXP)) >> 4);
            uint64_t bg = g_desc + (uint64_t)((t_off + P::kBG) >> 4);
            const uint32_t empty_bar = bar_empty((int)kb);
            pin64(bx);
            pin64(bg);
            uint64_t xb = 0;
            if (more) {
                xb = x_of(s + kRegions);
                need_tile((s + kRegions) >> 3);
            }
            HTR(2, s, 0);
            hand_over_sync(r);                                          // relu^T | M^T of stage s are in the region
            tc_fence_after();
            HTR(2, s, 1);
            if (tcp::elect_one()) {
                // rows 32p .. 32p + 31 of the stage sit at columns [32p, 32p + 16) (relu^T) and [32p + 16, 32p + 32) (M^T)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_ts_f16(acc2, region + (ks >> 1) * 32 + (ks & 1) * 8, bg + (uint64_t)((ks * 2 * P::kLboK) >> 4), idesc_f16(P::kNG, true), true);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_ts_f16(acc1, region + (ks >> 1) * 32 + 16 + (ks & 1) * 8, bx + (uint64_t)((ks * 2 * P::kLboK) >> 4), idx, true);
                if ((s & 7) >= 4) mma_commit(empty_bar);                // this issuer's last reads of the tile's shared-memory operands
                if (more) recompute(xb);                                // (in order behind grad(s), which reads the region)
                else mma_commit(bar_h(r));                              // the last gradient MMAs of the region
            }
            __syncwarp();
            HTR(2, s, 2);
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
        // ------------------------------------------------------------ consumers: group r = warp >> 2 serves region r
        // thread = hidden unit (TMEM lane) x the 64 rows of a stage, 32 at a time
        const int quad = warp & 3, cpart = warp >> 2;             // TMEM lane quadrant (hidden unit); region == (trunk, half) of the final read-out
        const int r = cpart;
        const int j_local = quad * 32 + lane32;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t region = tmem_lane + P::kColStage + (uint32_t)r * 64;
        const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll 1
        for (uint32_t s = (uint32_t)r; s < n_st; s += kRegions) {
            const uint32_t i = s >> 2;
            const bool tr = quad == 0 && r == 0;
            if (tr) HTR(0, s, 0);
            tcp::mbar_wait_c(bar_h(r), i & 1u);                     // H^T of stage s is in the region (and grad(s - 4) has read it)
            tc_fence_after();
            if (tr) HTR(0, s, 1);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                uint32_t hr[32];
                tmem_ld16(region + pass * 32, hr);
                tmem_ld16(region + pass * 32 + 16, hr + 16);
                tmem_ld_wait();
                // rows (2c, 2c + 1) -> one packed column: relu^T, and next to it the 0/1 mask (of the ROUNDED value: a
                // hidden unit whose positive pre-activation rounds to zero in fp16 counts as switched off in both)
                uint32_t re[16], mk[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    re[c] = pack2_relu(__uint_as_float(hr[2 * c]), __uint_as_float(hr[2 * c + 1]));
                    const __half2 m = __hgt2(*reinterpret_cast<const __half2*>(&re[c]), zero2);
                    mk[c] = *reinterpret_cast<const uint32_t*>(&m);
                }
                tmem_st16(region + pass * 32, re);
                tmem_st16(region + pass * 32 + 16, mk);
            }
            tcp::tmem_st_wait();
            if (tr) HTR(0, s, 2);
            tc_fence_before();
            hand_over_arrive(r);                                    // -> grad(s), recompute(s + 4)
            if (tr) HTR(0, s, 3);
        }
        // every gradient MMA of the region complete: its issuer's last commit; then the groups meet
        if (n_st >= (uint32_t)kRegions) tcp::mbar_wait_c(bar_h(r), (n_st >> 2) & 1u);
        tc_fence_before();
        __syncwarp();
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
