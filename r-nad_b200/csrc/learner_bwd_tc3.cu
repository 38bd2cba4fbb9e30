// Learner backward, third cut ("mask" formulation) - the default for max_actions <= 3.
//
// learner_bwd_tc2_kernel recomputes H^T and S^T = (g W2)^T per stage and lets the CUDA cores turn them into relu^T and
// dh^T = S^T * [H^T > 0]: two tensor-memory loads, two stores and four instructions per (hidden unit, row).  But
//     dW1[j, k] = sum_n [h_jn > 0] (sum_a W2[a, j] g_a[n]) x_k[n] = sum_a W2[a, j] * ( sum_n M[j, n] (g_a[n] x_k[n]) )
// so the first-layer gradient is the 0/1 MASK matrix M^T (hidden unit x row) times a B operand that the producer
// warps can build from global data alone: per row n the products g_a[n] * x_k[n] and g_a[n] itself (for db1), the
// A sets of KIN + 1 columns side by side (N = A (KIN + 1) for the policy trunk, KIN + 1 for the value trunk).  The
// second-layer weights come in ONCE per CTA, in fp32, when the accumulators leave tensor memory.  Per stage the
// consumers now read H^T only and write relu^T and M^T (2 instructions per element, both exact: max(h, 0) and
// sat(h * 2^127)), the S^T MMA is gone, and no product is rounded to tf32 twice (g W2 used to be).
//
// Everything else is learner_bwd_tc2_kernel's pipeline: ONE CTA per SM, all 512 tensor-memory columns, two stage
// buffers, 16 consumer warps + 2 issuer warps + 4 producer warps, accumulators resident over all tiles of the CTA,
// every accumulator touched by one issuing thread only (bit-reproducible).  Reference: loss.backward() of
// rnad.py:425 through nn/net.py:37-51.
#include "learner_bwd.cuh"

namespace rnad {
namespace tc {

namespace {

constexpr int kConsumers = 512, kProducers = 128, kIssuers = 64;
constexpr int kThreads3 = kConsumers + kProducers + kIssuers;
#ifndef RNAD_BWD3_GROUPS
#define RNAD_BWD3_GROUPS 2
#endif
constexpr int kGroups = RNAD_BWD3_GROUPS;   // consumer groups: 2 = one per stage buffer, 1 = all sixteen warps on every stage
static_assert(kGroups == 1 || kGroups == 2, "one or two consumer groups");

template <int A>
struct Plan3 : Shape<A> {
    using S = Shape<A>;
    static constexpr int kSet = S::KIN + 1;                              // columns of one g_a set: g_a x_0 .. g_a x_{KIN-1}, g_a
    static constexpr int kNXV = round_up(kSet, 16);                      // N of the value trunk's dW1 MMA
    static constexpr int kNXP = round_up(A * kSet, 16);                  // ... of the policy trunk's
    static constexpr int kNG = 16;                                       // N of the dW2 MMA (rows of BG: d_v, d_logit[0..A))
    static constexpr int kLbo = 144;                                     // as in BwdTcPlan: conflict-free transposing stores
    static constexpr int kSboT = (kTileM / 4) * kLbo;
    // the image of pack_bwd_tc_image_kernel: both trunks' first layers [2][256 x KP] tf32, then their biases [2][256] f32
    static constexpr int kSW1 = 0, kSB1 = 2 * S::kTrunkBytes;
    static constexpr int kImageLoad = kSB1 + 2 * kHidden * 4;
    static constexpr int kTile = round_up(kImageLoad, 128);
    static constexpr int kX = 0;                                         // within a tile buffer: observation tile [128 x KP]
    static constexpr int kBXV = kX + kTileM * S::KP * 4;                 // [kNXV x 128]  d_v x | d_v
    static constexpr int kBXP = kBXV + (kNXV / 8) * kSboT;               // [kNXP x 128]  d_logit[a] x | d_logit[a], a = 0..A-1
    static constexpr int kBG = kBXP + (kNXP / 8) * kSboT;                // [16 x 128]    d_v, d_logit[0..A)
    static constexpr int kTileBytes = round_up(kBG + (kNG / 8) * kSboT, 128);
    static constexpr int kRed = kTile + 2 * kTileBytes;
    static constexpr int kBar = kRed + 4 * 32;
    static constexpr int kTmem = kBar + 80;
    static constexpr int kBytes = kTmem + 16;
    // tensor memory: two stage buffers [H^T -> relu^T 64 | M^T 64], then the value trunk's accumulators
    // [D_w1 half 0 | half 1 | D_w2 half 0 | half 1], then the policy trunk's
    static constexpr int kAccV = 256, kAccP = kAccV + 2 * kNXV + 2 * kNG;
    static constexpr bool kFits = kAccP + 2 * kNXP + 2 * kNG <= 512 && kBytes <= 227 * 1024;
    __host__ __device__ static constexpr int acc(int trunk) { return trunk == 0 ? kAccV : kAccP; }
    __host__ __device__ static constexpr int nx(int trunk) { return trunk == 0 ? kNXV : kNXP; }
};

template <int A>
__global__ void __launch_bounds__(kThreads3, 1) learner_bwd_tc3_kernel(const float* __restrict__ obs, int64_t N, int T_split,
                                                                       int64_t B_split, const uint8_t* __restrict__ image,
                                                                       const float* __restrict__ w2v,
                                                                       const float* __restrict__ w2p,
                                                                       const float* __restrict__ d_logit,
                                                                       const float* __restrict__ d_v,
                                                                       float* __restrict__ partials) {
    using P = Plan3<A>;
    constexpr int KIN = P::KIN, KP = P::KP, kSet = P::kSet;
    constexpr int kSbo1 = (KP / 4) * 128;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane32 = tid & 31;
    const uint32_t bar_img = smem_u32(smem + P::kBar);
    auto bar_r = [&](int b) { return bar_img + 8 + 8 * b; };        // recompute into buffer b complete (and every MMA its issuer issued before)
    auto bar_c = [&](int b) { return bar_img + 24 + 8 * b; };       // the consumers are done with buffer b
    auto bar_full = [&](int b) { return bar_img + 40 + 8 * b; };    // tile operands of shared-memory buffer b written
    auto bar_empty = [&](int b) { return bar_img + 56 + 8 * b; };   // every MMA reading shared-memory buffer b complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_r(b), 1);
            mbar_init(bar_c(b), kConsumers / 32 / kGroups);
            mbar_init(bar_full(b), kProducers / 32);
            mbar_init(bar_empty(b), 2);                              // one commit per issuer
        }
        mbar_fence_init();
        tma_bulk_load(smem, image, P::kImageLoad, bar_img);
    }
    // operand rows that are never written stay zero (both tile buffers)
    for (int i = tid; i < 2 * P::kTileBytes / 4; i += kThreads3) reinterpret_cast<uint32_t*>(smem + P::kTile)[i] = 0u;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
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
    float* dst = partials + (int64_t)cta * P::kParams;

    if (tid >= kConsumers + kProducers) {
        // ------------------------------------------------------------ issuers: warp b issues the stages with s & 1 == b
        // stage s: tile k = s >> 3, hidden half = s & 1 (== buffer == issuer), trunk = (s >> 1) & 1, row half = (s >> 2) & 1
        const int b = warp - (kConsumers + kProducers) / 32;
        tcp::mbar_wait_c(bar_img, 0);
        int64_t seen_full = -1;
        auto need_tile = [&](int64_t k) {       // (whole warp) the producers have written tile k's operands
            if (k != seen_full) {
                tcp::mbar_wait_c(bar_full((int)(k & 1)), (uint32_t)(k >> 1) & 1u);
                seen_full = k;
            }
        };
        auto recompute = [&](int64_t s) {      // H^T of stage s into tensor-memory buffer b (elected lane)
            const int64_t k = s >> 3;
            const int trunk = (int)(s >> 1) & 1, rh = (int)(s >> 2) & 1;
            const uint32_t tile = smem_u32(smem + P::kTile + (int)(k & 1) * P::kTileBytes);
            const uint32_t d = tmem_base + (uint32_t)b * 128;
            const uint32_t a_base = smem_u32(smem + P::kSW1) + trunk * P::kTrunkBytes + b * (128 / 8) * kSbo1;
            const uint32_t b_base = tile + P::kX + rh * (64 / 8) * kSbo1;
#pragma unroll
            for (int ks = 0; ks < KP / 8; ++ks)
                mma_ss_n(d, make_desc<KP>(a_base + ks * 256), make_desc<KP>(b_base + ks * 256), idesc_tf32(64), ks > 0);
        };
        if (b < n_stages) {                     // fill the pipeline: stage b
            need_tile(0);
            tc_fence_after();
            if (tcp::elect_one()) {
                recompute(b);
                mma_commit(bar_r(b));
            }
            __syncwarp();
        }
#pragma unroll 1
        for (int64_t s = b; s < n_stages; s += 2) {
            const int trunk = (int)(s >> 1) & 1, rh = (int)(s >> 2) & 1;
            const int64_t k = s >> 3;
            const bool more = s + 2 < n_stages;
            if (more) need_tile((s + 2) >> 3);
            tcp::mbar_wait_c(bar_c(b), (uint32_t)(s >> 1) & 1u);            // relu^T / M^T of stage s are in buffer b
            tc_fence_after();
            // grad(s): D_w2 += relu^T BG^T, D_w1 += M^T BX^T (K = the stage's 64 rows); then recompute(s + 2) into the
            // buffer just read; one commit covers both groups
            const uint32_t tile = smem_u32(smem + P::kTile + (int)(k & 1) * P::kTileBytes);
            const uint64_t bx = desc_lbo_sbo(tile + (trunk == 0 ? P::kBXV : P::kBXP) + rh * 16 * P::kLbo, P::kLbo, P::kSboT);
            const uint64_t bg = desc_lbo_sbo(tile + P::kBG + rh * 16 * P::kLbo, P::kLbo, P::kSboT);
            const int nx = P::nx(trunk);
            const uint32_t acc = tmem_base + P::acc(trunk);
            const uint32_t idx = trunk == 0 ? idesc_tf32(P::kNXV) : idesc_tf32(P::kNXP);
            const uint32_t buf = tmem_base + (uint32_t)b * 128;
            if (tcp::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    tcp::mma_ts(acc + 2 * nx + b * P::kNG, buf + ks * 8, bg + (uint64_t)((ks * 2 * P::kLbo) >> 4),
                                idesc_tf32(P::kNG), true);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    tcp::mma_ts(acc + b * nx, buf + 64 + ks * 8, bx + (uint64_t)((ks * 2 * P::kLbo) >> 4), idx, true);
                if ((s & 7) >= 6) mma_commit(bar_empty((int)(k & 1)));   // this issuer's last reads of the tile's shared-memory operands
                if (more) recompute(s + 2);
                mma_commit(bar_r(b));
            }
            __syncwarp();
        }
    } else if (tid >= kConsumers) {
        // ------------------------------------------------------------ producers: one thread per tile row
        const int n = tid - kConsumers, pw = n >> 5;
        auto off_t = [](int c, int nn) { return (c >> 3) * P::kSboT + (nn >> 2) * P::kLbo + (c & 7) * 16 + (nn & 3) * 4; };
        float gsum[1 + A];
#pragma unroll
        for (int a = 0; a <= A; ++a) gsum[a] = 0.f;
        for (int64_t k = 0; k < my_tiles; ++k) {
            const int64_t u = my_first + k * my_stride;
            int64_t row = u * kTileM + n;
            bool active = row < N;
            if (split) {
                const int64_t tt = 2 * (u / tiles_per_t) + player, j = (u % tiles_per_t) * kTileM + n;
                row = tt * B_split + j;
                active = j < B_split;
            }
            float x[KIN], g[1 + A];
            load_row<KIN>(obs, active ? row : 0, active, x);
            g[0] = active ? __ldg(d_v + row) : 0.f;
#pragma unroll
            for (int a = 0; a < A; ++a) g[1 + a] = active ? __ldg(d_logit + row * A + a) : 0.f;
            const int tb = (int)(k & 1);
            if (k >= 2) tcp::mbar_wait_c(bar_empty(tb), (uint32_t)((k >> 1) - 1) & 1u);   // the MMAs of tile k - 2 are done with it
            uint8_t* tile = smem + P::kTile + tb * P::kTileBytes;
            store_operand_row<KIN, KP, P::kBiasInK>(tile + P::kX, n, x);
            // the masks' B operands: the row's gradient signal times its observation, and the signal itself (-> db1)
#pragma unroll
            for (int kk = 0; kk < KIN; ++kk)
                *reinterpret_cast<float*>(tile + P::kBXV + off_t(kk, n)) = to_tf32_fast(g[0] * x[kk]);
            *reinterpret_cast<float*>(tile + P::kBXV + off_t(KIN, n)) = to_tf32_fast(g[0]);
#pragma unroll
            for (int a = 0; a < A; ++a) {
#pragma unroll
                for (int kk = 0; kk < KIN; ++kk)
                    *reinterpret_cast<float*>(tile + P::kBXP + off_t(a * kSet + kk, n)) = to_tf32_fast(g[1 + a] * x[kk]);
                *reinterpret_cast<float*>(tile + P::kBXP + off_t(a * kSet + KIN, n)) = to_tf32_fast(g[1 + a]);
            }
#pragma unroll
            for (int a = 0; a <= A; ++a) {
                *reinterpret_cast<float*>(tile + P::kBG + off_t(a, n)) = to_tf32_fast(g[a]);
                gsum[a] += g[a];
            }
            fence_async_smem();
            __syncwarp();
            if (lane32 == 0) tcp::mbar_arrive(bar_full(tb));
        }
        // output-bias gradients: sums of g over the CTA's rows
        float* s_red = reinterpret_cast<float*>(smem + P::kRed);
#pragma unroll
        for (int a = 0; a <= A; ++a) {
            const float v = warp_sum(gsum[a]);
            if (lane32 == 0) s_red[pw * 8 + a] = v;
        }
    } else {
        // ------------------------------------------------------------ consumers: thread = hidden unit x 16 * kGroups rows of a stage
        // Two GROUPS of eight warps, one per stage buffer (group b takes the stages with s & 1 == b): a stage's
        // round trip - barrier wake-up, tcgen05.ld, arithmetic, tcgen05.st, arrival - is mostly latency, and with all
        // sixteen warps on the same stage nothing hid it (the consumers, not the tensor core, bounded the kernel).
        const int quad = warp & 3, cpart = warp >> 2;             // TMEM lane quadrant (hidden unit); (trunk, half) of the final read-out
        const int group = kGroups == 2 ? warp >> 3 : 0;
        const int cw = kGroups == 2 ? cpart & 1 : cpart;           // which kColsW columns (rows of the tile) of a 64-column stage
        constexpr int kColsW = 16 * kGroups;
        const int j_local = quad * 32 + lane32;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        tcp::mbar_wait_c(bar_img, 0);
        const float* b1 = reinterpret_cast<const float*>(smem + P::kSB1);
        float bias_j[4];                                           // [trunk][half] (zero where the bias rides in K)
#pragma unroll
        for (int c = 0; c < 4; ++c) bias_j[c] = P::kBiasInK ? 0.f : b1[(c >> 1) * kHidden + (c & 1) * 128 + j_local];
#pragma unroll 1
        for (int64_t s = group; s < n_stages; s += kGroups) {
            const int b = (int)(s & 1), trunk = (int)(s >> 1) & 1;
            tcp::mbar_wait_c(bar_r(b), (uint32_t)(s >> 1) & 1u);
            tc_fence_after();
            // ---- this thread's rows of its hidden unit: relu^T in place of H^T, the 0/1 mask next to it
            uint32_t hr[kColsW];
            const uint32_t th = tmem_lane + b * 128 + cw * kColsW;
#pragma unroll
            for (int q = 0; q < kGroups; ++q) tmem_ld16(th + q * 16, hr + q * 16);
            tmem_ld_wait();
            if (!P::kBiasInK) {
                const float bias = bias_j[trunk * 2 + b];
#pragma unroll
                for (int i = 0; i < kColsW; ++i) hr[i] = __float_as_uint(__uint_as_float(hr[i]) + bias);
            }
#pragma unroll
            for (int q = 0; q < kGroups; ++q) {
                uint32_t mk[16];
                uint32_t* h16 = hr + q * 16;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float h = __uint_as_float(h16[i]);
                    mk[i] = __float_as_uint(__saturatef(h * 1.7014118346046923e38f));   // 2^127: 1 for every normal h > 0, else 0
                    h16[i] = __float_as_uint(fmaxf(h, 0.f));
                }
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(th + q * 16),
                    "r"(h16[0]), "r"(h16[1]), "r"(h16[2]), "r"(h16[3]), "r"(h16[4]), "r"(h16[5]), "r"(h16[6]), "r"(h16[7]), "r"(h16[8]),
                    "r"(h16[9]), "r"(h16[10]), "r"(h16[11]), "r"(h16[12]), "r"(h16[13]), "r"(h16[14]), "r"(h16[15])
                    : "memory");
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(th + 64 + q * 16),
                    "r"(mk[0]), "r"(mk[1]), "r"(mk[2]), "r"(mk[3]), "r"(mk[4]), "r"(mk[5]), "r"(mk[6]), "r"(mk[7]), "r"(mk[8]),
                    "r"(mk[9]), "r"(mk[10]), "r"(mk[11]), "r"(mk[12]), "r"(mk[13]), "r"(mk[14]), "r"(mk[15])
                    : "memory");
            }
            tcp::tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane32 == 0) tcp::mbar_arrive(bar_c(b));
        }
        // every gradient MMA complete: the last commit of each issuer (a group has followed its own buffer's barrier
        // only, so it waits there and the groups meet at a named barrier)
        if (kGroups == 2) {
            if (n_stages >= 2) tcp::mbar_wait_c(bar_r(group), (uint32_t)(n_stages >> 1) & 1u);
            tc_fence_before();
            __syncwarp();
            asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
        } else if (n_stages >= 2) {
            tcp::mbar_wait_c(bar_r(0), (uint32_t)(n_stages >> 1) & 1u);
            tcp::mbar_wait_c(bar_r(1), (uint32_t)(n_stages >> 1) & 1u);
        }
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
            float* w1_dst = dst + P::kOffV0w + j * KIN;
#pragma unroll
            for (int kk = 0; kk < KIN; ++kk) w1_dst[kk] = w2 * __uint_as_float(w[kk]);
            dst[P::kOffV0b + j] = w2 * __uint_as_float(w[KIN]);
            uint32_t d2[16];
            tmem_ld16(acc + 2 * P::kNXV + half * P::kNG, d2);
            tmem_ld_wait();
            dst[P::kOffV1w + j] = __uint_as_float(d2[0]);
        } else {
            const uint32_t acc = tmem_lane + P::kAccP;
            uint32_t w[P::kNXP];
#pragma unroll
            for (int q = 0; q < P::kNXP / 16; ++q) tmem_ld16(acc + half * P::kNXP + q * 16, w + q * 16);
            tmem_ld_wait();
            float w2[A];
#pragma unroll
            for (int a = 0; a < A; ++a) w2[a] = __ldg(w2p + a * kHidden + j);
            float* w1_dst = dst + P::kOffP0w + j * KIN;
#pragma unroll
            for (int kk = 0; kk <= KIN; ++kk) {
                float v = w2[0] * __uint_as_float(w[kk]);
#pragma unroll
                for (int a = 1; a < A; ++a) v = fmaf(w2[a], __uint_as_float(w[a * kSet + kk]), v);
                if (kk < KIN) w1_dst[kk] = v;
                else dst[P::kOffP0b + j] = v;
            }
            uint32_t d2[16];
            tmem_ld16(acc + 2 * P::kNXP + half * P::kNG, d2);
            tmem_ld_wait();
#pragma unroll
            for (int a = 0; a < A; ++a) dst[P::kOffP1w + a * kHidden + j] = __uint_as_float(d2[1 + a]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid <= A) {
        const float* s_red = reinterpret_cast<const float*>(smem + P::kRed);
        const float v = (s_red[0 * 8 + tid] + s_red[1 * 8 + tid]) + (s_red[2 * 8 + tid] + s_red[3 * 8 + tid]);
        if (tid == 0) dst[P::kOffV1b] = v;
        else dst[P::kOffP1b + tid - 1] = v;
    }
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

template <int A>
int launch3(const float* obs, int64_t N, int T_split, int64_t B_split, const rnad_mlp_weights& w, const float* d_logit,
            const float* d_v, const uint8_t* image, float* partials, int blocks, cudaStream_t st) {
    using P = Plan3<A>;
    if constexpr (!P::kFits) {
        set_error("learner_backward_tc3: the accumulators of max_actions = %d do not fit tensor memory", A);
        return RNAD_EUNSUPPORTED;
    } else {
        // one CTA per SM (all 512 tensor-memory columns): more than half of the shared memory keeps a second one out
        const size_t smem = P::kBytes > 116 * 1024 ? P::kBytes : 116 * 1024;
        int rc = check_cuda(cudaFuncSetAttribute(learner_bwd_tc3_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(learner_bwd_tc3)");
        if (rc) return rc;
        rc = check_cuda(cudaFuncSetAttribute(learner_bwd_tc3_kernel<A>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             cudaSharedmemCarveoutMaxShared),
                        "cudaFuncSetAttribute(learner_bwd_tc3)");
        if (rc) return rc;
        learner_bwd_tc3_kernel<A><<<blocks, kThreads3, smem, st>>>(obs, N, T_split, B_split, image, w.value_fc1_w,
                                                                  w.policy_fc1_w, d_logit, d_v, partials);
        RNAD_CHECK_LAUNCH("learner_bwd_tc3_kernel");
        return RNAD_OK;
    }
}

}  // namespace

bool learner_backward_tc3_supported(int A) {
    switch (A) {
        case 2: return Plan3<2>::kFits;
        case 3: return Plan3<3>::kFits;
        case 4: return Plan3<4>::kFits;
    }
    return false;
}

int learner_backward_tc3(int A, const float* obs, int64_t N, int T_split, int64_t B_split, const rnad_mlp_weights& w,
                         const float* d_logit, const float* d_v, const uint8_t* image, float* partials, int blocks,
                         cudaStream_t st) {
    switch (A) {
        case 2: return launch3<2>(obs, N, T_split, B_split, w, d_logit, d_v, image, partials, blocks, st);
        case 3: return launch3<3>(obs, N, T_split, B_split, w, d_logit, d_v, image, partials, blocks, st);
        case 4: return launch3<4>(obs, N, T_split, B_split, w, d_logit, d_v, image, partials, blocks, st);
    }
    return RNAD_EINVAL;
}

}  // namespace tc
}  // namespace rnad
