// Learner-side net passes of RNaD.__learn, fused (SURVEY.md section 8 f#3): the C-ABI entry points, the workspace
// layout, and the one kernel still serving max_actions = 4.
//
//   rnad_learner_forward   the four forward_batch calls of rnad.py:373-380 in ONE pass over the trajectory: learner
//                          (logit, pi, log_pi, v), target net (v only), regularisation nets (log_pi only) = five
//                          trunk evaluations per step instead of eight -> learner_fwd_tc2.cu (both layers on
//                          tcgen05, fp16 operands).
//   rnad_learner_backward  parameter gradients of the learner net given d loss/d logit and d loss/d v (from
//   (_split)               rnad_learner_targets).  Dispatch:  split mode, max_actions <= 3 -> learner_bwd_f16.cu
//                          (fp16 operands, mask formulation, four in-place stage regions);  flat mode,
//                          max_actions <= 3 -> learner_bwd_tc3.cu (tf32, mask formulation);  max_actions = 4 ->
//                          learner_bwd_tc2_kernel here: the two trunks recomputed TRANSPOSED on the tensor core,
//                              dW2 = g^T relu(h),  dh = (g W2) * [h > 0],  dW1 = dh^T x
//                          with MMAs whose A operands (relu^T, dh^T) sit in tensor memory and whose accumulators
//                          stay there across all tiles of a CTA.  A second kernel adds the per-CTA partials in a
//                          fixed order (deterministic), per player in split mode.
//   (Earlier builds - a forward with its second layers on the CUDA cores, a backward that reduced on the CUDA cores
//    and one with two CTAs per SM - were removed at the end of round 2; `git log` has them.)
//
// Reference: nn/net.py:64-85 (forward_batch), learn/rnad.py:373-380, 424-425.
// In the reference these are 8 x T small GEMMs + ~40 elementwise launches forward and
// autograd's mirror image backward, with the (T*B x 256) activations written to and
// re-read from HBM about ten times per update.
#include <cstdlib>

#include "learner_bwd.cuh"
#include "tc_common.cuh"
#include "tc_pipe.cuh"

namespace rnad {

// pipelined forward with both layers on the tensor core (learner_fwd_tc2.cu)
bool learner_forward_tc2_supported(int A, int width);
int64_t learner_forward_tc2_image_bytes(int A);
int learner_forward_tc2(const float* obs, int64_t N, int A, const rnad_mlp_weights* net, const rnad_mlp_weights* target,
                        const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_, const rnad_learner_fwd_out* out,
                        void* workspace, cudaStream_t st, int mode, bool others_only);

namespace tc {

// -------------------------------------------------------- backward on the tensor core
//
// dW1 = dh^T x, db1 = dh^T 1 and dW2 = relu(h)^T g are contractions over the ROWS of a tile, while tensor memory
// holds an accumulator with rows as lanes.  So the trunks are recomputed TRANSPOSED: H^T[j][n] = W1[j,:] . X^T[:,n]
// (A operand = the first-layer weights, B operand = the observation tile, both K-major in shared memory) puts hidden
// unit j on lane j and row n on column n.  The thread of lane j turns its columns into relu(h)^T and
// dh^T = ((g W2) * [h > 0])^T in place, and those are A operands *in tensor memory* (K = rows) for
//     D_w1[j][c] += sum_n dh^T[j][n]   BX[c][n]      BX rows: x^T (2A^2 rows), ones     -> dW1, db1
//     D_w2[j][c] += sum_n relu^T[j][n] BG[c][n]      BG rows: d_v, d_logit[0..A)        -> dW2
// whose accumulators stay in tensor memory over all tiles of the CTA.  No activation ever leaves the SM and the
// CUDA cores do ~9 instructions per (row, hidden unit) instead of ~40.  Deterministic (fixed order everywhere).


// Two CTAs share an SM, one per trunk (256 TMEM columns each): while one waits for its MMAs the other runs its
// elementwise stage.  A stage is (128-unit half of the trunk) x (64-row half of the tile).
template <int A>
struct BwdTcPlan : Shape<A> {
    using S = Shape<A>;
    static constexpr int kNX = round_up(S::KIN + 1, 16);               // rows of BX = N of the dW1 / db1 MMA
    static constexpr int kNG = 16;                                     // rows of BG = N of the dW2 MMA
    static constexpr int kLbo = 144;                                   // K-chunk stride of BX / BG: 128 + 16 bytes of padding make the
                                                                       // transposing stores (32 rows n of one operand row c) conflict-free
    static constexpr int kSboT = (kTileM / 4) * kLbo;                  // 8-row groups of BX / BG
    // global image (pack_bwd_tc_image_kernel): both trunks' first layers, their biases, the transposed second layers
    static constexpr int kB = 0;                                       // [2][256 x KP] tf32, K-major
    static constexpr int kB1 = kB + 2 * S::kTrunkBytes;                // [2][256] f32
    static constexpr int kW2T = kB1 + 2 * kHidden * 4;                 // [2][256 x 8] tf32 K-major: row j = (w2v[j],0..) or (0,w2p[0][j],..)
    static constexpr int kImageBytes = kW2T + 2 * kHidden * 8 * 4;
    // shared memory of a CTA: its trunk's slices of the image, then the tile's operands
    static constexpr int kSW1 = 0;
    static constexpr int kSB1 = kSW1 + S::kTrunkBytes;
    static constexpr int kSW2T = kSB1 + kHidden * 4;
    static constexpr int kX = kSW2T + kHidden * 8 * 4;                 // observation tile [128 x KP] tf32, K-major
    static constexpr int kBX = kX + kTileM * S::KP * 4;
    static constexpr int kBG = kBX + (kNX / 8) * kSboT;
    static constexpr int kG = kBG + (kNG / 8) * kSboT;                 // g tile [128 x 8] tf32, K-major: d_v, d_logit[0..A)
    static constexpr int kRed = kG + kTileM * 32;                      // [4 warps][8] partial sums of g (output-bias gradients)
    static constexpr int kBar = kRed + 4 * 32;
    static constexpr int kTmem = kBar + 32;
    static constexpr int kBytes = kTmem + 16;
    // tensor memory (256 columns): [0,64) H^T -> relu^T, [64,128) S^T -> dh^T, then 2 x kNX of D_w1 and 2 x 16 of D_w2
    static constexpr int kColW1 = 128, kColW2 = 128 + 2 * kNX;
    static_assert(kColW2 + 2 * kNG <= 256, "accumulators do not fit the CTA's share of tensor memory");
    static_assert(2 * (kBytes + 1024) <= 228 * 1024, "two CTAs per SM do not fit shared memory");
    static_assert(kX % 16 == 0 && kBX % 16 == 0 && kBG % 16 == 0 && kG % 16 == 0 && kBar % 8 == 0, "alignment");
};

template <int A>
__global__ void pack_bwd_tc_image_kernel(rnad_mlp_weights w, uint8_t* __restrict__ image) {
    using P = BwdTcPlan<A>;
    const int thread = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    pack_trunk_operand<P::KIN, P::KP, P::kBiasInK>(w.value_fc0_w, w.value_fc0_b, image + P::kB, thread, n_threads);
    pack_trunk_operand<P::KIN, P::KP, P::kBiasInK>(w.policy_fc0_w, w.policy_fc0_b, image + P::kB + P::kTrunkBytes,
                                                   thread, n_threads);
    for (int j = thread; j < kHidden; j += n_threads) {
        reinterpret_cast<float*>(image + P::kB1)[j] = w.value_fc0_b[j];
        reinterpret_cast<float*>(image + P::kB1)[kHidden + j] = w.policy_fc0_b[j];
    }
    for (int e = thread; e < 2 * kHidden * 8; e += n_threads) {
        const int row = e / 8, k = e % 8, trunk = row / kHidden, j = row % kHidden;
        float v = 0.f;
        if (trunk == 0 && k == 0) v = w.value_fc1_w[j];
        if (trunk == 1 && k >= 1 && k <= A) v = w.policy_fc1_w[(k - 1) * kHidden + j];
        *reinterpret_cast<float*>(image + P::kW2T + trunk * kHidden * 32 + operand_offset<8>(j, k)) = to_tf32(v);
    }
}

// -------------------------------------------------------- backward, software-pipelined
//
// The same mathematics as learner_bwd_tc_kernel, re-cut so that the tensor core never waits for the CUDA cores and
// vice versa: ONE CTA per SM owns all 512 tensor-memory columns and both trunks, and a stage's H^T / S^T live in one
// of TWO 128-column buffers.  A tile (128 rows) is eight stages s = (trunk, 128-unit hidden half, 64-row half);
//     consumers (16 warps, thread = hidden unit x 16 rows):  wait recompute(s) -> relu^T / dh^T in place in buffer s & 1
//     issuer (one elected lane, after the consumers' barrier): grad(s) = 16 MMAs reading buffer s & 1, then
//                                                              recompute(s + 2) INTO that buffer, one commit
// so while the consumers work on stage s + 1 (the other buffer, whose recompute was issued a stage earlier) the tensor
// core runs grad(s) and recompute(s + 2).  MMAs of one thread execute in issue order, which is what lets
// recompute(s + 2) overwrite the buffer grad(s) reads.  Four producer warps (one thread per tile row) build the next
// tile's operands - the observation tile, x^T | 1, g and g^T - in the other half of a double-buffered shared-memory
// region behind full / empty mbarriers, so the stream of stages never drains at a tile boundary.
// [learner_bwd_tc_kernel: two CTAs per SM, MMA -> elementwise -> MMA serialised per stage, tensor pipe 31 % active,
//  147 us at cfg2; this kernel: 128 us, 35 %.  What bounds a stage (measured, scripts/microbench/tmem_ldst.cu and
//  mma_shapes.cu): ~350 cycles of MMAs + ~320 cycles in which the consumers read and rewrite the stage's 64 KB of
//  tensor memory (tcgen05.ld -> ALU -> tcgen05.st sustains ~200 B / cycle / SM each way, the load side being the slow
//  one) + ~300 issue cycles of relu / mask arithmetic per SM sub-partition, and the tensor core's own tensor-memory
//  traffic (A operands and accumulators) does not overlap with the consumers' - a variant with three buffers and two
//  stages of slack ran no faster (133 us), i.e. the stage time is the SUM of these, not their maximum.]
#ifdef RNAD_TRACE_BWD
// development aid: cycle stamps of CTA 0, stages 16..79: [role][stage - 16][event]; role 0 = consumer warp 0, 1 = consumer
// warp 15, 2 = issuer 0, 3 = issuer 1, 4 = producer warp 0 (per tile)
__device__ long long g_bwd_trace[5][64][8];
#define BTR(role, s, ev) do { if (blockIdx.x == 0 && (s) >= 16 && (s) < 80 && lane32 == 0) g_bwd_trace[role][(s) - 16][ev] = clock64(); } while (0)
#else
#define BTR(role, s, ev) do { } while (0)
#endif
constexpr int kBwd2Consumers = 512, kBwd2Producers = 128, kBwd2Issuers = 64;
constexpr int kBwd2Threads = kBwd2Consumers + kBwd2Producers + kBwd2Issuers;

template <int A>
struct BwdTc2Plan : Shape<A> {
    using S = Shape<A>;
    using PT = BwdTcPlan<A>;
    static constexpr int kNX = PT::kNX, kNG = PT::kNG, kLbo = PT::kLbo, kSboT = PT::kSboT;
    // shared memory: the whole weight image of pack_bwd_tc_image_kernel (both trunks), then two tile buffers
    static constexpr int kSW1 = PT::kB, kSB1 = PT::kB1, kSW2T = PT::kW2T;
    static constexpr int kTile = round_up(PT::kImageBytes, 128);
    static constexpr int kX = 0;                                        // within a tile buffer
    static constexpr int kBX = kX + kTileM * S::KP * 4;
    static constexpr int kBG = kBX + (kNX / 8) * kSboT;
    static constexpr int kG = kBG + (kNG / 8) * kSboT;
    static constexpr int kTileBytes = round_up(kG + kTileM * 32, 128);
    static constexpr int kRed = kTile + 2 * kTileBytes;                 // [4 producer warps][8] sums of g
    static constexpr int kBar = kRed + 4 * 32;                          // image, recompute[2], consumed[2], full[2], empty[2]
    static constexpr int kTmem = kBar + 80;
    static constexpr int kBytes = kTmem + 16;
    // tensor memory: two stage buffers [H^T 64 | S^T 64], then per trunk 2 x kNX of D_w1 and 2 x 16 of D_w2
    static constexpr int kAcc = 256, kAccTrunk = 2 * kNX + 2 * kNG;
    static_assert(kAcc + 2 * kAccTrunk <= 512, "accumulators do not fit tensor memory");
    static_assert(kBytes <= 227 * 1024, "shared memory plan does not fit");
    static_assert(kTile % 16 == 0 && kBX % 16 == 0 && kBG % 16 == 0 && kG % 16 == 0 && kBar % 8 == 0, "alignment");
};

// Stage s of a CTA's stream: tile k = s >> 3; within the tile  half = s & 1 (the 128-unit hidden half - also the
// tensor-memory buffer and the issuer warp of the stage), trunk = (s >> 1) & 1, row half = (s >> 2) & 1.  Consecutive
// stages therefore add into DIFFERENT accumulators and every accumulator is only ever touched by ONE issuing thread,
// whose MMAs execute in issue order: the sums are bit-reproducible although two warps issue.
template <int A>
__global__ void __launch_bounds__(kBwd2Threads, 1) learner_bwd_tc2_kernel(const float* __restrict__ obs, int64_t N,
                                                                           int T_split, int64_t B_split,
                                                                           const uint8_t* __restrict__ image,
                                                                           const float* __restrict__ d_logit,
                                                                           const float* __restrict__ d_v,
                                                                           float* __restrict__ partials) {
    using P = BwdTc2Plan<A>;
    constexpr int KIN = P::KIN, KP = P::KP;
    constexpr int kSbo1 = (KP / 4) * 128;
    static_assert(A <= 4, "g[n] is staged as 8 floats");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane32 = tid & 31;
    const uint32_t bar_img = smem_u32(smem + P::kBar);
    auto bar_r = [&](int b) { return bar_img + 8 + 8 * b; };        // recompute into buffer b complete (and every MMA its issuer issued before)
    auto bar_c = [&](int b) { return bar_img + 24 + 8 * b; };       // the consumers are done with buffer b (relu^T / dh^T in place)
    auto bar_full = [&](int b) { return bar_img + 40 + 8 * b; };    // tile operands of shared-memory buffer b written
    auto bar_empty = [&](int b) { return bar_img + 56 + 8 * b; };   // every MMA reading shared-memory buffer b complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P::kTmem);

    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        mbar_init(bar_img, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_r(b), 1);
            mbar_init(bar_c(b), kBwd2Consumers / 32);
            mbar_init(bar_full(b), kBwd2Producers / 32);
            mbar_init(bar_empty(b), 2);                              // one commit per issuer
        }
        mbar_fence_init();
        tma_bulk_load(smem, image, BwdTcPlan<A>::kImageBytes, bar_img);
    }
    // operand rows that are never written stay zero (both tile buffers)
    for (int i = tid; i < 2 * P::kTileBytes / 4; i += kBwd2Threads) reinterpret_cast<uint32_t*>(smem + P::kTile)[i] = 0u;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid < kBwd2Consumers) {
        // clear the gradient accumulators: columns [256, 256 + 2 * kAccTrunk), 16 at a time, split over the four column parts
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int c = P::kAcc + (warp >> 2) * 16; c < P::kAcc + 2 * P::kAccTrunk; c += 64)
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_base + c),
                "r"(0u)
                : "memory");
        tcp::tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // which tiles this CTA walks: as in learner_bwd_tc_kernel, with the CTA in the role of the CTA pair
    const int cta = blockIdx.x, n_ctas = gridDim.x;
    const bool split = T_split > 0;
    const int player = split ? (cta & 1) : 0;
    const int64_t tiles_per_t = split ? (B_split + kTileM - 1) / kTileM : 0;
    const int64_t my_first = split ? (cta >> 1) : cta, my_stride = split ? (n_ctas >> 1) : n_ctas;
    const int64_t num_tiles = split ? (int64_t)((T_split - player + 1) / 2) * tiles_per_t : (N + kTileM - 1) / kTileM;
    const int64_t my_tiles = my_first < num_tiles ? (num_tiles - 1 - my_first) / my_stride + 1 : 0;
    const int64_t n_stages = my_tiles * 8;
    float* dst = partials + (int64_t)cta * P::kParams;

    if (tid >= kBwd2Consumers + kBwd2Producers) {
        // ------------------------------------------------------------ issuers: warp b issues the stages with s & 1 == b
        const int b = warp - (kBwd2Consumers + kBwd2Producers) / 32;
        tcp::mbar_wait_c(bar_img, 0);
        int64_t seen_full = -1;
        auto need_tile = [&](int64_t k) {       // (whole warp) the producers have written tile k's operands
            if (k != seen_full) {
                tcp::mbar_wait_c(bar_full((int)(k & 1)), (uint32_t)(k >> 1) & 1u);
                seen_full = k;
            }
        };
        auto recompute = [&](int64_t s) {      // H^T | S^T of stage s into tensor-memory buffer b (elected lane)
            const int64_t k = s >> 3;
            const int trunk = (int)(s >> 1) & 1, rh = (int)(s >> 2) & 1;
            const uint32_t tile = smem_u32(smem + P::kTile + (int)(k & 1) * P::kTileBytes);
            const uint32_t d = tmem_base + (uint32_t)b * 128;
            const uint32_t a_base = smem_u32(smem + P::kSW1) + trunk * P::kTrunkBytes + b * (128 / 8) * kSbo1;
            const uint32_t b_base = tile + P::kX + rh * (64 / 8) * kSbo1;
#pragma unroll
            for (int ks = 0; ks < KP / 8; ++ks)
                mma_ss_n(d, make_desc<KP>(a_base + ks * 256), make_desc<KP>(b_base + ks * 256), idesc_tf32(64), ks > 0);
            mma_ss_n(d + 64, make_desc<8>(smem_u32(smem + P::kSW2T) + trunk * kHidden * 32 + b * 128 * 32),
                     make_desc<8>(tile + P::kG + rh * 64 * 32), idesc_tf32(64), false);
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
            BTR(2 + b, s, 0);
            if (more) need_tile((s + 2) >> 3);
            BTR(2 + b, s, 1);
            tcp::mbar_wait_c(bar_c(b), (uint32_t)(s >> 1) & 1u);            // relu^T / dh^T of stage s are in buffer b
            tc_fence_after();
            BTR(2 + b, s, 2);
            // grad(s): D_w2 += relu^T BG^T, D_w1 += dh^T BX^T (K = the stage's 64 rows); then recompute(s + 2) into the buffer
            // just read; one commit covers both groups
            const uint32_t tile = smem_u32(smem + P::kTile + (int)(k & 1) * P::kTileBytes);
            const uint64_t bx = desc_lbo_sbo(tile + P::kBX + rh * 16 * P::kLbo, P::kLbo, P::kSboT);
            const uint64_t bg = desc_lbo_sbo(tile + P::kBG + rh * 16 * P::kLbo, P::kLbo, P::kSboT);
            const uint32_t acc = tmem_base + P::kAcc + trunk * P::kAccTrunk;
            const uint32_t buf = tmem_base + (uint32_t)b * 128;
            if (tcp::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    tcp::mma_ts(acc + 2 * P::kNX + b * P::kNG, buf + ks * 8, bg + (uint64_t)((ks * 2 * P::kLbo) >> 4),
                                idesc_tf32(P::kNG), true);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    tcp::mma_ts(acc + b * P::kNX, buf + 64 + ks * 8, bx + (uint64_t)((ks * 2 * P::kLbo) >> 4),
                                idesc_tf32(P::kNX), true);
                if ((s & 7) >= 6) mma_commit(bar_empty((int)(k & 1)));   // this issuer's last reads of the tile's shared-memory operands
                if (more) recompute(s + 2);
                mma_commit(bar_r(b));
            }
            __syncwarp();
            BTR(2 + b, s, 3);
        }
    } else if (tid >= kBwd2Consumers) {
        // ------------------------------------------------------------ producers: one thread per tile row
        const int n = tid - kBwd2Consumers, pw = n >> 5;
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
            if (pw == 0) BTR(4, 8 * k, 0);
            if (k >= 2) tcp::mbar_wait_c(bar_empty(tb), (uint32_t)((k >> 1) - 1) & 1u);   // the MMAs of tile k - 2 are done with it
            if (pw == 0) BTR(4, 8 * k, 1);
            uint8_t* tile = smem + P::kTile + tb * P::kTileBytes;
            store_operand_row<KIN, KP, P::kBiasInK>(tile + P::kX, n, x);
#pragma unroll
            for (int kk = 0; kk < KIN; ++kk) *reinterpret_cast<float*>(tile + P::kBX + off_t(kk, n)) = to_tf32_fast(x[kk]);
            *reinterpret_cast<float*>(tile + P::kBX + off_t(KIN, n)) = 1.f;
            float g8[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) g8[a] = a <= A ? to_tf32_fast(g[a < 1 + A ? a : 0]) : 0.f;
#pragma unroll
            for (int a = 0; a <= A; ++a) {
                *reinterpret_cast<float*>(tile + P::kBG + off_t(a, n)) = g8[a];
                gsum[a] += g[a];
            }
            *reinterpret_cast<float4*>(tile + P::kG + operand_offset<8>(n, 0)) = make_float4(g8[0], g8[1], g8[2], g8[3]);
            *reinterpret_cast<float4*>(tile + P::kG + operand_offset<8>(n, 4)) = make_float4(g8[4], g8[5], g8[6], g8[7]);
            fence_async_smem();
            __syncwarp();
            if (lane32 == 0) tcp::mbar_arrive(bar_full(tb));
            if (pw == 0) BTR(4, 8 * k, 2);
        }
        // output-bias gradients: sums of g over the CTA's rows
        float* s_red = reinterpret_cast<float*>(smem + P::kRed);
#pragma unroll
        for (int a = 0; a <= A; ++a) {
            const float v = warp_sum(gsum[a]);
            if (lane32 == 0) s_red[pw * 8 + a] = v;
        }
    } else {
        // ------------------------------------------------------------ consumers: thread = hidden unit x 16 rows of a stage
        const int quad = warp & 3, cpart = warp >> 2;             // TMEM lane quadrant (hidden unit), 16-column (row) part of a stage
        const int j_local = quad * 32 + lane32;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        tcp::mbar_wait_c(bar_img, 0);
        const float* b1 = reinterpret_cast<const float*>(smem + P::kSB1);
        float bias_j[4];                                           // [trunk][half]
#pragma unroll
        for (int c = 0; c < 4; ++c) bias_j[c] = P::kBiasInK ? 0.f : b1[(c >> 1) * kHidden + (c & 1) * 128 + j_local];
#pragma unroll 1
        for (int64_t s = 0; s < n_stages; ++s) {
            const int b = (int)(s & 1), trunk = (int)(s >> 1) & 1;
            const float bias = bias_j[trunk * 2 + b];
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 0);
            tcp::mbar_wait_c(bar_r(b), (uint32_t)(s >> 1) & 1u);
            tc_fence_after();
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 1);
            // ---- this thread's 16 rows of its hidden unit: relu^T over H^T, dh^T = S^T where h > 0, both in place
            uint32_t hr[16], dh[16];
            const uint32_t th = tmem_lane + b * 128 + cpart * 16;
            tmem_ld16(th, hr);
            tmem_ld16(th + 64, dh);
            tmem_ld_wait();
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 2);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float h = __uint_as_float(hr[i]) + bias;
                const bool on = h > 0.f;
                hr[i] = on ? __float_as_uint(h) : 0u;
                dh[i] = on ? dh[i] : 0u;
            }
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(th),
                "r"(hr[0]), "r"(hr[1]), "r"(hr[2]), "r"(hr[3]), "r"(hr[4]), "r"(hr[5]), "r"(hr[6]), "r"(hr[7]), "r"(hr[8]),
                "r"(hr[9]), "r"(hr[10]), "r"(hr[11]), "r"(hr[12]), "r"(hr[13]), "r"(hr[14]), "r"(hr[15])
                : "memory");
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(th + 64),
                "r"(dh[0]), "r"(dh[1]), "r"(dh[2]), "r"(dh[3]), "r"(dh[4]), "r"(dh[5]), "r"(dh[6]), "r"(dh[7]), "r"(dh[8]),
                "r"(dh[9]), "r"(dh[10]), "r"(dh[11]), "r"(dh[12]), "r"(dh[13]), "r"(dh[14]), "r"(dh[15])
                : "memory");
            tcp::tmem_st_wait();
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 3);
            tc_fence_before();
            __syncwarp();
            if (lane32 == 0) tcp::mbar_arrive(bar_c(b));
            if (warp == 0 || warp == 15) BTR(warp == 0 ? 0 : 1, s, 4);
        }
        // every gradient MMA complete: the last commit of each issuer
        if (n_stages >= 2) {
            tcp::mbar_wait_c(bar_r(0), (uint32_t)(n_stages >> 1) & 1u);
            tcp::mbar_wait_c(bar_r(1), (uint32_t)(n_stages >> 1) & 1u);
        }
        tc_fence_after();

        // ---- this CTA's partial gradient, flat in state_dict order: column part c reads (trunk, half) = (c >> 1, c & 1)
        {
            const int trunk = cpart >> 1, half = cpart & 1;
            const int j = half * 128 + j_local;
            const uint32_t acc = tmem_lane + P::kAcc + trunk * P::kAccTrunk;
            uint32_t w[P::kNX];
#pragma unroll
            for (int q = 0; q < P::kNX / 16; ++q) tmem_ld16(acc + half * P::kNX + q * 16, w + q * 16);
            tmem_ld_wait();
            float* w1_dst = dst + (trunk == 0 ? P::kOffV0w : P::kOffP0w) + j * KIN;
#pragma unroll
            for (int kk = 0; kk < KIN; ++kk) w1_dst[kk] = __uint_as_float(w[kk]);
            dst[(trunk == 0 ? P::kOffV0b : P::kOffP0b) + j] = __uint_as_float(w[KIN]);
            uint32_t w2[16];
            tmem_ld16(acc + 2 * P::kNX + half * P::kNG, w2);
            tmem_ld_wait();
            if (trunk == 0) {
                dst[P::kOffV1w + j] = __uint_as_float(w2[0]);
            } else {
#pragma unroll
                for (int a = 0; a < A; ++a) dst[P::kOffP1w + a * kHidden + j] = __uint_as_float(w2[1 + a]);
            }
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

#ifdef RNAD_TRACE_BWD
}  // namespace tc
}  // namespace rnad
extern "C" __attribute__((visibility("default"))) int rnad_debug_bwd_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, rnad::tc::g_bwd_trace, sizeof(rnad::tc::g_bwd_trace));
}
namespace rnad {
namespace tc {
#endif

// Sum of the per-CTA partial gradients in a FIXED order (deterministic): eight lanes per parameter take the partials
// p = lane, lane + 8, ... (eight loads in flight per thread as well) and meet in a shuffle tree.
// Split mode (gridDim.y == 2): blockIdx.y = player; player p's partials are the rows p, p + 2, ... (pairs with index
// parity p) and its sum goes to flat_grad + p * n_params.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_parts, int n_params,
                                       float* __restrict__ flat_grad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 3, sub = t & 7;
    const int step = gridDim.y, first = blockIdx.y;      // 1, 0 in flat mode
    float acc = 0.f;
    if (i < n_params) {
        float a0 = 0.f, a1 = 0.f;
        int p = first + sub * step;
        for (; p + 8 * step < n_parts; p += 16 * step) {
            a0 += partials[(int64_t)p * n_params + i];
            a1 += partials[(int64_t)(p + 8 * step) * n_params + i];
        }
        if (p < n_parts) a0 += partials[(int64_t)p * n_params + i];
        acc = a0 + a1;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (i < n_params && sub == 0) flat_grad[(int64_t)blockIdx.y * n_params + i] = acc;
}


// the forward kernel's weight image (learner_fwd_tc2.cu) sits at the start of the workspace
template <int A>
int64_t fwd_image_reserve() {
    return round_up((int)learner_forward_tc2_image_bytes(A), 256);
}

// then the backward's images: the tf32 one (pack_bwd_tc_image_kernel) and, where that engine exists, the fp16 one
template <int A>
int64_t bwd_image_reserve() {
    return round_up(BwdTcPlan<A>::kImageBytes, 256) + round_up((int)learner_backward_f16_image_bytes(A), 256);
}

template <int A>
int64_t workspace_bytes() {
    return fwd_image_reserve<A>() + bwd_image_reserve<A>() + (int64_t)kMaxBwdCtas * Shape<A>::kParams * 4;
}

template <int A, typename Kernel>
int prepare(Kernel kernel, size_t smem, const char* what) {
    int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), what);
    if (rc) return rc;
    return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared), what);
}

// T_split > 0: the rows are a (T_split, B_split) trajectory and flat_grad receives TWO unnormalised gradients,
// player 0's (rows of even t) then player 1's.  mode (as in learner_fwd_tc2.cu): 0 pack + run, 1 prepacked, 2 pack only.
template <int A>
int launch_backward(const float* obs, int64_t N, int T_split, int64_t B_split, const rnad_mlp_weights& w,
                    const float* d_logit, const float* d_v, float* flat_grad, uint8_t* workspace, cudaStream_t st,
                    int mode = 0) {
    using P = Shape<A>;
    using PT = BwdTcPlan<A>;
    uint8_t* image = workspace + fwd_image_reserve<A>();
    uint8_t* image_f16 = image + round_up(PT::kImageBytes, 256);
    float* partials = reinterpret_cast<float*>(image + bwd_image_reserve<A>());
    int64_t blocks = (N + kTileM - 1) / kTileM;
    const int cap = sm_count() < kMaxBwdCtas ? sm_count() : kMaxBwdCtas;
    if (T_split > 0) {
        // an even number of CTAs, half of them per player; at least one each (a player without rows writes zeros)
        const int64_t per_player = (int64_t)((T_split + 1) / 2) * ((B_split + kTileM - 1) / kTileM);
        blocks = 2 * (per_player < cap / 2 ? per_player : cap / 2);
    } else if (blocks > cap) {
        blocks = cap;
    }
    const dim3 reduce_grid((P::kParams * 8 + 255) / 256, T_split > 0 ? 2 : 1);
    // split mode (one UNNORMALISED gradient per player: the learner step) runs on the fp16-operand engine
    // (learner_bwd_f16.cu) where it exists; RNAD_LEARNER_BWD_TF32 keeps the tf32 kernels there too, for A/B runs
    static const bool v2 = getenv("RNAD_LEARNER_BWD_V2") != nullptr;         // the S^T formulation where the mask one fits, for A/B runs
    static const bool tf32_only = getenv("RNAD_LEARNER_BWD_TF32") != nullptr || v2;
    const bool f16_engine = !tf32_only && learner_backward_f16_supported(A);
    if (f16_engine && (T_split > 0 || mode == 2)) {
        int rc = learner_backward_f16(A, obs, N, T_split, B_split, w, d_logit, d_v, image_f16, partials, (int)blocks, st, mode);
        if (rc) return rc;
        if (mode != 2) {
            reduce_partials_kernel<<<reduce_grid, 256, 0, st>>>(partials, (int)blocks, P::kParams, flat_grad);
            RNAD_CHECK_LAUNCH("reduce_partials_kernel");
        }
        return RNAD_OK;      // (pack only: the prepacked entry point is the split one, which this engine serves)
    }
    if (mode != 1) {
        pack_bwd_tc_image_kernel<A><<<32, 256, 0, st>>>(w, image);
        RNAD_CHECK_LAUNCH("pack_bwd_tc_image_kernel");
        if (mode == 2) return RNAD_OK;
    }
    static_assert(PT::kB == 0 && PT::kB1 == 2 * PT::kTrunkBytes, "learner_bwd_tc3.cu reads the head of this image");
    if (!v2 && learner_backward_tc3_supported(A)) {
        int rc = learner_backward_tc3(A, obs, N, T_split, B_split, w, d_logit, d_v, image, partials, (int)blocks, st);
        if (rc) return rc;
    } else {
        using P2 = BwdTc2Plan<A>;
        // one CTA per SM (all 512 tensor-memory columns): more than half of the shared memory keeps a second one out
        const size_t smem2 = P2::kBytes > 116 * 1024 ? P2::kBytes : 116 * 1024;
        int rc = prepare<A>(learner_bwd_tc2_kernel<A>, smem2, "cudaFuncSetAttribute(learner_bwd_tc2)");
        if (rc) return rc;
        learner_bwd_tc2_kernel<A><<<(int)blocks, kBwd2Threads, smem2, st>>>(obs, N, T_split, B_split, image, d_logit, d_v,
                                                                           partials);
        RNAD_CHECK_LAUNCH("learner_bwd_tc2_kernel");
    }
    reduce_partials_kernel<<<reduce_grid, 256, 0, st>>>(partials, (int)blocks, P::kParams, flat_grad);
    RNAD_CHECK_LAUNCH("reduce_partials_kernel");
    return RNAD_OK;
}

}  // namespace tc
}  // namespace rnad

using namespace rnad;

static bool weights_ok(const rnad_mlp_weights* w) {
    return w && w->value_fc0_w && w->value_fc0_b && w->value_fc1_w && w->value_fc1_b && w->policy_fc0_w &&
           w->policy_fc0_b && w->policy_fc1_w && w->policy_fc1_b;
}

extern "C" {

int rnad_learner_mlp_supported(int A, int width) { return (width == tc::kHidden && A >= 2 && A <= 4) ? 1 : 0; }

int64_t rnad_learner_mlp_workspace_bytes(int A, int width) {
    if (!rnad_learner_mlp_supported(A, width)) return 0;
    switch (A) {
        case 2: return tc::workspace_bytes<2>();
        case 3: return tc::workspace_bytes<3>();
        case 4: return tc::workspace_bytes<4>();
    }
    return 0;
}

int rnad_learner_param_count(int A, int width) {
    if (!rnad_learner_mlp_supported(A, width)) return 0;
    return 2 * width * (2 * A * A) + 2 * width + width + 1 + A * width + A;
}

static int learner_forward_impl(const float* observations, int64_t N, int A, const rnad_mlp_weights* net,
                                const rnad_mlp_weights* target, const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_,
                                const rnad_learner_fwd_out* out, void* workspace, void* stream, int prepacked,
                                int others_only = 0) {
    RNAD_REQUIRE(observations && out && workspace, "rnad_learner_forward: null pointer");
    RNAD_REQUIRE(weights_ok(net) && weights_ok(target) && weights_ok(reg) && weights_ok(reg_),
                 "rnad_learner_forward: null weight pointer");
    RNAD_REQUIRE(out->v_target && out->log_pi_reg && out->log_pi_reg_ &&
                     (others_only || (out->logit && out->pi && out->log_pi && out->v)),
                 "rnad_learner_forward: null output pointer");
    RNAD_REQUIRE(N >= 0, "rnad_learner_forward: negative row count");
    if (!rnad_learner_mlp_supported(A, net->width) || target->width != net->width || reg->width != net->width ||
        reg_->width != net->width) {
        set_error("rnad_learner_forward: needs four nets of width 256 and 2 <= max_actions <= 4");
        return RNAD_EUNSUPPORTED;
    }
    RNAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rnad_learner_forward: workspace must be 256-byte aligned");
    if (N == 0) return RNAD_OK;
    if (!learner_forward_tc2_supported(A, net->width)) {
        set_error("rnad_learner_forward: no forward engine for max_actions %d, width %d", A, net->width);
        return RNAD_EUNSUPPORTED;
    }
    return learner_forward_tc2(observations, N, A, net, target, reg, reg_, out, workspace, (cudaStream_t)stream,
                               prepacked ? 1 : 0, others_only != 0);
}

int rnad_learner_forward(const float* observations, int64_t N, int A, const rnad_mlp_weights* net,
                         const rnad_mlp_weights* target, const rnad_mlp_weights* reg, const rnad_mlp_weights* reg_,
                         const rnad_learner_fwd_out* out, void* workspace, void* stream) {
    return learner_forward_impl(observations, N, A, net, target, reg, reg_, out, workspace, stream, 0);
}

int rnad_learner_forward_prepacked(const float* observations, int64_t N, int A, const rnad_mlp_weights* net,
                                   const rnad_mlp_weights* target, const rnad_mlp_weights* reg,
                                   const rnad_mlp_weights* reg_, const rnad_learner_fwd_out* out, int others_only,
                                   void* workspace, void* stream) {
    return learner_forward_impl(observations, N, A, net, target, reg, reg_, out, workspace, stream, 1, others_only);
}

int rnad_learner_pack(int A, const rnad_mlp_weights* net, const rnad_mlp_weights* target, const rnad_mlp_weights* reg,
                      const rnad_mlp_weights* reg_, int others_only, void* workspace, void* stream) {
    RNAD_REQUIRE(workspace, "rnad_learner_pack: null pointer");
    RNAD_REQUIRE(weights_ok(net) && weights_ok(target) && weights_ok(reg) && weights_ok(reg_), "rnad_learner_pack: null weight pointer");
    if (!rnad_learner_mlp_supported(A, net->width) || target->width != net->width || reg->width != net->width ||
        reg_->width != net->width) {
        set_error("rnad_learner_pack: needs four nets of width 256 and 2 <= max_actions <= 4");
        return RNAD_EUNSUPPORTED;
    }
    RNAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rnad_learner_pack: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = learner_forward_tc2(nullptr, 0, A, net, target, reg, reg_, nullptr, workspace, st, 2, others_only != 0);
    if (rc) return rc;
    switch (A) {
        case 2: return tc::launch_backward<2>(nullptr, 0, 0, 0, *net, nullptr, nullptr, nullptr, (uint8_t*)workspace, st, 2);
        case 3: return tc::launch_backward<3>(nullptr, 0, 0, 0, *net, nullptr, nullptr, nullptr, (uint8_t*)workspace, st, 2);
        case 4: return tc::launch_backward<4>(nullptr, 0, 0, 0, *net, nullptr, nullptr, nullptr, (uint8_t*)workspace, st, 2);
    }
    return RNAD_EINVAL;
}

int rnad_learner_backward(const float* observations, int64_t N, int A, const rnad_mlp_weights* net,
                          const float* d_logit, const float* d_v, float* flat_grad, void* workspace, void* stream) {
    RNAD_REQUIRE(observations && d_logit && d_v && flat_grad && workspace, "rnad_learner_backward: null pointer");
    RNAD_REQUIRE(weights_ok(net), "rnad_learner_backward: null weight pointer");
    RNAD_REQUIRE(N >= 1, "rnad_learner_backward: empty batch");
    if (!rnad_learner_mlp_supported(A, net->width)) {
        set_error("rnad_learner_backward: needs width 256 and 2 <= max_actions <= 4");
        return RNAD_EUNSUPPORTED;
    }
    RNAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rnad_learner_backward: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    switch (A) {
        case 2: return tc::launch_backward<2>(observations, N, 0, 0, *net, d_logit, d_v, flat_grad, (uint8_t*)workspace, st);
        case 3: return tc::launch_backward<3>(observations, N, 0, 0, *net, d_logit, d_v, flat_grad, (uint8_t*)workspace, st);
        case 4: return tc::launch_backward<4>(observations, N, 0, 0, *net, d_logit, d_v, flat_grad, (uint8_t*)workspace, st);
    }
    return RNAD_EINVAL;
}

static int learner_backward_split_impl(const float* observations, int T, int64_t B, int A, const rnad_mlp_weights* net,
                                       const float* d_logit, const float* d_v, float* player_grads, void* workspace,
                                       void* stream, int prepacked) {
    RNAD_REQUIRE(observations && d_logit && d_v && player_grads && workspace, "rnad_learner_backward_split: null pointer");
    RNAD_REQUIRE(weights_ok(net), "rnad_learner_backward_split: null weight pointer");
    RNAD_REQUIRE(T >= 1 && B >= 1, "rnad_learner_backward_split: empty trajectory");
    if (!rnad_learner_mlp_supported(A, net->width)) {
        set_error("rnad_learner_backward_split: needs width 256 and 2 <= max_actions <= 4");
        return RNAD_EUNSUPPORTED;
    }
    RNAD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rnad_learner_backward_split: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = (int64_t)T * B;
    switch (A) {
        case 2: return tc::launch_backward<2>(observations, N, T, B, *net, d_logit, d_v, player_grads, (uint8_t*)workspace, st, prepacked);
        case 3: return tc::launch_backward<3>(observations, N, T, B, *net, d_logit, d_v, player_grads, (uint8_t*)workspace, st, prepacked);
        case 4: return tc::launch_backward<4>(observations, N, T, B, *net, d_logit, d_v, player_grads, (uint8_t*)workspace, st, prepacked);
    }
    return RNAD_EINVAL;
}

int rnad_learner_backward_split(const float* observations, int T, int64_t B, int A, const rnad_mlp_weights* net,
                                const float* d_logit, const float* d_v, float* player_grads, void* workspace,
                                void* stream) {
    return learner_backward_split_impl(observations, T, B, A, net, d_logit, d_v, player_grads, workspace, stream, 0);
}

int rnad_learner_backward_split_prepacked(const float* observations, int T, int64_t B, int A, const rnad_mlp_weights* net,
                                          const float* d_logit, const float* d_v, float* player_grads, void* workspace,
                                          void* stream) {
    return learner_backward_split_impl(observations, T, B, A, net, d_logit, d_v, player_grads, workspace, stream, 1);
}

}  // extern "C"
