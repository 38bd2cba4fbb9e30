// The tail of one learner step as ONE kernel, with the data-parallel gradient exchange inside it:
//
//   rnad_learner_tail     [exchange over NVLink peer memory] -> g = G_0 / N_0 + G_1 / N_1 -> clip_grad_norm_ ->
//                         Adam -> target-net average -> both loss values
//   rnad_step_control     the per-step scalars (rollout seed, alpha) into device memory, so that a whole learner step
//                         (rollout -> forward -> targets -> backward -> tail) replays as one CUDA graph
//   rnad_xchg_*           the exchange buffers: cudaMalloc'ed here, shared between the ranks' processes by CUDA IPC
//
// Reference: learn/rnad.py:456 (clip_grad_norm_), :514-515 (optimizer.step / zero_grad), :516-523 (target-net
// average), vtrace.py:370-374, 387-389 (the divisions by the per-player step counts).  The reference has no
// data-parallel code; north_star asks for ONE exchange of the learner gradients per step and nothing else on NVLink.
//
// Exchange.  Both losses are batch sums divided by per-player step counts N_p, and every trajectory row belongs to
// exactly one player, so the backward kernel leaves one UNNORMALISED gradient per player, G_0 and G_1
// (rnad_learner_backward_split).  A rank's message is the row [G_0 | G_1 | N_0, N_1 (each as hi, lo: exact in fp32) |
// the four loss numerators] = 2P + 8 floats.  Every rank PUSHES its row into slot (seq & 1)[rank] of every peer's
// buffer with plain stores through the NVLink mapping, fences (system scope), then writes the flag seq + 1 next to it;
// it then waits until its own flags of this slot show seq + 1 for every peer, and adds the `world` rows in rank order
// - the same order on every rank, so all ranks compute bit-identical parameters and never drift apart.  The division
// by the GLOBAL N_p happens after the sum: the result is the single-process gradient of the concatenated batch, on
// ragged trees too.  Two slots suffice: a rank can be at most one step ahead of a peer (it needs the peer's flag of
// step k to finish step k, and pushes step k + 2 into slot k & 1 only after the peer's flag of step k + 1, which the
// peer writes after it has consumed slot k & 1).  A peer that never arrives trips a 10 s timeout: the error word is
// set, the step completes with what is there, and the host raises.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace rnad {
namespace {

constexpr int kTailThreads = 1024;

__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__host__ __device__ constexpr int64_t xchg_row_floats(int n_params) { return (int64_t)((2 * n_params + 8 + 3) / 4 * 4); }
// [slot][source rank][row] floats, then [slot][source rank] flags
__host__ __device__ constexpr int64_t xchg_flag_offset_floats(int n_params, int world) {
    return 2 * (int64_t)world * xchg_row_floats(n_params);
}

__global__ void step_control_kernel(rnad_step_ctrl* ctrl, uint64_t seed, float alpha) {
    ctrl->seed = seed;
    ctrl->alpha = alpha;
}

// splitmix64 (Steele, Lea, Flood 2014): state += golden gamma, seed = the mixed state's upper 62 bits
__global__ void step_advance_kernel(rnad_step_ctrl* ctrl) {
    uint64_t state = ((uint64_t)ctrl->seed_state[1] << 32 | ctrl->seed_state[0]) + 0x9E3779B97F4A7C15ull;
    ctrl->seed_state[0] = (uint32_t)state;
    ctrl->seed_state[1] = (uint32_t)(state >> 32);
    uint64_t z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    ctrl->seed = (z ^ (z >> 31)) >> 2;
}

// The same, and the step's inputs with it: n floats from `src` - pinned host memory, which unified addressing makes
// device-readable at the same address - to `dst`, one 16-byte load per thread (a single round trip over PCIe instead
// of a copy node of its own in front of a one-thread kernel).
__global__ void step_advance_fetch_kernel(rnad_step_ctrl* ctrl, const float* __restrict__ src, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * i + 3 < n) {
        reinterpret_cast<float4*>(dst)[i] = __ldcv(reinterpret_cast<const float4*>(src) + i);
    } else {
        for (int k = 4 * i; k < n; ++k) dst[k] = __ldcv(src + k);
    }
    if (i == 0) {
        uint64_t state = ((uint64_t)ctrl->seed_state[1] << 32 | ctrl->seed_state[0]) + 0x9E3779B97F4A7C15ull;
        ctrl->seed_state[0] = (uint32_t)state;
        ctrl->seed_state[1] = (uint32_t)(state >> 32);
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        ctrl->seed = (z ^ (z >> 31)) >> 2;
    }
}

// The tail runs as ONE thread-block cluster of kTailCtas CTAs (8192 threads: one or two parameters per thread, so
// every pass is a single round of independent loads instead of a latency-bound loop) synchronised by the hardware
// cluster barrier; the squared gradient norm is reduced per CTA and the CTAs' partial sums are read by every CTA
// from its peers' shared memory (DSMEM) in rank order - the same value, bit for bit, in every thread.
constexpr int kTailCtas = 8;

__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ float ld_dsmem(const float* local_smem_ptr, uint32_t cta) {
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(local_smem_ptr), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(cta));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}

__global__ void __cluster_dims__(kTailCtas, 1, 1) __launch_bounds__(kTailThreads, 1) learner_tail_kernel(rnad_tail_args a) {
    __shared__ float s_red[kTailThreads / 32];
    __shared__ float s_partial;
    __shared__ float s_stats[8];
    __shared__ float s_adam[3];
    const int tid = threadIdx.x;
    const uint32_t cta = cluster_ctarank();
    const int gtid = (int)cta * kTailThreads + tid, n_threads = kTailCtas * kTailThreads;
    const int P = a.n_params;
    const int64_t row_floats = xchg_row_floats(P);
    const uint32_t seq = a.ctrl->seq;           // (rewritten only after the last cluster barrier)
    const int slot = (int)(seq & 1u);

    // this rank's row: G_0 | G_1 | counts | loss numerators
    auto my_value = [&](int i) -> float {
        if (i < 2 * P) return a.player_grads[i];
        const int k = i - 2 * P;
        if (k < 4) {
            const int32_t n = a.stats[1 + (k >> 1)];
            return (k & 1) ? (float)(n & 4095) : (float)(n >> 12);     // N = 4096 hi + lo, both exact in fp32 - and so are their sums
        }
        return k < 8 ? a.loss_sums[k - 4] : 0.f;
    };

    if (a.world > 1) {
        // ---- push the row to every rank's buffer (own included), then the flags
        for (int i = gtid; i < 2 * P + 8; i += n_threads) {
            const float v = my_value(i);
            for (int r = 0; r < a.world; ++r) a.xchg[r][((int64_t)slot * a.world + a.rank) * row_floats + i] = v;
        }
        __threadfence_system();
        cluster_sync();
        if (cta == 0 && tid < a.world && tid != a.rank) {
            uint32_t* flags = reinterpret_cast<uint32_t*>(a.xchg[tid] + xchg_flag_offset_floats(P, a.world));
            st_release_sys(flags + slot * a.world + a.rank, seq + 1u);
        }
        // ---- wait for every peer's row of this step (every CTA watches the flags itself)
        if (tid < a.world && tid != a.rank) {
            const uint32_t* flags = reinterpret_cast<const uint32_t*>(a.xchg[a.rank] + xchg_flag_offset_floats(P, a.world));
            const uint64_t t0 = global_timer_ns();
            while (ld_acquire_sys(flags + slot * a.world + tid) != seq + 1u) {
                if (global_timer_ns() - t0 > 10000000000ull) {      // 10 s: a rank is gone; do not hang the device
                    atomicOr(&a.ctrl->error, 1u << (tid & 31));
                    break;
                }
                __nanosleep(100);
            }
        }
        __syncthreads();
    }
    // sum over ranks of element i of the rows, in rank order (L1 is not coherent with the peers' stores: ld.cg)
    auto total = [&](int i) -> float {
        if (a.world == 1) return my_value(i);
        float acc = 0.f;
        for (int r = 0; r < a.world; ++r) acc += __ldcg(a.xchg[a.rank] + ((int64_t)slot * a.world + r) * row_floats + i);
        return acc;
    };
    // ---- pass 1: the gradient (vtrace.py:370-374, 387-389: each player's sum over its own step count) and its norm.
    //      Everything pass 2 needs is fetched here as well, and BEFORE the block waits for the step counts and Adam's
    //      scalars (two double-precision pow on one thread): one round of independent loads per thread, in flight
    //      under that arithmetic.
    constexpr int kMaxPer = 4;                   // up to 32,768 parameters
    float g[kMaxPer], g1[kMaxPer], m[kMaxPer], v[kMaxPer], p[kMaxPer], t[kMaxPer];
#pragma unroll
    for (int u = 0; u < kMaxPer; ++u) {
        const int i = gtid + u * n_threads;
        g[u] = g1[u] = m[u] = v[u] = p[u] = t[u] = 0.f;
        if (i < P) {
            g[u] = total(i);
            g1[u] = total(P + i);
            m[u] = a.exp_avg[i];
            v[u] = a.exp_avg_sq[i];
            p[u] = a.params[i];
            t[u] = a.target_params[i];
        }
    }
    if (tid < 8) s_stats[tid] = total(2 * P + tid);
    if (tid == 32) {     // Adam's scalars, once per CTA, in double like torch's host-side arithmetic
        const double step = (double)a.ctrl->adam_step + 1.0;
        const double bc1 = 1.0 - pow((double)a.beta1, step), bc2 = 1.0 - pow((double)a.beta2, step);
        s_adam[0] = (float)((double)a.lr / bc1);
        s_adam[1] = (float)sqrt(bc2);
        s_adam[2] = (float)step;
    }
    __syncthreads();
    const float n0 = fmaxf(s_stats[0] * 4096.f + s_stats[1], 1.f), n1 = fmaxf(s_stats[2] * 4096.f + s_stats[3], 1.f);
    float sq = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxPer; ++u) g[u] = g[u] / n0 + g1[u] / n1;      // (0 / n + 0 / n = 0 beyond the last parameter)
#pragma unroll
    for (int u = 0; u < kMaxPer; ++u) sq = fmaf(g[u], g[u], sq);
    sq = warp_sum(sq);
    if ((tid & 31) == 0) s_red[tid >> 5] = sq;
    __syncthreads();
    if (tid == 0) {
        float tot = 0.f;
        for (int w = 0; w < kTailThreads / 32; ++w) tot += s_red[w];
        s_partial = tot;
    }
    cluster_sync();
    float norm_sq = 0.f;
    for (uint32_t c = 0; c < kTailCtas; ++c) norm_sq += ld_dsmem(&s_partial, c);
    const float norm = sqrtf(norm_sq);
    // clip_grad_norm_ (rnad.py:456): coefficient min(max_norm / (norm + 1e-6), 1)
    const float coef = fminf(a.grad_clip / (norm + 1e-6f), 1.f);

    // ---- pass 2: Adam (torch.optim.Adam, no amsgrad / weight decay; rnad.py:514), then the target-net average (:516-523)
    const float step_size = s_adam[0], bc2_sqrt = s_adam[1];
    const float w1 = 1.f - a.beta1, w2 = 1.f - a.beta2;
    const float avg_new = a.gamma_averaging, avg_old = a.one_minus_gamma_averaging;
#pragma unroll
    for (int u = 0; u < kMaxPer; ++u) {
        const int i = gtid + u * n_threads;
        if (i < P) {
            const float gc = g[u] * coef;
            const float mm = w1 < 0.5f ? m[u] + w1 * (gc - m[u]) : gc - (gc - m[u]) * (1.f - w1);      // Tensor.lerp_
            const float vv = v[u] * a.beta2 + w2 * gc * gc;
            const float denom = sqrtf(vv) / bc2_sqrt + a.eps;
            const float pp = p[u] - step_size * (mm / denom);
            a.flat_grad[i] = gc;
            a.exp_avg[i] = mm;
            a.exp_avg_sq[i] = vv;
            a.params[i] = pp;
            a.target_params[i] = __fadd_rn(__fmul_rn(avg_new, pp), __fmul_rn(avg_old, t[u]));
        }
    }
    cluster_sync();                              // every CTA has read ctrl and its peers' partial sums
    if (cta == 0 && tid == 0) {
        // loss = sum over players of (its numerator / its step count); the NeuRD loss carries a minus sign (vtrace.py:429)
        const float loss_v = s_stats[4] / n0 + s_stats[5] / n1, loss_nerd = -(s_stats[6] / n0 + s_stats[7] / n1);
        const float err = (float)a.ctrl->error;
        a.losses[0] = loss_v;
        a.losses[1] = loss_nerd;
        a.losses[2] = norm;
        a.losses[3] = err;
        if (a.losses_host != nullptr) {
            a.losses_host[0] = loss_v;
            a.losses_host[1] = loss_nerd;
            a.losses_host[2] = norm;
            a.losses_host[3] = err;
        }
        a.ctrl->seq = seq + 1u;
        a.ctrl->adam_step = s_adam[2];
    }
}

}  // namespace
}  // namespace rnad

using namespace rnad;

extern "C" {

int rnad_step_control(rnad_step_ctrl* ctrl, uint64_t seed, float alpha, void* stream) {
    RNAD_REQUIRE(ctrl != nullptr, "rnad_step_control: null pointer");
    step_control_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctrl, seed, alpha);
    RNAD_CHECK_LAUNCH("step_control_kernel");
    return RNAD_OK;
}

int rnad_step_advance(rnad_step_ctrl* ctrl, void* stream) {
    RNAD_REQUIRE(ctrl != nullptr, "rnad_step_advance: null pointer");
    step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctrl);
    RNAD_CHECK_LAUNCH("step_advance_kernel");
    return RNAD_OK;
}

int rnad_step_advance_fetch(rnad_step_ctrl* ctrl, const float* src, float* dst, int64_t n, void* stream) {
    RNAD_REQUIRE(ctrl != nullptr && src != nullptr && dst != nullptr, "rnad_step_advance_fetch: null pointer");
    RNAD_REQUIRE(n >= 1 && n <= (1 << 30), "rnad_step_advance_fetch: %lld floats", (long long)n);
    RNAD_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
                 "rnad_step_advance_fetch: src and dst must be 16-byte aligned");
    const int threads = 256, quads = (int)((n + 3) / 4);
    step_advance_fetch_kernel<<<(quads + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(ctrl, src, dst, (int)n);
    RNAD_CHECK_LAUNCH("step_advance_fetch_kernel");
    return RNAD_OK;
}

int rnad_learner_tail(const rnad_tail_args* args, void* stream) {
    RNAD_REQUIRE(args != nullptr, "rnad_learner_tail: null pointer");
    RNAD_REQUIRE(args->n_params >= 1 && args->player_grads && args->stats && args->loss_sums && args->params &&
                     args->target_params && args->exp_avg && args->exp_avg_sq && args->flat_grad && args->losses &&
                     args->ctrl,
                 "rnad_learner_tail: null tensor pointer");
    RNAD_REQUIRE(args->world >= 1 && args->world <= RNAD_MAX_PEERS && args->rank >= 0 && args->rank < args->world,
                 "rnad_learner_tail: rank %d of %d (at most %d ranks)", args->rank, args->world, RNAD_MAX_PEERS);
    if (args->world > 1)
        for (int r = 0; r < args->world; ++r) RNAD_REQUIRE(args->xchg[r] != nullptr, "rnad_learner_tail: exchange buffer of rank %d is null", r);
    RNAD_REQUIRE(args->n_params <= 4 * kTailCtas * kTailThreads, "rnad_learner_tail: at most %d parameters", 4 * kTailCtas * kTailThreads);
    learner_tail_kernel<<<kTailCtas, kTailThreads, 0, (cudaStream_t)stream>>>(*args);
    RNAD_CHECK_LAUNCH("learner_tail_kernel");
    return RNAD_OK;
}

int64_t rnad_xchg_bytes(int n_params, int world) {
    if (n_params < 1 || world < 1 || world > RNAD_MAX_PEERS) return 0;
    return (xchg_flag_offset_floats(n_params, world) + 2 * world) * 4 + 256;
}

int rnad_xchg_create(int64_t bytes, void** ptr, unsigned char* handle) {
    RNAD_REQUIRE(ptr && handle && bytes > 0, "rnad_xchg_create: bad argument");
    int rc = check_cuda(cudaMalloc(ptr, (size_t)bytes), "cudaMalloc(exchange buffer)");
    if (rc) return rc;
    rc = check_cuda(cudaMemset(*ptr, 0, (size_t)bytes), "cudaMemset(exchange buffer)");
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    rc = check_cuda(cudaIpcGetMemHandle(&h, *ptr), "cudaIpcGetMemHandle");
    if (rc) return rc;
    static_assert(sizeof(h) == RNAD_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(handle, &h, sizeof(h));
    return RNAD_OK;
}

int rnad_xchg_open(const unsigned char* handle, void** ptr) {
    RNAD_REQUIRE(ptr && handle, "rnad_xchg_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return check_cuda(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

int rnad_xchg_close(void* ptr) { return ptr ? check_cuda(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle") : RNAD_OK; }

int rnad_xchg_destroy(void* ptr) { return ptr ? check_cuda(cudaFree(ptr), "cudaFree(exchange buffer)") : RNAD_OK; }

}  // extern "C"
