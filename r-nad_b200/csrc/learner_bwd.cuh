// Declarations shared by the learner's backward kernels (learner_mlp.cu, learner_bwd_tc3.cu): parameter layout, row
// loads, operand tiles, descriptor / tensor-memory helpers.
#pragma once

#include "tc_common.cuh"
#include "tc_pipe.cuh"

namespace rnad {
namespace tc {

constexpr int kMaxBwdCtas = 160;     // per-CTA partial gradients the workspace holds

template <int A>
struct Shape {
    static constexpr int KIN = 2 * A * A;
    static constexpr bool kBiasInK = (KIN % 8) != 0;
    static constexpr int KP = round_up(KIN + (kBiasInK ? 1 : 0), 8);
    static constexpr int kTrunkBytes = kHidden * KP * 4;
    // number of learner parameters, in state_dict order:
    // value_fc0.{weight,bias}, value_fc1.{weight,bias}, policy_fc0.{weight,bias}, policy_fc1.{weight,bias}
    static constexpr int kOffV0w = 0;
    static constexpr int kOffV0b = kOffV0w + kHidden * KIN;
    static constexpr int kOffV1w = kOffV0b + kHidden;
    static constexpr int kOffV1b = kOffV1w + kHidden;
    static constexpr int kOffP0w = kOffV1b + 1;
    static constexpr int kOffP0b = kOffP0w + kHidden * KIN;
    static constexpr int kOffP1w = kOffP0b + kHidden;
    static constexpr int kOffP1b = kOffP1w + A * kHidden;
    static constexpr int kParams = kOffP1b + A;
};

// row of the observation tensor -> registers (8-byte loads; KIN is even)
template <int KIN>
__device__ __forceinline__ void load_row(const float* __restrict__ obs, int64_t row, bool active, float (&x)[KIN]) {
    const float2* src = reinterpret_cast<const float2*>(obs + row * KIN);
#pragma unroll
    for (int i = 0; i < KIN / 2; ++i) {
        const float2 v = active ? __ldg(src + i) : make_float2(0.f, 0.f);
        x[2 * i] = v.x;
        x[2 * i + 1] = v.y;
    }
}

template <int KIN, int KP, bool kBiasInK>
__device__ __forceinline__ void store_operand_row(uint8_t* tile, int lane, const float (&x)[KIN]) {
#pragma unroll
    for (int q = 0; q < KP / 4; ++q) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = 4 * q + u;
            v[u] = k < KIN ? to_tf32_fast(x[k < KIN ? k : 0]) : ((kBiasInK && k == KIN) ? 1.f : 0.f);
        }
        *reinterpret_cast<float4*>(tile + operand_offset<KP>(lane, 4 * q)) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ void mma_ss_n(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}
__device__ __forceinline__ uint64_t desc_lbo_sbo(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}


// The "mask" formulation of the pipelined backward (learner_bwd_tc3.cu), the default where its accumulators fit
// tensor memory (max_actions <= 3).  `image`: the image of pack_bwd_tc_image_kernel (first layers + biases are read);
// leaves `blocks` per-CTA partial gradients in `partials` (split mode: CTAs of even / odd index serve player 0 / 1).
bool learner_backward_tc3_supported(int A);
int learner_backward_tc3(int A, const float* obs, int64_t N, int T_split, int64_t B_split, const rnad_mlp_weights& w,
                         const float* d_logit, const float* d_v, const uint8_t* image, float* partials, int blocks,
                         cudaStream_t st);

// The fp16-operand engine (learner_bwd_f16.cu), for UNNORMALISED gradients (split mode).  `image`: its own share of
// the workspace, learner_backward_f16_image_bytes(A) bytes; mode: 0 = pack the weight image, then run; 1 = prepacked;
// 2 = pack only.
bool learner_backward_f16_supported(int A);
int64_t learner_backward_f16_image_bytes(int A);
int learner_backward_f16(int A, const float* obs, int64_t N, int T_split, int64_t B_split, const rnad_mlp_weights& w,
                         const float* d_logit, const float* d_v, uint8_t* image, float* partials, int blocks,
                         cudaStream_t st, int mode);

}  // namespace tc
}  // namespace rnad
