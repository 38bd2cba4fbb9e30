// tcgen05 / TMEM / mbarrier / TMA helpers shared by the tensor-core kernels (sm_100a inline PTX).
#pragma once

#include "common.cuh"

namespace rnad {
namespace tc {

constexpr int kTileM = 128;      // rows (games / trajectory steps) per CTA tile == TMEM lanes
constexpr int kHidden = 256;     // width of each trunk == N of one MMA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// The same rounding (nearest, ties away from zero) for finite inputs with two integer instructions: ptxas expands
// cvt.rna.tf32.f32 into a branchy sequence with Inf / NaN handling, which matters where it runs per matrix element.
__device__ __forceinline__ float to_tf32_fast(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// canonical K-major, no-swizzle operand layout: 8-row x 16-byte core matrices,
// K chunks of a row group adjacent (LBO = 128 B), row groups SBO bytes apart
template <int KP>
__host__ __device__ __forceinline__ uint32_t operand_offset(int row, int k) {
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
                                | ((uint32_t)(kHidden >> 3) << 17)
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
        if (!done && clock64() - start > 4000000000LL) __trap();   // a lost completion must not hang the device
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

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one TMA bulk copy global -> shared, completion counted in bytes on `mbar`
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(mbar)
                 : "memory");
}

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)),
                 "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kCols) : "memory");
}

// first-layer weights of one trunk (W, KIN) + bias -> tf32 B operand [256 x KP] in core-matrix order
template <int KIN, int KP, bool kBiasInK>
__device__ __forceinline__ void pack_trunk_operand(const float* __restrict__ w, const float* __restrict__ bias,
                                                   uint8_t* __restrict__ dst, int thread, int n_threads) {
    for (int e = thread; e < kHidden * KP; e += n_threads) {
        const int j = e / KP, k = e % KP;
        float v = 0.f;
        if (k < KIN) v = w[j * KIN + k];
        else if (kBiasInK && k == KIN) v = bias[j];
        *reinterpret_cast<float*>(dst + operand_offset<KP>(j, k)) = to_tf32(v);
    }
}

// relu(acc [+ bias]) dotted with the second-layer weights, for one 32-column chunk
template <int A, bool kValuePass, bool kAddBias>
__device__ __forceinline__ void consume_chunk(const uint32_t (&r)[32], int col, const float* __restrict__ b1,
                                              const float* __restrict__ w2v, const float4* __restrict__ w2p,
                                              float (&vacc)[4], float (&lacc)[2][4]) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        float h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = __uint_as_float(r[i + u]);
        if (kAddBias) {
            const float4 b = *reinterpret_cast<const float4*>(b1 + col + i);
            h[0] += b.x;
            h[1] += b.y;
            h[2] += b.z;
            h[3] += b.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = fmaxf(h[u], 0.f);
        if (kValuePass) {
            const float4 w = *reinterpret_cast<const float4*>(w2v + col + i);
            vacc[0] = fmaf(h[0], w.x, vacc[0]);
            vacc[1] = fmaf(h[1], w.y, vacc[1]);
            vacc[2] = fmaf(h[2], w.z, vacc[2]);
            vacc[3] = fmaf(h[3], w.w, vacc[3]);
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 w = w2p[col + i + u];
                float(&acc)[4] = lacc[u & 1];
                acc[0] = fmaf(h[u], w.x, acc[0]);
                if constexpr (A > 1) acc[1] = fmaf(h[u], w.y, acc[1]);
                if constexpr (A > 2) acc[2] = fmaf(h[u], w.z, acc[2]);
                if constexpr (A > 3) acc[3] = fmaf(h[u], w.w, acc[3]);
            }
        }
    }
}

}  // namespace tc
}  // namespace rnad
