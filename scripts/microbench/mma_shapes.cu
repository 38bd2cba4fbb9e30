// Microbenchmark (development aid): cycles per tcgen05.mma for small shapes, A from smem (SS) or TMEM (TS).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_shapes mma_shapes.cu && ./mma_shapes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sbo(uint32_t addr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
__host__ __device__ constexpr uint32_t idesc_bf16(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }

template <int KIND>   // 0 tf32, 1 bf16
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{.reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p;}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
}

// MODE 0: SS, 1: TS.  N = MMA N.  REP MMAs back to back, all accumulating into the same D (or DISTINCT D columns).
template <int KIND, int MODE, int N, int REP, bool DISTINCT>
__global__ void bench(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i % 97);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x < 32) {
        const uint64_t a_desc = desc_sbo(smem_u32(smem), 256), b_desc = desc_sbo(smem_u32(smem) + 16384, 256);
        const uint32_t idesc = KIND == 0 ? idesc_tf32(N) : idesc_bf16(N);
        for (int round = 0; round < 3; ++round) {
            long long t0 = clock64(), t1 = 0;
            if (elect_one()) {
#pragma unroll
                for (int r = 0; r < REP; ++r) {
                    const uint32_t d = tmem + 256 + (DISTINCT ? (r % (256 / N)) * N : 0);
                    if (MODE == 0) mma_ss<KIND>(d, a_desc + (r & 3) * 16, b_desc + (r & 7) * 16, idesc, r >= (DISTINCT ? 256 / N : 1));
                    else mma_ts<KIND>(d, tmem + (r & 15) * 8, b_desc + (r & 7) * 16, idesc, r >= (DISTINCT ? 256 / N : 1));
                }
                t1 = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            mbar_wait(smem_u32(&bar), round & 1);
            long long t2 = clock64();
            if (t1 && round == 2) {
                out[0] = t1 - t0;
                out[1] = t2 - t0;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int KIND, int MODE, int N, int REP, bool DISTINCT>
void run(const char* name, int grid = 1) {
    long long* d;
    cudaMalloc(&d, 16);
    cudaMemset(d, 0, 16);
    auto k = bench<KIND, MODE, N, REP, DISTINCT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<<<grid, 128, 64 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-34s N=%3d rep=%3d grid=%3d  issue %6lld cyc (%5.1f/mma)  done %6lld cyc (%5.1f/mma)  %s\n", name, N, REP, grid, h[0],
           (double)h[0] / REP, h[1], (double)h[1] / REP, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}


// The rollout kernel's per-chunk tensor work: 16 x (TS, N=16, K=8) + 3 x (TS, N=128, K=8), REP items back to back,
// with LDST_WARPS other warps looping tcgen05.ld.x32 + tcgen05.st.x32 over the ring columns (the relu epilogue's traffic).
template <int LDST_WARPS, int COMMITS>
__global__ void bench_mix(long long* out, int items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[4];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i % 97);
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        for (int q = 0; q < 4; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[q])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint64_t b_desc = desc_sbo(smem_u32(smem) + 16384, 256);
        long long t0 = clock64(), t1 = 0;
        if (elect_one()) {
            for (int it = 0; it < items; ++it) {
                const uint32_t slot = tmem + (it % 3) * 128;
#pragma unroll
                for (int s = 0; s < 16; ++s) mma_ts<0>(tmem + 384, slot + s * 8, b_desc + s * 16, idesc_tf32(16), s > 0);
#pragma unroll
                for (int s = 0; s < 3; ++s) mma_ts<0>(slot, tmem + 400 + s * 8, b_desc + s * 16, idesc_tf32(128), s > 0);
                if (COMMITS >= 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[it % 3])) : "memory");
                if (COMMITS >= 2 && (it & 3) == 3) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[3])) : "memory");
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (t1) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
        stop = 1;
    } else if (warp <= LDST_WARPS) {
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((warp - 1) >> 2) * 64;
        long long n = 0;
        while (!stop) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr + (uint32_t)(n & 1) * 32)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr + (uint32_t)(n & 1) * 32),
                "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
                "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
                "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            ++n;
        }
        if ((threadIdx.x & 31) == 0) out[2 + warp] = n;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int LDST_WARPS, int COMMITS>
void run_mix(int items) {
    long long* d;
    cudaMalloc(&d, 8 * 32);
    cudaMemset(d, 0, 8 * 32);
    auto k = bench_mix<LDST_WARPS, COMMITS>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<<<1, 32 * (1 + LDST_WARPS), 64 * 1024>>>(d, items);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[32] = {0};
    cudaMemcpy(h, d, 8 * 32, cudaMemcpyDeviceToHost);
    long long iters = 0;
    for (int w = 1; w <= LDST_WARPS; ++w) iters += h[2 + w];
    printf("item mix (16 x TS N=16 + 3 x TS N=128), %d commits/item, %2d ld/st warps: issue %6.1f cyc/item, done %6.1f cyc/item; "
           "concurrent ld+st x32 pairs per item: %.2f %s\n", COMMITS, LDST_WARPS, (double)h[0] / items, (double)h[1] / items,
           (double)iters / items, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run_mix<0, 0>(64);
    run_mix<0, 1>(64);
    run_mix<0, 2>(64);
    run_mix<8, 0>(64);
    run_mix<8, 2>(64);
    run<0, 1, 128, 128, false>("tf32 TS same D");
    run<0, 1, 256, 128, false>("tf32 TS same D");

    run<0, 0, 16, 128, false>("tf32 SS same D");
    run<0, 0, 32, 128, false>("tf32 SS same D");
    run<0, 0, 64, 128, false>("tf32 SS same D");
    run<0, 0, 128, 128, false>("tf32 SS same D");
    run<0, 0, 256, 128, false>("tf32 SS same D");
    run<0, 0, 16, 128, true>("tf32 SS distinct D");
    run<0, 0, 64, 128, true>("tf32 SS distinct D");
    run<0, 1, 16, 128, false>("tf32 TS same D");
    run<0, 1, 32, 128, false>("tf32 TS same D");
    run<0, 1, 64, 128, false>("tf32 TS same D");
    run<0, 1, 16, 128, true>("tf32 TS distinct D");
    run<1, 1, 16, 128, false>("bf16 TS same D");
    run<1, 1, 64, 128, false>("bf16 TS same D");
    run<1, 0, 16, 128, false>("bf16 SS same D");
    run<1, 0, 256, 128, false>("bf16 SS same D");
    run<0, 1, 16, 128, false>("tf32 TS same D, 148 CTAs", 148);
    run<0, 1, 16, 16, false>("tf32 TS same D");
    run<0, 1, 16, 8, false>("tf32 TS same D");
    return 0;
}
