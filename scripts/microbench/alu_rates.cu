// Microbenchmark (development aid): issue rate of the conversion / pack instructions the fp16 epilogues use,
// against plain FP32 and min/max instructions.  One CTA of 4 * W warps per SM; cycles per warp-instruction per SMSP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rates alu_rates.cu && ./alu_rates
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b) {
    uint32_t d;
    if (OP == 0) asm volatile("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
    if (OP == 1) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
    if (OP == 2) asm volatile("add.f32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    if (OP == 3) asm volatile("max.f32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    if (OP == 4) asm volatile("set.gt.f16x2.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    if (OP == 5) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
    if (OP == 6) { unsigned short h; asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(__uint_as_float(a))); d = h; }
    if (OP == 7) asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    if (OP == 8) asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(__uint_as_float(a)));
    return d;
}

template <int OP>
__global__ void bench(long long* out, uint32_t seed) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = seed + threadIdx.x * 16 + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = op<OP>(r[i], r[(i + 5) & 15]);
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= r[i];
    if (acc == 0x12345678u) out[1] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name) {
    long long* d;
    cudaMalloc(&d, 16);
    for (int warps = 4; warps <= 16; warps *= 2) {
        bench<OP><<<1, warps * 32>>>(d, 1);
        bench<OP><<<1, warps * 32>>>(d, 1);
        long long h = 0;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / (256.0 * 16.0 * (warps / 4));   // cycles per warp-instruction per SMSP
        printf("%-36s warps/SMSP %d : %6.2f cycles per warp instruction\n", name, warps / 4, per);
    }
    cudaFree(d);
}

int main() {
    run<2>("add.f32 (FADD)");
    run<3>("max.f32 (FMNMX)");
    run<0>("cvt.rn.relu.f16x2.f32 (F2FP.RELU.PACK)");
    run<1>("cvt.rn.f16x2.f32 (F2FP.PACK)");
    run<5>("cvt.rn.satfinite.f16x2.f32");
    run<6>("cvt.rn.f16.f32 (single)");
    run<4>("set.gt.f16x2.f16x2 (HSET2)");
    run<7>("max.f16x2 (HMNMX2)");
    run<8>("cvt.rna.tf32.f32");
    return 0;
}
