// Micro-benchmark: how many cycles does ONE tcgen05.mma kind::tf32 (M=128, K=8) cost when many are issued back to back?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_pacing scripts/mma_pacing.cu && /tmp/mma_pacing
// Variants: operand A from shared memory (SS) or tensor memory (TS); N = 64 / 128 / 256; all MMAs into one accumulator or
// alternating between two.  One CTA per SM, one issuing thread, operands are whatever the memories contain (timing only).
// Background: the ncu source page of tc_gemm_ts_kernel shows the issuer stalled on `mio` with the tensor pipe ~55-60 % active
// (DESIGN.md section 7); this separates per-instruction cost from accumulator dependency.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode bit 0: TS, bit 1: alternate two accumulators, bit 2: kind::f16 with bf16 operands (K = 16 per instruction; TS only),
// bit 3: 4 tf32 MMAs followed by 4 bf16 MMAs into the same accumulator (the "tf32 main term + bf16 correction terms" mix; TS only)
__global__ void __launch_bounds__(128, 1) pacing_kernel(int N, int mode, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint64_t a_desc = desc_sw128(smem_u32(smem)), b_desc = desc_sw128(smem_u32(smem + 16384));
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const bool ts = mode & 1, alt = mode & 2, f16 = mode & 4, mix = mode & 8;
        const uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_tmem = tmem + 448;                      // 32 operand columns at the top of tensor memory
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint32_t d = tmem + ((alt && (ks & 1)) ? (uint32_t)N : 0u);
                if (mix) {
                    mma_ts(d, a_tmem + ks * 8, b_desc + ks * 2, idesc, 1);
                    mma_ts_f16(d, a_tmem + 32 + ks * 8, b_desc + 512 + ks * 2, idesc16, 1);
                } else if (f16) mma_ts_f16(d, a_tmem + ks * 8, b_desc + ks * 2, idesc16, 1);
                else if (ts) mma_ts(d, a_tmem + ks * 8, b_desc + ks * 2, idesc, 1);
                else mma_ss(d, a_desc + ks * 2, b_desc + ks * 2, idesc, 1);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!done);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = (t1 - t0) / (mix ? 2 : 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, sizeof(long long));
    const int smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(pacing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 2000;
    printf("tcgen05.mma kind::tf32 M=128 K=8, %d MMAs back to back per SM, all %d SMs busy\n", iters * 4, sms);
    printf("%-4s %-5s %-12s %10s %8s\n", "A", "N", "accumulators", "cyc/MMA", "floor");
    for (int n : {64, 128, 256}) {
        for (int mode : {0, 1, 2, 3, 5, 9}) {
            if ((mode & 2) && 2 * n > 448) continue;
            if ((mode & 8) && n > 128) continue;
            pacing_kernel<<<sms, 128, smem>>>(n, mode, iters, d_out);          // warm-up
            pacing_kernel<<<sms, 128, smem>>>(n, mode, iters, d_out);
            long long cyc = 0;
            cudaError_t e = cudaMemcpy(&cyc, d_out, sizeof(cyc), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            printf("%-4s %-5d %-12s %10.1f %8d\n", (mode & 1) ? "TS" : "SS", n,
                   (mode & 8) ? "tf32+bf16 mix" : (mode & 4) ? "bf16 K=16" : (mode & 2) ? "2 alternating" : "1", (double)cyc / (iters * 4.0),
                   128 * n / 256);
        }
    }
    return 0;
}
