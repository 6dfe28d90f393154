// umma_rate.cu — micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16) with both operands in
// shared memory (SS), M = 128, K = 16, for several N.  One CTA per SM, one issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t"
                 "@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

// mode 0: same A every MMA; mode 1: A start address walks (distinct 4 KB tiles, 16 B shifts)
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int a_tiles, long long* out, int ctas_per_sm) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < 160 * 1024 / 16 / ctas_per_sm; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const uint32_t bar = smem_u32(&s_bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512 / ctas_per_sm));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 150 * 1024 / ctas_per_sm;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            // A: 128 rows x 16 B per K-half; halves 2080 B apart; tiles walk by 4160 B
            const uint32_t a = a0 + (uint32_t)(i % a_tiles) * 4160u + (uint32_t)(i % 3) * 16u;
            umma(tmem + (uint32_t)((i % 4) * 16), umma_desc(a, 2080u, 128u), umma_desc(b0, (uint32_t)N * 16u, 128u), idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
        mbar_wait(bar, 0);
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512 / ctas_per_sm));
}

__device__ __forceinline__ void umma2(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, {%7, %7, %7, %7}, p;\n\t}\n"
                 :: "r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(1u), "r"(0u) : "memory");
}

// lean issue loop shaped like trunk_rows: per layer 10 columns x 3 MMAs (N=48), one commit per column
__global__ void __launch_bounds__(128, 1) k2(int layers, long long* out, int commits) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[12];
    const int tid = threadIdx.x;
    for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int c = 0; c < 11; ++c) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar[c])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint64_t ad = umma_desc(smem_u32(smem), 2176u, 128u), bd = umma_desc(smem_u32(smem) + 100 * 1024, 768u, 128u);
        const uint32_t alo = (uint32_t)ad, ahi = (uint32_t)(ad >> 32), blo0 = (uint32_t)bd, bhi = (uint32_t)(bd >> 32);
        const uint32_t id48 = (1u << 4) | (1u << 7) | (1u << 10) | ((48u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t bar = smem_u32(&s_bar[0]);
        long long t0 = clock64();
        for (int l = 0; l < layers; ++l) {
            const uint32_t blo = blo0 + (uint32_t)(l % 8) * 288u;
            const uint32_t dst = tmem + ((l & 1) ? 0u : 160u);
#pragma unroll
            for (int c = 0; c < 10; ++c) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
                    umma2(dst + (uint32_t)(16 * (c == 0 ? 0 : c - 1) - (c == 9 ? 16 : 0)), alo + (uint32_t)(c * 272 + dy), ahi, blo + (uint32_t)(dy * 96), bhi, id48);
                if (commits) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar + 8 * c) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar + 80) : "memory");
        mbar_wait(bar + 80, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

int main() {
    {
        long long* d; cudaMalloc(&d, 1024 * sizeof(long long));
        long long h[148];
        cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        for (int commits = 0; commits < 2; ++commits) {
            k2<<<148, 128, 160 * 1024>>>(200, d, commits);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("k2 err %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
            printf("lean issue, commits=%d: %.1f clk per layer (30 MMAs N=48; smem bound 1320) -> %.1f clk/MMA\n", commits, s / 148 / 200, s / 148 / 200 / 30);
        }
        cudaFree(d);
    }

    long long* d; cudaMalloc(&d, 1024 * sizeof(long long));
    long long h[1024];
    const int iters = 4000;
    for (int cps = 1; cps <= 4; cps *= 2) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024 / cps);
        for (int N : {16, 32, 48, 64, 96, 128, 160, 256}) {
            if (cps > 1 && N > 64) continue;
            for (int tiles : {1, 8}) {
                k<<<148 * cps, 128, 160 * 1024 / cps>>>(N, iters, tiles, d, cps);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("N=%d err %s\n", N, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 148 * cps * sizeof(long long), cudaMemcpyDeviceToHost);
                double s = 0; for (int i = 0; i < 148 * cps; ++i) s += h[i];
                printf("ctas/SM=%d N=%3d a_tiles=%d : %.1f clk/MMA per CTA (%.1f clk/MMA per SM), floor %d, smem-bytes %d -> %.1f B/clk/SM\n",
                       cps, N, tiles, s / (148 * cps) / iters, s / (148 * cps) / iters / cps, N / 2, 4096 + N * 32,
                       (4096 + N * 32) / (s / (148 * cps) / iters / cps));
            }
        }
    }
    return 0;
}
