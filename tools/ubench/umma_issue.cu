#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred q;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t"
                 "@q bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}
// whole warp executes; the elected lane issues
__device__ __forceinline__ void umma_w(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\telect.sync _|e, 0xffffffff;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, {%7, %7, %7, %7}, p;\n\t}\n"
                 :: "r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(1u), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit_w(uint32_t bar) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(bar) : "memory");
}
__global__ void __launch_bounds__(576, 1) k3(int layers, long long* out, int commits, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[12];
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 160 * 1024 / 16; i += 576) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int c = 0; c < 11; ++c) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar[c])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);
    if (warp == 1) {
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
                    umma_w(dst + (uint32_t)(16 * (c == 0 ? 0 : c - 1) - (c == 9 ? 16 : 0)), alo + (uint32_t)(c * 272 + dy), ahi, blo + (uint32_t)(dy * 96), bhi, id48);
                if (commits) commit_w(bar + 8 * c);
            }
        }
        commit_w(bar + 80);
        mbar_wait(bar + 80, 0);
        if (tid == 32) out[blockIdx.x] = clock64() - t0;
    } else if (warp >= 2) {
        // interference from 16 "epilogue" warps while the tensor pipe runs
        const uint32_t bar = smem_u32(&s_bar[0]);
        const int et = tid - 64;
        if (mode == 1) {            // all wait on the final barrier (try_wait polling)
            mbar_wait(bar + 80, 0);
        } else if (mode == 2) {     // operand-like stores: 2 x 16 B per thread, then proxy fence, repeated
            uint4 v = make_uint4(et, et, et, et);
            for (int it = 0; it < layers * 3; ++it) {
                uint8_t* p = smem + 110 * 1024 + ((it & 7) * 4352) + (et & 127) * 16;
                *reinterpret_cast<uint4*>(p) = v;
                *reinterpret_cast<uint4*>(p + 2176) = v;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __nanosleep(200);
            }
        } else if (mode == 3) {     // TMEM loads of 16 columns per thread, repeated
            const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 320u;
            uint32_t acc = 0;
            for (int it = 0; it < layers * 3; ++it) {
                uint32_t r[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta + (uint32_t)((it & 7) * 16)) : "memory");
                acc += r[0] + r[15];
                __nanosleep(200);
            }
            if (acc == 0x12345u) out[1000] = acc;
        } else if (mode == 4) {     // stores only, no proxy fence
            uint4 v = make_uint4(et, et, et, et);
            for (int it = 0; it < layers * 3; ++it) {
                uint8_t* p = smem + 110 * 1024 + ((it & 7) * 4352) + (et & 127) * 16;
                *reinterpret_cast<uint4*>(p) = v;
                *reinterpret_cast<uint4*>(p + 2176) = v;
                __nanosleep(200);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}
int main() {
    long long* d; cudaMalloc(&d, 1024 * sizeof(long long));
    long long h[148];
    cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int mode = 0; mode < 5; ++mode) {
        const int commits = 1;
        k3<<<148, 576, 160 * 1024>>>(200, d, commits, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("k3 err %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
        printf("warp-uniform issue, interference mode=%d: %.1f clk per layer (30 MMAs N=48; smem bound 1320) -> %.1f clk/MMA\n", mode, s / 148 / 200, s / 148 / 200 / 30);
    }
    return 0;
}
