// trunk.cu — fused convolutional trunk of AlphaSame (filters = 16, kernels = 1) on the
// 5th-generation tensor cores: tcgen05.mma with TMEM accumulators, activations resident in
// shared memory across all 2*blocks convolutions of an image.
//
// Replaces AlphaSame.process_grid (reference architectures.py:120-126): conv1 5x5 (1->16) ->
// blocks x [BN-ReLU-Conv3x3, BN-ReLU-Conv3x3, +skip] -> BN-ReLU -> Conv1x1 (16->1) -> BN-ReLU ->
// flatten(400), eval mode (BatchNorm folded to per-channel scale/bias on the host).
//
// Mapping (one CTA = one warpgroup = one image at a time, persistent over images):
//   * An image is kept as a zero-haloed 42 x 12 grid of "pixels"; pixel p = (y+1)*12 + (x+1).
//     An activation buffer holds two planes [pixel][8 channels] bf16 (16 B per pixel per plane):
//     this IS the canonical K-major no-swizzle UMMA operand layout (8 consecutive pixels x 16 B =
//     one 128 B core matrix), so the im2col view of tap (dy,dx) is just the same buffer with the
//     descriptor start address moved by (12*dy + dx) pixels.  No im2col copy exists.
//   * A 3x3 convolution of one image = 4 M-tiles (128 pixels each, p in [13, 525)) x 9 taps of
//     tcgen05.mma M=128 N=16 K=16 (bf16 x bf16 -> fp32 in TMEM), issued by one thread;
//     accumulators: 4 tiles x 16 TMEM columns.
//   * Epilogue: thread t of the warpgroup owns pixel 13 + 128*m + t of every tile m
//     (tcgen05.ld 32x32b.x16 gives it the 16 output channels of its pixel), applies bias/ReLU or
//     the residual add, and writes the next layer's operand (bf16) back in the same layout.  The
//     residual stream X stays in fp32 REGISTERS for the whole network (64 registers).
//   * The binary 5x5 stem is a table lookup: per kernel row, the 5 input bits select a
//     precomputed 16-channel partial sum (5 x 32 x 16 floats).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "trl_common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kPadW = 12;                     // 10 columns + 1 halo each side
constexpr int kFirstPixel = 13;               // (y=0, x=0)
constexpr int kBufPixels = 544;               // >= 4*128 + 2*12 + 2 + 1, multiple of 8
constexpr int kPlaneBytes = kBufPixels * 16;  // 8704
constexpr int kActBytes = 2 * kPlaneBytes;    // 17408: two 8-channel planes
constexpr int kWLayerBytes = 9 * 512;         // 9 taps x (16 x 16 bf16)
constexpr int kLutFloats = 5 * 32 * 16;
constexpr int kTmemCols = 64;                 // 4 tiles x 16 fp32 columns

// shared memory carve-up (bytes)
constexpr int kOffT = 0;
constexpr int kOffU = kOffT + kActBytes;
constexpr int kOffW = kOffU + kActBytes;              // double buffered
constexpr int kOffLut = kOffW + 2 * kWLayerBytes;
constexpr int kOffConst = kOffLut + kLutFloats * 4;   // per block 48 floats + 50 final
constexpr int kMaxBlocks = 40;
constexpr int kOffRows = kOffConst + (kMaxBlocks * 48 + 64) * 4;
constexpr int kOffBar = kOffRows + 48 * 4;
constexpr int kSmemBytes = kOffBar + 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (stride between the two 8-element K chunks) |
// SBO>>4 [32,46) (stride between 8-row core matrices) | version=1 [46,48) | layout NONE [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// cute::UMMA::InstrDescriptor for kind::f16: D=f32, A=B=bf16, K-major both, N=16, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a), "l"(b), "r"(kIdesc), "r"(accumulate), "r"(0u) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t"
        "@q bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&d)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// write the 16 channels of pixel p into an operand buffer (two 16-byte stores, one per plane)
__device__ __forceinline__ void store_pixel(uint8_t* buf, int p, const float (&v)[16]) {
    uint4 a, b;
    a.x = pack_bf16x2(v[0], v[1]);   a.y = pack_bf16x2(v[2], v[3]);
    a.z = pack_bf16x2(v[4], v[5]);   a.w = pack_bf16x2(v[6], v[7]);
    b.x = pack_bf16x2(v[8], v[9]);   b.y = pack_bf16x2(v[10], v[11]);
    b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
    *reinterpret_cast<uint4*>(buf + (size_t)p * 16) = a;
    *reinterpret_cast<uint4*>(buf + kPlaneBytes + (size_t)p * 16) = b;
}

// 36 MMAs of one 3x3 convolution: D[tile] = sum_taps A(tap, tile) * W(tap)
__device__ __forceinline__ void issue_conv(uint32_t act_saddr, uint32_t w_saddr, uint32_t tmem_base, uint32_t bar) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            const uint64_t a = umma_desc(act_saddr + (uint32_t)(128 * m + kPadW * dy + dx) * 16u, kPlaneBytes, 128u);
            const uint64_t b = umma_desc(w_saddr + (uint32_t)tap * 512u, 256u, 128u);
            umma_bf16(tmem_base + (uint32_t)(m * 16), a, b, tap > 0 ? 1u : 0u);
        }
    }
    umma_commit(bar);
}

__global__ void __launch_bounds__(kThreads)
alphasame_trunk_kernel(const __nv_bfloat16* __restrict__ grids, int n_images, int n_blocks,
                       const uint4* __restrict__ w_packed,   // [2*n_blocks][9*512 B]
                       const float* __restrict__ consts,     // [n_blocks*48 + 50]
                       const float* __restrict__ stem_lut,   // [5][32][16]
                       __nv_bfloat16* __restrict__ out,      // [n_images][400]
                       int* __restrict__ next_image) {       // work counter (zeroed before launch)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_img;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* bufT = smem + kOffT;
    uint8_t* bufU = smem + kOffU;
    float* s_lut = reinterpret_cast<float*>(smem + kOffLut);
    float* s_const = reinterpret_cast<float*>(smem + kOffConst);
    uint32_t* s_rows = reinterpret_cast<uint32_t*>(smem + kOffRows);
    const uint32_t bar = smem_u32(smem + kOffBar);
    const int n_layers = 2 * n_blocks;

    // ---- one-time setup ----
    for (int i = tid; i < 2 * kActBytes / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < kLutFloats; i += kThreads) s_lut[i] = stem_lut[i];
    for (int i = tid; i < n_blocks * 48 + 50; i += kThreads) s_const[i] = consts[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem_base;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);

    // pixel geometry of this thread: tile m -> padded pixel, interior flag, (y, x)
    int pix[4], py_[4], px_[4];
    bool inside[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        pix[m] = kFirstPixel + 128 * m + tid;
        py_[m] = pix[m] / kPadW - 1;
        px_[m] = pix[m] % kPadW - 1;
        inside[m] = (py_[m] >= 0 && py_[m] < 40 && px_[m] >= 0 && px_[m] < 10);
    }

    uint32_t phase = 0;
    // Images are handed out dynamically: a CTA that becomes resident late (e.g. while another
    // kernel shares the GPU) simply takes fewer images instead of delaying the whole launch.
    while (true) {
        // ---- input: 400 bf16 {0,1} -> bit rows with a 2-cell border (for the 5x5 stem) ----
        if (tid == 0) s_img = atomicAdd(next_image, 1);
        if (tid < 48) s_rows[tid] = 0;
        __syncthreads();
        const int img = s_img;
        if (img >= n_images) break;
        const __nv_bfloat16* gin = grids + (size_t)img * 400;
        for (int c = tid; c < 400; c += kThreads) {
            if (__bfloat162float(gin[c]) != 0.f) atomicOr(&s_rows[c / 10 + 2], 1u << (c % 10 + 2));
        }
        // first layer's weights
        for (int i = tid; i < kWLayerBytes / 16; i += kThreads)
            reinterpret_cast<uint4*>(smem + kOffW)[i] = w_packed[i];
        __syncthreads();

        // ---- stem: X = conv5x5(grid) by table lookup; T = relu(bn1_0(X)) ----
        float X[4][16];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            float t[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) { X[m][c] = 0.f; t[c] = 0.f; }
            if (inside[m]) {
#pragma unroll
                for (int r = 0; r < 5; ++r) {
                    const uint32_t pat = (s_rows[py_[m] + r] >> px_[m]) & 31u;
                    const float4* l = reinterpret_cast<const float4*>(s_lut + (r * 32 + pat) * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 v = l[q];
                        X[m][4 * q] += v.x; X[m][4 * q + 1] += v.y; X[m][4 * q + 2] += v.z; X[m][4 * q + 3] += v.w;
                    }
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) t[c] = fmaxf(fmaf(s_const[c], X[m][c], s_const[16 + c]), 0.f);
            }
            store_pixel(bufT, pix[m], t);
        }

        // ---- 2 * n_blocks convolutions ----
        for (int layer = 0; layer < n_layers; ++layer) {
            const bool second = layer & 1;
            uint8_t* src = second ? bufU : bufT;
            uint8_t* dst = second ? bufT : bufU;
            const uint32_t wbuf = smem_u32(smem + kOffW + (layer & 1) * kWLayerBytes);
            // operand writes (generic proxy) -> visible to the tensor core (async proxy)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                issue_conv(smem_u32(src), wbuf, tmem_base, bar);
            }
            // prefetch the next layer's weights into the other buffer while the MMAs run
            if (layer + 1 < n_layers) {
                const uint4* wn = w_packed + (size_t)(layer + 1) * (kWLayerBytes / 16);
                uint4* wd = reinterpret_cast<uint4*>(smem + kOffW + ((layer + 1) & 1) * kWLayerBytes);
                for (int i = tid; i < kWLayerBytes / 16; i += kThreads) wd[i] = wn[i];
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

            const float* cb = s_const + (layer >> 1) * 48;
            const bool last = (layer == n_layers - 1);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float d[16];
                tmem_ld16(tmem_lane + (uint32_t)(m * 16), d);
                float v[16];
                if (!second) {
                    // U = relu(conv1'(T) + c2)   (bn2 scale folded into the weights)
#pragma unroll
                    for (int c = 0; c < 16; ++c) v[c] = inside[m] ? fmaxf(d[c] + cb[32 + c], 0.f) : 0.f;
                    store_pixel(dst, pix[m], v);
                } else {
                    // X += conv2(U); T = relu(bn1_next(X))
#pragma unroll
                    for (int c = 0; c < 16; ++c) X[m][c] += d[c];
                    if (!last) {
                        const float* nb = cb + 48;
#pragma unroll
                        for (int c = 0; c < 16; ++c) v[c] = inside[m] ? fmaxf(fmaf(nb[c], X[m][c], nb[16 + c]), 0.f) : 0.f;
                        store_pixel(dst, pix[m], v);
                    }
                }
            }
        }

        // ---- head of the trunk: BN-ReLU, 1x1 conv to one channel, BN-ReLU, flatten ----
        const float* fc = s_const + n_blocks * 48;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            if (inside[m]) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) acc = fmaf(fc[32 + c], fmaxf(fmaf(fc[c], X[m][c], fc[16 + c]), 0.f), acc);
                const float y = fmaxf(fmaf(fc[48], acc, fc[49]), 0.f);
                out[(size_t)img * 400 + py_[m] * 10 + px_[m]] = __float2bfloat16(y);
            }
        }
        // all TMEM reads of this image are done before the next image's MMAs (fence + barrier above)
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols));
    }
}

}  // namespace

extern "C" int trl_alphasame_trunk(const void* grids_bf16, int n_images, int n_blocks, const void* w_packed,
                                   const float* consts, const float* stem_lut, void* out_bf16, void* stream) {
    if (n_images < 0 || n_blocks < 1 || n_blocks > kMaxBlocks || !grids_bf16 || !w_packed || !consts || !stem_lut || !out_bf16)
        return TRL_E_ARG;
    if (n_images == 0) return TRL_OK;
    static bool configured = false;
    if (!configured) {
        int rc = trl_check(cudaFuncSetAttribute(alphasame_trunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        if (rc) return rc;
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = sms * 3;  // 3 resident CTAs per SM (shared memory bound), persistent over images
    if (grid > n_images) grid = n_images;
    int* counter = (int*)trl_workspace(TRL_WS_TRUNK_COUNTER, 256);
    if (!counter) return TRL_E_NOMEM;
    int rc = trl_check(cudaMemsetAsync(counter, 0, sizeof(int), (cudaStream_t)stream));
    if (rc) return rc;
    alphasame_trunk_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)grids_bf16, n_images, n_blocks, (const uint4*)w_packed, consts, stem_lut,
        (__nv_bfloat16*)out_bf16, counter);
    return trl_check(cudaGetLastError());
}
