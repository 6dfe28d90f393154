// heads.cu — everything between the convolutional trunk and the policy GEMM of AlphaSame, fused.
//
// Replaces, in eval mode, the tail of AlphaSame.forward (reference architectures.py:128-142):
//   o16   = ReLU(BN1d(Linear(400 -> 16)(opponent trunk features)))            (osidedense)
//   x     = cat[own trunk features 400, own side 52, o16, opponent side 52, colour 1]   (521, padded to 528)
//   value = Sigmoid|Tanh(Linear(16 -> 1)(ReLU(BN1d(Linear(521 -> 16)(x)))))   (value_head; dropout = identity)
// The policy head Linear(521 -> 11583) stays a library GEMM on x (it is GEMM shaped: 50 GFLOP per
// 4096-leaf step).  One warp per leaf; lanes = (k parity, output): 32 lanes cover the 16 outputs of a
// layer twice, each half summing every other k, so the k-major weight rows are read conflict-free.
// BatchNorm1d is folded into the linear layers on the host (trunk.py).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "trl_common.cuh"

namespace {

constexpr int kFeat = 400, kSide = 52, kOpp = 16, kIn = 521, kPad = 528;
constexpr int kWarpsPerBlock = 16;   // one block per SM: the 59 KB of weights are staged once per block
// packed weights (fp32): Wo_t[400][16], bo[16], Wv_t[528][16], bv[16], w2[16], b2
constexpr int kOffWo = 0, kOffBo = kOffWo + kFeat * 16, kOffWv = kOffBo + 16, kOffBv = kOffWv + kPad * 16,
              kOffW2 = kOffBv + 16, kOffB2 = kOffW2 + 16, kWFloats = kOffB2 + 4;

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
alphasame_heads_kernel(const __nv_bfloat16* __restrict__ feats,    // [2G][400]: own grids then opponent grids (or a cache)
                       const int32_t* __restrict__ own_row,        // optional: feature rows of leaf g (< 0: skip)
                       const int32_t* __restrict__ opp_row,
                       const __nv_bfloat16* __restrict__ extras,   // [G][105]
                       int G, const float* __restrict__ weights, int use_tanh,
                       __nv_bfloat16* __restrict__ x_out,          // [G][528]
                       __nv_bfloat16* __restrict__ value_out) {    // [G]
    extern __shared__ __align__(16) float sm[];
    float* w = sm;
    float* xs_all = sm + kWFloats;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kWFloats; i += blockDim.x) w[i] = weights[i];
    __syncthreads();
    float* xs = xs_all + warp * (kPad + kFeat);
    float* fo = xs + kPad;
    const int o = lane & 15, h = lane >> 4;
    for (int g = blockIdx.x * kWarpsPerBlock + warp; g < G; g += gridDim.x * kWarpsPerBlock) {
        // ---- stage: own features, side inputs, opponent features ----
        const int ra = own_row ? own_row[g] : g, rb = own_row ? opp_row[g] : G + g;
        if (ra < 0) continue;   // warp-uniform
        const __nv_bfloat162* fa = reinterpret_cast<const __nv_bfloat162*>(feats + (size_t)ra * kFeat);
        const __nv_bfloat162* fb = reinterpret_cast<const __nv_bfloat162*>(feats + (size_t)rb * kFeat);
        for (int i = lane; i < kFeat / 2; i += 32) {
            const float2 a = __bfloat1622float2(fa[i]), b = __bfloat1622float2(fb[i]);
            xs[2 * i] = a.x; xs[2 * i + 1] = a.y;
            fo[2 * i] = b.x; fo[2 * i + 1] = b.y;
        }
        const __nv_bfloat16* ex = extras + (size_t)g * (2 * kSide + 1);
        for (int i = lane; i < 2 * kSide + 1; i += 32)
            xs[kFeat + i + (i >= kSide ? kOpp : 0)] = __bfloat162float(ex[i]);
        if (lane < kPad - kIn) xs[kIn + lane] = 0.f;
        __syncwarp();
        // ---- osidedense: 16 outputs x 400, lanes = (k parity, output) ----
        float acc = 0.f;
#pragma unroll 4
        for (int k = h; k < kFeat; k += 2) acc = fmaf(fo[k], w[kOffWo + k * 16 + o], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        const float o16 = fmaxf(acc + w[kOffBo + o], 0.f);
        if (h == 0) xs[kFeat + kSide + o] = __bfloat162float(__float2bfloat16(o16));   // rounded like the module's bf16 output
        __syncwarp();
        // ---- x out (bf16, 528 wide) ----
        __nv_bfloat162* xo = reinterpret_cast<__nv_bfloat162*>(x_out + (size_t)g * kPad);
        for (int i = lane; i < kPad / 2; i += 32) xo[i] = __floats2bfloat162_rn(xs[2 * i], xs[2 * i + 1]);
        // ---- value head ----
        acc = 0.f;
#pragma unroll 4
        for (int k = h; k < kPad; k += 2) acc = fmaf(xs[k], w[kOffWv + k * 16 + o], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        float v = fmaxf(acc + w[kOffBv + o], 0.f) * w[kOffW2 + o];
#pragma unroll
        for (int d = 8; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        v += w[kOffB2];
        if (lane == 0) value_out[g] = __float2bfloat16(use_tanh ? tanhf(v) : 1.f / (1.f + __expf(-v)));
        __syncwarp();
    }
}

}  // namespace

extern "C" int trl_alphasame_heads_weight_floats(void) { return kWFloats; }

static int launch_heads(const void* feats_bf16, const int32_t* own_row, const int32_t* opp_row, const void* extras_bf16,
                        int n_leaves, const float* weights, int use_tanh, void* x_out_bf16, void* value_out_bf16,
                        void* stream) {
    if (n_leaves < 0 || !feats_bf16 || !extras_bf16 || !weights || !x_out_bf16 || !value_out_bf16) return TRL_E_ARG;
    if (n_leaves == 0) return TRL_OK;
    const int smem = (kWFloats + kWarpsPerBlock * (kPad + kFeat)) * 4;
    static bool configured = false;
    if (!configured) {
        int rc = trl_check(cudaFuncSetAttribute(alphasame_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (rc) return rc;
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = (n_leaves + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > sms) blocks = sms;
    alphasame_heads_kernel<<<blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)feats_bf16, own_row, opp_row, (const __nv_bfloat16*)extras_bf16, n_leaves, weights, use_tanh,
        (__nv_bfloat16*)x_out_bf16, (__nv_bfloat16*)value_out_bf16);
    return trl_check(cudaGetLastError());
}

extern "C" int trl_alphasame_heads(const void* feats_bf16, const void* extras_bf16, int n_leaves, const float* weights,
                                   int use_tanh, void* x_out_bf16, void* value_out_bf16, void* stream) {
    return launch_heads(feats_bf16, nullptr, nullptr, extras_bf16, n_leaves, weights, use_tanh, x_out_bf16, value_out_bf16, stream);
}

extern "C" int trl_alphasame_heads_indexed(const void* cache_bf16, const int32_t* own_row, const int32_t* opp_row,
                                           const void* extras_bf16, int n_leaves, const float* weights, int use_tanh,
                                           void* x_out_bf16, void* value_out_bf16, void* stream) {
    if (!own_row || !opp_row) return TRL_E_ARG;
    return launch_heads(cache_bf16, own_row, opp_row, extras_bf16, n_leaves, weights, use_tanh, x_out_bf16, value_out_bf16, stream);
}
