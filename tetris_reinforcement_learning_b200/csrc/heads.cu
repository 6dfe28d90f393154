// heads.cu — everything between the convolutional trunk and the policy GEMM of AlphaSame, fused.
//
// Replaces, in eval mode, the tail of AlphaSame.forward (reference architectures.py:128-142):
//   o16   = ReLU(BN1d(Linear(400 -> 16)(opponent trunk features)))            (osidedense)
//   x     = cat[own trunk features 400, own side 52, o16, opponent side 52, colour 1]   (521, padded to 528)
//   value = Sigmoid|Tanh(Linear(16 -> 1)(ReLU(BN1d(Linear(521 -> 16)(x)))))   (value_head; dropout = identity)
// The policy head Linear(521 -> 11583) stays a library GEMM on x (it is GEMM shaped: 50 GFLOP per
// 4096-leaf step).  One warp per four leaves; lanes = (k mod 8, output quad): a lane accumulates 4 outputs x
// 4 leaves over every eighth k (two LDS.128 feed 16 independent FMAs), a butterfly sums the eight k lanes.
// BatchNorm1d is folded into the linear layers on the host (trunk.py).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "trl_common.cuh"

namespace {

constexpr int kFeat = 400, kSide = 52, kOpp = 16, kIn = 521, kPad = 528;
constexpr int kWarpsPerBlock = 8;    // x 4 leaves x 3.7 KB of staged inputs + 59 KB of weights per block
// packed weights (fp32): Wo_t[400][16], bo[16], Wv_t[528][16], bv[16], w2[16], b2
constexpr int kOffWo = 0, kOffBo = kOffWo + kFeat * 16, kOffWv = kOffBo + 16, kOffBv = kOffWv + kPad * 16,
              kOffW2 = kOffBv + 16, kOffB2 = kOffW2 + 16, kWFloats = kOffB2 + 4;

// kLeaves leaves per warp pass: the weight row of a k is loaded once and used for all of them
// (the kernel is bound by shared-memory loads: per k one weight word + one 16-byte vector of leaf inputs).
constexpr int kLeaves = 4;

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
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {   // weights -> shared memory: 16-byte loads, several in flight per thread (a scalar copy loop is
        // latency bound: 59 dependent L2 round trips per thread)
        const float4* src = reinterpret_cast<const float4*>(weights);
        float4* dst = reinterpret_cast<float4*>(w);
        constexpr int kVec = kWFloats / 4;
        static_assert(kWFloats % 4 == 0, "weights are copied as float4");
#pragma unroll 8
        for (int i = tid; i < kVec; i += kWarpsPerBlock * 32) dst[i] = src[i];
    }
    __syncthreads();
    // per warp: xs[k][leaf] (528 x 4) and fo[k][leaf] (400 x 4), leaf-interleaved so one LDS.128 feeds 4 FMAs
    float4* xs = reinterpret_cast<float4*>(sm + kWFloats) + warp * (kPad + kFeat);
    float4* fo = xs + kPad;
    float* xs_f = reinterpret_cast<float*>(xs);
    float* fo_f = reinterpret_cast<float*>(fo);
    const int og = lane & 3, kk = lane >> 2;   // output quad 4og..4og+3, k = kk mod 8
    const int n_quads = (G + kLeaves - 1) / kLeaves;
    for (int q = blockIdx.x * kWarpsPerBlock + warp; q < n_quads; q += gridDim.x * kWarpsPerBlock) {
        int ra[kLeaves], rb[kLeaves];
        bool any = false;
#pragma unroll
        for (int j = 0; j < kLeaves; ++j) {
            const int g = q * kLeaves + j;
            ra[j] = (g < G) ? (own_row ? own_row[g] : g) : -1;
            rb[j] = (g < G) ? (own_row ? opp_row[g] : G + g) : -1;
            any = any || ra[j] >= 0;
        }
        if (!any) continue;   // warp-uniform
        // ---- stage: own features, side inputs, opponent features (zeros for skipped leaves) ----
        // all global loads of the pass are issued before the first use (the pass is latency bound otherwise)
        constexpr int kFI = (kFeat / 2 + 31) / 32;            // 7 bf16x2 words per lane and feature row
        constexpr int kEI = (2 * kSide + 1 + 31) / 32;        // 4 side inputs per lane
        uint32_t ua[kLeaves][kFI], ub[kLeaves][kFI];
        uint16_t ue[kLeaves][kEI];
#pragma unroll
        for (int j = 0; j < kLeaves; ++j) {
            const uint32_t* fa = reinterpret_cast<const uint32_t*>(feats + (size_t)(ra[j] < 0 ? 0 : ra[j]) * kFeat);
            const uint32_t* fb = reinterpret_cast<const uint32_t*>(feats + (size_t)(ra[j] < 0 ? 0 : rb[j]) * kFeat);
            const uint16_t* ex = reinterpret_cast<const uint16_t*>(extras + (size_t)(ra[j] < 0 ? 0 : q * kLeaves + j) * (2 * kSide + 1));
#pragma unroll
            for (int t = 0; t < kFI; ++t) {
                const int i = lane + 32 * t;
                const bool ok = ra[j] >= 0 && i < kFeat / 2;
                ua[j][t] = ok ? fa[i] : 0u;
                ub[j][t] = ok ? fb[i] : 0u;
            }
#pragma unroll
            for (int t = 0; t < kEI; ++t) {
                const int i = lane + 32 * t;
                ue[j][t] = (ra[j] >= 0 && i < 2 * kSide + 1) ? ex[i] : (uint16_t)0;
            }
        }
#pragma unroll
        for (int j = 0; j < kLeaves; ++j) {
#pragma unroll
            for (int t = 0; t < kFI; ++t) {
                const int i = lane + 32 * t;
                if (i < kFeat / 2) {
                    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ua[j][t]));
                    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ub[j][t]));
                    xs_f[(2 * i) * kLeaves + j] = a.x; xs_f[(2 * i + 1) * kLeaves + j] = a.y;
                    fo_f[(2 * i) * kLeaves + j] = b.x; fo_f[(2 * i + 1) * kLeaves + j] = b.y;
                }
            }
#pragma unroll
            for (int t = 0; t < kEI; ++t) {
                const int i = lane + 32 * t;
                if (i < 2 * kSide + 1)
                    xs_f[(kFeat + i + (i >= kSide ? kOpp : 0)) * kLeaves + j] = __bfloat162float(__ushort_as_bfloat16(ue[j][t]));
            }
            if (lane < kOpp) xs_f[(kFeat + kSide + lane) * kLeaves + j] = 0.f;
            if (lane < kPad - kIn) xs_f[(kIn + lane) * kLeaves + j] = 0.f;
        }
        __syncwarp();
        // ---- osidedense: 16 outputs x 400 for 4 leaves; lanes = (k mod 8, output quad) ----
        // per k one LDS.128 of weights (4 outputs) and one LDS.128 of inputs (4 leaves) feed 16 independent FMAs;
        // 50 / 66 iterations per layer instead of 200 / 264, then a 3-step butterfly over the k lanes
        float acc[4][kLeaves];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < kLeaves; ++j) acc[a][j] = 0.f;
#pragma unroll 2
        for (int k = kk; k < kFeat; k += 8) {
            const float4 wk = *reinterpret_cast<const float4*>(&w[kOffWo + k * 16 + 4 * og]);
            const float4 f = fo[k];
            const float wv4[4] = {wk.x, wk.y, wk.z, wk.w}, fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < kLeaves; ++j) acc[a][j] = fmaf(fv[j], wv4[a], acc[a][j]);
        }
#pragma unroll
        for (int d = 4; d <= 16; d <<= 1)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < kLeaves; ++j) acc[a][j] += __shfl_xor_sync(0xffffffffu, acc[a][j], d);
        if (kk == 0) {   // rounded like the module's bf16 output
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float bo = w[kOffBo + 4 * og + a];
                xs[kFeat + kSide + 4 * og + a] = make_float4(__bfloat162float(__float2bfloat16(fmaxf(acc[a][0] + bo, 0.f))),
                                                             __bfloat162float(__float2bfloat16(fmaxf(acc[a][1] + bo, 0.f))),
                                                             __bfloat162float(__float2bfloat16(fmaxf(acc[a][2] + bo, 0.f))),
                                                             __bfloat162float(__float2bfloat16(fmaxf(acc[a][3] + bo, 0.f))));
            }
        }
        __syncwarp();
        // ---- x out (bf16, 528 wide) ----
#pragma unroll
        for (int j = 0; j < kLeaves; ++j) {
            if (ra[j] < 0) continue;
            __nv_bfloat162* xo = reinterpret_cast<__nv_bfloat162*>(x_out + (size_t)(q * kLeaves + j) * kPad);
            for (int i = lane; i < kPad / 2; i += 32)
                xo[i] = __floats2bfloat162_rn(xs_f[(2 * i) * kLeaves + j], xs_f[(2 * i + 1) * kLeaves + j]);
        }
        // ---- value head ----
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < kLeaves; ++j) acc[a][j] = 0.f;
#pragma unroll 2
        for (int k = kk; k < kPad; k += 8) {
            const float4 wk = *reinterpret_cast<const float4*>(&w[kOffWv + k * 16 + 4 * og]);
            const float4 f = xs[k];
            const float wv4[4] = {wk.x, wk.y, wk.z, wk.w}, fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < kLeaves; ++j) acc[a][j] = fmaf(fv[j], wv4[a], acc[a][j]);
        }
#pragma unroll
        for (int d = 4; d <= 16; d <<= 1)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < kLeaves; ++j) acc[a][j] += __shfl_xor_sync(0xffffffffu, acc[a][j], d);
        float v[kLeaves];
#pragma unroll
        for (int j = 0; j < kLeaves; ++j) {
            v[j] = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) v[j] = fmaf(fmaxf(acc[a][j] + w[kOffBv + 4 * og + a], 0.f), w[kOffW2 + 4 * og + a], v[j]);
            v[j] += __shfl_xor_sync(0xffffffffu, v[j], 1);
            v[j] += __shfl_xor_sync(0xffffffffu, v[j], 2);
            v[j] += w[kOffB2];
            if (lane == 0 && ra[j] >= 0)
                value_out[q * kLeaves + j] = __float2bfloat16(use_tanh ? tanhf(v[j]) : 1.f / (1.f + __expf(-v[j])));
        }
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
    const int smem = (kWFloats + kWarpsPerBlock * kLeaves * (kPad + kFeat)) * 4;
    static bool configured = false;
    if (!configured) {
        int rc = trl_check(cudaFuncSetAttribute(alphasame_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (rc) return rc;
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = ((n_leaves + kLeaves - 1) / kLeaves + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > sms) blocks = sms;
    return trl_launch_ex(alphasame_heads_kernel, dim3(blocks), dim3(kWarpsPerBlock * 32), (size_t)smem, (cudaStream_t)stream, false, false,
                         (const __nv_bfloat16*)feats_bf16, own_row, opp_row, (const __nv_bfloat16*)extras_bf16, n_leaves, weights,
                         use_tanh, (__nv_bfloat16*)x_out_bf16, (__nv_bfloat16*)value_out_bf16);
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
