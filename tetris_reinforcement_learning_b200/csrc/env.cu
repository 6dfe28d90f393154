// env.cu — batched env-step and game-setup kernels (one warp per game, the 400-byte game
// staged in shared memory as uint16 bitrows).  Rules: env_step.cuh.
#include <cuda_runtime.h>
#include <stdint.h>

#include "env_step.cuh"
#include "trl_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kGameWords = sizeof(TrlGame) / 4;  // 100

static_assert(sizeof(TrlPlayer) == 192, "TrlPlayer layout");
static_assert(sizeof(TrlGame) == 400, "TrlGame layout");
static_assert(sizeof(TrlStepOut) == 8, "TrlStepOut layout");

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
env_step_kernel(TrlGame* __restrict__ games, const uint16_t* __restrict__ moves, int n,
                TrlStepOut* __restrict__ out, int add_bag, uint64_t seed) {
    __shared__ __align__(16) TrlGame s_games[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = blockIdx.x * kWarpsPerBlock + wib;
    if (i >= n) return;
    const uint32_t mv = moves[i];
    if (mv == 0xFFFFu) {  // skipped item
        if (out && lane == 0) { TrlStepOut z = {0, 0, 0, 0, 0}; out[i] = z; }
        return;
    }
    uint32_t* sg = reinterpret_cast<uint32_t*>(&s_games[wib]);
    uint32_t* gg = reinterpret_cast<uint32_t*>(games + i);
    for (int w = lane; w < kGameWords; w += 32) sg[w] = gg[w];
    __syncwarp();
    if (lane == 0) {
        TrlStepOut o = trl_env_step_scalar(&s_games[wib], (int)mv, add_bag != 0, seed, 0u, &s_games[wib].rng_ctr);
        if (out) out[i] = o;
    }
    __syncwarp();
    for (int w = lane; w < kGameWords; w += 32) gg[w] = sg[w];
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
game_setup_kernel(TrlGame* __restrict__ games, int n, uint32_t first_game_id, uint32_t id_stride, uint64_t seed) {
    __shared__ __align__(16) TrlGame s_games[kWarpsPerBlock];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = blockIdx.x * kWarpsPerBlock + wib;
    if (i >= n) return;
    if (lane == 0) trl_game_setup_scalar(&s_games[wib], first_game_id + (uint32_t)i * id_stride, seed);
    __syncwarp();
    uint32_t* sg = reinterpret_cast<uint32_t*>(&s_games[wib]);
    uint32_t* gg = reinterpret_cast<uint32_t*>(games + i);
    for (int w = lane; w < kGameWords; w += 32) gg[w] = sg[w];
}

}  // namespace

extern "C" int trl_env_step(TrlGame* games, const uint16_t* moves, int n, TrlStepOut* out, int add_bag,
                            uint64_t seed, void* stream) {
    if (n < 0 || !games || !moves) return TRL_E_ARG;
    if (n == 0) return TRL_OK;
    env_step_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        games, moves, n, out, add_bag, seed);
    return trl_check(cudaGetLastError());
}

extern "C" int trl_game_setup(TrlGame* games, int n, uint32_t first_game_id, uint32_t id_stride, uint64_t seed,
                              void* stream) {
    if (n < 0 || !games) return TRL_E_ARG;
    if (n == 0) return TRL_OK;
    game_setup_kernel<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        games, n, first_game_id, id_stride, seed);
    return trl_check(cudaGetLastError());
}

extern "C" int trl_env_step_host(TrlGame* games, const uint16_t* moves, int n, TrlStepOut* out,
                                 int add_bag, uint64_t seed) {
    if (n < 0 || !games || !moves) return TRL_E_ARG;
    if (n == 0) return TRL_OK;
    cudaStream_t s = trl_host_stream();
    if (!s) return TRL_E_CUDA;
    size_t bytes = (size_t)n * (sizeof(TrlGame) + sizeof(TrlStepOut) + 2) + 64;
    char* ws = (char*)trl_workspace(TRL_WS_HOST_STAGE, bytes);
    if (!ws) return TRL_E_NOMEM;
    TrlGame* d_games = (TrlGame*)ws;
    TrlStepOut* d_out = (TrlStepOut*)(ws + (size_t)n * sizeof(TrlGame));
    uint16_t* d_moves = (uint16_t*)(ws + (size_t)n * (sizeof(TrlGame) + sizeof(TrlStepOut)));
    int rc = trl_check(cudaMemcpyAsync(d_games, games, (size_t)n * sizeof(TrlGame), cudaMemcpyHostToDevice, s));
    if (!rc) rc = trl_check(cudaMemcpyAsync(d_moves, moves, (size_t)n * 2, cudaMemcpyHostToDevice, s));
    if (!rc) rc = trl_env_step(d_games, d_moves, n, d_out, add_bag, seed, s);
    if (!rc) rc = trl_check(cudaMemcpyAsync(games, d_games, (size_t)n * sizeof(TrlGame), cudaMemcpyDeviceToHost, s));
    if (!rc && out) rc = trl_check(cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(TrlStepOut), cudaMemcpyDeviceToHost, s));
    if (!rc) rc = trl_check(cudaStreamSynchronize(s));
    return rc;
}

extern "C" int trl_game_setup_host(TrlGame* games, int n, uint32_t first_game_id, uint32_t id_stride, uint64_t seed) {
    if (n < 0 || !games) return TRL_E_ARG;
    if (n == 0) return TRL_OK;
    cudaStream_t s = trl_host_stream();
    if (!s) return TRL_E_CUDA;
    TrlGame* d_games = (TrlGame*)trl_workspace(TRL_WS_HOST_STAGE, (size_t)n * sizeof(TrlGame));
    if (!d_games) return TRL_E_NOMEM;
    int rc = trl_game_setup(d_games, n, first_game_id, id_stride, seed, s);
    if (!rc) rc = trl_check(cudaMemcpyAsync(games, d_games, (size_t)n * sizeof(TrlGame), cudaMemcpyDeviceToHost, s));
    if (!rc) rc = trl_check(cudaStreamSynchronize(s));
    return rc;
}
