// features.cu — network input encoding straight from packed bitboard states.
//
// Replaces ai.game_to_X / get_grids / get_pieces / get_stat / get_garbage / simplify_grid
// (reference ai.py:1364-1413): 11 features oriented to the side to move.  Output layout is
// what the batched net consumes:
//   grids  [2n][400]  rows 0..n-1 = side-to-move grid, rows n..2n-1 = opponent grid (0/1)
//   extras [n][105]   a_pieces(7x7 one-hot: active, held, 5 previews) a_b2b a_combo a_garbage
//                     o_pieces(49) o_b2b o_combo o_garbage color
// One warp per game; grid cells are produced from the uint16 bitrows with coalesced stores.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "features_dev.cuh"
#include "trl_common.cuh"
#include "trl_tables.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kCells = TRL_ROWS * TRL_COLS;  // 400
constexpr int kExtras = 105;

template <typename T> __device__ __forceinline__ T to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <typename T>
__global__ void __launch_bounds__(kWarps * 32)
encode_features_kernel(const TrlGame* __restrict__ games, const int32_t* __restrict__ index, int n,
                       T* __restrict__ grids, T* __restrict__ extras) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (i >= n) return;
    const int gi = index ? index[i] : i;
    if (gi < 0) return;  // nothing to encode for this item (its buffers keep their old content)
    const TrlGame& g = games[gi];
    const int turn = g.turn & 1;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const TrlPlayer& p = g.players[side == 0 ? turn : 1 - turn];
        T* out = grids + ((size_t)side * n + i) * kCells;
        for (int c = lane; c < kCells; c += 32) {
            const int row = c / TRL_COLS, col = c - row * TRL_COLS;
            out[c] = to_out<T>((float)((p.rows[row] >> col) & 1u));
        }
        T* ex = extras + (size_t)i * kExtras + side * 52;
        // 7x7 one-hot table (ai.py:1381-1392): slot 0 active, 1 held, 2..6 previews
        for (int c = lane; c < 49; c += 32) {
            const int slot = c / 7, mino = c - slot * 7;
            int piece = TRL_NONE;
            if (slot == 0) piece = p.piece;
            else if (slot == 1) piece = p.held;
            else if (slot - 2 < p.qlen) piece = p.queue[slot - 2];
            ex[c] = to_out<T>(piece == mino ? 1.f : 0.f);
        }
        if (lane == 0) {
            ex[49] = to_out<T>((float)p.b2b);
            ex[50] = to_out<T>((float)p.combo);
            ex[51] = to_out<T>((float)p.n_recv);
        }
    }
    if (lane == 0) extras[(size_t)i * kExtras + 104] = to_out<T>((float)turn);  // players[turn].color == turn
}

// Feature encoding for the trunk-feature cache (see include/trl.h): extras as above; the board
// cells only of the boards whose trunk features are not known yet, as a compact image list.
__global__ void __launch_bounds__(kWarps * 32)
encode_features_cached_kernel(const TrlGame* __restrict__ states, const int32_t* __restrict__ leaf_state,
                              const int32_t* __restrict__ leaf_parent, int n, __nv_bfloat16* __restrict__ cache,
                              __nv_bfloat16* __restrict__ images, int32_t* __restrict__ image_dest,
                              int32_t* __restrict__ n_images, __nv_bfloat16* __restrict__ extras,
                              int32_t* __restrict__ own_row, int32_t* __restrict__ opp_row, int32_t* __restrict__ row_of) {
    __shared__ int s_new[kWarps];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int i = blockIdx.x * kWarps + wib;
    const int si = (i < n) ? leaf_state[i] : -1;
    const int pi = (si >= 0) ? leaf_parent[i] : -1;
    const int n_new = (si < 0) ? 0 : ((pi < 0) ? 2 : 1);   // root: both boards; else only the mover's (= opponent of the side to move)
    // one atomic per BLOCK on the image counter (4096 same-address atomics would serialise in L2)
    if (lane == 0) s_new[wib] = n_new;
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) total += s_new[w];
        s_base = total ? atomicAdd(n_images, total) : 0;
    }
    __syncthreads();
    if (i >= n) return;
    if (si < 0) {
        if (lane == 0) { own_row[i] = -1; opp_row[i] = -1; }
        return;
    }
    int pos = s_base;
    for (int w = 0; w < wib; ++w) pos += s_new[w];
    TrlEncodeArgs E;
    E.cache = cache; E.images = images; E.image_dest = image_dest; E.n_images = n_images; E.extras = extras;
    E.own_row = own_row; E.opp_row = opp_row; E.row_of = row_of;
    trl_encode_cached_leaf(states[si], i, si, pi, pos, lane, E);
}

}  // namespace

extern "C" int trl_encode_features_cached(const TrlGame* states, const int32_t* leaf_state, const int32_t* leaf_parent,
                                          int n, void* cache_bf16, void* images_bf16, int32_t* image_dest,
                                          int32_t* n_images, void* extras_bf16, int32_t* own_row, int32_t* opp_row,
                                          int32_t* row_of, void* stream) {
    if (n < 0 || !states || !leaf_state || !leaf_parent || !cache_bf16 || !images_bf16 || !image_dest || !n_images ||
        !extras_bf16 || !own_row || !opp_row || !row_of)
        return TRL_E_ARG;
    if (n == 0) return TRL_OK;
    encode_features_cached_kernel<<<(n + kWarps - 1) / kWarps, kWarps * 32, 0, (cudaStream_t)stream>>>(
        states, leaf_state, leaf_parent, n, (__nv_bfloat16*)cache_bf16, (__nv_bfloat16*)images_bf16, image_dest,
        n_images, (__nv_bfloat16*)extras_bf16, own_row, opp_row, row_of);
    return trl_check(cudaGetLastError());
}

extern "C" int trl_encode_features(const TrlGame* games, const int32_t* index, int n, void* grids,
                                   void* extras, int dtype, void* stream) {
    if (n < 0 || !games || !grids || !extras || (dtype != 0 && dtype != 1)) return TRL_E_ARG;
    if (n == 0) return TRL_OK;
    const int blocks = (n + kWarps - 1) / kWarps;
    if (dtype == 0)
        encode_features_kernel<float><<<blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(
            games, index, n, (float*)grids, (float*)extras);
    else
        encode_features_kernel<__nv_bfloat16><<<blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(
            games, index, n, (__nv_bfloat16*)grids, (__nv_bfloat16*)extras);
    return trl_check(cudaGetLastError());
}
