// features_dev.cuh — per-leaf body of the cached feature encoder (features.cu), shared with the search
// kernel that runs it right after selecting a leaf (mcts.cu: expand + select + encode in one launch).
//
// Replaces ai.game_to_X for one leaf (reference ai.py:1364-1413) in the trunk-feature-cache form of
// include/trl.h (trl_encode_features_cached): extras, the 0/1 cells of the boards whose trunk features
// are unknown, the cache row of the side to move (inherited from the parent state), own_row / opp_row.  One warp per leaf.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "trl_tables.cuh"

struct TrlEncodeArgs {
    __nv_bfloat16* cache;        // [n_states * 2][400] trunk features per (state, player)
    __nv_bfloat16* images;       // [<= 2n][400] compact list of boards for the trunk
    int32_t* image_dest;         // [<= 2n] cache row each image's features go to
    int32_t* n_images;           // [1] running count (zero on entry of a step; the trunk kernel resets it)
    __nv_bfloat16* extras;       // [n][105]
    int32_t* own_row;            // [n]
    int32_t* opp_row;            // [n]
    int32_t* row_of;             // [n_states * 2] cache row that holds the features of (state, player)
};

// g: the leaf state (any address space), i: leaf index, si / pi: state index of the leaf / its parent
// (pi < 0: root, both boards are new), pos: first slot in `images` reserved for this leaf.
// inherit_row: row_of[parent][side to move] if the caller has loaded it already, else -1 (loaded here).
__device__ __forceinline__ void trl_encode_cached_leaf(const TrlGame& g, int i, int si, int pi, int pos, int lane,
                                                       const TrlEncodeArgs& E, int inherit_row = -1) {
    constexpr int kCells = TRL_ROWS * TRL_COLS, kExtras = 105;
    const int turn = g.turn & 1;
    int own_r = si * 2 + turn;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int pl = side == 0 ? turn : 1 - turn;
        const TrlPlayer& p = g.players[pl];
        const int row = si * 2 + pl;
        if (side == 0 && pi >= 0) {
            // the side to move did not move: its board is the parent's, so are its trunk features: the leaf
            // points at the row that holds them (no 800-byte copy)
            own_r = inherit_row >= 0 ? inherit_row : E.row_of[pi * 2 + pl];
            if (lane == 0) E.row_of[row] = own_r;
        } else {
            if (lane == 0) E.row_of[row] = row;
            const int k = pos + ((side == 1 && pi < 0) ? 1 : 0);
            // lane = board row: 10 cells = five words of two bf16 (1.0 = 0x3F80)
            uint32_t* out = reinterpret_cast<uint32_t*>(E.images + (size_t)k * kCells);
            for (int r = lane; r < TRL_ROWS; r += 32) {
                const uint32_t bits = p.rows[r];
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    out[r * 5 + q] = (((bits >> (2 * q)) & 1u) ? 0x3F80u : 0u) | (((bits >> (2 * q + 1)) & 1u) ? 0x3F800000u : 0u);
            }
            if (lane == 0) E.image_dest[k] = row;
        }
        __nv_bfloat16* ex = E.extras + (size_t)i * kExtras + side * 52;
        // 7x7 one-hot table (ai.py:1381-1392): slot 0 active, 1 held, 2..6 previews
        for (int c = lane; c < 49; c += 32) {
            const int slot = c / 7, mino = c - slot * 7;
            int piece = TRL_NONE;
            if (slot == 0) piece = p.piece;
            else if (slot == 1) piece = p.held;
            else if (slot - 2 < p.qlen) piece = p.queue[slot - 2];
            ex[c] = __float2bfloat16(piece == mino ? 1.f : 0.f);
        }
        if (lane == 0) {
            ex[49] = __float2bfloat16((float)p.b2b);
            ex[50] = __float2bfloat16((float)p.combo);
            ex[51] = __float2bfloat16((float)p.n_recv);
        }
    }
    if (lane == 0) {
        E.extras[(size_t)i * kExtras + 104] = __float2bfloat16((float)turn);   // players[turn].color == turn
        E.own_row[i] = own_r;
        E.opp_row[i] = si * 2 + (1 - turn);
    }
}
