// trl_tables.cuh — piece shape / kick / policy-plane tables and small board helpers shared
// by the sm_100a kernels.  Data restates reference const.py:72-136 (policy planes),
// :191-235 (SRS+ kick tables incl. 180), :238-281 (mino coordinates).
#pragma once
#include <stdint.h>
#include "../../include/trl.h"

#define TRL_MAP_W 14  // COLS + 4: bit mx of a validity row <=> origin x = mx - 2
#define TRL_MAP_H 44  // ROWS + 4: validity row my <=> origin y = my - 2
#define TRL_FULL_ROW 0x3FFu

enum : int { P_Z = 0, P_L = 1, P_O = 2, P_S = 3, P_I = 4, P_J = 5, P_T = 6 };

// Mino cells packed per (piece, rotation): nibble pairs (col, row) for 4 minos, little end
// first: bits [8m, 8m+4) = col offset, [8m+4, 8m+8) = row offset.
#define TRL_PK(c0, r0, c1, r1, c2, r2, c3, r3)                                          \
    ((uint32_t)(c0) | ((uint32_t)(r0) << 4) | ((uint32_t)(c1) << 8) | ((uint32_t)(r1) << 12) | \
     ((uint32_t)(c2) << 16) | ((uint32_t)(r2) << 20) | ((uint32_t)(c3) << 24) | ((uint32_t)(r3) << 28))

__device__ __constant__ uint32_t c_minos[7][4] = {
    /* Z */ {TRL_PK(0, 0, 1, 0, 1, 1, 2, 1), TRL_PK(1, 1, 1, 2, 2, 0, 2, 1),
             TRL_PK(0, 1, 1, 1, 1, 2, 2, 2), TRL_PK(0, 1, 0, 2, 1, 0, 1, 1)},
    /* L */ {TRL_PK(0, 1, 1, 1, 2, 0, 2, 1), TRL_PK(1, 0, 1, 1, 1, 2, 2, 2),
             TRL_PK(0, 1, 0, 2, 1, 1, 2, 1), TRL_PK(0, 0, 1, 0, 1, 1, 1, 2)},
    /* O */ {TRL_PK(0, 0, 0, 1, 1, 0, 1, 1), TRL_PK(0, 0, 0, 1, 1, 0, 1, 1),
             TRL_PK(0, 0, 0, 1, 1, 0, 1, 1), TRL_PK(0, 0, 0, 1, 1, 0, 1, 1)},
    /* S */ {TRL_PK(0, 1, 1, 0, 1, 1, 2, 0), TRL_PK(1, 0, 1, 1, 2, 1, 2, 2),
             TRL_PK(0, 2, 1, 1, 1, 2, 2, 1), TRL_PK(0, 0, 0, 1, 1, 1, 1, 2)},
    /* I */ {TRL_PK(0, 1, 1, 1, 2, 1, 3, 1), TRL_PK(2, 0, 2, 1, 2, 2, 2, 3),
             TRL_PK(0, 2, 1, 2, 2, 2, 3, 2), TRL_PK(1, 0, 1, 1, 1, 2, 1, 3)},
    /* J */ {TRL_PK(0, 0, 0, 1, 1, 1, 2, 1), TRL_PK(1, 0, 1, 1, 1, 2, 2, 0),
             TRL_PK(0, 1, 1, 1, 2, 1, 2, 2), TRL_PK(0, 2, 1, 0, 1, 1, 1, 2)},
    /* T */ {TRL_PK(0, 1, 1, 0, 1, 1, 2, 1), TRL_PK(1, 0, 1, 1, 1, 2, 2, 1),
             TRL_PK(0, 1, 1, 1, 1, 2, 2, 1), TRL_PK(0, 1, 1, 0, 1, 1, 1, 2)},
};

// Kick lists: c_kicks[isI][from_rot][kick_dir-1] ; n kicks then (kx, ky) pairs.
// new_rot = (from_rot + kick_dir) & 3; target = (x + kx, y - ky)  (player.py:88-92).
struct TrlKicks {
    int8_t n;
    int8_t k[6][2];
};

__device__ __constant__ TrlKicks c_kicks[2][4][3] = {
    {   // wallkicks (all pieces but I), const.py:191-212
        {{5, {{0, 0}, {-1, 0}, {-1, 1}, {0, -2}, {-1, -2}}},            // 0 -> 1
         {6, {{0, 0}, {0, 1}, {1, 1}, {-1, 1}, {1, 0}, {-1, 0}}},       // 0 -> 2
         {5, {{0, 0}, {1, 0}, {1, 1}, {0, -2}, {1, -2}}}},              // 0 -> 3
        {{5, {{0, 0}, {1, 0}, {1, -1}, {0, 2}, {1, 2}}},                // 1 -> 2
         {6, {{0, 0}, {1, 0}, {1, 2}, {1, 1}, {0, 2}, {0, 1}}},         // 1 -> 3
         {5, {{0, 0}, {1, 0}, {1, -1}, {0, 2}, {1, 2}}}},               // 1 -> 0
        {{5, {{0, 0}, {1, 0}, {1, 1}, {0, -2}, {1, -2}}},               // 2 -> 3
         {6, {{0, 0}, {0, -1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}}},    // 2 -> 0
         {5, {{0, 0}, {-1, 0}, {-1, 1}, {0, -2}, {-1, -2}}}},           // 2 -> 1
        {{5, {{0, 0}, {-1, 0}, {-1, -1}, {0, 2}, {-1, 2}}},             // 3 -> 0
         {6, {{0, 0}, {-1, 0}, {-1, 2}, {-1, 1}, {0, 2}, {0, 1}}},      // 3 -> 1
         {5, {{0, 0}, {-1, 0}, {-1, -1}, {0, 2}, {-1, 2}}}},            // 3 -> 2
    },
    {   // i_wallkicks, const.py:214-235
        {{5, {{0, 0}, {-2, 0}, {1, 0}, {-2, -1}, {1, 2}}},              // 0 -> 1
         {2, {{0, 0}, {0, 1}}},                                         // 0 -> 2
         {5, {{0, 0}, {-1, 0}, {2, 0}, {-1, 2}, {2, -1}}}},             // 0 -> 3
        {{5, {{0, 0}, {-1, 0}, {2, 0}, {-1, 2}, {2, -1}}},              // 1 -> 2
         {2, {{0, 0}, {1, 0}}},                                         // 1 -> 3
         {5, {{0, 0}, {2, 0}, {-1, 0}, {2, 1}, {-1, -2}}}},             // 1 -> 0
        {{5, {{0, 0}, {2, 0}, {-1, 0}, {2, 1}, {-1, -2}}},              // 2 -> 3
         {2, {{0, 0}, {0, -1}}},                                        // 2 -> 0
         {5, {{0, 0}, {1, 0}, {-2, 0}, {1, -2}, {-2, 1}}}},             // 2 -> 1
        {{5, {{0, 0}, {1, 0}, {-2, 0}, {1, -2}, {-2, 1}}},              // 3 -> 0
         {2, {{0, 0}, {-1, 0}}},                                        // 3 -> 1
         {5, {{0, 0}, {-2, 0}, {1, 0}, {-2, -1}, {1, 2}}}},             // 3 -> 2
    },
};

// policy planes (const.py:82-118): first plane, number of rotation planes, matrix size
__device__ __constant__ uint8_t c_plane_base[7] = {1, 7, 0, 3, 5, 11, 15};
__device__ __constant__ uint8_t c_plane_nrot[7] = {2, 4, 1, 2, 2, 4, 4};
__device__ __constant__ uint8_t c_matrix_size[7] = {3, 3, 2, 3, 4, 3, 3};
// inverse: plane -> piece type (rotation = plane - base for planes < 15)
__device__ __constant__ uint8_t c_plane_piece[27] = {2, 0, 0, 3, 3, 4, 4, 1, 1, 1, 1, 5, 5, 5, 5,
                                                     6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6};

__device__ __forceinline__ int trl_spawn_x(int type) { return type == P_O ? 4 : 3; }

// Empty-cell mask (10 bits) of board row `row`, 0 outside the board (board.py:23-30).
__device__ __forceinline__ uint32_t trl_empty_row(const uint16_t* rows, int row) {
    return ((unsigned)row < (unsigned)TRL_ROWS) ? (~(uint32_t)rows[row] & TRL_FULL_ROW) : 0u;
}

// Piece.get_mino_coords + Board.is_valid_position (piece.py:50-55, board.py:23-30).
__device__ __forceinline__ bool trl_fits(const uint16_t* rows, uint32_t minos, int x, int y) {
    bool ok = true;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        int c = x + (int)((minos >> (8 * m)) & 15u);
        int r = y + (int)((minos >> (8 * m + 4)) & 15u);
        ok = ok && ((unsigned)c < (unsigned)TRL_COLS) && ((unsigned)r < (unsigned)TRL_ROWS) &&
             (((rows[(unsigned)r < (unsigned)TRL_ROWS ? r : 0] >> (c & 15)) & 1u) == 0u);
    }
    return ok;
}

// Validity row: bit mx set <=> rotation `minos` with origin (mx-2, my-2) is in bounds and
// collides with nothing (move_generation.py:490-528).
__device__ __forceinline__ uint32_t trl_valid_row(const uint16_t* rows, uint32_t minos, int my) {
    uint32_t v = 0x3FFFu;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        int co = (int)((minos >> (8 * m)) & 15u);
        int ro = (int)((minos >> (8 * m + 4)) & 15u);
        v &= (trl_empty_row(rows, my - 2 + ro) << 2) >> co;
    }
    return v;
}

// Philox4x32-10 (counter RNG that replaces the reference's `random` draws).
__device__ __forceinline__ void trl_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                           uint32_t c3, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
