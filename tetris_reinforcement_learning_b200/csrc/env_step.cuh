// env_step.cuh — device-side rules of one placement (ruleset s2), shared by the env-step
// kernel and the search kernel's leaf materialisation.
//
// Replaces Game.make_move(move, add_bag, add_history=False) (reference game.py:40-118),
// Player.place_piece / hold_piece / create_next_piece / spawn_garbage (player.py:29-44,
// 109-205), Board.create_garbage (board.py:12-21), Stats.get_attack + update_b2b_level
// for ruleset 's2' (stats.py:30-43, 88-129) and Queue.generate_bag (piece_queue.py:17-20).
//
// The game lives in SHARED memory (uint16 bitrows); the functions below are scalar and are
// run by one lane of the warp that owns the game, the other lanes do the coalesced
// global<->shared staging of the 400-byte state.
#pragma once
#include "trl_tables.cuh"

// random.randint(0, 9) of player.py:185 as draw #ctr of Philox stream (seed, game_id, stream).
// stream 0 = real moves (ctr = TrlGame.rng_ctr); stream 1+k = leaf simulations inside the k-th
// search of the game (ctr = sequential draw index within that search).
__device__ __forceinline__ int trl_garbage_column(uint64_t seed, uint32_t game_id, uint32_t stream, uint32_t ctr) {
    uint32_t o[4];
    trl_philox(seed, ctr, game_id, 1u, stream, o);
    return (int)__umulhi(o[0], 10u);
}

// Queue.generate_bag: Fisher-Yates `for i in 6..1: j = randbelow(i+1); swap` over "ZLOSIJT".
__device__ __forceinline__ void trl_generate_bag(uint64_t seed, uint32_t game_id, uint32_t bag_ctr,
                                                 int player, uint8_t bag[7]) {
    uint32_t r[8];
    trl_philox(seed, bag_ctr, game_id, 2u, (uint32_t)player, r);
    trl_philox(seed, bag_ctr, game_id, 2u, (uint32_t)player | 0x100u, r + 4);
#pragma unroll
    for (int i = 0; i < 7; ++i) bag[i] = (uint8_t)i;
    int k = 0;
#pragma unroll
    for (int i = 6; i >= 1; --i) {
        int j = (int)__umulhi(r[k++], (uint32_t)(i + 1));
        uint8_t t = bag[i]; bag[i] = bag[j]; bag[j] = t;
    }
}

__device__ __forceinline__ bool trl_cell_free(const uint16_t* rows, int c, int r) {
    return (unsigned)c < (unsigned)TRL_COLS && (unsigned)r < (unsigned)TRL_ROWS && !((rows[r] >> c) & 1u);
}

// Player.create_piece (player.py:37-44)
__device__ __forceinline__ void trl_create_piece(TrlPlayer* p, int type) {
    if (trl_fits(p->rows, c_minos[type][0], trl_spawn_x(type), TRL_SPAWN_Y)) p->piece = (uint8_t)type;
    else p->game_over = 1;
}

// Player.create_next_piece (player.py:29-35)
__device__ __forceinline__ void trl_create_next_piece(TrlPlayer* p) {
    if (p->qlen > 0) {
        int next = p->queue[0];
        for (int i = 0; i + 1 < p->qlen; ++i) p->queue[i] = p->queue[i + 1];
        p->qlen--;
        trl_create_piece(p, next);
    } else if (p->held != TRL_NONE) {
        int next = p->held;
        p->held = TRL_NONE;
        trl_create_piece(p, next);
    }
}

// Player.hold_piece (player.py:190-201); false where the reference would raise.
__device__ __forceinline__ bool trl_hold_piece(TrlPlayer* p) {
    if (p->held == TRL_NONE) {
        if (p->piece == TRL_NONE) return false;
        p->held = p->piece;
        p->piece = TRL_NONE;
        trl_create_next_piece(p);
    } else {
        int tmp = p->held;
        p->held = p->piece;  // TRL_NONE stays "none"
        trl_create_piece(p, tmp);
    }
    return true;
}

// Game.add_bag_to_all (game.py:34-38)
__device__ __forceinline__ void trl_add_bag_to_all(TrlGame* g, uint64_t seed) {
    for (int pl = 0; pl < 2; ++pl) {
        uint8_t bag[7];
        trl_generate_bag(seed, g->game_id, g->bag_ctr, pl, bag);
        TrlPlayer* p = &g->players[pl];
        for (int i = 0; i < 7 && p->qlen < TRL_QUEUE_CAP; ++i) p->queue[p->qlen++] = bag[i];
    }
    g->bag_ctr++;
}

// Stats.get_attack for ruleset s2 in integer arithmetic (stats.py:88-129): every floor()
// argument there is a dyadic rational with denominator <= 16.
__device__ __forceinline__ int trl_get_attack_s2(int n, bool tspin, bool mini, bool all_clear,
                                                 int& combo, int& b2b, int& level) {
    if (n == 0) { combo = 0; return 0; }
    int attack = 0;
    const bool is_b2b = tspin || mini || n == 4;
    if (is_b2b) b2b += 1;
    else if (all_clear) { attack += 5; b2b += 1; }
    else { if (b2b >= 4) attack += b2b; b2b = -1; }  // surge
    // update_b2b_level (stats.py:30-43): the level never decreases
    const int thr[9] = {-1, 1, 3, 8, 24, 67, 185, 504, 1370};
#pragma unroll
    for (int i = 0; i < 9; ++i)
        if (b2b >= thr[i] && level < i) level = i;
    const int L = min(max(level, 0), 1);
    const int q = 4 + combo;
    if (n == 1) {
        if (!tspin) attack += (2 + combo) >> 2;
        else if (mini) attack += (b2b <= 0 && level <= 0) ? ((2 + combo) >> 2) : ((L * q) >> 2);
        else attack += (8 + L * q) >> 2;
    } else {
        int inner4 = (tspin ? 2 * n * (mini ? 1 : 4) : 4 * (1 << (n - 2))) + 4 * L * (is_b2b ? 1 : 0);
        attack += (q * inner4) >> 4;
    }
    combo += 1;
    return attack;
}

// Stats.get_attack for ruleset s1 (stats.py:49-86), same integer-arithmetic treatment: B2B only for
// quads and T-spins, no surge, the full b2b level enters the formulas, +10 for any all-clear.
__device__ __forceinline__ int trl_get_attack_s1(int n, bool tspin, bool mini, bool all_clear,
                                                 int& combo, int& b2b, int& level) {
    if (n == 0) { combo = 0; return 0; }
    int attack = 0;
    const bool is_b2b = tspin || n == 4;
    b2b = is_b2b ? b2b + 1 : -1;
    const int thr[9] = {-1, 1, 3, 8, 24, 67, 185, 504, 1370};
#pragma unroll
    for (int i = 0; i < 9; ++i)
        if (b2b >= thr[i] && level < i) level = i;
    const int q = 4 + combo;
    if (n == 1) {
        if (!tspin) attack += (2 + combo) >> 2;
        else if (mini) attack += (b2b <= 0 && level <= 0) ? ((2 + combo) >> 2) : ((level * q) >> 2);
        else attack += ((2 + level) * q) >> 2;
    } else {
        const int inner4 = (tspin ? 2 * n * (mini ? 1 : 4) : 4 * (1 << (n - 2))) + 4 * level * (is_b2b ? 1 : 0);
        attack += (q * inner4) >> 4;
    }
    combo += 1;
    if (all_clear) attack += 10;
    return attack;
}

// One full Game.make_move on a game in shared memory.  Scalar: call from ONE lane.
static __device__ __noinline__ TrlStepOut trl_env_step_scalar(TrlGame* g, int move, bool add_bag, uint64_t seed,
                                                       uint32_t stream, uint32_t* ctr) {
    TrlStepOut o;
    o.rows_cleared = 0; o.attack = 0; o.flags = 0; o.garbage_col = 0; o.status = 0;
    if ((unsigned)move >= (unsigned)TRL_POLICY_SIZE) { o.status = TRL_ST_BAD_MOVE; return o; }
    const int turn = g->turn & 1;
    TrlPlayer* p = &g->players[turn];
    TrlPlayer* opp = &g->players[1 - turn];

    // ---- Game.move_piece (game.py:40-64) ----
    const int plane = move / (TRL_POLICY_ROWS * TRL_POLICY_COLS);
    const int rem = move - plane * (TRL_POLICY_ROWS * TRL_POLICY_COLS);
    const int y = rem / TRL_POLICY_COLS;
    const int x = rem - y * TRL_POLICY_COLS - 2;
    const int type = c_plane_piece[plane];
    int rot, tsi = 0;
    if (plane >= 15) { rot = (plane - 15) & 3; tsi = (plane - 15) >> 2; }
    else rot = plane - c_plane_base[type];
    if (p->piece == TRL_NONE || p->piece != type) {
        o.flags |= 0x8;
        if (!trl_hold_piece(p)) { o.status = TRL_ST_BAD_MOVE; return o; }
    }
    if (p->piece != type) { o.status = TRL_ST_BAD_MOVE; return o; }
    bool roc = tsi >= 1, ulk = tsi == 2;
    if (p->game_over) return o;  // game.py:70 — nothing happens, not even the turn flip

    // ---- Player.place_piece (player.py:109-188) ----
    const uint32_t minos = c_minos[type][rot];
    int py = y;  // ghost_y (player.py:49-59)
    while (trl_fits(p->rows, minos, x, py + 1)) ++py;
    if (py != y) { roc = false; ulk = false; }

    bool tspin = false, mini = false, all_clear = false;
    if (type == P_T && roc) {  // 3-corner rule + 2-front-corner mini rule (player.py:122-140)
        const bool f0 = !trl_cell_free(p->rows, x, py), f1 = !trl_cell_free(p->rows, x + 2, py);
        const bool f2 = !trl_cell_free(p->rows, x + 2, py + 2), f3 = !trl_cell_free(p->rows, x, py + 2);
        tspin = ((int)f0 + (int)f1 + (int)f2 + (int)f3) >= 3;
        const uint32_t fm = (uint32_t)f0 | ((uint32_t)f1 << 1) | ((uint32_t)f2 << 2) | ((uint32_t)f3 << 3);
        const bool front = ((fm >> rot) & 1u) && ((fm >> ((rot + 1) & 3)) & 1u);
        if (!front && !ulk) mini = true;
    }
    const bool s1 = g->ruleset == TRL_RULESET_S1;
    if (!tspin && !s1) {  // s2 all-spin: immobile at its own (x, y) (player.py:145-151)
        const bool movable = trl_fits(p->rows, minos, x - 1, y) || trl_fits(p->rows, minos, x + 1, y) ||
                             trl_fits(p->rows, minos, x, y - 1) || trl_fits(p->rows, minos, x, y + 1);
        if (!movable) mini = true;
    }

    uint32_t touched = 0;  // bit r of a 64-bit set would be needed for 40 rows: keep lo/hi
    uint32_t touched_hi = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int c = x + (int)((minos >> (8 * m)) & 15u);
        const int r = py + (int)((minos >> (8 * m + 4)) & 15u);
        if ((unsigned)r < (unsigned)TRL_ROWS && (unsigned)c < (unsigned)TRL_COLS) {
            p->rows[r] |= (uint16_t)(1u << c);
            if (r < 32) touched |= 1u << r; else touched_hi |= 1u << (r - 32);
        }
    }
    p->piece = TRL_NONE;

    int n_cleared = 0;  // player.py:161-176: drop full touched rows, empty rows enter on top
    bool any_full = false;   // most placements clear nothing: skip the 40-row compaction then
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int r = py + (int)((minos >> (8 * m + 4)) & 15u);
        if ((unsigned)r < (unsigned)TRL_ROWS && (p->rows[r] & TRL_FULL_ROW) == TRL_FULL_ROW) any_full = true;
    }
    if (any_full) {
        int w = TRL_ROWS - 1;
        for (int r = TRL_ROWS - 1; r >= 0; --r) {
            const bool t = (r < 32) ? ((touched >> r) & 1u) : ((touched_hi >> (r - 32)) & 1u);
            const uint16_t v = p->rows[r];
            if (t && (v & TRL_FULL_ROW) == TRL_FULL_ROW) { ++n_cleared; continue; }
            p->rows[w--] = v;
        }
        while (w >= 0) p->rows[w--] = 0;
    }
    if (n_cleared > 0) {  // player.py:178-179
        all_clear = true;
        for (int r = n_cleared; r < TRL_ROWS; ++r)
            if (p->rows[r] & TRL_FULL_ROW) all_clear = false;
    }

    int combo = p->combo, b2b = p->b2b, level = p->b2b_level;
    const int attack = s1 ? trl_get_attack_s1(n_cleared, tspin, mini, all_clear, combo, b2b, level)
                          : trl_get_attack_s2(n_cleared, tspin, mini, all_clear, combo, b2b, level);
    p->combo = (int16_t)combo; p->b2b = (int16_t)b2b; p->b2b_level = (uint8_t)level;
    p->pieces += 1;

    int send_n = 0, send_col = 0;
    if (attack > 0) {  // one hole column per attack (player.py:184-186)
        send_col = trl_garbage_column(seed, g->game_id, stream, *ctr);
        *ctr += 1;
        send_n = attack;
    }

    // ---- Game.check_garbage (game.py:100-117) ----
    int n_recv = p->n_recv;
    const int cancel = min(send_n, n_recv);
    if (cancel > 0) {
        send_n -= cancel;
        n_recv -= cancel;
        for (int i = 0; i < n_recv; ++i) p->recv[i] = p->recv[i + cancel];
    }
    if (n_recv > 0 && n_cleared == 0) {  // Player.spawn_garbage -> Board.create_garbage (board.py:12-21)
        const int skip = max(n_recv - TRL_ROWS, 0);
        const int n = n_recv - skip;
        for (int r = 0; r < TRL_ROWS - n; ++r) p->rows[r] = p->rows[r + n];
        for (int i = 0; i < n; ++i)
            p->rows[TRL_ROWS - n + i] = (uint16_t)(TRL_FULL_ROW & ~(1u << p->recv[skip + i]));
        n_recv = 0;
        o.flags |= 0x20;
    }
    p->n_recv = (uint8_t)n_recv;
    for (int i = 0; i < send_n; ++i) {
        if (opp->n_recv >= TRL_RECV_CAP) { o.status |= TRL_ST_RECV_OVERFLOW; break; }
        opp->recv[opp->n_recv++] = (uint8_t)send_col;
    }

    trl_create_next_piece(p);                                // game.py:80
    if (add_bag && p->qlen < 5) trl_add_bag_to_all(g, seed);  // game.py:82-83
    if (add_bag && turn == 1) g->rounds += 1;                // game.py:86-87 (history length)
    g->turn = (uint8_t)(1 - turn);                           // game.py:89

    o.rows_cleared = (uint8_t)n_cleared;
    o.attack = (uint8_t)attack;
    o.flags |= (uint8_t)((tspin ? 1 : 0) | (mini ? 2 : 0) | (all_clear ? 4 : 0) | (p->game_over ? 0x10 : 0));
    o.garbage_col = (uint8_t)send_col;
    return o;
}

// Game() + Game.setup() (game.py:8-38).  Scalar.
__device__ __forceinline__ void trl_game_setup_scalar(TrlGame* g, uint32_t game_id, uint64_t seed, uint8_t ruleset = TRL_RULESET_S2) {
    uint32_t* w = reinterpret_cast<uint32_t*>(g);
    for (int i = 0; i < (int)(sizeof(TrlGame) / 4); ++i) w[i] = 0;
    g->game_id = game_id;
    g->ruleset = ruleset;
    for (int pl = 0; pl < 2; ++pl) {
        g->players[pl].b2b = -1;
        g->players[pl].piece = TRL_NONE;
        g->players[pl].held = TRL_NONE;
    }
    trl_add_bag_to_all(g, seed);
    for (int pl = 0; pl < 2; ++pl) trl_create_next_piece(&g->players[pl]);
    g->rounds = 1;
}
