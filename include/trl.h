/*
 * trl.h — C ABI of the B200-native self-play hot path (libtrl_b200.so).
 *
 * Drop-in boundary for the data-generation path of mat-lee/tetris-reinforcement-learning.
 * Every entry point names the reference interface it replaces (file:line into the
 * reference checkout).  The reference is pure Python, so the binding a maintainer adds
 * is a ctypes stub.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types in any signature.
 *   - `*_dev` style entry points (the default, no suffix) take caller-allocated DEVICE
 *     pointers and a `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *     They never allocate, never synchronise, are stream-ordered and CUDA-graph capturable.
 *   - `*_host` entry points take HOST pointers, stage through an internal device
 *     workspace and synchronise before returning (this is what a reference-side caller
 *     such as get_move_matrix(player) binds to).
 *   - return value: 0 = TRL_OK, negative = error (see TRL_E_*); never throws.
 *   - per-item anomalies (queue overflow, no legal move, ...) are reported in a per-item
 *     `status` word (TRL_ST_* bits), not in the return value.
 *
 * Encodings (reference: const.py:6-9, 68, 82-123, 238-281)
 *   - piece id = index into MINOS = "ZLOSIJT": Z0 L1 O2 S3 I4 J5 T6; 255 = none.
 *   - board = 40 rows x uint16; row 0 = top, row 39 = floor; bit c = column c occupied
 *     (reference keeps an object grid; only `!= 0` matters to the rules, board.py:23-30).
 *   - policy tensor (27, 39, 11): flat index = (plane*39 + row)*11 + (x+2), 11583 bits,
 *     bit-packed little-endian into 362 uint32 words (bit i -> word i>>5, bit i&31).
 *   - move = flat policy index (uint16) <-> reference tuple (plane, col=x, row) (ai.py:1022).
 */
#ifndef TRL_H_
#define TRL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRL_ROWS 40
#define TRL_COLS 10
#define TRL_SPAWN_Y 17           /* ROWS - SPAWN_ROW, const.py:6-7, piece.py:17-21 */
#define TRL_PREVIEWS 5           /* const.py PREVIEWS */
#define TRL_PLANES 27
#define TRL_POLICY_ROWS 39
#define TRL_POLICY_COLS 11
#define TRL_POLICY_SIZE 11583    /* 27*39*11, const.py:122-123 */
#define TRL_MASK_WORDS 362       /* ceil(11583/32) */
#define TRL_NONE 255
#define TRL_QUEUE_CAP 16
#define TRL_RECV_CAP 80

/* return codes */
#define TRL_OK 0
#define TRL_E_ARG (-1)           /* bad argument (NULL pointer, negative size, ...) */
#define TRL_E_CUDA (-2)          /* a CUDA runtime call failed; see trl_last_error() */
#define TRL_E_NOMEM (-3)

/* per-item status bits */
#define TRL_ST_QUEUE_OVERFLOW 0x1u   /* movegen exploration FIFO exceeded its capacity  */
#define TRL_ST_MOVES_TRUNC    0x2u   /* compact move list truncated at moves_cap        */
#define TRL_ST_NO_PIECE       0x4u   /* cur == none and alt == none (Game.no_move)      */
#define TRL_ST_RECV_OVERFLOW  0x8u   /* pending-garbage list exceeded TRL_RECV_CAP      */
#define TRL_ST_BAD_MOVE       0x10u  /* env step: move index out of range / undecodable */

/*
 * One player.  Mirrors the fields Player.copy() carries (player.py:217-233):
 * board, queue, stats.{pieces,b2b,b2b_level,combo}, game_over, piece, held_piece,
 * garbage_to_receive.  The active piece is stored as a type only: between moves it always
 * sits at its spawn location (piece.py:17-21), Game.move_piece overwrites x/y/rot
 * (game.py:56-64).  192 bytes.
 */
typedef struct TrlPlayer {
    uint16_t rows[TRL_ROWS];         /* 0   board bitrows                                 */
    int32_t  pieces;                 /* 80  stats.pieces                                  */
    int16_t  b2b;                    /* 84  stats.b2b (starts at -1)                      */
    int16_t  combo;                  /* 86  stats.combo                                   */
    uint8_t  qlen;                   /* 88  len(queue.pieces)                             */
    uint8_t  piece;                  /* 89  active piece type or TRL_NONE                 */
    uint8_t  held;                   /* 90  held piece type or TRL_NONE                   */
    uint8_t  game_over;              /* 91                                               */
    uint8_t  b2b_level;              /* 92  stats.b2b_level                               */
    uint8_t  n_recv;                 /* 93  len(garbage_to_receive)                       */
    uint8_t  pad_[2];                /* 94                                               */
    uint8_t  queue[TRL_QUEUE_CAP];   /* 96  queue.pieces, index 0 = next                  */
    uint8_t  recv[TRL_RECV_CAP];     /* 112 garbage_to_receive hole columns, 0 = first    */
} TrlPlayer;

/* Two-player game (game.py:6-32).  400 bytes. */
typedef struct TrlGame {
    TrlPlayer players[2];            /* 0                                                 */
    uint8_t  turn;                   /* 384 index of the side to move                     */
    uint8_t  pad_;
    uint16_t bag_ctr;                /* 386 number of 7-bag refills dealt so far          */
    uint32_t rounds;                 /* 388 len(history.states) (game.py:86-87)           */
    uint32_t rng_ctr;                /* 392 counter for the next garbage-column draw      */
    uint32_t game_id;                /* 396 Philox stream id (global game index)          */
} TrlGame;

/* Per-step outputs of the env step (what Player.place_piece computes, player.py:109-188). */
typedef struct TrlStepOut {
    uint8_t rows_cleared;            /* 0..4                                              */
    uint8_t attack;                  /* lines sent by Stats.get_attack (stats.py:88-129)  */
    uint8_t flags;                   /* bit0 tspin, bit1 mini, bit2 all_clear, bit3 held, */
                                     /* bit4 mover topped out, bit5 garbage received      */
    uint8_t garbage_col;             /* hole column drawn (valid iff attack > 0)          */
    uint32_t status;                 /* TRL_ST_* bits                                     */
} TrlStepOut;

/* ------------------------------------------------------------------------------------ */
/* library                                                                               */
/* ------------------------------------------------------------------------------------ */

/* ABI version of this header; bumped on any incompatible change. */
int trl_abi_version(void);
/* Text of the last CUDA error seen by this thread's calls ("" if none). */
const char* trl_last_error(void);
/* sizeof checks for bindings: returns sizeof(TrlPlayer) / sizeof(TrlGame). */
int trl_sizeof_player(void);
int trl_sizeof_game(void);

/* ------------------------------------------------------------------------------------ */
/* legal-placement enumeration                                                           */
/* replaces move_generation.get_move_matrix(player, algo='convolutional')                */
/*   (move_generation.py:752-789 -> MoveGenerator.generate_moves :77-105,                */
/*    _generate_moves_for_piece :107-149, _convolutional_algorithm :325-488,             */
/*    _build_validity_maps :490-528, _convert_placements_to_policy :650-749)             */
/* and ai.get_move_list's argwhere ordering (ai.py:1016-1024).                           */
/* ------------------------------------------------------------------------------------ */

/*
 * boards   [n][40] uint16   board of the player to move
 * cur      [n]     uint8    active piece type (TRL_NONE if player.piece is None)
 * alt      [n]     uint8    held piece if any, else queue[0] if any, else TRL_NONE
 *                           (move_generation.py:91-105); the alt piece is only legal if it
 *                           can spawn (Player.hold_piece -> create_piece, player.py:37-44)
 * mask_bits[n][362] uint32  OUT bit-packed (27,39,11) legal mask; may be NULL
 * moves    [n][moves_cap] uint16 OUT ascending flat policy indices (= np.argwhere order);
 *                           may be NULL
 * n_moves  [n]     uint16   OUT number of legal placements (always the full count); may be NULL
 * status   [n]     uint32   OUT TRL_ST_* bits; may be NULL
 */
int trl_movegen(const uint16_t* boards, const uint8_t* cur, const uint8_t* alt, int n,
                uint32_t* mask_bits, uint16_t* moves, int moves_cap, uint16_t* n_moves,
                uint32_t* status, void* stream);

/* Same, on the side-to-move player of packed games (cur/alt derived on device). */
int trl_movegen_games(const TrlGame* games, int n, uint32_t* mask_bits, uint16_t* moves,
                      int moves_cap, uint16_t* n_moves, uint32_t* status, void* stream);

/* HOST-buffer variant of trl_movegen: copies in, runs, copies out, synchronises. */
int trl_movegen_host(const uint16_t* boards, const uint8_t* cur, const uint8_t* alt, int n,
                     uint32_t* mask_bits, uint16_t* moves, int moves_cap, uint16_t* n_moves,
                     uint32_t* status);

/* ------------------------------------------------------------------------------------ */
/* env step                                                                              */
/* replaces Game.make_move(move, add_bag, add_history=False)                             */
/*   (game.py:40-118: move_piece, place, check_garbage; player.py:29-44, 109-205;        */
/*    board.py:12-21; stats.py:30-43, 88-129 [ruleset s2]; piece_queue.py:17-20)          */
/* ------------------------------------------------------------------------------------ */

/*
 * games  [n] TrlGame  IN/OUT, stepped in place
 * moves  [n] uint16   flat policy index of the move of the side to move; 0xFFFF = skip item
 * out    [n] TrlStepOut OUT; may be NULL
 * add_bag  != 0: refill both queues with a fresh 7-bag when the mover's queue drops below 5
 *          (game.py:82-83); bags come from Philox(seed, game_id, bag stream)
 * seed     Philox key; the garbage hole column (player.py:184-186) is
 *          draw(seed, game_id, games[i].rng_ctr++) — one draw per placement with attack > 0
 */
int trl_env_step(TrlGame* games, const uint16_t* moves, int n, TrlStepOut* out, int add_bag,
                 uint64_t seed, void* stream);
int trl_env_step_host(TrlGame* games, const uint16_t* moves, int n, TrlStepOut* out,
                      int add_bag, uint64_t seed);

/*
 * Game() + Game.setup() (game.py:8-38): empty boards, one 7-bag each (player 0 first),
 * first piece spawned, turn = 0, rounds = 1.  game_id[i] = first_game_id + i.
 */
int trl_game_setup(TrlGame* games, int n, uint32_t first_game_id, uint64_t seed, void* stream);
int trl_game_setup_host(TrlGame* games, int n, uint32_t first_game_id, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* TRL_H_ */
