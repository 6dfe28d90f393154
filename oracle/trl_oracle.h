/* trl_oracle.h — prototypes of the CPU oracle (test infrastructure; see trl_oracle.c). */
#ifndef TRL_ORACLE_H_
#define TRL_ORACLE_H_
#include <stdint.h>
#include "../include/trl.h"
#ifdef __cplusplus
extern "C" {
#endif
void trl_oracle_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
int  trl_oracle_garbage_column(uint64_t seed, uint32_t game_id, uint32_t stream, uint32_t ctr);
double trl_oracle_uniform(uint64_t seed, uint32_t game_id, uint32_t search_no, uint32_t purpose, uint32_t idx);
double trl_oracle_gamma(uint64_t seed, uint32_t game_id, uint32_t search_no, uint32_t child, double alpha);
void trl_oracle_generate_bag(uint64_t seed, uint32_t game_id, uint32_t bag_ctr, int player, uint8_t bag[7]);
int  trl_oracle_movegen(const uint16_t* rows, int cur, int alt, uint8_t* mask, uint32_t* status,
                        int* n_push_out, int* n_emit_out);
int  trl_oracle_get_attack_s2(int rows_cleared, int is_tspin, int is_mini, int is_all_clear,
                              int* combo, int* b2b, int* b2b_level);
int  trl_oracle_get_attack_s1(int rows_cleared, int is_tspin, int is_mini, int is_all_clear,
                              int* combo, int* b2b, int* b2b_level);
void trl_oracle_game_setup(TrlGame* g, uint32_t game_id, uint64_t seed);
void trl_oracle_env_step(TrlGame* g, int move, int add_bag, uint64_t seed, TrlStepOut* out);
void trl_oracle_env_step_rng(TrlGame* g, int move, int add_bag, uint64_t seed, uint32_t stream,
                              uint32_t* ctr, TrlStepOut* out);
long long trl_oracle_movegen_batch(const uint16_t* boards, const uint8_t* cur, const uint8_t* alt,
                                   int n, uint32_t* mask_bits, uint16_t* n_moves, uint32_t* status,
                                   int n_threads);
void trl_oracle_env_step_batch(TrlGame* games, const uint16_t* moves, int n, TrlStepOut* out,
                               int add_bag, uint64_t seed);
void trl_oracle_game_setup_batch(TrlGame* games, int n, uint32_t first_game_id, uint64_t seed);
#ifdef __cplusplus
}
#endif
#endif
