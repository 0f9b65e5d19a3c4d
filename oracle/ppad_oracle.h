/*
 * ppad_oracle.h -- CPU restatement of the P-PAD game logic.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * This is the parity oracle for the B200 kernels: a plain, scalar, array-based C restatement of
 * the reference's Python algorithm (nuwapi/P-PAD, files under /root/reference):
 *     ppad/pad/utils.py:26-70,86-111     cancel, detect_island
 *     ppad/pad/game.py:122-155,429-456   apply_action, swap
 *     ppad/pad/game.py:157-199           damage
 *     ppad/pad/game.py:201-221           drop
 *     ppad/pad/game.py:232-256,458-470   fill_board, random_orb
 *     ppad/pad/game.py:364-405           calculate_reward (cascade loop)
 *     ppad/pad/game.py:282-323           reset (rejection-sample a combo-free board)
 *     ppad/pad/player.py:38-66           PlayerAction.sample (legal set)
 *     ppad/agent/agent01.py:254-273      board_finger_to_state (one-hot)
 *     ppad/agent/utils.py:22-51          discount
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call it, and only as the checker / the CPU baseline -- never as a fallback for the CUDA path.
 *
 * Parity pin: the reference ships NO tests or golden vectors for this path,
 * so the oracle is pinned against outputs of the reference itself run in the build container:
 * tests/golden/gen_golden.py imports the unmodified reference and writes the fixtures that
 * tests/test_oracle_golden.py checks this file against.
 *
 * Boards are int8 [rows*cols] row-major, -1 = empty, 0..9 = orb type (game.py:236-246).
 */
#ifndef PPAD_ORACLE_H
#define PPAD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPAD_ORACLE_MAX_CELLS 64
#define PPAD_ORACLE_MAX_TEAM 8

/* per-board fault bits; values mirror include/ppad_b200.h (checked by tests/test_cabi_symbols.py) */
enum {
    PPAD_OR_REJECTED = 0x01, PPAD_OR_SKYFALL_EXHAUSTED = 0x02, PPAD_OR_CASCADE_CAP = 0x04,
    PPAD_OR_NO_LEGAL_MOVE = 0x08, PPAD_OR_BAD_ACTION = 0x10, PPAD_OR_RESET_CAP = 0x20,
    PPAD_OR_UNSUPPORTED_ORB = 0x40, PPAD_OR_EPISODE_DONE = 0x80
};

/* Philox stream ids (word 3 of the counter = stream << 24 | block) */
enum { PPAD_OR_STREAM_SKYFALL = 0, PPAD_OR_STREAM_ACTION = 1, PPAD_OR_STREAM_RESET = 2,
       PPAD_OR_STREAM_FINGER = 3, PPAD_OR_STREAM_POLICY = 4 };

typedef struct ppad_oracle_cfg {
    int32_t rows, cols;
    int32_t skyfall_damage;          /* game.py:45,387 */
    int32_t cascade_cap;             /* the reference has none; 0 = unlimited */
    int32_t reset_cap;               /* max rejection-sampling attempts in reset; 0 = unlimited */
    int32_t team_n;
    int32_t team_c1[PPAD_ORACLE_MAX_TEAM];   /* orb type of color_1, -1 = None */
    int32_t team_c2[PPAD_ORACLE_MAX_TEAM];
    double  team_atk[PPAD_ORACLE_MAX_TEAM];
    double  skyfall[10];             /* parameters.py:43 */
    uint64_t seed;                   /* Philox key */
} ppad_oracle_cfg;

void ppad_oracle_default_cfg(ppad_oracle_cfg *cfg);

/* Philox4x32-10 (Salmon et al., SC'11): ctr[4], key[2] -> out[4] */
void ppad_oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* the d-th 32-bit draw of (env, step, stream) */
uint32_t ppad_oracle_draw(uint64_t seed, uint64_t env, uint32_t step, uint32_t stream, uint32_t d);
/* categorical draw following game.py:458-470 with u = x / 2^32 */
int ppad_oracle_orb_from_u32(const double prob[10], uint32_t x);

/* utils.py:26-70.  Mutates board; writes (colour, N) per combo in reference order; returns #combos */
int ppad_oracle_cancel(int8_t *board, int rows, int cols, int32_t *colors, int32_t *counts);
/* game.py:201-221 */
void ppad_oracle_drop(int8_t *board, int rows, int cols);
/* game.py:157-199.  counts are the (integer valued) N of each combo.  NaN if a unit matches a
 * colour outside {blue, red, green, dark, light} (the reference raises KeyError there). */
double ppad_oracle_damage(const ppad_oracle_cfg *cfg, const int32_t *colors, const int32_t *counts, int n);

/* game.py:364-405 on a copy of `board`.  Skyfall orbs come from inj[0..n_inj) first, then from the
 * Philox stream (env, step, STREAM_SKYFALL) with flag SKYFALL_EXHAUSTED when n_inj >= 0 was given
 * and ran out (n_inj < 0 means "pure Philox", no flag).
 * combo_log (may be NULL) receives colour | N << 8 for up to log_cap combos. */
typedef struct ppad_oracle_result {
    double  reward;
    int32_t n_combos, depth, consumed, flags;
} ppad_oracle_result;
void ppad_oracle_resolve(const ppad_oracle_cfg *cfg, const int8_t *board, const uint8_t *inj, int n_inj,
                         uint64_t env, uint32_t step, ppad_oracle_result *res, int8_t *final_board,
                         uint16_t *combo_log, int log_cap);

/* game.py:122-155 + 429-456: returns 1 accepted, 0 rejected.  action 0..3 = up, down, left, right */
int ppad_oracle_apply_action(int8_t *board, int rows, int cols, int32_t finger[2], int action);
/* player.py:46-62 legal set as a 4-bit mask (bit a = action a allowed); prev = 0..4 (4 = pass) */
int ppad_oracle_legal_mask(int rows, int cols, const int32_t finger[2], int prev);
/* our Philox random policy: uniform over the legal set; returns -1 when the set is empty */
int ppad_oracle_sample_action(uint64_t seed, uint64_t env, uint32_t step, int legal_mask);

/* game.py:303-315: rejection-sample until cancel() finds nothing.  Orbs from inj first then Philox
 * (env, step, STREAM_RESET).  finger from STREAM_FINGER unless finger_in != NULL.  Returns flags. */
int ppad_oracle_reset(const ppad_oracle_cfg *cfg, const uint8_t *inj, int n_inj, uint64_t env, uint32_t step,
                      const int32_t *finger_in, int8_t *board, int32_t finger[2], int32_t *consumed);

/* agent01.py:254-273: [n, rows, cols, 7] float32 */
void ppad_oracle_onehot(const int8_t *boards, const int32_t *fingers, int64_t n, int rows, int cols, float *out);

/* agent/utils.py:22-51 without norm: in-place over rewards[n] */
void ppad_oracle_discount(double *rewards, int n, double gamma, int log10_flag);

/* One batched env step mirroring the kernel's contract (include/ppad_b200.h: ppad_step):
 *   action = actions ? actions[i] : (moves[i] >= max_moves > 0 ? pass : sampled);  4 = pass
 *   moves [n] uint16 (may be NULL) counts accepted moves of the episode
 *   move: apply (REJECTED flag + unchanged state when off-board), prev = action
 *   evaluate (pass, or eval_moves != 0): resolve on a copy -> reward/info
 *   pass && auto_reset: reset board/finger/prev (step-keyed Philox), flag EPISODE_DONE
 * boards [n, cells] int8, fingers [n,2] int32, prev [n] uint8 are updated in place.
 * info = n_combos | depth << 8 | consumed << 16 | flags << 24 (fields saturate at 255). */
void ppad_oracle_step_batch(const ppad_oracle_cfg *cfg, int64_t n, uint64_t env0, uint32_t step,
                            int8_t *boards, int32_t *fingers, uint8_t *prev, uint16_t *moves,
                            const uint8_t *actions, int eval_moves, int auto_reset, int max_moves,
                            const uint8_t *inj, int inj_stride, int n_inj,
                            uint8_t *action_out, double *reward, uint32_t *info, int8_t *final_boards);

/* Whole finger paths (include/ppad_b200.h: ppad_step_args.paths): lens[i] moves per board, one byte each */
void ppad_oracle_path_batch(const ppad_oracle_cfg *cfg, int64_t n, uint64_t env0, uint32_t step,
                            int8_t *boards, int32_t *fingers, uint8_t *prev, uint16_t *moves,
                            const uint8_t *paths, int path_stride, const uint16_t *lens, int eval,
                            const uint8_t *inj, int inj_stride, int n_inj,
                            uint8_t *action_out, double *reward, uint32_t *info, int8_t *final_boards);

#ifdef __cplusplus
}
#endif
#endif
