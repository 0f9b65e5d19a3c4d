/*
 * ppad_oracle.c -- CPU restatement of the P-PAD game logic.  TEST INFRASTRUCTURE, NOT PRODUCT.
 * See ppad_oracle.h for the scope, the reference file:line map and the parity pin.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: the float64 reward must not be FMA-contracted)
 */
#include "ppad_oracle.h"

#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------ defaults (parameters.py:43,54-89) */
void ppad_oracle_default_cfg(ppad_oracle_cfg *cfg)
{
    /* orb types: 0 red 1 blue 2 green 3 light 4 dark 5 heal (parameters.py:20-29) */
    static const int c1[6] = {4, 3, 1, 2, 0, 1};   /* dark, light, blue, green, red, blue */
    static const int c2[6] = {4, 3, 2, 0, 1, -1};  /* dark, light, green, red, blue, None */
    memset(cfg, 0, sizeof(*cfg));
    cfg->rows = 5;
    cfg->cols = 6;
    cfg->skyfall_damage = 1;
    cfg->cascade_cap = 255;
    cfg->reset_cap = 64;
    cfg->team_n = 6;
    for (int i = 0; i < 6; i++) {
        cfg->team_c1[i] = c1[i];
        cfg->team_c2[i] = c2[i];
        cfg->team_atk[i] = 1000.0;
    }
    for (int i = 0; i < 6; i++) cfg->skyfall[i] = 1.0 / 6;   /* the literal 1/6 of parameters.py:43 */
}

/* ------------------------------------------------------------------ Philox4x32-10 */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void ppad_oracle_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; r++) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

uint32_t ppad_oracle_draw(uint64_t seed, uint64_t env, uint32_t step, uint32_t stream, uint32_t d)
{
    uint32_t ctr[4] = {(uint32_t)env, (uint32_t)(env >> 32), step, (stream << 24) | (d >> 2)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    ppad_oracle_philox4x32(ctr, key, out);
    return out[d & 3];
}

/* game.py:458-470 with random_number = x / 2^32 */
int ppad_oracle_orb_from_u32(const double prob[10], uint32_t x)
{
    const double u = (double)x / 4294967296.0;
    double prob_sum = 0;
    for (int i = 0; i < 10; i++) {
        prob_sum += prob[i];
        if (prob_sum >= u && prob[i] > 0) return i;
    }
    return 9;
}

/* ------------------------------------------------------------------ cancel (utils.py:26-70, 86-111) */
static void flood(const int8_t *board, uint8_t *island, int rows, int cols, int start, int orb)
{
    /* explicit stack instead of the reference's recursion; the filled set is the same */
    int stack[PPAD_ORACLE_MAX_CELLS * 4];
    int sp = 0;
    stack[sp++] = start;
    while (sp) {
        const int p = stack[--sp];
        if (board[p] != orb || island[p]) continue;
        island[p] = 1;
        const int x = p / cols, y = p % cols;
        if (y + 1 < cols && !island[p + 1]) stack[sp++] = p + 1;
        if (y - 1 >= 0 && !island[p - 1]) stack[sp++] = p - 1;
        if (x + 1 < rows && !island[p + cols]) stack[sp++] = p + cols;
        if (x - 1 >= 0 && !island[p - cols]) stack[sp++] = p - cols;
    }
}

int ppad_oracle_cancel(int8_t *board, int rows, int cols, int32_t *colors, int32_t *counts)
{
    const int cells = rows * cols;
    int n = 0;
    for (int p = 0; p < cells; p++) {
        const int orb = board[p];
        if (orb == -1) continue;
        uint8_t island[PPAD_ORACLE_MAX_CELLS] = {0};
        uint8_t pruned[PPAD_ORACLE_MAX_CELLS] = {0};
        flood(board, island, rows, cols, p, orb);
        for (int k = 0; k + 3 <= rows; k++)           /* vertical windows, utils.py:47-50 */
            for (int l = 0; l < cols; l++)
                if (island[k * cols + l] + island[(k + 1) * cols + l] + island[(k + 2) * cols + l] == 3)
                    pruned[k * cols + l] = pruned[(k + 1) * cols + l] = pruned[(k + 2) * cols + l] = 1;
        for (int k = 0; k < rows; k++)                /* horizontal windows, utils.py:51-54 */
            for (int l = 0; l + 3 <= cols; l++)
                if (island[k * cols + l] + island[k * cols + l + 1] + island[k * cols + l + 2] == 3)
                    pruned[k * cols + l] = pruned[k * cols + l + 1] = pruned[k * cols + l + 2] = 1;
        int count = 0;
        for (int q = 0; q < cells; q++) count += pruned[q];
        if (count >= 3) {
            for (int q = 0; q < cells; q++)
                if (pruned[q]) board[q] = -1;
            if (colors) colors[n] = orb;
            if (counts) counts[n] = count;
            n++;
        }
    }
    return n;
}

/* ------------------------------------------------------------------ drop (game.py:201-221) */
void ppad_oracle_drop(int8_t *board, int rows, int cols)
{
    for (int y = 0; y < cols; y++) {
        int8_t col[PPAD_ORACLE_MAX_CELLS];
        int m = 0;
        for (int x = 0; x < rows; x++)
            if (board[x * cols + y] != -1) col[m++] = board[x * cols + y];
        for (int x = 0; x < rows; x++) board[x * cols + y] = -1;
        const int canceled = rows - m;
        for (int x = 0; x < m; x++) board[(x + canceled) * cols + y] = col[x];
    }
}

/* ------------------------------------------------------------------ damage (game.py:157-199) */
double ppad_oracle_damage(const ppad_oracle_cfg *cfg, const int32_t *colors, const int32_t *counts, int n)
{
    /* damage_tracker dict order: blue, red, green, dark, light (game.py:163-167) */
    static const int order[5] = {1, 0, 2, 4, 3};
    double tracker[10] = {0};
    for (int c = 0; c < n; c++) {
        const int color = colors[c];
        for (int u = 0; u < cfg->team_n; u++) {
            double multiplier;
            if (cfg->team_c1[u] == cfg->team_c2[u] && cfg->team_c2[u] == color) multiplier = 1.1;
            else if (cfg->team_c1[u] == color) multiplier = 1;
            else if (cfg->team_c2[u] == color) multiplier = 0.3;
            else continue;
            if (color > 4 || color < 0) return NAN;      /* KeyError in the reference */
            multiplier *= 1 + 0.25 * ((double)counts[c] - 3);
            tracker[color] += multiplier * cfg->team_atk[u];
        }
    }
    double total = 0;
    for (int i = 0; i < 5; i++) {
        tracker[order[i]] *= 1 + 0.25 * (n - 1);
        total += tracker[order[i]];
    }
    return total;
}

/* ------------------------------------------------------------------ orb source: injected then Philox */
typedef struct {
    const uint8_t *inj;
    int n_inj;          /* < 0: pure Philox */
    int cursor;         /* orbs handed out so far */
    int exhausted;
    uint64_t seed, env;
    uint32_t step, stream;
    const double *prob;
} orb_source;

static int next_orb(orb_source *s)
{
    const int d = s->cursor++;
    if (s->n_inj >= 0) {
        if (d < s->n_inj) return s->inj[d];
        s->exhausted = 1;
    }
    return ppad_oracle_orb_from_u32(s->prob, ppad_oracle_draw(s->seed, s->env, s->step, s->stream, (uint32_t)d));
}

/* game.py:232-256: row-major over empties (or every cell when reset) */
static void fill(int8_t *board, int cells, int reset, orb_source *s)
{
    for (int p = 0; p < cells; p++)
        if (board[p] == -1 || reset) board[p] = (int8_t)next_orb(s);
}

static int sat255(int v) { return v > 255 ? 255 : v; }

/* ------------------------------------------------------------------ calculate_reward (game.py:364-405) */
void ppad_oracle_resolve(const ppad_oracle_cfg *cfg, const int8_t *board_in, const uint8_t *inj, int n_inj,
                         uint64_t env, uint32_t step, ppad_oracle_result *res, int8_t *final_board,
                         uint16_t *combo_log, int log_cap)
{
    enum { MAXC = 8192 };
    static _Thread_local int32_t colors[MAXC], counts[MAXC];
    const int cells = cfg->rows * cfg->cols;
    int8_t board[PPAD_ORACLE_MAX_CELLS];
    memcpy(board, board_in, (size_t)cells);
    orb_source src = {inj, n_inj, 0, 0, cfg->seed, env, step, PPAD_OR_STREAM_SKYFALL, cfg->skyfall};
    int n = 0, depth = 0, flags = 0;
    for (;;) {
        if (n + cells > MAXC) { flags |= PPAD_OR_CASCADE_CAP; break; }
        const int m = ppad_oracle_cancel(board, cfg->rows, cfg->cols, colors + n, counts + n);
        if (m < 1) break;
        n += m;
        depth++;
        if (!cfg->skyfall_damage) break;
        if (cfg->cascade_cap > 0 && depth >= cfg->cascade_cap) { flags |= PPAD_OR_CASCADE_CAP; break; }
        ppad_oracle_drop(board, cfg->rows, cfg->cols);
        fill(board, cells, 0, &src);
    }
    if (src.exhausted) flags |= PPAD_OR_SKYFALL_EXHAUSTED;
    res->reward = ppad_oracle_damage(cfg, colors, counts, n);
    res->n_combos = n;
    res->depth = depth;
    res->consumed = src.cursor;
    res->flags = flags;
    if (final_board) memcpy(final_board, board, (size_t)cells);
    if (combo_log)
        for (int i = 0; i < n && i < log_cap; i++) combo_log[i] = (uint16_t)(colors[i] | (counts[i] << 8));
}

/* ------------------------------------------------------------------ moves (game.py:122-155, 429-456) */
int ppad_oracle_apply_action(int8_t *board, int rows, int cols, int32_t finger[2], int action)
{
    int tx = finger[0], ty = finger[1];
    switch (action) {
    case 0: tx -= 1; if (tx < 0) return 0; break;              /* up    */
    case 1: tx += 1; if (tx > rows - 1) return 0; break;       /* down  */
    case 2: ty -= 1; if (ty < 0) return 0; break;              /* left  */
    case 3: ty += 1; if (ty > cols - 1) return 0; break;       /* right */
    default: return 0;
    }
    const int a = finger[0] * cols + finger[1], b = tx * cols + ty;
    const int8_t t = board[a];
    board[a] = board[b];
    board[b] = t;
    finger[0] = tx;
    finger[1] = ty;
    return 1;
}

/* player.py:52-62: not the reverse of the previous action, not off the board */
int ppad_oracle_legal_mask(int rows, int cols, const int32_t finger[2], int prev)
{
    static const int opposite[5] = {1, 0, 3, 2, 4};
    int mask = 0;
    for (int a = 0; a < 4; a++) {
        int ok = 1;
        if (opposite[a] == prev) ok = 0;
        else if (finger[0] >= rows - 1 && a == 1) ok = 0;
        else if (finger[0] <= 0 && a == 0) ok = 0;
        else if (finger[1] >= cols - 1 && a == 3) ok = 0;
        else if (finger[1] <= 0 && a == 2) ok = 0;
        if (ok) mask |= 1 << a;
    }
    return mask;
}

int ppad_oracle_sample_action(uint64_t seed, uint64_t env, uint32_t step, int legal_mask)
{
    int n = 0;
    for (int a = 0; a < 4; a++) n += (legal_mask >> a) & 1;
    if (n == 0) return -1;
    /* word step & 3 of the block (env, step >> 2, ACTION): four consecutive steps share one Philox block */
    const uint32_t x = ppad_oracle_draw(seed, env, step >> 2, PPAD_OR_STREAM_ACTION, step & 3u);
    int pick = (int)(((uint64_t)x * (uint64_t)n) >> 32);
    for (int a = 0; a < 4; a++)
        if ((legal_mask >> a) & 1) {
            if (pick == 0) return a;
            pick--;
        }
    return -1;
}

/* ------------------------------------------------------------------ reset (game.py:282-323) */
int ppad_oracle_reset(const ppad_oracle_cfg *cfg, const uint8_t *inj, int n_inj, uint64_t env, uint32_t step,
                      const int32_t *finger_in, int8_t *board, int32_t finger[2], int32_t *consumed)
{
    const int cells = cfg->rows * cfg->cols;
    orb_source src = {inj, n_inj, 0, 0, cfg->seed, env, step, PPAD_OR_STREAM_RESET, cfg->skyfall};
    int flags = 0, attempts = 0;
    for (;;) {
        int8_t trial[PPAD_ORACLE_MAX_CELLS];
        fill(board, cells, 1, &src);
        attempts++;
        memcpy(trial, board, (size_t)cells);
        /* the reference keeps the board that cancel() left untouched, i.e. the combo-free one */
        if (ppad_oracle_cancel(trial, cfg->rows, cfg->cols, 0, 0) == 0) break;
        if (cfg->reset_cap > 0 && attempts >= cfg->reset_cap) { flags |= PPAD_OR_RESET_CAP; break; }
    }
    if (src.exhausted) flags |= PPAD_OR_SKYFALL_EXHAUSTED;
    if (finger_in) {
        finger[0] = finger_in[0];
        finger[1] = finger_in[1];
    } else {
        const uint32_t x0 = ppad_oracle_draw(cfg->seed, env, step, PPAD_OR_STREAM_FINGER, 0);
        const uint32_t x1 = ppad_oracle_draw(cfg->seed, env, step, PPAD_OR_STREAM_FINGER, 1);
        finger[0] = (int32_t)(((uint64_t)x0 * (uint64_t)cfg->rows) >> 32);
        finger[1] = (int32_t)(((uint64_t)x1 * (uint64_t)cfg->cols) >> 32);
    }
    if (consumed) *consumed = src.cursor;
    return flags;
}

/* ------------------------------------------------------------------ one-hot (agent01.py:254-273) */
void ppad_oracle_onehot(const int8_t *boards, const int32_t *fingers, int64_t n, int rows, int cols, float *out)
{
    const int cells = rows * cols;
    memset(out, 0, sizeof(float) * (size_t)n * (size_t)cells * 7);
    for (int64_t i = 0; i < n; i++) {
        float *o = out + i * cells * 7;
        for (int p = 0; p < cells; p++) {
            const int v = boards[i * cells + p];
            if (v >= 0 && v < 6) o[p * 7 + v] = 1.0f;
        }
        o[(fingers[2 * i] * cols + fingers[2 * i + 1]) * 7 + 6] = 1.0f;
    }
}

/* ------------------------------------------------------------------ discount (agent/utils.py:22-51) */
void ppad_oracle_discount(double *rewards, int n, double gamma, int log10_flag)
{
    int any_pos = 0;
    for (int i = 0; i < n; i++) any_pos |= rewards[i] > 0;
    if (!any_pos) return;
    if (log10_flag)
        for (int i = 0; i < n; i++)
            if (rewards[i] > 0) rewards[i] = log10(rewards[i]);
    double cur = 0.0;
    for (int i = n - 1; i >= 0; i--) {
        cur = cur * gamma + rewards[i];
        rewards[i] = cur;
    }
}

/* ------------------------------------------------------------------ batched step (kernel contract) */
void ppad_oracle_step_batch(const ppad_oracle_cfg *cfg, int64_t n, uint64_t env0, uint32_t step,
                            int8_t *boards, int32_t *fingers, uint8_t *prev, uint16_t *moves,
                            const uint8_t *actions, int eval_moves, int auto_reset, int max_moves,
                            const uint8_t *inj, int inj_stride, int n_inj,
                            uint8_t *action_out, double *reward, uint32_t *info, int8_t *final_boards)
{
    const int cells = cfg->rows * cfg->cols;
    for (int64_t i = 0; i < n; i++) {
        const uint64_t env = env0 + (uint64_t)i;
        int8_t *board = boards + i * cells;
        int32_t *finger = fingers + 2 * i;
        int flags = 0, action, evaluate = 0;
        if (actions) {
            action = actions[i];
        } else if (max_moves > 0 && moves && moves[i] >= max_moves) {
            action = 4;                             /* episode length cap: simulation.py:50-52,91-93 */
        } else {
            action = ppad_oracle_sample_action(cfg->seed, env, step,
                                               ppad_oracle_legal_mask(cfg->rows, cfg->cols, finger, prev[i]));
            if (action < 0) { action = 255; flags |= PPAD_OR_NO_LEGAL_MOVE; }
        }
        if (action_out) action_out[i] = (uint8_t)action;
        if (action == 4) {
            evaluate = 1;
        } else if (action < 4) {
            if (ppad_oracle_apply_action(board, cfg->rows, cfg->cols, finger, action)) {
                prev[i] = (uint8_t)action;          /* game.py:351 */
                if (moves && moves[i] < 65535) moves[i]++;
                evaluate = eval_moves;
            } else {
                flags |= PPAD_OR_REJECTED;          /* game.py:353 raises ValueError */
            }
        } else if (action == 5) {
            /* no-op: the board sits this step out (a finished episode in a lock-step batch) */
        } else if (!(flags & PPAD_OR_NO_LEGAL_MOVE)) {
            flags |= PPAD_OR_BAD_ACTION;
        }
        ppad_oracle_result res = {0.0, 0, 0, 0, 0};
        if (evaluate) {
            ppad_oracle_resolve(cfg, board, inj ? inj + (size_t)i * (size_t)inj_stride : 0, inj ? n_inj : -1,
                                env, step, &res, final_boards ? final_boards + i * cells : 0, 0, 0);
        } else if (final_boards) {
            memcpy(final_boards + i * cells, board, (size_t)cells);
        }
        flags |= res.flags;
        if (action == 4 && auto_reset) {
            flags |= PPAD_OR_EPISODE_DONE;
            flags |= ppad_oracle_reset(cfg, 0, -1, env, step, 0, board, finger, 0);
            prev[i] = 4;                            /* game.py:299 */
            if (moves) moves[i] = 0;
        }
        if (reward) reward[i] = res.reward;
        if (info)
            info[i] = (uint32_t)sat255(res.n_combos) | ((uint32_t)sat255(res.depth) << 8) |
                      ((uint32_t)sat255(res.consumed) << 16) | ((uint32_t)(flags & 0xFF) << 24);
    }
}

/* ------------------------------------------------------------------ whole finger paths (kernel contract)
 * Board i applies lens[i] moves paths[i * path_stride + t] (ids 0..3) like that many PAD.step(move) calls
 * (game.py:329-362; off-board moves are skipped and flagged), then is evaluated once when eval != 0. */
void ppad_oracle_path_batch(const ppad_oracle_cfg *cfg, int64_t n, uint64_t env0, uint32_t step,
                            int8_t *boards, int32_t *fingers, uint8_t *prev, uint16_t *moves,
                            const uint8_t *paths, int path_stride, const uint16_t *lens, int eval,
                            const uint8_t *inj, int inj_stride, int n_inj,
                            uint8_t *action_out, double *reward, uint32_t *info, int8_t *final_boards)
{
    const int cells = cfg->rows * cfg->cols;
    for (int64_t i = 0; i < n; i++) {
        int8_t *board = boards + i * cells;
        int32_t *finger = fingers + 2 * i;
        int flags = 0, last = 255;
        for (int t = 0; t < lens[i]; t++) {
            const int a = paths[(size_t)i * (size_t)path_stride + t] & 3;
            if (ppad_oracle_apply_action(board, cfg->rows, cfg->cols, finger, a)) {
                prev[i] = (uint8_t)a;
                last = a;
                if (moves && moves[i] < 65535) moves[i]++;
            } else {
                flags |= PPAD_OR_REJECTED;
            }
        }
        ppad_oracle_result res = {0.0, 0, 0, 0, 0};
        if (eval)
            ppad_oracle_resolve(cfg, board, inj ? inj + (size_t)i * (size_t)inj_stride : 0, inj ? n_inj : -1,
                                env0 + (uint64_t)i, step, &res, final_boards ? final_boards + i * cells : 0, 0, 0);
        else if (final_boards)
            memcpy(final_boards + i * cells, board, (size_t)cells);
        flags |= res.flags;
        if (action_out) action_out[i] = (uint8_t)last;
        if (reward) reward[i] = res.reward;
        if (info)
            info[i] = (uint32_t)sat255(res.n_combos) | ((uint32_t)sat255(res.depth) << 8) |
                      ((uint32_t)sat255(res.consumed) << 16) | ((uint32_t)(flags & 0xFF) << 24);
    }
}
