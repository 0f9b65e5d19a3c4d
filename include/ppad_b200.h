/*
 * ppad_b200.h -- C ABI of the B200-native batched Puzzle & Dragons environment.
 *
 * The reference (nuwapi/P-PAD) is pure Python and has no FFI: its boundary for this path is the
 * duck-typed `PAD` object (ppad/pad/game.py:28-470) plus two free functions
 * (ppad/pad/utils.py:26-83) and one agent method (ppad/agent/agent01.py:254-273).  Each entry point
 * below names the reference interface it replaces for a batch of n boards; the Python layer in
 * p-pad_b200/ rebuilds the reference's object surface on top of them via ctypes.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls only
 *     enqueue work, they never synchronise (ppad_pipe_wait is the one exception);
 *   - every call makes the context's device current for its own duration and restores the caller's on return;
 *   - return value: 0 = PPAD_OK, otherwise a ppad_status; ppad_strerror() names it and
 *     ppad_last_error() holds the detail (thread-local);
 *   - per-board faults never abort a launch: they are reported in the flags byte of `info`;
 *   - there is NO CPU path: without a CUDA device every compute entry point returns PPAD_ERR_CUDA.
 *
 * Packed state record ("state"), one per board, ppad_state_bytes_ex(rows, cols, orb_bits) bytes:
 *   orb_bits = 3 (default: orb types 0..6, everything the default skyfall of parameters.py:43 produces)
 *     5x6 (<= 32 cells): uint32 b0, b1, b2, meta                    (16 B, one uint4)
 *     6x7 (<= 64 cells): uint64 b0, b1, b2; uint32 meta, reserved   (32 B, two uint4)
 *     codes 0..6 = red, blue, green, light, dark, heal, jammer (game.py:236-246), 7 = empty (-1);
 *   orb_bits = 4 (all ten orb types of game.py:236-246: + 7 poison, 8 mortal poison, 9 bomb)
 *     5x6: uint32 b0, b1, b2, b3, meta, 3 x reserved                (32 B)
 *     6x7: uint64 b0, b1, b2, b3; uint32 meta, 3 x reserved         (48 B)
 *     codes 0..9 = the reference's orb types, 15 = empty (-1);
 *   bit i of plane bj = bit j of the orb code of cell i = row * cols + col;
 *   meta = finger cell [0,6) | previous action [6,9) (4 = 'pass') | accepted moves this episode [9,25).
 */
#ifndef PPAD_B200_H
#define PPAD_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPAD_ABI_VERSION 4
#define PPAD_MAX_TEAM_UNITS 8

typedef enum ppad_status {
    PPAD_OK = 0,
    PPAD_ERR_INVALID = 1,       /* bad argument (message in ppad_last_error) */
    PPAD_ERR_UNSUPPORTED = 2,   /* board size / orb type / dtype without a kernel instantiation */
    PPAD_ERR_CUDA = 3           /* CUDA runtime error, including "no device" */
} ppad_status;

/* actions: player.py:24, simulation.py:236 */
enum { PPAD_UP = 0, PPAD_DOWN = 1, PPAD_LEFT = 2, PPAD_RIGHT = 3, PPAD_PASS = 4,
       PPAD_NOOP = 5,      /* the board sits this step out: no move, no evaluation, no flag (lock-step batches) */
       PPAD_SAMPLE = 255 };

/* flags byte (info >> 24) */
enum {
    PPAD_FLAG_REJECTED = 0x01,          /* move off the board: state unchanged (game.py:353 raises ValueError); from
                                           ppad_pack / ppad_reset: a given finger was off the board and was clamped */
    PPAD_FLAG_SKYFALL_EXHAUSTED = 0x02, /* injected orbs ran out, the Philox stream took over */
    PPAD_FLAG_CASCADE_CAP = 0x04,       /* cascade stopped at cfg.cascade_cap iterations */
    PPAD_FLAG_NO_LEGAL_MOVE = 0x08,     /* random policy found an empty legal set */
    PPAD_FLAG_BAD_ACTION = 0x10,        /* action id outside 0..5 */
    PPAD_FLAG_RESET_CAP = 0x20,         /* reset gave up after cfg.reset_cap attempts */
    PPAD_FLAG_UNSUPPORTED_ORB = 0x40,   /* pack saw a value outside -1..6 (orb_bits 3) / -1..9 (orb_bits 4): stored as empty */
    PPAD_FLAG_EPISODE_DONE = 0x80       /* a pass was processed and the board auto-reset */
};

/* info word: n_combos | depth << 8 | consumed << 16 | flags << 24, each field saturating at 255 */
#define PPAD_INFO_COMBOS(i) ((i) & 0xFFu)
#define PPAD_INFO_DEPTH(i) (((i) >> 8) & 0xFFu)
#define PPAD_INFO_CONSUMED(i) (((i) >> 16) & 0xFFu)
#define PPAD_INFO_FLAGS(i) (((i) >> 24) & 0xFFu)

typedef enum ppad_dtype { PPAD_I8 = 1, PPAD_I32 = 4, PPAD_I64 = 8 } ppad_dtype;
typedef enum ppad_obs_dtype { PPAD_OBS_F32 = 0, PPAD_OBS_U8 = 1 } ppad_obs_dtype;

/* Layout of an injected skyfall slab (the orb sequence the reference's random_orb, game.py:458-470, is replaced by in
 * parity runs): n_inj orbs per board in slab_words uint32 per board.
 *   PPAD_SLAB_NIBBLES  4 bits per orb, 8 per word: orb d in word d >> 3 at bits 4 * (d & 7)            (16 B / 32 orbs)
 *   PPAD_SLAB_PLANES3  bit-sliced 3-bit codes, 32 orbs per group of three words {p0, p1, p2}: bit (d & 31) of word
 *                      3 * (d >> 5) + j is bit j of orb d; slab_words is a multiple of 3                  (12 B / 32 orbs);
 *                      contexts with orb_bits = 3 only (orb types 0..6).
 *   PPAD_SLAB_CHUNKS10 chunk-major over the batch, ten 3-bit codes per uint32 chunk, bit-sliced inside the chunk: chunk c
 *                      of board i is slab[c * slab_stride + i], and bit 10 j + k of it is bit j of orb 10 c + k (bits
 *                      30, 31 unused).  slab_words = chunks per board, n_inj <= 10 * slab_words, slab_stride >= n
 *                      words.  Made for slabs in pinned HOST memory (ppad_pipe_step in zero-copy mode): the kernel
 *                      fetches the first slab_eager chunks of every board up front (one 128-byte line per warp and
 *                      chunk) and a further chunk only when a board's cascade gets that far, so the host may offer 50
 *                      orbs or more per board while 4 * slab_eager bytes per board cross PCIe (a random-policy step
 *                      draws 3.5 orbs per board on average; one board in a hundred draws more than 20).
 *                      ppad_step and ppad_pipe_step (zero copy) only; orb_bits = 3 only.
 *   PPAD_SLAB_CHUNKS6  the same chunk-major, eager / on-demand scheme with uint16 chunks of SIX orbs as base-6 digits
 *                      (orb 6 c + k of board i = (chunk / 6^k) % 6, chunk = ((const uint16_t *)slab)[c * slab_stride + i]):
 *                      the six skyfall colours 0..5 in 15.5 of 16 bits.  slab_words = chunks per board, n_inj <= 6 *
 *                      slab_words, slab_stride >= n in uint16 units, slab_eager 1..8 (default 3: 18 orbs, 6 bytes per
 *                      board).  A chunk value >= 46656 is no encoding: flagged PPAD_FLAG_UNSUPPORTED_ORB.
 *   PPAD_SLAB_CHUNKS6T the same uint16 chunks, chunk-major per TILE of 256 boards: chunk c of board i =
 *                      ((const uint16_t *)slab)[((i / 256) * slab_words + c) * 256 + i % 256]; the buffer holds
 *                      ceil(n / 256) * slab_words * 256 elements and is 16-byte aligned; slab_stride is ignored.  The
 *                      eager chunks of a CTA's 256 boards are one contiguous block: full 128-byte lines across PCIe.
 * A code that is no orb type of the context (>= 7, or >= 10 with orb_bits = 4) is flagged PPAD_FLAG_UNSUPPORTED_ORB. */
typedef enum ppad_slab_format {
    PPAD_SLAB_NIBBLES = 0, PPAD_SLAB_PLANES3 = 1, PPAD_SLAB_CHUNKS10 = 2, PPAD_SLAB_CHUNKS6 = 3, PPAD_SLAB_CHUNKS6T = 4
} ppad_slab_format;

/* Environment constants: the keyword arguments of PAD.__init__ (game.py:39-50) that reach the hot path. */
typedef struct ppad_config {
    int32_t rows, cols;              /* dim_row, dim_col: (4,5), (5,6) and (6,7) are instantiated */
    int32_t skyfall_damage;          /* game.py:45,387: cascade with drop + skyfall, or a single cancel */
    int32_t cascade_cap;             /* the reference has none; 0 = unlimited */
    int32_t reset_cap;               /* rejection-sampling attempts in reset; 0 = unlimited */
    int32_t team_size;               /* game.py:47, parameters.py:54-89 */
    int32_t team_color_1[PPAD_MAX_TEAM_UNITS];   /* orb type 0..4, or -1 for None */
    int32_t team_color_2[PPAD_MAX_TEAM_UNITS];
    double  team_atk[PPAD_MAX_TEAM_UNITS];
    double  skyfall[10];             /* game.py:42, parameters.py:43; types 7..9 need orb_bits = 4 */
    uint64_t seed;                   /* Philox4x32-10 key (replaces random.seed(datetime.now()), game.py:72,302) */
    int32_t orb_bits;                /* 0 or 3: 3-bit orb codes (types 0..6); 4: 4-bit codes (types 0..9) */
    int32_t reserved;
} ppad_config;

typedef struct ppad_ctx ppad_ctx;   /* immutable after creation; usable from any stream of its device */

/* parameters.py:43,54-89 + PAD() defaults */
void ppad_default_config(ppad_config *cfg);
/* PAD.__init__ (game.py:39-120) validation; `device` is a CUDA ordinal */
int ppad_create(const ppad_config *cfg, int device, ppad_ctx **out);
void ppad_destroy(ppad_ctx *ctx);
int ppad_abi_version(void);
const char *ppad_strerror(int status);
const char *ppad_last_error(void);
int64_t ppad_state_bytes(int rows, int cols);                       /* orb_bits = 3 */
int64_t ppad_state_bytes_ex(int rows, int cols, int orb_bits);     /* -1: no such geometry */
/* kernels launched by this library since load (all contexts, all threads) */
uint64_t ppad_launch_count(void);

/* PAD.board / PAD.finger / action_space.previous_action (game.py:88,94,117) <-> packed state.
 * boards [n, rows, cols] of `dtype`; fingers [n, 2] (row, col) of `dtype`; prev [n] uint8 or NULL
 * (= 'pass'); moves [n] uint16 or NULL (= 0); flags [n] uint8 or NULL. */
int ppad_pack(const ppad_ctx *ctx, int64_t n, const void *boards, const void *fingers, int dtype,
              const uint8_t *prev, const uint16_t *moves, void *state, uint8_t *flags, void *stream);
int ppad_unpack(const ppad_ctx *ctx, int64_t n, const void *state, void *boards, void *fingers, int dtype,
                uint8_t *prev, uint16_t *moves, void *stream);

/* PAD.reset (game.py:282-323) for n boards: combo-free random boards by rejection, random finger,
 * previous action 'pass'.  Orbs come from the injected nibble slab first (slab != NULL: n_inj orbs per
 * board, slab_words uint32 per board, orb d at word d >> 3, bits 4 * (d & 7)) and then from the Philox
 * stream (env0 + i, step).  fingers_in [n, 2] int32 or NULL (= Philox).  flags [n] uint8 or NULL. */
int ppad_reset(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, const uint32_t *slab,
               int32_t slab_words, int32_t n_inj, const int32_t *fingers_in, void *state, uint8_t *flags,
               uint8_t *consumed, void *stream);

/* One env step of n boards: the fused hot path.
 *   action      = actions ? actions[i] : (max_moves > 0 && moves >= max_moves ? PASS : random legal move)
 *                 (PlayerAction.sample, player.py:38-66; PPAD_SAMPLE in actions[i] also samples)
 *   move        : PAD.apply_action + PAD.swap (game.py:122-155,429-456); off-board -> FLAG_REJECTED
 *   evaluate    : on PASS, or on every accepted move when eval_moves != 0 (the look-ahead pattern
 *                 env.step(a); env.calculate_reward(...) of simulation.py:117-119):
 *                 PAD.calculate_reward (game.py:364-405) = cancel -> drop -> fill_board until stable,
 *                 on a COPY of the board, then PAD.damage (game.py:157-199) in float64
 *   auto_reset  : after a PASS, PAD.reset for that board (step-keyed Philox), FLAG_EPISODE_DONE
 *   obs         : Agent01.board_finger_to_state (agent01.py:254-273) of the state written back
 * Outputs may be NULL.  final_state receives the board AFTER the cascade (debug / parity only).
 * combo_log [n, log_cap] uint16 = colour | N << 8 per combo in reference order. */
typedef struct ppad_step_args {
    int64_t n;
    uint64_t env0;                /* global id of board 0 of this slab (sharding: rank * n_local) */
    uint32_t step;                /* global step index: Philox counter word */
    int32_t eval_moves, auto_reset, max_moves;
    void *state;                  /* in/out */
    const uint8_t *actions;       /* [n] or NULL */
    const uint32_t *slab;         /* injected skyfall, [n, slab_words] or NULL */
    int32_t slab_words, n_inj;
    uint8_t *action_out;          /* [n] action actually taken */
    double *reward;               /* [n] float64 */
    uint32_t *info;               /* [n] */
    void *final_state;            /* [n] packed, or NULL */
    void *state_copy;             /* [n] packed: a second copy of the state written back (e.g. pinned host memory) or NULL */
    uint16_t *combo_log;          /* [n, log_cap] or NULL */
    int32_t log_cap;
    void *obs;                    /* [n, rows, cols, 7] or NULL */
    int32_t obs_dtype;            /* ppad_obs_dtype */
    int32_t skyfall_damage;       /* -1 = cfg.skyfall_damage; 0 / 1 = the skyfall_damage argument of
                                     PAD.calculate_reward (game.py:364) for this call */
    /* optional whole finger paths instead of one action (BASELINE configs[2]): paths [n, path_words]
     * uint32, 16 moves per word, move t = (paths[i][t >> 4] >> 2 * (t & 15)) & 3; board i applies
     * path_lens ? path_lens[i] : path_len moves like that many PAD.step(move) calls (off-board moves are
     * skipped and flagged REJECTED), then is evaluated once when eval_moves != 0.  action_out = last
     * accepted move (255 = none). */
    const uint32_t *paths;
    const uint16_t *path_lens;
    int32_t path_words, path_len;
    int32_t slab_format;          /* ppad_slab_format of `slab` */
    int32_t slab_stride;          /* chunk-major slabs: elements (uint32 / uint16) between consecutive chunks of one board
                                     (>= n); else ignored */
    uint16_t *consumed_out;       /* [n] skyfall orbs the cascade drew, saturating at 65535 (the info byte stops at
                                     255), or NULL */
    int32_t slab_eager;           /* chunk-major slabs: chunks per board fetched up front; 0 = default (CHUNKS10: 2 of
                                     1..4, CHUNKS6: 3 of 1..8) */
    int32_t reserved;
} ppad_step_args;
int ppad_step(const ppad_ctx *ctx, const ppad_step_args *args, void *stream);

/* `steps` consecutive env steps of the in-kernel random policy in ONE launch (the loop of
 * visualize_episode.py:28-33 / model_simulation, simulation.py:34-68, for a random policy): boards stay in
 * registers between steps.  Equivalent to ppad_step(actions = NULL, step = step0 + t, eval_moves, auto_reset,
 * max_moves) for t = 0..steps-1 without injected skyfall or observation; per board it reports the sum of the
 * rewards (float64, added in step order), the number of finished episodes, the number of combos and the OR of
 * the flags bytes.  Outputs may be NULL. */
int ppad_rollout(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step0, int32_t steps, int32_t eval_moves,
                 int32_t auto_reset, int32_t max_moves, void *state, double *reward_sum, uint32_t *episodes,
                 uint32_t *combos_sum, uint8_t *flags_or, void *stream);
/* The same, recording the trajectory (the random walks of smart_data.py:69-72 in one launch): traj_state
 * [steps, n] packed records = the state after every step, traj_action [steps, n] uint8 = the action taken;
 * time-major so that a warp writes consecutive records.  Either may be NULL. */
int ppad_rollout_record(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step0, int32_t steps, int32_t eval_moves,
                        int32_t auto_reset, int32_t max_moves, void *state, double *reward_sum, uint32_t *episodes,
                        uint32_t *combos_sum, uint8_t *flags_or, void *traj_state, uint8_t *traj_action, void *stream);

/* The same env step for callers whose inputs and results live in HOST memory -- what the reference's callers
 * have: PAD.step takes a Python value and appends to Python lists (game.py:329-362).  A pipe owns device
 * scratch and `chunks` streams; ppad_pipe_step cuts the batch into chunks, and chunk k copies its inputs
 * host -> device, runs the fused step kernel and copies its results device -> host on its own stream, so the
 * copies of neighbouring chunks overlap the kernels and each other (PCIe is full duplex).
 * ppad_pipe_step returns after enqueueing; ppad_pipe_wait(slot) blocks until the step last enqueued with that
 * slot is complete; ppad_pipe_join makes `stream` wait for it instead.  Steps of one pipe execute in the order
 * they were enqueued (each chunk keeps its stream); `slot` (0 or 1) only names the completion events, so that a
 * caller with two sets of host buffers can enqueue step t + 1 before it waits for step t.  Host buffers should be pinned.  The env state and the
 * observation stay in HBM (state_out_host, if given, receives a copy of the packed state records: the
 * reference's (board, finger) observation). */
typedef struct ppad_pipe ppad_pipe;
typedef struct ppad_host_step_args {
    int64_t n;
    uint64_t env0;
    uint32_t step;
    int32_t eval_moves, auto_reset, max_moves;
    int32_t skyfall_damage;          /* -1 = cfg.skyfall_damage */
    void *state;                     /* DEVICE [n] packed records, in/out */
    const uint8_t *actions_host;     /* [n] or NULL (= in-kernel random policy) */
    const uint32_t *slab_host;       /* [n, slab_words] injected skyfall or NULL */
    int32_t slab_words, n_inj;
    uint8_t *action_out_host;        /* [n] or NULL */
    double *reward_host;             /* [n] or NULL */
    uint32_t *info_host;             /* [n] or NULL */
    void *state_out_host;            /* [n] packed records or NULL */
    void *obs;                       /* DEVICE [n, rows, cols, 7] or NULL */
    int32_t obs_dtype;
    int32_t slab_format;             /* ppad_slab_format of slab_host */
    /* Compact result stream (results_compact != 0; pinned host buffers, zero-copy mode): what a host caller of
     * PAD.step gets back (game.py:329-362: the action it took, the reward of an evaluation, a fresh board after a
     * reset) without shipping zeros.  About half the boards of a random-policy step match nothing (reward 0, info 0):
     * they cost 4 bits.  Per tile of 256 consecutive boards (tile t = boards 256 t ..):
     *   nibbles_host   [tiles, 32] uint32: nibble i of tile t (word i >> 3, bits 4 * (i & 7)) =
     *                  action taken (0..4, 7 = none) | has_record << 3 for board 256 t + i
     *   tile_line_host [tiles] uint32: index of the first 128-byte line of the tile's record block in records_host
     *                  (blocks are claimed with an atomic counter: their order is not the tile order)
     *   records_host   the record blocks.  A tile with k boards whose info word is not 0, m of them with
     *                  PPAD_FLAG_EPISODE_DONE, holds, in board order:
     *                    k x float64 reward | k x uint32 info | pad to 16 B | m x packed state record (the fresh
     *                    episode's board) | pad to 128 B
     *                  records_capacity bytes; the worst case is ppad_compact_capacity(ctx, n) bytes
     *   summary_host   [4] uint32: 128-byte lines used, tiles, overflow (a block did not fit: its tile_line is
     *                  0xFFFFFFFF), 0 -- valid once ppad_pipe_wait returns
     * reward_host / info_host / action_out_host are ignored in this mode; state_out_host still works.
     *
     * results_compact = 2: the same stream with the records dictionary-coded (ppad_pipe_set_dictionary first).  The
     * (reward, info) pairs of a step repeat -- every board with one red combo of three orbs at depth 1 reports the
     * same twelve bytes -- so a record that is in the pipe's dictionary travels as its 16-bit index.  The record block
     * of a tile with k records, e of them not in the dictionary, m fresh episodes:
     *   k x uint16 code (dictionary index; 0xFFFF = escape) | pad to 8 B | e x float64 reward | e x uint32 info
     *   (the escapes, in board order) | pad to 16 B | m x packed state record | pad to 128 B
     * Decoding is exact whatever the dictionary holds: it only decides how many records escape. */
    int32_t results_compact;
    uint32_t *nibbles_host;
    uint32_t *tile_line_host;
    void *records_host;
    int64_t records_capacity;
    uint32_t *summary_host;
    int32_t no_join;                 /* 0: `after_stream` is made to wait for this step (work queued there later sees the
                                        state it leaves behind); 1: the caller orders later work itself (ppad_pipe_join /
                                        ppad_pipe_wait) -- saves one cross-stream dependency per step when steps are
                                        enqueued back to back */
    int32_t slab_stride;             /* chunk-major slabs: elements between consecutive chunks of one board (>= n) */
    int32_t slab_eager;              /* chunk-major slabs: chunks per board fetched up front; 0 = default */
    int32_t reserved2;
} ppad_host_step_args;
/* bytes of records_host that can never overflow for n boards (with or without auto_reset, with or without a dictionary) */
int64_t ppad_compact_capacity(const ppad_ctx *ctx, int64_t n);
/* The dictionary of results_compact = 2 -- a transport coding of what PAD.step appends to env.rewards (game.py:357-362)
 * and of the counters calculate_reward keeps (game.py:364-405); nothing of the game is decided by it: `count`
 * (<= PPAD_DICT_MAX) pairs (rewards[i], infos[i]), typically the most
 * frequent pairs of an earlier step's results (the host learns them; the library never guesses).  Code i of the stream
 * means pair i of THIS call's arrays: keep them for decoding.  A pair the hash table cannot place (more than 8
 * collisions in a row at load factor 1/4: practically never) simply keeps escaping.  Waits for the steps already
 * enqueued on the pipe; count = 0 removes the dictionary. */
#define PPAD_DICT_MAX 32768
int ppad_pipe_set_dictionary(ppad_pipe *pipe, const double *rewards, const uint32_t *infos, int32_t count);
int ppad_pipe_create(const ppad_ctx *ctx, int64_t n_max, int32_t slab_words_max, int32_t chunks, ppad_pipe **out);
void ppad_pipe_destroy(ppad_pipe *pipe);
/* `after_stream`: work already queued there (e.g. ppad_reset) is ordered before the chunks */
int ppad_pipe_step(ppad_pipe *pipe, const ppad_host_step_args *args, void *after_stream, int32_t slot);
/* enabled = 1: pinned host buffers are handed to the kernel directly (zero copy over unified addressing: the PCIe
 * traffic overlaps the kernel store by store); 2: only the results are written in place, the skyfall slab and the
 * actions travel by DMA chunk by chunk (no PCIe read inside a kernel); 0, or pageable buffers: everything goes
 * through the pipe's staging buffers and chunked cudaMemcpyAsync. */
int ppad_pipe_set_zero_copy(ppad_pipe *pipe, int32_t enabled);
int ppad_pipe_wait(ppad_pipe *pipe, int32_t slot);
int ppad_pipe_join(ppad_pipe *pipe, void *stream, int32_t slot);

/* Agent01.board_finger_to_state (agent01.py:254-273): state -> [n, rows, cols, 7] NHWC */
int ppad_encode_obs(const ppad_ctx *ctx, int64_t n, const void *state, void *obs, int obs_dtype, void *stream);

/* pad_utils.cancel (utils.py:26-70): one in-place cancel; n_combos [n] uint8; combo_log as above */
int ppad_cancel(const ppad_ctx *ctx, int64_t n, void *state, uint8_t *n_combos, uint16_t *combo_log,
                int32_t log_cap, void *stream);

/* pad_utils.detect_island (utils.py:86-111): island_out[i] = island_in[i] | the 4-connected island of cells of
 * orb_type[i] (-1: empty cells; anything else that no cell holds: nothing) around cell seed_cell[i] = row * cols + col of board i, as bit masks (bit = cell index).
 * Like the reference's recursion the flood does not enter cells already set in island_in (NULL = all clear); a seed
 * that does not hold orb_type leaves the mask as it was. */
int ppad_island(const ppad_ctx *ctx, int64_t n, const void *state, const uint8_t *seed_cell, const int8_t *orb_type,
                const uint64_t *island_in, uint64_t *island_out, void *stream);

/* pad_utils.illegal_actions (utils.py:73-83) / the support of PlayerAction.sample (player.py:46-62):
 * mask [n] uint8, bit a = action a is a legal non-reversing move; bits 4..7 = off-board moves. */
int ppad_legal_mask(const ppad_ctx *ctx, int64_t n, const void *state, uint8_t *mask, void *stream);

/* PlayerAction.sample (player.py:38-66): one uniformly random legal non-reversing move per board
 * (255 when there is none), from the same Philox draw the in-kernel policy of ppad_step uses, so
 * ppad_sample_actions(step) followed by ppad_step(actions, step) equals ppad_step(NULL, step). */
int ppad_sample_actions(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, const void *state,
                        int32_t include_pass, uint8_t *actions, void *stream);

/* PAD.drop (game.py:201-221) in place */
int ppad_drop(const ppad_ctx *ctx, int64_t n, void *state, void *stream);

/* PAD.fill_board (game.py:232-256) in place: empties (or every cell when reset_all != 0) in row-major
 * order, one orb each, from draw index draw0 of the (slab, Philox skyfall stream) source. */
int ppad_fill(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, int32_t draw0, int32_t reset_all,
              const uint32_t *slab, int32_t slab_words, int32_t n_inj, void *state, uint8_t *flags, uint8_t *consumed,
              void *stream);

/* PAD.damage (game.py:157-199) of explicit combo lists: combo_log [n, log_cap] = colour | N << 8 in
 * order, n_combos [n] -> reward [n] float64 */
int ppad_damage(const ppad_ctx *ctx, int64_t n, const uint16_t *combo_log, int32_t log_cap, const uint8_t *n_combos,
                double *reward, void *stream);

/* ---- the callers either side of the env step ---------------------------- */

typedef enum ppad_select_method { PPAD_SELECT_MAX = 0, PPAD_SELECT_BOLTZMANN = 1 } ppad_select_method;

/* permutation_mapping (smart_data.py:94-108) in place: maps [n, 8] uint8, orb type t -> maps[i][t] (t = 0..6;
 * empty cells stay empty; byte 7 is padding; with orb_bits = 4 the types 7..9 keep their value -- the reference
 * only ever permutes allowed_orbs = 0..5, solved_boards.py:25) */
int ppad_permute(const ppad_ctx *ctx, int64_t n, void *state, const uint8_t *maps, void *stream);


/* Agent01.act after the network (agent01.py:173-218): q [n, 5] float32 -> actions [n].
 * Off-board moves and the reverse of the previous action get -inf (agent01.py:173-190); epsilon-greedy
 * (agent01.py:193-201), then argmax or a Boltzmann draw exp(beta q) / sum (agent01.py:203-213) from the
 * Philox POLICY stream of (env0 + i, step).  max_moves > 0 forces PASS at the episode length cap
 * (simulation.py:50-52). */
int ppad_select_actions(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, const void *state, const float *q,
                        int32_t method, float beta, float epsilon, int32_t max_moves, uint8_t *actions, void *stream);

/* The policy-driven env step of a device-resident rollout as ONE kernel (the loop body of model_simulation,
 * projects/simulation.py:34-68, for n boards; BASELINE configs[4]):
 *   action  = max_moves > 0 && moves >= max_moves ? PASS (simulation.py:50-52)
 *             : Agent01.act after the network (agent01.py:173-218) on q[i], as ppad_select_actions
 *   ring    : row (ring_step % ring_steps) * n + i of the time-major replay ring gets the state BEFORE the step and
 *             the action (the (observation, action) pairs of simulation.py:60-61); ring_state == NULL: no replay
 *   step    : as ppad_step(actions = action, eval_moves, auto_reset, max_moves) with Philox (or injected) skyfall
 *   reward  : the ring row gets the step's reward (log10 of it when log10_reward and it is positive); when the
 *             board passed with a positive reward and gamma > 0, the rows of the episode's earlier steps still in
 *             the ring get the discounted returns cur = cur * gamma + 0 of ppad.discount (agent/utils.py:22-51)
 *   obs     : one-hot observation of the state written back (agent01.py:254-273), or NULL
 *   stats   : stats_rows [ppad_policy_stats_rows(n), 40] uint64 per-CTA accumulators, ADDED to by every call (zero
 *             them to start a window): board-steps, evaluated boards with a match, sum of rewards * 1024 (fixed
 *             point: exact and order independent), maximum reward (float64 bits), combos, skyfall orbs consumed,
 *             episodes finished, faults, cascade-depth histogram [16] and combo-count histogram [16] over the
 *             evaluated boards (the report block of simulation.py:336-349).  ppad_policy_stats sums the rows into
 *             out [40] (the vector a rank contributes to the NCCL all-gather). */
typedef struct ppad_policy_args {
    int64_t n;
    uint64_t env0;
    uint32_t step;
    int32_t eval_moves, auto_reset, max_moves;
    void *state;                  /* in/out */
    const float *q;               /* [n, 5] float32 */
    int32_t method;               /* ppad_select_method */
    float beta, epsilon;
    const uint32_t *slab;         /* injected skyfall [n, slab_words] or NULL */
    int32_t slab_words, n_inj, slab_format;
    int32_t obs_dtype;
    uint8_t *action_out;          /* [n] or NULL */
    double *reward;               /* [n] or NULL */
    uint32_t *info;               /* [n] or NULL */
    void *obs;                    /* [n, rows, cols, 7] or NULL */
    void *ring_state;             /* [ring_steps * n] packed records or NULL */
    uint8_t *ring_action;         /* [ring_steps * n] */
    double *ring_reward;          /* [ring_steps * n] */
    int64_t ring_steps, ring_step;   /* ring_step = rows pushed so far / n (0 for the first step) */
    double gamma;                 /* 0 = store plain rewards */
    int32_t log10_reward;
    int32_t ctas_per_sm;          /* 0 = default; register allocation target of the kernel (3 or 4), for tuning */
    uint64_t *stats_rows;         /* or NULL */
} ppad_policy_args;
int ppad_policy_step(const ppad_ctx *ctx, const ppad_policy_args *args, void *stream);
int64_t ppad_policy_stats_rows(int64_t n);
int ppad_policy_stats(const ppad_ctx *ctx, int64_t n, const uint64_t *stats_rows, uint64_t *out, void *stream);

/* ppad.discount (agent/utils.py:22-51) in place over n trajectories rewards [n, max_len] float64 with
 * lengths [n] int32 (NULL = max_len): optional log10 of positive rewards, reverse scan
 * cur = cur * gamma + r, optional normalisation. */
int ppad_discount(const ppad_ctx *ctx, int64_t n, int32_t max_len, double *rewards, const int32_t *lengths, double gamma,
                  int32_t log10_flag, int32_t norm, void *stream);

/* Experience replay ring (simulation.py:294-306) of (packed state, action, reward) rows.
 * push: row i (if keep == NULL or keep[i]) goes to ring slot (cursor + (slot ? slot[i] : i)) % capacity.
 * `count` = number of rows this push keeps (n when keep == NULL).  When count > capacity only the LAST `capacity`
 * kept rows are written (rows with slot index < count - capacity are skipped), so no two rows of one push share a
 * ring slot -- the reference truncates to MAX_DATA_SIZE and never mixes tuples (simulation.py:296-299).
 * sample: k uniform rows of the first `size` ring rows (Philox, keyed by step); index (may be NULL) gets the rows. */
int ppad_replay_push(const ppad_ctx *ctx, int64_t n, int64_t capacity, int64_t cursor, int64_t count, const void *state,
                     const uint8_t *action, const double *reward, const uint8_t *keep, const int64_t *slot,
                     void *r_state, uint8_t *r_action, double *r_reward, void *stream);
int ppad_replay_sample(const ppad_ctx *ctx, int64_t k, int64_t size, uint32_t step, const void *r_state,
                       const uint8_t *r_action, const double *r_reward, void *state, uint8_t *action, double *reward,
                       int64_t *index, void *stream);

#ifdef __cplusplus
}
#endif
#endif
