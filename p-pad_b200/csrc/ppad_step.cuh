// ppad_step.cuh -- the env-step kernels (one thread per board with the fused observation; persistent CTAs with a
// recirculating cascade pool for the state-only step; T fused steps of the random policy) and their per-geometry
// launchers.  Instantiated once per geometry by ppad_step_inst.cu (-DPPAD_GEOM=0..5), so the geometries compile in
// parallel; ppad_step.cu holds the C ABI on top of the launchers.
#pragma once
#include "ppad_device.cuh"

namespace ppad {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
struct StepKernelArgs {
    StepParams p;
    int64_t n;
    uint64_t env0;
    void *state;
    const uint8_t *actions;
    const uint32_t *slab;
    uint8_t *action_out;
    double *reward;
    uint32_t *info;
    void *final_state;
    void *state_copy;
    uint16_t *combo_log;
    uint16_t *consumed16;   // [n] skyfall orbs drawn by the cascade, saturating at 65535 (the info byte stops at 255), or NULL
    void *obs;
    int32_t obs_dtype;
    const uint32_t *paths;
    const uint16_t *path_lens;
    int32_t path_words, path_len;
    int32_t stage_slab;   // the slab is pinned HOST memory (zero copy): fetch each row once, up front.  1: 4 words per
                          // board, one uint4 per thread; 2: 3 words per board, the CTA's 3072 contiguous bytes cooperatively;
                          // 3: SLAB_CHUNKS10 (host or device memory): the first p.slab_eager chunks up front, further
                          // chunks when the cascade asks; 4: SLAB_CHUNKS6T: the same, the CTA's eager block cooperatively
    uint32_t *tile_counter;   // DEVICE [2]: {next tile, CTAs finished} of the warp-pool kernel; both 0 before the launch and
                              // left at 0 by the last CTA to finish (no memset between launches)
    // compact result stream (COMPACT instantiations: the host pipe), see include/ppad_b200.h ppad_host_step_args
    uint32_t *c_nibbles;      // [tiles, 32] words: nibble i of tile t = action (7 = none) | has_record << 3 of board 256 t + i
    uint32_t *c_tile_line;    // [tiles]: first 128-byte line of the tile's record block
    uint4 *c_records;         // the line stream
    uint32_t *c_summary;      // [4]: lines used, tiles, overflow, 0 -- written by the last CTA to finish
    uint32_t *c_counters;     // DEVICE [2]: next free line, finished tiles; left at 0 by the last CTA
    uint32_t c_capacity;      // lines the record buffer holds
    const uint4 *c_dict;      // DEVICE: result dictionary (open-addressed hash table, see ResultDict) or NULL
    uint32_t c_dict_mask;     // slots - 1
};

// Result dictionary of the compact stream (ppad_pipe_set_dictionary): the (reward, info) pairs of a step repeat -- a
// board with one red combo of three orbs at depth 1 always reports the same twelve bytes -- so the host hands the pipe
// the pairs it has seen most often, and a board whose pair is in the dictionary sends its 16-bit index instead.
// Open addressing, linear probing, at most DICT_PROBES slots per lookup (the builder drops what does not fit); slot =
// {reward bits lo, hi, info, index}; index 0xFFFFFFFF marks an empty slot.  The reward bit pattern 0x7FF8DEADBEEF0000
// never occurs (a NaN), so an all-ones slot is unambiguous.
constexpr int DICT_PROBES = 8;
constexpr uint32_t DICT_ESCAPE = 0xFFFFu;
__host__ __device__ inline uint32_t dict_hash(uint32_t lo, uint32_t hi, uint32_t info)
{
    uint32_t h = lo * 0x9E3779B1u;
    h = (h ^ (h >> 15) ^ hi) * 0x85EBCA77u;
    h = (h ^ (h >> 13) ^ info) * 0xC2B2AE3Du;
    return h ^ (h >> 16);
}
// The first probe is issued by the caller as soon as the step's results are known (dict_first), so that its trip to L2
// runs under the observation stores; dict_lookup finishes the search after them.
__device__ __forceinline__ uint4 dict_first(const uint4 *dict, uint32_t mask, double reward, uint32_t info)
{
    return __ldg(dict + (dict_hash((uint32_t)__double2loint(reward), (uint32_t)__double2hiint(reward), info) & mask));
}
__device__ __forceinline__ uint32_t dict_lookup(const uint4 *dict, uint32_t mask, double reward, uint32_t info, uint4 slot)
{
    const uint32_t lo = (uint32_t)__double2loint(reward), hi = (uint32_t)__double2hiint(reward);
    uint32_t h = dict_hash(lo, hi, info);
#pragma unroll 1
    for (int probe = 0; probe < DICT_PROBES; probe++) {
        if (slot.w == 0xFFFFFFFFu) break;
        if (slot.x == lo && slot.y == hi && slot.z == info) return slot.w;
        h++;
        slot = __ldg(dict + (h & mask));
    }
    return DICT_ESCAPE;
}

// Record block of one tile (= CTA of BLOCK boards) in the compact result stream: k records for the boards whose info
// word is not 0 (a board that was not evaluated, or matched nothing, has reward 0 and info 0), m <= k of them with a
// fresh episode (FLAG_EPISODE_DONE), in board order:
//   [k x float64 reward][k x uint32 info][pad to 16 bytes][m x packed state record], padded to 128 bytes.
// With a result dictionary (e <= k of the records are not in it):
//   [k x uint16 code, 0xFFFF = escape][pad to 8][e x float64 reward][e x uint32 info][pad to 16][m x state], padded to 128.
template <class G>
struct CompactLayout {
    static constexpr int SB = G::STATE_BYTES;
    static __host__ __device__ constexpr int off_info(int k) { return 8 * k; }
    static __host__ __device__ constexpr int off_state(int k) { return (12 * k + 15) & ~15; }
    static __host__ __device__ constexpr int lines(int k, int m) { return (off_state(k) + m * SB + 127) / 128; }
    static __host__ __device__ constexpr int d_off_reward(int k) { return (2 * k + 7) & ~7; }
    static __host__ __device__ constexpr int d_off_info(int k, int e) { return d_off_reward(k) + 8 * e; }
    static __host__ __device__ constexpr int d_off_state(int k, int e) { return (d_off_info(k, e) + 4 * e + 15) & ~15; }
    static __host__ __device__ constexpr int d_lines(int k, int e, int m) { return (d_off_state(k, e) + m * SB + 127) / 128; }
    static constexpr int MAX_STAGE = 128 * ((2 * BLOCK + 8 * BLOCK + 4 * BLOCK + BLOCK * SB + 127) / 128);
    static constexpr int MAX_STAGE_NO_RESET = 128 * ((2 * BLOCK + 8 * BLOCK + 4 * BLOCK + 127) / 128);
};

// Every thread of the CTA must call this (after its board's step is complete).
template <class G>
__device__ __forceinline__ void compact_results(const StepKernelArgs &a, bool valid, const StepOut &o, const State<G> &s,
                                                const uint4 &first_slot)
{
    typedef CompactLayout<G> CL;
    extern __shared__ uint4 c_stage4[];
    unsigned char *stage = reinterpret_cast<unsigned char *>(c_stage4);
    __shared__ int c_cnt[3][BLOCK / 32];
    __shared__ __align__(16) uint32_t c_nib[BLOCK / 8];
    __shared__ uint32_t c_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool has = valid && o.info != 0;             // reward != 0 implies n_combos != 0
    const bool done = has && (o.info >> 31) != 0;      // FLAG_EPISODE_DONE: the host wants the fresh board
    const bool use_dict = a.c_dict != nullptr;          // kernel-uniform
    uint32_t code = DICT_ESCAPE;
    if (use_dict && has) code = dict_lookup(a.c_dict, a.c_dict_mask, o.reward, o.info, first_slot);
    const bool esc = has && code == DICT_ESCAPE;
    const unsigned bh = __ballot_sync(0xFFFFFFFFu, has), bd = __ballot_sync(0xFFFFFFFFu, done),
                   be = __ballot_sync(0xFFFFFFFFu, esc);
    if (lane == 0) {
        c_cnt[0][warp] = __popc(bh);
        c_cnt[1][warp] = __popc(bd);
        c_cnt[2][warp] = __popc(be);
    }
    uint32_t nib = valid ? ((o.action <= ACT_NOOP ? (uint32_t)o.action : 7u) | (has ? 8u : 0u)) : 0u;   // none / bad id -> 7
    nib <<= 4 * (lane & 7);
    nib |= __shfl_xor_sync(0xFFFFFFFFu, nib, 1);
    nib |= __shfl_xor_sync(0xFFFFFFFFu, nib, 2);
    nib |= __shfl_xor_sync(0xFFFFFFFFu, nib, 4);
    if ((lane & 7) == 0) c_nib[warp * 4 + (lane >> 3)] = nib;
    __syncthreads();
    const unsigned below = (1u << lane) - 1u;
    int r = __popc(bh & below), rd = __popc(bd & below), re = __popc(be & below), k = 0, m = 0, e = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; w++) {
        const int ch = c_cnt[0][w], cd = c_cnt[1][w], ce = c_cnt[2][w];
        if (w < warp) {
            r += ch;
            rd += cd;
            re += ce;
        }
        k += ch;
        m += cd;
        e += ce;
    }
    const int off_reward = use_dict ? CL::d_off_reward(k) : 0;
    const int off_info = use_dict ? CL::d_off_info(k, e) : CL::off_info(k);
    const int off_state = use_dict ? CL::d_off_state(k, e) : CL::off_state(k);
    const int lines = use_dict ? CL::d_lines(k, e, m) : CL::lines(k, m);
    if (tid == 0) {
        uint32_t base = 0;
        if (lines) {
            base = atomicAdd(&a.c_counters[0], (uint32_t)lines);
            if (base + (uint32_t)lines > a.c_capacity) base = 0xFFFFFFFFu;   // overflow: block dropped, flagged below
        }
        c_base = base;
        a.c_tile_line[blockIdx.x] = base;
    }
    if (use_dict) {
        if (has) *reinterpret_cast<uint16_t *>(stage + 2 * r) = (uint16_t)code;
        if (esc) {
            *reinterpret_cast<double *>(stage + off_reward + 8 * re) = o.reward;
            *reinterpret_cast<uint32_t *>(stage + off_info + 4 * re) = o.info;
        }
    } else if (has) {
        *reinterpret_cast<double *>(stage + 8 * r) = o.reward;
        *reinterpret_cast<uint32_t *>(stage + off_info + 4 * r) = o.info;
    }
    if (done) StateIO<G>::store(stage + off_state, rd, s.b, s.meta);
    __syncthreads();
    const uint32_t base = c_base;
    if (base != 0xFFFFFFFFu)
        for (int j = tid; j < lines * 8; j += BLOCK) a.c_records[(size_t)base * 8 + j] = c_stage4[j];
    if (tid < BLOCK / 32) reinterpret_cast<uint4 *>(a.c_nibbles)[(size_t)blockIdx.x * (BLOCK / 32) + tid] = reinterpret_cast<const uint4 *>(c_nib)[tid];
    if (tid == 0) {
        if (base == 0xFFFFFFFFu) a.c_summary[2] = 1u;
        __threadfence();
        const uint32_t t = atomicAdd(&a.c_counters[1], 1u);
        if (t == gridDim.x - 1) {                      // every other CTA claimed its lines before it counted itself
            const uint32_t used = atomicExch(&a.c_counters[0], 0u);
            a.c_counters[1] = 0u;
            a.c_summary[0] = used < a.c_capacity ? used : a.c_capacity;
            a.c_summary[1] = gridDim.x;
        }
    }
}


// The step WITH the one-hot observation, one thread per board.  This variant is bound by the 840 B/board store
// stream, and keeping all eight warps of a CTA busy (each runs its own boards' cascades in a warp-uniform
// loop, then emits its observation) overlaps the stores with the integer work best: measured 0.159 ms per
// 2^20-board step against 0.162-0.166 ms for packing the cascades once and 0.19 ms for the pooled kernel.
// Four resident CTAs per SM (64 registers) is the measured optimum with fp32 obs: 0.1546 ms against 0.1556 (3),
// 0.1875 (2), 0.1685 (5, spills) and 0.1731 (6).  Emitting the observation before the cascades (0.167 ms), or
// before them in the odd warps and after them in the even warps (0.156 ms), is not faster either.
// (6x7 with 4-bit codes -- four 64-bit planes -- is the one geometry that does better with 3: 0.221 against 0.240 ms.)
// AR: the launch has auto_reset set.  Launches without it run an instantiation that does not contain the reset code
// at all; those with it get the unrolled Philox refill (a Boltzmann-policy rollout resets 29 % of its boards per step:
// 0.227 ms per 2^20-board step with fp32 obs against 0.286 ms with the plain refill, profiles/ar_bench.py).
// Resident CTAs per SM the register allocation aims at: 4 (64 registers) for the plain step; the instantiations that
// carry the auto-reset need more (154 B of spill stores per thread at 64 registers for 5x6) and get 3 (80 registers: no
// spills), like the widest geometry (6x7 with four 64-bit planes).
#if !defined(PPAD_MINB_64)
#define PPAD_MINB_64 4   // 64-bit planes, 3-bit codes (6x7) without auto-reset
#endif
template <class G, bool AR>
constexpr int step_min_ctas()
{
    return (AR || (sizeof(typename G::W) == 8 && G::NP == 4)) ? 3 : (sizeof(typename G::W) == 8 ? PPAD_MINB_64 : 4);
}

// COMPACT: the results leave as the compact stream of compact_results() instead of dense arrays (the host pipe).
template <class G, bool OBS, bool AR, bool COMPACT = false>
__global__ void __launch_bounds__(BLOCK, step_min_ctas<G, AR>()) step_obs_kernel(const __grid_constant__ StepKernelArgs a)
{
    geom_enter<G>(a.p.dyn);
    __shared__ ObsShared obs_sh;
    __shared__ uint32_t obs_stream[OBS && !G::DYNAMIC ? (BLOCK / 32) * ObsGeom<G>::WARP_WORDS : 1];
    if (OBS) {
        obs_init_luts(obs_sh);
        __syncthreads();
    }
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = i < a.n;
    State<G> s;
    clear_state<G>(s);
    int action = -1, path_len = 0;
    if (valid) {
        s = StateIO<G>::load(a.state, i);
        if (a.actions) {
            action = a.actions[i];
            if (action == PPAD_SAMPLE) action = -1;
        }
        if (a.paths) {
            path_len = a.path_lens ? a.path_lens[i] : a.path_len;
            path_len = path_len > 16 * a.path_words ? 16 * a.path_words : path_len;
        }
    }
    const uint32_t *slab = a.slab && valid ? a.slab + (size_t)i * a.p.slab_words : nullptr;
    const uint32_t *slab_eager = nullptr;
    __shared__ uint4 slab_rows[BLOCK];
    if (a.stage_slab == 3) {   // kernel-uniform
        // chunk-major slab: a chunk of the warp's 32 boards is one 128-byte line; the eager chunks are in flight together
        // with the state load and wait in the thread's own 16 bytes of shared memory
        slab = a.slab && valid ? a.slab + i : nullptr;
        if (slab) {
            uint32_t w[4] = {0, 0, 0, 0};
            if (a.p.slab_fmt == SLAB_CHUNKS6) {       // uint16 chunks: up to 8 eager ones fill the same 16 bytes
                const uint16_t *s16 = reinterpret_cast<const uint16_t *>(a.slab) + i;
                slab = reinterpret_cast<const uint32_t *>(s16);
#pragma unroll
                for (int c = 0; c < 8; c++)
                    if (c < a.p.slab_eager) w[c >> 1] |= (uint32_t)__ldg(s16 + (size_t)c * (size_t)a.p.slab_stride) << (16 * (c & 1));
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++)
                    if (c < a.p.slab_eager) w[c] = __ldg(slab + (size_t)c * (size_t)a.p.slab_stride);
            }
            slab_rows[threadIdx.x] = make_uint4(w[0], w[1], w[2], w[3]);
            slab_eager = reinterpret_cast<const uint32_t *>(&slab_rows[threadIdx.x]);
        }
    } else if (a.stage_slab == 4) {   // kernel-uniform
        // tile-major uint16 chunks: the eager chunks of the CTA's 256 boards are 512 * slab_eager contiguous bytes
        const uint16_t *tile = reinterpret_cast<const uint16_t *>(a.slab) + (size_t)blockIdx.x * (size_t)a.p.slab_words * BLOCK;
        if ((int)threadIdx.x < 32 * a.p.slab_eager) slab_rows[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(tile) + threadIdx.x);
        __syncthreads();
        if (valid) {
            slab = reinterpret_cast<const uint32_t *>(tile + threadIdx.x);
            slab_eager = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint16_t *>(slab_rows) + threadIdx.x);
        } else {
            slab = nullptr;
        }
    } else if (a.stage_slab == 1) {
        // Zero copy: the skyfall slab lives in pinned host memory.  The cascade reads its words lazily, which over
        // PCIe is one 4-byte round trip per lane and cascade iteration; fetch the board's whole 16-byte row once
        // instead (512 contiguous bytes per warp, in flight together with the state load) and let the cascade read
        // it from shared memory.
        if (slab) {
            slab_rows[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(slab));
            slab = reinterpret_cast<const uint32_t *>(&slab_rows[threadIdx.x]);
        }
        __syncwarp();
    } else if (a.stage_slab == 2) {
        // The same for 12-byte rows (SLAB_PLANES3, one group): the CTA's 256 rows are 3072 contiguous, 16-byte aligned
        // bytes; 192 threads fetch one uint4 each, then every thread reads its three words from shared memory
        // (stride 3 words: conflict free).
        const int64_t tile0 = (int64_t)blockIdx.x * BLOCK;
        const int rows = (int)(a.n - tile0 < BLOCK ? a.n - tile0 : BLOCK);
        const uint32_t *src = a.slab + (size_t)tile0 * 3;
        const int w0 = 4 * (int)threadIdx.x;
        if (w0 + 3 < 3 * rows) slab_rows[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(src) + threadIdx.x);
        else
            for (int w = w0; w < 3 * rows; w++) reinterpret_cast<uint32_t *>(slab_rows)[w] = __ldg(src + w);
        __syncthreads();
        if (slab) slab = reinterpret_cast<const uint32_t *>(slab_rows) + 3 * threadIdx.x;
    }
    // every lane takes part (the cascade loop inside is warp-uniform); lanes past the batch are not live
    Board<G> fin;
    const StepOut o = step_board<G, AR, AR>(valid, s, fin, a.p, a.env0 + (uint64_t)i, action, slab,
                                    a.combo_log && valid ? a.combo_log + (size_t)i * a.p.log_cap : nullptr,
                                    a.paths && valid ? a.paths + (size_t)i * a.path_words : nullptr, path_len, 0, nullptr, slab_eager);
    if (valid) {
        StateIO<G>::store(a.state, i, s.b, s.meta);
        if constexpr (!COMPACT) {
            if (a.action_out) a.action_out[i] = (uint8_t)o.action;
            if (a.reward) a.reward[i] = o.reward;
            if (a.info) a.info[i] = o.info;
        }
        if (a.final_state) StateIO<G>::store(a.final_state, i, fin, s.meta);
        if (a.state_copy) StateIO<G>::store(a.state_copy, i, s.b, s.meta);
        if (a.consumed16) a.consumed16[i] = (uint16_t)(o.consumed > 0xFFFF ? 0xFFFF : o.consumed);
    }
    uint4 first_slot = make_uint4(0, 0, 0, 0xFFFFFFFFu);
    if constexpr (COMPACT) {
        if (a.c_dict && valid && o.info != 0) first_slot = dict_first(a.c_dict, a.c_dict_mask, o.reward, o.info);
    }
    const int64_t warp0 = i - (threadIdx.x & 31);
    if (OBS && warp0 < a.n) {
        const int n_valid = (int)(a.n - warp0 < 32 ? a.n - warp0 : 32);
        emit_obs_any<G>(s, valid, n_valid, obs_stream + (threadIdx.x >> 5) * ObsGeom<G>::WARP_WORDS, obs_sh,
                        obs_warp_ptr<G>(a.obs, warp0, a.obs_dtype), a.obs_dtype);
    }
    if constexpr (COMPACT) compact_results<G>(a, valid, o, s, first_slot);
}

// T consecutive steps of the in-kernel random policy in one launch: the board stays in registers between
// steps, only the final state and per-board totals are written.  Equals T calls of ppad_step(actions = NULL,
// step = step0 + t) without injected skyfall / observation (tests/test_gpu_parity.py::test_fused_rollout_equals_single_steps).
struct RolloutArgs {
    StepParams p;
    int64_t n;
    uint64_t env0;
    int32_t steps;
    void *state;
    double *reward_sum;
    uint32_t *episodes;
    uint32_t *combos_sum;
    uint8_t *flags_or;
    void *traj_state;       // [steps, n] packed records after every step (time-major: coalesced), or NULL
    uint8_t *traj_action;   // [steps, n] action taken at every step, or NULL
};

template <class G, bool RECORD>   // RECORD: the trajectory outputs are compiled in only where they are asked for
__global__ void __launch_bounds__(BLOCK) rollout_kernel(const __grid_constant__ RolloutArgs a)
{
    geom_enter<G>(a.p.dyn);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = i < a.n;
    State<G> s;
    clear_state<G>(s);
    if (valid) s = StateIO<G>::load(a.state, i);
    double reward_sum = 0.0;
    uint32_t episodes = 0, combos = 0, flags = 0;
    U4 act_blk = {0, 0, 0, 0};
    for (int t = 0; t < a.steps; t++) {          // every lane runs every step: the cascade loop is warp-uniform
        Board<G> fin;
        const uint32_t step = a.p.key.step + (uint32_t)t;
        if (t == 0 || (step & 3u) == 0) {        // one Philox block holds the action draws of four steps
            RngKey key = board_key(a.p, a.env0 + (uint64_t)i);
            key.step = step;
            act_blk = action_block(key);
        }
        const StepOut o = step_board<G, true>(valid, s, fin, a.p, a.env0 + (uint64_t)i, -1, nullptr, nullptr, nullptr,
                                              0, (uint32_t)t, &act_blk);
        if constexpr (RECORD) {
            if (valid && a.traj_state) StateIO<G>::store(a.traj_state, (int64_t)t * a.n + i, s.b, s.meta);
            if (valid && a.traj_action) a.traj_action[(int64_t)t * a.n + i] = (uint8_t)o.action;
        }
        reward_sum = ppad::dadd(reward_sum, o.reward);
        episodes += (o.info >> 31) & 1u;         // FLAG_EPISODE_DONE
        combos += o.info & 0xFFu;
        flags |= o.info >> 24;
    }
    if (valid) {
        StateIO<G>::store(a.state, i, s.b, s.meta);
        if (a.reward_sum) a.reward_sum[i] = reward_sum;
        if (a.episodes) a.episodes[i] = episodes;
        if (a.combos_sum) a.combos_sum[i] = combos;
        if (a.flags_or) a.flags_or[i] = (uint8_t)flags;
    }
}

// ------------------------------------------------------------------------------------------------
// Persistent step kernel with a recirculating cascade pool.
//
// Cascade depth is geometric-like (about 45 % of the boards: 0 iterations, 38 %: 1, 12 %: 2, ...), so a warp
// that keeps its own 32 boards runs every iteration for its unluckiest lane at ~40 % lane occupancy, and
// packing only once leaves the deep tail to a few lanes.  Here each CTA is persistent and owns a pool of
// boards that are still cascading, parked in shared memory between iterations:
//   fill : when the pool is low, run the cheap convergent part (action, move, first match test, write-back) on
//          the next tile of BLOCK boards and append the boards that matched;
//   round: the top min(pool, BLOCK) boards run ONE cascade iteration on dense warps; finished boards write
//          their reward, the others go back to the pool.
// Deep boards simply stay in the pool while fresh tiles keep streaming in, so every round has (nearly) a
// full complement of lanes; only the final drain of each CTA runs at low occupancy.  Measured (2^20 boards):
// 0.089 ms per step against 0.112 ms for one-thread-per-board and 0.098 ms for packing once.
// The variant that also emits observations (step_obs_kernel) does NOT use a pool.  Its store stream keeps the
// SM's memory-instruction queue (MIO) backed up with STG.128; a cascade that lives purely in registers and
// ALU instructions overlaps that for free, but one that parks boards in shared memory queues its LDS/STS
// behind the stores.  Measured with fp32 obs: 0.159 ms one thread per board, 0.19 ms with this CTA pool,
// 0.21 ms with a barrier-free per-warp pool (which ties the CTA pool at 0.090 ms without obs).
// ------------------------------------------------------------------------------------------------
#if !defined(PPAD_POOL_BLOCK)
#define PPAD_POOL_BLOCK 256
#endif
#if !defined(PPAD_POOL_MINB)
#define PPAD_POOL_MINB (1024 / PPAD_POOL_BLOCK)
#endif
constexpr int PBLOCK = PPAD_POOL_BLOCK;        // threads per CTA of the pooled kernel
#if !defined(PPAD_POOL_HALVES)
#define PPAD_POOL_HALVES 3                     // pool capacity in units of PBLOCK / 2
#endif
constexpr int POOL_SLOTS = PBLOCK * PPAD_POOL_HALVES / 2;
constexpr int POOL_FILL_BELOW = POOL_SLOTS - PBLOCK;   // fill while pool <= this: pool + PBLOCK <= POOL_SLOTS always holds

template <class G>
struct CascadePool {
    typename G::W b0[POOL_SLOTS], b1[POOL_SLOTS], b2[POOL_SLOTS], b3[G::NP > 3 ? POOL_SLOTS : 1];
    typename G::W mm[POOL_SLOTS];         // match mask of the parked board (computed when it was parked)
    double t0[POOL_SLOTS], t1[POOL_SLOTS], t2[POOL_SLOTS], t3[POOL_SLOTS], t4[POOL_SLOTS];
    uint32_t counts[POOL_SLOTS];   // n_combos [0,16) | depth [16,24) | flags [24,32)
    uint32_t source[POOL_SLOTS];   // skyfall cursor [0,16) | injected slab exhausted [16]
    uint64_t owner[POOL_SLOTS];    // board index within the launch
    int warp_count[2][PBLOCK / 32];   // double buffered: consecutive compactions alternate, one barrier each

    __device__ __forceinline__ void put(int slot, const CascadeItem<G> &it, uint64_t own, typename G::W m)
    {
        b0[slot] = it.b.b0;
        b1[slot] = it.b.b1;
        b2[slot] = it.b.b2;
        PPAD_P4(b3[slot] = it.b.b3)
        mm[slot] = m;
        if (it.depth) {   // a board parked by the fill has not been tallied yet: its trackers are zero, not stored
            t0[slot] = it.tally.t0;
            t1[slot] = it.tally.t1;
            t2[slot] = it.tally.t2;
            t3[slot] = it.tally.t3;
            t4[slot] = it.tally.t4;
        }
        const uint32_t nc = it.tally.n_combos > 0xFFFF ? 0xFFFFu : (uint32_t)it.tally.n_combos;
        counts[slot] = nc | (sat255(it.depth) << 16) | ((it.flags & 0xFFu) << 24);
        source[slot] = (it.cursor > 0xFFFF ? 0xFFFFu : (uint32_t)it.cursor) | ((it.exhausted ? 1u : 0u) << 16);
        owner[slot] = own;
    }
    __device__ __forceinline__ uint64_t get(int slot, CascadeItem<G> &it, typename G::W &m) const
    {
        it.b.b0 = b0[slot];
        it.b.b1 = b1[slot];
        it.b.b2 = b2[slot];
        PPAD_P4(it.b.b3 = b3[slot])
        m = mm[slot];
        const uint32_t c = counts[slot], q = source[slot];
        it.tally.n_combos = (int)(c & 0xFFFFu);
        it.depth = (int)((c >> 16) & 0xFFu);
        it.tally.t0 = it.tally.t1 = it.tally.t2 = it.tally.t3 = it.tally.t4 = 0.0;
        if (it.depth) {
            it.tally.t0 = t0[slot];
            it.tally.t1 = t1[slot];
            it.tally.t2 = t2[slot];
            it.tally.t3 = t3[slot];
            it.tally.t4 = t4[slot];
        }
        it.flags = c >> 24;
        it.cursor = (int)(q & 0xFFFFu);
        it.exhausted = (q >> 16) & 1u;
        return owner[slot];
    }
    // CTA-wide stable compaction (ONE __syncthreads; every thread must call it, `phase` alternates 0/1 from call
    // to call: the counts of call k live in buffer k & 1, and call k + 2 can only overwrite them after every thread
    // has passed the barrier of call k + 1, i.e. has finished reading them)
    __device__ __forceinline__ int compact(bool keep, int &total, int phase)
    {
        const unsigned ballot = __ballot_sync(0xFFFFFFFFu, keep);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) warp_count[phase][warp] = __popc(ballot);
        __syncthreads();
        int base = 0, sum = 0;
#pragma unroll
        for (int w = 0; w < PBLOCK / 32; w++) {
            const int c = warp_count[phase][w];
            if (w < warp) base += c;
            sum += c;
        }
        total = sum;
        return base + __popc(ballot & ((1u << lane) - 1u));
    }
};

template <class G, bool AR>
__global__ void __launch_bounds__(PBLOCK, PPAD_POOL_MINB) step_pool_kernel(const __grid_constant__ StepKernelArgs a)
{
    geom_enter<G>(a.p.dyn);
    __shared__ CascadePool<G> pool;
    typedef typename G::W W;
    const int tid = threadIdx.x;
    const int64_t n_tiles = (a.n + PBLOCK - 1) / PBLOCK;
    const bool sd = a.p.skyfall_damage != 0;
    int64_t tile = blockIdx.x;
    int pool_count = 0;                          // CTA-uniform
    int phase = 0;                               // CTA-uniform: which count buffer the next compaction uses

    for (;;) {
        // ---- fill: the convergent part of the step for one fresh tile, when the pool is low
        const bool fill = tile < n_tiles && pool_count <= POOL_FILL_BELOW;
        if (!fill && pool_count == 0) break;    // no tiles left (pool_count == 0 <= POOL_FILL_BELOW) and pool drained
        if (fill) {
            const int64_t i = tile * PBLOCK + tid;
            const bool valid = i < a.n;
            State<G> s;
            CascadeItem<G> it;
            bool cascading = false;
            StepPrefix pre;
            pre.action = 255;
            pre.flags = 0;
            pre.evaluate = false;
            clear_state<G>(s);
            if (valid) {
                s = StateIO<G>::load(a.state, i);
                const RngKey key = board_key(a.p, a.env0 + (uint64_t)i);
                if (a.paths) {
                    int len = a.path_lens ? a.path_lens[i] : a.path_len;
                    len = len > 16 * a.path_words ? 16 * a.path_words : len;
                    pre = path_prefix<G>(s, a.p, a.paths + (size_t)i * a.path_words, len);
                } else {
                    int action = -1;
                    if (a.actions) {
                        action = a.actions[i];
                        if (action == PPAD_SAMPLE) action = -1;
                    }
                    pre = step_prefix<G>(s, a.p, key, action);
                }
            }
            it.init(s.b, pre.flags);             // the cascade works on a copy (game.py:366)
            W m0 = 0;
            if (valid && pre.evaluate) {
                W er, ed;
                m0 = match_mask<G>(it.b, er, ed);
                cascading = m0 != 0;
            }
            if constexpr (AR) {                   // launches with auto_reset; the reset is warp-cooperative
                const bool need = valid && pre.action == ACT_PASS;
                const uint32_t f = reset_boards_warp<G, true>(need, s, a.p, a.p.key.step, a.env0 + (uint64_t)i, nullptr, -1, -1, nullptr);
                if (need) it.flags |= FLAG_EPISODE_DONE | f;
            }
            if (valid) {
                StateIO<G>::store(a.state, i, s.b, s.meta);
                if (a.action_out) a.action_out[i] = (uint8_t)pre.action;
                if (a.final_state) StateIO<G>::store(a.final_state, i, it.b, s.meta);   // planes rewritten after a cascade
                if (a.state_copy) StateIO<G>::store(a.state_copy, i, s.b, s.meta);
                if (!cascading) {
                    if (a.reward) a.reward[i] = 0.0;
                    if (a.info) a.info[i] = it.info();
                    if (a.consumed16) a.consumed16[i] = 0;
                }
            }
            int added;
            const int slot = pool.compact(cascading, added, phase);
            phase ^= 1;
            if (cascading) pool.put(pool_count + slot, it, (uint64_t)i, m0);
            pool_count += added;
            tile += gridDim.x;
        }
        __syncthreads();                          // every put() of this fill and of the previous round is visible

        // ---- round: one cascade iteration for the top m boards of the pool, on dense warps
        const int m = pool_count < PBLOCK ? pool_count : PBLOCK;
        const int base = pool_count - m;
        CascadeItem<G> it;
        uint64_t own = 0;
        bool again = false;
        W mk = 0;
        if (tid < m) {
            own = pool.get(base + tid, it, mk);
            const RngKey key = board_key(a.p, a.env0 + own);
            again = cascade_round_m<G>(it, mk, sd, a.p.cascade_cap, a.p.team,
                                     a.slab ? a.slab + (size_t)own * a.p.slab_words : nullptr, a.p.n_inj, key, a.p.sky,
                                     a.combo_log ? a.combo_log + (size_t)own * a.p.log_cap : nullptr, a.p.log_cap, a.p.slab_fmt);
            if (!again) {
                if (a.reward) a.reward[own] = it.reward();
                if (a.info) a.info[own] = it.info();
                if (a.consumed16) a.consumed16[own] = (uint16_t)(it.cursor > 0xFFFF ? 0xFFFF : it.cursor);
                if (a.final_state) StateIO<G>::store_planes(a.final_state, own, it.b);   // meta: written by the fill
            }
        }

        if (m > 0) {                             // CTA-uniform
            int survivors;
            const int slot = pool.compact(again, survivors, phase);   // its barrier: every get() of the round is done
            phase ^= 1;
            if (again) pool.put(base + slot, it, own, mk);   // read again only after the barrier at the top of the round
            pool_count = base + survivors;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same recirculating pool at WARP granularity: every warp is autonomous -- it pulls tiles of 32 boards from a
// global counter, keeps its own pool of cascading boards in its slice of shared memory and never meets a CTA barrier
// (ncu on step_pool_kernel: the compaction barrier is 33 % of the stall samples).  Tiles are handed out dynamically,
// so the tail of the launch is one tile of 32 boards per warp instead of one of 256 per CTA.
// ------------------------------------------------------------------------------------------------
#if !defined(PPAD_WSLOTS)
#define PPAD_WSLOTS 64
#endif
constexpr int WSLOTS = PPAD_WSLOTS;                      // pool slots per warp: a fill needs 32 free ones

template <class G>
struct WarpPool {
    typename G::W b0[WSLOTS], b1[WSLOTS], b2[WSLOTS], b3[G::NP > 3 ? WSLOTS : 1];
    typename G::W mm[WSLOTS];         // match mask of the parked board (computed when it was parked)
    double t0[WSLOTS], t1[WSLOTS], t2[WSLOTS], t3[WSLOTS], t4[WSLOTS];
    uint32_t counts[WSLOTS];   // n_combos [0,16) | depth [16,24) | flags [24,32)
    uint32_t source[WSLOTS];   // skyfall cursor [0,16) | injected slab exhausted [16]
    uint64_t owner[WSLOTS];    // board index within the launch

    __device__ __forceinline__ void put(int slot, const CascadeItem<G> &it, uint64_t own, typename G::W m)
    {
        b0[slot] = it.b.b0;
        b1[slot] = it.b.b1;
        b2[slot] = it.b.b2;
        PPAD_P4(b3[slot] = it.b.b3)
        mm[slot] = m;
        if (it.depth) {   // a board parked by the fill has not been tallied yet: its trackers are zero, not stored
            t0[slot] = it.tally.t0;
            t1[slot] = it.tally.t1;
            t2[slot] = it.tally.t2;
            t3[slot] = it.tally.t3;
            t4[slot] = it.tally.t4;
        }
        const uint32_t nc = it.tally.n_combos > 0xFFFF ? 0xFFFFu : (uint32_t)it.tally.n_combos;
        counts[slot] = nc | (sat255(it.depth) << 16) | ((it.flags & 0xFFu) << 24);
        source[slot] = (it.cursor > 0xFFFF ? 0xFFFFu : (uint32_t)it.cursor) | ((it.exhausted ? 1u : 0u) << 16);
        owner[slot] = own;
    }
    __device__ __forceinline__ uint64_t get(int slot, CascadeItem<G> &it, typename G::W &m) const
    {
        it.b.b0 = b0[slot];
        it.b.b1 = b1[slot];
        it.b.b2 = b2[slot];
        PPAD_P4(it.b.b3 = b3[slot])
        m = mm[slot];
        const uint32_t c = counts[slot], q = source[slot];
        it.tally.n_combos = (int)(c & 0xFFFFu);
        it.depth = (int)((c >> 16) & 0xFFu);
        it.tally.t0 = it.tally.t1 = it.tally.t2 = it.tally.t3 = it.tally.t4 = 0.0;
        if (it.depth) {
            it.tally.t0 = t0[slot];
            it.tally.t1 = t1[slot];
            it.tally.t2 = t2[slot];
            it.tally.t3 = t3[slot];
            it.tally.t4 = t4[slot];
        }
        it.flags = c >> 24;
        it.cursor = (int)(q & 0xFFFFu);
        it.exhausted = (q >> 16) & 1u;
        return owner[slot];
    }
};

#if !defined(PPAD_WPOOL_MINB)
#define PPAD_WPOOL_MINB 4
#endif
template <class G, bool AR>
__global__ void __launch_bounds__(BLOCK, PPAD_WPOOL_MINB) step_wpool_kernel(const __grid_constant__ StepKernelArgs a)
{
    geom_enter<G>(a.p.dyn);
    extern __shared__ __align__(16) unsigned char wpool_raw[];   // BLOCK / 32 WarpPool<G> (dynamic: more than 48 KB)
    typedef typename G::W W;
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    WarpPool<G> &pool = reinterpret_cast<WarpPool<G> *>(wpool_raw)[threadIdx.x >> 5];
    const int64_t n_tiles = (a.n + 31) / 32;
    const bool sd = a.p.skyfall_damage != 0;
    int count = 0;                               // warp-uniform: boards in this warp's pool
    bool more = true;                            // warp-uniform: the global counter has not run out

    for (;;) {
        // ---- fill: the convergent part of the step for one fresh tile of 32 boards, when the pool has room.
        // (Fetching the next tile index one fill ahead to hide the atomic's round trip -- 17 % of the stall samples --
        // was slower, 0.0731 against 0.0692 ms: a warp then holds a tile it is not yet working on, and with seven
        // tiles per warp that costs more at the tail than the latency it hides.)
        if (more && count <= WSLOTS - 32) {
            unsigned t32 = 0;
            if (lane == 0) t32 = atomicAdd(a.tile_counter, 1u);
            const int64_t tile = (int64_t)__shfl_sync(full, t32, 0);
            if (tile >= n_tiles) {
                more = false;
            } else {
                const int64_t i = tile * 32 + lane;
                const bool valid = i < a.n;
                State<G> s;
                CascadeItem<G> it;
                bool cascading = false;
                StepPrefix pre;
                pre.action = 255;
                pre.flags = 0;
                pre.evaluate = false;
                clear_state<G>(s);
                if (valid) {
                    s = StateIO<G>::load(a.state, i);
                    const RngKey key = board_key(a.p, a.env0 + (uint64_t)i);
                    if (a.paths) {
                        int len = a.path_lens ? a.path_lens[i] : a.path_len;
                        len = len > 16 * a.path_words ? 16 * a.path_words : len;
                        pre = path_prefix<G>(s, a.p, a.paths + (size_t)i * a.path_words, len);
                    } else {
                        int action = -1;
                        if (a.actions) {
                            action = a.actions[i];
                            if (action == PPAD_SAMPLE) action = -1;
                        }
                        pre = step_prefix<G>(s, a.p, key, action);
                    }
                }
                it.init(s.b, pre.flags);             // the cascade works on a copy (game.py:366)
                W m0 = 0;
                if (valid && pre.evaluate) {
                    W er, ed;
                    m0 = match_mask<G>(it.b, er, ed);
                    cascading = m0 != 0;
                }
                if constexpr (AR) {                   // launches with auto_reset; the reset is warp-cooperative
                    const bool need = valid && pre.action == ACT_PASS;
                    const uint32_t f = reset_boards_warp<G, true>(need, s, a.p, a.p.key.step, a.env0 + (uint64_t)i, nullptr, -1, -1, nullptr);
                    if (need) it.flags |= FLAG_EPISODE_DONE | f;
                }
                if (valid) {
                    StateIO<G>::store(a.state, i, s.b, s.meta);
                    if (a.action_out) a.action_out[i] = (uint8_t)pre.action;
                    if (a.final_state) StateIO<G>::store(a.final_state, i, it.b, s.meta);   // planes rewritten after a cascade
                    if (a.state_copy) StateIO<G>::store(a.state_copy, i, s.b, s.meta);
                    if (!cascading) {
                        if (a.reward) a.reward[i] = 0.0;
                        if (a.info) a.info[i] = it.info();
                        if (a.consumed16) a.consumed16[i] = 0;
                    }
                }
                const unsigned ballot = __ballot_sync(full, cascading);
                if (cascading) pool.put(count + __popc(ballot & ((1u << lane) - 1u)), it, (uint64_t)i, m0);
                count += __popc(ballot);
                __syncwarp();
                if (count <= WSLOTS - 32) continue;   // still room for a whole tile: fill on, the rounds run on full warps
            }
        }
        if (count == 0) {
            if (!more) break;
            continue;
        }
        // ---- round: one cascade iteration for the top min(count, 32) boards of the pool
        const int m = count < 32 ? count : 32;
        const int base = count - m;
        CascadeItem<G> it;
        uint64_t own = 0;
        bool again = false;
        W mk = 0;
        if (lane < m) {
            own = pool.get(base + lane, it, mk);
            const RngKey key = board_key(a.p, a.env0 + own);
            again = cascade_round_m<G>(it, mk, sd, a.p.cascade_cap, a.p.team,
                                     a.slab ? a.slab + (size_t)own * a.p.slab_words : nullptr, a.p.n_inj, key, a.p.sky,
                                     a.combo_log ? a.combo_log + (size_t)own * a.p.log_cap : nullptr, a.p.log_cap, a.p.slab_fmt);
            if (!again) {
                if (a.reward) a.reward[own] = it.reward();
                if (a.info) a.info[own] = it.info();
                if (a.consumed16) a.consumed16[own] = (uint16_t)(it.cursor > 0xFFFF ? 0xFFFF : it.cursor);
                if (a.final_state) StateIO<G>::store_planes(a.final_state, own, it.b);   // meta: written by the fill
            }
        }
        __syncwarp();                                // every get() of the round is done before the survivors are put back
        const unsigned ballot = __ballot_sync(full, again);
        if (again) pool.put(base + __popc(ballot & ((1u << lane) - 1u)), it, own, mk);
        count = base + __popc(ballot);
        __syncwarp();
    }
    // the last CTA to finish leaves the counters at 0 for the launch that uses them next
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.tile_counter + 1, 1u) == gridDim.x - 1) {
            a.tile_counter[0] = 0u;
            a.tile_counter[1] = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-geometry launchers (explicitly instantiated in ppad_step_inst.cu)
// ------------------------------------------------------------------------------------------------
enum StepKernelChoice { STEP_KERNEL_OBS = 0, STEP_KERNEL_THREAD = 1, STEP_KERNEL_POOL = 2, STEP_KERNEL_WPOOL = 3 };

template <class G>
int launch_step_g(const ppad_ctx *ctx, const StepKernelArgs &a, int choice, cudaStream_t st)
{
    const bool ar = a.p.auto_reset != 0;
    if (a.c_records) {   // compact result stream (host pipe): one thread per board, with or without the observation
        typedef CompactLayout<G> CL;
        const size_t sm = ar ? CL::MAX_STAGE : CL::MAX_STAGE_NO_RESET;
        const unsigned grid = grid_for(a.n);
        if (choice == STEP_KERNEL_OBS) {
            if (ar) step_obs_kernel<G, true, true, true><<<grid, BLOCK, sm, st>>>(a);
            else step_obs_kernel<G, true, false, true><<<grid, BLOCK, sm, st>>>(a);
        } else {
            if (ar) step_obs_kernel<G, false, true, true><<<grid, BLOCK, sm, st>>>(a);
            else step_obs_kernel<G, false, false, true><<<grid, BLOCK, sm, st>>>(a);
        }
        return after_launch("step_obs_kernel<compact>");
    }
    if (choice == STEP_KERNEL_OBS) {
        if (ar) step_obs_kernel<G, true, true><<<grid_for(a.n), BLOCK, 0, st>>>(a);
        else step_obs_kernel<G, true, false><<<grid_for(a.n), BLOCK, 0, st>>>(a);
        return after_launch("step_obs_kernel");
    }
    if (choice == STEP_KERNEL_THREAD) {
        if (ar) step_obs_kernel<G, false, true><<<grid_for(a.n), BLOCK, 0, st>>>(a);
        else step_obs_kernel<G, false, false><<<grid_for(a.n), BLOCK, 0, st>>>(a);
        return after_launch("step_obs_kernel<no obs>");
    }
    if (choice == STEP_KERNEL_WPOOL) {
        // persistent warps, dynamic tiles of 32 boards
        const size_t sm = sizeof(WarpPool<G>) * (BLOCK / 32);
        static int cached_ctas[2] = {0, 0};      // per instantiation (the attribute is per function, all devices alike)
        int &ctas = cached_ctas[ar ? 1 : 0];
        if (ctas == 0) {
            if (ar) {
                cudaFuncSetAttribute(step_wpool_kernel<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, step_wpool_kernel<G, true>, BLOCK, sm);
            } else {
                cudaFuncSetAttribute(step_wpool_kernel<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, step_wpool_kernel<G, false>, BLOCK, sm);
            }
            if (ctas <= 0) ctas = 1;
        }
        const int64_t want = (a.n + BLOCK - 1) / BLOCK;
        int64_t grid = (int64_t)ctx->sm_count * (ctas > 0 ? ctas : 1);
        if (grid > want) grid = want;
        if (ar) step_wpool_kernel<G, true><<<(unsigned)grid, BLOCK, sm, st>>>(a);
        else step_wpool_kernel<G, false><<<(unsigned)grid, BLOCK, sm, st>>>(a);
        return after_launch("step_wpool_kernel");
    }
    // persistent: one wave of CTAs, as many as fit
    const int64_t tiles = (a.n + PBLOCK - 1) / PBLOCK;
    int64_t grid = (int64_t)ctx->sm_count * ctx->pool_ctas_per_sm;
    if (grid > tiles) grid = tiles;
    if (ar) step_pool_kernel<G, true><<<(unsigned)grid, PBLOCK, 0, st>>>(a);
    else step_pool_kernel<G, false><<<(unsigned)grid, PBLOCK, 0, st>>>(a);
    return after_launch("step_pool_kernel");
}

template <class G>
int launch_rollout_g(const RolloutArgs &a, bool record, cudaStream_t st)
{
    if (record) rollout_kernel<G, true><<<grid_for(a.n), BLOCK, 0, st>>>(a);
    else rollout_kernel<G, false><<<grid_for(a.n), BLOCK, 0, st>>>(a);
    return after_launch("rollout_kernel");
}

template <class G>
int pool_occupancy_g()
{
    int ctas = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, step_pool_kernel<G, true>, PBLOCK, 0);
    return ctas;
}

#define PPAD_STEP_LAUNCHERS(PREFIX, G)                                                                        \
    PREFIX template int launch_step_g<G>(const ppad_ctx *, const StepKernelArgs &, int, cudaStream_t);        \
    PREFIX template int launch_rollout_g<G>(const RolloutArgs &, bool, cudaStream_t);                         \
    PREFIX template int pool_occupancy_g<G>();

}   // namespace ppad
