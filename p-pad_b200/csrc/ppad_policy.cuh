// ppad_policy.cuh -- the policy-driven env step of a device-resident rollout (BASELINE configs[4]) as ONE kernel:
//
//   q[n, 5] (the agent's Q-values)                                   projects/simulation.py:96 (agent.act)
//     -> mask off-board / reversing moves, epsilon-greedy, argmax or Boltzmann draw     ppad/agent/agent01.py:173-218
//     -> episode length cap                                                             simulation.py:50-52
//     -> push (state BEFORE the step, action) into the replay ring                      simulation.py:60-61, 294-299
//     -> env step: move, or pass = cascade resolution + float64 reward                  ppad/pad/game.py:329-405
//     -> auto reset of the boards that passed                                           simulation.py:44, game.py:282-323
//     -> push the reward; when an episode ends, write the discounted returns            ppad/agent/utils.py:22-51
//        gamma^j * r back into the ring rows of its earlier steps                       (simulation.py:59: discount)
//     -> one-hot observation of the new state for the agent's next forward pass         agent01.py:254-273
//     -> per-CTA episode statistics (for the NCCL all-gather every K steps)             simulation.py:336-349
//
// Round 1 ran this as select -> state.clone() -> step -> push -> ~10 torch reductions (>= 5 launches and 4.9 ms per
// 2^23-board step, 0.23 of the HBM roofline).  Fused, the step touches per board: state in/out 32 B, q 20 B, obs
// 840 B, results 13 B, ring row 25 B = 930 B.
#pragma once
#include "ppad_step.cuh"

#if !defined(PPAD_POLICY_REFILL_UNROLL)
#define PPAD_POLICY_REFILL_UNROLL 1
#endif

namespace ppad {

constexpr int STATS_WORDS = 40;   // uint64 per CTA row, layout = dist.STAT_FIELDS + depth hist [16] + combo hist [16]
enum : int { ST_BOARDS = 0, ST_EVALUATED = 1, ST_REWARD_SUM = 2, ST_REWARD_MAX = 3, ST_COMBOS = 4, ST_CONSUMED = 5,
             ST_EPISODES = 6, ST_FAULTS = 7, ST_DEPTH_HIST = 8, ST_COMBO_HIST = 24 };
constexpr double STATS_REWARD_SCALE = 1024.0;   // reward sums are kept in fixed point: exact, order independent

struct PolicyArgs {
    StepParams p;
    int64_t n;
    uint64_t env0;
    void *state;
    const float *q;           // [n, 5]
    int32_t method;           // SELECT_MAX / SELECT_BOLTZMANN
    float beta, epsilon;
    const uint32_t *slab;     // injected skyfall or NULL (Philox)
    uint8_t *action_out;      // [n] or NULL
    double *reward;           // [n] or NULL
    uint32_t *info;           // [n] or NULL
    void *obs;                // [n, R, C, 7] or NULL
    int32_t obs_dtype;
    // replay ring, time-major: the row of (this step, board i) is ((ring_step % ring_steps) * n + i)
    void *r_state;            // [ring_steps * n] packed records, or NULL (no replay)
    uint8_t *r_action;
    double *r_reward;
    int64_t ring_steps, ring_step;
    int64_t ring_slot;        // ring_step % ring_steps (host-computed: no 64-bit modulo in the kernel)
    double gamma;             // > 0: discounted returns are written back when an episode ends (agent/utils.py:22-51)
    int32_t log10_reward;     // log10 of a positive final reward before discounting (utils.py:37-40)
    unsigned long long *stats;   // [grid, STATS_WORDS] per-CTA accumulators (added to, never reset here), or NULL
};

// MINB: resident CTAs per SM the register allocation aims at: 4 (64 registers, ~100 B of spills) or 3 (80 registers, no
// spills).  Measured per 2^20 boards (profiles/r02f_ab.txt): 0.2198 ms with 4 against 0.2250 ms with 3 once the code was
// made smaller (before: 0.269 against 0.254) -- occupancy beats the spills.
// INJ: the launch has an injected skyfall slab.  Without one (the usual rollout: Philox skyfall) the instantiation does
// not carry the slab paths of the skyfall at all -- this kernel is instruction-cache sensitive.
template <class G, bool OBS, int MINB, bool INJ>
__global__ void __launch_bounds__(BLOCK, MINB) policy_step_kernel(const __grid_constant__ PolicyArgs a)
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
    const uint64_t env = a.env0 + (uint64_t)i;
    State<G> s;
    clear_state<G>(s);
    int action = ACT_PASS;
    const int64_t row = a.ring_slot * a.n + i;
    if (valid) {
        s = StateIO<G>::load(a.state, i);
        // ---- Agent01.act after the network (agent01.py:173-218) + the episode length cap (simulation.py:50-52)
        if (a.p.max_moves > 0 && meta_moves(s.meta) >= a.p.max_moves) {
            action = ACT_PASS;
        } else {
            float qi[5];
#pragma unroll
            for (int k = 0; k < 5; k++) qi[k] = __ldg(a.q + 5 * i + k);
            RngKey key = board_key(a.p, env);
            const U4 blk = philox4x32_10(U4{key.env_lo, key.env_hi, key.step, STREAM_POLICY << 24}, key.k0, key.k1);
            action = select_action<G>(qi, meta_finger(s.meta), meta_prev(s.meta), a.method, a.beta, a.epsilon, blk.x, blk.y);
        }
        // ---- the ring row's observation is the state the action was chosen in (simulation.py:60-61)
        if (a.r_state) {
            StateIO<G>::store(a.r_state, row, s.b, s.meta);
            a.r_action[row] = (uint8_t)action;
        }
    }
    // ---- the env step (every lane takes part: the cascade loop and the reset are warp-cooperative)
    const int episode_len = meta_moves(s.meta);   // moves made before this step (a pass ends the episode)
    Board<G> fin;
    const uint32_t *slab = nullptr;
    if constexpr (INJ) slab = a.slab && valid ? a.slab + (size_t)i * a.p.slab_words : nullptr;
    const StepOut o = step_board<G, true, true, PPAD_POLICY_REFILL_UNROLL>(valid, s, fin, a.p, env, action, slab, nullptr);
    if (valid) {
        StateIO<G>::store(a.state, i, s.b, s.meta);
        if (a.action_out) a.action_out[i] = (uint8_t)o.action;
        if (a.reward) a.reward[i] = o.reward;
        if (a.info) a.info[i] = o.info;
        if (a.r_state) {
            // ---- reward of this row; at the end of an episode the discounted returns of its earlier rows:
            // discount() scans backwards cur = cur * gamma + r with r = 0 for the moves (utils.py:41-45)
            double r = o.reward;
            if (a.log10_reward && r > 0.0) r = log10(r);
            a.r_reward[row] = r;
            if (a.gamma > 0.0 && o.action == ACT_PASS && o.reward > 0.0) {   // all-non-positive episodes stay as they are (utils.py:32)
                int64_t back = episode_len < a.ring_steps - 1 ? episode_len : a.ring_steps - 1;   // rows still in the ring
                if (back > a.ring_step) back = a.ring_step;                                        // rows that were ever pushed
                double cur = r;
                int64_t slot = a.ring_slot;
                for (int64_t j = 0; j < back; j++) {
                    cur = dadd(dmul(cur, a.gamma), 0.0);
                    slot = slot == 0 ? a.ring_steps - 1 : slot - 1;
                    a.r_reward[slot * a.n + i] = cur;
                }
            }
        }
    }
    // ---- episode statistics: warp reductions, then a handful of 64-bit reductions (RED) into this CTA's row in global
    // memory -- no shared memory and no barrier, so a warp that is done does not wait for the CTA's slowest one
    if (a.stats) {   // kernel-uniform: every lane takes part in the reductions
        const unsigned full = 0xFFFFFFFFu;
        unsigned long long *row_st = a.stats + (size_t)blockIdx.x * STATS_WORDS;
        const uint32_t inf = valid ? o.info : 0u;
        const uint32_t combos = inf & 0xFFu, depth = (inf >> 8) & 0xFFu, consumed = (inf >> 16) & 0xFFu, flags = inf >> 24;
        const bool evaluated = valid && (o.action == ACT_PASS || a.p.eval_moves != 0);
        // histograms over the evaluated boards (the passes of a policy rollout); most of them match nothing (bin 0 of
        // both: counted per warp), the others add themselves
        if (evaluated && depth) {
            atomicAdd(&row_st[ST_DEPTH_HIST + (depth < 15 ? depth : 15)], 1ull);
            atomicAdd(&row_st[ST_COMBO_HIST + (combos < 15 ? combos : 15)], 1ull);
        }
        const unsigned n_zero = __popc(__ballot_sync(full, evaluated && depth == 0));
        const unsigned n_valid_w = __popc(__ballot_sync(full, valid));
        const unsigned n_eval = __popc(__ballot_sync(full, depth != 0));
        const unsigned n_done = __popc(__ballot_sync(full, (flags & FLAG_EPISODE_DONE) != 0));
        const unsigned n_fault = __popc(__ballot_sync(full, (flags & 0x7Fu) != 0));
        const unsigned sum_combos = __reduce_add_sync(full, combos), sum_consumed = __reduce_add_sync(full, consumed);
        unsigned long long fx = valid && depth ? (unsigned long long)llrint(o.reward * STATS_REWARD_SCALE) : 0ull;
        unsigned long long mx = valid && depth ? (unsigned long long)__double_as_longlong(o.reward) : 0ull;   // rewards are >= 0
        if (n_eval) {   // warp-uniform
#pragma unroll
            for (int off = 16; off; off >>= 1) {
                fx += __shfl_xor_sync(full, fx, off);
                const unsigned long long other = __shfl_xor_sync(full, mx, off);
                mx = other > mx ? other : mx;
            }
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&row_st[ST_BOARDS], (unsigned long long)n_valid_w);
            if (n_zero) {
                atomicAdd(&row_st[ST_DEPTH_HIST], (unsigned long long)n_zero);
                atomicAdd(&row_st[ST_COMBO_HIST], (unsigned long long)n_zero);
            }
            if (n_eval) {
                atomicAdd(&row_st[ST_EVALUATED], (unsigned long long)n_eval);
                atomicAdd(&row_st[ST_REWARD_SUM], fx);
                atomicMax(&row_st[ST_REWARD_MAX], mx);
                atomicAdd(&row_st[ST_COMBOS], (unsigned long long)sum_combos);
                atomicAdd(&row_st[ST_CONSUMED], (unsigned long long)sum_consumed);
            }
            if (n_done) atomicAdd(&row_st[ST_EPISODES], (unsigned long long)n_done);
            if (n_fault) atomicAdd(&row_st[ST_FAULTS], (unsigned long long)n_fault);
        }
    }
    const int64_t warp0 = i - (threadIdx.x & 31);
    if (OBS && warp0 < a.n) {
        const int n_valid = (int)(a.n - warp0 < 32 ? a.n - warp0 : 32);
        emit_obs_any<G>(s, valid, n_valid, obs_stream + (threadIdx.x >> 5) * ObsGeom<G>::WARP_WORDS, obs_sh,
                        obs_warp_ptr<G>(a.obs, warp0, a.obs_dtype), a.obs_dtype);
    }
}

template <class G, bool OBS, bool INJ>
void launch_policy_minb(const PolicyArgs &a, int minb, unsigned grid, cudaStream_t st)
{
#if defined(PPAD_POLICY_TUNING)      // A/B builds: also the 3-CTAs-per-SM (80 registers, no spills) instantiation
    if (minb == 3) {
        policy_step_kernel<G, OBS, 3, INJ><<<grid, BLOCK, 0, st>>>(a);
        return;
    }
#endif
    (void)minb;
    policy_step_kernel<G, OBS, 4, INJ><<<grid, BLOCK, 0, st>>>(a);
}

template <class G>
int launch_policy_g(const PolicyArgs &a, int minb, cudaStream_t st)
{
    const unsigned grid = grid_for(a.n);
    if (a.obs) {
        if (a.slab) launch_policy_minb<G, true, true>(a, minb, grid, st);
        else launch_policy_minb<G, true, false>(a, minb, grid, st);
    } else {
        if (a.slab) launch_policy_minb<G, false, true>(a, minb, grid, st);
        else launch_policy_minb<G, false, false>(a, minb, grid, st);
    }
    return after_launch("policy_step_kernel");
}

#define PPAD_POLICY_LAUNCHERS(PREFIX, G) PREFIX template int launch_policy_g<G>(const PolicyArgs &, int, cudaStream_t);

}   // namespace ppad
