// ppad_agent.cu -- the callers either side of the env step: action selection from Q-values, the
// discounted-return scan, the device-side experience replay ring; kernels + C ABI.
#include "ppad_device.cuh"

namespace ppad {

// Agent01.act after the network (agent01.py:173-218): Q-values -> one action per board
template <class G>
__global__ void __launch_bounds__(BLOCK) select_kernel(const DynGeomData dg, int64_t n, uint64_t env0, RngKey key, const void *state,
                                                       const float *q, int method, float beta, float epsilon,
                                                       int max_moves, uint8_t *actions)
{
    geom_enter<G>(dg);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const State<G> s = StateIO<G>::load(state, i);
    const uint64_t env = env0 + (uint64_t)i;
    key.env_lo = (uint32_t)env;
    key.env_hi = (uint32_t)(env >> 32);
    if (max_moves > 0 && meta_moves(s.meta) >= max_moves) {   // simulation.py:50-52: episode length cap
        actions[i] = ACT_PASS;
        return;
    }
    float qi[5];
#pragma unroll
    for (int a = 0; a < 5; a++) qi[a] = __ldg(q + 5 * i + a);
    const U4 blk = philox4x32_10(U4{key.env_lo, key.env_hi, key.step, STREAM_POLICY << 24}, key.k0, key.k1);
    actions[i] = (uint8_t)select_action<G>(qi, meta_finger(s.meta), meta_prev(s.meta), method, beta, epsilon, blk.x, blk.y);
}

// ppad.discount (agent/utils.py:22-51) for n trajectories stored row-major [n, max_len].  One thread walks one row;
// consecutive iterations of a lane touch consecutive doubles, so L1 / L2 turn the strided warp accesses into whole
// sectors: measured 0.065 ms for 2^20 x 51 (rows of 1..51 valid steps), i.e. HBM speed on the bytes touched.  Staging
// the rows through shared memory was 3x slower (0.21 ms: it moves the untouched tails too, at a quarter of the occupancy).
__global__ void __launch_bounds__(BLOCK) discount_kernel(int64_t n, int max_len, double *rewards, const int32_t *lengths,
                                                         double gamma, int log10_flag, int norm)
{
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    int len = lengths ? lengths[i] : max_len;
    len = len < 0 ? 0 : (len > max_len ? max_len : len);
    discount_trajectory(rewards + (size_t)i * max_len, len, gamma, log10_flag, norm);
}

// experience replay ring (simulation.py:294-306): rows are (packed state, action, reward)
__global__ void __launch_bounds__(BLOCK) replay_push_kernel(int64_t n, int words, int64_t capacity, int64_t cursor,
                                                            int64_t count, const uint32_t *state, const uint8_t *action,
                                                            const double *reward, const uint8_t *keep,
                                                            const int64_t *slot, uint32_t *r_state, uint8_t *r_action,
                                                            double *r_reward)
{
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    if (keep && !keep[i]) return;
    const int64_t si = slot ? slot[i] : i;
    if (si < count - capacity) return;   // a push larger than the ring keeps its last `capacity` rows: one writer per slot
    const int64_t d = (cursor + si) % capacity;
    for (int w = 0; w < words; w++) r_state[d * words + w] = state[i * words + w];
    r_action[d] = action[i];
    r_reward[d] = reward[i];
}

__global__ void __launch_bounds__(BLOCK) replay_sample_kernel(int64_t k, int words, int64_t size, RngKey key,
                                                              const uint32_t *r_state, const uint8_t *r_action,
                                                              const double *r_reward, uint32_t *state, uint8_t *action,
                                                              double *reward, int64_t *index)
{
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= k) return;
    key.env_lo = (uint32_t)i;
    key.env_hi = (uint32_t)((uint64_t)i >> 32);
    const uint32_t hi = philox_draw(key, STREAM_POLICY, 2), lo = philox_draw(key, STREAM_POLICY, 3);
    const uint64_t x = ((uint64_t)hi << 32) | lo;
    const int64_t j = (int64_t)__umul64hi(x, (uint64_t)size);   // uniform in [0, size)
    for (int w = 0; w < words; w++) state[i * words + w] = r_state[j * words + w];
    action[i] = r_action[j];
    reward[i] = r_reward[j];
    if (index) index[i] = j;
}

}   // namespace ppad

using namespace ppad;

extern "C" {

int ppad_select_actions(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, const void *state, const float *q,
                        int32_t method, float beta, float epsilon, int32_t max_moves, uint8_t *actions, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state || !q || !actions) return fail(PPAD_ERR_INVALID, "ppad_select_actions: null pointer");
    if (method != SELECT_MAX && method != SELECT_BOLTZMANN)
        return fail(PPAD_ERR_INVALID, "ERROR: Unknown action method!");
    RngKey key = ctx->base.key;
    key.step = step;
    PPAD_DISPATCH(ctx, (select_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(ctx->base.dyn, n, env0, key, state, q, method, beta, epsilon, max_moves, actions)));
    return after_launch("select_kernel");
}

int ppad_discount(const ppad_ctx *ctx, int64_t n, int32_t max_len, double *rewards, const int32_t *lengths, double gamma,
                  int32_t log10_flag, int32_t norm, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0 || max_len == 0) return PPAD_OK;
    if (!rewards || max_len < 0) return fail(PPAD_ERR_INVALID, "ppad_discount: bad argument");
    discount_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, max_len, rewards, lengths, gamma, log10_flag, norm);
    return after_launch("discount_kernel");
}

int ppad_replay_push(const ppad_ctx *ctx, int64_t n, int64_t capacity, int64_t cursor, int64_t count, const void *state,
                     const uint8_t *action, const double *reward, const uint8_t *keep, const int64_t *slot,
                     void *r_state, uint8_t *r_action, double *r_reward, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state || !action || !reward || !r_state || !r_action || !r_reward || capacity <= 0 || cursor < 0)
        return fail(PPAD_ERR_INVALID, "ppad_replay_push: bad argument");
    if (count < 0 || count > n || (!keep && !slot && count != n))
        return fail(PPAD_ERR_INVALID, "ppad_replay_push: count must be the number of kept rows (n without a keep mask)");
    const int words = (int)(ppad_state_bytes_ex(ctx->cfg.rows, ctx->cfg.cols, ctx->cfg.orb_bits) / 4);
    replay_push_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, words, capacity, cursor, count, (const uint32_t *)state,
                                                                        action, reward, keep, slot, (uint32_t *)r_state,
                                                                        r_action, r_reward);
    return after_launch("replay_push_kernel");
}

int ppad_replay_sample(const ppad_ctx *ctx, int64_t k, int64_t size, uint32_t step, const void *r_state,
                       const uint8_t *r_action, const double *r_reward, void *state, uint8_t *action, double *reward,
                       int64_t *index, void *stream)
{
    PPAD_ENTER(ctx, k);
    if (k == 0) return PPAD_OK;
    if (!r_state || !r_action || !r_reward || !state || !action || !reward || size <= 0)
        return fail(PPAD_ERR_INVALID, "ppad_replay_sample: bad argument");
    const int words = (int)(ppad_state_bytes_ex(ctx->cfg.rows, ctx->cfg.cols, ctx->cfg.orb_bits) / 4);
    RngKey key = ctx->base.key;
    key.step = step;
    replay_sample_kernel<<<grid_for(k), BLOCK, 0, (cudaStream_t)stream>>>(k, words, size, key, (const uint32_t *)r_state,
                                                                          r_action, r_reward, (uint32_t *)state, action,
                                                                          reward, index);
    return after_launch("replay_sample_kernel");
}

}   // extern "C"
