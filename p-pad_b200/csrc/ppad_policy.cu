// ppad_policy.cu -- C ABI of the fused policy step (ppad_policy.cuh; kernels instantiated per geometry in
// ppad_step_inst.cu) and the reduction of its per-CTA statistics rows.
#include "ppad_policy.cuh"

namespace ppad {

PPAD_POLICY_LAUNCHERS(extern, G56)
PPAD_POLICY_LAUNCHERS(extern, G67)
PPAD_POLICY_LAUNCHERS(extern, G45)
PPAD_POLICY_LAUNCHERS(extern, G56W)
PPAD_POLICY_LAUNCHERS(extern, G67W)
PPAD_POLICY_LAUNCHERS(extern, G45W)
PPAD_POLICY_LAUNCHERS(extern, GDYN)
PPAD_POLICY_LAUNCHERS(extern, GDYNW)

// Sums the per-CTA rows (max for the reward maximum) into out[STATS_WORDS] (zeroed before the launch).  A row is 320
// contiguous bytes: 6 rows x 40 words per pass of a CTA's 240 working threads, consecutive threads on consecutive words;
// one 64-bit reduction per word and CTA at the end.  (One CTA walking all rows took 180 us for 8192 rows: a third of
// the 2^21-board policy step it follows.)
constexpr int REDUCE_ROWS_PER_PASS = 6;
__global__ void __launch_bounds__(BLOCK) stats_reduce_kernel(const unsigned long long *rows, int64_t n_rows, unsigned long long *out)
{
    __shared__ unsigned long long part[REDUCE_ROWS_PER_PASS][STATS_WORDS];
    const int w = threadIdx.x % STATS_WORDS, lane_row = threadIdx.x / STATS_WORDS;
    const bool is_max = w == ST_REWARD_MAX;
    if (lane_row < REDUCE_ROWS_PER_PASS) {
        unsigned long long acc = 0;
        for (int64_t r = (int64_t)blockIdx.x * REDUCE_ROWS_PER_PASS + lane_row; r < n_rows; r += (int64_t)gridDim.x * REDUCE_ROWS_PER_PASS) {
            const unsigned long long v = rows[r * STATS_WORDS + w];
            acc = is_max ? (v > acc ? v : acc) : acc + v;
        }
        part[lane_row][w] = acc;
    }
    __syncthreads();
    if (threadIdx.x < STATS_WORDS) {
        unsigned long long acc = 0;
        for (int k = 0; k < REDUCE_ROWS_PER_PASS; k++) {
            const unsigned long long v = part[k][threadIdx.x];
            acc = is_max ? (v > acc ? v : acc) : acc + v;
        }
        if (is_max) atomicMax(&out[threadIdx.x], acc);
        else atomicAdd(&out[threadIdx.x], acc);
    }
}

}   // namespace ppad

using namespace ppad;

extern "C" {

int ppad_policy_step(const ppad_ctx *ctx, const ppad_policy_args *args, void *stream)
{
    if (!args) return fail(PPAD_ERR_INVALID, "ppad_policy_step: null args");
    PPAD_ENTER(ctx, args->n);
    if (args->n == 0) return PPAD_OK;
    if (!args->state || !args->q) return fail(PPAD_ERR_INVALID, "ppad_policy_step: null state / q");
    if (args->method != SELECT_MAX && args->method != SELECT_BOLTZMANN)
        return fail(PPAD_ERR_INVALID, "ERROR: Unknown action method!");
    if (((uintptr_t)args->state & 15) || (args->obs && ((uintptr_t)args->obs & 15)) ||
        (args->ring_state && ((uintptr_t)args->ring_state & 15)))
        return fail(PPAD_ERR_INVALID, "ppad_policy_step: state, obs and ring_state must be 16-byte aligned");
    if (args->obs && args->obs_dtype != PPAD_OBS_F32 && args->obs_dtype != PPAD_OBS_U8)
        return fail(PPAD_ERR_UNSUPPORTED, "ppad_policy_step: obs_dtype must be PPAD_OBS_F32 or PPAD_OBS_U8");
    if (args->ring_state && (!args->ring_action || !args->ring_reward || args->ring_steps <= 0 || args->ring_step < 0))
        return fail(PPAD_ERR_INVALID, "ppad_policy_step: the replay ring needs state, action, reward, ring_steps > 0, ring_step >= 0");
    if (args->slab_format != PPAD_SLAB_NIBBLES && args->slab_format != PPAD_SLAB_PLANES3)
        return fail(PPAD_ERR_INVALID, "ppad_policy_step: unknown slab_format");
    if (args->slab && args->slab_format == PPAD_SLAB_PLANES3) {
        if (ctx->geom >= 3) return fail(PPAD_ERR_UNSUPPORTED, "ppad_policy_step: PPAD_SLAB_PLANES3 needs a context with orb_bits = 3");
        if (args->slab_words <= 0 || args->slab_words % 3 || args->n_inj < 0 || args->n_inj > args->slab_words / 3 * 32)
            return fail(PPAD_ERR_INVALID, "ppad_policy_step: PPAD_SLAB_PLANES3 needs slab_words = 3 * groups and 0 <= n_inj <= 32 * groups");
    } else if (args->slab && (args->slab_words <= 0 || args->n_inj < 0 || args->n_inj > args->slab_words * 8))
        return fail(PPAD_ERR_INVALID, "ppad_policy_step: need 0 <= n_inj <= 8 * slab_words");
    PolicyArgs a;
    a.p = ctx->base;
    a.p.key.step = args->step;
    a.p.eval_moves = args->eval_moves;
    a.p.auto_reset = args->auto_reset;
    a.p.max_moves = args->max_moves;
    a.p.n_inj = args->slab ? args->n_inj : -1;
    a.p.slab_words = args->slab_words;
    a.p.slab_fmt = args->slab_format;
    a.n = args->n;
    a.env0 = args->env0;
    a.state = args->state;
    a.q = args->q;
    a.method = args->method;
    a.beta = args->beta;
    a.epsilon = args->epsilon;
    a.slab = args->slab;
    a.action_out = args->action_out;
    a.reward = args->reward;
    a.info = args->info;
    a.obs = args->obs;
    a.obs_dtype = args->obs_dtype;
    a.r_state = args->ring_state;
    a.r_action = args->ring_action;
    a.r_reward = args->ring_reward;
    a.ring_steps = args->ring_steps;
    a.ring_step = args->ring_step;
    a.ring_slot = args->ring_steps > 0 ? args->ring_step % args->ring_steps : 0;
    a.gamma = args->gamma;
    a.log10_reward = args->log10_reward;
    a.stats = reinterpret_cast<unsigned long long *>(args->stats_rows);
    int r = PPAD_OK;
    const int minb = args->ctas_per_sm > 0 ? args->ctas_per_sm : 4;
    PPAD_DISPATCH(ctx, (r = launch_policy_g<G>(a, minb, (cudaStream_t)stream)));
    return r;
}

int64_t ppad_policy_stats_rows(int64_t n) { return n < 0 ? -1 : (n + BLOCK - 1) / BLOCK; }

int ppad_policy_stats(const ppad_ctx *ctx, int64_t n, const uint64_t *stats_rows, uint64_t *out, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (!stats_rows || !out) return fail(PPAD_ERR_INVALID, "ppad_policy_stats: null pointer");
    const int64_t n_rows = ppad_policy_stats_rows(n);
    if (int r = cuda_check(cudaMemsetAsync(out, 0, STATS_WORDS * sizeof(uint64_t), (cudaStream_t)stream), "ppad_policy_stats")) return r;
    int64_t grid = (n_rows + REDUCE_ROWS_PER_PASS * 8 - 1) / (REDUCE_ROWS_PER_PASS * 8);   // >= 8 passes per CTA
    grid = grid < 1 ? 1 : (grid > 2 * ctx->sm_count ? 2 * ctx->sm_count : grid);
    stats_reduce_kernel<<<(unsigned)grid, BLOCK, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long *>(stats_rows),
                                                                           n_rows, reinterpret_cast<unsigned long long *>(out));
    return after_launch("stats_reduce_kernel");
}

}   // extern "C"
