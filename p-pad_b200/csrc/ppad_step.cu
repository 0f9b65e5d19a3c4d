// ppad_step.cu -- C ABI of the env step: ppad_step, ppad_rollout(_record) and the host-buffer pipe ppad_pipe_*.
// The kernels live in ppad_step.cuh and are instantiated per geometry in ppad_step_inst.cu.
#include "ppad_step.cuh"
#include <cstdlib>
#include <vector>

namespace ppad {

PPAD_STEP_LAUNCHERS(extern, G56)
PPAD_STEP_LAUNCHERS(extern, G67)
PPAD_STEP_LAUNCHERS(extern, G45)
PPAD_STEP_LAUNCHERS(extern, G56W)
PPAD_STEP_LAUNCHERS(extern, G67W)
PPAD_STEP_LAUNCHERS(extern, G45W)
PPAD_STEP_LAUNCHERS(extern, GDYN)
PPAD_STEP_LAUNCHERS(extern, GDYNW)

int pool_occupancy(int geom)
{
    ppad_ctx probe;
    probe.geom = geom;
    int ctas = 0;
    PPAD_DISPATCH(&probe, (ctas = pool_occupancy_g<G>()));
    return ctas;
}

static int dispatch_step(const ppad_ctx *ctx, const StepKernelArgs &a, int choice, cudaStream_t st)
{
    int r = PPAD_OK;
    PPAD_DISPATCH(ctx, (r = launch_step_g<G>(ctx, a, choice, st)));
    return r;
}

static int dispatch_rollout(const ppad_ctx *ctx, const RolloutArgs &a, bool record, cudaStream_t st)
{
    int r = PPAD_OK;
    PPAD_DISPATCH(ctx, (r = launch_rollout_g<G>(a, record, st)));
    return r;
}

}   // namespace ppad

using namespace ppad;

extern "C" {

struct CompactTarget {
    uint32_t *nibbles, *tile_line, *summary, *counters;
    void *records;
    int64_t capacity_bytes;
    const uint4 *dict;     // result dictionary (device hash table) or NULL
    uint32_t dict_mask;
};

static int launch_step(const ppad_ctx *ctx, const ppad_step_args *args, cudaStream_t st, bool coalesced_results = false,
                       bool stage_host_slab = false, const CompactTarget *compact = nullptr)
{
    if (!args->state) return fail(PPAD_ERR_INVALID, "ppad_step: null state");
    const bool six = args->slab_format == PPAD_SLAB_CHUNKS6 || args->slab_format == PPAD_SLAB_CHUNKS6T;
    const bool chunked = args->slab_format == PPAD_SLAB_CHUNKS10 || six;
    if (args->slab_format != PPAD_SLAB_NIBBLES && args->slab_format != PPAD_SLAB_PLANES3 && !chunked)
        return fail(PPAD_ERR_INVALID, "ppad_step: unknown slab_format");
    if (args->slab && chunked) {
        const int per = six ? 6 : 10;
        if (ctx->geom >= 3) return fail(PPAD_ERR_UNSUPPORTED, "ppad_step: chunk-major slabs need a context with orb_bits = 3");
        if (args->slab_words <= 0 || args->slab_words > 6000 || args->n_inj < 0 || args->n_inj > args->slab_words * per ||
            (args->slab_format != PPAD_SLAB_CHUNKS6T && (int64_t)args->slab_stride < args->n))
            return fail(PPAD_ERR_INVALID, "ppad_step: chunk-major slabs need 0 <= n_inj <= orbs per chunk * slab_words and slab_stride >= n");
        if (args->slab_format == PPAD_SLAB_CHUNKS6T && ((uintptr_t)args->slab & 15))
            return fail(PPAD_ERR_INVALID, "ppad_step: PPAD_SLAB_CHUNKS6T slab must be 16-byte aligned");
        if (args->paths || args->combo_log) return fail(PPAD_ERR_UNSUPPORTED, "ppad_step: chunk-major slabs with finger paths / combo log");
        if (six && ((uintptr_t)args->slab & 1))
            return fail(PPAD_ERR_INVALID, "ppad_step: PPAD_SLAB_CHUNKS6 slab must be 2-byte aligned");
    } else if (args->slab && args->slab_format == PPAD_SLAB_PLANES3) {
        if (ctx->geom >= 3) return fail(PPAD_ERR_UNSUPPORTED, "ppad_step: PPAD_SLAB_PLANES3 needs a context with orb_bits = 3");
        if (args->slab_words <= 0 || args->slab_words % 3 || args->n_inj < 0 || args->n_inj > args->slab_words / 3 * 32)
            return fail(PPAD_ERR_INVALID, "ppad_step: PPAD_SLAB_PLANES3 needs slab_words = 3 * groups and 0 <= n_inj <= 32 * groups");
    } else if (args->slab && (args->slab_words <= 0 || args->n_inj < 0 || args->n_inj > args->slab_words * 8))
        return fail(PPAD_ERR_INVALID, "ppad_step: need 0 <= n_inj <= 8 * slab_words");
    if (args->combo_log && args->log_cap <= 0) return fail(PPAD_ERR_INVALID, "ppad_step: combo_log needs log_cap > 0");
    if (args->obs && args->obs_dtype != PPAD_OBS_F32 && args->obs_dtype != PPAD_OBS_U8)
        return fail(PPAD_ERR_UNSUPPORTED, "ppad_step: obs_dtype must be PPAD_OBS_F32 or PPAD_OBS_U8");
    if (args->obs && ((uintptr_t)args->obs & 15)) return fail(PPAD_ERR_INVALID, "ppad_step: obs must be 16-byte aligned");
    if (((uintptr_t)args->state & 15) || (args->final_state && ((uintptr_t)args->final_state & 15)))
        return fail(PPAD_ERR_INVALID, "ppad_step: state must be 16-byte aligned");
    StepKernelArgs a;
    a.p = ctx->base;
    a.p.key.step = args->step;
    a.p.eval_moves = args->eval_moves;
    a.p.auto_reset = args->auto_reset;
    a.p.max_moves = args->max_moves;
    a.p.n_inj = args->slab ? args->n_inj : -1;
    a.p.slab_words = args->slab_words;
    a.p.log_cap = args->log_cap;
    a.p.slab_fmt = args->slab_format;
    a.p.slab_stride = chunked ? (args->slab_format == PPAD_SLAB_CHUNKS6T ? BLOCK : args->slab_stride) : 0;
    a.p.slab_eager = 0;
    a.p.slab_estride = args->slab_format == PPAD_SLAB_CHUNKS6T ? BLOCK : 1;
    if (args->slab && chunked) {
        const int most = six ? 8 : 4;   // the eager chunks of a board wait in 16 bytes
        if (args->slab_eager < 0 || args->slab_eager > most)
            return fail(PPAD_ERR_INVALID, "ppad_step: slab_eager must be 0 (default) or 1..4 (PPAD_SLAB_CHUNKS6: 1..8)");
        a.p.slab_eager = args->slab_eager ? args->slab_eager : (six ? 3 : 2);
        if (a.p.slab_eager > args->slab_words) a.p.slab_eager = args->slab_words;
    }
    if (args->skyfall_damage >= 0) a.p.skyfall_damage = args->skyfall_damage;
    a.n = args->n;
    a.env0 = args->env0;
    a.state = args->state;
    a.actions = args->actions;
    a.slab = args->slab;
    a.action_out = args->action_out;
    a.reward = args->reward;
    a.info = args->info;
    a.final_state = args->final_state;
    a.state_copy = args->state_copy;
    a.combo_log = args->combo_log;
    a.consumed16 = args->consumed_out;
    a.obs = args->obs;
    a.obs_dtype = args->obs_dtype;
    a.paths = args->paths;
    a.path_lens = args->path_lens;
    a.path_words = args->path_words;
    a.path_len = args->path_len;
    if (a.paths && (a.path_words <= 0 || a.path_len < 0))
        return fail(PPAD_ERR_INVALID, "ppad_step: paths need path_words > 0 and path_len >= 0");
    a.stage_slab = 0;
    if (stage_host_slab && a.slab && ((uintptr_t)a.slab & 15) == 0) {
        if (a.p.slab_fmt == PPAD_SLAB_NIBBLES && a.p.slab_words == 4) a.stage_slab = 1;
        if (a.p.slab_fmt == PPAD_SLAB_PLANES3 && a.p.slab_words == 3) a.stage_slab = 2;
    }
    if (a.slab && chunked) {
        a.stage_slab = args->slab_format == PPAD_SLAB_CHUNKS6T ? 4 : 3;
        if (a.stage_slab == 4) a.p.slab_fmt = PPAD_SLAB_CHUNKS6;   // to the per-board code: uint16 chunks, 256 apart
    }   // the first chunks up front, the rest on demand
    a.c_nibbles = a.c_tile_line = a.c_summary = a.c_counters = nullptr;
    a.c_records = nullptr;
    a.c_capacity = 0;
    a.c_dict = nullptr;
    a.c_dict_mask = 0;
    if (compact) {
        if (a.paths || a.combo_log) return fail(PPAD_ERR_UNSUPPORTED, "compact results: no finger paths / combo log");
        a.c_nibbles = compact->nibbles;
        a.c_tile_line = compact->tile_line;
        a.c_summary = compact->summary;
        a.c_counters = compact->counters;
        a.c_records = reinterpret_cast<uint4 *>(compact->records);
        const int64_t lines = compact->capacity_bytes / 128;
        a.c_capacity = (uint32_t)(lines > 0x7FFFFFFF ? 0x7FFFFFFF : lines);
        a.c_dict = compact->dict;
        a.c_dict_mask = compact->dict_mask;
    }
    a.tile_counter = nullptr;
    int choice = ctx->pool_kind == 1 ? STEP_KERNEL_WPOOL : STEP_KERNEL_POOL;
    if (args->obs) choice = STEP_KERNEL_OBS;
    else if (compact || coalesced_results || a.stage_slab >= 3 || (!a.p.eval_moves && !a.actions && !a.paths) || !a.p.skyfall_damage) {
        // results go straight to host memory: the pooled kernel's scattered 8-byte result writes would become tiny
        // PCIe transactions, one thread per board writes them as full lines.
        // Steps of the in-kernel random policy that only evaluate passes (eval_moves = 0, no actions given: plain
        // moves until the episode length cap) take this kernel, too: a warp whose boards all just move leaves early,
        // and there is nothing for the pool to pack (episode mode with one launch per step: 4.4e10 against 3.7e10
        // env.step calls/s).  Given actions may all be passes (calculate_reward of a whole slab): those stay pooled.
        // Without skyfall (skyfall_damage = 0: the look-ahead's and smart data's evaluations) a board is cancelled
        // once and never cascades, so there is nothing to pool either: 0.027 against 0.033 ms per 2^20 evaluations.
        choice = STEP_KERNEL_THREAD;
    }
    if (choice == STEP_KERNEL_WPOOL) {
        // a pair of counters per launch, handed out round robin: launches of one context may overlap on several streams
        // (256 pairs: a pair is reused only after 255 later launches were enqueued)
        a.tile_counter = ctx->tile_counters + 2 * (ctx->next_counter->fetch_add(1, std::memory_order_relaxed) % TILE_COUNTERS);
    }
    return dispatch_step(ctx, a, choice, st);
}

int ppad_step(const ppad_ctx *ctx, const ppad_step_args *args, void *stream)
{
    if (!args) return fail(PPAD_ERR_INVALID, "ppad_step: null args");
    PPAD_ENTER(ctx, args->n);
    if (args->n == 0) return PPAD_OK;
    return launch_step(ctx, args, (cudaStream_t)stream);
}

int ppad_rollout(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step0, int32_t steps, int32_t eval_moves,
                 int32_t auto_reset, int32_t max_moves, void *state, double *reward_sum, uint32_t *episodes,
                 uint32_t *combos_sum, uint8_t *flags_or, void *stream)
{
    return ppad_rollout_record(ctx, n, env0, step0, steps, eval_moves, auto_reset, max_moves, state, reward_sum, episodes,
                               combos_sum, flags_or, nullptr, nullptr, stream);
}

int ppad_rollout_record(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step0, int32_t steps, int32_t eval_moves,
                        int32_t auto_reset, int32_t max_moves, void *state, double *reward_sum, uint32_t *episodes,
                        uint32_t *combos_sum, uint8_t *flags_or, void *traj_state, uint8_t *traj_action, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0 || steps == 0) return PPAD_OK;
    if (!state || steps < 0) return fail(PPAD_ERR_INVALID, "ppad_rollout: bad argument");
    RolloutArgs a;
    a.p = ctx->base;
    a.p.key.step = step0;
    a.p.eval_moves = eval_moves;
    a.p.auto_reset = auto_reset;
    a.p.max_moves = max_moves;
    a.p.n_inj = -1;
    a.n = n;
    a.env0 = env0;
    a.steps = steps;
    a.state = state;
    a.reward_sum = reward_sum;
    a.episodes = episodes;
    a.combos_sum = combos_sum;
    a.flags_or = flags_or;
    a.traj_state = traj_state;
    a.traj_action = traj_action;
    if (traj_state && ((uintptr_t)traj_state & 15)) return fail(PPAD_ERR_INVALID, "ppad_rollout_record: traj_state must be 16-byte aligned");
    return dispatch_rollout(ctx, a, traj_state || traj_action, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// host-buffer pipeline: the env step for callers whose inputs and results live in HOST memory
// ------------------------------------------------------------------------------------------------
struct ppad_pipe {
    const ppad_ctx *ctx;
    int64_t n_max, chunk;
    int chunks, slab_words_max;
    cudaStream_t streams[16];
    cudaEvent_t done[2][16], fork;   // done[slot][chunk]: two steps may be in flight, one per host buffer set
    int zero_copy;   // 1: pinned host buffers are handed to the kernel directly (UVA), no staging copies
    int last_mode, last_slot;   // 0 none, 1 zero copy, 2 chunked: a mode switch must not reorder steps
    uint32_t *d_slab;
    uint8_t *d_actions, *d_action_out;
    double *d_reward;
    uint32_t *d_info;
    uint32_t *d_counters;   // compact result stream: {next free line, finished tiles}, kept at 0 between launches
    uint4 *d_dict;          // result dictionary: open-addressed hash table (ppad_step.cuh: dict_lookup) or NULL
    uint32_t dict_slots;    // power of two
    int32_t dict_count;     // entries placed
    cudaEvent_t joined;     // recorded after the last chunk of a step: after_stream waits for it
};

// NULL counts as pinned (absent buffer); anything cudaHostAlloc'ed / cudaHostRegister'ed is device-visible under UVA
static bool host_is_pinned(const void *ptr)
{
    if (!ptr) return true;
    // profiling aid (profiles/e2e_placement_probe.py): lets a probe put single buffers in device memory to see which of
    // the PCIe streams of a zero-copy step costs what; never set in production
    static const bool any_memory = std::getenv("PPAD_PIPE_ANY_MEMORY") != nullptr;
    if (any_memory) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

int ppad_pipe_create(const ppad_ctx *ctx, int64_t n_max, int32_t slab_words_max, int32_t chunks, ppad_pipe **out)
{
    if (!out) return fail(PPAD_ERR_INVALID, "ppad_pipe_create: null out");
    *out = nullptr;
    PPAD_ENTER(ctx, n_max);
    if (chunks <= 0) chunks = 4;
    if (chunks > 16) chunks = 16;
    if (slab_words_max < 0) return fail(PPAD_ERR_INVALID, "ppad_pipe_create: slab_words_max < 0");
    ppad_pipe *p = new (std::nothrow) ppad_pipe;
    if (!p) return fail(PPAD_ERR_INVALID, "out of memory");
    std::memset(p, 0, sizeof(*p));
    p->ctx = ctx;
    p->n_max = n_max;
    p->chunks = chunks;
    p->slab_words_max = slab_words_max;
    p->zero_copy = 1;
    p->chunk = ((n_max + chunks - 1) / chunks + BLOCK - 1) / BLOCK * BLOCK;   // CTA-aligned: obs stays 16-byte aligned
    if (p->chunk == 0) p->chunk = BLOCK;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < chunks && e == cudaSuccess; k++) {
        e = cudaStreamCreateWithFlags(&p->streams[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done[0][k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done[1][k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->fork, cudaEventDisableTiming);
    const size_t n = (size_t)(n_max > 0 ? n_max : 1);
    if (e == cudaSuccess && slab_words_max > 0) e = cudaMalloc(&p->d_slab, n * slab_words_max * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_actions, n);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_action_out, n);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_reward, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_info, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_counters, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(p->d_counters, 0, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->joined, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        ppad_pipe_destroy(p);
        return cuda_check(e, "ppad_pipe_create");
    }
    *out = p;
    return PPAD_OK;
}

void ppad_pipe_destroy(ppad_pipe *p)
{
    if (!p) return;
    int prev_device = -1;
    cudaGetDevice(&prev_device);
    cudaSetDevice(p->ctx->device);
    for (int k = 0; k < p->chunks; k++) {
        if (p->streams[k]) {
            cudaStreamSynchronize(p->streams[k]);
            cudaStreamDestroy(p->streams[k]);
        }
        if (p->done[0][k]) cudaEventDestroy(p->done[0][k]);
        if (p->done[1][k]) cudaEventDestroy(p->done[1][k]);
    }
    if (p->fork) cudaEventDestroy(p->fork);
    cudaFree(p->d_slab);
    cudaFree(p->d_actions);
    cudaFree(p->d_action_out);
    cudaFree(p->d_reward);
    cudaFree(p->d_info);
    cudaFree(p->d_counters);
    cudaFree(p->d_dict);
    if (p->joined) cudaEventDestroy(p->joined);
    if (prev_device >= 0) cudaSetDevice(prev_device);   // destroy may run from a finaliser: leave the caller's device alone
    delete p;
}

int ppad_pipe_step(ppad_pipe *p, const ppad_host_step_args *h, void *after_stream, int32_t slot)
{
    if (!p || !h) return fail(PPAD_ERR_INVALID, "ppad_pipe_step: null argument");
    if (slot < 0 || slot > 1) return fail(PPAD_ERR_INVALID, "ppad_pipe_step: slot must be 0 or 1");
    const ppad_ctx *ctx = p->ctx;
    PPAD_ENTER(ctx, h->n);
    if (h->n > p->n_max) return fail(PPAD_ERR_INVALID, "ppad_pipe_step: n exceeds the pipe's n_max");
    if (h->slab_host && (h->slab_words <= 0 || h->slab_words > p->slab_words_max))
        return fail(PPAD_ERR_INVALID, "ppad_pipe_step: slab_words exceeds the pipe's slab_words_max");
    if (!h->state) return fail(PPAD_ERR_INVALID, "ppad_pipe_step: null state");
    if (h->n == 0) return PPAD_OK;
    if (h->results_compact) {
        if (!h->nibbles_host || !h->tile_line_host || !h->records_host || !h->summary_host || h->records_capacity < 128)
            return fail(PPAD_ERR_INVALID, "ppad_pipe_step: compact results need nibbles, tile_line, records and summary buffers");
        if (((uintptr_t)h->records_host & 127) || ((uintptr_t)h->nibbles_host & 127))
            return fail(PPAD_ERR_INVALID, "ppad_pipe_step: records_host and nibbles_host must be 128-byte aligned");
        if (!p->zero_copy || p->zero_copy == 2 || !host_is_pinned(h->nibbles_host) || !host_is_pinned(h->tile_line_host) ||
            !host_is_pinned(h->records_host) || !host_is_pinned(h->summary_host) || !host_is_pinned(h->slab_host) ||
            !host_is_pinned(h->actions_host) || !host_is_pinned(h->state_out_host))
            return fail(PPAD_ERR_UNSUPPORTED, "ppad_pipe_step: compact results need pinned host buffers and zero-copy mode 1");
    }
    const size_t rec = (size_t)ppad_state_bytes_ex(ctx->cfg.rows, ctx->cfg.cols, ctx->cfg.orb_bits);
    const size_t obs_elem = h->obs_dtype == PPAD_OBS_F32 ? 4 : 1;
    const size_t obs_row = (size_t)ctx->cfg.rows * ctx->cfg.cols * 7 * obs_elem;
    // chunk streams start after whatever the caller has queued on after_stream (e.g. the reset)
    cudaError_t e = cudaEventRecord(p->fork, (cudaStream_t)after_stream);
    const bool zc = h->results_compact ||
                    (p->zero_copy && host_is_pinned(h->actions_host) && host_is_pinned(h->slab_host) &&
                     host_is_pinned(h->action_out_host) && host_is_pinned(h->reward_host) && host_is_pinned(h->info_host) &&
                     host_is_pinned(h->state_out_host));
    // zero_copy 1: the kernel reads the slab from and writes its results to the host buffers; 2 ("hybrid"): the slab
    // (and given actions) travel by DMA, chunk by chunk on the pipe's streams, and only the results are written in place
    // by the kernels -- no PCIe read sits inside a kernel, which matters on hosts whose read latency is high
    const bool hybrid = zc && p->zero_copy == 2;
    const int mode = zc && !hybrid ? 1 : 2;
    if (h->slab_host && h->slab_format >= PPAD_SLAB_CHUNKS10 && mode != 1)
        return fail(PPAD_ERR_UNSUPPORTED, "ppad_pipe_step: chunk-major slabs need pinned host buffers and zero-copy mode 1");
    if (p->last_mode && p->last_mode != mode)   // stream <-> board mapping changes: order behind the previous step
        for (int k = 0; k < p->chunks && e == cudaSuccess; k++)
            for (int j = 0; j < p->chunks && e == cudaSuccess; j++)
                e = cudaStreamWaitEvent(p->streams[k], p->done[p->last_slot][j], 0);
    p->last_mode = mode;
    p->last_slot = slot;
    if (e == cudaSuccess && zc && !hybrid) {
        // Zero copy: with unified addressing the kernel reads the skyfall slab from, and writes reward / info /
        // action / the state copy to, the pinned host buffers themselves.  The PCIe traffic then overlaps the
        // kernel store by store instead of chunk by chunk (measured 0.74 ms against 0.87 ms per 2^20-board step).
        cudaStream_t st = p->streams[0];
        e = cudaStreamWaitEvent(st, p->fork, 0);
        if (e != cudaSuccess) return cuda_check(e, "ppad_pipe_step");
        ppad_step_args a;
        std::memset(&a, 0, sizeof(a));
        a.n = h->n;
        a.env0 = h->env0;
        a.step = h->step;
        a.eval_moves = h->eval_moves;
        a.auto_reset = h->auto_reset;
        a.max_moves = h->max_moves;
        a.skyfall_damage = h->skyfall_damage;
        a.state = h->state;
        a.actions = h->actions_host;
        a.slab = h->slab_host;
        a.slab_words = h->slab_words;
        a.n_inj = h->n_inj;
        a.slab_format = h->slab_format;
        a.slab_stride = h->slab_stride;
        a.slab_eager = h->slab_eager;
        a.state_copy = h->state_out_host;
        a.obs = h->obs;
        a.obs_dtype = h->obs_dtype;
        CompactTarget ct;
        if (h->results_compact) {
            ct.nibbles = h->nibbles_host;
            ct.tile_line = h->tile_line_host;
            ct.summary = h->summary_host;
            ct.counters = p->d_counters;
            ct.records = h->records_host;
            ct.capacity_bytes = h->records_capacity;
            ct.dict = nullptr;
            ct.dict_mask = 0;
            if (h->results_compact == 2) {
                if (!p->d_dict) return fail(PPAD_ERR_INVALID, "ppad_pipe_step: results_compact = 2 needs ppad_pipe_set_dictionary first");
                ct.dict = p->d_dict;
                ct.dict_mask = p->dict_slots - 1;
            }
        } else {
            a.action_out = h->action_out_host;
            a.reward = h->reward_host;
            a.info = h->info_host;
        }
        if (int r = launch_step(ctx, &a, st, true, true, h->results_compact ? &ct : nullptr)) return r;
        for (int k = 0; k < p->chunks && e == cudaSuccess; k++) e = cudaEventRecord(p->done[slot][k], st);
        // the caller's stream sees the state this step leaves behind (a later step / encode / reset on it is ordered)
        if (e == cudaSuccess && !h->no_join) e = cudaStreamWaitEvent((cudaStream_t)after_stream, p->done[slot][0], 0);
        return cuda_check(e, "ppad_pipe_step");
    }
    for (int k = 0; k < p->chunks && e == cudaSuccess; k++) {
        const int64_t lo = (int64_t)k * p->chunk;
        if (lo >= h->n) break;
        const int64_t cnt = h->n - lo < p->chunk ? h->n - lo : p->chunk;
        cudaStream_t st = p->streams[k];
        e = cudaStreamWaitEvent(st, p->fork, 0);
        if (e != cudaSuccess) break;
        ppad_step_args a;
        std::memset(&a, 0, sizeof(a));
        a.n = cnt;
        a.env0 = h->env0 + (uint64_t)lo;
        a.step = h->step;
        a.eval_moves = h->eval_moves;
        a.auto_reset = h->auto_reset;
        a.max_moves = h->max_moves;
        a.skyfall_damage = h->skyfall_damage;
        a.state = (char *)h->state + (size_t)lo * rec;
        if (h->actions_host) {
            e = cudaMemcpyAsync(p->d_actions + lo, h->actions_host + lo, (size_t)cnt, cudaMemcpyHostToDevice, st);
            a.actions = p->d_actions + lo;
        }
        if (e == cudaSuccess && h->slab_host) {
            const size_t w = (size_t)h->slab_words;
            e = cudaMemcpyAsync(p->d_slab + (size_t)lo * w, h->slab_host + (size_t)lo * w, (size_t)cnt * w * 4,
                                cudaMemcpyHostToDevice, st);
            a.slab = p->d_slab + (size_t)lo * w;
            a.slab_words = h->slab_words;
            a.n_inj = h->n_inj;
            a.slab_format = h->slab_format;
        }
        if (e != cudaSuccess) break;
        if (hybrid) {
            a.action_out = h->action_out_host ? h->action_out_host + lo : nullptr;
            a.reward = h->reward_host ? h->reward_host + lo : nullptr;
            a.info = h->info_host ? h->info_host + lo : nullptr;
            a.state_copy = h->state_out_host ? (char *)h->state_out_host + (size_t)lo * rec : nullptr;
        } else {
            a.action_out = h->action_out_host ? p->d_action_out + lo : nullptr;
            a.reward = h->reward_host ? p->d_reward + lo : nullptr;
            a.info = h->info_host ? p->d_info + lo : nullptr;
        }
        if (h->obs) {
            a.obs = (char *)h->obs + (size_t)lo * obs_row;
            a.obs_dtype = h->obs_dtype;
        }
        if (int r = launch_step(ctx, &a, st, hybrid)) return r;
        if (hybrid) {
            e = cudaEventRecord(p->done[slot][k], st);
            continue;
        }
        if (h->reward_host)
            e = cudaMemcpyAsync(h->reward_host + lo, p->d_reward + lo, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && h->info_host)
            e = cudaMemcpyAsync(h->info_host + lo, p->d_info + lo, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && h->action_out_host)
            e = cudaMemcpyAsync(h->action_out_host + lo, p->d_action_out + lo, (size_t)cnt, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && h->state_out_host)
            e = cudaMemcpyAsync((char *)h->state_out_host + (size_t)lo * rec, (char *)h->state + (size_t)lo * rec,
                                (size_t)cnt * rec, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(p->done[slot][k], st);
    }
    for (int k = 0; k < p->chunks && e == cudaSuccess && !h->no_join; k++) {
        if ((int64_t)k * p->chunk >= h->n) break;
        e = cudaStreamWaitEvent((cudaStream_t)after_stream, p->done[slot][k], 0);
    }
    return cuda_check(e, "ppad_pipe_step");
}

int64_t ppad_compact_capacity(const ppad_ctx *ctx, int64_t n)
{
    if (!ctx || n < 0) return -1;
    const int64_t tiles = (n + BLOCK - 1) / BLOCK;
    const int64_t rec = ppad_state_bytes_ex(ctx->cfg.rows, ctx->cfg.cols, ctx->cfg.orb_bits);
    // worst case of the dictionary-coded stream: every board has a record, none of them is in the dictionary
    return tiles * (((int64_t)BLOCK * (2 + 12 + rec) + 16 + 127) / 128 * 128);
}

int ppad_pipe_set_dictionary(ppad_pipe *p, const double *rewards, const uint32_t *infos, int32_t count)
{
    if (!p || count < 0 || (count > 0 && (!rewards || !infos)))
        return fail(PPAD_ERR_INVALID, "ppad_pipe_set_dictionary: bad argument");
    if (count > PPAD_DICT_MAX) return fail(PPAD_ERR_INVALID, "ppad_pipe_set_dictionary: at most PPAD_DICT_MAX entries");
    uint32_t slots = 1024;
    while (slots < 4u * (uint32_t)count) slots <<= 1;      // load factor <= 1/4: short probe sequences
    std::vector<uint4> table(slots, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
    int placed = 0;
    for (int32_t i = 0; i < count; i++) {
        uint64_t bits;
        std::memcpy(&bits, &rewards[i], 8);
        const uint32_t lo = (uint32_t)bits, hi = (uint32_t)(bits >> 32);
        uint32_t h = dict_hash(lo, hi, infos[i]);
        for (int probe = 0; probe < DICT_PROBES; probe++, h++) {
            uint4 &slot = table[h & (slots - 1)];
            if (slot.w == 0xFFFFFFFFu) {
                slot = make_uint4(lo, hi, infos[i], (uint32_t)i);   // the code is the caller's index, placed or not
                placed++;
                break;
            }
            if (slot.x == lo && slot.y == hi && slot.z == infos[i]) break;   // duplicate entry: the first index wins
        }
    }
    DeviceGuard guard(p->ctx->device);
    // steps already enqueued may still read the old table
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < p->chunks && e == cudaSuccess; k++) e = cudaStreamSynchronize(p->streams[k]);
    if (e == cudaSuccess && p->d_dict) {
        e = cudaFree(p->d_dict);
        p->d_dict = nullptr;
    }
    p->dict_slots = 0;
    p->dict_count = 0;
    if (e == cudaSuccess && count > 0) {
        e = cudaMalloc(&p->d_dict, (size_t)slots * sizeof(uint4));
        if (e == cudaSuccess) e = cudaMemcpy(p->d_dict, table.data(), (size_t)slots * sizeof(uint4), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            p->dict_slots = slots;
            p->dict_count = placed;
        }
    }
    return cuda_check(e, "ppad_pipe_set_dictionary");
}

int ppad_pipe_set_zero_copy(ppad_pipe *p, int32_t enabled)
{
    if (!p) return fail(PPAD_ERR_INVALID, "ppad_pipe_set_zero_copy: null pipe");
    p->zero_copy = enabled == 2 ? 2 : (enabled ? 1 : 0);
    return PPAD_OK;
}

int ppad_pipe_wait(ppad_pipe *p, int32_t slot)
{
    if (!p || slot < 0 || slot > 1) return fail(PPAD_ERR_INVALID, "ppad_pipe_wait: bad argument");
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < p->chunks && e == cudaSuccess; k++) e = cudaEventSynchronize(p->done[slot][k]);
    return cuda_check(e, "ppad_pipe_wait");
}

int ppad_pipe_join(ppad_pipe *p, void *stream, int32_t slot)
{
    if (!p || slot < 0 || slot > 1) return fail(PPAD_ERR_INVALID, "ppad_pipe_join: bad argument");
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < p->chunks && e == cudaSuccess; k++)
        e = cudaStreamWaitEvent((cudaStream_t)stream, p->done[slot][k], 0);
    return cuda_check(e, "ppad_pipe_join");
}


}   // extern "C"
