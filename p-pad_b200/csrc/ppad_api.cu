// ppad_api.cu -- context management and the single-stage kernels behind the reference's individual methods
// (pack / unpack, reset, encode_obs, cancel, legal mask, sample, drop, fill, damage, permute) with their C ABI.
#include "ppad_device.cuh"

namespace ppad {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int pool_occupancy(int geom);   // ppad_step.cu

template <class G>
__global__ void __launch_bounds__(BLOCK) encode_kernel(const DynGeomData dg, int64_t n, const void *state, void *obs, int obs_dtype)
{
    geom_enter<G>(dg);
    __shared__ ObsShared obs_sh;
    __shared__ uint32_t obs_stream[G::DYNAMIC ? 1 : (BLOCK / 32) * ObsGeom<G>::WARP_WORDS];
    obs_init_luts(obs_sh);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = i < n;
    State<G> s;
    clear_state<G>(s);
    if (valid) s = StateIO<G>::load(state, i);
    const int64_t warp0 = i - (threadIdx.x & 31);
    if (warp0 < n) {
        const int n_valid = (int)(n - warp0 < 32 ? n - warp0 : 32);
        emit_obs_any<G>(s, valid, n_valid, obs_stream + (threadIdx.x >> 5) * ObsGeom<G>::WARP_WORDS, obs_sh,
                        obs_warp_ptr<G>(obs, warp0, obs_dtype), obs_dtype);
    }
}

struct ResetKernelArgs {
    StepParams p;
    int64_t n;
    uint64_t env0;
    const uint32_t *slab;
    const int32_t *fingers_in;
    void *state;
    uint8_t *flags;
    uint8_t *consumed;
};

template <class G>
__global__ void __launch_bounds__(BLOCK) reset_kernel(const __grid_constant__ ResetKernelArgs a)
{
    geom_enter<G>(a.p.dyn);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = i < a.n;                  // every lane takes part: the rejection sampling is warp-cooperative
    int finger = -1;
    uint32_t finger_flag = 0;
    if (valid && a.fingers_in) {
        // a finger off the board would spill into the neighbouring meta fields and drive shifts past the plane width:
        // clamp it and say so
        int fr = a.fingers_in[2 * i], fc = a.fingers_in[2 * i + 1];
        if ((unsigned)fr >= (unsigned)G::rows() || (unsigned)fc >= (unsigned)G::cols()) {
            finger_flag = FLAG_REJECTED;
            fr = fr < 0 ? 0 : (fr >= G::rows() ? G::rows() - 1 : fr);
            fc = fc < 0 ? 0 : (fc >= G::cols() ? G::cols() - 1 : fc);
        }
        finger = fr * G::cols() + fc;
    }
    State<G> s;
    clear_state<G>(s);
    int consumed = 0;
    const uint32_t flags = reset_boards_warp<G, true>(valid, s, a.p, a.p.key.step, a.env0 + (uint64_t)i,
                                                a.slab && valid ? a.slab + (size_t)i * a.p.slab_words : nullptr,
                                                a.p.n_inj, finger, &consumed);
    if (!valid) return;
    StateIO<G>::store(a.state, i, s.b, s.meta);
    if (a.flags) a.flags[i] = (uint8_t)(flags | finger_flag);
    if (a.consumed) a.consumed[i] = (uint8_t)sat255(consumed);
}

// plain int boards -> state.  Like unpack below, through shared memory: the CTA reads its contiguous input region with
// consecutive threads on consecutive elements and parks one code byte per cell (0x80 = a value that is no orb type);
// then every thread packs its own board from there.
template <class G, class T>
__global__ void __launch_bounds__(BLOCK) pack_kernel(const DynGeomData dg, int64_t n, const T *boards, const T *fingers, const uint8_t *prev,
                                                     const uint16_t *moves, void *state, uint8_t *flags)
{
    geom_enter<G>(dg);
    __shared__ uint8_t codes[BLOCK * G::MAX_CELLS];
    const int64_t tile0 = (int64_t)blockIdx.x * BLOCK;
    const int64_t left = n - tile0;
    const int count = (int)(left < BLOCK ? left : BLOCK) * G::cells();
    const T *in = boards + (size_t)tile0 * G::cells();
    for (int e = threadIdx.x; e < count; e += BLOCK) {
        const long long v = (long long)in[e];
        codes[e] = v == -1 ? (uint8_t)G::EMPTY : ((v >= 0 && v < G::NTYPES) ? (uint8_t)v : (uint8_t)0x80);
    }
    __syncthreads();
    const int64_t i = tile0 + threadIdx.x;
    if (i >= n) return;
    int8_t cells[G::MAX_CELLS];
    bool bad = false;
#pragma unroll
    for (int c = 0; c < G::cells(); c++) {
        const uint8_t v = codes[threadIdx.x * G::cells() + c];
        bad |= v == 0x80;
        cells[c] = v == 0x80 || v == G::EMPTY ? (int8_t)-1 : (int8_t)v;
    }
    Board<G> b;
    pack_cells<G, int8_t>(b, cells);
    long long fr = (long long)fingers[2 * i], fc = (long long)fingers[2 * i + 1];
    uint32_t fl = bad ? (uint32_t)FLAG_UNSUPPORTED_ORB : 0u;
    if (fr < 0 || fr >= G::rows() || fc < 0 || fc >= G::cols()) {   // see reset_kernel: clamp and flag
        fl |= FLAG_REJECTED;
        fr = fr < 0 ? 0 : (fr >= G::rows() ? G::rows() - 1 : fr);
        fc = fc < 0 ? 0 : (fc >= G::cols() ? G::cols() - 1 : fc);
    }
    int pv = prev ? prev[i] : ACT_PASS;
    if (pv > ACT_PASS) pv = ACT_PASS;                     // 3-bit field; anything else reads as 'pass'
    StateIO<G>::store(state, i, b, meta_pack((int)(fr * G::cols() + fc), pv, moves ? moves[i] : 0));
    if (flags) flags[i] = (uint8_t)fl;
}

// state -> plain int boards.  A board is CELLS elements of T (240 B as int64), so one thread per board would write
// 32 scattered runs per warp instruction; instead the CTA parks its 256 boards' planes in shared memory and then
// writes its contiguous output region element by element, consecutive threads -> consecutive addresses.
template <class G, class T>
__global__ void __launch_bounds__(BLOCK) unpack_kernel(const DynGeomData dg, int64_t n, const void *state, T *boards, T *fingers, uint8_t *prev,
                                                       uint16_t *moves)
{
    geom_enter<G>(dg);
    typedef typename G::W W;
    __shared__ W planes[G::NP][BLOCK];
    const int64_t tile0 = (int64_t)blockIdx.x * BLOCK;
    const int64_t i = tile0 + threadIdx.x;
    if (i < n) {
        const State<G> s = StateIO<G>::load(state, i);
        planes[0][threadIdx.x] = s.b.b0;
        planes[1][threadIdx.x] = s.b.b1;
        planes[2][threadIdx.x] = s.b.b2;
        PPAD_P4(planes[3][threadIdx.x] = s.b.b3)
        const int f = meta_finger(s.meta);
        if (fingers) {
            fingers[2 * i] = (T)(f / G::cols());
            fingers[2 * i + 1] = (T)(f % G::cols());
        }
        if (prev) prev[i] = (uint8_t)meta_prev(s.meta);
        if (moves) moves[i] = (uint16_t)meta_moves(s.meta);
    }
    if (!boards) return;
    __syncthreads();
    const int64_t left = n - tile0;
    const int count = (int)(left < BLOCK ? left : BLOCK) * G::cells();
    T *out = boards + (size_t)tile0 * G::cells();
    for (int e = threadIdx.x; e < count; e += BLOCK) {
        const int b = e / G::cells(), c = e - b * G::cells();
        int code = 0;
#pragma unroll
        for (int j = 0; j < G::NP; j++) code |= (int)((planes[j][b] >> c) & 1) << j;
        out[e] = (T)(code == G::EMPTY ? -1 : code);
    }
}

template <class G>
__global__ void __launch_bounds__(BLOCK) cancel_kernel(const DynGeomData dg, int64_t n, void *state, uint8_t *n_combos, uint16_t *combo_log,
                                                       int log_cap)
{
    geom_enter<G>(dg);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    State<G> s = StateIO<G>::load(state, i);
    LogSink sink;
    sink.log = combo_log ? combo_log + (size_t)i * log_cap : nullptr;
    sink.cap = log_cap;
    sink.n = 0;
    const int m = cancel_board<G>(s.b, [&](int colour, int N) { sink.put(colour, N); });
    StateIO<G>::store(state, i, s.b, s.meta);
    if (n_combos) n_combos[i] = (uint8_t)sat255(m);
}

// pad_utils.detect_island (utils.py:86-111): the 4-connected island of cells of `type` around the seed cell.  The
// reference's recursion starts at the seed if it holds `type`, marks it, and enters every neighbour of that type whose
// mark is still 0; cells already marked in `island_in` are therefore never entered (nor expanded), the seed itself is
// expanded even when it is marked.  Result = island_in | cells reached.
template <class G>
__global__ void __launch_bounds__(BLOCK) island_kernel(const DynGeomData dg, int64_t n, const void *state, const uint8_t *seed, const int8_t *type,
                                                       const uint64_t *island_in, uint64_t *island_out)
{
    geom_enter<G>(dg);
    typedef typename G::W W;
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const State<G> s = StateIO<G>::load(state, i);
    const W pre = island_in ? (W)island_in[i] & G::board() : (W)0;
    const int cell = seed[i];
    const int t = type[i] == -1 ? G::EMPTY : type[i];      // -1 = empty cells: the reference floods those, too
    W out = pre;
    if (cell < G::cells() && t >= 0 && code_at<G>(s.b, cell) == t) {
        // er / ed: same code as the right / lower neighbour (empty cells included, unlike match_mask)
        W xr = (s.b.b0 ^ (s.b.b0 >> 1)) | (s.b.b1 ^ (s.b.b1 >> 1)) | (s.b.b2 ^ (s.b.b2 >> 1));
        W xd = (s.b.b0 ^ (s.b.b0 >> G::cols())) | (s.b.b1 ^ (s.b.b1 >> G::cols())) | (s.b.b2 ^ (s.b.b2 >> G::cols()));
        PPAD_P4(xr |= s.b.b3 ^ (s.b.b3 >> 1); xd |= s.b.b3 ^ (s.b.b3 >> G::cols()))
        W er = ~xr & G::not_lastcol(), ed = ~xd & G::not_lastrow();
        const W bit = G::ONE << cell;
        const W open = ~pre | bit;                        // cells the flood may stand on
        er &= open & (open >> 1);
        ed &= open & (open >> G::cols());
        out |= closure<G>(bit, er, ed);
    }
    island_out[i] = (uint64_t)out;
}

template <class G>
__global__ void __launch_bounds__(BLOCK) legal_kernel(const DynGeomData dg, int64_t n, const void *state, uint8_t *mask)
{
    geom_enter<G>(dg);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const State<G> s = StateIO<G>::load(state, i);
    const int f = meta_finger(s.meta);
    const int off = (~legal_mask<G>(f, ACT_PASS)) & 15;   // off-board moves only (utils.py:73-83)
    mask[i] = (uint8_t)(legal_mask<G>(f, meta_prev(s.meta)) | (off << 4));
}

// PlayerAction.sample (player.py:38-66): same draw as the in-kernel policy of step_kernel
template <class G>
__global__ void __launch_bounds__(BLOCK) sample_kernel(const DynGeomData dg, int64_t n, uint64_t env0, RngKey key, const void *state,
                                                       int include_pass, uint8_t *actions)
{
    geom_enter<G>(dg);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const State<G> s = StateIO<G>::load(state, i);
    const uint64_t env = env0 + (uint64_t)i;
    key.env_lo = (uint32_t)env;
    key.env_hi = (uint32_t)(env >> 32);
    int mask = legal_mask<G>(meta_finger(s.meta), meta_prev(s.meta));
    // with include_pass the reference rejects 'pass' right after a pass / reset (player.py:52-53)
    if (include_pass && meta_prev(s.meta) != ACT_PASS) mask |= 16;
    const int a = sample_from_mask(mask, action_word(action_block(key), key.step));
    actions[i] = (uint8_t)(a < 0 ? 255 : a);
}

// PAD.drop (game.py:201-221), in place
template <class G>
__global__ void __launch_bounds__(BLOCK) drop_kernel(const DynGeomData dg, int64_t n, void *state)
{
    geom_enter<G>(dg);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    State<G> s = StateIO<G>::load(state, i);
    drop_board<G>(s.b);
    StateIO<G>::store(state, i, s.b, s.meta);
}

// PAD.fill_board (game.py:232-256), in place; draws start at index p.draw0 of the skyfall stream
template <class G>
__global__ void __launch_bounds__(BLOCK) fill_kernel(const __grid_constant__ ResetKernelArgs a, int reset_all)
{
    geom_enter<G>(a.p.dyn);
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= a.n) return;
    RngKey key = a.p.key;
    const uint64_t env = a.env0 + (uint64_t)i;
    key.env_lo = (uint32_t)env;
    key.env_hi = (uint32_t)(env >> 32);
    OrbSource src;
    src.init(a.slab ? a.slab + (size_t)i * a.p.slab_words : nullptr, a.p.n_inj, STREAM_SKYFALL);
    src.cursor = a.p.draw0;
    State<G> s = StateIO<G>::load(a.state, i);
    if (reset_all) refill_all<G>(s.b, src, key, a.p.sky);
    else fill_board<G>(s.b, src, key, a.p.sky);
    StateIO<G>::store(a.state, i, s.b, s.meta);
    if (a.flags) a.flags[i] = (uint8_t)(src.exhausted ? FLAG_SKYFALL_EXHAUSTED : 0);
    if (a.consumed) a.consumed[i] = (uint8_t)sat255(src.cursor - a.p.draw0);
}

// PAD.damage (game.py:157-199) of explicit combo lists
__global__ void __launch_bounds__(BLOCK) damage_kernel(int64_t n, const __grid_constant__ TeamTable team,
                                                       const uint16_t *combo_log, int log_cap, const uint8_t *n_combos,
                                                       double *reward)
{
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const int m = n_combos[i] < log_cap ? n_combos[i] : log_cap;
    reward[i] = damage_of_log(team, combo_log + (size_t)i * log_cap, m);
}

// permutation_mapping (smart_data.py:94-108): every orb type t of board i becomes maps[i][t]
template <class G>
__global__ void __launch_bounds__(BLOCK) permute_kernel(const DynGeomData dg, int64_t n, void *state, const uint8_t *maps)
{
    geom_enter<G>(dg);
    typedef typename G::W W;
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    State<G> s = StateIO<G>::load(state, i);
    const uint2 m2 = *reinterpret_cast<const uint2 *>(maps + 8 * i);
    const uint64_t lut = (uint64_t)m2.x | ((uint64_t)m2.y << 32);
    // bit-sliced: the cells of type t all move to type maps[t] at once (a plane at a time instead of a cell at a time);
    // empty cells and types 7..9 keep their code
    W p0 = 0, p1 = 0, p2 = 0, p3 = 0, moved = 0;
    auto route = [&](W cells, int t) {
        const int to = (int)((lut >> (8 * t)) & 7);
        moved |= cells;
        p0 |= (to & 1) ? cells : (W)0;
        p1 |= (to & 2) ? cells : (W)0;
        p2 |= (to & 4) ? cells : (W)0;
    };
    route(colour_plane<G, 0>(s.b) & G::board(), 0);
    route(colour_plane<G, 1>(s.b) & G::board(), 1);
    route(colour_plane<G, 2>(s.b) & G::board(), 2);
    route(colour_plane<G, 3>(s.b) & G::board(), 3);
    route(colour_plane<G, 4>(s.b) & G::board(), 4);
    route(colour_plane<G, 5>(s.b) & G::board(), 5);
    route(colour_plane<G, 6>(s.b) & G::board(), 6);
    const W kept = G::board() & ~moved;
    s.b.b0 = p0 | (s.b.b0 & kept);
    s.b.b1 = p1 | (s.b.b1 & kept);
    s.b.b2 = p2 | (s.b.b2 & kept);
    PPAD_P4(s.b.b3 = p3 | (s.b.b3 & kept))
    StateIO<G>::store(state, i, s.b, s.meta);
}

}   // namespace ppad

using namespace ppad;

template <class G, class T>
static void launch_pack(const ppad_ctx *ctx, int64_t n, const void *boards, const void *fingers, const uint8_t *prev, const uint16_t *moves,
                        void *state, uint8_t *flags, cudaStream_t st)
{
    pack_kernel<G, T><<<grid_for(n), BLOCK, 0, st>>>(ctx->base.dyn, n, (const T *)boards, (const T *)fingers, prev, moves, state, flags);
}

template <class G, class T>
static void launch_unpack(const ppad_ctx *ctx, int64_t n, const void *state, void *boards, void *fingers, uint8_t *prev, uint16_t *moves,
                          cudaStream_t st)
{
    unpack_kernel<G, T><<<grid_for(n), BLOCK, 0, st>>>(ctx->base.dyn, n, state, (T *)boards, (T *)fingers, prev, moves);
}

extern "C" {

void ppad_default_config(ppad_config *cfg)
{
    // parameters.py:54-89 (default_team) and :43 (default_skyfall); orb types per parameters.py:20-29
    static const int c1[6] = {4, 3, 1, 2, 0, 1};
    static const int c2[6] = {4, 3, 2, 0, 1, -1};
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->rows = 5;
    cfg->cols = 6;
    cfg->skyfall_damage = 1;
    cfg->cascade_cap = 255;
    cfg->reset_cap = 64;
    cfg->team_size = 6;
    for (int i = 0; i < 6; i++) {
        cfg->team_color_1[i] = c1[i];
        cfg->team_color_2[i] = c2[i];
        cfg->team_atk[i] = 1000.0;
    }
    for (int i = 0; i < 6; i++) cfg->skyfall[i] = 1.0 / 6;
}

int ppad_abi_version(void) { return PPAD_ABI_VERSION; }

const char *ppad_strerror(int status)
{
    switch (status) {
    case PPAD_OK: return "ok";
    case PPAD_ERR_INVALID: return "invalid argument";
    case PPAD_ERR_UNSUPPORTED: return "unsupported configuration";
    case PPAD_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
    }
}

const char *ppad_last_error(void) { return g_last_error.c_str(); }

int64_t ppad_state_bytes_ex(int rows, int cols, int orb_bits)
{
    const bool wide = orb_bits == 4;
    if (orb_bits != 0 && orb_bits != 3 && !wide) return -1;
    if (rows == 5 && cols == 6) return wide ? G56W::STATE_BYTES : G56::STATE_BYTES;
    if (rows == 6 && cols == 7) return wide ? G67W::STATE_BYTES : G67::STATE_BYTES;
    if (rows == 4 && cols == 5) return wide ? G45W::STATE_BYTES : G45::STATE_BYTES;
    if (rows >= 1 && cols >= 1 && rows * cols <= 64) return wide ? GDYNW::STATE_BYTES : GDYN::STATE_BYTES;
    return -1;
}

int64_t ppad_state_bytes(int rows, int cols) { return ppad_state_bytes_ex(rows, cols, 3); }

uint64_t ppad_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int ppad_create(const ppad_config *cfg, int device, ppad_ctx **out)
{
    if (!cfg || !out) return fail(PPAD_ERR_INVALID, "null argument");
    *out = nullptr;
    ppad_ctx *ctx = new (std::nothrow) ppad_ctx;
    if (!ctx) return fail(PPAD_ERR_INVALID, "out of memory");
    std::memset(ctx, 0, sizeof(*ctx));
    ctx->cfg = *cfg;
    ctx->device = device;
    std::string err;
    if (const int r = build_params(*cfg, ctx->base, ctx->geom, err)) {
        delete ctx;
        return fail(r, err);
    }
    int count = 0;
    const cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        delete ctx;
        return fail(PPAD_ERR_CUDA, std::string("no usable CUDA device ") + std::to_string(device) + ": " +
                                       (e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range") +
                                       " (this library has no CPU path)");
    }
    ctx->sm_count = 148;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    {
        DeviceGuard guard(device);   // the occupancy query wants the device current; the caller's device is restored
        ctx->pool_ctas_per_sm = pool_occupancy(ctx->geom);
    }
    if (ctx->pool_ctas_per_sm <= 0) ctx->pool_ctas_per_sm = 1;
    {
        DeviceGuard guard(device);
        if (cudaMalloc(&ctx->tile_counters, 2 * TILE_COUNTERS * sizeof(uint32_t)) != cudaSuccess ||
            cudaMemset(ctx->tile_counters, 0, 2 * TILE_COUNTERS * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            delete ctx;
            return fail(PPAD_ERR_CUDA, "ppad_create: allocating the tile counters failed");
        }
    }
    ctx->next_counter = new std::atomic<uint32_t>(0);
    const char *kind = std::getenv("PPAD_POOL_KIND");
    ctx->pool_kind = kind ? std::atoi(kind) : 1;
    *out = ctx;
    return PPAD_OK;
}

void ppad_destroy(ppad_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->tile_counters) {
        DeviceGuard guard(ctx->device);
        cudaFree(ctx->tile_counters);
    }
    delete ctx->next_counter;
    delete ctx;
}

int ppad_pack(const ppad_ctx *ctx, int64_t n, const void *boards, const void *fingers, int dtype, const uint8_t *prev,
              const uint16_t *moves, void *state, uint8_t *flags, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!boards || !fingers || !state) return fail(PPAD_ERR_INVALID, "ppad_pack: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case PPAD_I8: PPAD_DISPATCH(ctx, (launch_pack<G, int8_t>(ctx, n, boards, fingers, prev, moves, state, flags, st))); break;
    case PPAD_I32: PPAD_DISPATCH(ctx, (launch_pack<G, int32_t>(ctx, n, boards, fingers, prev, moves, state, flags, st))); break;
    case PPAD_I64: PPAD_DISPATCH(ctx, (launch_pack<G, int64_t>(ctx, n, boards, fingers, prev, moves, state, flags, st))); break;
    default: return fail(PPAD_ERR_UNSUPPORTED, "ppad_pack: dtype must be PPAD_I8, PPAD_I32 or PPAD_I64");
    }
    return after_launch("pack_kernel");
}

int ppad_unpack(const ppad_ctx *ctx, int64_t n, const void *state, void *boards, void *fingers, int dtype, uint8_t *prev,
                uint16_t *moves, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state) return fail(PPAD_ERR_INVALID, "ppad_unpack: null state");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case PPAD_I8: PPAD_DISPATCH(ctx, (launch_unpack<G, int8_t>(ctx, n, state, boards, fingers, prev, moves, st))); break;
    case PPAD_I32: PPAD_DISPATCH(ctx, (launch_unpack<G, int32_t>(ctx, n, state, boards, fingers, prev, moves, st))); break;
    case PPAD_I64: PPAD_DISPATCH(ctx, (launch_unpack<G, int64_t>(ctx, n, state, boards, fingers, prev, moves, st))); break;
    default: return fail(PPAD_ERR_UNSUPPORTED, "ppad_unpack: dtype must be PPAD_I8, PPAD_I32 or PPAD_I64");
    }
    return after_launch("unpack_kernel");
}

int ppad_reset(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, const uint32_t *slab, int32_t slab_words,
               int32_t n_inj, const int32_t *fingers_in, void *state, uint8_t *flags, uint8_t *consumed, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state) return fail(PPAD_ERR_INVALID, "ppad_reset: null state");
    if (slab && (slab_words <= 0 || n_inj < 0 || n_inj > slab_words * 8))
        return fail(PPAD_ERR_INVALID, "ppad_reset: need 0 <= n_inj <= 8 * slab_words");
    ResetKernelArgs a;
    a.p = ctx->base;
    a.p.key.step = step;
    a.p.n_inj = slab ? n_inj : -1;
    a.p.slab_words = slab_words;
    a.n = n;
    a.env0 = env0;
    a.slab = slab;
    a.fingers_in = fingers_in;
    a.state = state;
    a.flags = flags;
    a.consumed = consumed;
    PPAD_DISPATCH(ctx, (reset_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(a)));
    return after_launch("reset_kernel");
}

int ppad_encode_obs(const ppad_ctx *ctx, int64_t n, const void *state, void *obs, int obs_dtype, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state || !obs) return fail(PPAD_ERR_INVALID, "ppad_encode_obs: null pointer");
    if (obs_dtype != PPAD_OBS_F32 && obs_dtype != PPAD_OBS_U8)
        return fail(PPAD_ERR_UNSUPPORTED, "ppad_encode_obs: obs_dtype must be PPAD_OBS_F32 or PPAD_OBS_U8");
    if ((uintptr_t)obs & 15) return fail(PPAD_ERR_INVALID, "ppad_encode_obs: obs must be 16-byte aligned");
    PPAD_DISPATCH(ctx, (encode_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(ctx->base.dyn, n, state, obs, obs_dtype)));
    return after_launch("encode_kernel");
}

int ppad_cancel(const ppad_ctx *ctx, int64_t n, void *state, uint8_t *n_combos, uint16_t *combo_log, int32_t log_cap,
                void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state) return fail(PPAD_ERR_INVALID, "ppad_cancel: null state");
    if (combo_log && log_cap <= 0) return fail(PPAD_ERR_INVALID, "ppad_cancel: combo_log needs log_cap > 0");
    PPAD_DISPATCH(ctx, (cancel_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(ctx->base.dyn, n, state, n_combos, combo_log,
                                                                                         log_cap)));
    return after_launch("cancel_kernel");
}

int ppad_island(const ppad_ctx *ctx, int64_t n, const void *state, const uint8_t *seed_cell, const int8_t *orb_type,
                const uint64_t *island_in, uint64_t *island_out, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state || !seed_cell || !orb_type || !island_out) return fail(PPAD_ERR_INVALID, "ppad_island: null pointer");
    PPAD_DISPATCH(ctx, (island_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(ctx->base.dyn, n, state, seed_cell, orb_type,
                                                                                         island_in, island_out)));
    return after_launch("island_kernel");
}

int ppad_legal_mask(const ppad_ctx *ctx, int64_t n, const void *state, uint8_t *mask, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state || !mask) return fail(PPAD_ERR_INVALID, "ppad_legal_mask: null pointer");
    PPAD_DISPATCH(ctx, (legal_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(ctx->base.dyn, n, state, mask)));
    return after_launch("legal_kernel");
}

int ppad_sample_actions(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, const void *state,
                        int32_t include_pass, uint8_t *actions, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state || !actions) return fail(PPAD_ERR_INVALID, "ppad_sample_actions: null pointer");
    RngKey key = ctx->base.key;
    key.step = step;
    PPAD_DISPATCH(ctx, (sample_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(ctx->base.dyn, n, env0, key, state,
                                                                                         include_pass, actions)));
    return after_launch("sample_kernel");
}

int ppad_drop(const ppad_ctx *ctx, int64_t n, void *state, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state) return fail(PPAD_ERR_INVALID, "ppad_drop: null state");
    PPAD_DISPATCH(ctx, (drop_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(ctx->base.dyn, n, state)));
    return after_launch("drop_kernel");
}

int ppad_fill(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, int32_t draw0, int32_t reset_all,
              const uint32_t *slab, int32_t slab_words, int32_t n_inj, void *state, uint8_t *flags, uint8_t *consumed,
              void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state) return fail(PPAD_ERR_INVALID, "ppad_fill: null state");
    if (draw0 < 0) return fail(PPAD_ERR_INVALID, "ppad_fill: draw0 must be >= 0");
    if (slab && (slab_words <= 0 || n_inj < 0 || n_inj > slab_words * 8))
        return fail(PPAD_ERR_INVALID, "ppad_fill: need 0 <= n_inj <= 8 * slab_words");
    ResetKernelArgs a;
    a.p = ctx->base;
    a.p.key.step = step;
    a.p.n_inj = slab ? n_inj : -1;
    a.p.slab_words = slab_words;
    a.p.draw0 = draw0;
    a.n = n;
    a.env0 = env0;
    a.slab = slab;
    a.fingers_in = nullptr;
    a.state = state;
    a.flags = flags;
    a.consumed = consumed;
    PPAD_DISPATCH(ctx, (fill_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(a, reset_all)));
    return after_launch("fill_kernel");
}

int ppad_damage(const ppad_ctx *ctx, int64_t n, const uint16_t *combo_log, int32_t log_cap, const uint8_t *n_combos,
                double *reward, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!combo_log || !n_combos || !reward || log_cap <= 0) return fail(PPAD_ERR_INVALID, "ppad_damage: bad argument");
    damage_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, ctx->base.team, combo_log, log_cap, n_combos, reward);
    return after_launch("damage_kernel");
}

int ppad_permute(const ppad_ctx *ctx, int64_t n, void *state, const uint8_t *maps, void *stream)
{
    PPAD_ENTER(ctx, n);
    if (n == 0) return PPAD_OK;
    if (!state || !maps) return fail(PPAD_ERR_INVALID, "ppad_permute: null pointer");
    if ((uintptr_t)maps & 7) return fail(PPAD_ERR_INVALID, "ppad_permute: maps must be 8-byte aligned");
    PPAD_DISPATCH(ctx, (permute_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(ctx->base.dyn, n, state, maps)));
    return after_launch("permute_kernel");
}

}   // extern "C"
