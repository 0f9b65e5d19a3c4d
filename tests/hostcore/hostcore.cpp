// hostcore.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the per-board device code (p-pad_b200/csrc/ppad_core.cuh, __host__ __device__) with g++
// so that the bit-sliced core can be fuzzed against the CPU oracle in the GPU-less build container
// (tests/test_hostcore_vs_oracle.py).  Nothing in the product loads this library; the shipped
// libppad_b200.so has device code only.
#include <cstdint>
#include <cstring>
#include <string>

#include "../../p-pad_b200/csrc/ppad_core.cuh"
#include "../../p-pad_b200/csrc/ppad_params.h"

using namespace ppad;

// geom: 0 = 5x6, 1 = 6x7, 2 = 4x5 (3-bit orb codes); 3..5 = the same boards with 4-bit codes; 6 / 7 = run-time
// dimensions (GeomDyn: the caller sets dyn_geom() first)
#define HC_DISPATCH(geom, EXPR)                                  \
    do {                                                         \
        if ((geom) == 0) { typedef Geom<5, 6> G; EXPR; }         \
        else if ((geom) == 1) { typedef Geom<6, 7> G; EXPR; }    \
        else if ((geom) == 2) { typedef Geom<4, 5> G; EXPR; }    \
        else if ((geom) == 3) { typedef Geom<5, 6, 4> G; EXPR; } \
        else if ((geom) == 4) { typedef Geom<6, 7, 4> G; EXPR; } \
        else if ((geom) == 5) { typedef Geom<4, 5, 4> G; EXPR; } \
        else if ((geom) == 6) { typedef GeomDyn<3> G; EXPR; }    \
        else { typedef GeomDyn<4> G; EXPR; }                     \
    } while (0)

static int shape_geom(int rows, int cols, int orb_bits)
{
    int g;
    if (rows == 5 && cols == 6) g = 0;
    else if (rows == 6 && cols == 7) g = 1;
    else if (rows == 4 && cols == 5) g = 2;
    else if (rows >= 1 && cols >= 1 && rows * cols <= 64) {
        dyn_geom() = make_dyn_geom(rows, cols);
        return orb_bits == 4 ? 7 : 6;
    } else return -1;
    return orb_bits == 4 ? g + 3 : g;
}

namespace {

template <class G>
void step_batch(const StepParams &base, int64_t n, uint64_t env0, int8_t *boards, int32_t *fingers, uint8_t *prev,
                uint16_t *moves, const uint8_t *actions, const uint32_t *slab, uint8_t *action_out, double *reward,
                uint32_t *info, int8_t *final_boards, uint16_t *combo_log, const uint32_t *paths = nullptr,
                int path_words = 0, const uint16_t *path_lens = nullptr, bool tile_major = false)
{
    for (int64_t i = 0; i < n; i++) {
        State<G> s;
        pack_cells<G, int8_t>(s.b, boards + i * G::cells());
        s.meta = meta_pack(fingers[2 * i] * G::cols() + fingers[2 * i + 1], prev[i], moves ? moves[i] : 0);
        int action = actions ? actions[i] : -1;
        if (action == 255) action = -1;
        Board<G> fin;
        // SLAB_CHUNKS10 is chunk-major: board i starts at slab + i, its chunks lie slab_stride words apart, and the
        // first slab_eager chunks are handed over as a copy (what the step kernel does with the lines it fetches up front)
        // (SLAB_CHUNKS6T, tile-major, is SLAB_CHUNKS6 to the per-board code: see hostcore_step_batch)
        const bool chunked = base.slab_fmt == SLAB_CHUNKS10 || base.slab_fmt == SLAB_CHUNKS6;
        const uint32_t *row = slab ? slab + (chunked ? (size_t)i : (size_t)i * base.slab_words) : nullptr;
        uint32_t eager[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        if (base.slab_fmt == SLAB_CHUNKS6 && slab) {   // uint16 chunks
            const uint16_t *r16 = reinterpret_cast<const uint16_t *>(slab) +
                                  (tile_major ? (size_t)(i / 256) * base.slab_words * 256 + (size_t)(i % 256) : (size_t)i);
            row = reinterpret_cast<const uint32_t *>(r16);
            for (int c = 0; c < base.slab_eager; c++)
                reinterpret_cast<uint16_t *>(eager)[c] = r16[(size_t)c * (size_t)base.slab_stride];
        } else if (chunked && row)
            for (int c = 0; c < base.slab_eager; c++) eager[c] = row[(size_t)c * (size_t)base.slab_stride];
        const StepOut o = step_board<G>(true, s, fin, base, env0 + (uint64_t)i, action, row,
                                        combo_log ? combo_log + (size_t)i * base.log_cap : nullptr,
                                        paths ? paths + (size_t)i * path_words : nullptr, paths ? path_lens[i] : 0, 0,
                                        nullptr, chunked && row ? eager : nullptr);
        unpack_cells<G, int8_t>(s.b, boards + i * G::cells());
        const int f = meta_finger(s.meta);
        fingers[2 * i] = f / G::cols();
        fingers[2 * i + 1] = f % G::cols();
        prev[i] = (uint8_t)meta_prev(s.meta);
        if (moves) moves[i] = (uint16_t)meta_moves(s.meta);
        if (action_out) action_out[i] = (uint8_t)o.action;
        if (reward) reward[i] = o.reward;
        if (info) info[i] = o.info;
        if (final_boards) unpack_cells<G, int8_t>(fin, final_boards + i * G::cells());
    }
}

template <class G>
void reset_batch(const StepParams &base, int64_t n, uint64_t env0, const uint32_t *slab, const int32_t *fingers_in,
                 int8_t *boards, int32_t *fingers, uint8_t *flags, int32_t *consumed)
{
    for (int64_t i = 0; i < n; i++) {
        RngKey key = base.key;
        const uint64_t env = env0 + (uint64_t)i;
        key.env_lo = (uint32_t)env;
        key.env_hi = (uint32_t)(env >> 32);
        OrbSource src;
        src.init(slab ? slab + (size_t)i * base.slab_words : nullptr, base.n_inj, STREAM_RESET);
        State<G> s;
        const int fin = fingers_in ? fingers_in[2 * i] * G::cols() + fingers_in[2 * i + 1] : -1;
        const uint32_t fl = reset_board<G>(s, base.reset_cap, src, key, base.sky, fin);
        unpack_cells<G, int8_t>(s.b, boards + i * G::cells());
        const int f = meta_finger(s.meta);
        fingers[2 * i] = f / G::cols();
        fingers[2 * i + 1] = f % G::cols();
        if (flags) flags[i] = (uint8_t)fl;
        if (consumed) consumed[i] = src.cursor;
    }
}

template <class G>
void cancel_batch(int64_t n, int8_t *boards, uint8_t *n_combos, uint16_t *combo_log, int log_cap)
{
    for (int64_t i = 0; i < n; i++) {
        Board<G> b;
        pack_cells<G, int8_t>(b, boards + i * G::cells());
        LogSink sink;
        sink.log = combo_log ? combo_log + (size_t)i * log_cap : nullptr;
        sink.cap = log_cap;
        sink.n = 0;
        const int m = cancel_board<G>(b, [&](int colour, int N) { sink.put(colour, N); });
        unpack_cells<G, int8_t>(b, boards + i * G::cells());
        if (n_combos) n_combos[i] = (uint8_t)sat255(m);
    }
}

template <class G>
void drop_batch(int64_t n, int8_t *boards)
{
    for (int64_t i = 0; i < n; i++) {
        Board<G> b;
        pack_cells<G, int8_t>(b, boards + i * G::cells());
        drop_board<G>(b);
        unpack_cells<G, int8_t>(b, boards + i * G::cells());
    }
}

template <class G>
void obs_batch(int64_t n, const int8_t *boards, const int32_t *fingers, uint8_t *out)
{
    if constexpr (G::DYNAMIC) {
        (void)n; (void)boards; (void)fingers; (void)out;      // the run-time geometries emit the obs in device code only
    } else {
    constexpr int NB = G::OBS_ELEMS, NW = (NB + 31) / 32 + 1;
    for (int64_t i = 0; i < n; i++) {
        State<G> s;
        pack_cells<G, int8_t>(s.b, boards + i * G::cells());
        s.meta = meta_pack(fingers[2 * i] * G::cols() + fingers[2 * i + 1], 4, 0);
        uint32_t w[NW] = {0};
        obs_mask<G, NW>(s, w);
        for (int e = 0; e < NB; e++) out[i * NB + e] = (w[e >> 5] >> (e & 31)) & 1;
        for (int e = NB; e < 32 * NW; e++)
            if ((w[e >> 5] >> (e & 31)) & 1) out[i * NB] = 99;   // bits beyond the board must stay clear
    }
    }
}

std::string g_err;

}   // namespace

extern "C" {

const char *hostcore_last_error() { return g_err.c_str(); }

int hostcore_step_batch(const ppad_config *cfg, int64_t n, uint64_t env0, uint32_t step, int8_t *boards, int32_t *fingers,
                        uint8_t *prev, uint16_t *moves, const uint8_t *actions, int eval_moves, int auto_reset,
                        int max_moves, const uint32_t *slab, int slab_words, int n_inj, uint8_t *action_out,
                        double *reward, uint32_t *info, int8_t *final_boards, uint16_t *combo_log, int log_cap,
                        const uint32_t *paths, int path_words, const uint16_t *path_lens, int slab_format)
{
    StepParams p;
    int geom;
    if (int r = build_params(*cfg, p, geom, g_err)) return r;
    p.key.step = step;
    p.slab_fmt = slab_format;
    p.eval_moves = eval_moves;
    p.auto_reset = auto_reset;
    p.max_moves = max_moves;
    p.n_inj = slab ? n_inj : -1;
    p.slab_words = slab_words;
    const bool tile_major = slab_format == SLAB_CHUNKS6T;      // what launch_step does: uint16 chunks, 256 apart
    if (tile_major) p.slab_fmt = slab_format = SLAB_CHUNKS6;
    const bool chunk_major = slab_format == SLAB_CHUNKS10 || slab_format == SLAB_CHUNKS6;
    p.slab_stride = chunk_major ? (tile_major ? 256 : (int32_t)n) : 0;
    p.slab_estride = 1;
    const int eager = slab_format == SLAB_CHUNKS6 ? 3 : 2;
    p.slab_eager = chunk_major ? (slab_words < eager ? slab_words : eager) : 0;
    p.log_cap = log_cap;
    dyn_geom() = p.dyn;
    HC_DISPATCH(geom, (step_batch<G>(p, n, env0, boards, fingers, prev, moves, actions, slab, action_out, reward, info,
                                     final_boards, combo_log, paths, path_words, path_lens, tile_major)));
    return 0;
}

int hostcore_reset(const ppad_config *cfg, int64_t n, uint64_t env0, uint32_t step, const uint32_t *slab, int slab_words,
                   int n_inj, const int32_t *fingers_in, int8_t *boards, int32_t *fingers, uint8_t *flags,
                   int32_t *consumed)
{
    StepParams p;
    int geom;
    if (int r = build_params(*cfg, p, geom, g_err)) return r;
    p.key.step = step;
    p.n_inj = slab ? n_inj : -1;
    p.slab_words = slab_words;
    dyn_geom() = p.dyn;
    HC_DISPATCH(geom, (reset_batch<G>(p, n, env0, slab, fingers_in, boards, fingers, flags, consumed)));
    return 0;
}

int hostcore_cancel(int rows, int cols, int orb_bits, int64_t n, int8_t *boards, uint8_t *n_combos, uint16_t *combo_log,
                    int log_cap)
{
    const int geom = shape_geom(rows, cols, orb_bits);
    if (geom < 0) return 2;
    HC_DISPATCH(geom, (cancel_batch<G>(n, boards, n_combos, combo_log, log_cap)));
    return 0;
}

int hostcore_drop(int rows, int cols, int orb_bits, int64_t n, int8_t *boards)
{
    const int geom = shape_geom(rows, cols, orb_bits);
    if (geom < 0) return 2;
    HC_DISPATCH(geom, (drop_batch<G>(n, boards)));
    return 0;
}

int hostcore_obs(int rows, int cols, int orb_bits, int64_t n, const int8_t *boards, const int32_t *fingers, uint8_t *out)
{
    const int geom = shape_geom(rows, cols, orb_bits);
    if (geom < 0) return 2;
    HC_DISPATCH(geom, (obs_batch<G>(n, boards, fingers, out)));
    return 0;
}

int hostcore_sky_table(const ppad_config *cfg, uint32_t thr[10], uint32_t *valid)
{
    StepParams p;
    int geom;
    if (int r = build_params(*cfg, p, geom, g_err)) return r;
    std::memcpy(thr, p.sky.thr, sizeof(p.sky.thr));
    *valid = p.sky.valid;
    return 0;
}

int hostcore_uniform_k(const ppad_config *cfg)
{
    StepParams p;
    int geom;
    if (build_params(*cfg, p, geom, g_err)) return -1;
    return (int)p.sky.uniform_k;
}

int hostcore_orb_from_u32(const ppad_config *cfg, uint32_t x)
{
    StepParams p;
    int geom;
    if (build_params(*cfg, p, geom, g_err)) return -1;
    return geom >= 3 ? orb_from_u32<10>(p.sky, x) : orb_from_u32<7>(p.sky, x);
}

void hostcore_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    U4 c = {ctr[0], ctr[1], ctr[2], ctr[3]};
    const U4 o = philox4x32_10(c, key[0], key[1]);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}

}   // extern "C"
