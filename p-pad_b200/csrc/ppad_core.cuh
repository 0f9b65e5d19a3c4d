// ppad_core.cuh -- per-board game logic on bit-sliced boards (sm_100a device code).
//
// One thread owns one board.  A board of R x C cells (R*C <= 64) lives in three machine words
// b0,b1,b2: bit i of plane j is bit j of the 3-bit orb code of cell i = row*C + col (row 0 = top,
// gravity towards larger rows; reference orientation, ppad/pad/game.py:30-38).  Codes 0..6 are the
// orb types red, blue, green, light, dark, heal, jammer (game.py:236-246); 7 = empty (-1).
// Bits >= R*C of every plane are zero.  With this layout every step of the reference's hot path
// is a handful of LOP3/SHF instructions:
//   * "same colour as my right/down neighbour" is three XORs and an OR across the planes,
//   * 3-runs are two shift-ANDs of that edge mask (utils.py:47-54),
//   * islands (utils.py:86-111) are the closure of a seed under the two edge masks,
//   * gravity (game.py:201-221) moves whole rows of bits, skyfall (game.py:232-256) walks the
//     empties in ascending bit order, which IS the reference's row-major order.
//
// The functions are __host__ __device__ so that tests/hostcore/ can compile the very same code
// with g++ and fuzz it against the CPU oracle without a GPU.  That host build is test
// infrastructure only: the shipped library (ppad_kernels.cu) contains device code only and has no
// CPU path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PPAD_HD __host__ __device__ __forceinline__
#define PPAD_HDC __host__ __device__ constexpr
#define PPAD_TIGHT_DEFAULT false
#else
#define PPAD_HD inline
#define PPAD_HDC constexpr
#define PPAD_TIGHT_DEFAULT true   // the host build (tests/hostcore) fuzzes the unrolled Philox refill
#endif

// tuning knobs (profiles/variants.py builds A/B variants with -D...)
#define PPAD_PRAGMA_(x) _Pragma(#x)
#define PPAD_PRAGMA(x) PPAD_PRAGMA_(x)
#if defined(PPAD_PHILOX_NOINLINE) && defined(__CUDACC__)
#define PPAD_PHILOX_FN __host__ __device__ __noinline__
#else
#define PPAD_PHILOX_FN PPAD_HD
#endif

namespace ppad {

// ---------------------------------------------------------------------------------------------
// small portability shims (device intrinsics / host builtins)
// ---------------------------------------------------------------------------------------------
PPAD_HD int popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
PPAD_HD int popc(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
// index of the lowest set bit (x != 0)
PPAD_HD int ctz(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
PPAD_HD int ctz(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
// IEEE-754 round-to-nearest double ops that can never be contracted into an FMA: the reward
// (game.py:157-199) is order- and contraction-sensitive.  Host build uses -ffp-contract=off.
PPAD_HD double dmul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
PPAD_HD double dadd(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
// Warp-level convergence helpers.  The per-board code is full of data-dependent loops; without explicit
// reconvergence points the compiler is free to let lanes drift apart for the rest of the kernel (measured:
// 4.5 active lanes per warp instruction).  Every lane of the warp must reach these calls.
PPAD_HD bool warp_any(bool pred)
{
#if defined(__CUDA_ARCH__)
    return __any_sync(0xFFFFFFFFu, pred) != 0;
#else
    return pred;
#endif
}
PPAD_HD void warp_converge()
{
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}

PPAD_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// ---------------------------------------------------------------------------------------------
// per-board fault bits and info word (mirrors include/ppad_b200.h)
// ---------------------------------------------------------------------------------------------
enum : uint32_t {
    FLAG_REJECTED = 0x01, FLAG_SKYFALL_EXHAUSTED = 0x02, FLAG_CASCADE_CAP = 0x04, FLAG_NO_LEGAL_MOVE = 0x08,
    FLAG_BAD_ACTION = 0x10, FLAG_RESET_CAP = 0x20, FLAG_UNSUPPORTED_ORB = 0x40, FLAG_EPISODE_DONE = 0x80
};
enum : uint32_t { STREAM_SKYFALL = 0, STREAM_ACTION = 1, STREAM_RESET = 2, STREAM_FINGER = 3, STREAM_POLICY = 4 };
enum : int { ACT_UP = 0, ACT_DOWN = 1, ACT_LEFT = 2, ACT_RIGHT = 3, ACT_PASS = 4, ACT_NOOP = 5 };

PPAD_HD uint32_t sat255(int v) { return v > 255 ? 255u : (uint32_t)v; }
PPAD_HD uint32_t make_info(int n_combos, int depth, int consumed, uint32_t flags)
{
    return sat255(n_combos) | (sat255(depth) << 8) | (sat255(consumed) << 16) | ((flags & 0xFFu) << 24);
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
template <bool WIDE> struct WordOf { typedef uint32_t type; };
template <> struct WordOf<true> { typedef uint64_t type; };

// NP_ = bit planes per cell: 3 (orb types 0..6, 7 = empty: everything the default skyfall produces) or 4 (orb
// types 0..9 of game.py:236-246 incl. poison, mortal poison and bomb, 15 = empty).
// The board dimensions are read through accessor functions (rows(), cols(), cells(), board(), ...): compile-time
// constants for the specialised geometries below, run-time values for GeomDyn.
template <int R_, int C_, int NP_ = 3>
struct Geom {
    static constexpr bool DYNAMIC = false;
    static constexpr int R = R_, C = C_, CELLS = R_ * C_, NP = NP_;
    static constexpr int MAX_CELLS = CELLS;               // for array sizes
    static constexpr int EMPTY = (1 << NP_) - 1;          // code of an empty cell (-1 in the reference)
    static constexpr int NTYPES = NP_ == 3 ? 7 : 10;      // orb types 0 .. NTYPES - 1
    static_assert(CELLS <= 64 && R_ >= 3 && C_ >= 3 && (NP_ == 3 || NP_ == 4), "unsupported board size");
    typedef typename WordOf<(CELLS > 32)>::type W;
    static constexpr int WBITS = 8 * (int)sizeof(W);
    static constexpr W ONE = 1;
    static constexpr W BOARD = CELLS == WBITS ? ~(W)0 : (((W)1 << (CELLS % WBITS)) - 1);
    static constexpr W col_mask(int c)
    {
        W m = 0;
        for (int r = 0; r < R_; r++) m |= (W)1 << (r * C_ + c);
        return m;
    }
    static constexpr W row_mask(int r) { return (((W)1 << C_) - 1) << (r * C_); }
    static constexpr W NOT_LASTCOL = BOARD & ~col_mask(C_ - 1);
    static constexpr W NOT_LASTROW = BOARD & ~row_mask(R_ - 1);
    static constexpr int OBS_CH = 7;                      // agent01.py:263 fixes 7 channels
    static constexpr int OBS_ELEMS = CELLS * OBS_CH;      // 210 (5x6) / 294 (6x7)
    // bytes of one packed state record in HBM: NP planes + the meta word, padded to 16 bytes
    static constexpr int STATE_BYTES = (NP_ * (int)sizeof(W) + 4 + 15) / 16 * 16;
    PPAD_HDC static int rows() { return R_; }
    PPAD_HDC static int cols() { return C_; }
    PPAD_HDC static int cells() { return CELLS; }
    PPAD_HDC static int obs_elems() { return OBS_ELEMS; }
    PPAD_HDC static W board() { return BOARD; }
    PPAD_HDC static W not_lastcol() { return NOT_LASTCOL; }
    PPAD_HDC static W not_lastrow() { return NOT_LASTROW; }
};

// Run-time board dimensions: any dim_row x dim_col with at most 64 cells (PAD.__init__ accepts every size with at
// least 13 orbs, game.py:54-68).  The slower path behind the specialised geometries: 64-bit planes, loops that are
// not unrolled, a simple observation emitter.  The dimensions of the launch live in shared memory (one copy per CTA,
// written by geom_enter() at the top of every kernel); the host build keeps them in a thread-local.
struct DynGeomData {
    int32_t rows, cols, cells, obs_elems;
    uint64_t board, not_lastcol, not_lastrow;
};

PPAD_HD DynGeomData &dyn_geom()
{
#if defined(__CUDA_ARCH__)
    __shared__ DynGeomData d;
    return d;
#else
    static thread_local DynGeomData d;
    return d;
#endif
}

inline DynGeomData make_dyn_geom(int rows, int cols)
{
    DynGeomData d;
    d.rows = rows;
    d.cols = cols;
    d.cells = rows * cols;
    d.obs_elems = 7 * rows * cols;
    d.board = d.cells >= 64 ? ~0ull : ((1ull << d.cells) - 1);
    uint64_t lastcol = 0, lastrow = 0;
    for (int r = 0; r < rows; r++) lastcol |= 1ull << (r * cols + cols - 1);
    for (int c = 0; c < cols; c++) lastrow |= 1ull << ((rows - 1) * cols + c);
    d.not_lastcol = d.board & ~lastcol;
    d.not_lastrow = d.board & ~lastrow;
    return d;
}

template <int NP_ = 3>
struct GeomDyn {
    static constexpr bool DYNAMIC = true;
    static constexpr int NP = NP_, MAX_CELLS = 64;
    static constexpr int EMPTY = (1 << NP_) - 1;
    static constexpr int NTYPES = NP_ == 3 ? 7 : 10;
    typedef uint64_t W;
    static constexpr int WBITS = 64;
    static constexpr W ONE = 1;
    static constexpr int OBS_CH = 7;
    static constexpr int STATE_BYTES = (NP_ * 8 + 4 + 15) / 16 * 16;   // 32 / 48: the record layout of the 64-bit geometries
    PPAD_HD static int rows() { return dyn_geom().rows; }
    PPAD_HD static int cols() { return dyn_geom().cols; }
    PPAD_HD static int cells() { return dyn_geom().cells; }
    PPAD_HD static int obs_elems() { return dyn_geom().obs_elems; }
    PPAD_HD static W board() { return dyn_geom().board; }
    PPAD_HD static W not_lastcol() { return dyn_geom().not_lastcol; }
    PPAD_HD static W not_lastrow() { return dyn_geom().not_lastrow; }
};

// first statement of every kernel (all threads): publishes the launch's board dimensions for GeomDyn
template <class G>
PPAD_HD void geom_enter(const DynGeomData &d)
{
    if constexpr (G::DYNAMIC) {
#if defined(__CUDA_ARCH__)
        if (threadIdx.x == 0) dyn_geom() = d;
        __syncthreads();
#else
        dyn_geom() = d;
#endif
    }
}

template <class W, int NP>
struct Planes {
    W b0, b1, b2;
};
template <class W>
struct Planes<W, 4> {
    W b0, b1, b2, b3;
};
template <class G>
struct Board : Planes<typename G::W, G::NP> {};
// statement for the fourth plane of the 4-bit geometries
#define PPAD_P4(...)                    \
    if constexpr (G::NP > 3) {          \
        __VA_ARGS__;                    \
    }

// meta word: finger cell [0,6) | previous action [6,9) | accepted moves this episode [9,25)
PPAD_HD uint32_t meta_pack(int finger, int prev, int moves)
{
    return (uint32_t)finger | ((uint32_t)prev << 6) | ((uint32_t)moves << 9);
}
PPAD_HD int meta_finger(uint32_t m) { return (int)(m & 63u); }
PPAD_HD int meta_prev(uint32_t m) { return (int)((m >> 6) & 7u); }
PPAD_HD int meta_moves(uint32_t m) { return (int)((m >> 9) & 0xFFFFu); }

template <class G>
struct State {
    Board<G> b;
    uint32_t meta;
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Counter = (env_lo, env_hi, step, stream << 24 | block),
// key = seed: results do not depend on how envs are sharded over GPUs.
// ---------------------------------------------------------------------------------------------
struct U4 {
    uint32_t x, y, z, w;
};

PPAD_PHILOX_FN U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t lo0 = 0xD2511F53u * c.x, hi0 = mulhi32(0xD2511F53u, c.x);
        const uint32_t lo1 = 0xCD9E8D57u * c.z, hi1 = mulhi32(0xCD9E8D57u, c.z);
        U4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

struct RngKey {
    uint32_t k0, k1, env_lo, env_hi, step;
};

// the d-th 32-bit draw of (env, step, stream)
PPAD_HD uint32_t philox_draw(const RngKey &k, uint32_t stream, uint32_t d)
{
    U4 c;
    c.x = k.env_lo;
    c.y = k.env_hi;
    c.z = k.step;
    c.w = (stream << 24) | (d >> 2);
    const U4 o = philox4x32_10(c, k.k0, k.k1);
    const uint32_t l = d & 3u;
    return l == 0 ? o.x : (l == 1 ? o.y : (l == 2 ? o.z : o.w));
}

// The random policy's draw of (env, step): word step & 3 of the block (env, step >> 2, ACTION stream).  Four
// consecutive steps share one Philox block, which a fused multi-step rollout computes once for all four.
PPAD_HD U4 action_block(const RngKey &k)
{
    U4 c;
    c.x = k.env_lo;
    c.y = k.env_hi;
    c.z = k.step >> 2;
    c.w = STREAM_ACTION << 24;
    return philox4x32_10(c, k.k0, k.k1);
}
PPAD_HD uint32_t action_word(const U4 &b, uint32_t step)
{
    const uint32_t l = step & 3u;
    return l == 0 ? b.x : (l == 1 ? b.y : (l == 2 ? b.z : b.w));
}

// Categorical skyfall draw.  game.py:458-470 returns the first i with cumsum(prob)[i] >= u and
// prob[i] > 0; with u = x / 2^32 that is x <= thr[i] for thr[i] = floor(cumsum[i] * 2^32) (host
// computes thr, tests/test_host_logic.py checks it against the float rule).  valid bit i = prob[i] > 0.
struct SkyfallTable {
    uint32_t thr[10];
    uint32_t valid;
    uint32_t uniform_k;   // K > 0: the valid types are 0 .. K-1 with thr[k-1] == floor(k * 2^32 / K) (host-verified), see below
};

template <int NT = 7>   // number of orb types of the geometry (7 or 10)
PPAD_HD int orb_from_u32(const SkyfallTable &t, uint32_t x)
{
    // thr is non-decreasing and a type with prob 0 repeats its predecessor's threshold, so the first i with
    // x <= thr[i] is a valid type and equals the NUMBER of thresholds below x -- except for x = 0 when the
    // leading types have prob 0 (thr = 0): then it is the first valid type.  The host guarantees
    // thr = 0xFFFFFFFF from the last valid type on, so the count never passes it.
    if (t.uniform_k) {   // kernel-uniform
        // K equally likely types 0 .. K-1 (the reference's default skyfall, parameters.py:43, has K = 6): the count of
        // thresholds floor(k 2^32 / K) below x is floor(x K / 2^32), except that x == k 2^32 / K exactly (an integer
        // threshold, reached when the low word of x K is 0) still belongs to the type below.  One widening multiply
        // instead of NT - 1 compares -- 19 % of the instructions of a reset-heavy step (profiles/r02b).
        const uint64_t wide = (uint64_t)x * t.uniform_k;
        return (int)(wide >> 32) - (((uint32_t)wide == 0u && x != 0u) ? 1 : 0);
    }
    int orb = 0;
#pragma unroll
    for (int i = 0; i < NT - 1; i++) orb += x > t.thr[i] ? 1 : 0;
    return x ? orb : ctz(t.valid);
}

// ---------------------------------------------------------------------------------------------
// Orb source: an injected per-board sequence first, then the Philox stream at the same draw index d.
// Two slab layouts (include/ppad_b200.h: ppad_slab_format):
//   SLAB_NIBBLES  4-bit nibbles, 8 per uint32: orb d in word d >> 3 at bit 4 * (d & 7)
//   SLAB_PLANES3  bit-sliced 3-bit codes, 32 orbs per group of three uint32 {p0, p1, p2}: bit (d & 31) of word
//                 3 * (d >> 5) + j is bit j of orb d -- 12 bytes per 32 orbs (orb types 0..6 only)
//   SLAB_CHUNKS10 chunks of ten 3-bit codes in one uint32, bit-sliced inside the chunk: bit 10 j + k of chunk c is bit j
//                 of orb 10 c + k (bits 30, 31 unused); the chunks of one board lie `stride` words apart (chunk-major
//                 over the batch: chunk c of board i at slab[c * stride + i]), so that a warp reads one chunk of its
//                 32 boards as one line.  Made for a slab that stays in pinned HOST memory: the first n_eager chunks
//                 are fetched up front by the kernel (`eager`, full lines, all in flight together), every further
//                 chunk only when the cascade gets there -- the host can offer 50 orbs or more per board while 8 bytes
//                 per board cross PCIe (orb types 0..6 only).  A read that waits on PCIe in the middle of a cascade
//                 stalls its whole warp, so n_eager is chosen to make those rare (2: one board in a hundred).
//   SLAB_CHUNKS6  the same chunk-major, eager / on-demand scheme with chunks of SIX orbs as base-6 digits in one uint16
//                 (orb 6 c + k = (chunk c / 6^k) % 6: the six skyfall colours 0..5 at log2(6^6) = 15.5 of 16 bits);
//                 slab and stride are in uint16 units.  Three eager chunks = 18 orbs in 6 bytes per board.  A chunk
//                 value >= 6^6 is no encoding and is flagged like a bad code.
// Injected codes that are not orb types of the geometry (>= NT) are handed on as they are and flagged
// (bad -> FLAG_UNSUPPORTED_ORB).
// ---------------------------------------------------------------------------------------------
//   SLAB_CHUNKS6T the same uint16 chunks, chunk-major per TILE of 256 boards instead of over the whole batch: chunk c of
//                 board i at ((i / 256) * chunks_per_board + c) * 256 + i % 256.  The eager chunks of a CTA's 256
//                 boards are then ONE contiguous block (1536 bytes for three chunks: twelve full 128-byte lines
//                 instead of 24 half lines), fetched cooperatively.  To the per-board code it is SLAB_CHUNKS6 with a
//                 chunk stride of 256.
// A chunk value >= 6^6 is no encoding and is flagged like a bad code.
enum : int { SLAB_NIBBLES = 0, SLAB_PLANES3 = 1, SLAB_CHUNKS10 = 2, SLAB_CHUNKS6 = 3, SLAB_CHUNKS6T = 4 };
PPAD_HD uint32_t div6_u16(uint32_t w) { return (w * 43691u) >> 18; }   // floor(w / 6), exact for w < 65536

struct OrbSource {
    const uint32_t *slab;   // this board's words, or nullptr
    int n_inj;              // injected orbs available; < 0 = pure Philox (no EXHAUSTED flag)
    int cursor;             // orbs handed out so far
    uint32_t stream;
    uint32_t word;          // cached slab word (nibble layout)
    int word_idx;
    U4 blk;                 // cached Philox block
    int blk_idx;
    bool exhausted;
    int fmt;                // SLAB_NIBBLES / SLAB_PLANES3 / SLAB_CHUNKS10
    uint32_t bad;           // bit 0: an injected code was not an orb type of the geometry
    const uint32_t *eager;  // SLAB_CHUNKS10: the first n_eager chunks of this board (fetched up front by the caller)
    int n_eager;
    int stride;             // SLAB_CHUNKS10: words between consecutive chunks of this board
    int estride;            // elements between consecutive EAGER chunks of this board where the caller keeps them

    PPAD_HD void init(const uint32_t *slab_, int n_inj_, uint32_t stream_, int fmt_ = SLAB_NIBBLES,
                      const uint32_t *eager_ = nullptr, int n_eager_ = 0, int stride_ = 0, int estride_ = 1)
    {
        eager = eager_;
        n_eager = eager_ ? n_eager_ : 0;
        stride = stride_;
        estride = estride_;
        slab = slab_;
        n_inj = slab_ ? n_inj_ : -1;
        cursor = 0;
        stream = stream_;
        word = 0;
        word_idx = -1;
        blk_idx = -1;
        exhausted = false;
        fmt = fmt_;
        bad = 0;
        blk.x = blk.y = blk.z = blk.w = 0;
    }

    PPAD_HD uint32_t chunk(int c) const { return c < n_eager ? eager[c] : slab[(size_t)c * (size_t)stride]; }
    PPAD_HD uint32_t chunk16(int c)   // SLAB_CHUNKS6: slab, eager and stride in uint16 units
    {
        const uint32_t w = c < n_eager ? reinterpret_cast<const uint16_t *>(eager)[c * estride]
                                       : reinterpret_cast<const uint16_t *>(slab)[(size_t)c * (size_t)stride];
        bad |= w >= 46656u ? 1u : 0u;
        return w;
    }

    template <int NT = 7>
    PPAD_HD int next(const RngKey &key, const SkyfallTable &sky)
    {
        const int d = cursor++;
        if (d < n_inj) {
            int code;
            if (fmt == SLAB_PLANES3) {
                // the general path is rare (a slab that runs out in the middle of a fill; reset from a slab): no caching
                const uint32_t *g = slab + 3 * (d >> 5);
                const int k = d & 31;
                code = (int)(((g[0] >> k) & 1u) | (((g[1] >> k) & 1u) << 1) | (((g[2] >> k) & 1u) << 2));
            } else if (fmt == SLAB_CHUNKS6) {
                const int c = d / 6;
                uint32_t w = chunk16(c);
                for (int k = d - 6 * c; k > 0; k--) w = div6_u16(w);
                code = (int)(w - 6u * div6_u16(w));
            } else if (fmt == SLAB_CHUNKS10) {
                const int c = d / 10, k = d - 10 * c;
                const uint32_t w = chunk(c);
                code = (int)(((w >> k) & 1u) | (((w >> (10 + k)) & 1u) << 1) | (((w >> (20 + k)) & 1u) << 2));
            } else {
                const int wi = d >> 3;
                if (wi != word_idx) {
                    word = slab[wi];
                    word_idx = wi;
                }
                code = (int)((word >> (4 * (d & 7))) & 15u);
            }
            bad |= code >= NT ? 1u : 0u;
            return code;
        }
        if (n_inj >= 0) exhausted = true;
        const int bi = d >> 2;
        if (bi != blk_idx) {
            U4 c;
            c.x = key.env_lo;
            c.y = key.env_hi;
            c.z = key.step;
            c.w = (stream << 24) | (uint32_t)bi;
            blk = philox4x32_10(c, key.k0, key.k1);
            blk_idx = bi;
        }
        const int l = d & 3;
        const uint32_t x = l == 0 ? blk.x : (l == 1 ? blk.y : (l == 2 ? blk.z : blk.w));
        return orb_from_u32<NT>(sky, x);
    }
};

// ---------------------------------------------------------------------------------------------
// Team table for the reward (game.py:157-199, parameters.py:54-89): for each attack colour
// (orb types 0..4) the (multiplier, atk) of the matching units in team order.
// ---------------------------------------------------------------------------------------------
#define PPAD_MAX_TEAM 8
struct TeamTable {
    double mult[5][PPAD_MAX_TEAM];
    double atk[5][PPAD_MAX_TEAM];
    int32_t n[8];   // n[0..4] used
};

struct Tally {
    double t0, t1, t2, t3, t4;   // damage_tracker by orb type: red, blue, green, light, dark
    int n_combos;

    PPAD_HD void init()
    {
        t0 = t1 = t2 = t3 = t4 = 0.0;
        n_combos = 0;
    }
    // one combo (game.py:169-190): tracker[colour] += (mult * (1 + 0.25 (N - 3))) * atk per unit
    PPAD_HD void add(const TeamTable &team, int colour, int N)
    {
        n_combos++;
        if (colour > 4) return;   // heal / jammer: counts as a combo, no unit matches
        const double f = 1.0 + 0.25 * (double)(N - 3);   // exact
        const int nu = team.n[colour];
        // pick this colour's tracker once, add the units in team order, put it back: plain selects around a short
        // loop instead of a colour switch per unit (which compiled to indirect branches that split the warp)
        double tr = colour == 0 ? t0 : (colour == 1 ? t1 : (colour == 2 ? t2 : (colour == 3 ? t3 : t4)));
        for (int u = 0; u < nu; u++) tr = dadd(tr, dmul(dmul(team.mult[colour][u], f), team.atk[colour][u]));
        t0 = colour == 0 ? tr : t0;
        t1 = colour == 1 ? tr : t1;
        t2 = colour == 2 ? tr : t2;
        t3 = colour == 3 ? tr : t3;
        t4 = colour == 4 ? tr : t4;
    }
    // the same with the colour known at compile time (the colour-major cancel below): no selects at all
    template <int COLOUR>
    PPAD_HD void add_c(const TeamTable &team, int N)
    {
        n_combos++;
        if (COLOUR > 4) return;
        const double f = 1.0 + 0.25 * (double)(N - 3);   // exact
        double &tr = COLOUR == 0 ? t0 : (COLOUR == 1 ? t1 : (COLOUR == 2 ? t2 : (COLOUR == 3 ? t3 : t4)));
        constexpr int CI = COLOUR > 4 ? 0 : COLOUR;
        const int nu = team.n[CI];
#pragma unroll 1
        for (int u = 0; u < nu; u++) tr = dadd(tr, dmul(dmul(team.mult[CI][u], f), team.atk[CI][u]));
    }
    // game.py:193-199: every tracker *= 1 + 0.25 (#combos - 1); sum in dict order blue, red,
    // green, dark, light starting from the int 0
    PPAD_HD double total() const
    {
        const double m = 1.0 + 0.25 * (double)(n_combos - 1);
        double s = dmul(t1, m);
        s = dadd(s, dmul(t0, m));
        s = dadd(s, dmul(t2, m));
        s = dadd(s, dmul(t4, m));
        s = dadd(s, dmul(t3, m));
        return s;
    }
};

// ---------------------------------------------------------------------------------------------
// board primitives
// ---------------------------------------------------------------------------------------------
template <class G>
PPAD_HD typename G::W empties(const Board<G> &b)
{
    typename G::W e = b.b0 & b.b1 & b.b2;   // bits >= CELLS are zero in every plane
    PPAD_P4(e &= b.b3)
    return e;
}

template <class G>
PPAD_HD int code_at(const Board<G> &b, int cell)
{
    int code = (int)(((b.b0 >> cell) & 1) | (((b.b1 >> cell) & 1) << 1) | (((b.b2 >> cell) & 1) << 2));
    PPAD_P4(code |= (int)((b.b3 >> cell) & 1) << 3)
    return code;
}

template <class G>
PPAD_HD void set_code(Board<G> &b, int cell, int code)
{
    typedef typename G::W W;
    const W bit = G::ONE << cell;
    b.b0 = (b.b0 & ~bit) | ((W)(code & 1) << cell);
    b.b1 = (b.b1 & ~bit) | ((W)((code >> 1) & 1) << cell);
    b.b2 = (b.b2 & ~bit) | ((W)((code >> 2) & 1) << cell);
    PPAD_P4(b.b3 = (b.b3 & ~bit) | ((W)((code >> 3) & 1) << cell))
}

// swap the codes of cells i and j (game.py:429-456)
template <class G>
PPAD_HD void swap_cells(Board<G> &b, int i, int j)
{
    typedef typename G::W W;
    W x;
    x = ((b.b0 >> i) ^ (b.b0 >> j)) & 1;
    b.b0 ^= (x << i) | (x << j);
    x = ((b.b1 >> i) ^ (b.b1 >> j)) & 1;
    b.b1 ^= (x << i) | (x << j);
    x = ((b.b2 >> i) ^ (b.b2 >> j)) & 1;
    b.b2 ^= (x << i) | (x << j);
    PPAD_P4(x = ((b.b3 >> i) ^ (b.b3 >> j)) & 1; b.b3 ^= (x << i) | (x << j))
}

// Edge masks and matched cells.  er bit i: cell i is non-empty and has the colour of cell i+1 in
// the same row; ed bit i: same for the cell below.  A 3-run is two consecutive edges
// (utils.py:47-54).  Returns the mask of all cells lying in some 3-run.
template <class G>
PPAD_HD typename G::W match_mask(const Board<G> &b, typename G::W &er, typename G::W &ed)
{
    typedef typename G::W W;
    const int C = G::cols();
    const W ne = ~empties<G>(b);
    W xr = (b.b0 ^ (b.b0 >> 1)) | (b.b1 ^ (b.b1 >> 1)) | (b.b2 ^ (b.b2 >> 1));
    W xd = (b.b0 ^ (b.b0 >> C)) | (b.b1 ^ (b.b1 >> C)) | (b.b2 ^ (b.b2 >> C));
    PPAD_P4(xr |= b.b3 ^ (b.b3 >> 1); xd |= b.b3 ^ (b.b3 >> C))
    er = ~xr & ne & G::not_lastcol();
    ed = ~xd & ne & G::not_lastrow();
    const W h = er & (er >> 1);
    const W v = ed & (ed >> C);
    return h | (h << 1) | (h << 2) | v | (v << C) | (v << (2 * C));
}

// closure of g under the same-colour adjacency (the island flood fill of utils.py:86-111)
template <class G>
PPAD_HD typename G::W closure(typename G::W g, typename G::W er, typename G::W ed)
{
    typedef typename G::W W;
    const int C = G::cols();
    for (;;) {
        const W n = g | ((g & er) << 1) | ((g >> 1) & er) | ((g & ed) << C) | ((g >> C) & ed);
        if (n == g) return g;
        g = n;
    }
}

// One cancel() (utils.py:26-70): emits (colour, N) per combo in the reference's order -- ascending
// row-major index of each island's first cell -- then erases the matched cells.
// Islands are taken over ALL cells of a colour (matched or not), so two runs bridged by unmatched
// orbs of the same colour are one combo (hard part 2).
template <class G, class F>
PPAD_HD int cancel_board(Board<G> &b, F &&on_combo)
{
    typedef typename G::W W;
    W er, ed;
    const W m = match_mask<G>(b, er, ed);
    if (m == 0) return 0;
    W todo = closure<G>(m, er, ed);   // every island that holds a matched cell
    int n = 0;
    while (todo) {
        const W seed = todo & (~todo + 1);          // lowest cell = that island's first cell
        const W comp = closure<G>(seed, er, ed);
        on_combo(code_at<G>(b, ctz(seed)), popc(comp & m));
        todo &= ~comp;
        n++;
    }
    b.b0 |= m;
    b.b1 |= m;
    b.b2 |= m;
    PPAD_P4(b.b3 |= m)
    return n;
}

// The cells of orb type COLOUR (bits >= CELLS may be set; AND with a board mask before use).
template <class G, int COLOUR>
PPAD_HD typename G::W colour_plane(const Board<G> &b)
{
    typename G::W c = ((COLOUR & 1) ? b.b0 : ~b.b0) & ((COLOUR & 2) ? b.b1 : ~b.b1) & ((COLOUR & 4) ? b.b2 : ~b.b2);
    PPAD_P4(c &= (COLOUR & 8) ? b.b3 : ~b.b3)
    return c;
}

// cancel() straight into the reward tally, colour-major: the per-colour trackers of damage() (game.py:163-190)
// only see the combos of their own colour, so the reference's result depends on the ORDER of combos only
// within one colour -- and one colour almost always has a single combo per cancel (two need >= 6 matched cells
// of that colour), which needs no flood fill at all: N = popc(matched & colour).  Colours with >= 6 matched
// cells are left to one generic loop that peels their islands in the reference's order (utils.py:35-44).
// m/er/ed are match_mask() of b, m != 0.
template <class G, int COLOUR>
PPAD_HD void cancel_colour(const Board<G> &b, typename G::W m, typename G::W &slow, Tally &tally, const TeamTable &team)
{
    typedef typename G::W W;
    const W mc = m & colour_plane<G, COLOUR>(b);
    const int cnt = popc(mc);
    if (cnt >= 6) slow |= mc;
    else if (cnt) tally.add_c<COLOUR>(team, cnt);   // every combo holds >= 3 matched cells: fewer than 6 is ONE island
}

// With a combo log (parity / debugging; kernel-uniform) every colour takes the generic loop, which then emits the
// combos in the reference's global order.
template <class G>
PPAD_HD void cancel_tally(Board<G> &b, typename G::W m, Tally &tally, const TeamTable &team, uint16_t *log = nullptr,
                          int log_cap = 0)
{
    typedef typename G::W W;
    W slow = m;
    if (!log) {
        slow = 0;
        cancel_colour<G, 0>(b, m, slow, tally, team);
        cancel_colour<G, 1>(b, m, slow, tally, team);
        cancel_colour<G, 2>(b, m, slow, tally, team);
        cancel_colour<G, 3>(b, m, slow, tally, team);
        cancel_colour<G, 4>(b, m, slow, tally, team);
        cancel_colour<G, 5>(b, m, slow, tally, team);
        cancel_colour<G, 6>(b, m, slow, tally, team);
        PPAD_P4(cancel_colour<G, 7>(b, m, slow, tally, team); cancel_colour<G, 8>(b, m, slow, tally, team);
                cancel_colour<G, 9>(b, m, slow, tally, team))
    }
    if (slow) {
        // the edge masks are only needed here (about one board in twenty): recomputed rather than carried through the
        // straight-line colour blocks above, where they would cost two registers on the hot path
        W er, ed;
        match_mask<G>(b, er, ed);
        // usually those cells still form ONE island (an L, T or cross, a run of 6): one flood fill settles it
        const W first = closure<G>(slow & (~slow + 1), er, ed);
        W todo = first;
        if (slow & ~first) todo |= closure<G>(slow & ~first, er, ed);   // all islands that hold those matched cells
        while (todo) {
            const W seed = todo & (~todo + 1);      // lowest cell = that island's first cell
            const W comp = (seed & first) ? first : closure<G>(seed, er, ed);
            const int colour = code_at<G>(b, ctz(seed)), N = popc(comp & m);
            if (log && tally.n_combos < log_cap) log[tally.n_combos] = (uint16_t)(colour | (N << 8));
            tally.add(team, colour, N);
            todo &= ~comp;
        }
    }
    b.b0 |= m;
    b.b1 |= m;
    b.b2 |= m;
    PPAD_P4(b.b3 |= m)
}

// gravity (game.py:201-221): every orb with an empty cell directly below moves down one row,
// repeated until nothing moves
template <class G>
PPAD_HD void drop_board(Board<G> &b)
{
    typedef typename G::W W;
    const int C = G::cols();
    for (;;) {
        const W e = empties<G>(b);
        const W fall = ~e & (e >> C) & G::board();
        if (fall == 0) return;
        const W dst = fall << C;
        b.b0 = (b.b0 & ~dst) | ((b.b0 & fall) << C) | fall;
        b.b1 = (b.b1 & ~dst) | ((b.b1 & fall) << C) | fall;
        b.b2 = (b.b2 & ~dst) | ((b.b2 & fall) << C) | fall;
        PPAD_P4(b.b3 = (b.b3 & ~dst) | ((b.b3 & fall) << C) | fall)
    }
}

// skyfall (game.py:232-256): empties in ascending cell index = row-major, one draw each
template <class G>
PPAD_HD void fill_board(Board<G> &b, OrbSource &src, const RngKey &key, const SkyfallTable &sky)
{
    typedef typename G::W W;
    W e = empties<G>(b);
    if (e == 0) return;
    if (src.cursor + popc(e) <= src.n_inj) {
        // every orb comes from the injected slab: a tight loop over the orb stream, no source bookkeeping
        int d = src.cursor;
        uint32_t badacc = 0;   // bit 0 of the OR of "this code is no orb type" over the orbs taken
        if (src.fmt == SLAB_PLANES3) {
            const uint32_t *g = src.slab + 3 * (d >> 5);
            uint32_t s0 = g[0] >> (d & 31), s1 = g[1] >> (d & 31), s2 = g[2] >> (d & 31);
            do {
                const W bit = e & (~e + 1);
                e ^= bit;
                // an empty cell has its bit set in all three planes: clear the planes where the code has a 0
                b.b0 ^= bit & ((W)(s0 & 1u) - 1);
                b.b1 ^= bit & ((W)(s1 & 1u) - 1);
                b.b2 ^= bit & ((W)(s2 & 1u) - 1);
                PPAD_P4(b.b3 ^= bit)           // 3-bit codes: plane 3 of a 4-plane board is always 0
                badacc |= s0 & s1 & s2;        // code 7 is no orb type
                d++;
                s0 >>= 1;
                s1 >>= 1;
                s2 >>= 1;
                if ((d & 31) == 0 && e) {
                    g = src.slab + 3 * (d >> 5);
                    s0 = g[0];
                    s1 = g[1];
                    s2 = g[2];
                }
            } while (e);
        } else if (src.fmt == SLAB_CHUNKS6) {
            int c = d / 6, left = 6 * (c + 1) - d;     // orbs left in the chunk at hand
            uint32_t w = src.chunk16(c);
            for (int k = left; k < 6; k++) w = div6_u16(w);   // drop the digits already handed out
            do {
                const W bit = e & (~e + 1);
                e ^= bit;
                const uint32_t q = div6_u16(w), digit = w - 6u * q;
                w = q;
                b.b0 ^= bit & ((W)(digit & 1u) - 1);
                b.b1 ^= bit & ((W)((digit >> 1) & 1u) - 1);
                b.b2 ^= bit & ((W)(digit >> 2) - 1);
                PPAD_P4(b.b3 ^= bit)
                d++;
                if (--left == 0 && e) {
                    c++;
                    left = 6;
                    w = src.chunk16(c);
                }
            } while (e);
        } else if (src.fmt == SLAB_CHUNKS10) {
            int c = d / 10, left = 10 * (c + 1) - d;   // orbs left in the chunk at hand
            uint32_t w = src.chunk(c) >> (10 - left);
            do {
                const W bit = e & (~e + 1);
                e ^= bit;
                b.b0 ^= bit & ((W)(w & 1u) - 1);
                b.b1 ^= bit & ((W)((w >> 10) & 1u) - 1);
                b.b2 ^= bit & ((W)((w >> 20) & 1u) - 1);
                PPAD_P4(b.b3 ^= bit)
                badacc |= w & (w >> 10) & (w >> 20);
                d++;
                w >>= 1;
                if (--left == 0 && e) {
                    c++;
                    left = 10;
                    w = src.chunk(c);
                }
            } while (e);
        } else {
            uint32_t word = src.slab[d >> 3] >> (4 * (d & 7));
            do {
                const W bit = e & (~e + 1);
                e ^= bit;
                b.b0 ^= bit & ((W)(word & 1u) - 1);
                b.b1 ^= bit & ((W)((word >> 1) & 1u) - 1);
                b.b2 ^= bit & ((W)((word >> 2) & 1u) - 1);
                PPAD_P4(b.b3 ^= bit & ((W)((word >> 3) & 1u) - 1))
                if constexpr (G::NP == 3) badacc |= (word >> 3) | (word & (word >> 1) & (word >> 2));   // code >= 7
                else badacc |= (word >> 3) & ((word >> 1) | (word >> 2));                                  // code >= 10
                d++;
                word >>= 4;
                if ((d & 7) == 0 && e) word = src.slab[d >> 3];
            } while (e);
        }
        src.bad |= badacc & 1u;
        src.cursor = d;
        src.word_idx = -1;
        return;
    }
    if (src.n_inj < 0) {
        // pure Philox skyfall: one block per four orbs, the words of a block taken by compile-time index
        int d = src.cursor;
        do {
            U4 c;
            c.x = key.env_lo;
            c.y = key.env_hi;
            c.z = key.step;
            c.w = (src.stream << 24) | (uint32_t)(d >> 2);
            const U4 blk = philox4x32_10(c, key.k0, key.k1);
            const int l0 = d & 3;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (k >= l0 && e) {
                    const int code = orb_from_u32<G::NTYPES>(sky, k == 0 ? blk.x : (k == 1 ? blk.y : (k == 2 ? blk.z : blk.w)));
                    const W bit = e & (~e + 1);
                    e ^= bit;
                    b.b0 ^= bit & ((W)(code & 1) - 1);
                    b.b1 ^= bit & ((W)((code >> 1) & 1) - 1);
                    b.b2 ^= bit & ((W)((code >> 2) & 1) - 1);
                    PPAD_P4(b.b3 ^= bit & ((W)((code >> 3) & 1) - 1))
                    d++;
                }
            }
        } while (e);
        src.cursor = d;
        src.blk_idx = -1;
        return;
    }
    while (e) {
        const W bit = e & (~e + 1);
        const int code = src.template next<G::NTYPES>(key, sky);
        if (!(code & 1)) b.b0 ^= bit;
        if (!(code & 2)) b.b1 ^= bit;
        if (!(code & 4)) b.b2 ^= bit;
        PPAD_P4(if (!(code & 8)) b.b3 ^= bit)
        e ^= bit;
    }
}

// fill_board(reset=True): every cell gets a draw, row-major.  TIGHT adds the unrolled pure-Philox loop: it pays
// in the kernels that reset all the time (rollout, reset: 0.49 -> 0.29 ms per 2^20 boards) but its code size
// slows the step kernels down (0.080 -> 0.087 ms) even though they rarely reset, so they do without it.
// UNROLL: Philox blocks per loop iteration of the TIGHT path.  2 is fastest where the refill dominates (reset and rollout
// kernels); the fused policy step, whose instruction footprint already overflows the instruction cache (ncu: no_instruction
// is its top stall), is faster with 1.
template <class G, bool TIGHT = PPAD_TIGHT_DEFAULT, int UNROLL = 2>
PPAD_HD void refill_all(Board<G> &b, OrbSource &src, const RngKey &key, const SkyfallTable &sky)
{
    typedef typename G::W W;
    W p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    if (TIGHT && src.n_inj < 0) {
        // pure Philox (the auto-reset of rollouts): one block per four draws, words taken by compile-time index
        const uint32_t blk0 = (uint32_t)src.cursor >> 2;
        const int l0 = src.cursor & 3;             // the first block may start in the middle
        const int nblk = (l0 + G::cells() + 3) / 4;
#pragma unroll UNROLL
        for (int q = 0; q < nblk; q++) {
            U4 c;
            c.x = key.env_lo;
            c.y = key.env_hi;
            c.z = key.step;
            c.w = (src.stream << 24) | (blk0 + (uint32_t)q);
            const U4 blk = philox4x32_10(c, key.k0, key.k1);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = 4 * q + k - l0;
                if (i >= 0 && i < G::cells()) {
                    const int code = orb_from_u32<G::NTYPES>(sky, k == 0 ? blk.x : (k == 1 ? blk.y : (k == 2 ? blk.z : blk.w)));
                    p0 |= (W)(code & 1) << i;
                    p1 |= (W)((code >> 1) & 1) << i;
                    p2 |= (W)((code >> 2) & 1) << i;
                    PPAD_P4(p3 |= (W)((code >> 3) & 1) << i)
                }
            }
        }
        src.cursor += G::cells();
        src.blk_idx = -1;
    } else {
        for (int i = 0; i < G::cells(); i++) {
            const int code = src.template next<G::NTYPES>(key, sky);
            p0 |= (W)(code & 1) << i;
            p1 |= (W)((code >> 1) & 1) << i;
            p2 |= (W)((code >> 2) & 1) << i;
            PPAD_P4(p3 |= (W)((code >> 3) & 1) << i)
        }
    }
    b.b0 = p0;
    b.b1 = p1;
    b.b2 = p2;
    PPAD_P4(b.b3 = p3)
}

// ---------------------------------------------------------------------------------------------
// calculate_reward (game.py:364-405) on a copy of the board
// ---------------------------------------------------------------------------------------------
struct LogSink {
    uint16_t *log;   // colour | N << 8 per combo, may be nullptr
    int cap, n;
    PPAD_HD void put(int colour, int N)
    {
        if (log && n < cap) log[n] = (uint16_t)(colour | (N << 8));
        n++;
    }
};

// The state one board carries from one cascade iteration to the next.  The kernel parks it in shared
// memory between iterations so that boards still cascading can be re-packed into dense warps.
template <class G>
struct CascadeItem {
    Board<G> b;
    Tally tally;
    int depth;
    int cursor;        // skyfall orbs drawn so far
    bool exhausted;    // the injected slab ran out
    uint32_t flags;

    PPAD_HD void init(const Board<G> &board, uint32_t flags0)
    {
        b = board;
        tally.init();
        depth = 0;
        cursor = 0;
        exhausted = false;
        flags = flags0;
    }
    PPAD_HD double reward() const { return tally.n_combos ? tally.total() : 0.0; }
    PPAD_HD uint32_t info() const
    {
        return make_info(tally.n_combos, depth, cursor, flags | (exhausted ? (uint32_t)FLAG_SKYFALL_EXHAUSTED : 0u));
    }
};

// One iteration of the loop of calculate_reward (game.py:371-397), given the match mask m != 0 (and the edge
// masks) of it.b: cancel; stop if cascades are off or at the cap; else drop, fill and look for matches again.
// Returns true when the refilled board matches again, with its masks in m / er / ed for the next iteration.
template <class G>
PPAD_HD bool cascade_round_m(CascadeItem<G> &it, typename G::W &m, bool skyfall_damage, int cascade_cap, const TeamTable &team, const uint32_t *slab, int n_inj,
                             const RngKey &key, const SkyfallTable &sky, uint16_t *log, int log_cap, int slab_fmt = SLAB_NIBBLES,
                             const uint32_t *slab_eager = nullptr, int slab_n_eager = 0, int slab_stride = 0,
                             int slab_estride = 1)
{
    cancel_tally<G>(it.b, m, it.tally, team, log, log_cap);
    it.depth++;
    if (!skyfall_damage) return false;
    if (cascade_cap > 0 && it.depth >= cascade_cap) {
        it.flags |= FLAG_CASCADE_CAP;
        return false;
    }
    drop_board<G>(it.b);
    OrbSource src;
    src.init(slab, n_inj, STREAM_SKYFALL, slab_fmt, slab_eager, slab_n_eager, slab_stride, slab_estride);
    src.cursor = it.cursor;
    src.exhausted = it.exhausted;
    fill_board<G>(it.b, src, key, sky);
    it.cursor = src.cursor;
    it.exhausted = src.exhausted;
    if (src.bad) it.flags |= FLAG_UNSUPPORTED_ORB;
    typename G::W er, ed;
    m = match_mask<G>(it.b, er, ed);
    return m != 0;
}

// damage (game.py:157-199) of an explicit combo list: log[i] = colour | N << 8
PPAD_HD double damage_of_log(const TeamTable &team, const uint16_t *log, int n)
{
    Tally tally;
    tally.init();
    for (int i = 0; i < n; i++) tally.add(team, log[i] & 0xFF, log[i] >> 8);
    return tally.n_combos ? tally.total() : 0.0;
}

// ---------------------------------------------------------------------------------------------
// moves, legality, random policy
// ---------------------------------------------------------------------------------------------
// player.py:52-62: bit a set = action a allowed (not the reverse of prev, not off the board)
template <class G>
PPAD_HD int legal_mask(int finger, int prev)
{
    const int r = finger / G::cols(), c = finger - r * G::cols();
    int m = 15;
    if (r <= 0) m &= ~1;
    if (r >= G::rows() - 1) m &= ~2;
    if (c <= 0) m &= ~4;
    if (c >= G::cols() - 1) m &= ~8;
    if (prev < 4) m &= ~(1 << (prev ^ 1));   // opposite of up(0)/down(1)/left(2)/right(3)
    return m;
}

// index of the pick-th set bit of a 4-bit mask, as a table of 16 masks x 4 picks x 2 bits in four words
PPAD_HDC uint32_t nth_bit_lut_word(int w)
{
    uint32_t v = 0;
    for (int e = 0; e < 16; e++) {                 // entry 16 w + e = mask * 4 + pick
        const int mask = (16 * w + e) >> 2, pick = (16 * w + e) & 3;
        int seen = 0, a = 0;
        for (int i = 0; i < 4; i++)
            if ((mask >> i) & 1) {
                if (seen == pick) a = i;
                seen++;
            }
        v |= (uint32_t)a << (2 * e);
    }
    return v;
}

// uniform over the legal set (bits 0..4; bit 4 = pass when the caller allows it), one Philox draw
// (the action stream's word of this step); -1 when the set is empty
PPAD_HD int sample_from_mask(int mask, uint32_t x)
{
    constexpr uint32_t L0 = nth_bit_lut_word(0), L1 = nth_bit_lut_word(1), L2 = nth_bit_lut_word(2), L3 = nth_bit_lut_word(3);
    const int n = popc((uint32_t)mask);
    if (n == 0) return -1;
    const int pick = (int)mulhi32(x, (uint32_t)n);
    if ((mask & 16) && pick == n - 1) return ACT_PASS;   // pass is the highest bit, so the last pick
    const int e = (mask & 15) * 4 + pick;                  // pick <= 3 here
    const uint32_t w = (e & 32) ? ((e & 16) ? L3 : L2) : ((e & 16) ? L1 : L0);
    return (int)((w >> (2 * (e & 15))) & 3u);
}

// apply_action (game.py:122-155): false = 'rejected' (off the board), state untouched
template <class G>
PPAD_HD bool apply_move(State<G> &s, int action)
{
    // straight-line: up = row - 1, down = row + 1, left = col - 1, right = col + 1 (game.py:127-155)
    const int f = meta_finger(s.meta);
    const int r = f / G::cols(), c = f - r * G::cols();
    const int dr = (action == ACT_DOWN ? 1 : 0) - (action == ACT_UP ? 1 : 0);
    const int dc = (action == ACT_RIGHT ? 1 : 0) - (action == ACT_LEFT ? 1 : 0);
    const int nr = r + dr, nc = c + dc;
    if ((unsigned)nr >= (unsigned)G::rows() || (unsigned)nc >= (unsigned)G::cols()) return false;
    const int t = nr * G::cols() + nc;
    swap_cells<G>(s.b, f, t);
    int moves = meta_moves(s.meta);
    if (moves < 0xFFFF) moves++;
    s.meta = meta_pack(t, action, moves);
    return true;
}

// random finger of a reset (game.py:311-315): draws 0 and 1 of the FINGER stream = words x, y of its block 0
template <class G>
PPAD_HD int random_finger(const RngKey &key)
{
    U4 c;
    c.x = key.env_lo;
    c.y = key.env_hi;
    c.z = key.step;
    c.w = STREAM_FINGER << 24;
    const U4 o = philox4x32_10(c, key.k0, key.k1);
    return (int)mulhi32(o.x, (uint32_t)G::rows()) * G::cols() + (int)mulhi32(o.y, (uint32_t)G::cols());
}

// reset (game.py:303-315): refill the whole board until it holds no 3-run
template <class G>
PPAD_HD uint32_t reset_board(State<G> &s, int reset_cap, OrbSource &src, const RngKey &key, const SkyfallTable &sky,
                             int finger_in)
{
    typedef typename G::W W;
    uint32_t flags = 0;
    int attempts = 0;
    for (;;) {
        refill_all<G>(s.b, src, key, sky);
        attempts++;
        W er, ed;
        if (match_mask<G>(s.b, er, ed) == 0) break;
        if (reset_cap > 0 && attempts >= reset_cap) {
            flags |= FLAG_RESET_CAP;
            break;
        }
    }
    if (src.exhausted) flags |= FLAG_SKYFALL_EXHAUSTED;
    int finger = finger_in;
    if (finger < 0) finger = random_finger<G>(key);
    s.meta = meta_pack(finger, ACT_PASS, 0);   // previous_action = 'pass' (game.py:299)
    return flags;
}

// ---------------------------------------------------------------------------------------------
// one env step of one board (the kernel body; see include/ppad_b200.h: ppad_step)
// ---------------------------------------------------------------------------------------------
struct StepParams {
    RngKey key;              // env_lo/env_hi are overwritten per board
    SkyfallTable sky;
    TeamTable team;
    int32_t skyfall_damage, cascade_cap, reset_cap;
    int32_t eval_moves, auto_reset, max_moves;
    int32_t n_inj, slab_words, log_cap;
    int32_t draw0;   // first draw index of a stand-alone fill (ppad_fill)
    int32_t slab_fmt;   // SLAB_NIBBLES / SLAB_PLANES3 / SLAB_CHUNKS10: layout of the injected skyfall slab of this launch
    int32_t slab_stride;   // SLAB_CHUNKS10: words between consecutive chunks of one board
    int32_t slab_eager;    // SLAB_CHUNKS10: chunks per board the kernel fetches up front (1..4)
    int32_t slab_estride;  // elements between the eager chunks of one board where the kernel keeps them (1; tile-major: 256)
    DynGeomData dyn;    // board dimensions for GeomDyn (ignored by the specialised geometries)
};

struct StepOut {
    double reward;
    uint32_t info;
    int action;
    int consumed;   // skyfall orbs drawn by the cascade, not saturated (the info byte stops at 255)
};

struct StepPrefix {
    int action;
    uint32_t flags;
    bool evaluate;
};

// Everything of one env step except the cascade: pick the action, apply the move, decide whether the
// board is evaluated.  `s` is updated in place (a pass leaves it alone, game.py:366).
template <class G>
PPAD_HD StepPrefix step_prefix(State<G> &s, const StepParams &p, const RngKey &key, int action_in,
                               const U4 *act_blk = nullptr)   // action_block(key) if the caller has it at hand
{
    StepPrefix o;
    o.flags = 0;
    o.evaluate = false;
    int action = action_in;   // < 0: sample
    if (action < 0) {
        if (p.max_moves > 0 && meta_moves(s.meta) >= p.max_moves) {
            action = ACT_PASS;   // episode length cap (simulation.py:50-52)
        } else {
            action = sample_from_mask(legal_mask<G>(meta_finger(s.meta), meta_prev(s.meta)),
                                      action_word(act_blk ? *act_blk : action_block(key), key.step));
            if (action < 0) {
                action = 255;
                o.flags |= FLAG_NO_LEGAL_MOVE;
            }
        }
    }
    if (action == ACT_PASS) {
        o.evaluate = true;
    } else if (action < ACT_PASS) {
        if (apply_move<G>(s, action)) o.evaluate = p.eval_moves != 0;
        else o.flags |= FLAG_REJECTED;
    } else if (action == ACT_NOOP) {
        // the board sits this step out (a finished episode in a lock-step batch): no move, no evaluation, no flag
    } else if (!(o.flags & FLAG_NO_LEGAL_MOVE)) {
        o.flags |= FLAG_BAD_ACTION;
    }
    o.action = action;
    return o;
}

// A whole finger path in one go (BASELINE configs[2]: "max-length finger paths"): len moves, 2 bits each,
// 16 per uint32 (move t in word t >> 4 at bits 2 * (t & 15)), applied like len calls of step(move);
// off-board moves are skipped and flagged.  The board is evaluated afterwards when eval_moves is set.
template <class G>
PPAD_HD StepPrefix path_prefix(State<G> &s, const StepParams &p, const uint32_t *path, int len)
{
    StepPrefix o;
    o.flags = 0;
    o.action = 255;
    uint32_t word = 0;
    for (int t = 0; t < len; t++) {
        if ((t & 15) == 0) word = path[t >> 4];
        const int a = (int)((word >> (2 * (t & 15))) & 3u);
        if (apply_move<G>(s, a)) o.action = a;
        else o.flags |= FLAG_REJECTED;
    }
    o.evaluate = p.eval_moves != 0;
    return o;
}

// PAD.reset after a processed pass (auto_reset): returns the extra flags
template <class G>
PPAD_HD uint32_t step_auto_reset(State<G> &s, const StepParams &p, const RngKey &key)
{
    OrbSource src;
    src.init(nullptr, -1, STREAM_RESET);
    return FLAG_EPISODE_DONE | reset_board<G>(s, p.reset_cap, src, key, p.sky, -1);
}

PPAD_HD RngKey board_key(const StepParams &p, uint64_t env)
{
    RngKey key = p.key;
    key.env_lo = (uint32_t)env;
    key.env_hi = (uint32_t)(env >> 32);
    return key;
}

#if defined(__CUDACC__)
// reset for a whole warp (device only; EVERY lane must call it, `need` says whose board is reset).
// Rejection sampling is geometric (about 35 % of random 5x6 boards are combo-free), so 32 lanes retrying
// on their own wait for the unluckiest one (~10 attempts).  Because the orb stream is counter based, ANY
// lane can generate attempt j of ANY board: the lanes are dealt round-robin over the boards that still need
// a board and try consecutive attempt indices speculatively; per board the accepted attempt with the LOWEST
// index wins, which is exactly the one the sequential loop of game.py:303-308 would have stopped at -- same
// board, same number of orbs consumed.  Typically 3-4 warp iterations instead of ~10.
template <class G, bool TIGHT = false, int UNROLL = 2>
__device__ __forceinline__ uint32_t reset_boards_warp(bool need, State<G> &s, const StepParams &p, uint32_t step,
                                                      uint64_t env, const uint32_t *slab, int n_inj, int finger_in,
                                                      int *consumed_out)
{
    typedef typename G::W W;
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    RngKey my_key = board_key(p, env);
    my_key.step = step;
    const int cap = p.reset_cap;
    const int my_n_inj = slab ? n_inj : -1;
    int tried = 0;             // attempts of MY board already generated (by anyone)
    uint32_t flags = 0;
    int consumed = 0;
    unsigned needy = __ballot_sync(full, need);
    while (needy) {
        const int cnt = __popc(needy);
        const int k = lane % cnt, h = lane / cnt;                 // I help the k-th needy board with its h-th next attempt
        const int owner = __fns(needy, 0, k + 1);
        const unsigned long long env_o = __shfl_sync(full, (unsigned long long)env, owner);
        const unsigned long long slab_o = __shfl_sync(full, (unsigned long long)(uintptr_t)slab, owner);
        const int att = __shfl_sync(full, tried, owner) + h;
        const bool in_range = cap <= 0 || att < cap;
        Board<G> b;
        b.b0 = b.b1 = b.b2 = 0;
        PPAD_P4(b.b3 = 0)
        bool ok = false;
        if (in_range) {
            OrbSource src;
            src.init(reinterpret_cast<const uint32_t *>((uintptr_t)slab_o), n_inj, STREAM_RESET);   // init() ignores n_inj without a slab
            src.cursor = att * G::cells();
            RngKey key_o = board_key(p, env_o);
            key_o.step = step;
            refill_all<G, TIGHT, UNROLL>(b, src, key_o, p.sky);
            W er, ed;
            ok = match_mask<G>(b, er, ed) == 0;
        }
        const bool last = cap > 0 && att == cap - 1;              // the sequential loop gives up here and keeps the board
        const bool accept = in_range && (ok || last);
        // among the lanes that accept for the same owner, the lowest lane holds the lowest attempt index
        const unsigned grp = __match_any_sync(full, accept ? owner : 32 + lane);
        unsigned winners = __ballot_sync(full, accept && (__ffs(grp) - 1 == lane));
        bool got = false;
        while (winners) {
            const int w = __ffs(winners) - 1;
            winners &= winners - 1;
            const int o = __shfl_sync(full, owner, w);
            const W w0 = __shfl_sync(full, b.b0, w), w1 = __shfl_sync(full, b.b1, w), w2 = __shfl_sync(full, b.b2, w);
            W w3 = 0;
            PPAD_P4(w3 = __shfl_sync(full, b.b3, w))
            const int a_w = __shfl_sync(full, att, w);
            const int ok_w = __shfl_sync(full, (int)ok, w);
            if (lane == o) {
                s.b.b0 = w0;
                s.b.b1 = w1;
                s.b.b2 = w2;
                PPAD_P4(s.b.b3 = w3)
                consumed = (a_w + 1) * G::cells();
                if (!ok_w) flags |= FLAG_RESET_CAP;
                got = true;
            }
        }
        if (need && got) need = false;
        else if (need) tried += 32 / cnt + ((int)__popc(needy & ((1u << lane) - 1u)) < 32 % cnt ? 1 : 0);
        needy = __ballot_sync(full, need);
    }
    if (consumed > 0) {
        if (my_n_inj >= 0 && consumed > my_n_inj) flags |= FLAG_SKYFALL_EXHAUSTED;
        int finger = finger_in;
        if (finger < 0) finger = random_finger<G>(my_key);
        s.meta = meta_pack(finger, ACT_PASS, 0);
    }
    if (consumed_out) *consumed_out = consumed;
    return flags;
}
#endif

// The whole step of one board.  On the GPU EVERY lane of the warp must call this (lanes past the end of the
// batch pass live = false): the cascade loop is warp-uniform -- it runs while any lane still cascades and
// the lanes reconverge after every iteration.
// TIGHT_RESET: the auto-reset uses the unrolled Philox refill; HAS_RESET = false compiles the auto-reset out (kernels
// launched with auto_reset = 0 do not carry its code).
template <class G, bool TIGHT_RESET = false, bool HAS_RESET = true, int RESET_UNROLL = 2>
PPAD_HD StepOut step_board(bool live, State<G> &s, Board<G> &final_board, const StepParams &p, uint64_t env,
                           int action_in, const uint32_t *slab, uint16_t *log, const uint32_t *path = nullptr,
                           int path_len = 0, uint32_t step_offset = 0, const U4 *act_blk = nullptr,
                           const uint32_t *slab_eager = nullptr)
{
    RngKey key = board_key(p, env);
    key.step += step_offset;   // fused multi-step rollouts: step index = p.key.step + t
    StepPrefix pre;
    pre.action = 255;
    pre.flags = 0;
    pre.evaluate = false;
    if (live) pre = path ? path_prefix<G>(s, p, path, path_len) : step_prefix<G>(s, p, key, action_in, act_blk);
    warp_converge();
    if (!warp_any(live && (pre.evaluate || pre.action == ACT_PASS))) {
        // a plain move for the whole warp (episode mode: 50 of 51 steps): no cascade, no reset, nothing to tally
        final_board = s.b;
        StepOut quick;
        quick.action = pre.action;
        quick.reward = 0.0;
        quick.info = make_info(0, 0, 0, pre.flags);
        quick.consumed = 0;
        return quick;
    }
    CascadeItem<G> it;
    it.init(s.b, pre.flags);
    bool cascading = live && pre.evaluate;
    typename G::W m = 0, er = 0, ed = 0;
    if (cascading) {
        m = match_mask<G>(it.b, er, ed);
        cascading = m != 0;
    }
    while (warp_any(cascading)) {
        if (cascading)
            cascading = cascade_round_m<G>(it, m, p.skyfall_damage != 0, p.cascade_cap, p.team, slab, p.n_inj,
                                           key, p.sky, log, p.log_cap, p.slab_fmt, slab_eager, p.slab_eager, p.slab_stride,
                                           p.slab_estride);
        warp_converge();
    }
#if defined(__CUDA_ARCH__)
    if constexpr (HAS_RESET) {
        if (p.auto_reset) {   // kernel-uniform
            const bool need = live && pre.action == ACT_PASS;
            const uint32_t f = reset_boards_warp<G, TIGHT_RESET, RESET_UNROLL>(need, s, p, key.step, env, nullptr, -1, -1, nullptr);
            if (need) it.flags |= FLAG_EPISODE_DONE | f;
        }
    }
#else
    if (live && pre.action == ACT_PASS && p.auto_reset) it.flags |= step_auto_reset<G>(s, p, key);
#endif
    warp_converge();
    final_board = it.b;
    StepOut out;
    out.action = pre.action;
    out.reward = it.reward();
    out.info = it.info();
    out.consumed = it.cursor;
    return out;
}

// ---------------------------------------------------------------------------------------------
// one-hot observation as a bit mask (agent01.py:254-273): bit 7 * cell + channel of w[], channels
// 0..5 = (orb type == channel), channel 6 = finger.  Built a board row at a time: the C cells of colour
// plane k in row r are spread to stride 7 by ONE widening multiply -- bit j of x = rowbits << k lands at
// j + k + 6 m for every term 2^(6 m) of 0x41041041, and m = j is the wanted 7 j + k; the other products
// fall on bits the mask drops, and no two products share a bit (j < 6), so there are no carries.
// w must be zero-initialised and hold (7 * CELLS + 31) / 32 + 1 words.
// ---------------------------------------------------------------------------------------------
PPAD_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, int s)   // high word of (hi:lo) << s, 0 <= s < 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}

template <class G, int NW>
PPAD_HD void obs_mask(const State<G> &s, uint32_t (&w)[NW])
{
    typedef typename G::W W;
    constexpr int C = G::cols(), RB = 7 * C;                       // obs bits per board row
    constexpr int CM = C < 6 ? C : 6;                         // columns handled by the multiply
    W plane[7];
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
        plane[k] = ((k & 1) ? s.b.b0 : ~s.b.b0) & ((k & 2) ? s.b.b1 : ~s.b.b1) & ((k & 4) ? s.b.b2 : ~s.b.b2) & G::board();
        PPAD_P4(plane[k] &= ~s.b.b3)   // types 0..5 have bit 3 clear; 6..9 and empty give all-zero colour channels
    }
    plane[6] = G::ONE << meta_finger(s.meta);                 // channel 6 = finger
#pragma unroll
    for (int r = 0; r < G::rows(); r++) {
        uint64_t row = 0;
#pragma unroll
        for (int k = 0; k < 7; k++) {
            uint64_t keep = 0;
#pragma unroll
            for (int j = 0; j < CM; j++) keep |= (uint64_t)1 << (7 * j + k);
            const uint32_t x = (uint32_t)((plane[k] >> (r * C)) & ((1u << CM) - 1)) << k;
            row |= ((uint64_t)x * 0x41041041ull) & keep;
            if (C > 6) row |= (uint64_t)((plane[k] >> (r * C + 6)) & 1) << (42 + k);
        }
        const int o = RB * r, wi = o >> 5, sft = o & 31;
        const uint32_t lo = (uint32_t)row, hi = (uint32_t)(row >> 32);
        w[wi] |= lo << sft;
        w[wi + 1] |= funnel_l(lo, hi, sft);
        if (sft + RB > 64) w[wi + 2] |= hi >> (32 - sft);
    }
}

// ---------------------------------------------------------------------------------------------
// conversions between plain int boards and packed state (value -1 -> 7; 0..6 kept; others flagged)
// ---------------------------------------------------------------------------------------------
template <class G, class T>
PPAD_HD uint32_t pack_cells(Board<G> &b, const T *cells)
{
    typedef typename G::W W;
    W p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    uint32_t flags = 0;
    for (int i = 0; i < G::cells(); i++) {
        const long long v = (long long)cells[i];
        int code;
        if (v == -1) code = G::EMPTY;
        else if (v >= 0 && v < G::NTYPES) code = (int)v;
        else {
            code = G::EMPTY;
            flags |= FLAG_UNSUPPORTED_ORB;
        }
        p0 |= (W)(code & 1) << i;
        p1 |= (W)((code >> 1) & 1) << i;
        p2 |= (W)((code >> 2) & 1) << i;
        PPAD_P4(p3 |= (W)((code >> 3) & 1) << i)
    }
    b.b0 = p0;
    b.b1 = p1;
    b.b2 = p2;
    PPAD_P4(b.b3 = p3)
    return flags;
}

template <class G, class T>
PPAD_HD void unpack_cells(const Board<G> &b, T *cells)
{
    for (int i = 0; i < G::cells(); i++) {
        const int code = code_at<G>(b, i);
        cells[i] = (T)(code == G::EMPTY ? -1 : code);
    }
}

}   // namespace ppad
