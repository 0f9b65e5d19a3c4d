// ppad_device.cuh -- what every translation unit of libppad_b200.so shares: packed state I/O, the warp-cooperative
// one-hot observation emitter, the context object, the geometry dispatch and the host-side error plumbing.
//
// Thread-per-board: a warp owns 32 consecutive boards, state I/O is one (or two) 16-byte vector
// accesses per board, fully coalesced.  The one-hot observation -- the only large stream, 840 B
// per 5x6 board against 16 B of state -- is emitted warp-cooperatively: every lane builds the
// 210-bit occupancy mask of its board, the warp lays the 32 masks end to end in shared memory as one
// 6720-bit stream, and then each lane expands 4 bits at a time through a 16-entry float4 table into
// 16-byte streaming stores to consecutive addresses (512 contiguous bytes per warp instruction).
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC (one object per
// translation unit, see __graft_entry__.build)
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/ppad_b200.h"
#include "ppad_core.cuh"
#include "ppad_agent.cuh"
#include "ppad_params.h"

#if !defined(PPAD_OBS_UNROLL)
#define PPAD_OBS_UNROLL 4
#endif

namespace ppad {

constexpr int BLOCK = 256;

// ------------------------------------------------------------------------------------------------
// packed state I/O.  State buffers are read AND written by the same kernels (in place), so the loads are plain
// coherent L2 loads (ld.global.cg), not the read-only path (ld.global.nc is only defined for data nobody writes
// during the kernel); read-only inputs (slab, actions, Q-values) keep __ldg.
// ------------------------------------------------------------------------------------------------
template <class G, bool WIDE = (sizeof(typename G::W) == 8), int NP = G::NP>
struct StateIO;

template <class G>
__device__ __forceinline__ void clear_state(State<G> &s)
{
    s.b.b0 = s.b.b1 = s.b.b2 = 0;
    PPAD_P4(s.b.b3 = 0)
    s.meta = 0;
}

// 32-bit planes, 3-bit codes: {b0, b1, b2, meta} = one uint4
template <class G>
struct StateIO<G, false, 3> {
    static __device__ __forceinline__ State<G> load(const void *base, int64_t i)
    {
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(base) + i);
        State<G> s;
        s.b.b0 = v.x;
        s.b.b1 = v.y;
        s.b.b2 = v.z;
        s.meta = v.w;
        return s;
    }
    static __device__ __forceinline__ void store(void *base, int64_t i, const Board<G> &b, uint32_t meta)
    {
        reinterpret_cast<uint4 *>(base)[i] = make_uint4(b.b0, b.b1, b.b2, meta);
    }
    // the board planes only; the meta word of the record is left as it is
    static __device__ __forceinline__ void store_planes(void *base, int64_t i, const Board<G> &b)
    {
        uint32_t *p = reinterpret_cast<uint32_t *>(base) + 4 * i;
        p[0] = b.b0;
        p[1] = b.b1;
        p[2] = b.b2;
    }
};

// 64-bit planes, 3-bit codes: {b0, b1}, {b2, meta, 0} = two uint4
template <class G>
struct StateIO<G, true, 3> {
    static __device__ __forceinline__ State<G> load(const void *base, int64_t i)
    {
        const uint4 lo = __ldcg(reinterpret_cast<const uint4 *>(base) + 2 * i);
        const uint4 hi = __ldcg(reinterpret_cast<const uint4 *>(base) + 2 * i + 1);
        State<G> s;
        s.b.b0 = (uint64_t)lo.x | ((uint64_t)lo.y << 32);
        s.b.b1 = (uint64_t)lo.z | ((uint64_t)lo.w << 32);
        s.b.b2 = (uint64_t)hi.x | ((uint64_t)hi.y << 32);
        s.meta = hi.z;
        return s;
    }
    static __device__ __forceinline__ void store(void *base, int64_t i, const Board<G> &b, uint32_t meta)
    {
        uint4 *p = reinterpret_cast<uint4 *>(base) + 2 * i;
        p[0] = make_uint4((uint32_t)b.b0, (uint32_t)(b.b0 >> 32), (uint32_t)b.b1, (uint32_t)(b.b1 >> 32));
        p[1] = make_uint4((uint32_t)b.b2, (uint32_t)(b.b2 >> 32), meta, 0u);
    }
    static __device__ __forceinline__ void store_planes(void *base, int64_t i, const Board<G> &b)
    {
        uint64_t *p = reinterpret_cast<uint64_t *>(base) + 4 * i;
        p[0] = b.b0;
        p[1] = b.b1;
        p[2] = b.b2;
    }
};

// 32-bit planes, 4-bit codes: {b0, b1, b2, b3}, {meta, 0, 0, 0} = two uint4
template <class G>
struct StateIO<G, false, 4> {
    static __device__ __forceinline__ State<G> load(const void *base, int64_t i)
    {
        const uint4 lo = __ldcg(reinterpret_cast<const uint4 *>(base) + 2 * i);
        const uint4 hi = __ldcg(reinterpret_cast<const uint4 *>(base) + 2 * i + 1);
        State<G> s;
        s.b.b0 = lo.x;
        s.b.b1 = lo.y;
        s.b.b2 = lo.z;
        s.b.b3 = lo.w;
        s.meta = hi.x;
        return s;
    }
    static __device__ __forceinline__ void store(void *base, int64_t i, const Board<G> &b, uint32_t meta)
    {
        uint4 *p = reinterpret_cast<uint4 *>(base) + 2 * i;
        p[0] = make_uint4(b.b0, b.b1, b.b2, b.b3);
        p[1] = make_uint4(meta, 0u, 0u, 0u);
    }
    static __device__ __forceinline__ void store_planes(void *base, int64_t i, const Board<G> &b)
    {
        reinterpret_cast<uint4 *>(base)[2 * i] = make_uint4(b.b0, b.b1, b.b2, b.b3);
    }
};

// 64-bit planes, 4-bit codes: {b0, b1}, {b2, b3}, {meta, 0, 0, 0} = three uint4
template <class G>
struct StateIO<G, true, 4> {
    static __device__ __forceinline__ State<G> load(const void *base, int64_t i)
    {
        const uint4 *p = reinterpret_cast<const uint4 *>(base) + 3 * i;
        const uint4 a = __ldcg(p), c = __ldcg(p + 1), d = __ldcg(p + 2);
        State<G> s;
        s.b.b0 = (uint64_t)a.x | ((uint64_t)a.y << 32);
        s.b.b1 = (uint64_t)a.z | ((uint64_t)a.w << 32);
        s.b.b2 = (uint64_t)c.x | ((uint64_t)c.y << 32);
        s.b.b3 = (uint64_t)c.z | ((uint64_t)c.w << 32);
        s.meta = d.x;
        return s;
    }
    static __device__ __forceinline__ void store(void *base, int64_t i, const Board<G> &b, uint32_t meta)
    {
        uint4 *p = reinterpret_cast<uint4 *>(base) + 3 * i;
        p[0] = make_uint4((uint32_t)b.b0, (uint32_t)(b.b0 >> 32), (uint32_t)b.b1, (uint32_t)(b.b1 >> 32));
        p[1] = make_uint4((uint32_t)b.b2, (uint32_t)(b.b2 >> 32), (uint32_t)b.b3, (uint32_t)(b.b3 >> 32));
        p[2] = make_uint4(meta, 0u, 0u, 0u);
    }
    static __device__ __forceinline__ void store_planes(void *base, int64_t i, const Board<G> &b)
    {
        uint4 *p = reinterpret_cast<uint4 *>(base) + 3 * i;
        p[0] = make_uint4((uint32_t)b.b0, (uint32_t)(b.b0 >> 32), (uint32_t)b.b1, (uint32_t)(b.b1 >> 32));
        p[1] = make_uint4((uint32_t)b.b2, (uint32_t)(b.b2 >> 32), (uint32_t)b.b3, (uint32_t)(b.b3 >> 32));
    }
};

// ------------------------------------------------------------------------------------------------
// one-hot observation emission (agent01.py:254-273), warp-cooperative
// ------------------------------------------------------------------------------------------------
template <class G>
struct ObsGeom {
    static constexpr int NB = G::MAX_CELLS * 7;      // bits (= elements) per board (GeomDyn: the most a board can have)
    static constexpr int NW = (NB + 31) / 32 + 1;    // words of a shifted per-board mask
    static constexpr int CNT_LO = NB / 32;           // a board owns CNT_LO or CNT_LO + 1 stream words
    static constexpr int WARP_WORDS = NB;            // 32 boards * NB bits / 32
    static_assert(NB % 2 == 0, "obs tail handling assumes an even element count per board");
};

struct ObsShared {
    float4 lut_f32[16];
    uint32_t lut_u8[16];
};

__device__ __forceinline__ void obs_init_luts(ObsShared &sh)
{
    if (threadIdx.x < 16) {
        const unsigned b = threadIdx.x;
        sh.lut_f32[b] = make_float4((b & 1) ? 1.f : 0.f, (b & 2) ? 1.f : 0.f, (b & 4) ? 1.f : 0.f, (b & 8) ? 1.f : 0.f);
        sh.lut_u8[b] = (b & 1) | ((b & 2) << 7) | ((b & 4) << 14) | ((b & 8) << 21);
    }
}

// Every lane of the warp must call this (lanes without a board pass valid = false).
// `stream` is this warp's NB-word shared buffer; obs_warp points at the obs of the warp's first board.
template <class G>
__device__ __forceinline__ void emit_obs(const State<G> &s, bool valid, int n_valid, uint32_t *stream,
                                         const ObsShared &sh, void *obs_warp, int obs_dtype)
{
    typedef ObsGeom<G> OG;
    constexpr int NB = OG::NB, NW = OG::NW;
    const int lane = threadIdx.x & 31;

    // 1. this board's NB-bit mask: bit 7 * cell + channel
    uint32_t w[NW];
#pragma unroll
    for (int j = 0; j < NW; j++) w[j] = 0;
    if (valid) obs_mask<G, NW>(s, w);

    // 2. place it at bit NB * lane of the warp's stream; neighbours share one boundary word
    const int off = NB * lane, word0 = off >> 5, sft = off & 31;
    const int cnt = ((NB * (lane + 1)) >> 5) - word0;
    uint32_t t[NW];
    t[0] = w[0] << sft;
#pragma unroll
    for (int j = 1; j < NW; j++) t[j] = __funnelshift_l(w[j - 1], w[j], sft);
    const uint32_t tail = cnt == OG::CNT_LO ? t[OG::CNT_LO] : t[OG::CNT_LO + 1];
    uint32_t carry = __shfl_up_sync(0xFFFFFFFFu, tail, 1);
    if (lane == 0) carry = 0;
    t[0] |= carry;
#pragma unroll
    for (int j = 0; j <= OG::CNT_LO; j++)
        if (j < cnt) stream[word0 + j] = t[j];
    __syncwarp();

    // 3. expand: 4 stream bits -> one 16-byte store, consecutive lanes -> consecutive addresses
    const int valid_elems = n_valid * NB;
    if (obs_dtype == PPAD_OBS_F32) {
        float4 *out4 = reinterpret_cast<float4 *>(obs_warp);
        const int sh4 = (lane & 7) * 4;
        if (n_valid == 32) {
            // full warp: straight-line code, 5 instructions per 16-byte store (LDS, rotate, mask, LDS.128, STG.128)
            const char *lut = reinterpret_cast<const char *>(sh.lut_f32);
            const uint32_t *src = stream + (lane >> 3);
            float4 *dst = out4 + lane;
            const int rot = (sh4 + 28) & 31;                      // rotate right so the nibble sits at bits 4..7
            constexpr int FULL = (8 * NB) / 32, REST = (8 * NB) % 32;
            // unrolled by 4, not fully (52 iterations): the same speed in the plain step (0.1535 against 0.1541 ms) and
            // 6 % faster in the fused policy step, whose code no longer fits the instruction cache (profiles/r02f_ab.txt)
            PPAD_PRAGMA(unroll PPAD_OBS_UNROLL)
            for (int it = 0; it < FULL; it++) {
                const uint32_t idx16 = __funnelshift_r(src[4 * it], src[4 * it], rot) & 0xF0u;
                __stcs(dst + 32 * it, *reinterpret_cast<const float4 *>(lut + idx16));
            }
            if (REST && lane < REST) {
                const uint32_t idx16 = __funnelshift_r(src[4 * FULL], src[4 * FULL], rot) & 0xF0u;
                __stcs(dst + 32 * FULL, *reinterpret_cast<const float4 *>(lut + idx16));
            }
        } else {
            for (int q = lane; q < 8 * NB; q += 32) {
                const int e = 4 * q;
                if (e >= valid_elems) break;
                const float4 v = sh.lut_f32[(stream[q >> 3] >> sh4) & 15u];
                if (e + 4 <= valid_elems) __stcs(out4 + q, v);
                else __stcs(reinterpret_cast<float2 *>(out4 + q), make_float2(v.x, v.y));
            }
        }
    } else {
        uint4 *out4 = reinterpret_cast<uint4 *>(obs_warp);
        const int sh16 = (lane & 1) * 16;
        for (int q = lane; q < 2 * NB; q += 32) {
            const int e = 16 * q;
            if (e >= valid_elems) break;
            const uint32_t b16 = (stream[q >> 1] >> sh16) & 0xFFFFu;
            const uint4 v = make_uint4(sh.lut_u8[b16 & 15u], sh.lut_u8[(b16 >> 4) & 15u], sh.lut_u8[(b16 >> 8) & 15u],
                                       sh.lut_u8[b16 >> 12]);
            if (e + 16 <= valid_elems) {
                __stcs(out4 + q, v);
            } else {
                uint8_t *o = reinterpret_cast<uint8_t *>(out4 + q);
                for (int k = 0; k < valid_elems - e; k++) o[k] = (uint8_t)((b16 >> k) & 1u);
            }
        }
    }
    __syncwarp();
}

// The observation of a warp's boards for run-time board dimensions (GeomDyn): board by board, the planes of the board's
// lane are broadcast and the 32 lanes write its 7 * cells elements to consecutive addresses.  Every lane must call it.
template <class G>
__device__ __forceinline__ void emit_obs_dyn(const State<G> &s, bool valid, int n_valid, void *obs_warp, int obs_dtype)
{
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, ne = G::obs_elems();
    (void)valid;
    for (int b = 0; b < n_valid; b++) {
        const typename G::W p0 = __shfl_sync(full, s.b.b0, b), p1 = __shfl_sync(full, s.b.b1, b), p2 = __shfl_sync(full, s.b.b2, b);
        typename G::W p3 = 0;
        PPAD_P4(p3 = __shfl_sync(full, s.b.b3, b))
        const int finger = __shfl_sync(full, meta_finger(s.meta), b);
        for (int e = lane; e < ne; e += 32) {
            const int cell = e / 7, ch = e - 7 * cell;
            const int code = (int)(((p0 >> cell) & 1) | (((p1 >> cell) & 1) << 1) | (((p2 >> cell) & 1) << 2) | (((p3 >> cell) & 1) << 3));
            const bool on = ch < 6 ? code == ch : cell == finger;
            if (obs_dtype == PPAD_OBS_F32) reinterpret_cast<float *>(obs_warp)[(size_t)b * ne + e] = on ? 1.f : 0.f;
            else reinterpret_cast<uint8_t *>(obs_warp)[(size_t)b * ne + e] = on ? 1 : 0;
        }
    }
}

// the observation of the warp's boards, whichever geometry
template <class G>
__device__ __forceinline__ void emit_obs_any(const State<G> &s, bool valid, int n_valid, uint32_t *stream, const ObsShared &sh,
                                             void *obs_warp, int obs_dtype)
{
    if constexpr (G::DYNAMIC) emit_obs_dyn<G>(s, valid, n_valid, obs_warp, obs_dtype);
    else emit_obs<G>(s, valid, n_valid, stream, sh, obs_warp, obs_dtype);
}

template <class G>
__device__ __forceinline__ void *obs_warp_ptr(void *obs, int64_t warp_board0, int obs_dtype)
{
    const size_t elems = (size_t)warp_board0 * G::obs_elems();
    return obs_dtype == PPAD_OBS_F32 ? (void *)(reinterpret_cast<float *>(obs) + elems)
                                     : (void *)(reinterpret_cast<uint8_t *>(obs) + elems);
}

// ------------------------------------------------------------------------------------------------
// host side: errors, launch accounting, geometry dispatch
// ------------------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;      // defined in ppad_api.cu
extern std::atomic<uint64_t> g_launches;

inline int fail(int status, const std::string &msg)
{
    g_last_error = msg;
    return status;
}

inline int cuda_check(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return PPAD_OK;
    return fail(PPAD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

inline int after_launch(const char *name)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cuda_check(cudaGetLastError(), name);
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + BLOCK - 1) / BLOCK); }

typedef Geom<4, 5> G45;
typedef Geom<5, 6> G56;
typedef Geom<6, 7> G67;
typedef Geom<4, 5, 4> G45W;   // 4-bit orb codes: all ten orb types of game.py:236-246
typedef Geom<5, 6, 4> G56W;
typedef Geom<6, 7, 4> G67W;
typedef GeomDyn<3> GDYN;      // any dim_row x dim_col with <= 64 cells at run time
typedef GeomDyn<4> GDYNW;

// geometry ids of ppad_ctx::geom (build_params): 0 = 5x6, 1 = 6x7, 2 = 4x5; 3..5 = the same with 4-bit orb codes;
// 6 / 7 = run-time dimensions (GeomDyn) with 3- / 4-bit codes
template <int ID> struct GeomOf;
template <> struct GeomOf<0> { typedef G56 type; };
template <> struct GeomOf<1> { typedef G67 type; };
template <> struct GeomOf<2> { typedef G45 type; };
template <> struct GeomOf<3> { typedef G56W type; };
template <> struct GeomOf<4> { typedef G67W type; };
template <> struct GeomOf<5> { typedef G45W type; };
template <> struct GeomOf<6> { typedef GDYN type; };
template <> struct GeomOf<7> { typedef GDYNW type; };

}   // namespace ppad

struct ppad_ctx {
    ppad_config cfg;
    int device;
    int geom;   // 0 = 5x6, 1 = 6x7, 2 = 4x5; 3..5 = the same with 4-bit orb codes; 6 / 7 = run-time dimensions
    ppad::StepParams base;   // base.dyn: the board dimensions for the run-time geometries
    int sm_count;
    int pool_ctas_per_sm;   // resident CTAs of the persistent step_pool_kernel
    uint32_t *tile_counters;            // DEVICE [TILE_COUNTERS]: dynamic tile counters of the warp-pool kernel, one per launch
    std::atomic<uint32_t> *next_counter;   // round robin over them (launches of one context may overlap on several streams)
    int pool_kind;          // 1 = warp pool (default), 0 = CTA pool (env PPAD_POOL_KIND, for A/B runs)
};
constexpr int TILE_COUNTERS = 256;

#define PPAD_DISPATCH(ctx, EXPR)                                   \
    do {                                                           \
        if ((ctx)->geom == 0) { typedef ppad::G56 G; EXPR; }       \
        else if ((ctx)->geom == 1) { typedef ppad::G67 G; EXPR; }  \
        else if ((ctx)->geom == 2) { typedef ppad::G45 G; EXPR; }  \
        else if ((ctx)->geom == 3) { typedef ppad::G56W G; EXPR; } \
        else if ((ctx)->geom == 4) { typedef ppad::G67W G; EXPR; } \
        else if ((ctx)->geom == 5) { typedef ppad::G45W G; EXPR; } \
        else if ((ctx)->geom == 6) { typedef ppad::GDYN G; EXPR; } \
        else { typedef ppad::GDYNW G; EXPR; }                      \
    } while (0)

namespace ppad {

// Makes the context's device current for the duration of one API call and restores the caller's afterwards.
struct DeviceGuard {
    int prev, dev;
    cudaError_t err;
    explicit DeviceGuard(int device) : prev(-1), dev(device), err(cudaSuccess)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
        if (prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};

inline int check_enter(const ppad_ctx *ctx, int64_t n)
{
    if (!ctx) return fail(PPAD_ERR_INVALID, "null context");
    if (n < 0 || n > ((int64_t)1 << 40)) return fail(PPAD_ERR_INVALID, "bad batch size");
    return PPAD_OK;
}

}   // namespace ppad

// first statement of every compute entry point: validates, then switches to the context's device until the call returns
#define PPAD_ENTER(ctx, n)                                   \
    if (int r_ = ppad::check_enter((ctx), (n))) return r_;   \
    ppad::DeviceGuard guard_((ctx)->device);                 \
    if (guard_.err != cudaSuccess) return ppad::cuda_check(guard_.err, "cudaSetDevice")
