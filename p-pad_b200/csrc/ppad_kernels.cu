// ppad_kernels.cu -- sm_100a kernels and the C ABI of include/ppad_b200.h.
//
// Thread-per-board: a warp owns 32 consecutive boards, state I/O is one (or two) 16-byte vector
// accesses per board, fully coalesced.  The one-hot observation -- the only large stream, 840 B
// per 5x6 board against 16 B of state -- is emitted warp-cooperatively: every lane builds the
// 210-bit occupancy mask of its board, the warp lays the 32 masks end to end in shared memory as one
// 6720-bit stream, and then each lane expands 4 bits at a time through a 16-entry float4 table into
// 16-byte streaming stores to consecutive addresses (512 contiguous bytes per warp instruction).
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared
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

using namespace ppad;

namespace {

constexpr int BLOCK = 256;

// ------------------------------------------------------------------------------------------------
// packed state I/O
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
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(base) + i);
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
        const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(base) + 2 * i);
        const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(base) + 2 * i + 1);
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
        const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(base) + 2 * i);
        const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(base) + 2 * i + 1);
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
        const uint4 a = __ldg(p), c = __ldg(p + 1), d = __ldg(p + 2);
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
    static constexpr int NB = G::OBS_ELEMS;          // bits (= elements) per board
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
#pragma unroll
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

template <class G>
__device__ __forceinline__ void *obs_warp_ptr(void *obs, int64_t warp_board0, int obs_dtype)
{
    const size_t elems = (size_t)warp_board0 * G::OBS_ELEMS;
    return obs_dtype == PPAD_OBS_F32 ? (void *)(reinterpret_cast<float *>(obs) + elems)
                                     : (void *)(reinterpret_cast<uint8_t *>(obs) + elems);
}

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
    void *obs;
    int32_t obs_dtype;
    const uint32_t *paths;
    const uint16_t *path_lens;
    int32_t path_words, path_len;
    int32_t stage_slab;   // 1: the slab is pinned HOST memory (zero copy), 4 words per board: fetch each row once, up front
};

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
template <class G, bool OBS, bool AR>
__global__ void __launch_bounds__(BLOCK, (sizeof(typename G::W) == 8 && G::NP == 4) ? 3 : 4) step_obs_kernel(const __grid_constant__ StepKernelArgs a)
{
    __shared__ ObsShared obs_sh;
    __shared__ uint32_t obs_stream[OBS ? (BLOCK / 32) * ObsGeom<G>::WARP_WORDS : 1];
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
    __shared__ uint4 slab_rows[BLOCK];
    if (a.stage_slab) {   // kernel-uniform
        // Zero copy: the skyfall slab lives in pinned host memory.  The cascade reads its words lazily, which over
        // PCIe is one 4-byte round trip per lane and cascade iteration; fetch the board's whole 16-byte row once
        // instead (512 contiguous bytes per warp, in flight together with the state load) and let the cascade read
        // it from shared memory.
        if (slab) {
            slab_rows[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(slab));
            slab = reinterpret_cast<const uint32_t *>(&slab_rows[threadIdx.x]);
        }
        __syncwarp();
    }
    // every lane takes part (the cascade loop inside is warp-uniform); lanes past the batch are not live
    Board<G> fin;
    const StepOut o = step_board<G, AR, AR>(valid, s, fin, a.p, a.env0 + (uint64_t)i, action, slab,
                                    a.combo_log && valid ? a.combo_log + (size_t)i * a.p.log_cap : nullptr,
                                    a.paths && valid ? a.paths + (size_t)i * a.path_words : nullptr, path_len);
    if (valid) {
        StateIO<G>::store(a.state, i, s.b, s.meta);
        if (a.action_out) a.action_out[i] = (uint8_t)o.action;
        if (a.reward) a.reward[i] = o.reward;
        if (a.info) a.info[i] = o.info;
        if (a.final_state) StateIO<G>::store(a.final_state, i, fin, s.meta);
        if (a.state_copy) StateIO<G>::store(a.state_copy, i, s.b, s.meta);
    }
    const int64_t warp0 = i - (threadIdx.x & 31);
    if (OBS && warp0 < a.n) {
        const int n_valid = (int)(a.n - warp0 < 32 ? a.n - warp0 : 32);
        emit_obs<G>(s, valid, n_valid, obs_stream + (threadIdx.x >> 5) * ObsGeom<G>::WARP_WORDS, obs_sh,
                    obs_warp_ptr<G>(a.obs, warp0, a.obs_dtype), a.obs_dtype);
    }
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
constexpr int POOL_SLOTS = PBLOCK + PBLOCK / 2;
constexpr int POOL_FILL_BELOW = PBLOCK / 2;    // fill while pool <= this: pool + PBLOCK <= POOL_SLOTS always holds

template <class G>
struct CascadePool {
    typename G::W b0[POOL_SLOTS], b1[POOL_SLOTS], b2[POOL_SLOTS], b3[G::NP > 3 ? POOL_SLOTS : 1];
    double t0[POOL_SLOTS], t1[POOL_SLOTS], t2[POOL_SLOTS], t3[POOL_SLOTS], t4[POOL_SLOTS];
    uint32_t counts[POOL_SLOTS];   // n_combos [0,16) | depth [16,24) | flags [24,32)
    uint32_t source[POOL_SLOTS];   // skyfall cursor [0,16) | injected slab exhausted [16]
    uint64_t owner[POOL_SLOTS];    // board index within the launch
    int warp_count[2][PBLOCK / 32];   // double buffered: consecutive compactions alternate, one barrier each

    __device__ __forceinline__ void put(int slot, const CascadeItem<G> &it, uint64_t own)
    {
        b0[slot] = it.b.b0;
        b1[slot] = it.b.b1;
        b2[slot] = it.b.b2;
        PPAD_P4(b3[slot] = it.b.b3)
        t0[slot] = it.tally.t0;
        t1[slot] = it.tally.t1;
        t2[slot] = it.tally.t2;
        t3[slot] = it.tally.t3;
        t4[slot] = it.tally.t4;
        const uint32_t nc = it.tally.n_combos > 0xFFFF ? 0xFFFFu : (uint32_t)it.tally.n_combos;
        counts[slot] = nc | (sat255(it.depth) << 16) | ((it.flags & 0xFFu) << 24);
        source[slot] = (it.cursor > 0xFFFF ? 0xFFFFu : (uint32_t)it.cursor) | ((it.exhausted ? 1u : 0u) << 16);
        owner[slot] = own;
    }
    __device__ __forceinline__ uint64_t get(int slot, CascadeItem<G> &it) const
    {
        it.b.b0 = b0[slot];
        it.b.b1 = b1[slot];
        it.b.b2 = b2[slot];
        PPAD_P4(it.b.b3 = b3[slot])
        it.tally.t0 = t0[slot];
        it.tally.t1 = t1[slot];
        it.tally.t2 = t2[slot];
        it.tally.t3 = t3[slot];
        it.tally.t4 = t4[slot];
        const uint32_t c = counts[slot], q = source[slot];
        it.tally.n_combos = (int)(c & 0xFFFFu);
        it.depth = (int)((c >> 16) & 0xFFu);
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
            if (valid && pre.evaluate) {
                W er, ed;
                cascading = match_mask<G>(it.b, er, ed) != 0;
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
                }
            }
            int added;
            const int slot = pool.compact(cascading, added, phase);
            phase ^= 1;
            if (cascading) pool.put(pool_count + slot, it, (uint64_t)i);
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
        if (tid < m) {
            own = pool.get(base + tid, it);
            const RngKey key = board_key(a.p, a.env0 + own);
            again = cascade_round<G>(it, sd, a.p.cascade_cap, a.p.team,
                                     a.slab ? a.slab + (size_t)own * a.p.slab_words : nullptr, a.p.n_inj, key, a.p.sky,
                                     a.combo_log ? a.combo_log + (size_t)own * a.p.log_cap : nullptr, a.p.log_cap);
            if (!again) {
                if (a.reward) a.reward[own] = it.reward();
                if (a.info) a.info[own] = it.info();
                if (a.final_state) StateIO<G>::store_planes(a.final_state, own, it.b);   // meta: written by the fill
            }
        }

        if (m > 0) {                             // CTA-uniform
            int survivors;
            const int slot = pool.compact(again, survivors, phase);   // its barrier: every get() of the round is done
            phase ^= 1;
            if (again) pool.put(base + slot, it, own);   // read again only after the barrier at the top of the round
            pool_count = base + survivors;
        }
    }
}

template <class G>
__global__ void __launch_bounds__(BLOCK) encode_kernel(int64_t n, const void *state, void *obs, int obs_dtype)
{
    __shared__ ObsShared obs_sh;
    __shared__ uint32_t obs_stream[(BLOCK / 32) * ObsGeom<G>::WARP_WORDS];
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
        emit_obs<G>(s, valid, n_valid, obs_stream + (threadIdx.x >> 5) * ObsGeom<G>::WARP_WORDS, obs_sh,
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
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    const bool valid = i < a.n;                  // every lane takes part: the rejection sampling is warp-cooperative
    int finger = -1;
    if (valid && a.fingers_in) finger = a.fingers_in[2 * i] * G::C + a.fingers_in[2 * i + 1];
    State<G> s;
    clear_state<G>(s);
    int consumed = 0;
    const uint32_t flags = reset_boards_warp<G, true>(valid, s, a.p, a.p.key.step, a.env0 + (uint64_t)i,
                                                a.slab && valid ? a.slab + (size_t)i * a.p.slab_words : nullptr,
                                                a.p.n_inj, finger, &consumed);
    if (!valid) return;
    StateIO<G>::store(a.state, i, s.b, s.meta);
    if (a.flags) a.flags[i] = (uint8_t)flags;
    if (a.consumed) a.consumed[i] = (uint8_t)sat255(consumed);
}

// plain int boards -> state.  Like unpack below, through shared memory: the CTA reads its contiguous input region with
// consecutive threads on consecutive elements and parks one code byte per cell (0x80 = a value that is no orb type);
// then every thread packs its own board from there.
template <class G, class T>
__global__ void __launch_bounds__(BLOCK) pack_kernel(int64_t n, const T *boards, const T *fingers, const uint8_t *prev,
                                                     const uint16_t *moves, void *state, uint8_t *flags)
{
    __shared__ uint8_t codes[BLOCK * G::CELLS];
    const int64_t tile0 = (int64_t)blockIdx.x * BLOCK;
    const int64_t left = n - tile0;
    const int count = (int)(left < BLOCK ? left : BLOCK) * G::CELLS;
    const T *in = boards + (size_t)tile0 * G::CELLS;
    for (int e = threadIdx.x; e < count; e += BLOCK) {
        const long long v = (long long)in[e];
        codes[e] = v == -1 ? (uint8_t)G::EMPTY : ((v >= 0 && v < G::NTYPES) ? (uint8_t)v : (uint8_t)0x80);
    }
    __syncthreads();
    const int64_t i = tile0 + threadIdx.x;
    if (i >= n) return;
    int8_t cells[G::CELLS];
    bool bad = false;
#pragma unroll
    for (int c = 0; c < G::CELLS; c++) {
        const uint8_t v = codes[threadIdx.x * G::CELLS + c];
        bad |= v == 0x80;
        cells[c] = v == 0x80 || v == G::EMPTY ? (int8_t)-1 : (int8_t)v;
    }
    Board<G> b;
    pack_cells<G, int8_t>(b, cells);
    const int finger = (int)fingers[2 * i] * G::C + (int)fingers[2 * i + 1];
    StateIO<G>::store(state, i, b, meta_pack(finger, prev ? prev[i] : ACT_PASS, moves ? moves[i] : 0));
    if (flags) flags[i] = (uint8_t)(bad ? FLAG_UNSUPPORTED_ORB : 0);
}

// state -> plain int boards.  A board is CELLS elements of T (240 B as int64), so one thread per board would write
// 32 scattered runs per warp instruction; instead the CTA parks its 256 boards' planes in shared memory and then
// writes its contiguous output region element by element, consecutive threads -> consecutive addresses.
template <class G, class T>
__global__ void __launch_bounds__(BLOCK) unpack_kernel(int64_t n, const void *state, T *boards, T *fingers, uint8_t *prev,
                                                       uint16_t *moves)
{
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
            fingers[2 * i] = (T)(f / G::C);
            fingers[2 * i + 1] = (T)(f % G::C);
        }
        if (prev) prev[i] = (uint8_t)meta_prev(s.meta);
        if (moves) moves[i] = (uint16_t)meta_moves(s.meta);
    }
    if (!boards) return;
    __syncthreads();
    const int64_t left = n - tile0;
    const int count = (int)(left < BLOCK ? left : BLOCK) * G::CELLS;
    T *out = boards + (size_t)tile0 * G::CELLS;
    for (int e = threadIdx.x; e < count; e += BLOCK) {
        const int b = e / G::CELLS, c = e - b * G::CELLS;
        int code = 0;
#pragma unroll
        for (int j = 0; j < G::NP; j++) code |= (int)((planes[j][b] >> c) & 1) << j;
        out[e] = (T)(code == G::EMPTY ? -1 : code);
    }
}

template <class G>
__global__ void __launch_bounds__(BLOCK) cancel_kernel(int64_t n, void *state, uint8_t *n_combos, uint16_t *combo_log,
                                                       int log_cap)
{
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

template <class G>
__global__ void __launch_bounds__(BLOCK) legal_kernel(int64_t n, const void *state, uint8_t *mask)
{
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    const State<G> s = StateIO<G>::load(state, i);
    const int f = meta_finger(s.meta);
    const int off = (~legal_mask<G>(f, ACT_PASS)) & 15;   // off-board moves only (utils.py:73-83)
    mask[i] = (uint8_t)(legal_mask<G>(f, meta_prev(s.meta)) | (off << 4));
}

// PlayerAction.sample (player.py:38-66): same draw as the in-kernel policy of step_kernel
template <class G>
__global__ void __launch_bounds__(BLOCK) sample_kernel(int64_t n, uint64_t env0, RngKey key, const void *state,
                                                       int include_pass, uint8_t *actions)
{
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
__global__ void __launch_bounds__(BLOCK) drop_kernel(int64_t n, void *state)
{
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
__global__ void __launch_bounds__(BLOCK) permute_kernel(int64_t n, void *state, const uint8_t *maps)
{
    typedef typename G::W W;
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    State<G> s = StateIO<G>::load(state, i);
    const uint2 m2 = *reinterpret_cast<const uint2 *>(maps + 8 * i);
    const uint64_t lut = (uint64_t)m2.x | ((uint64_t)m2.y << 32);
    W p0 = 0, p1 = 0, p2 = 0, p3 = 0;
#pragma unroll
    for (int c = 0; c < G::CELLS; c++) {
        const int code = code_at<G>(s.b, c);
        const int to = code > 6 ? code : (int)((lut >> (8 * code)) & 7);   // empty cells and types 7..9 stay
        p0 |= (W)(to & 1) << c;
        p1 |= (W)((to >> 1) & 1) << c;
        p2 |= (W)((to >> 2) & 1) << c;
        PPAD_P4(p3 |= (W)((to >> 3) & 1) << c)
    }
    s.b.b0 = p0;
    s.b.b1 = p1;
    s.b.b2 = p2;
    PPAD_P4(s.b.b3 = p3)
    StateIO<G>::store(state, i, s.b, s.meta);
}

// Agent01.act after the network (agent01.py:173-218): Q-values -> one action per board
template <class G>
__global__ void __launch_bounds__(BLOCK) select_kernel(int64_t n, uint64_t env0, RngKey key, const void *state,
                                                       const float *q, int method, float beta, float epsilon,
                                                       int max_moves, uint8_t *actions)
{
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
                                                            const uint32_t *state, const uint8_t *action,
                                                            const double *reward, const uint8_t *keep,
                                                            const int64_t *slot, uint32_t *r_state, uint8_t *r_action,
                                                            double *r_reward)
{
    const int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    if (keep && !keep[i]) return;
    const int64_t d = (cursor + (slot ? slot[i] : i)) % capacity;
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

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

int fail(int status, const std::string &msg)
{
    g_last_error = msg;
    return status;
}

int cuda_check(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return PPAD_OK;
    return fail(PPAD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

int after_launch(const char *name)
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

}   // namespace

struct ppad_ctx {
    ppad_config cfg;
    int device;
    int geom;   // 0 = 5x6, 1 = 6x7, 2 = 4x5; 3..5 = the same with 4-bit orb codes
    StepParams base;
    int sm_count;
    int pool_ctas_per_sm;   // resident CTAs of the persistent step_pool_kernel
};

#define PPAD_DISPATCH(ctx, EXPR)                             \
    do {                                                     \
        if ((ctx)->geom == 0) { typedef G56 G; EXPR; }       \
        else if ((ctx)->geom == 1) { typedef G67 G; EXPR; }  \
        else if ((ctx)->geom == 2) { typedef G45 G; EXPR; }  \
        else if ((ctx)->geom == 3) { typedef G56W G; EXPR; } \
        else if ((ctx)->geom == 4) { typedef G67W G; EXPR; } \
        else { typedef G45W G; EXPR; }                       \
    } while (0)

template <class G, class T>
static void launch_pack(int64_t n, const void *boards, const void *fingers, const uint8_t *prev, const uint16_t *moves,
                        void *state, uint8_t *flags, cudaStream_t st)
{
    pack_kernel<G, T><<<grid_for(n), BLOCK, 0, st>>>(n, (const T *)boards, (const T *)fingers, prev, moves, state, flags);
}

template <class G, class T>
static void launch_unpack(int64_t n, const void *state, void *boards, void *fingers, uint8_t *prev, uint16_t *moves,
                          cudaStream_t st)
{
    unpack_kernel<G, T><<<grid_for(n), BLOCK, 0, st>>>(n, state, (T *)boards, (T *)fingers, prev, moves);
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
    cudaSetDevice(device);
    ctx->pool_ctas_per_sm = 0;
    PPAD_DISPATCH(ctx, (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->pool_ctas_per_sm, step_pool_kernel<G, true>, PBLOCK, 0)));
    if (ctx->pool_ctas_per_sm <= 0) ctx->pool_ctas_per_sm = 1;
    *out = ctx;
    return PPAD_OK;
}

void ppad_destroy(ppad_ctx *ctx) { delete ctx; }

static int enter(const ppad_ctx *ctx, int64_t n)
{
    if (!ctx) return fail(PPAD_ERR_INVALID, "null context");
    if (n < 0 || n > ((int64_t)1 << 40)) return fail(PPAD_ERR_INVALID, "bad batch size");
    return cuda_check(cudaSetDevice(ctx->device), "cudaSetDevice");
}

int ppad_pack(const ppad_ctx *ctx, int64_t n, const void *boards, const void *fingers, int dtype, const uint8_t *prev,
              const uint16_t *moves, void *state, uint8_t *flags, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!boards || !fingers || !state) return fail(PPAD_ERR_INVALID, "ppad_pack: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case PPAD_I8: PPAD_DISPATCH(ctx, (launch_pack<G, int8_t>(n, boards, fingers, prev, moves, state, flags, st))); break;
    case PPAD_I32: PPAD_DISPATCH(ctx, (launch_pack<G, int32_t>(n, boards, fingers, prev, moves, state, flags, st))); break;
    case PPAD_I64: PPAD_DISPATCH(ctx, (launch_pack<G, int64_t>(n, boards, fingers, prev, moves, state, flags, st))); break;
    default: return fail(PPAD_ERR_UNSUPPORTED, "ppad_pack: dtype must be PPAD_I8, PPAD_I32 or PPAD_I64");
    }
    return after_launch("pack_kernel");
}

int ppad_unpack(const ppad_ctx *ctx, int64_t n, const void *state, void *boards, void *fingers, int dtype, uint8_t *prev,
                uint16_t *moves, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!state) return fail(PPAD_ERR_INVALID, "ppad_unpack: null state");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case PPAD_I8: PPAD_DISPATCH(ctx, (launch_unpack<G, int8_t>(n, state, boards, fingers, prev, moves, st))); break;
    case PPAD_I32: PPAD_DISPATCH(ctx, (launch_unpack<G, int32_t>(n, state, boards, fingers, prev, moves, st))); break;
    case PPAD_I64: PPAD_DISPATCH(ctx, (launch_unpack<G, int64_t>(n, state, boards, fingers, prev, moves, st))); break;
    default: return fail(PPAD_ERR_UNSUPPORTED, "ppad_unpack: dtype must be PPAD_I8, PPAD_I32 or PPAD_I64");
    }
    return after_launch("unpack_kernel");
}

int ppad_reset(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, const uint32_t *slab, int32_t slab_words,
               int32_t n_inj, const int32_t *fingers_in, void *state, uint8_t *flags, uint8_t *consumed, void *stream)
{
    if (int r = enter(ctx, n)) return r;
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

static int launch_step(const ppad_ctx *ctx, const ppad_step_args *args, cudaStream_t st, bool coalesced_results = false,
                       bool stage_host_slab = false)
{
    if (!args->state) return fail(PPAD_ERR_INVALID, "ppad_step: null state");
    if (args->slab && (args->slab_words <= 0 || args->n_inj < 0 || args->n_inj > args->slab_words * 8))
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
    a.obs = args->obs;
    a.obs_dtype = args->obs_dtype;
    a.paths = args->paths;
    a.path_lens = args->path_lens;
    a.path_words = args->path_words;
    a.path_len = args->path_len;
    if (a.paths && (a.path_words <= 0 || a.path_len < 0))
        return fail(PPAD_ERR_INVALID, "ppad_step: paths need path_words > 0 and path_len >= 0");
    a.stage_slab = (stage_host_slab && a.slab && a.p.slab_words == 4 && ((uintptr_t)a.slab & 15) == 0) ? 1 : 0;
    const bool ar = a.p.auto_reset != 0;
    if (args->obs) {
        if (ar) PPAD_DISPATCH(ctx, (step_obs_kernel<G, true, true><<<grid_for(a.n), BLOCK, 0, st>>>(a)));
        else PPAD_DISPATCH(ctx, (step_obs_kernel<G, true, false><<<grid_for(a.n), BLOCK, 0, st>>>(a)));
        return after_launch("step_obs_kernel");
    }
    if (coalesced_results || (!a.p.eval_moves && !a.actions && !a.paths) || !a.p.skyfall_damage) {
        // results go straight to host memory: the pooled kernel's scattered 8-byte result writes would become tiny
        // PCIe transactions, one thread per board writes them as full lines.
        // Steps of the in-kernel random policy that only evaluate passes (eval_moves = 0, no actions given: plain
        // moves until the episode length cap) take this kernel, too: a warp whose boards all just move leaves early,
        // and there is nothing for the pool to pack (episode mode with one launch per step: 4.4e10 against 3.7e10
        // env.step calls/s).  Given actions may all be passes (calculate_reward of a whole slab): those stay pooled.
        // Without skyfall (skyfall_damage = 0: the look-ahead's and smart data's evaluations) a board is cancelled
        // once and never cascades, so there is nothing to pool either: 0.027 against 0.033 ms per 2^20 evaluations.
        if (ar) PPAD_DISPATCH(ctx, (step_obs_kernel<G, false, true><<<grid_for(a.n), BLOCK, 0, st>>>(a)));
        else PPAD_DISPATCH(ctx, (step_obs_kernel<G, false, false><<<grid_for(a.n), BLOCK, 0, st>>>(a)));
        return after_launch("step_obs_kernel<no obs>");
    }
    // persistent: one wave of CTAs, as many as fit
    const int64_t tiles = (a.n + PBLOCK - 1) / PBLOCK;
    int64_t grid = (int64_t)ctx->sm_count * ctx->pool_ctas_per_sm;
    if (grid > tiles) grid = tiles;
    if (ar) PPAD_DISPATCH(ctx, (step_pool_kernel<G, true><<<(unsigned)grid, PBLOCK, 0, st>>>(a)));
    else PPAD_DISPATCH(ctx, (step_pool_kernel<G, false><<<(unsigned)grid, PBLOCK, 0, st>>>(a)));
    return after_launch("step_pool_kernel");
}

int ppad_step(const ppad_ctx *ctx, const ppad_step_args *args, void *stream)
{
    if (!args) return fail(PPAD_ERR_INVALID, "ppad_step: null args");
    if (int r = enter(ctx, args->n)) return r;
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
    if (int r = enter(ctx, n)) return r;
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
    if (traj_state || traj_action) PPAD_DISPATCH(ctx, (rollout_kernel<G, true><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(a)));
    else PPAD_DISPATCH(ctx, (rollout_kernel<G, false><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(a)));
    return after_launch("rollout_kernel");
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
};

// NULL counts as pinned (absent buffer); anything cudaHostAlloc'ed / cudaHostRegister'ed is device-visible under UVA
static bool host_is_pinned(const void *ptr)
{
    if (!ptr) return true;
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
    if (int r = enter(ctx, n_max)) return r;
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
    if (prev_device >= 0) cudaSetDevice(prev_device);   // destroy may run from a finaliser: leave the caller's device alone
    delete p;
}

int ppad_pipe_step(ppad_pipe *p, const ppad_host_step_args *h, void *after_stream, int32_t slot)
{
    if (!p || !h) return fail(PPAD_ERR_INVALID, "ppad_pipe_step: null argument");
    if (slot < 0 || slot > 1) return fail(PPAD_ERR_INVALID, "ppad_pipe_step: slot must be 0 or 1");
    const ppad_ctx *ctx = p->ctx;
    if (int r = enter(ctx, h->n)) return r;
    if (h->n > p->n_max) return fail(PPAD_ERR_INVALID, "ppad_pipe_step: n exceeds the pipe's n_max");
    if (h->slab_host && (h->slab_words <= 0 || h->slab_words > p->slab_words_max))
        return fail(PPAD_ERR_INVALID, "ppad_pipe_step: slab_words exceeds the pipe's slab_words_max");
    if (!h->state) return fail(PPAD_ERR_INVALID, "ppad_pipe_step: null state");
    if (h->n == 0) return PPAD_OK;
    const size_t rec = (size_t)ppad_state_bytes_ex(ctx->cfg.rows, ctx->cfg.cols, ctx->cfg.orb_bits);
    const size_t obs_elem = h->obs_dtype == PPAD_OBS_F32 ? 4 : 1;
    const size_t obs_row = (size_t)ctx->cfg.rows * ctx->cfg.cols * 7 * obs_elem;
    // chunk streams start after whatever the caller has queued on after_stream (e.g. the reset)
    cudaError_t e = cudaEventRecord(p->fork, (cudaStream_t)after_stream);
    const bool zc = p->zero_copy && host_is_pinned(h->actions_host) && host_is_pinned(h->slab_host) &&
                    host_is_pinned(h->action_out_host) && host_is_pinned(h->reward_host) && host_is_pinned(h->info_host) &&
                    host_is_pinned(h->state_out_host);
    // zero_copy 1: the kernel reads the slab from and writes its results to the host buffers; 2 ("hybrid"): the slab
    // (and given actions) travel by DMA, chunk by chunk on the pipe's streams, and only the results are written in place
    // by the kernels -- no PCIe read sits inside a kernel, which matters on hosts whose read latency is high
    const bool hybrid = zc && p->zero_copy == 2;
    const int mode = zc && !hybrid ? 1 : 2;
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
        a.action_out = h->action_out_host;
        a.reward = h->reward_host;
        a.info = h->info_host;
        a.state_copy = h->state_out_host;
        a.obs = h->obs;
        a.obs_dtype = h->obs_dtype;
        if (int r = launch_step(ctx, &a, st, true, true)) return r;
        for (int k = 0; k < p->chunks && e == cudaSuccess; k++) e = cudaEventRecord(p->done[slot][k], st);
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
    return cuda_check(e, "ppad_pipe_step");
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

int ppad_encode_obs(const ppad_ctx *ctx, int64_t n, const void *state, void *obs, int obs_dtype, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!state || !obs) return fail(PPAD_ERR_INVALID, "ppad_encode_obs: null pointer");
    if (obs_dtype != PPAD_OBS_F32 && obs_dtype != PPAD_OBS_U8)
        return fail(PPAD_ERR_UNSUPPORTED, "ppad_encode_obs: obs_dtype must be PPAD_OBS_F32 or PPAD_OBS_U8");
    if ((uintptr_t)obs & 15) return fail(PPAD_ERR_INVALID, "ppad_encode_obs: obs must be 16-byte aligned");
    PPAD_DISPATCH(ctx, (encode_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, state, obs, obs_dtype)));
    return after_launch("encode_kernel");
}

int ppad_cancel(const ppad_ctx *ctx, int64_t n, void *state, uint8_t *n_combos, uint16_t *combo_log, int32_t log_cap,
                void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!state) return fail(PPAD_ERR_INVALID, "ppad_cancel: null state");
    if (combo_log && log_cap <= 0) return fail(PPAD_ERR_INVALID, "ppad_cancel: combo_log needs log_cap > 0");
    PPAD_DISPATCH(ctx, (cancel_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, state, n_combos, combo_log,
                                                                                         log_cap)));
    return after_launch("cancel_kernel");
}

int ppad_legal_mask(const ppad_ctx *ctx, int64_t n, const void *state, uint8_t *mask, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!state || !mask) return fail(PPAD_ERR_INVALID, "ppad_legal_mask: null pointer");
    PPAD_DISPATCH(ctx, (legal_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, state, mask)));
    return after_launch("legal_kernel");
}

int ppad_sample_actions(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, const void *state,
                        int32_t include_pass, uint8_t *actions, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!state || !actions) return fail(PPAD_ERR_INVALID, "ppad_sample_actions: null pointer");
    RngKey key = ctx->base.key;
    key.step = step;
    PPAD_DISPATCH(ctx, (sample_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, env0, key, state,
                                                                                         include_pass, actions)));
    return after_launch("sample_kernel");
}

int ppad_drop(const ppad_ctx *ctx, int64_t n, void *state, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!state) return fail(PPAD_ERR_INVALID, "ppad_drop: null state");
    PPAD_DISPATCH(ctx, (drop_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, state)));
    return after_launch("drop_kernel");
}

int ppad_fill(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, int32_t draw0, int32_t reset_all,
              const uint32_t *slab, int32_t slab_words, int32_t n_inj, void *state, uint8_t *flags, uint8_t *consumed,
              void *stream)
{
    if (int r = enter(ctx, n)) return r;
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
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!combo_log || !n_combos || !reward || log_cap <= 0) return fail(PPAD_ERR_INVALID, "ppad_damage: bad argument");
    damage_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, ctx->base.team, combo_log, log_cap, n_combos, reward);
    return after_launch("damage_kernel");
}

int ppad_permute(const ppad_ctx *ctx, int64_t n, void *state, const uint8_t *maps, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!state || !maps) return fail(PPAD_ERR_INVALID, "ppad_permute: null pointer");
    if ((uintptr_t)maps & 7) return fail(PPAD_ERR_INVALID, "ppad_permute: maps must be 8-byte aligned");
    PPAD_DISPATCH(ctx, (permute_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, state, maps)));
    return after_launch("permute_kernel");
}

int ppad_select_actions(const ppad_ctx *ctx, int64_t n, uint64_t env0, uint32_t step, const void *state, const float *q,
                        int32_t method, float beta, float epsilon, int32_t max_moves, uint8_t *actions, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!state || !q || !actions) return fail(PPAD_ERR_INVALID, "ppad_select_actions: null pointer");
    if (method != SELECT_MAX && method != SELECT_BOLTZMANN)
        return fail(PPAD_ERR_INVALID, "ERROR: Unknown action method!");
    RngKey key = ctx->base.key;
    key.step = step;
    PPAD_DISPATCH(ctx, (select_kernel<G><<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(
                           n, env0, key, state, q, method, beta, epsilon, max_moves, actions)));
    return after_launch("select_kernel");
}

int ppad_discount(const ppad_ctx *ctx, int64_t n, int32_t max_len, double *rewards, const int32_t *lengths, double gamma,
                  int32_t log10_flag, int32_t norm, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0 || max_len == 0) return PPAD_OK;
    if (!rewards || max_len < 0) return fail(PPAD_ERR_INVALID, "ppad_discount: bad argument");
    discount_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, max_len, rewards, lengths, gamma, log10_flag, norm);
    return after_launch("discount_kernel");
}

int ppad_replay_push(const ppad_ctx *ctx, int64_t n, int64_t capacity, int64_t cursor, const void *state,
                     const uint8_t *action, const double *reward, const uint8_t *keep, const int64_t *slot,
                     void *r_state, uint8_t *r_action, double *r_reward, void *stream)
{
    if (int r = enter(ctx, n)) return r;
    if (n == 0) return PPAD_OK;
    if (!state || !action || !reward || !r_state || !r_action || !r_reward || capacity <= 0 || cursor < 0)
        return fail(PPAD_ERR_INVALID, "ppad_replay_push: bad argument");
    const int words = (int)(ppad_state_bytes_ex(ctx->cfg.rows, ctx->cfg.cols, ctx->cfg.orb_bits) / 4);
    replay_push_kernel<<<grid_for(n), BLOCK, 0, (cudaStream_t)stream>>>(n, words, capacity, cursor, (const uint32_t *)state,
                                                                        action, reward, keep, slot, (uint32_t *)r_state,
                                                                        r_action, r_reward);
    return after_launch("replay_push_kernel");
}

int ppad_replay_sample(const ppad_ctx *ctx, int64_t k, int64_t size, uint32_t step, const void *r_state,
                       const uint8_t *r_action, const double *r_reward, void *state, uint8_t *action, double *reward,
                       int64_t *index, void *stream)
{
    if (int r = enter(ctx, k)) return r;
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
