// ppad_agent.cuh -- the steps right before and after the env step in a rollout:
// action selection from Q-values with the reference's masking rules (ppad/agent/agent01.py:173-218),
// the discounted-return scan (ppad/agent/utils.py:22-51) and the device-side experience replay ring
// (projects/simulation.py:294-306).  Device code, included by ppad_kernels.cu.
#pragma once
#include "ppad_core.cuh"

namespace ppad {

enum : int { SELECT_MAX = 0, SELECT_BOLTZMANN = 1 };

// agent01.py:173-190: -inf on off-board moves (with the reference's elif pairing) and on the reverse of
// the last action; 'pass' (4) is never masked.  Returns the bit mask of allowed actions.
template <class G>
PPAD_HD int agent_allowed_mask(int finger, int last_action)
{
    const int r = finger / G::cols(), c = finger - r * G::cols();
    int m = 31;
    if (r == 0) m &= ~1;
    else if (r == G::rows() - 1) m &= ~2;
    if (c == 0) m &= ~4;
    else if (c == G::cols() - 1) m &= ~8;
    if (last_action < 4) m &= ~(1 << (last_action ^ 1));
    return m;
}

// One action per board from q[5] (float32, as Agent01.predict returns it).
//   epsilon > 0 and u0 < epsilon : uniform over the allowed actions (agent01.py:193-201)
//   SELECT_MAX                   : first maximum (np.argmax, agent01.py:205-206)
//   SELECT_BOLTZMANN             : p ~ exp(beta * q), inverse-CDF draw like np.random.choice (agent01.py:207-213);
//                                  the maximum is subtracted before exp (same probabilities, no overflow)
// u0, u1 are draws 0 and 1 of the Philox POLICY stream, mapped to [0, 1) as x / 2^32.
template <class G>
__device__ __forceinline__ int select_action(const float *q, int finger, int last_action, int method, float beta,
                                             float epsilon, uint32_t x0, uint32_t x1)
{
    const int allowed = agent_allowed_mask<G>(finger, last_action);
    if (epsilon > 0.f && (double)epsilon > (double)x0 * (1.0 / 4294967296.0)) {
        return sample_from_mask(allowed, x1);
    }
    float qm[5];
    float best = -INFINITY;
    int arg = 4;
#pragma unroll
    for (int a = 4; a >= 0; a--) {
        qm[a] = ((allowed >> a) & 1) ? q[a] : -INFINITY;
        if (qm[a] >= best) {   // descending scan with >= keeps the FIRST maximum
            best = qm[a];
            arg = a;
        }
    }
    if (method == SELECT_MAX) return arg;
    float p[5];
    float sum = 0.f;
#pragma unroll
    for (int a = 0; a < 5; a++) {
        p[a] = ((allowed >> a) & 1) ? expf(beta * (qm[a] - best)) : 0.f;
        sum += p[a];
    }
    // np.random.choice: first a with cdf[a] > u, cdf = cumsum(p / sum).  Compared as cumsum(p) > u * sum in float64
    // (the same predicate up to the last rounding of the division): no float64 divisions on the hot path.
    const double us = (double)x1 * (1.0 / 4294967296.0) * (double)sum;
    double cum = 0.0;
    int pick = arg;
    bool found = false;
#pragma unroll
    for (int a = 0; a < 5; a++) {
        cum += (double)p[a];
        if (!found && ((allowed >> a) & 1) && cum > us) {
            pick = a;
            found = true;
        }
    }
    return pick;
}

// discount (agent/utils.py:22-51) of one trajectory r[0..len): optional log10 of the positive rewards,
// reverse scan cur = cur * gamma + r, optional normalisation.  All-non-positive trajectories are
// returned unchanged (utils.py:34).
__device__ __forceinline__ void discount_trajectory(double *r, int len, double gamma, int log10_flag, int norm)
{
    bool any_pos = false;
    for (int i = 0; i < len; i++) any_pos |= r[i] > 0.0;
    if (!any_pos) return;
    double cur = 0.0;
    for (int i = len - 1; i >= 0; i--) {
        double v = r[i];
        if (log10_flag && v > 0.0) v = log10(v);
        cur = dadd(dmul(cur, gamma), v);
        r[i] = cur;
    }
    if (norm) {
        double mean = 0.0;
        for (int i = 0; i < len; i++) mean += r[i];
        mean /= (double)len;
        double var = 0.0;
        for (int i = 0; i < len; i++) var += (r[i] - mean) * (r[i] - mean);
        const double sd = sqrt(var / (double)len);
        if (sd > 0.0)
            for (int i = 0; i < len; i++) r[i] = (r[i] - mean) / sd;
    }
}

}   // namespace ppad
