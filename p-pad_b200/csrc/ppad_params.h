// ppad_params.h -- host-side translation of ppad_config (the PAD.__init__ keyword arguments,
// ppad/pad/game.py:39-68) into the tables the kernels read.  Host code only.
#pragma once
#include <cmath>
#include <cstring>
#include <string>

#include "../../include/ppad_b200.h"
#include "ppad_core.cuh"

namespace ppad {

// returns PPAD_OK or an error status with `err` set; geom: 0 = 5x6, 1 = 6x7, 2 = 4x5 with 3-bit orb codes,
// 3..5 = the same boards with 4-bit codes (cfg.orb_bits == 4: orb types 7..9 allowed); 6 / 7 = any other
// dim_row x dim_col with at most 64 cells (run-time dimensions, 3- / 4-bit codes)
inline int build_params(const ppad_config &cfg, StepParams &p, int &geom, std::string &err)
{
    std::memset(&p, 0, sizeof(p));
    // PAD.__init__ checks, game.py:54-68 (same messages)
    double sum = 0;
    for (int i = 0; i < 10; i++) {
        if (cfg.skyfall[i] < 0) {
            err = "ERROR: Skyfall rates should be a positive number!";
            return PPAD_ERR_INVALID;
        }
        sum += cfg.skyfall[i];
    }
    if (std::fabs(sum - 1) > 1e-4) {
        err = "ERROR: Skyfall rates should add up to 1!";
        return PPAD_ERR_INVALID;
    }
    if (cfg.cols < 0 || cfg.rows < 0) {
        err = "ERROR: Board dimensions should be positive integers!";
        return PPAD_ERR_INVALID;
    }
    if (cfg.cols * cfg.rows < 13) {
        err = "ERROR: The board should allow at least 13 orbs. Please increase board size!";
        return PPAD_ERR_INVALID;
    }
    if (cfg.rows < 1 || cfg.cols < 1 || cfg.cols * cfg.rows > 64) {
        err = "boards of more than 64 cells do not fit the packed state (one bit per cell and plane in a 64-bit word)";
        return PPAD_ERR_UNSUPPORTED;
    }
    if (cfg.orb_bits != 0 && cfg.orb_bits != 3 && cfg.orb_bits != 4) {
        err = "orb_bits must be 3 (orb types 0..6) or 4 (orb types 0..9)";
        return PPAD_ERR_INVALID;
    }
    const bool wide = cfg.orb_bits == 4;
    if (cfg.rows == 5 && cfg.cols == 6) geom = wide ? 3 : 0;
    else if (cfg.rows == 6 && cfg.cols == 7) geom = wide ? 4 : 1;
    else if (cfg.rows == 4 && cfg.cols == 5) geom = wide ? 5 : 2;
    else geom = wide ? 7 : 6;            // run-time dimensions: the generic (slower) instantiation
    if (!wide && (cfg.skyfall[7] > 0 || cfg.skyfall[8] > 0 || cfg.skyfall[9] > 0)) {
        err = "orb types 7..9 (poison, mortal poison, bomb) do not fit the 3-bit state: create the context with orb_bits = 4";
        return PPAD_ERR_UNSUPPORTED;
    }
    if (cfg.team_size < 0 || cfg.team_size > PPAD_MAX_TEAM_UNITS) {
        err = "team_size out of range";
        return PPAD_ERR_INVALID;
    }
    p.key.k0 = (uint32_t)cfg.seed;
    p.key.k1 = (uint32_t)(cfg.seed >> 32);
    // skyfall thresholds: first i with cumsum[i] >= x / 2^32 and prob[i] > 0 (game.py:464-470);
    // cumsum * 2^32 is exact in double, so x <= floor(cumsum * 2^32) is the same predicate
    double cum = 0;
    int last_valid = -1;
    for (int i = 0; i < 10; i++) {
        cum += cfg.skyfall[i];
        const double t = std::floor(cum * 4294967296.0);
        p.sky.thr[i] = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
        if (cfg.skyfall[i] > 0) {
            p.sky.valid |= 1u << i;
            last_valid = i;
        }
    }
    if (last_valid < 0 || p.sky.thr[last_valid] != 0xFFFFFFFFu) {
        err = "skyfall rates sum too far below 1: the reference's fall-through (game.py:470) would emit bombs";
        return PPAD_ERR_UNSUPPORTED;
    }
    // K equally likely leading types: orb_from_u32 may then replace the compares by one multiply.  Enabled only when
    // the thresholds the float rule produced ARE floor(k 2^32 / K) -- checked here, not assumed.
    {
        int k_valid = 0;
        while (k_valid < 10 && cfg.skyfall[k_valid] > 0) k_valid++;
        bool uniform = k_valid >= 2 && (int)(last_valid + 1) == k_valid;
        for (int i = 0; uniform && i < k_valid; i++) uniform = cfg.skyfall[i] == cfg.skyfall[0];
        for (int k = 1; uniform && k < k_valid; k++)
            uniform = p.sky.thr[k - 1] == (uint32_t)(((uint64_t)k << 32) / (uint64_t)k_valid);
        for (int i = k_valid; uniform && i < 10; i++) uniform = p.sky.thr[i] == 0xFFFFFFFFu;
        p.sky.uniform_k = uniform ? (uint32_t)k_valid : 0u;
    }
    // team table, game.py:171-180: per attack colour the (multiplier, atk) of matching units in team order
    for (int u = 0; u < cfg.team_size; u++) {
        const int a = cfg.team_color_1[u], b = cfg.team_color_2[u];
        if (a > 4 || b > 4 || a < -1 || b < -1) {
            err = "team colours must be red, blue, green, light, dark or None "
                  "(the reference's damage_tracker has no other keys, game.py:163-167)";
            return PPAD_ERR_UNSUPPORTED;
        }
        for (int colour = 0; colour < 5; colour++) {
            double m;
            if (a == b && b == colour) m = 1.1;
            else if (a == colour) m = 1;
            else if (b == colour) m = 0.3;
            else continue;
            const int k = p.team.n[colour]++;
            p.team.mult[colour][k] = m;
            p.team.atk[colour][k] = cfg.team_atk[u];
        }
    }
    if (cfg.cascade_cap < 0 || cfg.cascade_cap > 255) {
        err = "cascade_cap must be 0 (no cap) or 1..255: the depth of a cascade is reported, and parked, in 8 bits";
        return PPAD_ERR_INVALID;
    }
    if (cfg.reset_cap < 0) {
        err = "reset_cap must be >= 0";
        return PPAD_ERR_INVALID;
    }
    p.dyn = make_dyn_geom(cfg.rows, cfg.cols);
    p.skyfall_damage = cfg.skyfall_damage;
    p.cascade_cap = cfg.cascade_cap;
    p.reset_cap = cfg.reset_cap;
    return PPAD_OK;
}

}   // namespace ppad
