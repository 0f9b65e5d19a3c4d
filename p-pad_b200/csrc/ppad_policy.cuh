// ppad_policy.cuh -- placeholder, filled in below
#pragma once
#include "ppad_step.cuh"
#define PPAD_POLICY_LAUNCHERS(PREFIX, G)
