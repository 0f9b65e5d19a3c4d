// ppad_step_inst.cu -- the env-step kernels of ONE geometry (compile with -DPPAD_GEOM=0..5; ids as in GeomOf).
#include "ppad_step.cuh"
#include "ppad_policy.cuh"

#if !defined(PPAD_GEOM)
#error "compile with -DPPAD_GEOM=<geometry id 0..5>"
#endif

namespace ppad {
typedef GeomOf<PPAD_GEOM>::type GInst;
PPAD_STEP_LAUNCHERS(, GInst)
PPAD_POLICY_LAUNCHERS(, GInst)
}   // namespace ppad
