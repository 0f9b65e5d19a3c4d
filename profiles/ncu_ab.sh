#!/bin/bash
# ncu --set full of the four bench kernels for one variant library: profiles/ncu_ab.sh build/variants/NAME.so
lib=$1; name=$(basename $lib .so)
PPAD_AB_LIB=$lib python profiles/ncu_variants.py > /dev/null 2>&1 || exit 1
PPAD_AB_LIB=$lib ncu --set full --import-source on --clock-control none -k regex:"step_pool_kernel|step_obs_kernel|encode_kernel" -s 29 -c 4 \
  -f -o gpurun_out/prof_$name python profiles/ncu_variants.py > gpurun_out/ncu_$name.log 2>&1
