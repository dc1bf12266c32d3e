#!/bin/bash
# Timing experiment (NOT a product path): the same kernel sources built with -DLS_ABLATE, where LS2FM_ABLATE bit 0 drops the table-gradient
# scatter, bit 1 replaces the hash-table loads by constants, bit 2 drops the weight-gradient MMAs.  Results are WRONG by construction; only
# the kernel durations mean something: they show how much of each kernel is memory work that is not hidden.   tools/ablate.sh (on the GPU box)
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DLS_ABLATE \
     -o gpurun_out/libls2fm_ablate.so level-s2fm_official_b200/csrc/ls2fm_api.cu 2> /dev/null
for m in 0 1 2 4 3 5 7; do
  LS2FM_LIB=$PWD/gpurun_out/libls2fm_ablate.so LS2FM_ABLATE=$m python bench.py --no-cpu-baseline --no-gpu-eager --steps 10 2> /dev/null | python -c "
import json,sys
d=json.load(sys.stdin); k=d['roofline']['kernel_ms_per_step']
print('ablate=$m  step %.3f ms | field_backward %.3f  field_forward %.3f  sampler %.3f' % (d['ms_per_step'], k['field_backward'], k['field_forward'], k['sample_error_bounded']))"
done
rm -f gpurun_out/libls2fm_ablate.so
