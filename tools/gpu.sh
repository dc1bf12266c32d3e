#!/bin/bash
# build the sm_100a library here (nvcc cross-compiles), then run a command on the B200 box:  tools/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python level-s2fm_official_b200/build.py > /dev/null
exec timeout $(( $1 + 1900 )) /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
