#!/bin/bash
# Round-end measurement pass on the GPU box: full GPU test suite, smoke, every bench workload with its baseline legs, reference arm,
# ncu launch list of the eager step.  Outputs under gpurun_out/final/.   tools/gpu.sh 1700 'bash tools/final_run.sh'
cd "$(dirname "$0")/.."
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > $O/gputest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python bench.py > $O/c2.json 2> $O/c2.err
python bench.py --regime trained > $O/c2_trained.json 2> /dev/null
python bench.py --graph off --no-cpu-baseline --no-gpu-eager > $O/c2_graph_off.json 2> /dev/null
for w in c1k uniform128 c3 c4 c5 dual; do python bench.py --workload $w > $O/$w.json 2> $O/$w.err; done
for w in image grid ba_sfm; do python bench.py --workload $w --no-cpu-baseline > $O/$w.json 2> $O/$w.err; done
python bench.py --impl reference --steps 5 --warmup 1 > $O/reference_arm.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --graph off --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-eager > $O/ncu_bench.log 2>&1
echo done > $O/done
