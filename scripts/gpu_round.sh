#!/bin/bash
# one GPU call: parity tests, bench (C2, with the CPU baseline), ncu launch list of one eager step, C3 / C4 bench lines.
# Usage: scripts/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cut -c1-260 gpurun_out/bench_$tag.json
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_$tag.json 2>/dev/null
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_$tag.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref_$tag.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 1500 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
python scripts/timeline.py > gpurun_out/timeline_$tag.log 2>&1; rm -f gpurun_out/timeline.json; mv gpurun_out/timeline.txt gpurun_out/timeline_$tag.txt
tail -9 gpurun_out/timeline_$tag.log
