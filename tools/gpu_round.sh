#!/bin/bash
# One GPU-box visit: parity suite, bench, launch list, full ncu captures of the two kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 600 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render -s 5 -c 1 -o gpurun_out/render python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 5 -c 1 -o gpurun_out/step python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > /dev/null 2>&1
python tools/throughput_all.py > gpurun_out/throughput.jsonl 2>> gpurun_out/bench.err; cat gpurun_out/throughput.jsonl | cut -c1-300
