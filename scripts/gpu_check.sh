#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list, ncu full captures of the top kernels.
# usage: gpurun --timeout 1700 -- 'bash scripts/gpu_check.sh <tag>'
# The ncu passes run the engine eager and serial (--no-graphs --serial): one kernel per launch, nothing beside it.
TAG=${1:-dev}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
cat gpurun_out/pytest_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err; head -c 400 gpurun_out/bench_$TAG.json; echo
NCUB="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-graphs --serial"
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    $NCUB > gpurun_out/ncu_bench_stdout_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pw_gemm -s 60 -c 3 -f -o gpurun_out/prof_gemm_$TAG \
    $NCUB > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_tile -s 40 -c 2 -f -o gpurun_out/prof_dw_$TAG \
    $NCUB > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_kernel -c 1 -f -o gpurun_out/prof_ransac_$TAG \
    $NCUB > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr2_ -c 12 -f -o gpurun_out/prof_corresp_$TAG \
    $NCUB > /dev/null 2>&1
ls -la gpurun_out | tail -12
