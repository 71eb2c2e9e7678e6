#!/bin/bash
# Round measurement pass on one B200 (run under gpurun): bench + per-shape GEMM table + step profile + deformable-path timings
# + ncu launch list of the bench command.  Outputs under gpurun_out/.
set -u
TAG=${1:-r01}
python bench.py --steps 10 --warmup 3 --prof-out gpurun_out/${TAG}_gemm_shapes_b8.txt > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
python tools/step_profile.py 8 bf16 > gpurun_out/${TAG}_step_profile_b8.txt 2>&1
python tools/bench_deform.py > gpurun_out/${TAG}_deform.json 2>&1
tail -30 gpurun_out/${TAG}_deform.json
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_b8.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_run.log 2>&1
python tools/agg_launches.py gpurun_out/${TAG}_launches_b8.csv 25
