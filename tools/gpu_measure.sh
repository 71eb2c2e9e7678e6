#!/bin/bash
# Round measurement pass on one B200 (run under gpurun): benches of every workload + per-shape GEMM table + step profiles.
# Outputs under gpurun_out/ (copy what should be judged into profiles/).  `LAUNCHES=1` also captures the ncu launch list
# of the bench command (slow: ~10 min).
set -u
TAG=${1:-r01}
python bench.py --steps 10 --warmup 3 --prof-out gpurun_out/${TAG}_gemm_shapes_b8.txt > gpurun_out/${TAG}_bench_transmil_b8.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_transmil_b8.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null
for p in bf16 bf16x3; do
  python bench.py --workload deform2d --precision $p --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_deform2d_$p.json 2>/dev/null
  cut -c1-220 gpurun_out/${TAG}_bench_deform2d_$p.json
done
python bench.py --workload deform1d --precision bf16 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_deform1d_bf16.json 2>/dev/null
cut -c1-220 gpurun_out/${TAG}_bench_deform1d_bf16.json
python bench.py --workload pathomic --precision bf16 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_pathomic_bf16.json 2>/dev/null
cut -c1-220 gpurun_out/${TAG}_bench_pathomic_bf16.json
python tools/step_profile.py 8 bf16 > gpurun_out/${TAG}_step_profile_transmil_b8.txt 2>&1
python tools/step_profile.py 8 bf16 deform2d > gpurun_out/${TAG}_step_profile_deform2d_bf16.txt 2>&1
python tools/stress_100k.py 2>/dev/null | tail -1 > gpurun_out/${TAG}_stress_100k.json
cat gpurun_out/${TAG}_stress_100k.json
python tools/diag_deform_tc.py time > gpurun_out/${TAG}_deform_tc_vs_simt.txt 2>&1
python tools/prof_all.py 3 bf16 > gpurun_out/${TAG}_kernel_times_plain.txt 2>&1
if [ "${LAUNCHES:-0}" = "1" ]; then
  # launch list of ONE steady-state train step of the bench command: skip the 3 warm-up steps (~411 launches each)
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1240 -c 420 --csv --log-file gpurun_out/${TAG}_launches_b8.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_run.log 2>&1
  python tools/agg_launches.py gpurun_out/${TAG}_launches_b8.csv 30 > gpurun_out/${TAG}_launches_b8_by_kernel.txt
  head -12 gpurun_out/${TAG}_launches_b8_by_kernel.txt
fi
