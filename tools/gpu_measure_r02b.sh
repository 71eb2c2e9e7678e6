#!/bin/bash
# Final measurement pass of round 2 on one B200 (run under gpurun): bench lines of every workload, per-shape GEMM table, step
# profile, ncu launch list of the bench command, ncu --set full of the round's new kernels.  Outputs under gpurun_out/
# (the files that should be judged are copied into profiles/ afterwards).
set -u
TAG=${1:-r02b}
O=gpurun_out
python bench.py --steps 20 --warmup 5 --prof-out $O/${TAG}_gemm_shapes_b8.txt > $O/${TAG}_bench_transmil_b8.json 2> $O/${TAG}_bench.err
cut -c1-300 $O/${TAG}_bench_transmil_b8.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2>/dev/null
python bench.py --workload deform2d --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_deform2d_bf16.json 2>/dev/null
python bench.py --workload deform1d --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_deform1d_bf16.json 2>/dev/null
python bench.py --workload pathomic1d --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pathomic1d_bf16.json 2>/dev/null
python bench.py --workload pathomic --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_pathomic_bf16.json 2>/dev/null
python tools/step_profile.py 8 bf16 2>&1 | grep -v -i warn > $O/${TAG}_step_profile_transmil_b8.txt
python tools/step_profile.py 8 bf16 deform1d 2>&1 | grep -v -i warn > $O/${TAG}_step_profile_deform1d_bf16.txt
python tools/time_deform1d.py 8 --lean > $O/${TAG}_deform1d_kernels.txt 2>&1
python tools/bench_a3.py 8 > $O/${TAG}_a3_kernels.txt 2>&1
python tools/bench_ppeg.py 8 2>&1 | tail -2 > $O/${TAG}_ppeg_kernels.txt
if [ "${NCU:-1}" = "1" ]; then
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --secondary '' --graph 0 > $O/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 620 --csv --log-file $O/${TAG}_launches_b8.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --secondary '' --graph 0 > $O/${TAG}_ncu_run.log 2>&1
  python tools/agg_launches.py $O/${TAG}_launches_b8.csv 30 > $O/${TAG}_launches_b8_by_kernel.txt
  head -12 $O/${TAG}_launches_b8_by_kernel.txt
  ncu --set full --clock-control none --import-source on -k regex:"a3_fwd_kernel" -c 1 -s 3 -o $O/${TAG}_a3_fwd -f python tools/bench_a3.py 8 > $O/${TAG}_ncu_a3.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"a3_bwd_kernel" -c 1 -s 3 -o $O/${TAG}_a3_bwd -f python tools/bench_a3.py 8 >> $O/${TAG}_ncu_a3.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"fwd_pwl_kernel" -c 1 -s 2 -o $O/${TAG}_pwl_fwd -f python tools/time_deform1d.py 1 > $O/${TAG}_ncu_pwl.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"bwd_pwl_kernel" -c 1 -s 2 -o $O/${TAG}_pwl_bwd -f python tools/time_deform1d.py 1 >> $O/${TAG}_ncu_pwl.log 2>&1
fi
