#!/bin/bash
# Round measurement pass on one B200 (run under gpurun): benches of every workload + per-shape GEMM table + step profiles +
# ncu launch list / full captures.  Outputs under gpurun_out/ (copy what should be judged into profiles/).
set -u
TAG=${1:-r02}
O=gpurun_out
python bench.py --steps 20 --warmup 5 --prof-out $O/${TAG}_gemm_shapes_b8.txt > $O/${TAG}_bench_transmil_b8.json 2> $O/${TAG}_bench.err
cut -c1-400 $O/${TAG}_bench_transmil_b8.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference_arm.json 2>/dev/null
for b in 1 2; do
  python bench.py --batch $b --steps 20 --warmup 5 --no-cpu-baseline --secondary '' > $O/${TAG}_bench_transmil_b$b.json 2>/dev/null
  python bench.py --batch $b --steps 20 --warmup 5 --no-cpu-baseline --secondary '' --graph 0 > $O/${TAG}_bench_transmil_b${b}_eager.json 2>/dev/null
done
python bench.py --workload deform2d --precision bf16 --steps 5 --warmup 3 > $O/${TAG}_bench_deform2d_bf16.json 2>/dev/null
python bench.py --workload deform2d --precision bf16x3 --steps 5 --warmup 3 > $O/${TAG}_bench_deform2d_bf16x3.json 2>/dev/null
python bench.py --workload deform1d --precision bf16 --steps 3 --warmup 3 > $O/${TAG}_bench_deform1d_bf16.json 2>/dev/null
python tools/step_profile.py 8 bf16 > $O/${TAG}_step_profile_transmil_b8.txt 2>&1
python tools/step_profile.py 8 bf16 deform2d > $O/${TAG}_step_profile_deform2d_bf16.txt 2>&1
python tools/time_deform_fwd.py 8 lean:0 old:2 split:1 > $O/${TAG}_deform_kernels.txt 2>&1
python tools/bench_gather.py 64 > $O/${TAG}_gather_roofline.txt 2>&1
./tools/ubench_tmem.bin > $O/${TAG}_ubench_tmem.txt 2>&1
python tools/prof_all.py 3 bf16 > $O/${TAG}_kernel_times_plain.txt 2>&1
if [ "${NCU:-1}" = "1" ]; then
  # launch list of steady-state train steps of the bench command (eager launches: the graph replays the same kernels)
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --secondary '' --graph 0 > $O/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 620 --csv --log-file $O/${TAG}_launches_b8.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --secondary '' --graph 0 > $O/${TAG}_ncu_run.log 2>&1
  python tools/agg_launches.py $O/${TAG}_launches_b8.csv 30 > $O/${TAG}_launches_b8_by_kernel.txt
  head -12 $O/${TAG}_launches_b8_by_kernel.txt
  ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|nystrom_a1_bwd|deform_attn" -c 9 -o $O/${TAG}_prof_all -f \
      python tools/prof_all.py 1 bf16 > $O/${TAG}_ncu_prof_all.log 2>&1
  ncu --set full --clock-control none -k regex:"gather_fwd_vec|gather_bwd_vec|gather_bwd_sorted" -c 3 -o $O/${TAG}_gather -f \
      python tools/bench_gather.py 64 > $O/${TAG}_ncu_gather.log 2>&1
fi
