#!/bin/bash
# One GPU-box pass that refreshes the evidence under profiles/: bench lines of every workload, the reference arm,
# the ncu launch list and one ncu --set full capture of the C2 step kernel.  Outputs land in gpurun_out/.
TAG=${1:-v6}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_c2_n1_$TAG.json 2> gpurun_out/bench_c2_n1_$TAG.err
for w in c1 c3 c4 c5; do
  python bench.py --workload $w > gpurun_out/bench_${w}_n1_$TAG.json 2> gpurun_out/bench_${w}_n1_$TAG.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_reference_arm_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c2_launches_$TAG.csv \
  python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads --no-ncu-side-run > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:meanfield_step_vec -s 6 -c 2 -f -o gpurun_out/step_vec_c2_$TAG \
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads --no-ncu-side-run > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
