#!/bin/bash
# Re-capture of the kernels changed late in round 2 (finalize_quad, Poisson sampler, thin-layer kernels) + launch lists.
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
COMMON="--no-e2e --no-cpu-baseline --no-other-workloads --no-ncu-side-run"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_c2_launches.csv \
  python bench.py --steps 20 --warmup 3 $COMMON > $O/${TAG}_c2_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_c5_launches.csv \
  python bench.py --workload c5 --steps 10 --warmup 3 $COMMON > $O/${TAG}_c5_launches.log 2>&1
full() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o $O/${TAG}_$name \
    python bench.py "$@" --steps 3 --warmup 3 $COMMON > $O/${TAG}_$name.log 2>&1
  ncu -i $O/${TAG}_$name.ncu-rep --page raw --csv > $O/${TAG}_${name}_raw.csv 2>/dev/null
  rm -f $O/${TAG}_$name.ncu-rep
}
full vae_mid "vae_mid_(fwd|bwd)_mma" 6 2 --workload c5
full finalize_c5 "finalize_(quad_)?kernel" 3 2 --workload c5
full poisson "poisson_(select|compact)" 8 2 --workload c2
python scripts/kernel_times.py --workload c5 > $O/${TAG}_c5_kernel_times.txt 2> /dev/null
ls $O | grep $TAG
