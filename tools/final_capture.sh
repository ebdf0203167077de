#!/bin/bash
# Evidence pass for profiles/ (one B200, under gpurun): launch list of the bench command, one full capture of the bench
# kernel (C2) and of the narrow kernel (C5 shard), DRAM bytes over a range of 512 overlapping launches, (the SASS excerpt needs no GPU: tools/sass_excerpt.py).
# usage: bash tools/final_capture.sh <tag>      (writes gpurun_out/<tag>_*)
set -x
T=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --batch 64 --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.json 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:spmv_tiles --launch-skip 20 -c 1 -o gpurun_out/${T}_c2_full \
    python tools/profile_run.py --spmv 30 > /dev/null 2>&1
ncu -i gpurun_out/${T}_c2_full.ncu-rep --page details > gpurun_out/${T}_ncu_full_spmv_tiles_kernel_c2.txt 2>/dev/null
ncu -i gpurun_out/${T}_c2_full.ncu-rep --page raw --csv > gpurun_out/${T}_c2_full_raw.csv 2>/dev/null
ncu --replay-mode app-range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__bytes.sum.per_second \
    --clock-control none --csv --page raw --log-file gpurun_out/${T}_range_c2.csv python tools/profile_run.py --range 512 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:spmv_tiles --launch-skip 4 -c 1 -o gpurun_out/${T}_c5_full \
    python tests/c5_probe.py --impl float_pob --no-check --steps 3 > /dev/null 2>&1
ncu -i gpurun_out/${T}_c5_full.ncu-rep --page details > gpurun_out/${T}_ncu_full_c5shard_narrow.txt 2>/dev/null
python tools/ncu_traffic.py --kernel-csv gpurun_out/${T}_c2_full_raw.csv --range-csv gpurun_out/${T}_range_c2.csv --range-spmvs 512 \
    --out gpurun_out/${T}_traffic.json
rm -f gpurun_out/${T}_c2_full.ncu-rep gpurun_out/${T}_c5_full.ncu-rep
ls -la gpurun_out | tail -12
