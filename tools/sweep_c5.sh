set -x
export PYTHONUNBUFFERED=1
for il in 0 1; do for lib in libhsb_ra1.so libhisparse_b200.so; do for impl in fixed float_pob; do
  echo "== il=$il lib=$lib impl=$impl"
  HSB_TILE_INTERLEAVE=$il HSB_LIB=$PWD/hisparse_b200/$lib timeout 200 python tools/c5_probe.py --impl $impl --no-check 2>&1 | tail -1
done; done; done
echo "== parity interleaved"
timeout 300 python tools/c5_probe.py --impl fixed 2>&1 | tail -1
echo "== C2"
HSB_LIB=$PWD/hisparse_b200/libhsb_ra1.so timeout 300 python bench.py --no-cpu-baseline --steps 5 --batch 256 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ra1', d['ms_per_spmv'], d['value'])"
timeout 300 python bench.py --no-cpu-baseline --steps 5 --batch 256 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ra2', d['ms_per_spmv'], d['value'])"
echo "== ncu C5 float"
timeout 600 ncu --set full --clock-control none -k regex:spmv_tiles --launch-skip 4 -c 1 -o gpurun_out/r01c_c5_float -f python tools/c5_probe.py --impl float_pob --no-check --steps 3 2>&1 | tail -3
ls -la gpurun_out/
