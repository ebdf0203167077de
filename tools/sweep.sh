export PYTHONUNBUFFERED=1
cd "$(dirname "$0")/.."
python tools/sweep_time.py
for sc in 1.0 2.0 2.5; do HSB_SLICE_COST=$sc python tools/sweep_time.py; done
for v in pf3 pf5 pf6 pf5ra2 w24pf5 w16pf8 w16c2pf4 w16c2pf6; do
  HSB_LIB=$PWD/hisparse_b200/libhsb_$v.so timeout 120 python tools/sweep_time.py 2>&1 | tail -1
done
