#!/bin/bash
# same-box A/B of tuning builds on the C5 shard (fp32): tools/c5_variants.sh tag1 tag2 ...   ("base" = the shipped library)
cd "$(dirname "$0")/.."
for tag in "$@"; do
  if [ "$tag" = base ]; then unset HSB_LIB; else export HSB_LIB=$PWD/hisparse_b200/libhsb_$tag.so; fi
  for rep in 1 2; do
    echo -n "$tag: "; timeout 300 python tests/c5_probe.py --impl float_pob --no-check --steps 10 2>&1 | tail -1 | python -c "import sys,json; d=json.load(sys.stdin); print(d['ms_per_spmv'], d['ms_isolated_launch'], d['layout'])"
  done
done
