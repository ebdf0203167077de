#!/bin/bash
# compute-sanitizer over tests/sanitize_walk.py: tools/sanitize_all.sh <out file>
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/compute_sanitizer.txt}
echo "# compute-sanitizer on tests/sanitize_walk.py (B200)" > $out
for tool in memcheck racecheck synccheck; do
  echo "## $tool" >> $out
  timeout 900 compute-sanitizer --tool $tool python tests/sanitize_walk.py 2>&1 | grep -E "COMPUTE-SANITIZER|SUMMARY|sanitize_run|Error|error|hazard|Traceback|assert" | head -40 >> $out
done
cat $out
