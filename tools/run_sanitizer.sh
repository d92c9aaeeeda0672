#!/bin/bash
# compute-sanitizer over tests/sanitizer_smoke.py with every tool (GPU box; output: gpurun_out/<tag>_compute_sanitizer.txt)
OUT=gpurun_out/${1:-r03}_compute_sanitizer.txt
: > $OUT
for tool in memcheck racecheck synccheck initcheck; do
    echo "== compute-sanitizer --tool $tool python tests/sanitizer_smoke.py" >> $OUT
    timeout 600 compute-sanitizer --tool $tool python tests/sanitizer_smoke.py 2>&1 | tail -45 >> $OUT
    echo >> $OUT
done
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $OUT | sort | uniq -c
