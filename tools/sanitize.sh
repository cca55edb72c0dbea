#!/bin/bash
# compute-sanitizer over a small pass of every kernel family (memcheck / racecheck on small batches).
# ONE tool per GPU call (the profiling guide: several sanitizer tools in one call have left a B200 unusable):
#   gpurun --timeout 1500 -- bash tools/sanitize.sh memcheck
#   gpurun --timeout 1500 -- bash tools/sanitize.sh racecheck
# Writes gpurun_out/sanitize_<tool>.log and appends a summary line to gpurun_out/sanitize_summary.txt
# (copied to profiles/r02_sanitize.md once read).
set -u
tool=${1:-memcheck}
OUT=${OUT:-gpurun_out}
mkdir -p "$OUT"
# racecheck / initcheck are slow: the pass is tiny (tools/sanitize_pass.py)
timeout 1300 compute-sanitizer --tool "$tool" --print-limit 20 python tools/sanitize_pass.py > "$OUT/sanitize_$tool.log" 2>&1
rc=$?
echo "$tool: exit $rc; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitize_$tool.log" | tail -1); pass: $(grep -c '^\[pass\]' "$OUT/sanitize_$tool.log") stages ok" | tee -a "$OUT/sanitize_summary.txt"
tail -5 "$OUT/sanitize_$tool.log"
