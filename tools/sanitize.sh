#!/bin/bash
# compute-sanitizer over a small pass of every kernel family (memcheck / racecheck on small batches).
# Run on a GPU box:   gpurun --timeout 1500 -- bash tools/sanitize.sh
# Writes gpurun_out/sanitize_<tool>.log and a summary line per tool to gpurun_out/sanitize_summary.txt
# (copied to profiles/r02_sanitize.md by hand once read).
set -u
OUT=${OUT:-gpurun_out}
mkdir -p "$OUT"
: > "$OUT/sanitize_summary.txt"
for tool in memcheck racecheck synccheck initcheck; do
    # racecheck / initcheck are slow: the pass is tiny (tools/sanitize_pass.py)
    timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_pass.py > "$OUT/sanitize_$tool.log" 2>&1
    rc=$?
    echo "$tool: exit $rc; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/sanitize_$tool.log" | tail -1); pass: $(grep -c '^\[pass\]' "$OUT/sanitize_$tool.log") stages ok" >> "$OUT/sanitize_summary.txt"
done
cat "$OUT/sanitize_summary.txt"
