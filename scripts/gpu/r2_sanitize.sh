#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  IODINE_NO_GRAPH=1 timeout 1200 compute-sanitizer --tool $tool python scripts/sanitize.py > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "== $tool"; grep -E "SUMMARY|ok " gpurun_out/r2_sanitize_$tool.log | tail -9
done
