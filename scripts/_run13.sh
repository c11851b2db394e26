for tool in memcheck racecheck synccheck; do
  echo "=== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py 2000 > gpurun_out/san_$tool.log 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|dist_eq|hazard|Invalid|error" gpurun_out/san_$tool.log | head -12
done
