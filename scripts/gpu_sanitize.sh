mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "ok$|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -30
done
