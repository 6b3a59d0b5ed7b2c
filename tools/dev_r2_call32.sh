mkdir -p gpurun_out
for cap in 0 4 3 2; do
  echo "=== MSFL_LM_MAX_CTAS=$cap"
  MSFL_LM_MAX_CTAS=$cap timeout 600 python tools/dev_two_streams.py 2048 20 2>&1 | grep -v "^poses\|^\[" | tail -7
done
