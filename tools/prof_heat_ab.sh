#!/bin/bash
# heat 10x10 (CTA-team kernel) throughput for each library variant on the command line, one warm-up + one timed run each,
# interleaved twice so that run-to-run drift shows
for rep in 1 2; do
  for v in "$@"; do
    echo -n "$v "
    B200IDA_LIB=variants/libb200ida_$v.so timeout 200 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import bench_configs
d = bench_configs.run_heat(32768, reps=1)
print(round(d["instances_per_s"]), d["all_success"])
PY
  done
done
