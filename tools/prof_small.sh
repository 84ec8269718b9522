#!/bin/bash
# heat M x M (warp-team kernel) throughput for each library variant given on the command line
for v in "$@"; do
  for M in 4 5; do
    echo -n "$v M=$M "
    B200IDA_LIB=variants/libb200ida_$v.so timeout 120 python tools/prof_heat.py 65536 $M 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read().strip()); print(round(d['instances_per_s']), d['all_success'])"
  done
done
