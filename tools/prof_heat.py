#!/usr/bin/env python3
"""one heat-2D ensemble run (n = 100, CTA-per-system kernel) sized for an ncu capture"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import bench_configs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M = int(sys.argv[2]) if len(sys.argv) > 2 else 10
print(bench_configs.run_heat(B, reps=0, M=M))
