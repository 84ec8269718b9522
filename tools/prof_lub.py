#!/usr/bin/env python3
"""Per-phase clock64 profile of the blocked getrf (lu_blocked.cuh) on a debug build of the library:

    cd rust_ida_b200/csrc && make
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -DLUB_PROFILE \
         -c linalg_batched.cu -o ../../build/linalg_prof.o
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/lib_lubprof.so ida_engine.o ../../build/linalg_prof.o \
         --cudart static
    B200IDA_LIB=build/lib_lubprof.so python tools/prof_lub.py 100        # n = 100, benchmark batch

Prints, for CTA 0, the cycles warp 0 and warp 1 spent in each phase of lub_factor (panel, [S], strip, panel || rest,
barrier waits, staging) and the breakdown of a panel column step (load+keys, reduction, pivot+reciprocal, scale,
update of the remaining panel columns, syncwarp, permutation).  The production build contains none of this."""
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from rust_ida_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
batch = math.ceil(9 * (1 << 20) / (n * n))
L = _lib.lib()
g = torch.Generator(device="cpu").manual_seed(1234 + n)
a = torch.rand((batch, n, n), generator=g, dtype=torch.float64) / n
a += torch.flip(torch.eye(n, dtype=torch.float64), dims=[0])[None]
A = a.transpose(1, 2).contiguous().cuda()
piv = torch.empty((n * batch,), dtype=torch.int64, device="cuda")
info = torch.empty((batch,), dtype=torch.int32, device="cuda")
out = (ctypes.c_ulonglong * 32)()
L.b200ida_debug_lub_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.b200ida_debug_lub_prof(None, 1)
st = torch.cuda.current_stream().cuda_stream
_lib.check(L.b200ida_getrf_batched(A.data_ptr(), piv.data_ptr(), info.data_ptr(), n, batch, _lib.LAYOUT_AOS, _lib.MEM_DEVICE, st))
torch.cuda.synchronize()
L.b200ida_debug_lub_prof(out, 0)
names = ["panel0", "S", "S-wait", "strip", "strip-wait", "panel|rest", "end-wait", "stage-in"]
tot0 = sum(out[i] for i in range(8))
for w, off in (("warp0", 0), ("warp1", 8)):
    print(w, {names[i]: int(out[off + i]) for i in range(8)})
print("panel internals (warp0 lane0, all panels):", dict(zip(["load+keys", "redux", "pv+rcp", "scale", "updates", "syncwarp", "permute+moves"], [int(out[16 + i]) for i in range(7)])))
print("cycles CTA0 total (warp0 phases)", tot0, "nsteps", (n + 7) // 8)
