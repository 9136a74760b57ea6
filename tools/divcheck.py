import sys, ctypes as C
sys.path.insert(0,'/root/repo')
import cascadeclassifier_b200 as ccb
from cascadeclassifier_b200 import _lib
ctx=ccb.Context(0)
for steps in (1,2):
    tot=0
    for seed in range(4):
        m=C.c_uint64(0)
        _lib.check(ctx.lib.cc_debug_div_check(ctx.h, steps, seed*7919+1, 4_000_000_000, C.byref(m)))
        tot+=m.value
    print("steps",steps,"mismatches in 1.6e10 pairs:",tot)
