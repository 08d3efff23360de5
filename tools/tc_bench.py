import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from surf_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH)
lib.surf_tc_bench.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_longlong)]
torch.zeros(1).cuda()
out = (C.c_longlong * 2)()
for N in (64, 128):
    for mode in (4, 5, 6):
        lib.surf_tc_bench(N, 64, mode, out)
        print("N=%3d TS %d issuer threads x 64 mma: done %6d clk -> %.1f clk per mma aggregate" % (N, mode - 3, out[1], out[1] / (64 * (mode - 3))))
for mode, name in ((5, "2 issuers plain"), (10, "2 issuers + commit/6"), (11, "2 issuers + wait + commit/6"), (12, "2 issuers + commit/6 + distinct addr")):
    lib.surf_tc_bench(128, 60, mode, out)
    print("N=128 %-40s 60 mma each: %.1f clk per mma per thread" % (name, out[1] / 60))
for N in ():
    for mode, name in ((0, "TS 1acc"), (1, "TS 2acc"), (2, "SS 1acc"), (3, "SS 2acc")):
        if N == 256 and mode in (1, 3):
            continue
        for reps in (8, 64):
            lib.surf_tc_bench(N, reps, mode, out)
            print("N=%3d %-8s reps=%3d  issue %6d clk (%.1f/mma)  done %6d clk (%.1f/mma)" % (N, name, reps, out[0], out[0] / reps, out[1], out[1] / reps))
