"""Cycles per 8-element epilogue group (16 warps = 4 per SM sub-partition), arithmetic only."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from surf_b200 import _lib
lib = C.CDLL(_lib.LIB_PATH)
lib.surf_epi_bench.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_longlong)]
torch.zeros(1).cuda()
out = (C.c_longlong * 2)()
names = {0: "full (softplus + e-code + sign + split)", 1: "MUFU only-ish (softplus, no split)", 2: "no MUFU (FMA stand-ins) + code + split",
         3: "softplus + split", 4: "softplus only", 11: "full + tcgen05.ld/st", 12: "  + wait::st", 13: "  + fence + elected arrive",
         14: "  + code store to global", 15: "  same, loads pipelined one group ahead",
         21: "level 14 + concurrent TS MMAs", 22: "level 14 + concurrent SS MMAs", 23: "level 14 + one thread parked in mbar_wait", 24: "level 14 + one warp parked in mbar_wait"}
for v in (0, 1, 2, 3, 4, 11, 12, 13, 14, 15, 21, 22, 23, 24):
    out[0] = out[1] = 0
    rc = lib.surf_epi_bench(v, 400, out)
    assert rc == 0, rc
    print("variant %d %-45s %7.1f clk per group of 8 (4 warps / SMSP)  -> %5.0f clk per 128x128 layer" % (v, names[v], out[0] / 400, out[0] / 400 * 4) + (("   [%d MMAs meanwhile: one per %.0f clk]" % (out[1], out[0] / max(out[1], 1))) if v in (21, 22) else ""))
