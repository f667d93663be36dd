"""Micro-benchmark of the tcgen05 GEMM (iefvad_bench_gemm): mainloop-only vs the real epilogues, per tile width.
Usage (on the GPU box): python scripts/bench_gemm.py [M]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iefvad_b200 import _lib  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
EPI = {0: "discard", 1: "f32", 2: "refine(resid,f32,hi,lo)", 3: "relu->hi,lo", 4: "qkv", 5: "fp16 relu->h",
       6: "fp16 refine(resid,f32,h)"}
print(f"{'N':>5} {'K':>5} {'split':>5} {'BN':>4} {'stg':>3} {'epilogue':>24} {'ms':>8} {'TF/s alg':>9} {'TF/s mma':>9}")
cases = []
for bn in (512, 256):
    for nsplit in (1, 3):
        for st in (0, 3):
            cases.append((768, 768, nsplit, bn, st, 0))
        for epi in ((1, 2, 3, 5, 6) if nsplit == 1 else (1, 2, 3)):
            cases.append((768, 768, nsplit, bn, 0, epi))
    cases.append((2304, 768, 1, bn, 0, 0))
    cases.append((2304, 768, 1, bn, 0, 4))
    cases.append((1536, 768, 3, bn, 0, 0))
    cases.append((1536, 768, 3, bn, 0, 1))
for N, K, nsplit, bn, st, epi in cases:
    ms = C.c_float()
    _lib.check(_lib.lib.iefvad_bench_gemm(M, N, K, nsplit, bn, st, epi, 20, C.byref(ms)))
    fl = 2.0 * M * N * K
    print(f"{N:5d} {K:5d} {nsplit:5d} {bn:4d} {st:3d} {EPI[epi]:>24} {ms.value:8.4f} {fl / ms.value / 1e9:9.1f} "
          f"{fl * nsplit / ms.value / 1e9:9.1f}", flush=True)
