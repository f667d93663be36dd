"""Run one configuration of the tcgen05 GEMM micro-benchmark (for ncu captures).
Usage: python scripts/one_gemm.py M N K nsplit tile_n stages epi_kind [iters]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iefvad_b200 import _lib  # noqa: E402

M, N, K, nsplit, bn, st, epi = (int(x) for x in sys.argv[1:8])
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 5
ms = C.c_float()
_lib.check(_lib.lib.iefvad_bench_gemm(M, N, K, nsplit, bn, st, epi, iters, C.byref(ms)))
fl = 2.0 * M * N * K
print(f"M={M} N={N} K={K} nsplit={nsplit} bn={bn} stages={st} epi={epi}: {ms.value:.4f} ms "
      f"{fl / ms.value / 1e9:.1f} TF/s alg")
