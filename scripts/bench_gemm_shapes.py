"""tcgen05 GEMM micro-benchmark over shapes / epilogue kinds / ring depths:
python scripts/bench_gemm_shapes.py M "N K epi stages" ["N K epi stages" ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iefvad_b200 import _lib  # noqa: E402

M = int(sys.argv[1])
for spec in sys.argv[2:]:
    N, K, epi, stages = (int(x) for x in spec.split())
    ms = C.c_float()
    _lib.check(_lib.lib.iefvad_bench_gemm(M, N, K, 1, 512, stages, epi, 30, C.byref(ms)))
    print(f"M={M} N={N} K={K} epi_kind={epi} stages={stages or 'max'}: {ms.value * 1e3:.1f} us  "
          f"{2.0 * M * N * K / ms.value / 1e9:.0f} TF/s", flush=True)
