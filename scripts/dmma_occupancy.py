"""DMMA.8x8x4 throughput against resident warps per SM sub-partition and independent accumulator tiles per warp
(jc_fp64_peak_tflops sub-modes).  Answers: how many ready warps does the FP64 tensor path need to stay full?"""
from jax_cosmo_b200 import _native

for wide in (0, 1):
    for wps in (1, 2, 3, 4):
        t = _native.fp64_peak_tflops(1 | (wps << 4) | (wide << 8), 0.3)
        print("acc tiles/warp %2d  warps/SMSP %d  DMMA %.2f TFLOP/s" % (16 if wide else 4, wps, t))
