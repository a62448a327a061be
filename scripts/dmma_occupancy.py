"""DMMA.8x8x4 throughput against resident warps per SM sub-partition and independent accumulator tiles per warp
(jc_fp64_peak_tflops sub-modes).  Answers: how many ready warps does the FP64 tensor path need to stay full?"""
from jax_cosmo_b200 import _native

NAMES = {0: '4 acc tiles, shared operands', 1: '16 acc tiles, shared operands', 2: '2x8 register tile', 3: '2x8 tile, fragments from smem + DMUL'}
for wide in (0, 1, 2, 3):
    for wps in (1, 2, 3, 4):
        t = _native.fp64_peak_tflops(1 | (wps << 4) | (wide << 8), 0.3)
        print("%-40s warps/SMSP %d  DMMA %.2f TFLOP/s" % (NAMES[wide], wps, t))
