"""Print the FP64 roofline probes of the current GPU: DFMA-only, DMMA-only and both interleaved."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from jax_cosmo_b200 import _native  # noqa: E402

torch.cuda.init()
for mode, name in ((0, "DFMA"), (1, "DMMA m8n8k4"), (2, "DFMA+DMMA interleaved")):
    print("%-24s %.2f TFLOP/s" % (name, _native.fp64_peak_tflops(mode, 0.5)))
