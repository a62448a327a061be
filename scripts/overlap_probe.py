"""Does the contraction co-reside usefully with the power kernel?  Times, on the bench tracer set,
  (a) power(chunk A) then contraction(chunk B) back to back on one stream,
  (b) the same two kernels on two streams (8-warp contraction kernel limited to one CTA per SM, so that two
      power CTAs fit next to it: 128 x 256 + 2 x 64 x 256 registers = one register file).
Both chunks are fully prepared before the timed region; the work is identical in (a) and (b)."""
import sys

import torch

sys.path.insert(0, ".")
import jax_cosmo_b200 as jc  # noqa: E402
from jax_cosmo_b200 import _native  # noqa: E402
from oracle import scenarios as sc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
scn = sc.scenario("cfg5", sc.PLANCK15, sc.ELL_CFG2, [sc.sources(10, 1.0), sc.lenses(10, 1.0)])
probes = sc.build_probes(scn, jc)
plan = _native.get_plan(probes, scn["ell"], None, None)
lib = _native.load_library()
rows = torch.as_tensor(sc.config5_cosmologies(2 * B), device="cuda")
nbytes = plan.workspace_bytes(B)
wsA = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
wsB = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
clA = torch.empty((B, plan.P, plan.L), dtype=torch.float64, device="cuda")
clB = torch.empty_like(clA)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def stages(mask, r, cl, ws, stream):
    _native.check(lib.jc_debug_stages_f64(plan._h, mask, r.data_ptr(), B, cl.data_ptr(), ws.data_ptr(), ws.numel() * 8,
                                          stream.cuda_stream), "jc_debug_stages_f64")


cur = torch.cuda.current_stream()
stages(1 | 2 | 4 | 8 | 16, rows[:B], clA, wsA, cur)      # both chunks fully prepared (V of chunk B is what the contraction reads)
stages(1 | 2 | 4 | 8 | 16, rows[B:], clB, wsB, cur)
ref = clB.clone()
torch.cuda.synchronize()


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    for _ in range(n):
        fn()
    e1.record(cur)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def sequential(cmask):
    stages(8, rows[:B], clA, wsA, cur)
    stages(cmask, rows[B:], clB, wsB, cur)


def concurrent(cmask):
    ev = torch.cuda.Event()
    ev.record(cur)
    s1.wait_event(ev); s2.wait_event(ev)
    stages(cmask, rows[B:], clB, wsB, s2)   # contraction first: its CTAs take one slot per SM
    stages(8, rows[:B], clA, wsA, s1)
    d1, d2 = torch.cuda.Event(), torch.cuda.Event()
    d1.record(s1); d2.record(s2)
    cur.wait_event(d1); cur.wait_event(d2)


print("chunk %d cosmologies" % B)
print("power alone                         %.3f ms" % timed(lambda: stages(8, rows[:B], clA, wsA, cur)))
print("contraction alone (TMA persistent)  %.3f ms" % timed(lambda: stages(16, rows[B:], clB, wsB, cur)))
print("contraction alone (8 warps, 1/SM)   %.3f ms" % timed(lambda: stages(32, rows[B:], clB, wsB, cur)))
print("sequential power + TMA contraction  %.3f ms" % timed(lambda: sequential(16)))
print("two streams power || 8-warp contr.  %.3f ms" % timed(lambda: concurrent(32)))
print("two streams power || TMA contr.     %.3f ms" % timed(lambda: concurrent(16)))
print("results equal:", bool(torch.equal(clB, ref)))
