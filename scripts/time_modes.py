"""Per-mode timing of the fused kernel at the config-2 size (CUDA events, 10 repetitions, median)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from baybe_b200 import AcqConfig, DeviceGP, sobol_normal_samples
from baybe_b200.synthetic import numeric_grid_workload

dev = torch.device("cuda", 0)
w = numeric_grid_workload(N=1_000_000, d=20, n=256, seed=0)
gp = DeviceGP(device=dev, **w.gp_kwargs())
x = torch.from_numpy(w.candidates).to(dev, torch.float32)
z = sobol_normal_samples(512, 1, 1234)[:, 0].to(dev, torch.float32)
bf = gp.best_f(AcqConfig(kind="qLogEI"))


def med(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


print("posterior (mu,var)      %.3f ms" % med(lambda: gp.posterior(x)))
for kind in ("qLogEI", "qEI", "UCB", "PM"):
    acq = AcqConfig(kind=kind, best_f=bf)
    zz = z if acq.is_mc else None
    print("score %-6s key only   %.3f ms" % (kind, med(lambda: gp.score(acq, x, zz, want_scores=False))))
    print("score %-6s + scores   %.3f ms" % (kind, med(lambda: gp.score(acq, x, zz, want_scores=True))))

# test-only counters: rows outside / inside the tabulated qLogEI envelope
import ctypes as C
from baybe_b200 import _lib
cap = 64
buf = torch.zeros(2 * cap + 8, dtype=torch.int64, device=dev)
_lib.load().bb_debug_set_trace(C.c_void_p(buf.data_ptr()), cap)
gp.score(AcqConfig(kind="qLogEI", best_f=bf), x, z, want_scores=False)
torch.cuda.synchronize()
_lib.load().bb_debug_set_trace(None, 0)
print("rows outside envelope", int(buf[1 + 2 * cap]), "inside", int(buf[2 + 2 * cap]))
