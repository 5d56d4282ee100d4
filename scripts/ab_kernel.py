"""Same-box A/B timing of library builds: resident config-2 scoring pass (key init + kernels), L2 flushed, CUDA events.
usage: BB_LIB_VARIANT=<name> python scripts/ab_kernel.py   (no variable: the current build)"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from baybe_b200 import AcqConfig, DeviceGP, sobol_normal_samples
from baybe_b200.synthetic import numeric_grid_workload

dev = torch.device("cuda", 0)
w = numeric_grid_workload(N=1_000_000, d=20, n=256)
gp = DeviceGP(device=dev, **w.gp_kwargs())
x = torch.from_numpy(w.candidates).to(dev, torch.float32)
z = sobol_normal_samples(512, 1, 1234)[:, 0].to(dev, torch.float32)
acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    _, key = gp.score(acq, x, z, want_scores=False)
torch.cuda.synchronize()
ts = []
for _ in range(30):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, key = gp.score(acq, x, z, want_scores=False)
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"variant {os.environ.get('BB_LIB_VARIANT', 'current'):12s} median {ts[len(ts)//2]:.4f} ms  min {ts[0]:.4f}  max {ts[-1]:.4f}  key {int(key.item())}", flush=True)
