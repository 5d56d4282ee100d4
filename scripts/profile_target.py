"""Short driver for ncu captures: a few launches of one kernel at the config-2 size."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from baybe_b200 import AcqConfig, DeviceGP, sobol_normal_samples
from baybe_b200.synthetic import numeric_grid_workload

which = sys.argv[1] if len(sys.argv) > 1 else "fused"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dev = torch.device("cuda", 0)
if which == "wide":  # config-4 shard shape: bit-packed 2048-bit fingerprints, n = 512
    from baybe_b200.synthetic import fingerprint_workload
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 262_144
    w = fingerprint_workload(N=4096, d=2048, n=512, seed=1)
    gp = DeviceGP(device=dev, **w.gp_kwargs())
    g = torch.Generator(device="cuda").manual_seed(0)
    bits = torch.rand((N, 256, 8), device=dev, generator=g) < 0.05
    packed = (bits.to(torch.uint8) << torch.arange(8, device=dev, dtype=torch.uint8)).sum(dim=2).to(torch.uint8)
    z = sobol_normal_samples(512, 1, 1234)[:, 0]
    acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
    for _ in range(3):
        gp.score(acq, packed, z, want_scores=False)
    torch.cuda.synchronize()
    print("done wide")
    sys.exit(0)
w = numeric_grid_workload(N=N, d=20, n=256)
gp = DeviceGP(device=dev, **w.gp_kwargs())
x = torch.from_numpy(w.candidates).to(dev, torch.float32)
z = sobol_normal_samples(512, 1, 1234)[:, 0]
acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
for _ in range(4):
    if which == "fused":
        gp.score(acq, x, z, want_scores=False)
    elif which == "posterior":
        gp.posterior(x)
    elif which == "kmat":
        gp.kernel_matrix(x)
torch.cuda.synchronize()
print("done", which)
