"""Stand-alone K(X*, X) (bb_kernel_matrix) at the config-2 size: parity against the oracle on a row sample,
CUDA-event timing with the L2 flushed, achieved HBM bandwidth on the algorithmic 4d + 4n bytes per candidate."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import oracle
from baybe_b200 import DeviceGP
from baybe_b200.synthetic import numeric_grid_workload, task_workload
from tests.helpers import oracle_model

dev = torch.device("cuda", 0)
out = {}
for name, w in (("cfg2_1M", numeric_grid_workload(N=1_000_000, d=20, n=256)),
                ("n100_d7_rbf_scaled", numeric_grid_workload(N=50_001, d=7, n=100, family="rbf", outputscale=2.5,
                                                            lengthscale=np.linspace(0.4, 1.5, 7), seed=3)),
                ("task4", task_workload(N_per_task=20_000, n_tasks=4, d_num=6, n_per_task=40, seed=2))):
    om = oracle_model(w)
    gp = DeviceGP(device=dev, **w.gp_kwargs())
    x = torch.from_numpy(w.candidates).to(dev, torch.float32)
    K = gp.kernel_matrix(x)
    torch.cuda.synchronize()
    rows = np.random.default_rng(0).choice(len(w.candidates), size=2000, replace=False)
    rows[:3] = [0, len(w.candidates) - 1, len(w.candidates) // 2]
    Xn = (torch.from_numpy(w.candidates[rows]) - om.lo) / om.rng
    ref = oracle.kernel_matrix(om.spec, Xn, om.Xn)
    err = float((K[torch.from_numpy(rows).to(dev)].double().cpu() - ref).abs().max())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(7):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gp.kernel_matrix(x)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts[2:]))
    N, n, d = len(w.candidates), gp.n, gp.d
    out[name] = {"N": N, "n": n, "d": d, "max_abs_err": err, "ms": ms,
                 "GBps_algorithmic": N * (4 * d + 4 * n) / ms / 1e6}
    print(name, json.dumps(out[name]), flush=True)
    del gp, x, K
print("KMAT_DONE")
