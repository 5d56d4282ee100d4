"""First-light check of the fused_ts kernel against the oracle on a few shapes (run under `timeout`)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import oracle
from baybe_b200 import AcqConfig, DeviceGP, sobol_normal_samples
from baybe_b200.engine import decode_best
from baybe_b200.synthetic import mixed_small_workload, numeric_grid_workload, task_workload
from tests.helpers import oracle_model, score_bounds

dev = torch.device("cuda", 0)
cases = {
    "n64_d4": lambda: numeric_grid_workload(N=300, d=4, n=50, seed=1),
    "n128_d7_rbf": lambda: numeric_grid_workload(N=1500, d=7, n=100, family="rbf", outputscale=2.5,
                                                 lengthscale=np.linspace(0.4, 1.5, 7), seed=3),
    "cfg2_small": lambda: numeric_grid_workload(N=6000, d=20, n=256),
    "n192_d12_m32": lambda: numeric_grid_workload(N=2000, d=12, n=180, family="matern32", seed=4, lengthscale=0.8),
    "cfg1": mixed_small_workload,
    "task4": lambda: task_workload(N_per_task=800, n_tasks=4, d_num=6, n_per_task=40, seed=2),
    "cfg2_100k": lambda: numeric_grid_workload(N=100_000, d=20, n=256),
}
only = sys.argv[1:] or list(cases)
z = sobol_normal_samples(512, 1, seed=1234)
for name in only:
    w = cases[name]()
    om = oracle_model(w)
    gp = DeviceGP(device=dev, **w.gp_kwargs())
    x = torch.from_numpy(w.candidates).to(dev, torch.float32)
    t0 = time.time()
    mu, var = gp.posterior(x)
    torch.cuda.synchronize()
    mu_ref, var_ref = oracle.posterior(om, w.candidates)
    emu = float((mu.double().cpu() - mu_ref).abs().max())
    evar = float((var.double().cpu() - var_ref).abs().max())
    print(f"{name}: n_pad={gp.model.n_pad} posterior |dmu| {emu:.2e} |dvar| {evar:.2e} "
          f"(var range {float(var_ref.min()):.2e}..{float(var_ref.max()):.2e}) {time.time() - t0:.2f}s", flush=True)
    for kind in ("qLogEI", "qUCB", "UCB"):
        oacq = oracle.AcqSpec(kind)
        oacq.best_f = oracle.best_f_from_training(om, w.train_x, oacq)
        acq = AcqConfig(kind=kind, best_f=gp.best_f(AcqConfig(kind=kind)))
        scores, key = gp.score(acq, x, z[:, 0] if acq.is_mc else None)
        torch.cuda.synchronize()
        ref, bound = score_bounds(om, oacq, w.candidates, z[:, 0] if acq.is_mc else None)
        err = (scores.double().cpu() - ref).abs()
        val, idx = decode_best(key)
        print(f"   {kind}: max|err| {float(err.max()):.2e} rows over bound {int((err > bound).sum())} "
              f"argmax gpu {idx} oracle {int(ref.argmax())} first-max-of-scores {int(scores.argmax())}", flush=True)
print("FIRST_LIGHT_DONE")
