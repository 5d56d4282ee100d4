"""Wall time of the MAP hyper-parameter fit: host (float64 torch autograd on the CPU) vs device
(bb_fit_eval under the same scipy L-BFGS-B driver), config-2 training set (n = 256, d = 20) and n = 512."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from baybe_b200.surrogates import fit_map, fit_map_hyperparameters_device  # noqa: E402
from tests.helpers import HostMLL  # noqa: E402  (float64 autograd twin, test infrastructure)
from baybe_b200.synthetic import numeric_grid_workload  # noqa: E402

for n in (256, 512):
    w = numeric_grid_workload(N=4096, d=20, n=n, seed=0)
    y = (w.train_y - w.train_y.mean()) / w.train_y.std(ddof=1)
    active = list(range(20))
    fit_map_hyperparameters_device(w.train_x[:32], y[:32], active, None, 1, 5, device="cuda:0")  # warm-up
    t0 = time.perf_counter()
    dev = fit_map_hyperparameters_device(w.train_x, y, active, None, 1, 200, device="cuda:0")
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    host = fit_map(w.train_x, y, active, None, 1, 200, mll_factory=HostMLL)
    t2 = time.perf_counter()
    print(json.dumps(dict(n=n, d=20, device_s=t1 - t0, host_s=t2 - t1, device_evals=dev["n_eval"],
                          device_ms_per_eval=(t1 - t0) / dev["n_eval"] * 1e3, host_iters=host["n_iter"],
                          objective_device=dev["objective"], objective_host=host["objective"],
                          host_threads=torch.get_num_threads())))
