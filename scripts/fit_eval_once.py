import sys
import numpy as np
sys.path.insert(0, ".")
from baybe_b200.surrogates import DeviceMLL
from baybe_b200.synthetic import numeric_grid_workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = numeric_grid_workload(N=4096, d=20, n=n, seed=0)
y = (w.train_y - w.train_y.mean()) / w.train_y.std(ddof=1)
m = DeviceMLL(w.train_x, y, None, 1, "matern52", "cuda:0")
th = np.concatenate([np.full(20, 0.9), [0.01, 0.0, 1.0]])
for _ in range(3):
    print(m(th)[0])
