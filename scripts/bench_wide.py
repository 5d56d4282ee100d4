"""Timing of the wide-feature path (wide.cu) on one GPU: BASELINE config-4 shard shape (bit-packed
2048-bit fingerprints, n = 512, ScaleKernel(RBF), qLogEI) and a float descriptor space (d = 128).
Prints one JSON object per case; CUDA-event timing, L2 flushed between repetitions."""
from __future__ import annotations

import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from baybe_b200 import AcqConfig, DeviceGP, sobol_normal_samples  # noqa: E402
from baybe_b200.synthetic import fingerprint_workload, numeric_grid_workload, pack_bits  # noqa: E402


def timed(fn, reps=5, warm=2):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    dev = torch.device("cuda:0")
    z = sobol_normal_samples(512, 1, seed=1234)[:, 0].to(dev, torch.float32)
    out = []
    # ---- config 4 shard: 1.25M x 2048 bits, n = 512 ----
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
    w = fingerprint_workload(N=4096, d=2048, n=512, seed=1)
    gp = DeviceGP(device=dev, **w.gp_kwargs())
    g = torch.Generator(device="cuda").manual_seed(0)
    bits = (torch.rand((N, 2048 // 8, 8), device=dev, generator=g) < 0.05)
    packed = (bits.to(torch.uint8) << torch.arange(8, device=dev, dtype=torch.uint8)).sum(dim=2).to(torch.uint8)
    del bits
    acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
    K = torch.empty((min(N, 262144), 512), dtype=torch.float32, device=dev)
    t_k = timed(lambda: gp.kernel_matrix(packed[: K.shape[0]]))
    t_s = timed(lambda: gp.score(acq, packed, z, want_scores=False))
    t_p = timed(lambda: gp.posterior(packed))
    flops = 2.0 * 512 * 2048
    out.append(dict(case="cfg4_shard_bits", N=N, d=2048, n=512, ms_score=t_s, ms_posterior=t_p,
                    cand_per_s=N / t_s * 1e3, ms_kmat_262k=t_k,
                    kmat_dense_tflops=K.shape[0] * flops / t_k / 1e9,
                    # the bit-linear form issues TWO fp16 split products (W hi, W mid): issued MMA work = 2x dense
                    kmat_mma_tflops=2 * K.shape[0] * flops / t_k / 1e9))
    del packed, gp
    # ---- float descriptors: 1M x 128, n = 256, Matern-5/2 ----
    N2 = 1_000_000
    w2 = numeric_grid_workload(N=N2, d=128, n=256, seed=3, lengthscale=3.0)
    gp2 = DeviceGP(device=dev, **w2.gp_kwargs())
    x = torch.from_numpy(w2.candidates).to(dev, torch.float32)
    acq2 = AcqConfig(kind="qLogEI", best_f=gp2.best_f(AcqConfig(kind="qLogEI")))
    t_s2 = timed(lambda: gp2.score(acq2, x, z, want_scores=False))
    t_k2 = timed(lambda: gp2.kernel_matrix(x[:262144]))
    out.append(dict(case="float_d128_n256", N=N2, ms_score=t_s2, cand_per_s=N2 / t_s2 * 1e3, ms_kmat_262k=t_k2,
                    kmat_mma_tflops=6 * 262144 * 2.0 * 256 * 128 / t_k2 / 1e9))
    del x, gp2
    # ---- config 5: 4 tasks x 250k, d = 20 + task column, n = 512 ----
    from baybe_b200.synthetic import task_workload
    w5 = task_workload(N_per_task=250_000, n_tasks=4, d_num=20, n_per_task=128, seed=0)
    gp5 = DeviceGP(device=dev, **w5.gp_kwargs())
    x5 = torch.from_numpy(w5.candidates).to(dev, torch.float32)
    acq5 = AcqConfig(kind="qLogEI", best_f=gp5.best_f(AcqConfig(kind="qLogEI")))
    t_s5 = timed(lambda: gp5.score(acq5, x5, z, want_scores=False))
    t_p5 = timed(lambda: gp5.posterior(x5))
    out.append(dict(case="cfg5_tasks", wide=int(gp5.model.wide), N=x5.shape[0], ms_score=t_s5, ms_posterior=t_p5,
                    cand_per_s=x5.shape[0] / t_s5 * 1e3))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
