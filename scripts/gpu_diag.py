"""Stage-by-stage GPU diagnostic (prints max errors vs the CPU oracle). Run under `timeout`."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import oracle
from baybe_b200 import AcqConfig, DeviceGP, sobol_normal_samples
from baybe_b200.engine import decode_best
from baybe_b200.synthetic import numeric_grid_workload, task_workload, mixed_small_workload
from tests.helpers import oracle_model

dev = torch.device("cuda", 0)
print(torch.cuda.get_device_name(0), flush=True)
for name, w in [("cfg2_small", numeric_grid_workload(N=6000, d=20, n=256)),
                ("cfg1", mixed_small_workload()),
                ("n512", numeric_grid_workload(N=3000, d=20, n=512, seed=5)),
                ("task4", task_workload(N_per_task=800, n_tasks=4, d_num=6, n_per_task=40, seed=2))]:
    print("=====", name, flush=True)
    om = oracle_model(w)
    t0 = time.time(); gp = DeviceGP(device=dev, **w.gp_kwargs()); torch.cuda.synchronize()
    print(f"model build {time.time()-t0:.3f}s jitter_tries={gp.model.jitter_tries} r_scale={gp.model.r_scale} n_pad={gp.model.n_pad}", flush=True)
    X = torch.from_numpy(w.candidates)
    Xn = (X - om.lo) / om.rng
    Kref = oracle.kernel_matrix(om.spec, Xn, om.Xn)
    K = gp.kernel_matrix(X).double().cpu()
    print("kernel_matrix max err", float((K - Kref).abs().max()), flush=True)
    mu_ref, var_ref = oracle.posterior(om, w.candidates)
    mu, var = gp.posterior_simt(X); torch.cuda.synchronize()
    print("simt  mu err", float((mu.double().cpu() - mu_ref).abs().max()), "var err", float((var.double().cpu() - var_ref).abs().max()), flush=True)
    mu, var = gp.posterior(X); torch.cuda.synchronize()
    emu = (mu.double().cpu() - mu_ref).abs(); evar = (var.double().cpu() - var_ref).abs()
    print("tcgen mu err", float(emu.max()), "var err", float(evar.max()), "var range", float(var_ref.min()), float(var_ref.max()), flush=True)
    if float(evar.max()) > 1e-3:
        bad = torch.argsort(evar, descending=True)[:8]
        print(" worst rows", bad.tolist(), var[bad].tolist(), var_ref[bad].tolist())
        print(" first rows", var[:8].tolist(), var_ref[:8].tolist())
    z = sobol_normal_samples(512, 1, 1234)
    for kind in ["qLogEI", "qEI", "UCB", "LogEI"]:
        oacq = oracle.AcqSpec(kind=kind); oacq.best_f = oracle.best_f_from_training(om, w.train_x, oacq)
        acq = AcqConfig(kind=kind, best_f=gp.best_f(AcqConfig(kind=kind)))
        sc, key = gp.score(acq, X.to(dev, torch.float32), z[:, 0] if acq.is_mc else None); torch.cuda.synchronize()
        ref = oracle.acq_values(om, oacq, w.candidates, z[:, 0] if acq.is_mc else None)
        e = (sc.double().cpu() - ref).abs()
        print(f"{kind}: best_f {acq.best_f:.6f}/{oacq.best_f:.6f} score err max {float(e.max()):.3e} med {float(e.median()):.3e} argmax gpu {decode_best(key)} ref {int(ref.argmax())} {float(ref.max()):.5f}", flush=True)
    # joint
    P = 3
    pend = w.candidates[:P]; cand = w.candidates[P:P + 1000]
    zj = sobol_normal_samples(512, 1 + P, 99)
    oacq = oracle.AcqSpec(kind="qLogEI"); oacq.best_f = oracle.best_f_from_training(om, w.train_x, oacq)
    got = gp.score_joint(AcqConfig(kind="qLogEI", best_f=oacq.best_f), torch.from_numpy(cand), pend, zj).double().cpu()
    ref = oracle.acq_values_joint(om, oacq, cand, pend, zj)
    print("joint qLogEI err", float((got - ref).abs().max()), flush=True)

# timing at 1M
w = numeric_grid_workload(N=1_000_000, d=20, n=256)
gp = DeviceGP(device=dev, **w.gp_kwargs())
x = torch.from_numpy(w.candidates).to(dev, torch.float32)
z = sobol_normal_samples(512, 1, 1234)[:, 0]
acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
for label, fn in [("fused qLogEI", lambda: gp.score(acq, x, z, want_scores=False)),
                  ("posterior", lambda: gp.posterior(x)),
                  ("kernel_matrix", lambda: gp.kernel_matrix(x))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label}: {ms:.3f} ms per 1M candidates -> {1e6/ms*1e3:.3e} cand/s", flush=True)
