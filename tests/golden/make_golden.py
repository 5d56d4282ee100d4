"""Generate the golden fixtures in this directory from the CPU oracle.

    python -m tests.golden.make_golden

The reference's own implementation of this path (botorch/gpytorch) cannot be imported offline, so
these vectors are ORACLE outputs on seeded synthetic workloads: they pin the oracle (and, through
the GPU parity tests, the CUDA path) against drift, not against BoTorch."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

import oracle
from baybe_b200.synthetic import (fingerprint_workload, mixed_small_workload, numeric_grid_workload,
                                  task_workload)
from tests.helpers import oracle_model

HERE = Path(__file__).parent
WORKLOADS = {
    "cfg1": mixed_small_workload,
    "cfg2_slice": lambda: numeric_grid_workload(N=512, d=20, n=256),
    "task": lambda: task_workload(N_per_task=96, n_tasks=4, d_num=6, n_per_task=24, seed=2),
    # BASELINE config 4 shape (2048-bit fingerprints, n = 512, ScaleKernel(RBF)) at 600 candidates
    "cfg4_slice": lambda: fingerprint_workload(N=600, d=2048, n=512, seed=1),
}
ROWS = 48


def evaluate(w) -> dict:
    om = oracle_model(w)
    X = w.candidates[:ROWS]
    mu, var = oracle.posterior(om, X)
    out = {"mean": mu.tolist(), "variance": var.tolist()}
    z = oracle.sobol_normal_samples(512, 1, 1234)[:, 0]
    for kind in ("qLogEI", "qEI", "qUCB", "UCB", "LogEI"):
        a = oracle.AcqSpec(kind)
        a.best_f = oracle.best_f_from_training(om, w.train_x, a)
        out[kind] = oracle.acq_values(om, a, X, z if a.is_mc else None).tolist()
        if kind == "qLogEI":
            out["best_f"] = [a.best_f]
            idx, vals = oracle.optimize_acqf_discrete(om, a, w.candidates[:256], q=3, sampler_seed=1234)
            out["greedy_idx"] = idx
            out["greedy_val"] = vals
    return out


def main():
    for name, make in WORKLOADS.items():
        w = make()
        payload = {"workload": w.name, "rows": ROWS, "generator": "tests/golden/make_golden.py",
                   "note": "oracle outputs (float64); parity with botorch is unpinned",
                   "values": evaluate(w)}
        (HERE / f"{name}.json").write_text(json.dumps(payload, indent=1))
        print("wrote", name)


if __name__ == "__main__":
    main()
