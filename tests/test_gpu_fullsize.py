"""Full-size GPU tests (BASELINE config 2: 1,000,000 x 20 candidates, n = 256) through
size-independent properties, plus an oracle spot-check on a random row sample:
  * scores of a 4096-row random sample agree with the float64 oracle,
  * the fused arg-max equals the first maximum of the returned score vector,
  * sharding invariance: max over per-shard keys (with global offsets) == key of the whole set,
  * idempotence / determinism: two passes give bit-identical scores,
  * layout invariance at scale (fp32 row-major vs the reference's fp64 column-major),
  * the keep mask removes exactly the masked rows from contention.
BASELINE config 5 (4 tasks x 250k, ICM kernel) is checked the same way."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import oracle
from baybe_b200 import AcqConfig, DeviceGP, sobol_normal_samples
from baybe_b200.engine import decode_best, unpack_best
from baybe_b200.synthetic import numeric_grid_workload, task_workload
from tests.helpers import oracle_model, score_bounds

pytestmark = pytest.mark.gpu


def _check_fullsize(w, dev, n_shards=3):
    om = oracle_model(w)
    gp = DeviceGP(device=dev, **w.gp_kwargs())
    N = len(w.candidates)
    z = sobol_normal_samples(512, 1, seed=1234)
    oacq = oracle.AcqSpec("qLogEI")
    oacq.best_f = oracle.best_f_from_training(om, w.train_x, oacq)
    acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
    x32 = torch.from_numpy(w.candidates).to(dev, torch.float32)
    scores, key = gp.score(acq, x32, z[:, 0])
    val, idx = decode_best(key)
    # (1) oracle spot-check
    rows = np.random.default_rng(0).choice(N, size=4096, replace=False)
    ref, bound = score_bounds(om, oacq, w.candidates[rows], z[:, 0])  # hard per-row bound, no outlier allowance
    got = scores[torch.from_numpy(rows).to(dev)].double().cpu()
    err = (got - ref).abs()
    assert bool((err <= bound).all()), (float(err.max()), int(torch.argmax(err - bound)))
    # (2) arg-max == first maximum of the score vector
    assert idx == int(torch.argmax(scores)) and val == float(scores[idx])
    # (3) sharding invariance with global offsets
    bounds = np.linspace(0, N, n_shards + 1).astype(int)
    keys = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        _, k = gp.score(acq, x32[lo:hi], z[:, 0], index_offset=int(lo), want_scores=False)
        keys.append(int(k.item()))
    assert unpack_best(max(keys)) == (val, idx)
    # (4) determinism
    scores2, key2 = gp.score(acq, x32, z[:, 0])
    assert torch.equal(scores, scores2) and int(key2.item()) == int(key.item())
    # (5) layout invariance: the reference's float64 column-major matrix, no conversion pass
    x64c = torch.from_numpy(w.candidates).to(dev).t().contiguous().t()
    scores3, key3 = gp.score(acq, x64c, z[:, 0])
    assert torch.equal(scores, scores3) and int(key3.item()) == int(key.item())
    # (6) keep mask
    keep = torch.ones(N, dtype=torch.uint8, device=dev)
    top = torch.topk(scores, 3).indices
    keep[top[:2]] = 0
    _, key4 = gp.score(acq, x32, z[:, 0], keep=keep, want_scores=False)
    masked = scores.clone()
    masked[top[:2]] = -float("inf")
    assert decode_best(key4)[1] == int(torch.argmax(masked))
    return gp, scores


def test_config2_one_million_candidates(cuda_device):
    w = numeric_grid_workload(N=1_000_000, d=20, n=256)
    gp, scores = _check_fullsize(w, cuda_device)
    # recommended index: the float64 oracle over ALL 1,000,000 rows must pick the same row
    om = oracle_model(w)
    oacq = oracle.AcqSpec("qLogEI")
    oacq.best_f = oracle.best_f_from_training(om, w.train_x, oacq)
    z = sobol_normal_samples(512, 1, seed=1234)
    ref_all = oracle.acq_values(om, oacq, w.candidates, z[:, 0], chunk=32768)
    assert int(torch.argmax(ref_all)) == int(torch.argmax(scores)), (
        int(torch.argmax(ref_all)), int(torch.argmax(scores)), torch.topk(ref_all, 3).values.tolist())
    # posterior at the training rows: variance collapses to ~noise level, mean interpolates
    mu, var = gp.posterior(torch.from_numpy(w.train_x))
    assert float(var.max()) < 0.05 * float(np.var(w.train_y)) and float(var.min()) > 0
    assert float((mu.cpu().double() - torch.from_numpy(w.train_y)).abs().max()) < 0.2 * float(np.std(w.train_y))
    # top-k consistency
    vals, idx = torch.ops.baybe_b200.topk(scores, None, 8)
    ref_vals, _ = torch.topk(scores, 8)
    assert torch.equal(vals, ref_vals)


@pytest.mark.parametrize("overlapped", [True, False])
def test_streamed_host_matrix_equals_resident_scoring(cuda_device, overlapped, monkeypatch):
    """The e2e paths -- ONE gated launch that consumes the rows while the copy stream delivers them
    (bb_score_fused_overlapped), or row blocks scored launch by launch (bb_score_fused_host) -- must be
    bit-identical to the resident one."""
    monkeypatch.setattr(DeviceGP, "OVERLAPPED_HOST_PASS", overlapped)
    w = numeric_grid_workload(N=600_001, d=20, n=256, seed=3)  # ragged last block
    gp = DeviceGP(device=cuda_device, **w.gp_kwargs())
    z = sobol_normal_samples(512, 1, seed=5)
    acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
    x_host = torch.from_numpy(w.candidates).to(torch.float32).pin_memory()
    keep = torch.ones(len(x_host), dtype=torch.uint8, device=cuda_device)
    keep[::7] = 0
    s_res, k_res = gp.score(acq, x_host.to(cuda_device), z[:, 0], keep=keep, index_offset=1000)
    s_str, k_str = gp.score(acq, x_host, z[:, 0], keep=keep, index_offset=1000)
    assert torch.equal(s_res, s_str) and int(k_res.item()) == int(k_str.item())
    gp.check_host_pass()
    s64, k64 = gp.score(acq, torch.from_numpy(w.candidates), z[:, 0], keep=keep, index_offset=1000)  # pageable fp64
    assert torch.equal(s_res, s64) and int(k_res.item()) == int(k64.item())
    s32, k32 = gp.score(acq, torch.from_numpy(w.candidates).to(torch.float32), z[:, 0], keep=keep, index_offset=1000)  # pageable fp32
    assert torch.equal(s_res, s32) and int(k_res.item()) == int(k32.item())
    gp.check_host_pass()


@pytest.mark.parametrize("overlapped", [True, False])
def test_level_coded_candidates_score_bit_identically(cuda_device, overlapped, monkeypatch):
    """The compact exact form of a discrete space (4- or 8-bit level codes + value table) from pinned host codes --
    expanded inside the gated single-launch pass (bb_score_fused_overlapped) or block by block (bb_decode_codes +
    bb_score_fused_host) -- and from device-resident codes."""
    from baybe_b200.bits import decode_levels, encode_levels

    monkeypatch.setattr(DeviceGP, "OVERLAPPED_HOST_PASS", overlapped)

    for levels, bits_expected in ((11, 4), (40, 8)):
        w = numeric_grid_workload(N=300_017, d=20, n=256, seed=4, levels=levels)  # ragged last block
        gp = DeviceGP(device=cuda_device, **w.gp_kwargs())
        z = sobol_normal_samples(512, 1, seed=5)
        acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
        codes, table, bits = encode_levels(w.candidates)
        assert bits == bits_expected and codes.shape == (len(w.candidates), 10 if bits == 4 else 20)
        assert np.array_equal(decode_levels(codes, table, bits, 20), w.candidates.astype(np.float32))
        x32 = torch.from_numpy(w.candidates).to(cuda_device, torch.float32)
        keep = torch.ones(len(x32), dtype=torch.uint8, device=cuda_device)
        keep[::5] = 0
        s_res, k_res = gp.score(acq, x32, z[:, 0], keep=keep, index_offset=77)
        ch = torch.from_numpy(codes).pin_memory()
        s_host, k_host = gp.score_coded(acq, ch, table, bits, z[:, 0], keep=keep, index_offset=77)
        assert torch.equal(s_res, s_host) and int(k_res.item()) == int(k_host.item())
        s_dev, k_dev = gp.score_coded(acq, ch.to(cuda_device), table, bits, z[:, 0], keep=keep, index_offset=77)
        assert torch.equal(s_res, s_dev) and int(k_res.item()) == int(k_dev.item())
        _, k2 = gp.score_coded(acq, ch, table, bits, z[:, 0], keep=keep, index_offset=77, want_scores=False)  # buffers reused
        assert int(k2.item()) == int(k_res.item())
        gp.check_host_pass()
    with pytest.raises(ValueError):
        gp.score_coded(acq, ch[:, :5].contiguous(), table, bits, z[:, 0])


def test_overlapped_pass_small_ragged_and_masked(cuda_device):
    """bb_score_fused_overlapped on sets smaller than its first copy block, with a ragged last tile, several
    acquisition kinds and the posterior-free key only."""
    from baybe_b200.bits import encode_levels

    for N in (1, 130, 777, 20_001):
        w = numeric_grid_workload(N=N, d=20, n=100, seed=N)
        gp = DeviceGP(device=cuda_device, **w.gp_kwargs())
        z = sobol_normal_samples(256, 1, seed=9)
        for kind in ("qLogEI", "qEI", "UCB"):
            acq = AcqConfig(kind=kind, best_f=gp.best_f(AcqConfig(kind=kind)))
            x32 = torch.from_numpy(w.candidates).to(torch.float32)
            s_res, k_res = gp.score(acq, x32.to(cuda_device), z[:, 0], index_offset=5)
            s_h, k_h = gp.score(acq, x32.pin_memory(), z[:, 0], index_offset=5)
            assert torch.equal(s_res, s_h) and int(k_res.item()) == int(k_h.item())
            codes, table, bits = encode_levels(w.candidates)
            s_c, k_c = gp.score_coded(acq, torch.from_numpy(codes).pin_memory(), table, bits, z[:, 0], index_offset=5)
            assert torch.equal(s_res, s_c) and int(k_res.item()) == int(k_c.item())
            gp.check_host_pass()


def test_config5_four_tasks_one_million_candidates(cuda_device):
    w = task_workload(N_per_task=250_000, n_tasks=4, d_num=20, n_per_task=64, seed=0)
    _check_fullsize(w, cuda_device, n_shards=4)


def test_config4_shard_bit_packed_fingerprints(cuda_device):
    """One GPU's shard of BASELINE config 4 (10M x 2048-bit fingerprints over 8 GPUs = 1.25M rows here,
    n = 512, ScaleKernel(RBF), qLogEI) through the wide-feature path: oracle spot-check on 1024 random
    rows, arg-max consistency, shard/offset invariance, determinism, duplicated rows score identically."""
    from baybe_b200.bits import unpack_bits
    from baybe_b200.synthetic import fingerprint_workload

    dev = cuda_device
    N = 1_250_000
    w = fingerprint_workload(N=4096, d=2048, n=512, seed=1)
    om = oracle_model(w)
    gp = DeviceGP(device=dev, **w.gp_kwargs())
    assert gp.model.wide == 1
    g = torch.Generator(device="cuda").manual_seed(0)
    bits = torch.rand((N, 256, 8), device=dev, generator=g) < 0.05
    packed = (bits.to(torch.uint8) << torch.arange(8, device=dev, dtype=torch.uint8)).sum(dim=2).to(torch.uint8)
    del bits
    packed[N - 1] = packed[17]  # duplicates far apart (different blocks of the K* workspace)
    packed[N // 2] = packed[17]
    z = sobol_normal_samples(512, 1, seed=1234)
    oacq = oracle.AcqSpec("qLogEI")
    oacq.best_f = oracle.best_f_from_training(om, w.train_x, oacq)
    acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
    scores, key = gp.score(acq, packed, z[:, 0])
    val, idx = decode_best(key)
    assert idx == int(torch.argmax(scores).item()) and val == float(scores[idx].item())
    assert torch.isfinite(scores).all()
    assert scores[17] == scores[N - 1] == scores[N // 2]
    rows = np.random.default_rng(0).choice(N, size=1024, replace=False)
    sub = unpack_bits(packed[torch.from_numpy(rows).to(dev)].cpu().numpy(), 2048)
    ref, bound = score_bounds(om, oacq, sub, z[:, 0])
    got = scores[torch.from_numpy(rows).to(dev)].double().cpu()
    err = (got - ref).abs()
    assert bool((err <= bound).all()), (float(err.max()), int(torch.argmax(err - bound)))
    mu, var = gp.posterior(packed[torch.from_numpy(rows).to(dev)])
    mu_ref, var_ref = oracle.posterior(om, sub)
    assert float((mu.double().cpu() - mu_ref).abs().max()) <= 5e-5 * max(1.0, float(mu_ref.abs().max()))
    assert float((var.double().cpu() - var_ref).abs().max()) <= 2e-5 * om.y_std**2
    # shards with global offsets reduce to the same key; second pass is bit-identical
    bounds = [0, 400_000, 900_001, N]
    keys = [gp.score(acq, packed[a:b], z[:, 0], index_offset=a, want_scores=False)[1] for a, b in zip(bounds, bounds[1:])]
    assert int(torch.stack(keys).max()) == int(key)
    scores2, key2 = gp.score(acq, packed, z[:, 0])
    assert torch.equal(scores, scores2) and int(key2) == int(key)
