"""world_size-2 gloo test (CPU) of the sharded arg-max protocol used with N>1 GPUs: every rank
scores a contiguous row shard, packs (score, global index) into one int64 key, a MAX all-reduce
picks the winner, ties go to the lowest global index; sequential greedy rounds mask the winner on
its owning rank only.  Scores come from the CPU oracle here -- the protocol is what is tested."""
from __future__ import annotations

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from baybe_b200.engine import pack_best, unpack_best
from baybe_b200.recommenders import shard_bounds


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, scores: np.ndarray, q: int, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(len(scores), rank, world)
    local = scores[lo:hi].copy()
    keep = np.ones(hi - lo, dtype=bool)
    chosen = []
    for _ in range(q):
        best = -(1 << 63)
        for i in np.nonzero(keep)[0]:
            if not np.isnan(local[i]):
                best = max(best, pack_best(float(local[i]), lo + int(i)))
        key = torch.tensor([best], dtype=torch.int64)
        dist.all_reduce(key, op=dist.ReduceOp.MAX)
        val, idx = unpack_best(int(key.item()))
        chosen.append(idx)
        if lo <= idx < hi:
            keep[idx - lo] = False
    if rank == 0:
        out.put(chosen)
    dist.destroy_process_group()


def test_sharded_argmax_matches_single_process_greedy():
    rng = np.random.default_rng(0)
    scores = rng.standard_normal(1001).astype(np.float32)
    scores[[5, 700]] = scores.max() + 1.0  # tie across shards -> lowest global index first
    scores[17] = np.nan
    q = 4
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, scores, q, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = out.get()
    ref = []
    s = torch.from_numpy(np.nan_to_num(scores, nan=-np.inf)).clone()
    for _ in range(q):
        j = int(torch.argmax(s))  # first maximum
        ref.append(j)
        s[j] = -float("inf")
    assert got == ref and got[:2] == [5, 700]


def _topk_worker(rank: int, world: int, port: int, scores: np.ndarray, k: int, out):
    from baybe_b200.recommenders import merge_topk_across_ranks

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(len(scores), rank, world)
    local = torch.from_numpy(scores[lo:hi])
    v, i = torch.topk(local, min(k, hi - lo))  # stands in for the per-rank bb_topk (a GPU op)
    pad = k - v.numel()
    v = torch.cat([v, torch.full((pad,), -float("inf"))])
    i = torch.cat([i + lo, torch.full((pad,), -1, dtype=torch.int64)])
    gv, gi = merge_topk_across_ranks(v, i, k)
    if rank == 1:  # every rank holds the same merged result; check the non-zero rank
        out.put((gv.tolist(), gi.tolist()))
    dist.destroy_process_group()


def test_topk_merge_across_ranks():
    """SURVEY.md 8e / BASELINE config 5: all-gather of per-rank top-k, ties to the lowest global index."""
    rng = np.random.default_rng(1)
    scores = rng.standard_normal(777).astype(np.float32)
    scores[[3, 500]] = scores.max() + 2.0  # tie across the two shards
    k = 6
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_topk_worker, args=(r, 2, port, scores, k, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gv, gi = out.get()
    order = np.lexsort((np.arange(len(scores)), -scores))[:k]  # value descending, index ascending
    assert gi == order.tolist() and gi[:2] == [3, 500]
    assert np.allclose(gv, scores[order])


def test_topk_merge_single_process_and_padding():
    from baybe_b200.recommenders import merge_topk_across_ranks

    v = torch.tensor([2.0, 2.0, -float("inf")])
    i = torch.tensor([9, 4, -1])
    gv, gi = merge_topk_across_ranks(v, i, 3)
    assert gi.tolist() == [4, 9, -1] and gv[:2].tolist() == [2.0, 2.0]


def _seed_worker(rank: int, world: int, port: int, out):
    from baybe_b200.recommenders import _broadcast_seed

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)  # the usual seed+rank pattern: the ranks' own draws differ
    own = int(torch.randint(0, 1_000_000, (1,)).item())
    got = _broadcast_seed(own, torch.device("cpu"))
    out.put((rank, own, got))
    dist.destroy_process_group()


def test_sampler_seed_is_broadcast_from_rank_zero():
    """every rank must score its shard with the same base samples, whatever its own RNG state."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_seed_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(out.get() for _ in range(2))
    assert res[0][1] != res[1][1]  # different local draws ...
    assert res[0][2] == res[1][2] == res[0][1]  # ... one shared seed: rank 0's


def _greedy_worker(rank: int, world: int, port: int, out):
    """greedy_select across two ranks with the oracle-backed engine: shard-local scoring, global arg-max, winner rows
    fetched from the owning rank, pending points growing round by round."""
    import oracle
    from baybe_b200.engine import AcqConfig
    from baybe_b200.recommenders import greedy_select, shard_bounds as sb
    from baybe_b200.synthetic import numeric_grid_workload
    from tests.helpers import OracleBackedGP, oracle_model

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = numeric_grid_workload(N=301, d=4, n=20, seed=3)
    gp = OracleBackedGP(**w.gp_kwargs())
    cfg = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
    lo, hi = sb(len(w.candidates), rank, world)
    x = torch.from_numpy(w.candidates[lo:hi]).to(torch.float32)
    chosen, vals = greedy_select(gp, cfg, x, 4, 3, None, seed=5, n_samples=64, offset=lo)
    if rank == 1:
        om = oracle_model(w)
        oacq = oracle.AcqSpec("qLogEI", best_f=cfg.best_f)
        ref, _ = oracle.optimize_acqf_discrete(om, oacq, w.candidates.astype(np.float32).astype(np.float64), q=3,
                                               sampler_seed=5, n_samples=64)
        out.put((chosen, ref))
    dist.destroy_process_group()


def test_sharded_greedy_selection_matches_single_process_oracle():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_greedy_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    chosen, ref = out.get()
    assert chosen == ref


def _escape_hatch_worker(rank: int, world: int, port: int, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from baybe_b200.peers import NcclReduce

    red = NcclReduce(torch.device("cpu"))  # the class only needs a process group; gloo + CPU tensors here
    got = []
    for epoch, (a, b) in enumerate([((0.5, 10), (0.5, 3)), ((-1.0, 7), (2.0, 900_000)), (None, (0.25, 5))]):
        mine = (a, b)[rank]
        key = -(1 << 63) if mine is None else pack_best(*mine)
        got.append(unpack_best(int(red.allreduce_best(torch.tensor([key], dtype=torch.int64)).item())))
    red.check()
    if rank == 0:
        out.put(got)
    dist.destroy_process_group()


def test_nccl_escape_hatch_reduces_like_the_peer_kernel():
    """BB_PEER_REDUCE=0 path (baybe_b200/peers.py::NcclReduce): same contract as bb_allreduce_best -- maximum of
    the packed keys, ties to the lowest index, empty keys ignored."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_escape_hatch_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() == [(0.5, 3), (2.0, 900_000), (0.25, 5)]
