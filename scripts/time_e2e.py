"""End-to-end pass timing (host codes / host fp32 rows -> key on the host), per publication mode.
usage: BB_GATE_PUBLISH=<0|1|2> python scripts/time_e2e.py [overlapped 0|1]"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from baybe_b200 import AcqConfig, DeviceGP, sobol_normal_samples
from baybe_b200.bits import encode_levels
from baybe_b200.synthetic import numeric_grid_workload

DeviceGP.OVERLAPPED_HOST_PASS = (sys.argv[1] if len(sys.argv) > 1 else "1") == "1"
dev = torch.device("cuda", 0)
w = numeric_grid_workload(N=1_000_000, d=20, n=256)
gp = DeviceGP(device=dev, **w.gp_kwargs())
z = sobol_normal_samples(512, 1, 1234)[:, 0].to(dev, torch.float32)
acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
codes, table, bits = encode_levels(w.candidates)
ch = torch.from_numpy(codes).pin_memory()
tab = torch.from_numpy(table).to(dev)
x32 = torch.from_numpy(w.candidates).to(torch.float32).pin_memory()
xd = x32.to(dev)
key_host = torch.empty(1, dtype=torch.int64).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(fn, k=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    wall = 0.0
    for _ in range(k):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        wall += time.perf_counter() - t0
        tot += e0.elapsed_time(e1)
    return tot / k, wall / k * 1e3

def fin(key):
    key_host.copy_(key, non_blocking=True)
    torch.cuda.current_stream().synchronize()

mode = os.environ.get("BB_GATE_PUBLISH", "0")
print("overlapped", DeviceGP.OVERLAPPED_HOST_PASS, "publish", mode)
print(" resident      ms (events, wall):", run(lambda: fin(gp.score(acq, xd, z, want_scores=False)[1])))
print(" host codes4   ms (events, wall):", run(lambda: fin(gp.score_coded(acq, ch, tab, bits, z, want_scores=False)[1])))
print(" host fp32     ms (events, wall):", run(lambda: fin(gp.score(acq, x32, z, want_scores=False)[1]), k=5))
# a side stream as the compute stream (not the legacy default stream)
st = torch.cuda.Stream(device=dev)
with torch.cuda.stream(st):
    def fin2(key):
        key_host.copy_(key, non_blocking=True)
        st.synchronize()
    print(" host codes4 on a side stream ms:", run(lambda: fin2(gp.score_coded(acq, ch, tab, bits, z, want_scores=False)[1])))
gp.check_host_pass()
print("E2E_DONE")
