"""Pipeline timeline of CTA 0 (tiles 6..8) of the fused kernel from the test-only event trace."""
import ctypes as C, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from baybe_b200 import AcqConfig, DeviceGP, sobol_normal_samples, _lib
from baybe_b200.synthetic import numeric_grid_workload

mode = sys.argv[1] if len(sys.argv) > 1 else "posterior"
dev = torch.device("cuda", 0)
w = numeric_grid_workload(N=1_000_000, d=20, n=256)
gp = DeviceGP(device=dev, **w.gp_kwargs())
x = torch.from_numpy(w.candidates).to(dev, torch.float32)
z = sobol_normal_samples(512, 1, 1234)[:, 0]
acq = AcqConfig(kind="qLogEI", best_f=gp.best_f(AcqConfig(kind="qLogEI")))
run = (lambda: gp.posterior(x)) if mode == "posterior" else (lambda: gp.score(acq, x, z, want_scores=False))
for _ in range(2): run()
torch.cuda.synchronize()
cap = 4000
buf = torch.zeros(1 + 2 * cap, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.bb_debug_set_trace(C.c_void_p(buf.data_ptr()), cap)
run(); torch.cuda.synchronize()
lib.bb_debug_set_trace(None, 0)
h = buf.cpu()
n = min(int(h[0]), cap)
ev = sorted(((int(h[2 + 2 * i]), int(h[1 + 2 * i])) for i in range(n) if int(h[2 + 2 * i]) != 0))
t0 = ev[0][0]
if len(sys.argv) > 2 and sys.argv[2] == "ts":  # event ids of fused_ts.cu
    names = {100: "C chunk in regs", 110: "C chunk arrived", 120: "C stage_a2 done", 121: "C  a2 stored", 122: "C  a2 quarter-synced", 130: "E vsub ok", 140: "E v_empty arrive", 142: "E  table part done", 143: "E  exact rows done",
             150: "E acquisition done", 160: "C d2 slab0(next) ok", 161: "C d2 slab1 ok", 200: "M v_empty ok",
             210: "M a_full ok", 220: "M V chunk issued", 250: "M DIST issued slab", 260: "M a2_full ok"}
    for clk, code in ev:
        it, e = divmod(code, 1000)
        base = max(k for k in names if k <= e)
        print(f"{clk - t0:8d}  tile {it}  {names[base]:22s} {e - base}")
    sys.exit(0)
names = {100: "C chunk start", 110: "C D2 in regs", 120: "C k+split done", 130: "C mc slice done", 140: "C a_empty ok",
         150: "C a_full arrive", 160: "C finish_prev done", 161: "C stage_a2 done", 170: "C dsub ok", 180: "C epilogue bar",
         200: "M d_empty ok", 210: "M a_full ok", 220: "M r_full ok", 240: "M chunk issued", 250: "M dist begin",
         251: "M a2_full ok", 252: "M d2_empty ok", 253: "M dist issued", 300: "P r_empty ok"}
for clk, code in ev:
    it, e = divmod(code, 1000)
    base = max(k for k in names if k <= e)
    print(f"{clk - t0:8d}  tile {it}  {names[base]:20s} {e - base}")
