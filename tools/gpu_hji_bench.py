#!/usr/bin/env python
"""GPU box: stand-alone HJI value/gradient lookup micro-benchmark: the 13x13x9^5 float32 grid (319 MB, one
32-byte record per node) resident in HBM, M uniformly random in-grid queries.  Algorithmic traffic = 128 corners x 32 B = 4096 B per
query; prints one JSON line with queries/s, achieved GB/s and the fraction of the measured HBM copy bandwidth."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pigeon.jl_b200 as p
from pigeon.jl_b200 import synthetic

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
trajs = synthetic.synthetic_trajectories(n_traj=1, n_nodes=50)
g = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), trajs, 8)
knots, V, gV = synthetic.analytic_hji_grid()
g.set_HJI_cache(p.HJICache(knots, V, gV))
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
g.set_stream(stream.cuda_stream)
gen = torch.Generator(device=dev); gen.manual_seed(0x5049474E)
lo = torch.tensor([r[0] for r in synthetic.HJI_RANGES], dtype=torch.float64, device=dev)
hi = torch.tensor([r[1] for r in synthetic.HJI_RANGES], dtype=torch.float64, device=dev)
x = (lo[:, None] + (hi - lo)[:, None] * torch.rand((7, M), dtype=torch.float64, device=dev, generator=gen)).contiguous()     # field-major [7][M]
dV = torch.empty(M, dtype=torch.float64, device=dev)
dG = torch.empty((7, M), dtype=torch.float64, device=dev)
for _ in range(2):
    g.hji_lookup_device(M, x.data_ptr(), dV.data_ptr(), dG.data_ptr())
torch.cuda.synchronize()
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    g.hji_lookup_device(M, x.data_ptr(), dV.data_ptr(), dG.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
# spot check against the analytic function (multilinear interpolation of a smooth function on this grid: error << cell size^2)
idx = torch.arange(0, M, max(1, M // 4096), device=dev)
xs = x[:, idx].T.cpu().numpy()
val, grad = synthetic.analytic_hji_value(xs)
err_V = float(np.max(np.abs(dV[idx].cpu().numpy() - val)))
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
peak = peaks.get("hbm_gbs", 6650.0)
best = min(ms)
gbs = M * 4096 / (best * 1e-3) / 1e9
print(json.dumps({"kernel": "k_hji_lookup", "queries": M, "grid_nodes": int(np.prod(synthetic.HJI_DIMS)), "grid_bytes": int(np.prod(synthetic.HJI_DIMS)) * 32,
                  "ms_best": best, "ms_all": ms, "queries_per_s": M / (best * 1e-3), "algorithmic_bytes_per_query": 4096, "achieved_GBps": gbs,
                  "hbm_peak_GBps": peak, "frac_of_hbm_peak": gbs / peak, "io_bytes_per_query": 7 * 8 + 8 * 8,
                  "max_abs_err_V_vs_analytic": err_V}))
g.close()
