#!/usr/bin/env python
"""GPU box: stand-alone HJI lookup micro-benchmark: 2^24 random in-grid queries on the 13x13x9^5 grid (319 MB), input order
against cell order.  `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum python tools/gpu_hji_bench.py 22` gives the DRAM
bytes per kernel (argument: log2 of the query count)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pigeon.jl_b200 as p
from pigeon.jl_b200 import synthetic
M = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
dev = torch.device("cuda", 0)
knots, V, gV = synthetic.analytic_hji_grid()
m = p.BatchedCoupledTrajectoryTrackingMPC(p.X1(), [p.straight_trajectory(30.0, 5.0)], 1)
m.set_HJI_cache(p.HJICache(knots, V, gV))
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); m.set_stream(stream.cuda_stream)      # the events must sit on the stream the library launches on
g = torch.Generator(device=dev); g.manual_seed(7)
lo = torch.tensor([r[0] for r in synthetic.HJI_RANGES], dtype=torch.float64, device=dev)[:, None]
hi = torch.tensor([r[1] for r in synthetic.HJI_RANGES], dtype=torch.float64, device=dev)[:, None]
x = (lo + (hi - lo) * (0.001 + 0.998 * torch.rand((7, M), dtype=torch.float64, device=dev, generator=g))).contiguous()
Vo = torch.empty(M, dtype=torch.float64, device=dev); go = torch.empty((7, M), dtype=torch.float64, device=dev)
res = {}
for mode, name in ((0, "input order"), (1, "cell order"), (2, "cell order + TMA tiles")):
    m.set_hji_lookup_order(mode)
    for _ in range(2):
        m.hji_lookup_device(M, x.data_ptr(), Vo.data_ptr(), go.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record(stream); m.hji_lookup_device(M, x.data_ptr(), Vo.data_ptr(), go.data_ptr()); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[name] = (best, Vo.clone(), go.clone())
    print(f"{name}: {M} queries in {best:.3f} ms = {M / best / 1e3:.1f} M queries/s = {M * 4096 / best / 1e6:.1f} GB/s algorithmic")
print("bit-identical:", all(bool(torch.equal(res["input order"][1], res[k][1]) and torch.equal(res["input order"][2], res[k][2])) for k in res))
m.close()
