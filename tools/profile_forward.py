"""Small driver for ncu: N eager forwards of a benchmark workload (no CUDA graph,
so every kernel is a plain launch)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import sudo_rm_rf_b200 as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="improved_u16_512")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
m = bench.model_class(w["variant"])(**w["kw"]).cuda().eval()
x = torch.rand(a.batch or w["B"], 1, w["T"], device="cuda")
with torch.no_grad():
    for _ in range(a.iters):
        y = m(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
