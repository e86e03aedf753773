"""Diagnostic: per-role timeline of CTA 0 of the tcgen05 GEMM (needs the SDR_MMA_TRACE build:
SDR_B200_LIB=sudo_rm_rf_b200/csrc/build/lib_tr.so)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sudo_rm_rf_b200 import _native as N

lib = N.lib()
raw = C.CDLL(N.LIB_PATH)
dev = torch.device("cuda")
P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, M, K, res in (("proj", 512, 256, False), ("res_conv", 256, 512, True)):
    S, L = 32, 3200
    x = torch.randn(S, K, L, device=dev)
    W = torch.randn(M, K, device=dev) / K ** 0.5
    bias = torch.randn(M, device=dev)
    y = torch.randn(S, M, L, device=dev)
    xd = x.double().reshape(S, -1)
    st = torch.stack([xd.sum(1), (xd * xd).sum(1)], 1).contiguous()
    g, b, sl = torch.ones(K, device=dev), torch.zeros(K, device=dev), torch.full((1,), 0.25, device=dev)
    nin = N.SdrNormIn(st.data_ptr(), g.data_ptr(), b.data_ptr(), sl.data_ptr(), float(K * L)) if res else N.SdrNormIn(0, 0, 0, 0, 1.0)
    wpk = torch.empty(lib.sdr_pointwise_mma_packed_bytes(M, K), dtype=torch.uint8, device=dev)
    N.check(lib.sdr_pointwise_mma_pack(P(W), M, K, P(wpk), sp))
    sto = torch.zeros(S, 2, dtype=torch.float64, device=dev)
    buf = (C.c_ulonglong * (4 * 256 * 8))()
    for rep in range(3):
        torch.cuda.synchronize()
        raw.sdr_debug_read_trace(buf, 1)
        N.check(lib.sdr_pointwise_mma(P(x), C.byref(nin), P(wpk), P(bias), P(y) if res else P(None), P(None), 0, P(y),
                                      P(None) if res else P(sto), S, M, K, L, 0, sp))
        torch.cuda.synchronize()
    raw.sdr_debug_read_trace(buf, 0)
    t = np.array(buf, dtype=np.uint64).reshape(4, 256, 8).astype(np.int64)
    for k in range(2):
        e0, e1 = t[3, 64 + k, 0], t[3, 64 + k, 1]
        print(f"== {name} tile {2 + k}: epilogue {(e1 - e0) / 1000:.2f} us; per 16-column chunk: [ld+wait us, process us]")
        prev = e0
        row = []
        for c in range(16):
            a0, a1 = t[3, k * 32 + c, 0], t[3, k * 32 + c, 1]
            nxt = t[3, k * 32 + c + 1, 0] if c < 15 else e1
            row.append(f"[{(a1 - a0) / 1000:.2f} {(nxt - a1) / 1000:.2f}]")
        print("   " + " ".join(row))
