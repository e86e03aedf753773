"""Diagnostic: per-role clock64 timeline of pair 0's leader CTA of the tcgen05 GEMM.
Needs the trace build:  tools/build_variants.sh "trace=-DSDR_MMA_TRACE=1"  then
SDR_B200_LIB=$PWD/variants/trace.so python tools/trace_gemm.py"""
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
GHZ = float(os.environ.get("SM_GHZ", "1.9"))
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
    buf = torch.zeros(6 * 256 * 4, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    raw.sdr_debug_set_trace(P(buf))
    for rep in range(3):
        flush.fill_(1)
        buf.zero_()
        torch.cuda.synchronize()
        N.check(lib.sdr_pointwise_mma(P(x), C.byref(nin), P(wpk), P(bias), P(y) if res else P(None), P(None), 0, P(y),
                                      P(None) if res else P(sto), S, M, K, L, 0, sp))
        torch.cuda.synchronize()
    raw.sdr_debug_set_trace(C.c_void_p(0))
    t = buf.cpu().numpy().reshape(6, 256, 4).astype(np.int64)
    t0 = t[4, 0, 0]
    us = lambda c: (c - t0) / (GHZ * 1e3)
    KB = K // 64
    print(f"==== {name}: M={M} K={K} ({KB} k-blocks / tile); times in us since the MMA thread's first wait (clock {GHZ} GHz)")
    ntile = 6
    print("-- MMA thread per tile: [wait tempty from..to] then per k-block (wait_full_from, got_full, issued)")
    for ti in range(ntile):
        row = " ".join(f"({us(t[3, ti * KB + k, 0]):.2f} {us(t[3, ti * KB + k, 1]):.2f} {us(t[3, ti * KB + k, 2]):.2f})" for k in range(KB))
        print(f"  tile {ti}: tempty [{us(t[4, ti, 0]):.2f}..{us(t[4, ti, 1]):.2f}]  {row}")
    print("-- transform warp 0 per k-block: (wait_raw_from, raw_landed, a_slot_free, published)")
    for ti in range(ntile):
        print("  tile %d: " % ti + " ".join(
            f"({us(t[0, ti * KB + k, 0]):.2f} {us(t[0, ti * KB + k, 1]):.2f} {us(t[0, ti * KB + k, 2]):.2f} {us(t[0, ti * KB + k, 3]):.2f})"
            for k in range(KB)))
    print("-- raw loader per k-block: (wait_slot_from, slot_free)")
    for ti in range(ntile):
        print("  tile %d: " % ti + " ".join(f"({us(t[1, ti * KB + k, 0]):.2f} {us(t[1, ti * KB + k, 1]):.2f})" for k in range(KB)))
    print("-- weight loader per k-block: (wait_slot_from, slot_free)")
    for ti in range(ntile):
        print("  tile %d: " % ti + " ".join(f"({us(t[2, ti * KB + k, 0]):.2f} {us(t[2, ti * KB + k, 1]):.2f})" for k in range(KB)))
    print("-- epilogue warp 0 per tile: (wait_tfull_from, accumulator_ready, drained)")
    print("  " + " ".join(f"({us(t[5, ti, 0]):.2f} {us(t[5, ti, 1]):.2f} {us(t[5, ti, 2]):.2f})" for ti in range(ntile + 2)))
    # summary over the steady state
    nt = min(10, 255 // KB)
    per_tile = (t[3, (nt - 1) * KB, 0] - t[3, KB, 0]) / (nt - 2) / (GHZ * 1e3)
    wait_full = sum(t[3, i, 1] - t[3, i, 0] for i in range(KB, (nt - 1) * KB)) / (nt - 2) / (GHZ * 1e3)
    wait_tempty = sum(t[4, i, 1] - t[4, i, 0] for i in range(1, nt - 1)) / (nt - 2) / (GHZ * 1e3)
    print(f"   steady state per tile: {per_tile:.2f} us; MMA thread waits on operands {wait_full:.2f} us, on the accumulator {wait_tempty:.2f} us")
