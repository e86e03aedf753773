#!/usr/bin/env python
"""Benchmark of the SuDoRM-RF forward path (BASELINE.json metric: mixtures/sec,
4 s @ 8 kHz, 2 sources).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's kernels)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm (CPU oracle port)

One "step" = one forward pass of the separator over one synthetic batch.  At
N > 1 the batch is sharded by giving every rank its own batch of the same size
(weak scaling, no collective inside the step; the packed weights are broadcast
ONCE with NCCL before timing).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

# BASELINE.json configs (SURVEY.md §8 shapes).  The headline / default workload is
# configs[1]: Improved SuDoRM-RF U16/512, batch 32 x 4 s @ 8 kHz on one B200.
WORKLOADS = {
    "improved_u16_512": dict(variant="improved", B=32, T=32000, kw=dict(
        out_channels=256, in_channels=512, num_blocks=16, upsampling_depth=5,
        enc_kernel_size=21, enc_num_basis=512, num_sources=2)),
    "improved_u8_512": dict(variant="improved", B=1, T=32000, kw=dict(
        out_channels=256, in_channels=512, num_blocks=8, upsampling_depth=5,
        enc_kernel_size=21, enc_num_basis=512, num_sources=2)),
    "improved_u36_2048": dict(variant="improved", B=64, T=32000, kw=dict(
        out_channels=512, in_channels=512, num_blocks=36, upsampling_depth=6,
        enc_kernel_size=21, enc_num_basis=2048, num_sources=2)),
    "groupcomm_u8_512": dict(variant="groupcomm", B=16, T=32000, kw=dict(
        out_channels=256, in_channels=512, num_blocks=8, upsampling_depth=5,
        enc_kernel_size=21, enc_num_basis=512, num_sources=2, group_size=16)),
    "improved_u36_4096_16k": dict(variant="improved", B=32, T=64000, kw=dict(
        out_channels=512, in_channels=512, num_blocks=36, upsampling_depth=6,
        enc_kernel_size=21, enc_num_basis=4096, num_sources=2)),
}
METRIC = "mixtures_per_sec_forward_4s_8kHz_2src"
UNIT = "mixtures/s"


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------
# algorithmic bytes / flops model, fp32, every intermediate
# written once and read once
# --------------------------------------------------------------------------
def algorithmic_model(w):
    kw = w["kw"]
    N, Co, Ci = kw["enc_num_basis"], kw["out_channels"], kw["in_channels"]
    U, D, S, K = kw["num_blocks"], kw["upsampling_depth"], kw["num_sources"], kw["enc_kernel_size"]
    hop = K // 2
    q = hop * 2 ** D
    Tp = max(q, -(-w["T"] // q) * q)
    L = Tp // hop
    kappa = 5 + sum(2.0 ** -(d - 1) + 2.0 ** -d for d in range(1, D)) + sum(2.0 ** -d for d in range(D))
    a_blk = 4 * L * (3 * Co + kappa * Ci)
    gc = w["variant"] == "groupcomm"
    if gc:
        a_blk += 4 * L * 4 * Co
    a_front = 4 * (Tp + 2 * N * L + Co * L)
    a_back = 4 * (Co * L + N * L + S * Tp)
    a_mix = a_front + U * a_blk + a_back
    G = kw.get("group_size", 1) if gc else 1
    gemm = 2 * Co * Ci / G
    flops = 2 * L * (K * N + N * Co + U * (gemm + 5 * Ci * (2 - 2.0 ** (1 - D))) + Co * S * N + K * S * S * N)
    if gc:
        n = Co // G
        H = 3 * n
        flops += U * 2 * L * (3 * G * n * H + H * H)
    return dict(L=L, Tp=Tp, a_blk=a_blk, a_mix=a_mix, flops=flops,
                res_bytes=4 * L * (Ci + 2 * Co), res_flops=2 * Co * Ci * L / G)


# --------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md)
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [s.strip() for s in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle port of the reference forward
# --------------------------------------------------------------------------
def cpu_forward_rate(w, sample_B, steps, warmup, budget_s=None):
    """mixtures/s of the CPU oracle port (same torch op sequence as the reference's
    forward, improved_sudormrf.py:283-301) on all host cores."""
    from oracle import sudormrf_oracle as O
    cfg = O.Config(variant=w["variant"], **w["kw"])
    sd = O.make_state_dict(cfg, seed=0, perturbed=False)
    x = torch.rand(sample_B, 1, w["T"], generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        for _ in range(warmup):
            O.forward(cfg, sd, x)
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            O.forward(cfg, sd, x)
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return sample_B * done / dt, dt / done, done


def run_reference(args, w, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = 2
    cores = torch.get_num_threads()
    rate, sec_per_step, done = cpu_forward_rate(w, sample_B, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (torch.rand mixtures, default-init weights)",
        "config": {"workload": wl_name, "batch_per_step": sample_B, "samples": w["T"],
                   "note": "CPU oracle port of the reference forward (the Python reference cannot "
                           "travel to the GPU box); bounded sample of the workload per step"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{done} forwards of batch {sample_B} x {w['T']} samples"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------
def run_b200(args, w, wl_name):
    import ctypes as C
    import torch.distributed as dist
    import sudo_rm_rf_b200 as P
    from sudo_rm_rf_b200 import _engine, _native
    from oracle import sudormrf_oracle as O      # weights generator + cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (B200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, T = (args.batch or w["B"]), w["T"]
    cls = P.SuDORMRF if w["variant"] == "improved" else P.GroupCommSudoRmRf
    model = cls(**w["kw"])
    cfg_o = O.Config(variant=w["variant"], **w["kw"])
    if rank == 0:
        model.load_state_dict(O.make_state_dict(cfg_o, seed=0, perturbed=False))
    model = model.to(dev).eval()
    if world > 1:
        # ONE broadcast of all weights over NVLink (NCCL), outside the timed step; the step itself
        # has no collective: every rank separates its own shard of the batch
        from sudo_rm_rf_b200 import sharding
        sharding.broadcast_parameters(model, src=0)
    cfg = _engine.make_config(model)
    launches_per_step = _native.lib().sdr_forward_launch_count(C.byref(cfg))

    g = torch.Generator().manual_seed(1234 + rank)
    host_x = torch.rand(B, 1, T, generator=g).pin_memory()      # the reference's bench input (notebook :128)
    x = host_x.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    stream = torch.cuda.Stream(device=dev)
    with torch.no_grad(), torch.cuda.stream(stream):
        y = model(x)                      # packs weights, sizes the workspace
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            y = model(x)
        for _ in range(max(args.warmup, 3)):
            graph.replay()
        stream.synchronize()

        def barrier():
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        # ---- device-resident timing: K graph replays, L2 flushed between steps ----
        sampler = ClockSampler(local) if rank == 0 else None
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
        barrier()
        if sampler:
            sampler.start()
        wall0 = time.perf_counter()
        for s, e in evs:
            flush.fill_(1)
            s.record(stream)
            graph.replay()
            e.record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if sampler else None
        dev_ms = sum(s.elapsed_time(e) for s, e in evs)

        # ---- end to end: pinned host mixtures in, pinned host estimates out, every step ----
        host_y = torch.empty(B, w["kw"]["num_sources"], T).pin_memory()
        for _ in range(3):
            model.forward_host(host_x, host_y)
        stream.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            model.forward_host(host_x, host_y)
            stream.synchronize()           # the step's result is read on the host
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)

        # ---- dominant kernel, timed alone with CUDA events on the launching stream ----
        roof = None
        if rank == 0:
            roof = time_dominant_kernel(w, B, stream, flush, dev)

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()

    if rank == 0:
        peak, peak_src = load_peaks()
        am = algorithmic_model(w)
        n_params = sum(p.numel() for p in model.parameters())
        total_mix = B * world * args.steps
        value = total_mix / (dev_ms / 1e3)
        fwd_bytes = B * am["a_mix"] + 4 * n_params
        fwd_gbs = fwd_bytes / (dev_ms / args.steps / 1e3) / 1e9
        if world == 1:      # the CPU leg is timed at N=1 only (torchrun pins OMP threads to 1 per rank)
            cpu_rate, cpu_sec, cpu_done = cpu_forward_rate(w, 2, steps=64, warmup=1, budget_s=12.0)
            cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                            "sample": f"{cpu_done} forwards of batch 2 x {T} samples ({cpu_sec:.2f} s each), same model"}
        else:
            cpu_baseline = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (torch.rand mixtures, default-init weights)",
            "config": {"workload": wl_name, "batch_per_gpu": B, "global_batch": B * world,
                       "samples": T, "sources": w["kw"]["num_sources"], "parallelism": f"dp{world}",
                       "l2": "256 MiB flush write between timed steps; one step streams "
                             f"{fwd_bytes / 1e9:.1f} GB (algorithmic) through a >1 GB workspace >> 126 MB L2",
                       "timing": "CUDA events around each CUDA-graph replay, summed over K steps, max over ranks",
                       "wall_s_timed_loop": wall},
            "e2e": {"value": total_mix / (e2e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": B * T * 4 * world,
                    "d2h_bytes_per_step": B * w["kw"]["num_sources"] * T * 4 * world,
                    "api": "model.forward_host (sdr_forward_host): pinned host in/out, sync per step"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": roof,
            "forward_hbm": {"algorithmic_bytes_per_step": fwd_bytes, "achieved": fwd_gbs, "peak": peak,
                            "unit": "GB/s", "frac": fwd_gbs / peak, "peak_source": peak_src,
                            "gflop_per_mixture": am["flops"] / 1e9},
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def time_dominant_kernel(w, B, stream, flush, dev):
    """The dominant kernel of the step is the res_conv pointwise GEMM (+skip) of the
    U-ConvBlocks: time it alone at the benchmark shape with CUDA events."""
    import ctypes as C
    from sudo_rm_rf_b200 import _native as N
    kw = w["kw"]
    gc = w["variant"] == "groupcomm"
    G = kw.get("group_size", 1) if gc else 1
    am = algorithmic_model(w)
    L = am["L"]
    samples, M, K = B * G, kw["out_channels"] // G, kw["in_channels"] // G
    gen = torch.Generator(device=dev).manual_seed(0)
    xin = torch.randn(samples, K, L, device=dev, generator=gen)
    Wt = torch.randn(M, K, device=dev, generator=gen) / K ** 0.5
    bias = torch.randn(M, device=dev, generator=gen)
    gamma = torch.ones(K, device=dev)
    beta = torch.zeros(K, device=dev)
    slope = torch.full((1,), 0.25, device=dev)
    xd = xin.double().reshape(samples, -1)
    stats = torch.stack([xd.sum(1), (xd * xd).sum(1)], 1).contiguous()
    res = torch.randn(samples, M, L, device=dev, generator=gen)
    nin = N.SdrNormIn(stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), slope.data_ptr(), float(K * L))
    sp = C.c_void_p(stream.cuda_stream)

    nbytes = N.lib().sdr_pointwise_mma_packed_bytes(M, K)
    wpk = None
    if nbytes:
        wpk = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        N.check(N.lib().sdr_pointwise_mma_pack(C.c_void_p(Wt.data_ptr()), M, K, C.c_void_p(wpk.data_ptr()), sp))

    def launch():
        if wpk is not None:       # the path the forward takes for this shape (tcgen05)
            N.check(N.lib().sdr_pointwise_mma(
                C.c_void_p(xin.data_ptr()), C.byref(nin), C.c_void_p(wpk.data_ptr()),
                C.c_void_p(bias.data_ptr()), C.c_void_p(res.data_ptr()), C.c_void_p(0), 0,
                C.c_void_p(res.data_ptr()), C.c_void_p(0), samples, M, K, L, 0, sp))
        else:
            N.check(N.lib().sdr_pointwise(
                C.c_void_p(xin.data_ptr()), C.byref(nin), C.c_void_p(Wt.data_ptr()),
                C.c_void_p(bias.data_ptr()), C.c_void_p(res.data_ptr()), C.c_void_p(0), 0,
                C.c_void_p(res.data_ptr()), C.c_void_p(0), samples, M, K, L, 0, sp))

    for _ in range(3):
        launch()
    reps = 10
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in evs:
        flush.fill_(2)
        s.record(stream)
        launch()
        e.record(stream)
    stream.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in evs)
    avg = sum(ms) / len(ms)
    peak, peak_src = load_peaks()
    traffic = None      # dram__bytes_read+write per launch of this kernel at this shape, from the committed ncu capture
    tpath = os.path.join(REPO, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("shape") == {"samples": samples, "M": M, "K": K, "L": L}:
            traffic = t.get("dram_bytes_per_launch")
    bytes_per_launch = B * am["res_bytes"]
    achieved = bytes_per_launch / (avg / 1e3) / 1e9
    return {"kernel": "pointwise GEMM res_conv+skip (" + ("pw_mma_kernel, tcgen05 bf16x3" if wpk is not None else "pw_gemm_kernel, FFMA") + ")",
            "bound": "hbm",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg,
            "tflops_fp32_equivalent": B * am["res_flops"] / (avg / 1e3) / 1e12,
            "shape": {"samples": samples, "M": M, "K": K, "L": L}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="improved_u16_512", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_b200(args, w, args.workload)


if __name__ == "__main__":
    main()
