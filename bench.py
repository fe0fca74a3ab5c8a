#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 quantized-linear hot path.

Metric (BASELINE.json): NF4 gemv_4bit GB/s, on BASELINE config[1]: Llama-2-7B-shape NF4 gemv, batch 1,
blocksize 64, double-quantised absmax, bf16 activations.

One "step" = one pass of gemv_4bit over a stack of LAYERS distinct 7B decoder layers
(per layer: 4x 4096x4096, 2x 11008x4096, 1x 4096x11008 = 7 GEMVs).  The stack holds
LAYERS * 104.6 MB of packed weights (> the 126 MB L2 for LAYERS >= 2), so every launch streams its
weights from HBM -- "inputs larger than L2", no flush needed.

value      = algorithmic bytes of the whole job / device time (CUDA events, max over ranks); the
             algorithmic bytes per GEMV are N*K/2 + N*K/64 + 4*N*K/16384 + 2K + 2N (+1088 B codebooks),
             SURVEY.md 8d -- never the bytes the kernel chose to move.
e2e        = same metric through the public API (bnb_b200.matmul_4bit, what Linear4bit.forward calls) with
             the step's activations copied from pinned host memory and the outputs copied back, all
             inside the timed region.
roofline   = dominant kernel (k_gemv4_bc<bf16, nested>): algorithmic bytes per launch / average launch
             duration measured with CUDA events on the launch stream over the timed region; peak =
             MEASURED_PEAKS.json hbm_gbs (else the profiling guide's 6650 GB/s fallback).
cpu_baseline / --impl reference = the reference's own CPU path for this metric (BASELINE.json):
             sycl/cpu_ops.cpp dequantize_cpu (compiled into oracle/_ref) + torch CPU matmul on the
             dequantised weight, timed on this box's host cores on a bounded sample.

N > 1 (torchrun): every matrix is N-sharded (rows r*N/g..), each rank runs its shard, outputs are
all-gathered with NCCL (the only collective the path has); "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "bitsandbytes-sycl_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

# DRAM bytes per algorithmic byte of the GEMV kernel: a CONSTANT taken from the committed ncu --set full capture of the
# named round (dram__bytes_read.sum + dram__bytes_write.sum of one 11008x4096 launch), not measured inside this run
NCU_TRAFFIC_RATIO = 6704385024.0 / (2 * 3346012160.0)
NCU_TRAFFIC_SOURCE = "ncu-derived constant, round 2: dram__bytes_read + write over the 448 GEMV launches of two bench steps / algorithmic bytes = 1.0018 (profiles/r2_gemv_dram_traffic.txt)"
GEMV_KERNEL_NAME = "k_gemv4_bc<bf16,nested>"
LAYER_SHAPES_7B = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
LAYER_SHAPES_70B = [(8192, 8192), (1024, 8192), (1024, 8192), (8192, 8192), (28672, 8192), (28672, 8192), (8192, 28672)]


def algorithmic_bytes(N, K, blocksize=64):
    nblocks = N * K // blocksize
    return N * K // 2 + nblocks + 4 * ((nblocks + 255) // 256) + 2 * K + 2 * N + 1088


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML through pynvml
    (sub-millisecond queries, the timed region lasts tens of ms); falls back to polling nvidia-smi."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []      # (sm_mhz, reasons bitmask or list)
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.timed = threading.Event()   # set while the timed region runs: only those samples count
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            n = self.nvml
            names = [("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else 0x8),
                     ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
            while not self.stop_flag.is_set():
                try:
                    mhz = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                    try:
                        mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    except Exception:
                        mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    if self.timed.is_set():
                        self.samples.append((mhz, [k for k, bit in names if mask & bit]))
                except Exception:
                    pass
                self.stop_flag.wait(0.001)
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                if len(f) >= 6 and f[0].isdigit():
                    self.max_mhz = int(f[1]) if f[1].isdigit() else self.max_mhz
                    self.samples.append((int(f[0]), [nm for i, nm in enumerate(names) if f[2 + i].lower().startswith("active")]))
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = sorted({r for s in self.samples for r in s[1]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU path for this metric
# ---------------------------------------------------------------------------------------------------
_CPU_SAMPLE = {}


def _cpu_sample():
    """one 4096x4096 layer quantised with the reference's 8-bit dynamic code, blocksize 64 (setup, untimed)"""
    if not _CPU_SAMPLE:
        import numpy as np
        import torch
        from oracle import oracle as orc
        N = K = 4096
        torch.manual_seed(0)
        W = (torch.randn(N, K) * 0.02).numpy().ravel()
        # the reference's own create_dynamic_map() table (fixture generated from the reference source): the reference arm
        # loads nothing of this repo's product library
        code = np.ascontiguousarray(np.load(os.path.join(ROOT, "tests", "golden", "ref_python_tables.npz"))["dynamic_map"], dtype=np.float32)
        q, absmax = orc.quantize_blockwise(W, "fp32", code, 64, "8bit")
        kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_cpu.so")) else "port"
        _CPU_SAMPLE.update(N=N, K=K, code=code, q=q, absmax=absmax, x=torch.randn(1, K), kind=kind,
                           deq=orc.dequantize_cpu_reference if kind == "reference" else orc.dequantize_cpu_port)
        _CPU_SAMPLE["deq"](code, q, absmax, 64)  # warm
    return _CPU_SAMPLE


def cpu_reference_gemv(budget_s=12.0, max_reps=200):
    """dequantize_cpu (reference sycl/cpu_ops.cpp:7-14, 8-bit dynamic code, blocksize 64 -- the reference has
    no 4-bit CPU path) + torch CPU matmul x @ Wdeq.T for a 4096x4096 layer; GB/s is
    quoted on the NF4 algorithmic bytes of that layer so the unit matches the GPU arm."""
    import torch
    c = _cpu_sample()
    N, K = c["N"], c["K"]
    t0 = time.perf_counter()
    reps = 0
    while reps < max_reps and (time.perf_counter() - t0) < budget_s:
        Wd = torch.from_numpy(c["deq"](c["code"], c["q"], c["absmax"], 64)).view(N, K)
        torch.matmul(c["x"], Wd.t())
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    gbs = algorithmic_bytes(N, K) / dt / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": torch.get_num_threads(), "kind": c["kind"],
            "sample": f"{reps} x (dequantize_cpu 4096x4096 bs64 [1 thread] + torch CPU matmul [{torch.get_num_threads()} threads]); "
                      f"{dt * 1e3:.1f} ms per 4096x4096 GEMV", "ms_per_gemv": dt * 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    steps, warm = max(args.steps, 1), args.warmup
    per_step_budget = min(8.0, 150.0 / (steps + warm))
    for _ in range(warm):
        cpu_reference_gemv(budget_s=per_step_budget, max_reps=20)
    t0 = time.perf_counter()
    vals = [cpu_reference_gemv(budget_s=per_step_budget, max_reps=20) for _ in range(steps)]
    total = time.perf_counter() - t0
    v = sum(x["value"] for x in vals) / len(vals)
    base = dict(vals[-1])
    base["value"] = v
    line = {"impl": "reference", "metric": "nf4_gemv_4bit_GBps", "value": v, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": total / steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "llama2-7b-nf4-gemv-b1 (bounded sample: one 4096x4096 layer per rep)",
                       "blocksize": 64, "l2": "n/a (CPU)"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
DECODER_GROUPS = [[0, 1, 2], [3], [4, 5], [6]]      # q/k/v | o | gate/up | down: linears that read the same x


def _timed(fn, steps, warmup, world, dev):
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _capture(fn, no_graph=False):
    """CUDA graph of one step (the inner loop is launch-bound: ~3 us kernels vs ~20 us of Python per call); falls
    back to eager when capture is not possible."""
    import torch
    if no_graph:
        return fn, False
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        torch.cuda.synchronize()
        return g.replay, True
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches\n")
        torch.cuda.synchronize()
        return fn, False


def run_stack(args, workload, layers, full, world, rank, dev, sampler=None):
    """Build one NF4 linear stack (llama2-7b / llama3-70b shapes), time `steps` passes of gemv_4bit over it.
    full=True: also the end-to-end number through the public API and the shared-input variant (headline run)."""
    import torch
    import torch.distributed as dist

    import bnb_b200
    from bnb_b200 import functional as F
    from bnb_b200.parallel import all_gather_features, shard_quantized_weight

    shapes = LAYER_SHAPES_7B if workload == "llama2-7b" else LAYER_SHAPES_70B
    if workload == "llama3-70b":
        layers = min(layers, 8)      # 8 layers = 3.5 GB of NF4+DQ weights: > L2 per GPU even when sharded 8 ways
    dtype = torch.bfloat16
    B = args.batch

    # ---- build the quantised stack (full matrices are quantised, then row-sharded: SURVEY 8e)
    torch.manual_seed(1234)
    mats = []       # (packed_shard, state_shard, N, K)
    full_first = []  # N > 1: the unsharded first layer, for the bit-identity check before timing
    alg_bytes = 0
    for layer in range(layers):
        for (N, K) in shapes:
            W = (torch.randn(N, K, device=dev, dtype=torch.float32) * 0.02).to(dtype)
            q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
            del W
            if world > 1:
                if layer == 0:
                    full_first.append((q, st))
                q, st = shard_quantized_weight(q, st, world, rank)
            mats.append((q, st, N, K))
            alg_bytes += algorithmic_bytes(N, K)
    xs = [torch.randn(B, K, device=dev, dtype=dtype) for (_, _, _, K) in mats]
    outs = [torch.empty(B, st.shape[0], device=dev, dtype=dtype) for (_, st, _, _) in mats]
    gathered = [torch.empty(world, B, st.shape[0], device=dev, dtype=dtype) for (_, st, _, _) in mats] if world > 1 else None
    per_layer = len(shapes)
    # multi-GPU: the output all-gather is fused into the GEMV epilogue (peer stores over NVLink into symmetric
    # memory); one barrier per group of linears that feed the same consumer in a decoder layer (q/k/v | o | gate/up |
    # down) stands for the point where that consumer would wait.  --collective nccl keeps one NCCL all-gather per linear.
    peers = None
    collective = "none"
    if world > 1:
        collective = "nccl-allgather-per-linear"
        if args.collective == "fused":
            try:
                from bnb_b200.parallel import PeerOutputBuffers, sharded_gemm_push, sharded_gemv_push, sharded_gemv_push_multi
                peers = PeerOutputBuffers([N for (_, _, N, _) in mats], dtype, dev, batch=B)
                collective = "fused-epilogue-p2p-stores+symm-barrier-per-consumer-group"
                if args.barrier == "pdl":
                    peers.enable_fast_barrier()
                    collective = "fused-epilogue-p2p-stores+pdl-chained-peer-barrier-per-consumer-group"
            except Exception as e:  # noqa: BLE001
                sys.stderr.write(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather\n")
                peers = None
    groups = DECODER_GROUPS if per_layer == 7 else [[j] for j in range(per_layer)]
    group_ends = {g[-1] for g in groups}
    fuse_sharded = peers is not None and per_layer == 7 and args.fuse_same_input and B == 1
    if fuse_sharded:
        collective += "; q/k/v and gate/up share one launch"

    def sharded_step():
        for base in range(0, len(mats), per_layer):
            for grp in groups:
                ids = [base + p for p in grp]
                if fuse_sharded and len(ids) > 1:
                    sharded_gemv_push_multi(xs[ids[0]], [mats[i][0] for i in ids], [mats[i][1] for i in ids], peers, ids)
                elif B > 1:      # fused 4-bit GEMM whose epilogue stores the slice into every rank's gathered buffer
                    for i in ids:
                        sharded_gemm_push(xs[i], mats[i][0], mats[i][1], peers, i)
                else:
                    for i in ids:
                        sharded_gemv_push(xs[i], mats[i][0], mats[i][1], peers, i)
                peers.barrier()

    def step_eager():
        if peers is not None:
            sharded_step()
            return
        for i, (q, st, N, K) in enumerate(mats):
            if B == 1:
                F.gemv_4bit(xs[i], q.t(), out=outs[i], state=st)
            else:
                outs[i] = bnb_b200.matmul_4bit(xs[i], q.t(), quant_state=st)
            if world > 1:
                all_gather_features(outs[i], world, None, gathered[i])

    # ---- N > 1: every rank's gathered vectors must be bit-identical to the single-GPU kernel on the full matrix
    parity_checked = None
    if peers is not None:
        peers.buf.zero_()
        dist.barrier()
        torch.cuda.synchronize()
        sharded_step()
        peers.barrier()
        torch.cuda.synchronize()
        ok = True
        x_of = {p: (g[0] if fuse_sharded else p) for g in groups for p in g}    # a fused group reads the x of its first linear
        for i, (qf, stf) in enumerate(full_first):
            if B == 1:
                ref = F.gemv_4bit(xs[x_of[i]], qf.t(), state=stf)
                ok = ok and torch.equal(ref.view(torch.int16), peers.full(i).view(torch.int16))
            else:
                # batch > 1: the split-K of the fused GEMM depends on the number of row tiles, so the sharded run sums in
                # another order than the unsharded one -- equal within the accumulation tolerance, not bit for bit
                ref = F.gemm_4bit(xs[i], qf.t(), stf)
                got = peers.full(i)
                err = float((got.double() - ref.double()).norm() / ref.double().norm())
                ok = ok and err < 2e-3 and bool(torch.isfinite(got).all())
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity_checked = bool(int(flag.item()))
        if not parity_checked:
            raise SystemExit("[bench] N-sharded outputs differ from the single-GPU kernel: not timing a wrong result")
        del full_first[:]

    step_kernel_only, graphed = _capture(step_eager, args.no_graph)
    if peers is not None:
        launches_per_step = (len(mats) // per_layer) * (len(groups) if fuse_sharded else per_layer)
        our_kernels_per_step = launches_per_step + (len(mats) // per_layer) * len(groups)      # + one barrier kernel per group
    else:
        launches_per_step = our_kernels_per_step = len(mats)

    res = {"workload": workload, "layers": layers, "alg_bytes": alg_bytes, "n_gemv": len(mats), "graphed": graphed,
           "collective": collective, "launches_per_step": launches_per_step, "our_kernels_per_step": our_kernels_per_step,
           "parity_checked": parity_checked, "batch": B}

    if sampler is not None:
        sampler.start()
        sampler.timed.set()
    ms_total = _timed(step_kernel_only, args.steps, args.warmup, world, dev)
    if sampler is not None:
        # the timed region lasts tens of ms: keep the same load running ~0.3 s longer so the clock record has enough
        # samples of the GPU under exactly this load (these extra steps are not part of any number)
        pass
    if full:
        for _ in range(min(2000, max(1, int(300.0 / max(ms_total / args.steps, 1e-3))))):
            step_kernel_only()
        torch.cuda.synchronize()
    if sampler is not None:
        sampler.timed.clear()
        sampler.stop_flag.set()
    res["ms_per_step"] = ms_total / args.steps
    if not full:
        return res

    # ---- additive: the linears of a layer that read the same activations (q/k/v, gate/up) in ONE launch each
    # (cgemm_4bit_inference_nested_multi_*): 4 launches per layer instead of 7.  Reported beside the per-call metric.
    if world == 1 and B == 1 and per_layer == 7:
        def step_fused():
            for base in range(0, len(mats), per_layer):
                for grp in groups:
                    ids = [base + p for p in grp]
                    if len(ids) == 1:
                        i = ids[0]
                        F.gemv_4bit(xs[i], mats[i][0].t(), out=outs[i], state=mats[i][1])
                    else:
                        F.gemv_4bit_multi(xs[ids[0]], [mats[i][0].t() for i in ids], [mats[i][1] for i in ids],
                                          outs=[outs[i] for i in ids])
        fused_step, _ = _capture(step_fused, args.no_graph)
        res["fused_launches"] = (len(mats) // per_layer) * len(groups)
        res["fused_bytes"] = alg_bytes - sum((len(g) - 1) * 2 * mats[g[0]][3] for g in groups) * (len(mats) // per_layer)
        res["fused_ms"] = _timed(fused_step, args.steps, args.warmup, world, dev) / args.steps

    # ---- e2e: public API end to end: pinned host activations -> device, Linear4bit's matmul_4bit per matrix, outputs -> host
    k_total = sum(K for (_, _, _, K) in mats)
    n_total = sum(N for (_, _, N, _) in mats)
    x_host = torch.randn(B, k_total).to(dtype).pin_memory()
    x_dev = torch.empty(B, k_total, device=dev, dtype=dtype)
    y_dev = torch.empty(B, n_total, device=dev, dtype=dtype)
    y_host = torch.empty(B, n_total, dtype=dtype).pin_memory()
    offs_n, offs_k = [0], [0]
    for (_, _, N, K) in mats:
        offs_n.append(offs_n[-1] + N)
        offs_k.append(offs_k[-1] + K)

    def e2e_body():
        x_dev.copy_(x_host, non_blocking=True)
        if peers is not None:
            for base in range(0, len(mats), per_layer):
                for grp in groups:
                    ids = [base + p for p in grp]
                    xv = x_dev[:, offs_k[ids[0]]:offs_k[ids[0] + 1]]
                    if fuse_sharded and len(ids) > 1:
                        sharded_gemv_push_multi(xv, [mats[i][0] for i in ids], [mats[i][1] for i in ids], peers, ids)
                    elif B > 1:
                        for i in ids:
                            sharded_gemm_push(x_dev[:, offs_k[i]:offs_k[i + 1]], mats[i][0], mats[i][1], peers, i)
                    else:
                        for i in ids:
                            sharded_gemv_push(x_dev[:, offs_k[i]:offs_k[i + 1]], mats[i][0], mats[i][1], peers, i)
                    peers.barrier()
            if B == 1 and peers.offsets[-1] + peers.sizes[-1] == n_total:
                y_host.copy_(peers.buf[:n_total].view(B, n_total), non_blocking=True)   # gathered vectors, straight from symmetric memory
            else:
                for j, (_, _, Nj, _) in enumerate(mats):
                    y_dev[:, offs_n[j]:offs_n[j] + Nj] = peers.full(j)
                y_host.copy_(y_dev, non_blocking=True)
            return
        for i, (q, st, N, K) in enumerate(mats):
            if world == 1 and B == 1:
                # `out=` is part of the reference signature (matmul_4bit(A, B, quant_state, out, bias)): the GEMV
                # writes straight into its slice of the step's output buffer
                bnb_b200.matmul_4bit(x_dev[:, offs_k[i]:offs_k[i + 1]], q.t(), quant_state=st, out=y_dev[:, offs_n[i]:offs_n[i + 1]])
            else:
                y = bnb_b200.matmul_4bit(x_dev[:, offs_k[i]:offs_k[i + 1]], q.t(), quant_state=st)
                if world > 1:
                    y = all_gather_features(y, world)
                y_dev[:, offs_n[i]:offs_n[i + 1]] = y.reshape(B, N)
        y_host.copy_(y_dev, non_blocking=True)

    e2e_launch, e2e_graphed = _capture(e2e_body, args.no_graph)

    def step_e2e():
        e2e_launch()
        torch.cuda.current_stream().synchronize()
        return float(y_host[0, 0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, args.steps // 2)
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    res.update(e2e_s_per_step=e2e_s / e2e_steps, e2e_graphed=e2e_graphed, h2d=x_host.numel() * 2, d2h=y_host.numel() * 2)
    return res


def run_igemmlt_extras(steps, warmup):
    """BASELINE metric component 2: int8 igemmlt TOPS on config 3 (4096 tokens x 4096 -> 16384, threshold 6.0), through the
    reference ABI (cigemmlt_turing_32: col32 x col_turing -> col32 int32), the B200-native row-major entry
    (cigemm_rowmajor_32) and the fused Linear8bitLt.forward.  Rank 0, one GPU, outside the headline's timed region."""
    import torch

    import bnb_b200
    from bnb_b200 import functional as F

    m, k, n = 4096, 4096, 16384
    ops = 2.0 * m * n * k
    peak_bf16 = 1551.4
    try:
        peak_bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    torch.manual_seed(3)
    CA = torch.randint(-127, 128, (m, k), dtype=torch.int8, device="cuda")
    CB = torch.randint(-127, 128, (n, k), dtype=torch.int8, device="cuda")
    out32 = torch.empty(m, n, dtype=torch.int32, device="cuda")
    C32A, SA = F.transform(CA, "col32")
    CxB, SB = F.transform(CB, "col_turing")
    o32, So = F.igemmlt(C32A, CxB, SA, SB)
    A16 = torch.randn(m, k, device="cuda").half()
    A16[:, [7, 100, 2000, 3000]] = 8.0                  # four outlier feature columns (>= threshold 6.0)
    lin = bnb_b200.nn.Linear8bitLt(k, n, bias=True, has_fp16_weights=False, threshold=6.0).cuda().half()
    CA_host = CA.cpu().pin_memory()
    A16_host = A16.cpu().pin_memory()
    y_host = torch.empty(16, n, dtype=torch.float16).pin_memory()
    c_host = torch.empty(16 * n, dtype=torch.int32).pin_memory()

    def ev_time(fn, reps):
        for _ in range(max(3, warmup)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def host_time(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    def entry(ms, ms_e2e, h2d, d2h, api, launches):
        tops = ops / (ms * 1e-3) / 1e12
        return {"TOPS": tops, "ms": ms, "ops": ops, "api": api, "gpu_launches_per_call": launches,
                "roofline": {"bound": "tensor", "achieved": tops, "unit": "TOP/s",
                             "peak_2x_measured_bf16": 2 * peak_bf16, "frac_of_2x_measured_bf16": tops / (2 * peak_bf16),
                             "peak_nominal_int8": 4500.0, "frac_of_nominal_int8": tops / 4500.0},
                "e2e": {"value": ops / (ms_e2e * 1e-3) / 1e12, "unit": "TOP/s", "ms": ms_e2e, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "note": "operand A from pinned host memory in, first 16 output rows back"}}

    reps = max(5, min(20, steps))
    out = {"config": "4096 tokens x (4096 -> 16384), int8, threshold 6.0 (BASELINE config 3)", "l2": "operands + output 352 MB > L2"}
    with torch.no_grad():
        # (1) reference ABI
        f1 = lambda: F.igemmlt(C32A, CxB, SA, SB, out=o32, Sout=So)   # noqa: E731

        def f1e():
            C32A.view(-1)[: m * k].copy_(CA_host.view(-1), non_blocking=True)   # same bytes, any layout: the copy is what is timed
            F.igemmlt(C32A, CxB, SA, SB, out=o32, Sout=So)
            c_host.copy_(o32.view(-1)[: 16 * n], non_blocking=True)
        out["cigemmlt_turing_32"] = entry(ev_time(f1, reps), host_time(f1e, reps), m * k, 16 * n * 4,
                                          "F.igemmlt(col32, col_turing) -> cigemmlt_turing_32 (reference ABI, pythonInterface.cpp:298)", None)
        # (2) row-major entry
        f2 = lambda: F.igemmlt(CA, CB, ((m, k), "row"), ((n, k), "row"), out=out32, Sout=((m, n), "row"))   # noqa: E731

        def f2e():
            CA.copy_(CA_host, non_blocking=True)
            F.igemmlt(CA, CB, ((m, k), "row"), ((n, k), "row"), out=out32, Sout=((m, n), "row"))
            c_host.copy_(out32.view(-1)[: 16 * n], non_blocking=True)
        out["cigemm_rowmajor_32"] = entry(ev_time(f2, reps), host_time(f2e, reps), m * k, 16 * n * 4,
                                          "F.igemmlt(row, row) -> cigemm_rowmajor_32 (additive, TMA-native layout)", 1)
        # (3) fused Linear8bitLt forward (quantise + outliers + GEMM + dequant epilogue)
        f3 = lambda: lin(A16)   # noqa: E731

        def f3e():
            A16.copy_(A16_host, non_blocking=True)
            y = lin(A16)
            y_host.copy_(y[:16], non_blocking=True)
        out["linear8bitlt_forward"] = entry(ev_time(f3, reps), host_time(f3e, reps), m * k * 2, 16 * n * 2,
                                            "bnb_b200.nn.Linear8bitLt.forward (has_fp16_weights=False, threshold 6.0, bias)", None)
    return out


def run_gemm4_extras(steps, warmup):
    """BASELINE config 4: fused NF4 GEMM, batch 16-256 on the Llama-3-8B MLP shape 14336x4096 (bf16, blocksize 64), through
    `functional.gemm_4bit` (what `MatMul4Bit.forward` calls for batch > 1), beside the reference's composition
    (`dequantize_4bit` + `F.linear`, _functions.py:490-518).  Six rotating copies of the packed weight (176 MB > L2), one
    CUDA graph of six calls per measurement.  Rank 0, one GPU, outside the headline's timed region."""
    import torch

    from bnb_b200 import functional as F

    peak_hbm, _ = measured_peak_gbs()
    N, K = 14336, 4096
    torch.manual_seed(5)
    W = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=False, quant_type="nf4")
    del W
    qs = [q.clone() for _ in range(6)]
    reps = max(5, min(20, steps))

    def graph_us(fns):
        def step():
            for f in fns:
                f()
        run, graphed = _capture(step)
        for _ in range(max(3, warmup)):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * len(fns)), graphed

    out = {"config": "NF4 GEMM 14336x4096 bf16, blocksize 64 (BASELINE config 4)", "l2": "6 rotating weight copies, 176 MB > L2",
           "api": "functional.gemm_4bit -> cgemm_4bit_bf16 (additive C-ABI entry)"}
    with torch.no_grad():
        for batch in (16, 32, 64, 256):
            x = torch.randn(batch, K, device="cuda").bfloat16()
            outs = [torch.empty(batch, N, dtype=torch.bfloat16, device="cuda") for _ in range(6)]
            if F.gemm_4bit(x, qs[0].t(), st, out=outs[0]) is None:
                out[f"batch{batch}"] = {"error": "fused kernel refused the shape"}
                continue
            us, graphed = graph_us([(lambda i=i: F.gemm_4bit(x, qs[i].t(), st, out=outs[i])) for i in range(6)])
            us_ref, _ = graph_us([(lambda i=i: torch.nn.functional.linear(x, F.dequantize_4bit(qs[i], st))) for i in range(6)])
            nbytes = N * K // 2 + 4 * N * K // 64 + 2 * batch * (K + N)
            out[f"batch{batch}"] = {"us": us, "GBps": nbytes / us / 1e3, "hbm_frac": nbytes / us / 1e3 / peak_hbm,
                                    "TFLOPs": 2.0 * batch * N * K / us / 1e6, "reference_composition_us": us_ref,
                                    "speedup_vs_composition": us_ref / us, "launch": "cuda-graph" if graphed else "eager"}
            del outs
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--layers", type=int, default=32, help="decoder layers in the stack (32 = the whole Llama-2-7B)")
    ap.add_argument("--workload", default="llama2-7b", choices=["llama2-7b", "llama3-70b"])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the igemmlt, gemm_4bit and 70B-stack extra keys")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"],
                    help="N>1: output all-gather fused into the GEMV epilogue (peer stores), or one NCCL all-gather per linear")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph of the step")
    ap.add_argument("--barrier", default="pdl", choices=["pdl", "symm"],
                    help="N>1: consumer-group barrier = cbnb_peer_barrier (a link of the PDL chain) or torch symmetric-memory barrier")
    ap.add_argument("--no-fuse-same-input", dest="fuse_same_input", action="store_false",
                    help="N>1: one launch per linear instead of one per group of linears that read the same x (q/k/v, gate/up)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    head = run_stack(args, args.workload, args.layers, True, world, rank, dev, sampler)
    torch.cuda.empty_cache()

    extras = {}
    if not args.no_extras and args.batch == 1:
        if args.workload == "llama2-7b":
            try:
                r70 = run_stack(args, "llama3-70b", 8, False, world, rank, dev)
                gbs = r70["alg_bytes"] / (r70["ms_per_step"] * 1e-3) / 1e9
                extras["llama3_70b"] = {
                    "value": gbs, "unit": "GB/s", "ms_per_step": r70["ms_per_step"], "layers": r70["layers"], "n_gpus": world,
                    "tok_per_s": 1e3 / (r70["ms_per_step"] * 80.0 / r70["layers"]),
                    "tok_per_s_note": "batch-1 decode rate of the 80-layer Llama-3-70B LINEAR stack (q/k/v/o/gate/up/down NF4 GEMVs "
                                      "only; attention, norms and sampling are outside SURVEY section 8), extrapolated from the timed "
                                      f"{r70['layers']}-layer stack",
                    "algorithmic_bytes_per_step": r70["alg_bytes"], "launches_per_step": r70["launches_per_step"],
                    "collective": r70["collective"], "parity_checked": r70["parity_checked"],
                    "frac_of_hbm_peak_per_gpu": gbs / world / measured_peak_gbs()[0]}
            except Exception as e:  # noqa: BLE001
                extras["llama3_70b"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
        if world > 1:
            dist.barrier()
        if rank == 0:
            try:
                extras["igemmlt"] = run_igemmlt_extras(args.steps, args.warmup)
            except Exception as e:  # noqa: BLE001
                extras["igemmlt"] = {"error": f"{type(e).__name__}: {e}"}
            try:
                extras["gemm_4bit"] = run_gemm4_extras(args.steps, args.warmup)
            except Exception as e:  # noqa: BLE001
                extras["gemm_4bit"] = {"error": f"{type(e).__name__}: {e}"}
        if world > 1:
            dist.barrier()

    if rank == 0:
        ms_per_step = head["ms_per_step"]
        alg_bytes = head["alg_bytes"]
        value = alg_bytes / (ms_per_step * 1e-3) / 1e9
        peak, peak_kind = measured_peak_gbs()
        # dominant kernel: the GEMV; bytes per launch and launch time averaged over the launches actually issued (per rank)
        bytes_per_launch = alg_bytes / head["launches_per_step"] / world
        launch_ms = ms_per_step / head["launches_per_step"]
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        B = args.batch
        line = {
            "metric": "nf4_gemv_4bit_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}-nf4-gemv-b{B}", "layers": head["layers"], "gemvs_per_step": head["n_gemv"],
                       "blocksize": 64, "nested_absmax": True, "algorithmic_bytes_per_step": alg_bytes,
                       "l2": f"inputs larger than L2 ({alg_bytes / 1e6:.0f} MB of weights per step, distinct per launch)",
                       "launch": "cuda-graph" if head["graphed"] else "eager", "graph_shape": "chain",
                       "parallelism": f"n-shard{world}" if world > 1 else "single", "collective": head["collective"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": bytes_per_launch * NCU_TRAFFIC_RATIO if B == 1 else None,
                         "traffic_source": NCU_TRAFFIC_SOURCE,
                         "kernel": GEMV_KERNEL_NAME, "peak_source": peak_kind,
                         "bytes_per_launch": bytes_per_launch, "launch_us": launch_ms * 1e3,
                         "launches_per_step": head["launches_per_step"]},
            "e2e": {"value": alg_bytes / head["e2e_s_per_step"] / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": head["h2d"], "d2h_bytes_per_step": head["d2h"],
                    "launch": "cuda-graph" if head["e2e_graphed"] else "eager",
                    "api": "bnb_b200.matmul_4bit (Linear4bit.forward path), pinned host x in, y out"},
            "gpu_launches": head["our_kernels_per_step"] * args.steps,
            "clocks": sampler.summary(),
        }
        if head["parity_checked"] is not None:
            line["parity_checked"] = head["parity_checked"]
        if "fused_ms" in head:
            fv = head["fused_bytes"] / (head["fused_ms"] * 1e-3) / 1e9
            line["fused_same_input"] = {
                "value": fv, "unit": "GB/s", "ms_per_step": head["fused_ms"],
                "launches_per_step": head["fused_launches"], "frac_of_hbm_peak": fv / peak,
                "api": "bnb_b200.functional.gemv_4bit_multi: q/k/v and gate/up of a layer share one launch "
                       "(bit-identical outputs); NOT the headline value, which stays one call per matrix"}
        line.update(extras)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_gemv()
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": None, "kind": "reference",
                                    "sample": "reported at N=1 only"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL / symmetric-memory teardown after graph capture has hung at exit on this stack: every number is
        # out, so synchronise, flush and leave without running the destructors
        try:
            dist.barrier()
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
