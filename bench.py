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

NCU_TRAFFIC_RATIO = 23.306496 / 23.291200   # measured DRAM bytes per algorithmic byte (ncu, 11008x4096 GEMV)
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
        import torch
        from oracle import oracle as orc
        from bnb_b200.functional import create_dynamic_map
        N = K = 4096
        torch.manual_seed(0)
        W = (torch.randn(N, K) * 0.02).numpy().ravel()
        code = create_dynamic_map().numpy()
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
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"],
                    help="N>1: output all-gather fused into the GEMV epilogue (peer stores), or one NCCL all-gather per linear")
    ap.add_argument("--sync", default="barrier", choices=["kernel", "barrier"],
                    help="N>1 fused collective: ordering folded into the GEMV kernels, or one symmetric-memory barrier launch per consumer group")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph of the step")
    ap.add_argument("--barrier", default="pdl", choices=["pdl", "symm"],
                    help="N>1: consumer-group barrier = cbnb_peer_barrier (a link of the PDL chain) or torch symmetric-memory barrier")
    ap.add_argument("--no-fuse-same-input", dest="fuse_same_input", action="store_false",
                    help="N>1: one launch per linear instead of one per group of linears that read the same x (q/k/v, gate/up)")
    ap.add_argument("--graph-shape", default="chain", choices=["decoder", "chain"],
                    help="N=1: dependency structure of the step. chain = all launches on one stream, overlapped by "
                         "programmatic dependent launch (default, measured faster: 2392 vs 2148 GB/s); decoder = the "
                         "layer's own structure, {q,k,v} in parallel -> o -> {gate,up} in parallel -> down, on three "
                         "capture streams (cross-stream graph edges are full dependencies without PDL)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import bnb_b200
    from bnb_b200 import functional as F
    from bnb_b200.parallel import all_gather_features, shard_quantized_weight

    shapes = LAYER_SHAPES_7B if args.workload == "llama2-7b" else LAYER_SHAPES_70B
    if args.workload == "llama3-70b":
        args.layers = min(args.layers, 8)      # 8 layers = 3.5 GB of NF4+DQ weights: > L2 per GPU even when sharded 8 ways
    dtype = torch.bfloat16
    B = args.batch

    # ---- build the quantised stack (full matrices are quantised, then row-sharded: SURVEY 8e)
    torch.manual_seed(1234)
    mats = []       # (packed_shard, state_shard, N, K)
    alg_bytes = 0
    for layer in range(args.layers):
        for (N, K) in shapes:
            W = (torch.randn(N, K, device=dev, dtype=torch.float32) * 0.02).to(dtype)
            q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
            del W
            if world > 1:
                q, st = shard_quantized_weight(q, st, world, rank)
            mats.append((q, st, N, K))
            alg_bytes += algorithmic_bytes(N, K)
    xs = [torch.randn(B, K, device=dev, dtype=dtype) for (_, _, _, K) in mats]
    outs = [torch.empty(B, st.shape[0], device=dev, dtype=dtype) for (_, st, _, _) in mats]
    gathered = [torch.empty(world, B, st.shape[0], device=dev, dtype=dtype) for (_, st, _, _) in mats] if world > 1 else None
    n_launch_per_step = len(mats)
    # multi-GPU: the output all-gather is fused into the GEMV epilogue (peer stores over NVLink into symmetric
    # memory); one signal-pad barrier per group of linears that feed the same consumer in a decoder layer
    # (q/k/v | o | gate/up | down) stands for the point where that consumer would wait.  --collective nccl keeps
    # the plain NCCL all-gather per linear.
    peers = None
    collective = "none"
    if world > 1:
        collective = "nccl-allgather-per-linear"
        if args.collective == "fused" and B == 1:
            try:
                from bnb_b200.parallel import PeerOutputBuffers, sharded_gemv_push
                peers = PeerOutputBuffers([N for (_, _, N, _) in mats], dtype, dev)
                collective = "fused-epilogue-p2p-stores+symm-barrier-per-consumer-group"
                if args.barrier == "pdl" and args.sync != "kernel":
                    peers.enable_fast_barrier()
                    collective = "fused-epilogue-p2p-stores+pdl-chained-peer-barrier-per-consumer-group"
                if args.sync == "kernel":
                    peers.enable_kernel_sync(ngroups=args.layers * 4 if len(shapes) == 7 else len(mats))
                    collective = "fused-epilogue-p2p-stores+in-kernel-signal/wait-per-consumer-group"
            except Exception as e:  # noqa: BLE001
                sys.stderr.write(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather\n")
                peers = None
    per_layer = len(shapes)
    group_ends = {2, 3, 5, 6} if per_layer == 7 else set(range(per_layer))

    group_starts = {0, 3, 4, 6} if per_layer == 7 else set(range(per_layer))
    kernel_sync = peers is not None and args.sync == "kernel"
    syncs = None
    if kernel_sync:   # one descriptor per linear: group index, "ends a group" -> signal, "starts a group" -> wait
        syncs, gi = [], 0
        for i in range(len(mats)):
            pos = i % per_layer
            syncs.append(peers.sync_desc(gi, pos in group_ends, pos in group_starts))
            if pos in group_ends:
                gi += 1

    def fused_linear(i, x, q, st):
        if kernel_sync:
            sharded_gemv_push(x, q, st, peers, i, syncs[i])
            if i == len(mats) - 1:
                peers.bump_epoch()      # next pass counts on: sequence numbers never repeat across graph replays
        else:
            sharded_gemv_push(x, q, st, peers, i)
            if (i % per_layer) in group_ends:
                peers.barrier()

    # N=1: the linears of a decoder layer that read the same activations (q/k/v, gate/up) do not depend on each other:
    # they are launched on parallel streams (parallel branches of the captured graph), everything else stays ordered.
    decoder_groups = [[0, 1, 2], [3], [4, 5], [6]] if per_layer == 7 else [[j] for j in range(per_layer)]
    structured = world == 1 and args.graph_shape == "decoder" and per_layer == 7
    side_streams = [torch.cuda.Stream(device=dev) for _ in range(2)] if structured else []

    def run_structured(call):
        main = torch.cuda.current_stream()
        for base in range(0, len(mats), per_layer):
            for grp in decoder_groups:
                if len(grp) == 1:
                    call(base + grp[0])
                    continue
                fork = torch.cuda.Event()
                fork.record(main)
                for j, pos in enumerate(grp):
                    if j == 0:
                        call(base + pos)
                    else:
                        sst = side_streams[j - 1]
                        sst.wait_event(fork)
                        with torch.cuda.stream(sst):
                            call(base + pos)
                for j in range(1, len(grp)):
                    join = torch.cuda.Event()
                    join.record(side_streams[j - 1])
                    main.wait_event(join)

    fuse_sharded = peers is not None and not kernel_sync and per_layer == 7 and args.fuse_same_input
    if fuse_sharded:
        from bnb_b200.parallel import sharded_gemv_push_multi
        collective += "; q/k/v and gate/up share one launch"

    def step_eager():
        if fuse_sharded:     # N-sharded: one launch per group of linears that read the same x, one barrier per group
            for base in range(0, len(mats), per_layer):
                for grp in decoder_groups:
                    ids = [base + p for p in grp]
                    if len(ids) == 1:
                        sharded_gemv_push(xs[ids[0]], mats[ids[0]][0], mats[ids[0]][1], peers, ids[0])
                    else:
                        sharded_gemv_push_multi(xs[ids[0]], [mats[i][0] for i in ids], [mats[i][1] for i in ids], peers, ids)
                    peers.barrier()
            return
        if structured and B == 1:
            run_structured(lambda i: F.gemv_4bit(xs[i], mats[i][0].t(), out=outs[i], state=mats[i][1]))
            return
        for i, (q, st, N, K) in enumerate(mats):
            if peers is not None:
                fused_linear(i, xs[i], q, st)
                continue
            if B == 1:
                F.gemv_4bit(xs[i], q.t(), out=outs[i], state=st)
            else:
                outs[i] = bnb_b200.matmul_4bit(xs[i], q.t(), quant_state=st)
            if world > 1:
                all_gather_features(outs[i], world, None, gathered[i])

    def capture(fn):
        """CUDA graph of one step (the inner loop is launch-bound: ~3 us kernels vs ~20 us of Python per
        call); falls back to eager when capture is not possible."""
        if args.no_graph:
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

    step_kernel_only, graphed = capture(step_eager)

    # ---- e2e buffers: pinned host activations in, outputs back
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
        # public API end to end: pinned host activations -> device, Linear4bit's matmul_4bit per matrix, outputs -> host
        x_dev.copy_(x_host, non_blocking=True)
        if structured and B == 1:
            run_structured(lambda i: bnb_b200.matmul_4bit(x_dev[:, offs_k[i]:offs_k[i + 1]], mats[i][0].t(), quant_state=mats[i][1],
                                                          out=y_dev[:, offs_n[i]:offs_n[i + 1]]))
            y_host.copy_(y_dev, non_blocking=True)
            return
        ko = no = 0
        for i, (q, st, N, K) in enumerate(mats):
            if peers is not None:
                fused_linear(i, x_dev[:, ko:ko + K], q, st)
            else:
                if world == 1 and B == 1:
                    # `out=` is part of the reference signature (matmul_4bit(A, B, quant_state, out, bias)): the GEMV
                    # writes straight into its slice of the step's output buffer
                    bnb_b200.matmul_4bit(x_dev[:, ko:ko + K], q.t(), quant_state=st, out=y_dev[:, no:no + N])
                else:
                    y = bnb_b200.matmul_4bit(x_dev[:, ko:ko + K], q.t(), quant_state=st)
                    if world > 1:
                        y = all_gather_features(y, world)
                    y_dev[:, no:no + N] = y.reshape(B, N)
            ko += K
            no += N
        if kernel_sync:
            peers.barrier()      # the host read below consumes every gathered vector of the pass
        if peers is not None and peers.offsets[-1] + peers.sizes[-1] == n_total:
            y_host.copy_(peers.buf[:n_total].view(B, n_total), non_blocking=True)   # gathered vectors, straight from symmetric memory
        else:
            if peers is not None:
                for j, (_, _, Nj, _) in enumerate(mats):
                    y_dev[:, offs_n[j]:offs_n[j] + Nj] = peers.full(j)
            y_host.copy_(y_dev, non_blocking=True)

    e2e_launch, e2e_graphed = capture(e2e_body)

    def step_e2e():
        e2e_launch()
        torch.cuda.current_stream().synchronize()
        return float(y_host[0, 0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
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

    # ---- additive: the linears of a layer that read the same activations (q/k/v, gate/up) in ONE launch each
    # (cgemm_4bit_inference_nested_multi_*): 4 launches per layer instead of 7.  Reported beside the per-call metric.
    fused_step = None
    if world == 1 and B == 1 and per_layer == 7:
        def step_fused():
            for base in range(0, len(mats), per_layer):
                for grp in decoder_groups:
                    ids = [base + p for p in grp]
                    if len(ids) == 1:
                        i = ids[0]
                        F.gemv_4bit(xs[i], mats[i][0].t(), out=outs[i], state=mats[i][1])
                    else:
                        F.gemv_4bit_multi(xs[ids[0]], [mats[i][0].t() for i in ids], [mats[i][1] for i in ids],
                                          outs=[outs[i] for i in ids])
        fused_step, fused_graphed = capture(step_fused)
        fused_launches = (len(mats) // per_layer) * len(decoder_groups)
        fused_bytes = alg_bytes - sum((len(g) - 1) * 2 * mats[g[0]][3] for g in decoder_groups) * (len(mats) // per_layer)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.timed.set()
    ms_total = timed(step_kernel_only, args.steps, args.warmup)
    # the timed region lasts tens of ms: keep the same load running ~0.3 s longer so the clock record has enough
    # samples of the GPU under exactly this load (these extra steps are not part of any number; every rank runs
    # the same count, ms_total being the max over ranks)
    for _ in range(min(2000, max(1, int(300.0 / max(ms_total / args.steps, 1e-3))))):
        step_kernel_only()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.timed.clear()
        sampler.stop_flag.set()
    fused_ms = timed(fused_step, args.steps, args.warmup) / args.steps if fused_step is not None else None
    # e2e: host timer around the same loop (copies + API calls + sync are inside)
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

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = alg_bytes / (ms_per_step * 1e-3) / 1e9
        peak, peak_kind = measured_peak_gbs()
        # dominant kernel: one launch == one GEMV; bytes per launch averaged over the stack (per rank)
        bytes_per_launch = alg_bytes / n_launch_per_step / world
        launch_ms = ms_per_step / n_launch_per_step
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": "nf4_gemv_4bit_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}-nf4-gemv-b{B}", "layers": args.layers, "gemvs_per_step": n_launch_per_step,
                       "blocksize": 64, "nested_absmax": True, "algorithmic_bytes_per_step": alg_bytes,
                       "l2": f"inputs larger than L2 ({alg_bytes / 1e6:.0f} MB of weights per step, distinct per launch)",
                       "launch": "cuda-graph" if graphed else "eager",
                       "graph_shape": ("decoder: {q,k,v} parallel -> o -> {gate,up} parallel -> down" if structured and B == 1 else "chain"),
                       "parallelism": f"n-shard{world}" if world > 1 else "single", "collective": collective},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read+write per launch of this kernel from the committed ncu --set full capture
                         # (profiles/r1_ncu_summaries.txt: 23.306 MB for the 23.291 MB 11008x4096 GEMV, 121.23 MB
                         # for the 121.24 MB 28672x8192 one): DRAM traffic == algorithmic bytes (+0.07 %)
                         "traffic": bytes_per_launch * NCU_TRAFFIC_RATIO if B == 1 else None,
                         "traffic_source": "ncu dram bytes / algorithmic bytes = 1.0007 (profiles/r1_ncu_summaries.txt)",
                         "kernel": "k_gemv4_bc<bf16,nested>", "peak_source": peak_kind,
                         "bytes_per_launch": bytes_per_launch, "launch_us": launch_ms * 1e3},
            "e2e": {"value": alg_bytes / (e2e_s / e2e_steps) / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": x_host.numel() * 2, "d2h_bytes_per_step": y_host.numel() * 2,
                    "launch": "cuda-graph" if e2e_graphed else "eager",
                    "api": "bnb_b200.matmul_4bit (Linear4bit.forward path), pinned host x in, y out"},
            "gpu_launches": n_launch_per_step * args.steps,
            "clocks": sampler.summary(),
        }
        if fused_ms is not None:
            line["fused_same_input"] = {
                "value": fused_bytes / (fused_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": fused_ms,
                "launches_per_step": fused_launches, "frac_of_hbm_peak": fused_bytes / (fused_ms * 1e-3) / 1e9 / peak,
                "api": "bnb_b200.functional.gemv_4bit_multi: q/k/v and gate/up of a layer share one launch "
                       "(bit-identical outputs); NOT the headline value, which stays one call per matrix"}
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
