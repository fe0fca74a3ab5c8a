#!/usr/bin/env python
"""Multi-GPU check of the N-sharded GEMV with the all-gather fused into the epilogue (run under torchrun, one
rank per GPU):  torchrun --nproc-per-node 2 tools/check_multi_gpu.py
Every rank must end up with the full output vector, bit-identical to the single-GPU kernel on the full matrix."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from bnb_b200 import functional as F  # noqa: E402
from bnb_b200.parallel import PeerOutputBuffers, shard_quantized_weight, sharded_gemv_push, all_gather_features  # noqa: E402

shapes = [(4096, 4096), (11008, 4096), (4096, 11008), (8192, 8192)]
torch.manual_seed(7)            # same seed on every rank: identical full matrices
full, shards, xs = [], [], []
for (N, K) in shapes:
    W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
    full.append((q, st))
    shards.append(shard_quantized_weight(q, st, world, rank))
    xs.append(torch.randn(1, K, device=dev).bfloat16())
peers = PeerOutputBuffers([N for (N, _) in shapes], torch.bfloat16, dev)
peers.enable_kernel_sync(ngroups=len(shapes))
ok = True
for it in range(4):
    peers.buf.zero_()
    dist.barrier()
    torch.cuda.synchronize()
    if it % 2 == 0:      # ordering by one barrier launch
        for i, (qs, sts) in enumerate(shards):
            sharded_gemv_push(xs[i], qs, sts, peers, i)
        peers.barrier()
    else:                # ordering folded into the kernels: every linear is its own consumer group; the trailing
        for i, (qs, sts) in enumerate(shards):   # barrier launch only because the HOST looks at the result next
            sharded_gemv_push(xs[i], qs, sts, peers, i, peers.sync_desc(i, True, True))
        peers.bump_epoch()
        peers.barrier()
    torch.cuda.synchronize()
    for i, (q, st) in enumerate(full):
        ref = F.gemv_4bit(xs[i], q.t(), state=st)
        got = peers.full(i)
        same = torch.equal(ref.view(torch.int16), got.view(torch.int16))
        # and the NCCL route gives the same bits
        y = F.gemv_4bit(xs[i], shards[i][0].t(), state=shards[i][1])
        nc = all_gather_features(y, world)
        same_nccl = torch.equal(ref.view(torch.int16), nc.reshape(1, -1).view(torch.int16))
        if not (same and same_nccl):
            ok = False
            print(f"rank {rank} iter {it} shape {shapes[i]}: fused {same} nccl {same_nccl}", flush=True)
# ---- several linears that share x in ONE launch (q/k/v style: three matrices, same K, same x), fused all-gather
from bnb_b200.parallel import sharded_gemv_push_multi  # noqa: E402
mshapes = [(4096, 4096), (1024, 4096), (1024, 4096)]
mfull, mshards = [], []
for (N, K) in mshapes:
    W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    q, st = F.quantize_4bit(W, blocksize=64, compress_statistics=True, quant_type="nf4")
    mfull.append((q, st))
    mshards.append(shard_quantized_weight(q, st, world, rank))
xm = torch.randn(1, 4096, device=dev).bfloat16()
mpeers = PeerOutputBuffers([N for (N, _) in mshapes], torch.bfloat16, dev)
mpeers.buf.zero_()
dist.barrier()
torch.cuda.synchronize()
for it in range(12):
    if it == 2:
        mpeers.enable_fast_barrier()      # from here on: cbnb_peer_barrier (a link of the PDL chain) instead of torch's
    xi = (xm * (1.0 + 0.25 * it)).bfloat16()
    sharded_gemv_push_multi(xi, [s[0] for s in mshards], [s[1] for s in mshards], mpeers, [0, 1, 2])
    mpeers.barrier()
    got = [mpeers.full(i).clone() for i in range(3)]    # stream-ordered after the barrier: every peer's slice is in
    mpeers.barrier()                                     # nobody overwrites before everybody has read
    torch.cuda.synchronize()
    for i, (q, st) in enumerate(mfull):
        ref = F.gemv_4bit(xi, q.t(), state=st)
        if not torch.equal(ref.view(torch.int16), got[i].view(torch.int16)):
            ok = False
            print(f"rank {rank} multi-launch iter {it} shape {mshapes[i]}: MISMATCH", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("multi-gpu fused all-gather:", "OK (bit-identical on every rank)" if int(flag.item()) else "MISMATCH", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if int(flag.item()) else 1)
