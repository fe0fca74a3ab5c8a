"""N > 1 host logic on CPU: two processes over gloo (world_size 2).  The weight is quantised once (CPU oracle: NF4,
blocksize 64, nested absmax), every rank takes its row shard with bnb_b200.parallel.shard_quantized_weight, computes
its slice of the output from its shard alone (oracle de-nest + dequantise), and all_gather_features must hand every rank
the full output vector -- equal to the product with the unsharded weight.  No GPU, no compute call into the CUDA library."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build_full(N, K):
    from oracle import oracle as orc
    from bnb_b200 import functional as F
    rng = np.random.default_rng(11)
    W = (rng.standard_normal((N, K)) * 0.02).astype(np.float32)
    q, absmax = orc.quantize_blockwise(W.ravel(), "fp32", None, 64, "nf4")
    offset = np.float32(absmax.mean())
    code2 = F.create_dynamic_map().numpy()
    qabs, absmax2 = orc.quantize_blockwise((absmax - offset).astype(np.float32), "fp32", code2, 256, "8bit")
    s2 = F.QuantState(absmax=torch.from_numpy(absmax2.copy()), code=torch.from_numpy(code2.copy()), blocksize=256, dtype=torch.float32)
    st = F.QuantState(absmax=torch.from_numpy(qabs.copy()), shape=torch.Size((N, K)), code=torch.from_numpy(orc.nf4_table().copy()),
                      blocksize=64, quant_type="nf4", dtype=torch.float32, offset=torch.tensor(float(offset)), state2=s2)
    return torch.from_numpy(q.copy()).reshape(-1, 1), st


def _dequant(packed, st):
    from oracle import oracle as orc
    n = st.shape[0] * st.shape[1]
    absmax = orc.denest_absmax(st.absmax.numpy(), st.state2.absmax.numpy(), st.state2.code.numpy(),
                               np.float32(st.offset.item()), st.state2.blocksize)
    w = orc.dequantize_blockwise(packed.numpy().ravel(), absmax, n, "fp32", None, 64, "nf4")
    return torch.from_numpy(np.asarray(w, np.float32).reshape(st.shape[0], st.shape[1]).copy())


def _worker(rank, world, port, N, K, q):
    try:
        for p in (ROOT, os.path.join(ROOT, "bitsandbytes-sycl_b200")):
            if p not in sys.path:
                sys.path.insert(0, p)
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        from bnb_b200.parallel import all_gather_features, shard_bounds, shard_quantized_weight
        packed, st = _build_full(N, K)                       # same seed on every rank: identical full weight
        r0, r1 = shard_bounds(N, world, rank)
        assert (r0, r1) == (rank * N // world, (rank + 1) * N // world)
        p_sh, st_sh = shard_quantized_weight(packed, st, world, rank)
        # a shard is a plain slice of the full statistics (bit-exact), with the GLOBAL offset and code tables
        bpr, blocks = K // 2, K // 64
        assert torch.equal(p_sh.reshape(-1), packed.reshape(-1)[r0 * bpr:r1 * bpr])
        assert torch.equal(st_sh.absmax, st.absmax[r0 * blocks:r1 * blocks])
        assert torch.equal(st_sh.state2.absmax, st.state2.absmax[r0 * blocks // 256:(r1 * blocks + 255) // 256])
        assert float(st_sh.offset) == float(st.offset) and tuple(st_sh.shape) == (r1 - r0, K)
        torch.manual_seed(5)
        x = torch.randn(3, K)
        W_full, W_sh = _dequant(packed, st), _dequant(p_sh, st_sh)
        assert torch.equal(W_sh, W_full[r0:r1])              # de-nesting a shard == slicing the de-nested full weight
        y_local = x @ W_sh.t()
        y = all_gather_features(y_local, world)               # [3, N] on every rank, feature-major concatenation
        ref = x @ W_full.t()
        assert y.shape == ref.shape
        assert torch.equal(y[:, r0:r1], y_local)
        assert torch.allclose(y, ref, rtol=1e-5, atol=1e-6)
        # pre-allocated gather buffer variant
        buf = torch.empty(world, 3, (r1 - r0))
        y2 = all_gather_features(y_local, world, None, buf)
        assert torch.equal(y2, y)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("N,K", [(64, 1024), (128, 512)])
def test_n_sharding_and_output_gather_over_gloo_world2(N, K):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}:\n{msg}"
