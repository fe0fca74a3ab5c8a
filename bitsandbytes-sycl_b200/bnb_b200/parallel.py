"""N-sharded (column-parallel) 4-bit linear over the GPUs of one NVLink/NVSwitch box.

The reference has no multi-device code; the north star adds exactly one thing: output features are
independent, so rank r owns rows [r*N/g, (r+1)*N/g) of the quantised weight and the only data-path
exchange is an all-gather of the [batch, N/g] output slices.  K is never split, so no reduction exists and
results are bit-identical to the single-GPU kernel.

Sharding is done AFTER quantising the full [N, K] weight, so the nested statistics (global
`offset = absmax.mean()`, `state2`) are the ones the reference would produce; a shard is a plain slice:
packed bytes [N/g, K/2], absmax entries (N/g)*(K/blocksize), and for nested absmax the matching run of
256-entry groups (whole groups only -- asserted).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import functional as F
from .functional import QuantState


def shard_bounds(N: int, world: int, rank: int) -> Tuple[int, int]:
    assert N % world == 0, f"output features {N} must divide by world size {world}"
    per = N // world
    return rank * per, (rank + 1) * per


def shard_quantized_weight(packed: torch.Tensor, state: QuantState, world: int, rank: int):
    """Slice a quantised [N, K] weight (packed [(N*K+1)//2, 1] uint8 + QuantState) into rank's row shard."""
    N, K = state.shape
    r0, r1 = shard_bounds(N, world, rank)
    assert (K % 2) == 0 and (K % state.blocksize) == 0, "rows must hold whole bytes and whole absmax blocks"
    bpr = K // 2
    p = packed.reshape(-1)[r0 * bpr:r1 * bpr].reshape(-1, 1).contiguous()
    blocks_per_row = K // state.blocksize
    b0, b1 = r0 * blocks_per_row, r1 * blocks_per_row
    if state.nested:
        bs2 = state.state2.blocksize
        assert b0 % bs2 == 0 and (b1 % bs2 == 0 or r1 == N), "nested absmax groups must not straddle ranks"
        s2 = QuantState(absmax=state.state2.absmax[b0 // bs2:(b1 + bs2 - 1) // bs2].contiguous(),
                        code=state.state2.code, blocksize=bs2, dtype=state.state2.dtype)
        st = QuantState(absmax=state.absmax[b0:b1].contiguous(), shape=torch.Size((r1 - r0, K)), code=state.code,
                        blocksize=state.blocksize, quant_type=state.quant_type, dtype=state.dtype,
                        offset=state.offset, state2=s2)
    else:
        st = QuantState(absmax=state.absmax[b0:b1].contiguous(), shape=torch.Size((r1 - r0, K)), code=state.code,
                        blocksize=state.blocksize, quant_type=state.quant_type, dtype=state.dtype)
    return p, st


class ShardedLinear4bit:
    """One N-shard of a 4-bit linear + the output all-gather.  `forward(x)` with x replicated [batch, K]
    returns the full [batch, N] on every rank."""

    def __init__(self, packed: torch.Tensor, state: QuantState, group: Optional[dist.ProcessGroup] = None,
                 bias: Optional[torch.Tensor] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.N, self.K = state.shape
        self.packed, self.state = shard_quantized_weight(packed, state, self.world, self.rank)
        r0, r1 = shard_bounds(self.N, self.world, self.rank)
        self.bias = None if bias is None else bias[r0:r1].contiguous()

    def local_forward(self, x: torch.Tensor) -> torch.Tensor:
        from .autograd._functions import matmul_4bit
        return matmul_4bit(x, self.packed.t(), quant_state=self.state, bias=self.bias)

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = self.local_forward(x)
        if self.world == 1:
            return y
        return all_gather_features(y, self.world, self.group, out)

    __call__ = forward


def all_gather_features(y_local: torch.Tensor, world: int, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[batch, N/g] per rank -> [batch, N] on every rank (feature-major concatenation)."""
    y2 = y_local.reshape(-1, y_local.shape[-1]).contiguous()
    batch, per = y2.shape
    gathered = torch.empty((world, batch, per), dtype=y2.dtype, device=y2.device) if out is None else out
    dist.all_gather_into_tensor(gathered.view(world * batch, per), y2, group=group)
    return gathered.permute(1, 0, 2).reshape(*y_local.shape[:-1], world * per)


class PeerOutputBuffers:
    """Symmetric-memory output vectors of an N-sharded GEMV stack (batch 1) on the GPUs of one NVLink box.

    Every rank holds the FULL [1, N_i] output of each linear i in one symmetric allocation
    (torch.distributed._symmetric_memory: cuMem + peer mappings over NVLink).  Rank r's GEMV writes its slice
    [r*N_i/g, (r+1)*N_i/g) into its own copy AND -- through the peer-mapped addresses handed to the kernel -- into
    every peer's copy: the output all-gather of SURVEY 8e happens in the GEMV epilogue, no NCCL launch.
    `barrier()` (symmetric-memory signal-pad barrier, one small kernel) is what a consumer of the gathered
    vectors waits on."""

    def __init__(self, sizes: List[int], dtype: torch.dtype, device: torch.device, group=None, batch: int = 1):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.sizes = list(sizes)
        self.batch = int(batch)                  # rows of every gathered output ([batch, N_i], row-major)
        self.offsets = []
        off = 0
        for n in self.sizes:
            assert n % self.world == 0
            self.offsets.append(off)
            off += (self.batch * n + 63) // 64 * 64   # keep every output 128-byte aligned
        self.buf = symm_mem.empty(max(off, 64), dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.itemsize = self.buf.element_size()
        self.base_ptrs = [int(p) for p in self.hdl.buffer_ptrs]

    def full(self, i: int) -> torch.Tensor:
        return self.buf[self.offsets[i]:self.offsets[i] + self.batch * self.sizes[i]].view(self.batch, self.sizes[i])

    def local_slice(self, i: int) -> torch.Tensor:
        """This rank's columns of output i: [batch, N_i / world], row stride N_i (contiguous when batch == 1)."""
        per = self.sizes[i] // self.world
        if self.batch == 1:
            o = self.offsets[i] + self.rank * per
            return self.buf[o:o + per].view(1, per)
        return self.full(i)[:, self.rank * per:(self.rank + 1) * per]

    def peer_ptrs(self, i: int) -> List[int]:
        per = self.sizes[i] // self.world
        o = (self.offsets[i] + self.rank * per) * self.itemsize
        return [self.base_ptrs[p] + o for p in range(self.world) if p != self.rank]

    def barrier(self, channel: int = 0) -> None:
        if getattr(self, "_fb", None) is not None:
            self.fast_barrier()
        else:
            self.hdl.barrier(channel=channel)

    # ---- barrier as a link of the PDL chain (cbnb_peer_barrier): the next GEMV's prologue overlaps the exchange
    def enable_fast_barrier(self) -> None:
        import ctypes as ct
        import torch.distributed._symmetric_memory as symm_mem

        sig = symm_mem.empty(64, dtype=torch.int32, device=self.buf.device)
        sig.zero_()
        hdl = symm_mem.rendezvous(sig, self.group)
        counter = torch.zeros(1, dtype=torch.int32, device=self.buf.device)
        torch.cuda.synchronize()
        hdl.barrier(channel=0)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        others = [p for p in range(self.world) if p != self.rank]
        mine = [ptrs[p] + 4 * (self.rank if self.rank < p else self.rank - 1) for p in others]   # my compact slot at peer p
        self._fb = (sig, hdl, counter, (ct.c_void_p * len(mine))(*mine), len(mine))

    def fast_barrier(self) -> None:
        import ctypes as ct
        sig, _, counter, arr, n = self._fb
        F.lib.cbnb_set_stream(ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        F.lib.cbnb_peer_barrier(ct.c_void_p(counter.data_ptr()), ct.c_void_p(sig.data_ptr()), arr, ct.c_int32(n))

    # ---- in-kernel ordering (no barrier launch): signal slots in symmetric memory + a per-pass epoch
    def enable_kernel_sync(self, ngroups: int) -> None:
        import ctypes as ct
        import torch.distributed._symmetric_memory as symm_mem

        self.ngroups = ngroups
        self.sig = symm_mem.empty(64, dtype=torch.int32, device=self.buf.device)
        self.sig.zero_()
        self.sig_hdl = symm_mem.rendezvous(self.sig, self.group)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.buf.device)
        self.counter = torch.zeros(1, dtype=torch.int32, device=self.buf.device)
        torch.cuda.synchronize()
        self.sig_hdl.barrier(channel=0)
        sig_ptrs = [int(p) for p in self.sig_hdl.buffer_ptrs]
        others = [p for p in range(self.world) if p != self.rank]
        self._sig_peer = []
        for p in others:   # my compact slot inside peer p's array: my index among p's peers
            mine = self.rank if self.rank < p else self.rank - 1
            self._sig_peer.append(sig_ptrs[p] + 4 * mine)
        self._ct = ct

    def sync_desc(self, gidx: int, do_signal: bool, do_wait: bool) -> "F.GemvSync":
        s = F.GemvSync()
        s.sig_local = self.sig.data_ptr()
        for k, a in enumerate(self._sig_peer):
            s.sig_peer[k] = a
        s.epoch = self.epoch.data_ptr()
        s.cta_counter = self.counter.data_ptr()
        s.gidx, s.ngroups, s.do_signal, s.do_wait = gidx, self.ngroups, int(do_signal), int(do_wait)
        return s

    def bump_epoch(self) -> None:
        """once per pass over the stack (stream-ordered, graph-capturable)"""
        import ctypes as ct
        F.lib.cbnb_set_stream(ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        F.lib.cbnb_epoch_bump(ct.c_void_p(self.epoch.data_ptr()))


def sharded_gemv_push(x: torch.Tensor, packed_shard: torch.Tensor, state_shard: QuantState, peers: PeerOutputBuffers,
                      i: int, sync=None) -> torch.Tensor:
    """Rank-local GEMV of linear i whose epilogue stores the output slice into every rank's full vector."""
    return F.gemv_4bit(x, packed_shard.t(), out=peers.local_slice(i), state=state_shard, peer_outs=peers.peer_ptrs(i),
                       sync=sync)


def sharded_gemm_push(x: torch.Tensor, packed_shard: torch.Tensor, state_shard: QuantState, peers: PeerOutputBuffers,
                      i: int, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batch > 1: rank-local fused 4-bit GEMM of linear i whose epilogue stores the [batch, N_i / world] slice into every
    rank's gathered [batch, N_i] buffer (functional.gemm_4bit with peer outputs -> cgemm_4bit_push_*)."""
    y = F.gemm_4bit(x, packed_shard.t(), state_shard, bias=bias, out=peers.local_slice(i), peer_outs=peers.peer_ptrs(i))
    if y is None:
        raise RuntimeError("sharded_gemm_push: the fused kernel does not take this shape")
    return y


def sharded_gemv_push_multi(x: torch.Tensor, packed_shards, state_shards, peers: PeerOutputBuffers, idxs) -> list:
    """The linears `idxs` of a layer that share x (q/k/v, gate/up), rank-local slices, in ONE launch whose epilogue
    stores every slice into every rank's full vectors (functional.gemv_4bit_multi with peer outputs)."""
    return F.gemv_4bit_multi(x, [p.t() for p in packed_shards], state_shards, outs=[peers.local_slice(i) for i in idxs],
                             peer_outs=[peers.peer_ptrs(i) for i in idxs])
