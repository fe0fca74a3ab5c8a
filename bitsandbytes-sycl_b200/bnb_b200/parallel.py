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
