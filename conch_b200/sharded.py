"""Column-(N-)sharded mixed-precision GEMM across the GPUs of one node (BASELINE config 5).

The reference has no distributed code at all; this is the north_star's multi-GPU mode:
rank r of G owns output columns [r*N/G, (r+1)*N/G) -- i.e. contiguous copies of `w_q_packed[:, slice]`, `w_s[:, slice]`,
`w_zp[:, slice]` made ONCE -- `x` is replicated, every rank runs the single-GPU op on its shard and the `[M, N/G]`
outputs are all-gathered (`torch.distributed.all_gather_into_tensor`, NCCL over NVLink 5 / NVSwitch).  K is never split,
so the result is bit-identical to the single-GPU op.
"""

from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def shard_columns(t: torch.Tensor | None, rank: int, world: int) -> torch.Tensor | None:
    """Contiguous copy of this rank's column slice of a `[rows, N]` tensor (scalar zero-points are replicated)."""
    if t is None or t.numel() == 1:
        return t
    n = t.shape[1]
    if n % world or (n // world) % 64:
        raise ValueError(f"N={n} cannot be split into {world} shards that are multiples of 64 columns")
    w = n // world
    return t[:, rank * w : (rank + 1) * w].contiguous()


class ShardedMixedPrecisionLinear:
    """Holds one rank's shard of a quantised weight and computes `x @ dequant(W)` for the FULL N.

    `gemm_fn` defaults to the B200 op; tests inject a CPU function to exercise the plumbing under gloo.
    """

    def __init__(
        self,
        w_q_packed: torch.Tensor,
        w_s: torch.Tensor,
        w_zp: torch.Tensor | None,
        weight_size_bits: int,
        weight_bias: int,
        group_size: int,
        *,
        group: dist.ProcessGroup | None = None,
        already_sharded: bool = False,
        gemm_fn: Callable[..., torch.Tensor] | None = None,
    ) -> None:
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if already_sharded:
            self.w_q_packed, self.w_s, self.w_zp = w_q_packed, w_s, w_zp
        else:
            self.w_q_packed = shard_columns(w_q_packed, self.rank, self.world)
            self.w_s = shard_columns(w_s, self.rank, self.world)
            self.w_zp = shard_columns(w_zp, self.rank, self.world)
        self.bits, self.bias, self.group_size = weight_size_bits, weight_bias, group_size
        self.n_local = self.w_q_packed.shape[1]
        self.n_total = self.n_local * self.world
        if gemm_fn is None:
            from conch_b200.ops.quantization.gemm import mixed_precision_gemm

            gemm_fn = mixed_precision_gemm
        self._gemm = gemm_fn
        self._gather_buf: torch.Tensor | None = None
        self._local_buf: torch.Tensor | None = None

    def local(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """This rank's `[M, N/G]` slice of the output (no communication)."""
        return self._gemm(x, self.w_q_packed, self.w_s, self.w_zp, self.bits, self.bias, self.group_size, **kwargs)

    def gathered(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """`[G, M, N/G]`: shard-major layout straight out of the all-gather (no extra copy)."""
        y = self.local(x, **kwargs)
        if self.world == 1:
            return y.unsqueeze(0)
        m = y.shape[0]
        if self._gather_buf is None or self._gather_buf.shape[0] != self.world * m or self._gather_buf.dtype != y.dtype:
            # concatenation along dim 0 (what both NCCL and gloo expect), viewed as [G, M, N/G] below
            self._gather_buf = torch.empty((self.world * m, self.n_local), device=y.device, dtype=y.dtype)
        dist.all_gather_into_tensor(self._gather_buf, y.contiguous(), group=self.group)
        return self._gather_buf.view(self.world, m, self.n_local)

    def __call__(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """`[M, N]` on every rank, same values as the unsharded op."""
        g = self.gathered(x, **kwargs)
        if self.world == 1:
            return g[0]
        m = g.shape[1]
        if m == 1:
            return g.reshape(1, self.n_total)  # [G, 1, N/G] is already [1, N] in memory
        return g.permute(1, 0, 2).reshape(m, self.n_total)
