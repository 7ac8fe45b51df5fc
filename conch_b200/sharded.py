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
        self._symm: dict[tuple[int, torch.dtype], tuple[torch.Tensor, object, list[int]]] = {}
        self._flags: tuple[torch.Tensor, object, list[int]] | None = None
        self._push_fast: dict[tuple[int, torch.dtype], tuple[torch.Tensor, object, torch.Tensor]] = {}

    def local(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """This rank's `[M, N/G]` slice of the output (no communication)."""
        return self._gemm(x, self.w_q_packed, self.w_s, self.w_zp, self.bits, self.bias, self.group_size, **kwargs)

    def gathered(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """`[G, M, N/G]`: shard-major layout straight out of the all-gather (no extra copy)."""
        out, work = self.gathered_async(x, **kwargs)
        if work is not None:
            work.wait()
        return out

    def gathered_async(self, x: torch.Tensor, **kwargs) -> tuple[torch.Tensor, "dist.Work | None"]:
        """Like `gathered`, but the all-gather is only ENQUEUED (on the process group's own stream): returns the
        `[G, M, N/G]` view and the pending work.  The caller may launch independent work -- e.g. the next layer's GEMM --
        before `work.wait()`, so the NVLink transfer overlaps it; the view is valid after the wait."""
        y = self.local(x, **kwargs)
        if self.world == 1:
            return y.unsqueeze(0), None
        m = y.shape[0]
        if self._gather_buf is None or self._gather_buf.shape[0] != self.world * m or self._gather_buf.dtype != y.dtype:
            # concatenation along dim 0 (what both NCCL and gloo expect), viewed as [G, M, N/G] below
            self._gather_buf = torch.empty((self.world * m, self.n_local), device=y.device, dtype=y.dtype)
        work = dist.all_gather_into_tensor(self._gather_buf, y.contiguous(), group=self.group, async_op=True)
        return self._gather_buf.view(self.world, m, self.n_local), work

    # ---- fused gather: the GEMM kernel's own epilogue stores into every rank's gather buffer (P2P over NVLink) ----------
    PUSH_MAX_ROWS = 8  # the decode kernels are what AUTO picks up to here for every shape (cb200_mpgemm_push is decode-only)

    def _symmetric(self, m: int, dtype: torch.dtype, device: torch.device, key=None):
        """`[G, M, N/G]` gather buffer in symmetric memory (every rank maps every rank's copy), allocated once per (M, dtype).
        COLLECTIVE on first use: all ranks must call with the same M and dtype."""
        key = key or (m, dtype)
        if key not in self._symm:
            import torch.distributed._symmetric_memory as symm_mem

            buf = symm_mem.empty((self.world, m, self.n_local), dtype=dtype, device=device)
            group = self.group if self.group is not None else dist.group.WORLD
            hdl = symm_mem.rendezvous(buf, group.group_name)
            slot = self.rank * m * self.n_local * buf.element_size()  # this rank's slot inside EVERY rank's buffer
            peers = [int(hdl.buffer_ptrs[p]) + slot for p in range(self.world) if p != self.rank]
            self._symm[key] = (buf, hdl, peers)
        return self._symm[key]

    def _flag_array(self, device: torch.device):
        """`[world]` uint32 completion flags in symmetric memory (slot r is written by rank r's kernels).  COLLECTIVE on first use."""
        if self._flags is None:
            import torch.distributed._symmetric_memory as symm_mem

            flags = symm_mem.empty((max(self.world, 8) + 8,), dtype=torch.int32, device=device)  # [world] counters + the wait's own
            flags.zero_()
            group = self.group if self.group is not None else dist.group.WORLD
            hdl = symm_mem.rendezvous(flags, group.group_name)
            hdl.barrier(channel=1)  # everybody's flags are zero before anybody signals
            peer_flags = [int(hdl.buffer_ptrs[p]) + 4 * self.rank for p in range(self.world) if p != self.rank]
            self._flags = (flags, hdl, peer_flags)
        return self._flags

    # ---- side streams (per device, shared by all layers): the decoupled push and the copy-engine gather run there --------
    _side: dict[int, "torch.cuda.Stream"] = {}
    _peer_streams: dict[int, list["torch.cuda.Stream"]] = {}
    _hi_prio: dict[int, "torch.cuda.Stream"] = {}

    @classmethod
    def _side_stream(cls, device: torch.device) -> "torch.cuda.Stream":
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in cls._side:
            cls._side[idx] = torch.cuda.Stream(device=device)
        return cls._side[idx]

    def pushed_async(self, x: torch.Tensor, *, signalled: bool = True, **kwargs) -> torch.Tensor:
        """The DECOUPLED form of `pushed_signalled()`: the GEMV stores only this rank's slot (the plain single-GPU kernel, at
        single-GPU speed) and a `cb200_push` kernel on a side stream then copies the slot to every peer with 16-byte P2P stores
        and raises the same completion counters.  The token's GEMV chain therefore never contains a kernel with outstanding
        NVLink stores -- a grid that wrote to peers is not complete, and the next kernel's `griddepcontrol.wait` does not
        return, before the remote stores are acknowledged -- while every GEMV's gathered output still becomes visible on its
        own (`wait()`).  `join()` makes the current stream wait for all pushes enqueued so far (end of a token / of a graph
        capture)."""
        m = x.shape[0]
        if self.world == 1:
            return self.local(x, **kwargs).unsqueeze(0)
        out_dtype = kwargs.get("output_dtype") or x.dtype
        buf, _, peers = self._symmetric(m, out_dtype, x.device)
        self._gemm(x, self.w_q_packed, self.w_s, self.w_zp, self.bits, self.bias, self.group_size, out=buf[self.rank], **kwargs)
        import ctypes

        from conch_b200 import _lib

        side = self._side_stream(x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))  # fork: the push depends on THIS GEMV only
        nbytes = m * self.n_local * buf.element_size()
        peer_arr = (ctypes.c_void_p * len(peers))(*peers)
        if signalled:
            flags, _, peer_flags = self._flag_array(x.device)
            flag_arr = (ctypes.c_void_p * len(peer_flags))(*peer_flags)
            local_flag = flags.data_ptr() + 4 * self.rank
        else:
            flag_arr, local_flag = None, None
        with torch.cuda.device(x.device):
            _lib.raise_for_status(_lib.load().cb200_push(buf[self.rank].data_ptr(), nbytes, peer_arr, len(peers), flag_arr, local_flag,
                                                         side.cuda_stream))  # fmt: skip
        return buf

    @classmethod
    def join(cls, device: torch.device) -> None:
        """The current stream waits for every side-stream push enqueued so far on `device`."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx in cls._side:
            torch.cuda.current_stream(device).wait_stream(cls._side[idx])

    def gathered_push_async(self, x: torch.Tensor, slot: int = 0, *, mode: str = "auto", max_ctas: int = 0,
                            **kwargs) -> tuple[torch.Tensor, "torch.cuda.Event | None"]:
        """Large-M all-gather that OVERLAPS the next GEMM: the GEMM writes this rank's `[M, N/G]` slice straight into its slot
        of a symmetric `[G, M, N/G]` buffer; the slice then travels to the same slot of every peer's buffer from a side stream,
        and a device-side barrier across the ranks closes the gather.  An NCCL all-gather kernel does not overlap here: it is
        enqueued behind the next GEMM's CTAs and runs in that GEMM's tail (measured: end to end = GEMM + gather).  `mode`:
          "multicast"  `cb200_push_bulk` on a HIGH-PRIORITY stream, one `multimem.st` per 16 bytes to the NVLS multicast address
                       of the buffer (the switch replicates: the sender's egress is the slice once) -- "auto" when the symmetric-
                       memory handle has a multicast pointer;
          "p2p"        the same kernel with one plain store per peer -- "auto" otherwise;
          "ce"         one `cudaMemcpyAsync` per peer on per-peer streams (copy engines, no SM at all).
        `max_ctas` bounds the SMs the push takes from the next GEMM (0: 16 for multicast, 32 for p2p).  Returns the buffer (two
        alternate: `slot` 0 / 1) and an event; `current_stream().wait_event(event)` before reading the peers' slots, and before
        the call after next on the same slot may overwrite the local one."""
        m = x.shape[0]
        out_dtype = kwargs.get("output_dtype") or x.dtype
        if self.world == 1:
            return self.local(x, **kwargs).unsqueeze(0), None
        key = (("bulk", slot, m), out_dtype)
        buf, hdl, peers = self._symmetric(m, out_dtype, x.device, key=key)
        self._gemm(x, self.w_q_packed, self.w_s, self.w_zp, self.bits, self.bias, self.group_size, out=buf[self.rank], **kwargs)
        import ctypes

        from conch_b200 import _lib

        lib = _lib.load()
        idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
        cur = torch.cuda.current_stream(x.device)
        if idx not in self._hi_prio:
            self._hi_prio[idx] = torch.cuda.Stream(device=x.device, priority=-1)
        side = self._hi_prio[idx]
        nbytes = m * self.n_local * buf.element_size()
        src = buf[self.rank].data_ptr()
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if mode == "auto":
            mode = "multicast" if mc else "p2p"
        with torch.cuda.device(x.device):
            if mode == "ce":
                streams = self._peer_streams.setdefault(idx, [])
                while len(streams) < len(peers):
                    streams.append(torch.cuda.Stream(device=x.device))
                for st, dst in zip(streams, peers):
                    st.wait_stream(cur)
                    _lib.raise_for_status(lib.cb200_peer_copy(dst, src, nbytes, st.cuda_stream))
                for st in streams[: len(peers)]:
                    side.wait_stream(st)
            else:
                side.wait_stream(cur)
                peer_arr = (ctypes.c_void_p * len(peers))(*peers)
                if mode == "multicast":
                    if not mc:
                        raise RuntimeError("this symmetric-memory allocation has no multicast address (NVLS unavailable)")
                    mc_dst = mc + self.rank * nbytes
                    _lib.raise_for_status(lib.cb200_push_bulk(src, nbytes, peer_arr, len(peers), mc_dst, max_ctas or 16, side.cuda_stream))
                elif mode == "p2p":
                    _lib.raise_for_status(lib.cb200_push_bulk(src, nbytes, peer_arr, len(peers), None, max_ctas or 32, side.cuda_stream))
                else:
                    raise ValueError(f"unknown gather mode {mode!r}")
            with torch.cuda.stream(side):
                hdl.barrier(channel=2 + slot)  # every rank's slice has landed everywhere
                done = torch.cuda.Event()
                done.record(side)
        return buf, done

    def pushed_signalled(self, x: torch.Tensor, *, after: "ShardedMixedPrecisionLinear | None" = None, **kwargs) -> torch.Tensor:
        """`pushed()` with PER-GEMM visibility and no cross-GPU barrier: when the whole `[M, N/G]` slice has reached every rank,
        the kernel raises this rank's completion flag on every GPU (`cb200_mpgemm_push_signal`).  `wait()` then blocks the
        CURRENT STREAM (not the host) until all ranks' flags of this call have arrived: the gathered `[G, M, N/G]` buffer is
        complete, with the latency of one NVLink store instead of a barrier or an NCCL collective."""
        m = x.shape[0]
        if self.world == 1:
            return self.local(x, **kwargs).unsqueeze(0)
        if m > self.PUSH_MAX_ROWS:
            raise NotImplementedError(f"pushed_signalled() is for the decode regime (M <= {self.PUSH_MAX_ROWS}); use gathered() for M={m}")
        out_dtype = kwargs.get("output_dtype") or x.dtype
        buf, _, peers = self._symmetric(m, out_dtype, x.device)
        flags, _, peer_flags = self._flag_array(x.device)
        if after is not None:
            # `x` is (derived from) `after`'s gathered output: the kernel itself waits for all ranks' completion counters of
            # `after`'s latest pushed_signalled() -- cb200_mpgemm_chain -- instead of a wait() kernel in between.  One chained
            # consumer per signalled producer (they share `after`'s wait counter).
            aflags = after._flag_array(x.device)[0]
            kwargs["x_ready"] = (aflags.data_ptr(), self.world, aflags.data_ptr() + 4 * max(self.world, 8))
        self._gemm(x, self.w_q_packed, self.w_s, self.w_zp, self.bits, self.bias, self.group_size, out=buf[self.rank],
                   peer_out_ptrs=peers, signal=(peer_flags, flags.data_ptr() + 4 * self.rank), **kwargs)  # fmt: skip
        return buf

    def wait(self) -> None:
        """Enqueue (current stream) the wait for every rank's completion counter of the latest `pushed_signalled()` call.
        Exactly ONE wait per pushed_signalled() on every rank (the counters count both); graph-capturable and replay-safe."""
        if self.world == 1 or self._flags is None:
            return
        from conch_b200 import _lib

        flags = self._flags[0]
        stream = torch.cuda.current_stream(flags.device).cuda_stream
        with torch.cuda.device(flags.device):
            expected = flags.data_ptr() + 4 * max(self.world, 8)  # the wait's own counter lives behind the [world] flags
            _lib.raise_for_status(_lib.load().cb200_wait_flags(flags.data_ptr(), self.world, expected, stream))

    def pushed(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """GEMM on the shard whose epilogue stores the `[M, N/G]` result into slot `rank` of EVERY rank's gather buffer
        (`cb200_mpgemm_push`: plain P2P stores over NVLink / NVSwitch from the decode kernel -- no all-gather kernel, no
        NCCL latency).  Returns this rank's `[G, M, N/G]` buffer; the peers' slots are complete after `sync()`, one
        cross-GPU barrier that can cover any number of pushed layers (e.g. once per token).  Re-using the buffer for the
        next token while a peer may still read it needs a second `sync()` -- the caller's protocol.  M <= 8 only
        (decode regime); larger M should use `gathered()` (bandwidth-bound: NCCL is the right tool)."""
        m = x.shape[0]
        if not kwargs:  # the decode loop's call: everything below was resolved by the first call with this (M, dtype)
            fast = self._push_fast.get((m, x.dtype))
            if fast is not None:
                self._gemm(x, self.w_q_packed, self.w_s, self.w_zp, self.bits, self.bias, self.group_size, out=fast[0], peer_out_ptrs=fast[1])
                return fast[2]
        if self.world == 1:
            return self.local(x, **kwargs).unsqueeze(0)
        if m > self.PUSH_MAX_ROWS:
            raise NotImplementedError(f"pushed() is for the decode regime (M <= {self.PUSH_MAX_ROWS}); use gathered() for M={m}")
        out_dtype = kwargs.get("output_dtype") or x.dtype
        buf, _, peers = self._symmetric(m, out_dtype, x.device)
        fast = self._push_fast.get((m, out_dtype))
        if fast is None:  # this rank's slot as a tensor and the peer pointers as a C array, made once per (M, dtype)
            import ctypes

            fast = self._push_fast[(m, out_dtype)] = (buf[self.rank], (ctypes.c_void_p * len(peers))(*peers), buf)
        self._gemm(x, self.w_q_packed, self.w_s, self.w_zp, self.bits, self.bias, self.group_size,
                   out=fast[0], peer_out_ptrs=fast[1], **kwargs)  # fmt: skip
        return buf

    def pushed_local_view(self, m: int, dtype: torch.dtype = torch.float16) -> torch.Tensor:
        """This rank's own `[M, N/G]` slot of the symmetric gather buffer used by `pushed()`."""
        return self._symm[(m, dtype)][0][self.rank]

    def sync(self) -> None:
        """Cross-GPU barrier on the current stream (device side, symmetric-memory signal pads): after it, everything the
        peers pushed before THEIR sync() is visible here."""
        if self.world == 1:
            return
        if not self._symm:
            raise RuntimeError("sync() before any pushed() call")
        next(iter(self._symm.values()))[1].barrier(channel=0)

    def __call__(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        """`[M, N]` on every rank, same values as the unsharded op."""
        g = self.gathered(x, **kwargs)
        if self.world == 1:
            return g[0]
        m = g.shape[1]
        # a fresh tensor for every M (for M = 1 the buffer itself is [1, N] in memory, but a view of it would be overwritten by the
        # next call on this layer while M > 1 returns a copy: same ownership either way)
        if m == 1:
            return g.reshape(1, self.n_total).clone()
        return g.permute(1, 0, 2).reshape(m, self.n_total)
