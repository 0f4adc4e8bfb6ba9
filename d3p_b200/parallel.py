"""Sharding one minibatch over the GPUs of a box (new; the reference is single-device).

One process per GPU (``torch.distributed``, NCCL over NVLink).  All ranks hold the same
``DPSVIState`` and the same batchifier state, so
  * the sampler runs redundantly on every rank (ALU-only, no traffic) and yields the bit-identical
    index list everywhere — no collective for indices;
  * rank r computes per-example gradients only for batch positions ``[r*ceil(B/G), (r+1)*ceil(B/G))``
    (per-example Threefry keys are addressed by position, so they match the single-GPU run);
  * the ``P + 2`` clipped sums (gradient, loss, valid count) are all-reduced once per step;
  * every rank draws the same ChaCha noise and applies the same optimizer step.
The sum order differs from the single-GPU run, so parameters agree to fp32 reassociation error.
"""
import torch
import torch.distributed as dist

from . import _native as _n


class PeerWindow:
    """This rank's NVLink peer-memory window (``d3p_comm_*``): the finalize kernel publishes the
    ``P + 2`` clipped sums there and reads the peers' copies, so a sharded step needs neither a
    reduce kernel nor an NCCL launch.  ``torch.distributed`` is used once, to swap the IPC handles."""

    def __init__(self, rank, world_size, max_params, group=None, max_records=0, _connect=True):
        import ctypes as C
        self.rank, self.world_size, self.group = rank, world_size, group
        self.max_records = int(max_records)
        self.max_params = int(max_params)
        self._local_group = None
        handle = (C.c_uint8 * 64)()
        self._comm = C.c_void_p()
        _n.check(_n.lib().d3p_comm_create(rank, world_size, int(max_params), self.max_records, C.byref(self._comm),
                                          handle), "comm_create")
        if world_size > 1 and _connect:
            handles = [None] * world_size
            dist.all_gather_object(handles, bytes(handle), group=group)
            blob = (C.c_uint8 * (64 * world_size)).from_buffer_copy(b"".join(handles))
            rc = _n.lib().d3p_comm_connect(self._comm, blob)
            # every rank learns whether every mapping succeeded, so that a failure raises everywhere
            # (a rank raising alone would leave the others waiting in the next collective)
            rcs = [None] * world_size
            dist.all_gather_object(rcs, int(rc), group=group)
            if any(r != _n.OK for r in rcs):
                self.close()
                _n.check(next(r for r in rcs if r != _n.OK), "comm_connect (on some rank)")

    @classmethod
    def local_group(cls, world_size, max_params, max_records=0):
        """``world_size`` windows owned by THIS process and connected by pointer (``d3p_comm_connect_local``):
        logical ranks that share one device (each driven on its own stream) or several devices of one
        process with peer access enabled.  No ``torch.distributed``.  Used by the one-GPU tests of the
        exchange protocol; the launch order across ranks is the caller's business (a rank's kernel spins
        until the peers' kernels have been launched)."""
        import ctypes as C
        wins = [cls(r, world_size, max_params, None, max_records, _connect=False) for r in range(world_size)]
        ptrs = (C.c_void_p * world_size)()
        for r, w in enumerate(wins):
            p = C.c_void_p()
            _n.check(_n.lib().d3p_comm_window(w._comm, C.byref(p), None), "comm_window")
            ptrs[r] = p.value
        if world_size > 1:
            for w in wins:
                _n.check(_n.lib().d3p_comm_connect_local(w._comm, ptrs), "comm_connect_local")
        for w in wins:
            w._local_group = wins
        return wins

    @property
    def ptr(self):
        return self._comm

    def with_records(self, n_records):
        """A window that can also carry the sharded Poisson sampler for ``n_records`` records (collective:
        every rank must call it at the same point; the old window is closed)."""
        if n_records <= self.max_records:
            return self
        if self._local_group is not None:
            raise ValueError("a local_group window must be created with max_records large enough for the sampler")
        self.check()
        if self.world_size > 1:
            dist.barrier(group=self.group)
        self.close()
        new = PeerWindow(self.rank, self.world_size, self.max_params, self.group, max_records=n_records)
        if getattr(self, "sampler_margin", None) is not None:
            new.set_sampler_margin(self.sampler_margin)
        if getattr(self, "timeout_ms", None) is not None:
            new.set_timeout_ms(self.timeout_ms)
        return new

    def timeouts(self):
        """Exchanges that timed out on this rank so far (host-mapped counter, no device synchronisation: covers
        the work that has completed).  Non-zero is fatal: the step that saw it was poisoned with NaN and every
        later call through this window raises ``D3PNativeError`` (``D3P_ERR_PEER_TIMEOUT``)."""
        import ctypes as C
        out = C.c_uint32(0)
        _n.check(_n.lib().d3p_comm_timeouts(self._comm, C.byref(out)), "comm_timeouts")
        return out.value

    def timeout_detail(self):
        """(all, clipped-sum exchange, sampler counts) time-outs seen by this rank's kernels."""
        import ctypes as C
        out = (C.c_uint32 * 4)()
        _n.check(_n.lib().d3p_comm_timeout_detail(self._comm, out), "comm_timeout_detail")
        return tuple(out)[:3]

    def set_timeout_ms(self, ms):
        """How long a kernel waits for a peer's words before it gives up (default 10 s)."""
        _n.check(_n.lib().d3p_comm_set_timeout_ms(self._comm, int(ms)), "comm_set_timeout_ms")
        self.timeout_ms = int(ms)

    def set_sampler_margin(self, tiles):
        """Sharded Poisson sampler: tiles drawn redundantly on each side of a rank's slice (same on all ranks)."""
        _n.check(_n.lib().d3p_comm_set_sampler_margin(self._comm, int(tiles)), "comm_set_sampler_margin")
        self.sampler_margin = int(tiles)

    def check(self, synchronize=True):
        """Raise on every rank's own evidence if an exchange timed out (``synchronize=True`` drains the device first)."""
        if synchronize:
            torch.cuda.synchronize()
        n = self.timeouts()
        if n:
            raise _n.D3PNativeError(f"rank {self.rank}: {n} peer-memory exchange time-out(s) {self.timeout_detail()} "
                                    "(all, clipped sums, sampler counts); this step was poisoned "
                                    "(NaN) and the replicas can no longer be trusted")

    def close(self):
        if self._comm:
            _n.lib().d3p_comm_destroy(self._comm)
            self._comm = None


def shard_dpsvi(svi, rank=None, world_size=None, group=None, backend="p2p", max_params=None, window=None):
    """Make ``svi.update`` / ``svi.run_epoch`` process only this rank's slice of every batch.

    ``backend="p2p"`` (default): the clipped sums meet inside the finalize kernel over NVLink peer
    memory (``PeerWindow``); ``backend="nccl"``: reduce kernel + ``ncclAllReduce`` of ``P + 2`` floats.

    A step that waits longer than the window's time-out for a peer is poisoned (NaN loss and parameters) and
    every later ``update`` / ``run_epoch`` raises ``D3PNativeError``; ``svi.peer_window.check()`` asks
    explicitly (call it at least once per epoch)."""
    if window is not None:          # a window built by the caller (e.g. PeerWindow.local_group)
        svi.shard = (window.rank, window.world_size, None)
        svi.peer_window = window
        return svi
    rank = dist.get_rank(group) if rank is None else rank
    world_size = dist.get_world_size(group) if world_size is None else world_size
    if backend not in ("p2p", "nccl", "auto"):
        raise ValueError("backend must be 'p2p', 'nccl' or 'auto'")
    if backend in ("p2p", "auto"):
        if max_params is None:
            if svi.family is None:
                raise ValueError("max_params is required when DPSVI has no model family")
            max_params = svi.family.n_params
        window, err = None, None
        try:
            window = PeerWindow(rank, world_size, max_params, group)
        except Exception as e:      # e.g. no peer access between the GPUs of this box
            if backend == "p2p":
                raise
            err = e
        if backend == "auto" and world_size > 1:
            # all ranks must take the same path: agree on whether every window came up
            ok = torch.tensor([1 if window is not None else 0], device=torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                if window is not None:
                    window.close()
                window = None
        if window is not None:
            svi.shard = (rank, world_size, None)
            svi.peer_window = window
            return svi
        import warnings
        warnings.warn(f"peer-memory window unavailable ({err}); falling back to the NCCL all-reduce")
    buf = {}

    def reduce_fn(ws, n_partials, P):
        key = (ws.device, P)
        if key not in buf:
            buf[key] = torch.empty(P + 2, dtype=torch.float32, device=ws.device)
        out = buf[key]
        _n.check(_n.lib().d3p_reduce_partials_f32(_n.ptr(ws), n_partials, P, _n.ptr(out), _n.stream_ptr()),
                 "reduce_partials")
        if world_size > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out, 1

    svi.shard = (rank, world_size, reduce_fn)
    return svi


def position_range(B, rank, world_size):
    per = (B + world_size - 1) // world_size
    return min(B, rank * per), min(B, (rank + 1) * per)
