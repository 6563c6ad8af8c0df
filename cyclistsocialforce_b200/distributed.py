"""Agent-range sharding of one crowd over several GPUs (one process per GPU).

Rank r owns agents [lo_r, hi_r) -- state, destination queues and the per-agent kernels
are local -- and evaluates its targets against *all* sources.  The one exchange step per
simulation step is an all-gather of the 16 B/agent pair payload (x, y, cos psi, sin psi)
through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
Independent scenarios need no communication at all: give every rank its own ``Engine``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous, balanced agent ranges: [(lo, hi)] per rank."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def balanced_bounds(weights, world):
    """Contiguous agent ranges of (nearly) equal total weight: [(lo, hi)] per rank.  With the agents
    numbered along a space-filling curve and ``weights`` = estimated pair work per agent this is a
    load-balanced spatial decomposition (ranks at the rim of a crowd have fewer neighbours per agent
    and get more agents)."""
    w = np.asarray(weights, float)
    n = w.shape[0]
    c = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(c, c[-1] * r / world))
        cuts.append(min(max(k, cuts[-1] + 1), n - (world - r)))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def neighbour_work(x, y, radius, cell=None, overhead=0.1):
    """Estimated pair work per agent: the number of road users within ``radius`` (cell histogram +
    disc-shaped box filter), plus ``overhead`` x the mean as the per-agent fixed cost."""
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    cell = cell or radius / 5.0
    ix = np.floor((x - x.min()) / cell).astype(np.int64)
    iy = np.floor((y - y.min()) / cell).astype(np.int64)
    nx, ny = int(ix.max()) + 1, int(iy.max()) + 1
    hist = np.zeros((nx, ny))
    np.add.at(hist, (ix, iy), 1.0)
    k = int(np.ceil(radius / cell))
    acc = np.zeros_like(hist)
    for dx in range(-k, k + 1):
        for dy in range(-k, k + 1):
            if (dx * dx + dy * dy) * cell * cell > (radius + cell) ** 2:
                continue
            src = hist[max(0, -dx):nx - max(0, dx), max(0, -dy):ny - max(0, dy)]
            acc[max(0, dx):nx - max(0, -dx), max(0, dy):ny - max(0, -dy)] += src
    w = acc[ix, iy]
    return w + overhead * w.mean()


class PayloadExchange:
    """all-gather of the pair payload; handles unequal shard sizes."""

    def __init__(self, n_global, rank, world, group=None, bounds=None):
        self.bounds = list(bounds) if bounds is not None else shard_bounds(n_global, world)
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = self.bounds[rank]
        self.equal = len({hi - lo for lo, hi in self.bounds}) == 1
        self.calls = 0
        self._slab = self._mine = None

    def __call__(self, payload):
        """payload: (n_global, 4) tensor whose rows [lo, hi) were just written by this rank."""
        if self.world == 1:
            return
        self.calls += 1
        if self.equal:
            dist.all_gather_into_tensor(payload, payload[self.lo:self.hi], group=self.group)
        else:
            # ragged shards: gather max-sized slabs, then copy the valid rows out
            m = max(hi - lo for lo, hi in self.bounds)
            if self._slab is None or self._slab.device != payload.device or self._slab.dtype != payload.dtype:
                self._slab = payload.new_zeros((self.world, m, payload.shape[1]))
                self._mine = payload.new_zeros((m, payload.shape[1]))
            self._mine[:self.hi - self.lo] = payload[self.lo:self.hi]
            dist.all_gather_into_tensor(self._slab.view(-1, payload.shape[1]), self._mine, group=self.group)
            for r, (lo, hi) in enumerate(self.bounds):
                if r != self.rank:
                    payload[lo:hi] = self._slab[r, :hi - lo]


class _RawDeviceArray:
    """``__cuda_array_interface__`` wrapper: lets torch view memory this package allocated itself."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False),
                                             version=2)


class PeerExchange:
    """The same exchange over NVLink peer memory, no collective call on the step path
    (csrc/csf_peer.cu): every rank owns one IPC-shared buffer [payload | data flags | read flags |
    sequence words]; after K2/K3 a rank stores its new payload range straight into every peer's
    buffer and raises a flag there; the next step's first kernel waits for all flags.  The three
    kernels (wait / signal-read / push) have fixed arguments, so they are captured into the step's
    CUDA graph.  One node only (CUDA IPC); ``torch.distributed`` is used once, to swap the handles."""

    capturable = True
    fused = True        # the Engine's fused step kernels wait / signal / push themselves (``comm``)

    def __init__(self, n_global, rank, world, dtype, device, group=None, bounds=None):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        self._lib_mod = _lib
        if world > _lib.CSF_MAX_PEERS:
            raise ValueError(f"PeerExchange supports at most {_lib.CSF_MAX_PEERS} ranks")
        self.bounds = list(bounds) if bounds is not None else shard_bounds(n_global, world)
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = self.bounds[rank]
        self.device = torch.device(device)
        self.f32 = dtype == torch.float32
        self.elem_bytes = 16 if self.f32 else 32
        self.calls = 0
        pay = (n_global * self.elem_bytes + 255) // 256 * 256
        self.off_data, self.off_read, self.off_seq = pay, pay + 256, pay + 512
        self.bytes = pay + 1024
        base = C.c_void_p()
        hb = self.lib.csf_peer_handle_bytes()
        handle = (C.c_ubyte * hb)()
        # every step below is collective: a failure on one rank (no peer access, IPC not permitted ...) is
        # agreed on by all ranks before anybody raises, so that callers can fall back to PayloadExchange
        err = None
        self.base = None
        try:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.csf_peer_alloc(self.bytes, C.byref(base), handle), "csf_peer_alloc")
            self.base = int(base.value)
        except Exception as e:          # noqa: BLE001
            err = e
        self.peers = [None] * world
        self.peers[rank] = self.base
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle) if err is None else None, group=group)
            if err is None and any(h is None for h in handles):
                err = RuntimeError("PeerExchange: a peer could not allocate its buffer")
            if err is None:
                try:
                    with torch.cuda.device(self.device):
                        for p in range(world):
                            if p == rank:
                                continue
                            ptr = C.c_void_p()
                            buf = (C.c_ubyte * hb).from_buffer_copy(handles[p])
                            _lib.check(self.lib.csf_peer_open(buf, C.byref(ptr)), "csf_peer_open")
                            self.peers[p] = int(ptr.value)
                except Exception as e:  # noqa: BLE001
                    err = e
            flags = [None] * world
            dist.all_gather_object(flags, err is None, group=group)
            if not all(flags):
                for p, ptr in enumerate(self.peers):
                    if p != rank and ptr:
                        self.lib.csf_peer_close(C.c_void_p(ptr))
                if self.base:
                    self.lib.csf_peer_free(C.c_void_p(self.base))
                self.base = None
                raise RuntimeError(f"PeerExchange unavailable on this node: {err or 'a peer failed'}")
        elif err is not None:
            raise err
        comm = _lib.CsfPeerComm()
        comm.world, comm.rank = world, rank
        for p in range(world):
            comm.payload[p] = self.peers[p]
            comm.data_flags[p] = self.peers[p] + self.off_data
            comm.read_flags[p] = self.peers[p] + self.off_read
        comm.seq = self.base + self.off_seq
        # host-mapped mirror of the status word: Engine.step() polls it without a copy
        self.status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        comm.status_host = self.status_host.data_ptr()
        self.comm = comm
        self._C = C
        shape, typestr = ((n_global, 4), "<i4") if self.f32 else ((n_global, 4), "<f8")
        self._raw = _RawDeviceArray(self.base, shape, typestr)
        self.payload = torch.as_tensor(self._raw, device=self.device)
        self._seq = torch.as_tensor(_RawDeviceArray(self.base + self.off_seq, (4,), "<i4"), device=self.device)

    def payload_tensor(self):
        """The (n_global, 4) payload array inside the shared buffer: the Engine uses it as its payload."""
        return self.payload

    def _st(self):
        return self._C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def begin_step(self):
        self._lib_mod.check(self.lib.csf_peer_wait_data(self._C.byref(self.comm), self._st()), "csf_peer_wait_data")

    def after_pair(self):
        self._lib_mod.check(self.lib.csf_peer_signal_read(self._C.byref(self.comm), self._st()),
                            "csf_peer_signal_read")

    def __call__(self, payload):
        """One stand-alone exchange epoch for a payload range written by something else than the fused
        step (Engine.pack): signal "done reading", then push."""
        if self.world == 1:
            return
        assert payload.data_ptr() == self.base, "PeerExchange: the engine must use payload_tensor() as its payload"
        self.calls += 1
        self.after_pair()
        self._lib_mod.check(self.lib.csf_peer_push(self._C.byref(self.comm), self.lo, self.hi - self.lo,
                                                   self.elem_bytes, self._st()), "csf_peer_push")

    def poll_status(self):
        return bool(self.status_host[0] != 0)

    def check_status(self):
        st = int(self._seq[3].item())
        if st:
            raise RuntimeError(f"PeerExchange: a wait on a peer flag timed out (status {st})")

    def close(self):
        if self.base is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        for p, ptr in enumerate(self.peers):
            if p != self.rank and ptr:
                self.lib.csf_peer_close(self._C.c_void_p(ptr))
        if self.world > 1:
            dist.barrier(group=self.group)
        self.payload = self._seq = self._raw = None
        self.lib.csf_peer_free(self._C.c_void_p(self.base))
        self.base = None


def gather_rows_host(local_rows: np.ndarray, n_global: int, rank: int, world: int, group=None):
    """Host-side helper for tests: concatenate per-rank numpy rows on every rank."""
    if world == 1:
        return local_rows
    objs = [None] * world
    dist.all_gather_object(objs, local_rows, group=group)
    return np.concatenate(objs, axis=0)
