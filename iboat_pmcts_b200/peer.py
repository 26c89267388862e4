"""The ranks' update blocks exchanged through peer-mapped windows (``pmcts_peer_*``, csrc/pmcts_peer.cu).

Stands where the reference pushes every worker's buffer of results into the ``Manager().dict()`` all workers share
(code/solver/worker.py:348-378), for one process per GPU on one node: a put narrows the wave's histograms and stores
them into every rank's window over NVLink; ``wait`` orders the engine's stream after the arrival of all ranks' blocks
and returns them in place ([world][block] like an all-gather's output).
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _native as N


class _Window:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = dict(shape=(int(nbytes),), typestr="|u1", data=(int(ptr), False), version=2)


class PeerExchange:
    def __init__(self, eng, world, rank, block_bytes):
        self.eng, self.world, self.rank, self.block_bytes = eng, int(world), int(rank), int(block_bytes)
        self._lib = N.lib()
        self._h = C.c_void_p()
        self.handle = np.zeros(N.PEER_HANDLE_BYTES, np.uint8)
        N.check(self._lib.pmcts_peer_create(eng._h, self.world, self.rank, self.block_bytes, C.byref(self._h),
                                            self.handle.ctypes.data))
        self._views = {}

    def connect(self, handles):
        """``handles``: uint8 [world][64], the ``handle`` of every rank in rank order."""
        handles = np.ascontiguousarray(handles, np.uint8).reshape(self.world, N.PEER_HANDLE_BYTES)
        N.check(self._lib.pmcts_peer_connect(self._h, handles.ctypes.data))

    def connect_over(self, dist, group=None):
        """Exchanges the handles with one ``torch.distributed`` all-gather and maps the other ranks' windows."""
        on_device = dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", self.eng.device) if on_device else torch.device("cpu")
        mine = torch.from_numpy(self.handle.copy()).to(dev)
        out = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine, group=group)
        self.connect(torch.stack(out).cpu().numpy())

    def put_bins(self, bins, width):
        """uint32 counters (an int32 tensor) narrowed to ``width`` bytes each, into every rank's window."""
        N.check(self._lib.pmcts_peer_put_bins(self.eng._h, self._h, bins.data_ptr(), int(bins.numel()), int(width)))

    def put(self, block):
        """A device tensor's bytes as they are."""
        N.check(self._lib.pmcts_peer_put(self.eng._h, self._h, block.data_ptr(), int(block.numel() * block.element_size())))

    def wait(self):
        """The engine's stream waits for every rank's block of the last put; returns them as a uint8 tensor
        [world * block_bytes] that lives in the window (valid until the put after next)."""
        ptr = C.c_void_p()
        N.check(self._lib.pmcts_peer_wait(self.eng._h, self._h, C.byref(ptr)))
        if ptr.value not in self._views:
            self._views[ptr.value] = torch.as_tensor(_Window(ptr.value, self.world * self.block_bytes),
                                                     device=torch.device("cuda", self.eng.device))
        return self._views[ptr.value]

    def poll(self):
        n = C.c_int32(0)
        N.check(self._lib.pmcts_peer_poll(self.eng._h, self._h, C.byref(n)))
        return int(n.value)

    def verify(self, dist=None, group=None, deadline_s=20.0):
        """Bring-up check before anything waits on a stream: every rank puts a block of its own pattern, the HOST polls
        (with a deadline) until all of them have arrived here, then the blocks are compared.  True on every rank or
        on none (one all-reduce)."""
        n = self.block_bytes
        dev = torch.device("cuda", self.eng.device)
        self.eng.use_torch_stream()  # (the pattern is written on torch's current stream: the put follows it there)
        pattern = ((torch.arange(n, device=dev, dtype=torch.int64) * 7 + 31 * self.rank + 5) % 251).to(torch.uint8)
        self.put(pattern)
        self.eng.synchronize()
        ok, t0 = False, time.time()
        while time.time() - t0 < deadline_s:
            if self.poll() == self.world:
                ok = True
                break
            time.sleep(0.002)
        if ok:
            got = self.wait().view(self.world, n)  # (every block is here: the waits pass at once)
            self.eng.synchronize()
            for r in range(self.world):
                want = ((torch.arange(n, device=dev, dtype=torch.int64) * 7 + 31 * r + 5) % 251).to(torch.uint8)
                ok = ok and bool(torch.equal(got[r], want))
        if dist is not None and self.world > 1:
            on_device = dist.get_backend(group) == "nccl"
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev if on_device else "cpu")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            ok = bool(flag.item())
        return ok

    def close(self):
        if self._h:
            self._lib.pmcts_peer_destroy(self._h)
            self._h = C.c_void_p()


def open_exchange(eng, dist, world, rank, block_bytes, group=None):
    """A verified ``PeerExchange`` over the process group, or ``(None, reason)`` on EVERY rank when any rank cannot
    allocate, map or reach the windows (no peer access between the devices, IPC closed in the container ...): the
    caller then keeps the library all-gather (``pmcts_allgather_updates``)."""
    on_device = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", eng.device) if on_device else torch.device("cpu")

    def all_ok(ok):
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return bool(flag.item())

    px, why = None, ""
    try:
        px = PeerExchange(eng, world, rank, block_bytes)
    except Exception as e:  # noqa: BLE001 -- reported, and the other ranks are told
        why = str(e)
    mine = torch.from_numpy(px.handle.copy() if px else np.zeros(N.PEER_HANDLE_BYTES, np.uint8)).to(dev)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    if not all_ok(px is not None):
        if px:
            px.close()
        return None, why or "another rank could not allocate its window"
    try:
        px.connect(torch.stack(out).cpu().numpy())
        ok = True
    except Exception as e:  # noqa: BLE001
        ok, why = False, str(e)
    if not all_ok(ok):
        dist.barrier(group=group)
        px.close()
        return None, why or "another rank could not map the windows"
    if not px.verify(dist, group):
        dist.barrier(group=group)
        px.close()
        return None, "blocks put into the windows did not arrive within the deadline"
    return px, ""
