"""Multi-GPU plumbing for the scene-sharded paths.

AP evaluation shards by scene: NMS, exact IoU, argmax and TP flags are per-scene
(csrc/eval.cu); only the per-class score-sorted cumulative sums are global.  The
one exchange step is an all-gather of the compact class-major (score, tp-bits)
records plus an all-reduce of ``npos`` -- 5 bytes per record instead of the
reference's all-gather of every output tensor and the whole input batch
(engine.py:207-208 via utils/dist.py:159-176).  Pseudo-label generation shards
by scene with no data-path collective.  One process per GPU, ``torch.distributed``
(NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _capi as C


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_range(n_items, rank=None, world=None):
    """Contiguous block of scenes for this rank (sizes differ by at most one)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_records(rec_score, rec_tp, npos):
    """rec_score fp32 [C,Nl], rec_tp uint8 [C,Nl] (local, Nl may differ per rank),
    npos int64 [C] -> global ([C, W*Nmax], [C, W*Nmax], npos summed); absent slots
    carry score -inf and are ignored by the reduce."""
    if not is_distributed():
        return rec_score, rec_tp, npos
    world = dist.get_world_size()
    C, nl = rec_score.shape
    dev = rec_score.device
    n = torch.tensor([nl], dtype=torch.int64, device=dev)
    dist.all_reduce(n, op=dist.ReduceOp.MAX)
    nmax = int(n.item())
    if nl < nmax:
        pad_s = torch.full((C, nmax - nl), float("-inf"), dtype=rec_score.dtype, device=dev)
        pad_t = torch.zeros((C, nmax - nl), dtype=rec_tp.dtype, device=dev)
        rec_score = torch.cat([rec_score, pad_s], 1)
        rec_tp = torch.cat([rec_tp, pad_t], 1)
    rec_score, rec_tp = rec_score.contiguous(), rec_tp.contiguous()
    gs = torch.empty((world * C, nmax), dtype=rec_score.dtype, device=dev)   # rank-major
    gt = torch.empty((world * C, nmax), dtype=rec_tp.dtype, device=dev)
    dist.all_gather_into_tensor(gs, rec_score)
    dist.all_gather_into_tensor(gt, rec_tp)
    npos = npos.clone()
    dist.all_reduce(npos, op=dist.ReduceOp.SUM)
    return (gs.view(world, C, nmax).permute(1, 0, 2).reshape(C, world * nmax).contiguous(),
            gt.view(world, C, nmax).permute(1, 0, 2).reshape(C, world * nmax).contiguous(), npos)


def all_reduce_count(n, device):
    """Sum of a per-rank integer (kept-box counts of the pseudo-label sweep, label_formatter.py:174,179)."""
    if not is_distributed():
        return int(n)
    if dist.get_backend() != "nccl":      # gloo (CPU tests): the collective runs on host tensors
        device = torch.device("cpu")
    t = torch.tensor([int(n)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


class SymmetricBuffer(object):
    """One zero-filled device allocation per rank, mapped into every peer process (CUDA IPC through the C ABI's
    ``ovdet_symm_*``), so that kernels exchange by storing into each other's memory over NVLink -- the transport of
    ``ovdet_apx_reduce``.  ``torch.distributed`` only carries the 64-byte handles, once.

    ``peer_ptrs[r]`` is rank r's buffer as mapped here (``peer_ptrs[rank]`` is the local allocation).  Creation is
    collective over ``group``."""

    def __init__(self, nbytes, group=None, device=None):
        self.nbytes = int(nbytes)
        self.group = group
        self.world = dist.get_world_size(group) if is_distributed() else 1
        self.rank = dist.get_rank(group) if is_distributed() else 0
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        L = C.lib()
        p = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            C.check(L.ovdet_symm_alloc(self.nbytes, ctypes.byref(p)))
        self.ptr = p.value
        self._opened = []
        self.peer_ptrs = [None] * self.world
        self.peer_ptrs[self.rank] = self.ptr
        if self.world > 1:
            h = (ctypes.c_ubyte * C.SYMM_HANDLE_BYTES)()
            C.check(L.ovdet_symm_export(self.ptr, h))
            handles = exchange_bytes(bytes(h), group, self.device)
            with torch.cuda.device(self.device):
                for r, hb in enumerate(handles):
                    if r == self.rank:
                        continue
                    q = ctypes.c_void_p()
                    C.check(L.ovdet_symm_open(ctypes.create_string_buffer(hb, len(hb)), ctypes.byref(q)))
                    self.peer_ptrs[r] = q.value
                    self._opened.append(q.value)
        self.peers_array = (ctypes.c_void_p * self.world)(*self.peer_ptrs)

    def close(self):
        L = C.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            if self.world > 1:   # nobody unmaps or frees while a peer may still be storing into it
                dist.barrier(self.group)
            for q in self._opened:
                L.ovdet_symm_close(q)
            self._opened = []
            if self.world > 1:
                dist.barrier(self.group)
            if self.ptr:
                L.ovdet_symm_free(self.ptr)
                self.ptr = None

    @staticmethod
    def local_ranks(nbytes, n, device=None):
        """n buffers of ONE process wired to each other (no IPC): the exchange kernels of n virtual ranks can then be
        driven on n streams of a single GPU -- how the multi-rank protocol is tested on one device."""
        bufs = []
        for r in range(n):
            b = SymmetricBuffer.__new__(SymmetricBuffer)
            b.nbytes, b.group, b.world, b.rank = int(nbytes), None, n, r
            b.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            p = ctypes.c_void_p()
            with torch.cuda.device(b.device):
                C.check(C.lib().ovdet_symm_alloc(b.nbytes, ctypes.byref(p)))
            b.ptr, b._opened = p.value, []
            bufs.append(b)
        for b in bufs:
            b.peer_ptrs = [x.ptr for x in bufs]
            b.peers_array = (ctypes.c_void_p * n)(*b.peer_ptrs)
            b.world = n
        for b in bufs:
            b.close = (lambda bb=b: (C.lib().ovdet_symm_free(bb.ptr), setattr(bb, "ptr", None)) if bb.ptr else None)
        return bufs


def exchange_bytes(payload, group=None, device=None):
    """All-gather of one small byte string per rank (IPC handles) -> list of ``bytes`` in rank order.  Goes through a
    uint8 tensor on the backend's native device (CUDA for NCCL, host for gloo)."""
    if not is_distributed():
        return [payload]
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = (torch.device("cuda", torch.cuda.current_device()) if device is None else device) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    out = torch.empty((world, len(payload)), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group) if backend == "nccl" else \
        dist.all_gather(list(out.unbind(0)), mine, group=group)
    host = out.cpu()
    return [bytes(host[r].tolist()) for r in range(world)]
