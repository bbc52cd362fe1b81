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
import torch
import torch.distributed as dist


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
    t = torch.tensor([int(n)], dtype=torch.int64, device=device)
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
