"""CPU, world_size 2, gloo: the host side of the scene-sharded paths -- the record all-gather of the sort fallback
(dist.gather_records) reproduces the single-process records and their AP; the handle exchange of the symmetric buffers,
the count all-reduce and the scan sharding of the pseudo-label sweep.  (The device-side exchange itself -- push / flag /
merge kernels -- is tested on the GPU with virtual ranks: tests/test_gpu_parity.py::test_ap_exchange_virtual_ranks.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _host_ap(score, tp, npos):
    v = score > -np.inf
    o = np.argsort(-score[v], kind="stable")
    t = tp[v][o].astype(np.float64)
    tpc, fpc = np.cumsum(t), np.cumsum(1 - t)
    rec = tpc / npos if npos > 0 else np.zeros_like(tpc)
    return oracle.voc_ap(rec, tpc / np.maximum(tpc + fpc, np.finfo(np.float64).eps))


def _make(seed, C, n):
    g = torch.Generator().manual_seed(seed)
    score = (torch.randperm(C * n, generator=g).float().reshape(C, n) + 0.5) / (C * n)
    score[torch.rand((C, n), generator=g) < 0.25] = float("-inf")
    tp = (torch.rand((C, n), generator=g) < 0.1).to(torch.uint8)
    tp[score == float("-inf")] = 0
    return score, tp


def _worker(rank, world, port, q):
    try:
        _worker_impl(rank, world, port, q)
    except Exception as e:  # surface worker failures instead of a queue timeout
        if rank == 0:
            q.put(e)
        raise


def _worker_impl(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ovdet_b200  # noqa: F401
    from ovdet_b200 import dist as D
    C, n_total = 5, 1000
    score, tp = _make(0, C, n_total)
    lo, hi = D.shard_range(n_total)
    if rank == 1:
        lo -= 137  # make the shards unequal on purpose (363 / 637)
    if rank == 0:
        hi -= 137
    npos_local = tp[:, lo:hi].sum(1).to(torch.int64) + 2
    gs, gt, npos = D.gather_records(score[:, lo:hi].contiguous(), tp[:, lo:hi].contiguous(), npos_local)
    # the plumbing of the device-side exchange that runs on the host: the IPC handles of the symmetric buffers travel as
    # one small byte string per rank (dist.exchange_bytes), kept-box counts of the sharded sweeps as one all-reduce
    handles = D.exchange_bytes(bytes([rank + 1]) * 64)
    assert handles == [bytes([1]) * 64, bytes([2]) * 64]
    assert D.all_reduce_count(10 + rank, torch.device("cpu")) == 21
    # LabelFormatter.save(distributed=True): scans cut across the ranks, counts summed (label_formatter.py:169-174)
    from ovdet_b200.utils.label_formatter import LabelFormatter
    fmt = LabelFormatter("", "", "", ["scan%d" % i for i in range(7)])
    seen = []
    fmt.gen_pseudo = lambda i: (seen.append(i), i + 1)[1]        # scan i "acquires" i + 1 boxes
    total = fmt.save(distributed=True)
    assert total == sum(range(1, 8)) and seen == list(range(*D.shard_range(7)))
    if rank == 0:
        q.put((gs.numpy(), gt.numpy(), npos.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_records_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    if isinstance(res, Exception):
        raise res
    gs, gt, npos = res
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    C, n_total = 5, 1000
    score, tp = _make(0, C, n_total)
    score, tp = score.numpy(), tp.numpy()
    want_npos = tp[:, :363].sum(1) + 2 + tp[:, 363:].sum(1) + 2
    np.testing.assert_array_equal(npos, want_npos)
    for c in range(C):
        a = sorted(zip(gs[c][gs[c] > -np.inf].tolist(), gt[c][gs[c] > -np.inf].tolist()))
        b = sorted(zip(score[c][score[c] > -np.inf].tolist(), tp[c][score[c] > -np.inf].tolist()))
        assert a == b  # same multiset of (score, tp) records, padding ignored
        assert _host_ap(gs[c], gt[c], npos[c]) == _host_ap(score[c], tp[c], npos[c])


def test_shard_range_covers_everything():
    import ovdet_b200  # noqa: F401
    from ovdet_b200.dist import shard_range
    for n in (0, 1, 7, 5050, 100000):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
