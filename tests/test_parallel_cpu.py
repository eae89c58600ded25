"""N > 1 path on CPU: world_size-2 gloo processes shard the work items, 'edit' them and gather the results."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pnpinversion_b200 import parallel


def test_shard_bounds_cover_everything_once():
    for n, w in ((700, 8), (700, 1), (5, 8), (16, 4), (0, 2)):
        spans = [parallel.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_sizes(700, 8) == [88, 88, 88, 88, 87, 87, 87, 87]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        items = torch.randn(n_items, 4, 8, 8, generator=g) if rank == 0 else None
        mine = parallel.scatter_items(items, n_items, (4, 8, 8), torch.float32, "cpu", dist)
        lo, hi = parallel.shard_bounds(n_items, rank, world)
        assert mine.shape[0] == hi - lo
        edited = mine * 2.0 + 1.0  # stand-in for the per-image edit: any function of the item alone
        full = parallel.gather_items(edited, n_items, dist)
        dist.barrier()
        if rank == 0:
            q.put(full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 8])
def test_two_rank_gloo_scatter_edit_gather(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=900)  # spawned interpreters import torch: minutes on a starved host
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    ref = torch.randn(n_items, 4, 8, 8, generator=g) * 2.0 + 1.0
    assert torch.equal(full, ref)


def test_edit_lanes_assign_jobs_round_robin_keep_order_and_propagate_errors():
    """Host logic of parallel.EditLanes on a CPU device (no CUDA streams): job i runs on lane i % lanes, the jobs of one
    lane run in order on that lane's editor, results come back in job order, an exception in a lane reaches the caller."""
    import threading

    import pytest

    from pnpinversion_b200.parallel import EditLanes

    made = []

    class FakeEditor:
        def __init__(self):
            self.idx = len(made)
            self.seen = []
            self.threads = set()
            made.append(self)

    lanes = EditLanes(FakeEditor, lanes=3, device="cpu")
    assert len(lanes) == 3 and [e.idx for e in lanes.editors] == [0, 1, 2]

    def job(i):
        def run(ed):
            ed.seen.append(i)
            ed.threads.add(threading.get_ident())
            return (ed.idx, i * i)
        return run

    out = lanes.run([job(i) for i in range(8)])
    assert out == [(i % 3, i * i) for i in range(8)]
    assert [e.seen for e in lanes.editors] == [[0, 3, 6], [1, 4, 7], [2, 5]]
    assert all(len(e.threads) == 1 for e in lanes.editors)  # the jobs of a lane share one host thread
    assert lanes.run([]) == []

    def boom(ed):
        raise ValueError("lane failure")

    with pytest.raises(ValueError, match="lane failure"):
        lanes.run([job(0), boom, job(2)])
