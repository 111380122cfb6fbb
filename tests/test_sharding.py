"""N>1 host logic on CPU: contiguous sharding + the single all-gather of the final box list,
world_size 2 over gloo."""
import importlib
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module(PKG_NAME + ".sharding")
    lo, hi = sh.shard_range(n_total, rank, world)
    # a fake per-image detection list whose content encodes the global image index
    dets = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1, 1).expand(hi - lo, 100, 6).contiguous()
    out = sh.gather_detections(dets, n_total=n_total)
    out2 = sh.gather_detections(dets)  # sizes discovered by a collective
    counts = torch.arange(lo, hi, dtype=torch.int32)
    d3, c3 = sh.gather_variable(dets, counts)
    q.put((rank, out[:, 0, 0].tolist(), out2[:, 0, 0].tolist(), c3.tolist(), tuple(d3.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_two_rank_gather(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [float(i) for i in range(n_total)]
    for rank, a, b, c, shp in res:
        assert a == want and b == want, (rank, a, b)
        assert c == list(range(n_total)) and shp == (n_total, 100, 6)


def test_shard_range_covers_everything():
    sh = importlib.import_module(PKG_NAME + ".sharding")
    for n in (0, 1, 7, 32, 256, 1024):
        for world in (1, 2, 4, 8):
            spans = [sh.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def test_bind_host_to_gpu_is_optional():
    """bench.py pins every rank to its GPU's NUMA node before allocating pinned buffers; without NVML / without a GPU / inside a
    cpuset that forbids it the helper reports (None, None) (or an unchanged set) and never raises or shrinks the set to nothing."""
    import os
    from conftest import PKG_NAME
    import importlib
    sh = importlib.import_module(PKG_NAME + ".sharding")
    before = os.sched_getaffinity(0)
    old, new = sh.bind_host_to_gpu(0)
    try:
        assert (old is None and new is None) or (len(new) >= 1 and new <= old == before)
    finally:
        os.sched_setaffinity(0, before)
