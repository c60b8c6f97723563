"""World-size-2 gloo test (CPU) of the N>1 host logic: shard ranges tile the env ids and the per-step observation
all-gather reassembles the env-major block on every rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401  (sys.path)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rcs_b200.shard import gather_observations, shard_range
    b, e = shard_range(n_total, rank, world)
    local = torch.arange(b, e, dtype=torch.float64).unsqueeze(1).repeat(1, 30) + 0.5 * rank * 0
    full = gather_observations(local, n_total)
    ok = bool(torch.equal(full[:, 0], torch.arange(n_total, dtype=torch.float64))) and full.shape == (n_total, 30)
    q.put((rank, b, e, ok))
    dist.destroy_process_group()


def _run(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    return res


def test_even_shards_all_gather():
    res = _run(64)
    assert res == [(0, 0, 32, True), (1, 32, 64, True)]


def test_ragged_shards_all_gather():
    res = _run(33)
    assert res == [(0, 0, 16, True), (1, 16, 33, True)]


def test_shard_ranges_tile():
    from rcs_b200.shard import shard_range
    for n in (1, 7, 4096, 65536):
        for w in (1, 2, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
