"""World-size-2 gloo tests (CPU) of the N>1 host logic: shard ranges tile the env ids, the per-step observation
all-gather reassembles the env-major block on every rank, and ShardedVectorEnv (the product's multi-GPU env) cuts global
actions down to its block, steps it and returns the gathered observation -- even and ragged splits."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401  (sys.path)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class _FakeLocalEnv:
    """Stands in for SimVectorEnv on a machine without a GPU: row e of the packed observation is
    [global env id, step count, sum of the action row, 0 ...]."""

    def __init__(self, begin, end, obs_dim=30):
        self.begin, self.num_envs, self.obs_dim, self.dev = begin, end - begin, obs_dim, torch.device("cpu")
        self.t = 0

    def _rows(self, act_sum, obs_out):
        o = obs_out if obs_out is not None else torch.zeros((self.num_envs, self.obs_dim), dtype=torch.float64)
        o.zero_()
        o[:, 0] = torch.arange(self.begin, self.begin + self.num_envs, dtype=torch.float64)
        o[:, 1] = self.t
        o[:, 2] = act_sum
        return o

    def reset_packed(self, obs_out=None):
        self.t = 0
        return self._rows(0.0, obs_out)

    def step_packed(self, action, obs_out=None):
        assert action["joints"].shape[0] == self.num_envs
        self.t += 1
        return self._rows(action["joints"].sum(dim=1), obs_out)

    def unpack(self, o):
        return {"id": o[:, 0], "t": o[:, 1], "a": o[:, 2]}, {}, o[:, 26] != 0


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rcs_b200.envs.sharded import ShardedVectorEnv
    from rcs_b200.shard import gather_observations, shard_range
    b, e = shard_range(n_total, rank, world)
    local = torch.arange(b, e, dtype=torch.float64).unsqueeze(1).repeat(1, 30)
    full = gather_observations(local, n_total)
    ok = bool(torch.equal(full[:, 0], torch.arange(n_total, dtype=torch.float64))) and full.shape == (n_total, 30)
    # the product's sharded env over a stand-in local env
    env = ShardedVectorEnv(_FakeLocalEnv(b, e), n_total)
    obs, _ = env.reset()
    ok &= bool(torch.equal(obs["id"], torch.arange(n_total, dtype=torch.float64))) and float(obs["t"].max()) == 0
    act = {"joints": torch.arange(n_total, dtype=torch.float64).unsqueeze(1).repeat(1, 7)}  # global actions on every rank
    pend = [env.step_async(act) for _ in range(3)]                                          # three steps in flight, two buffers
    for t, p in enumerate(pend[1:], start=2):
        o, rew, term, trunc, info = p.result()
        ok &= bool(torch.equal(o["id"], torch.arange(n_total, dtype=torch.float64)))
        ok &= bool(torch.equal(o["a"], 7 * torch.arange(n_total, dtype=torch.float64))) and bool((o["t"] == t).all())
        ok &= rew.shape == (n_total,) and not bool(term.any())
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
