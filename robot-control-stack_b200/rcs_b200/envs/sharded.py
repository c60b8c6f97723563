"""Environment-sharded vector env: one process per GPU, contiguous blocks of environments per rank.

Environments are independent (each is its own mjData in the reference, src/sim/sim.h:75-76), so a step has no
collective inside it. The one exchange step of the path is the vectorised observation return: after every
env.step() each rank's packed observation rows ([n_local, obs_dim] float64, info flags included) are all-gathered so
that every rank holds the [N_total, obs_dim] block. The kernel writes its rows straight into this rank's slice of the
gather buffer (no staging copy), the all-gather runs in place on a side stream, and with `step_async` it overlaps
the next step's launch: buffers alternate, and a buffer is only rewritten after its gather has completed.

`ShardedVectorEnv` needs from the local env only `num_envs`, `dev`, `obs_dim`, `step_packed(action, obs_out)`,
`reset_packed(obs_out)` and `unpack(rows)`: `SimVectorEnv` on a GPU, a stand-in in the world-size-2 gloo test.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from rcs_b200.shard import shard_range


class PendingObs:
    """Observation of one sharded step whose all-gather may still be in flight on the side stream."""

    def __init__(self, env: "ShardedVectorEnv", rows: torch.Tensor, event):
        self._env, self._rows, self._event = env, rows, event

    def rows(self) -> torch.Tensor:
        """[N_total, obs_dim] packed rows; the current stream waits for the gather first."""
        if self._event is not None:
            torch.cuda.current_stream(self._rows.device).wait_event(self._event)
            self._event = None
        return self._env._strip(self._rows)

    def result(self):
        obs, info, truncated = self._env.local.unpack(self.rows())
        n = self._env.n_total
        zeros = torch.zeros(n, dtype=torch.float64, device=self._rows.device)
        return obs, zeros, torch.zeros(n, dtype=torch.bool, device=self._rows.device), truncated, info


class ShardedVectorEnv:
    def __init__(self, local_env, n_total: int, group=None, nbuf: int = 2):
        self.local, self.n_total, self.group = local_env, int(n_total), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.begin, self.end = shard_range(self.n_total, self.rank, self.world)
        assert local_env.num_envs == self.end - self.begin, "the local env must hold exactly this rank's shard"
        self.dev = local_env.dev
        self.obs_dim = local_env.obs_dim
        self.sizes = [shard_range(self.n_total, r, self.world) for r in range(self.world)]
        self.n_max = max(e - b for b, e in self.sizes)
        self.ragged = any(e - b != self.n_max for b, e in self.sizes)
        # gather buffers [world, n_max, obs_dim]; this rank's kernel writes rows [rank, :n_local]
        self._buf = [torch.zeros((self.world, self.n_max, self.obs_dim), dtype=torch.float64, device=self.dev) for _ in range(nbuf)]
        self._busy = [None] * nbuf  # event of the gather that last read / wrote each buffer
        self._turn = 0
        self._cuda = self.dev.type == "cuda"
        self._comm = torch.cuda.Stream(self.dev) if self._cuda and self.world > 1 else None
        self.num_envs = self.n_total
        self.action_space = getattr(local_env, "action_space", None)

    # ------------------------------------------------------------------ helpers
    def _strip(self, rows: torch.Tensor) -> torch.Tensor:
        if not self.ragged:
            return rows.view(self.world * self.n_max, self.obs_dim)
        return torch.cat([rows[r, : e - b] for r, (b, e) in enumerate(self.sizes)], dim=0)

    def local_action(self, action: dict) -> dict:
        """Actions given for all N_total environments are cut down to this rank's block; local ones pass through."""
        out = {}
        for k, v in action.items():
            out[k] = v[self.begin:self.end] if v.shape[0] == self.n_total and self.n_total != self.local.num_envs else v
        return out

    def _next_buffer(self):
        i = self._turn
        self._turn = (self._turn + 1) % len(self._buf)
        if self._busy[i] is not None and self._cuda:  # its previous gather must be over before the kernel rewrites it
            torch.cuda.current_stream(self.dev).wait_event(self._busy[i])
        return i, self._buf[i]

    def _gather(self, i: int, buf: torch.Tensor):
        """In-place all-gather of buf[rank] into buf, on the side stream when there is one. Returns the event (or None)."""
        if self.world == 1:
            return None
        mine = buf[self.rank]
        if self._comm is None:
            dist.all_gather_into_tensor(buf.view(-1), mine.reshape(-1), group=self.group)
            return None
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.dev))  # the step kernel has been enqueued before this point
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ready)
            dist.all_gather_into_tensor(buf.view(-1), mine.reshape(-1), group=self.group)
            done = torch.cuda.Event()
            done.record(self._comm)
        self._busy[i] = done
        return done

    # ------------------------------------------------------------------ gym API
    def reset(self, seed=None, options=None):
        i, buf = self._next_buffer()
        n = self.local.num_envs
        if n == self.n_max:
            self.local.reset_packed(obs_out=buf[self.rank])
        else:
            buf[self.rank, :n].copy_(self.local.reset_packed())
        ev = self._gather(i, buf)
        return PendingObs(self, buf, ev).result()[0], {}

    def step_async(self, action: dict) -> PendingObs:
        """Launch this rank's step and the observation all-gather; the returned handle yields the [N_total, ...]
        observation when asked (so the caller may launch the next step first)."""
        i, buf = self._next_buffer()
        n = self.local.num_envs
        if n == self.n_max:
            self.local.step_packed(self.local_action(action), obs_out=buf[self.rank])
        else:  # the smaller shard of a ragged split: its rows do not fill the slice
            buf[self.rank, :n].copy_(self.local.step_packed(self.local_action(action)))
        return PendingObs(self, buf, self._gather(i, buf))

    def step(self, action: dict):
        return self.step_async(action).result()

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self.local
