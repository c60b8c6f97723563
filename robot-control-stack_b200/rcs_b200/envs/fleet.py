"""Mixed fleets (BASELINE.json configs[4]): several robot types / scenes simulated side by side.

A kernel launch is homogeneous (one compiled scene per launch: SURVEY.md 8e "C5's mixed fleet shards by (robot type, env
range) so each kernel launch is homogeneous"), so a fleet is a set of homogeneous vector envs -- each built by the
reference-shaped creators -- that are stepped CONCURRENTLY: every group owns a CUDA stream, `step()` enqueues all groups'
fused launches before it waits for any of them, and the caller's stream then waits for all of them. On a GPU that one
group does not fill (4096 environments are a single wave of the FR3 kernel) the groups overlap; larger groups queue behind
each other with no host synchronisation in between.
"""
from __future__ import annotations

import torch


class FleetVectorEnv:
    def __init__(self, builders: dict, device: int = 0):
        """builders: name -> callable returning a vector env (e.g. lambda: SimEnvCreator()(..., num_envs=N, device=d)). Every
        env is constructed under its own stream so that its launches are issued there."""
        self.dev = torch.device("cuda", device)
        self.streams, self.envs = {}, {}
        for name, make in builders.items():
            s = torch.cuda.Stream(self.dev)
            with torch.cuda.stream(s):
                self.envs[name] = make()
            self.streams[name] = s
        torch.cuda.synchronize(self.dev)
        self.num_envs = {k: e.num_envs for k, e in self.envs.items()}

    def _fan_out(self, fn):
        cur = torch.cuda.current_stream(self.dev)
        start = torch.cuda.Event()
        start.record(cur)
        out = {}
        for name, env in self.envs.items():      # enqueue every group's launch first ...
            s = self.streams[name]
            with torch.cuda.stream(s):
                s.wait_event(start)               # the actions were produced on the caller's stream
                out[name] = fn(name, env)
        for s in self.streams.values():          # ... then make the caller's stream wait for all of them
            done = torch.cuda.Event()
            done.record(s)
            cur.wait_event(done)
        return out

    def reset(self, seed=None, options=None):
        res = self._fan_out(lambda name, env: env.reset(seed=seed, options=options))
        return {k: v[0] for k, v in res.items()}, {k: v[1] for k, v in res.items()}

    def step(self, actions: dict):
        """actions: name -> action dict of that group. Returns name -> (obs, reward, terminated, truncated, info)."""
        return self._fan_out(lambda name, env: env.step(actions[name]))

    def step_packed(self, actions: dict):
        return self._fan_out(lambda name, env: env.step_packed(actions[name]))

    def sample_actions(self):
        return {k: e.action_space.sample() for k, e in self.envs.items()}

    def close(self):
        for e in self.envs.values():
            e.close()
