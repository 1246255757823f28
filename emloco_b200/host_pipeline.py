"""The rollout with HOST buffers on both sides - the `rl_device = cpu` data flow of the reference (`RLGPUEnv.step`,
pacer/pacer/run.py:148-160 -> `VecTaskPython.step`, env/tasks/vec_task.py:125-134: observations, rewards and dones are moved
`.to(self.rl_device)` every step, and the agent hands observations back to the model) - software-pipelined.

A single env group makes the step a serial chain: host->device copy of the step's inputs, the step itself, device->host copy
of its results, and only then can the host issue the next step (the next step's input IS what came back).  PCIe is full
duplex and the copy engines are independent of the SMs, so the envs are split into `groups` independent groups (envs never
interact: own collision group, no inter-env observations), each with its own `Rollout`, CUDA stream, copy stream
and pinned buffers: while group A computes, group B's results travel to the host and B's next inputs travel to the device.
Causality per group is kept - step k+1 of a group is issued only after the host holds step k's results of that group.

    pipe = HostRolloutPipeline(4096, groups=2)
    pipe.submit(g)         # enqueue H2D(obs, noise) -> next step -> D2H(next obs, rew, reset | actions, neglogp, values) for group g
    pipe.wait(g)           # host blocks until group g's results are in pipe.host[g]; obs buffers swap roles (no host memcpy)

Pinned buffers are allocated after `dist.bind_to_gpu_numa_node` so that they live on the GPU's NUMA node.
"""
from __future__ import annotations

import torch

from .policy import ACTIONS, OBS
from .rollout import Rollout


class HostRolloutPipeline:
    def __init__(self, num_envs, groups=2, device=0, seed=0, horizon=32, graphs=True, **rollout_kw):
        if num_envs % groups:
            raise ValueError("num_envs must be divisible by the number of groups")
        self.N, self.G, self.n, self.T, self.graphs = int(num_envs), int(groups), int(num_envs) // int(groups), int(horizon), bool(graphs)
        dev = torch.device("cuda", device)
        torch.cuda.set_device(dev)
        self.R = [Rollout(self.n, device=device, seed=seed + 7919 * g, horizon=horizon, **rollout_kw) for g in range(self.G)]
        self.stream = [torch.cuda.Stream(device=dev) for _ in range(self.G)]
        self.copy = [torch.cuda.Stream(device=dev) for _ in range(self.G)]
        self.env_done = [torch.cuda.Event() for _ in range(self.G)]
        self.done = [torch.cuda.Event() for _ in range(self.G)]
        pin = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt).pin_memory()
        n = self.n
        self.host = [dict(obs_in=pin(n, OBS), noise=pin(n, ACTIONS).normal_(), obs=pin(n, OBS), rew=pin(n), reset=pin(n, dt=torch.int64),
                          actions=pin(n, ACTIONS), neglogp=pin(n), values=pin(n, 1)) for _ in range(self.G)]
        for g, R in enumerate(self.R):
            self.host[g]["obs_in"].copy_(R.sim.obs)
        torch.cuda.synchronize()
        self.next = [0] * self.G            # next horizon slot of every group

    @property
    def h2d_bytes_per_step(self):
        return sum(h["obs_in"].numel() * 4 + h["noise"].numel() * 4 for h in self.host)

    @property
    def d2h_bytes_per_step(self):
        return sum(sum(h[k].numel() * h[k].element_size() for k in ("obs", "rew", "reset", "actions", "neglogp", "values")) for h in self.host)

    def warm(self):
        """Eager steps (one-time initialisations), then one full horizon through the graphed path so every slot is captured."""
        for g, R in enumerate(self.R):
            with torch.cuda.stream(self.stream[g]):
                for k in range(3):
                    R.step(k)
                R.finish()
        torch.cuda.synchronize()
        for k in range(self.T if self.graphs else 3):
            for g in range(self.G):
                self.submit(g, k)
            for g in range(self.G):
                self.wait(g)

    def submit(self, g, n=None):
        R, h, st, cp = self.R[g], self.host[g], self.stream[g], self.copy[g]
        n = (self.next[g] if n is None else n) % self.T
        self.next[g] = n + 1

        def read_back_env():
            # the env step is done: next obs / reward / dones travel back on the group's copy stream while critic,
            # discriminator and the bookkeeping of the same step still run
            self.env_done[g].record()
            with torch.cuda.stream(cp):
                cp.wait_event(self.env_done[g])
                h["obs"].copy_(R.sim.obs, non_blocking=True); h["rew"].copy_(R.sim.rew, non_blocking=True)
                h["reset"].copy_(R.sim.reset, non_blocking=True)

        with torch.cuda.stream(st):
            R.sim.obs.copy_(h["obs_in"], non_blocking=True)        # the policy reads the observations the host handed over
            R.noise.copy_(h["noise"], non_blocking=True)
            if self.graphs and R._warmed:      # a Rollout's very first step runs eagerly (one-time initialisations)
                R.step_graphed_host_noise(n, after_env_step=read_back_env)
            else:
                R.step(n, noise=R.noise, host_obs=True)
                read_back_env()
            if n == self.T - 1:
                (R.finish_graphed if self.graphs else R.finish)()
            h["actions"].copy_(R.mb["actions"][n], non_blocking=True)
            h["neglogp"].copy_(R.mb["neglogpacs"][n], non_blocking=True); h["values"].copy_(R.mb["values"][n], non_blocking=True)
            st.wait_stream(cp)
            self.done[g].record()

    def wait(self, g):
        """Blocks until group g's step results are in host memory; what came back becomes the next step's input."""
        self.done[g].synchronize()
        h = self.host[g]
        h["obs_in"], h["obs"] = h["obs"], h["obs_in"]
        return h

    def run(self, steps):
        """`steps` control steps of all groups, pipelined: every group always has its next step queued while the host waits
        for another group's results."""
        if steps <= 0:
            return
        for g in range(self.G):
            self.submit(g)
        for i in range(1, steps):
            for g in range(self.G):
                self.wait(g)
                self.submit(g)
        for g in range(self.G):
            self.wait(g)

    def close(self):
        for R in self.R:
            R.close()
