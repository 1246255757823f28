"""The training epoch of the reference agent: `AMPValueAgent.train_epoch` (pacer/pacer/learning/amp_continuous_value.py:180-274)
= play_steps (rollout.Rollout) -> AMP demo / replay buffers (amp_continuous.py:621-649,694-710, learning/replay_buffer.py) ->
prepare_dataset (common_agent.py:426-456,685-696; amp_continuous.py:191-201) -> mini_epochs x minibatches of calc_gradients
(update.PPOUpdate) -> replay store.  One process per GPU; the only collective is the gradient all-reduce inside PPOUpdate.

Host-side work here is what the reference also does on the host between the two hot loops: drawing permutations, gathering
minibatch rows, normalising the 131 072 advantages of an epoch.  The hot loops themselves run in the CUDA kernels.
"""
from __future__ import annotations

import torch

from . import _lib
from .policy import AMP_OBS, RunningMeanStd, normalize
from .rollout import Rollout
from .update import PPOUpdate


class ReplayBuffer:
    """learning/replay_buffer.py:3-84 (circular store, permutation-driven sampling), tensors on the device."""

    def __init__(self, buffer_size, device, generator=None):
        self._head, self._total_count, self._buffer_size, self._device = 0, 0, int(buffer_size), device
        self._data_buf, self._gen = None, generator
        self._sample_idx = torch.randperm(self._buffer_size, device=device, generator=generator)
        self._sample_head = 0

    def get_buffer_size(self):
        return self._buffer_size

    def get_total_count(self):
        return self._total_count

    def store(self, data_dict):
        if self._data_buf is None:
            self._data_buf = {k: torch.zeros((self._buffer_size,) + tuple(v.shape[1:]), device=self._device) for k, v in data_dict.items()}
        n = next(iter(data_dict.values())).shape[0]
        assert n <= self._buffer_size
        for key, buf in self._data_buf.items():
            v = data_dict[key]
            assert v.shape[0] == n
            store_n = min(n, self._buffer_size - self._head)
            buf[self._head:self._head + store_n] = v[:store_n]
            if n - store_n > 0:
                buf[0:n - store_n] = v[store_n:]
        self._head = (self._head + n) % self._buffer_size
        self._total_count += n

    def sample(self, n):
        idx = torch.arange(self._sample_head, self._sample_head + n, device=self._device) % self._buffer_size
        rand_idx = self._sample_idx[idx]
        if self._total_count < self._buffer_size:
            rand_idx = rand_idx % self._head
        out = {k: v[rand_idx] for k, v in self._data_buf.items()}
        self._sample_head += n
        if self._sample_head >= self._buffer_size:
            self._sample_idx[:] = torch.randperm(self._buffer_size, device=self._device, generator=self._gen)
            self._sample_head = 0
        return out


class AMPValueAgent:
    def __init__(self, num_envs, horizon=32, minibatch_size=16384, amp_minibatch_size=None, mini_epochs=6, amp_batch_size=None,
                 amp_obs_demo_buffer_size=200000, amp_replay_buffer_size=200000, amp_replay_keep_prob=0.01, fetch_amp_obs_demo=None,
                 normalize_advantage=True, seed=0, device=0, update_cfg=None, graphed=True, motion_lib=None, **rollout_kw):
        self.R = Rollout(num_envs, device=device, horizon=horizon, seed=seed, tensor_cores=True, **rollout_kw)
        R = self.R
        self.T, self.N = R.T, R.N
        self.batch_size = self.T * self.N
        self.minibatch_size = int(minibatch_size)
        if self.batch_size % self.minibatch_size:
            raise ValueError("batch_size (horizon * num_envs) must be divisible by minibatch_size (rl_games asserts the same)")
        self.amp_minibatch_size = int(amp_minibatch_size or minibatch_size)
        self.mini_epochs, self.graphed, self.normalize_advantage = int(mini_epochs), bool(graphed), bool(normalize_advantage)
        self.up = PPOUpdate(R.net, R.obs_norm, R.amp_norm, self.minibatch_size, self.amp_minibatch_size, cfg=update_cfg)
        self.up.adopt_into(R.nets)
        dev = R.state.device
        self.gen = torch.Generator(device=dev).manual_seed(seed + 77)
        self.amp_batch_size = int(amp_batch_size or max(self.amp_minibatch_size // 2, 1))
        self._demo = ReplayBuffer(amp_obs_demo_buffer_size, dev, self.gen)
        self._replay = ReplayBuffer(amp_replay_buffer_size, dev, self.gen)
        self._keep_prob = float(amp_replay_keep_prob)
        # demos and reset states come from the motion library (motion_lib.MotionLibSMPL: `task.fetch_amp_obs_demo(n)`,
        # humanoid_amp.py:168-220, and `_sample_ref_state`, :295-317).  Default: synthetic walking clips (the AMASS data the
        # reference loads are not redistributable); pass `motion_lib=` built from real clips, or a `fetch_amp_obs_demo` callable.
        if motion_lib is None and fetch_amp_obs_demo is None:
            from .motion_lib import MotionLibSMPL
            from .synthetic import synthetic_motion_lib
            motion_lib = MotionLibSMPL(synthetic_motion_lib(64, seed), device=device, seed=seed + 5)
        self.motion_lib = motion_lib
        self.fetch_amp_obs_demo = fetch_amp_obs_demo or (lambda n: motion_lib.fetch_amp_obs_demo(n, dt=2.0 / 60.0))
        for _ in range(-(-self._demo.get_buffer_size() // self.amp_batch_size)):        # _init_amp_demo_buf (:636-644)
            self._demo.store({"amp_obs": self.fetch_amp_obs_demo(self.amp_batch_size)})
        self.epoch = 0
        self._scratch = torch.zeros(4, device=dev, dtype=torch.float64)

    def _value_norm(self, x):
        """value_mean_std(x) in training mode (common_agent.py:440-442): normalise with the current statistics, then absorb x."""
        vn = self.R.value_norm
        mean, var = vn.f32()
        y = normalize(x, mean, var, vn.epsilon)
        self.up._rms_update(vn, x)
        return y

    def train_epoch(self):
        R, up, T, N = self.R, self.up, self.T, self.N
        if self.motion_lib is not None:
            # fresh mocap start states for the envs that reset during this horizon (one draw per env and horizon; the reference
            # draws at every reset, humanoid_amp.py:295-317 - same distribution, consumed by the device-side reset)
            self.motion_lib.sample_reset_state(R.init_root, R.init_dof)
        b = R.play_steps(graphed=self.graphed)                                          # :183-186
        fl = lambda t: t.reshape(T * N, *t.shape[2:])
        self._demo.store({"amp_obs": self.fetch_amp_obs_demo(self.amp_batch_size)})     # _update_amp_demos (:646-649)
        amp_obs = fl(b["amp_obs"])
        demo = self._demo.sample(T * N)["amp_obs"]                                      # :192-194
        replay = amp_obs if self._replay.get_total_count() == 0 else self._replay.sample(T * N)["amp_obs"]     # :196-200
        # prepare_dataset (common_agent.py:426-456): advantages from the un-normalised returns / values, then both normalised
        returns, values = fl(b["returns"]), fl(b["values"])
        adv = (returns - values).sum(dim=1)
        if self.normalize_advantage:
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)                               # _calc_advs (:685-696)
        old_values = self._value_norm(values.contiguous())
        ret_n = self._value_norm(returns.contiguous())
        data = dict(obs=fl(b["obses"]), actions=fl(b["actions"]), old_logp_actions=fl(b["neglogpacs"].unsqueeze(-1))[:, 0], advantages=adv,
                    returns=ret_n, old_values=old_values, mu=fl(b["mus"]), sigma=torch.exp(R.net.sigma.detach()).expand(T * N, -1),
                    amp_obs=amp_obs, amp_obs_replay=replay, amp_obs_demo=demo)
        mb, amb = self.minibatch_size, self.amp_minibatch_size
        infos = []
        idx = torch.randperm(self.batch_size, device=adv.device, generator=self.gen)    # AMPDataset._idx_buf (amp_datasets.py:8)
        for _ in range(self.mini_epochs):                                               # :214-216
            for i in range(self.batch_size // mb):
                sel = idx[i * mb:(i + 1) * mb]
                batch = {k: v[sel].contiguous() for k, v in data.items() if not k.startswith("amp_obs")}
                for k in ("amp_obs", "amp_obs_replay", "amp_obs_demo"):                 # input_dict[k][0:amp_minibatch_size] (:289-295)
                    batch[k] = data[k][sel[:amb]].contiguous()
                up.step(batch)
            idx = torch.randperm(self.batch_size, device=adv.device, generator=self.gen)         # reshuffled when exhausted (:24-26)
            infos.append(up.info())
        self._store_replay(amp_obs)                                                     # :264
        self.epoch += 1
        return dict(epoch=self.epoch, frames=self.batch_size, **{k: sum(i[k] for i in infos) / len(infos) for k in infos[0]})

    def _store_replay(self, amp_obs):
        """_store_replay_amp_obs (amp_continuous.py:694-710)."""
        size = self._replay.get_buffer_size()
        if self._replay.get_total_count() > size:
            keep = torch.rand(amp_obs.shape[0], device=amp_obs.device, generator=self.gen) < self._keep_prob
            amp_obs = amp_obs[keep]
        if amp_obs.shape[0] > size:
            amp_obs = amp_obs[torch.randperm(amp_obs.shape[0], device=amp_obs.device, generator=self.gen)[:size]]
        self._replay.store({"amp_obs": amp_obs})

    def close(self):
        self.R.close()
