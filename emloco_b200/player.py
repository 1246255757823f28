"""Inference / LocoVal-correlation loop: `AMPPlayerContinuousValue.run` of the reference
(pacer/pacer/learning/amp_value_players.py:37-275, launched by `run.py --test`, eval_policy.sh:12-21).

The reference plays ONE env on the host: per episode it scores the trajectory with LocoVal at the first step (:128-137),
accumulates the discounted reward ((r_loc + r_pow) * 0.5 + r_disc * 0.25) * gamma^(n+1) (:144-160), snapshots it at
`step_to_pred` or at an earlier end (:177-193), compares prediction and normalised return by MSE (:195-198) and finally
reports their correlation (:263-279).  Here all N envs of a `Rollout` play at once with deterministic actions (the mean of
the policy, rl_games' `is_determenistic`), the bookkeeping runs in one kernel per step (`emloco_player_record`) and finished
episodes land in a device-side result list; plotting / video / the live plotter are out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .rollout import Rollout
from .sim import _ptr, _stream


class AMPPlayerContinuousValue:
    def __init__(self, rollout: Rollout, plot_val_reward=True, inversion_penalty_scale=0.3, min_reward=-10.0, max_reward=100.0,
                 capacity=1 << 20):
        self.R = rollout
        self.plot_val_reward = bool(plot_val_reward)
        self.inversion_penalty_scale = float(inversion_penalty_scale)
        self.min_reward, self.max_reward = float(min_reward), float(max_reward)       # :56-57
        self.gamma, self.step_to_pred = float(rollout.gamma), int(rollout.rcfg.step_to_pred)
        N, dev = rollout.N, rollout.state.device
        self.state = torch.zeros(11, N, device=dev)
        self.state[2].fill_(1.0)
        self.capacity = int(capacity)
        self.results = torch.zeros(self.capacity, 8, device=dev)
        self.count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.zero_noise = torch.zeros(N, 69, device=dev)
        self._k = 0

    def step(self):
        """One control step of every env: env_reset(done) -> get_action (mu) -> env_step -> bookkeeping."""
        R = self.R
        R.step(self._k % R.T, noise=self.zero_noise)
        self._k += 1
        _lib.check(_lib.load().emloco_player_record(
            _ptr(R.sim.rew), _ptr(R.sim.rew_raw), _ptr(R.sim.reset), _ptr(R._cur["logit"]), _ptr(R.locoval_scores), _ptr(R.inverted),
            _ptr(self.state), R.N, _ptr(self.results), _ptr(self.count), self.capacity, int(self.plot_val_reward),
            self.inversion_penalty_scale, R.disc_reward_scale, self.gamma, self.step_to_pred, self.min_reward, self.max_reward, _stream()),
            "emloco_player_record")

    def games_played(self):
        return int(self.count.item())

    def run(self, n_games, max_steps=None):
        """Plays until `n_games` episodes have finished.  -> dict(games, value_loss (mean squared error of LocoVal against the
        normalised return, the reference's `av value loss`), corr_total / corr_loc / corr_pow / corr_disc (np.corrcoef of the
        predictions with the returns, :271-279), vals, rewards, steps)."""
        max_steps = max_steps or (int(np.ceil(n_games / self.R.N)) + 2) * 400
        for i in range(max_steps):
            self.step()
            if (i & 15) == 15 and self.games_played() >= n_games:
                break
        n = min(self.games_played(), self.capacity)
        res = self.results[:n].cpu().numpy()
        res = res[np.argsort(res[:, 0], kind="stable")]
        vals, rewards, norm = res[:, 1], res[:, 2], res[:, 3]
        corr = lambda a, b: float(np.corrcoef(a, b)[0, 1]) if n > 1 and a.std() > 0 and b.std() > 0 else float("nan")
        return dict(games=n, value_loss=float(((vals - norm) ** 2).mean()) if n else float("nan"), corr_total=corr(vals, rewards),
                    corr_loc=corr(vals, res[:, 4]), corr_pow=corr(vals, res[:, 5]), corr_disc=corr(vals, res[:, 6]), vals=vals, rewards=rewards,
                    steps=res[:, 7], env=res[:, 0].astype(np.int64))
