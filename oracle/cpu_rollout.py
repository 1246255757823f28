"""TEST INFRASTRUCTURE ONLY - the whole rollout step on the CPU, assembled from the oracle pieces
(oracle_np.py restatements of the reference's torch code + physics_oracle.c).  Used by
  * tests/ as the end-to-end checker of `emloco_b200.rollout.Rollout`,
  * bench.py's `cpu_baseline` leg and `--impl reference` arm (the reference's Isaac Gym CPU pipeline cannot run: its
    binaries are absent, SURVEY 8c - so this "port" is the CPU number that can be had).
The product never imports this module.

One step = the body of the `for n in range(horizon_length)` loop of AMPValueAgent.play_steps
(pacer/pacer/learning/amp_continuous_value.py:44-118) for n envs.
"""
from __future__ import annotations

import numpy as np

from . import oracle_np as O
from . import physics_oracle as PO

F = np.float32


def weights_from_state_dict(sd):
    """`a2c_network.*`-named tensors (torch or numpy) -> the P (policy) and D (discriminator) dicts of oracle_np."""
    g = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k], F)
    lay = lambda p, i: (g(f"{p}.{i}.weight"), g(f"{p}.{i}.bias"))
    lin = lambda p: (g(f"{p}.weight"), g(f"{p}.bias"))
    P = dict(task=[lay("_task_mlp", 0), lay("_task_mlp", 2)], actor=[lay("actor_mlp", 0), lay("actor_mlp", 2)],
             mu=lin("mu"), sigma=g("sigma"), critic=[lay("critic_mlp", 0), lay("critic_mlp", 2)], value=lin("value"),
             tv=[lay("_task_value_mlp", 0), lay("_task_value_mlp", 2)], tv_out=lin("_value_logits"))
    D = dict(mlp=[lay("_disc_mlp", 0), lay("_disc_mlp", 2)], logit=lin("_disc_logits"))
    return P, D


class CpuRollout:
    def __init__(self, model_arrays, state, P, D, obs_stats=None, amp_stats=None, value_stats=(0.0, 1.0), gamma=0.99,
                 disc_scale=2.0, step_to_pred=144, inv_penalty=0.3, traj_flags=None, traj_pool=None,
                 traj_seed=0):
        A = model_arrays
        self.A = A
        self.M = PO.make_model(A["parent"], A["offset"], A["mass"], A["com"], A["inertia6"], A["kp_joint"], A["kd_joint"],
                               A["arm_joint"], A["geom_type"], A["geom_a"], A["geom_b"], A["geom_r"])
        self.cfg = PO.make_cfg(1.0 / 120.0)
        n = state["root"].shape[0]
        self.n = n
        self.init_root = state["root"].astype(np.float64)
        self.init_dof = state["dof"].reshape(n, 69, 2).astype(np.float64)
        self.verts = state["verts"].astype(F)
        self.P, self.D = dict(P), dict(D)
        self.P["mean"], self.P["var"] = obs_stats if obs_stats is not None else (np.zeros(1422), np.ones(1422))
        self.D["mean"], self.D["var"] = amp_stats if amp_stats is not None else (np.zeros(3090), np.ones(3090))
        self.v_mean, self.v_var = F(value_stats[0]), F(value_stats[1])
        self.inverted = None      # task.inverted (bool [n]); set by the caller when trajectories can be inverted
        self.gamma, self.disc_scale, self.step_to_pred, self.inv_penalty = F(gamma), disc_scale, step_to_pred, inv_penalty
        self.height = np.zeros((1080, 1080), np.int16)
        self.betas = np.zeros((n, 17), F)
        # sim state
        self.root = np.zeros((n, 13)); self.jq = np.zeros((n, 23, 4)); self.jw = np.zeros((n, 69))
        self.progress = np.zeros(n, np.int64)
        self.reset = np.ones(n, np.int64); self.terminate = np.zeros(n, np.int64)
        self.amp_buf = np.zeros((n, 15, 206), F)
        self.contact = np.zeros((n, 24, 3)); self.dof_force = np.zeros((n, 69))
        self.state = np.zeros((6, n), F); self.state[3] = 1
        self.obs = np.zeros((n, 1422), F)
        self.traj_flags = None
        self.reset_done()
        # TrajGenerator.reset for the envs that reset from now on (numpy draws; the device path uses Philox - same distribution)
        self.traj_flags, self.traj_pool, self.traj_rng = traj_flags, traj_pool, np.random.default_rng(traj_seed)

    # env_reset(done_indices): humanoid.py:455-481 + humanoid_amp.py:284-293,499-502 with a fixed initial state
    def reset_done(self):
        ids = np.nonzero(self.reset)[0]
        if len(ids) == 0:
            return
        self.root[ids] = self.init_root[ids]
        self.jq[ids] = PO.expmap_to_quat(self.init_dof[ids, :, 0]).reshape(len(ids), 23, 4)
        self.jw[ids] = self.init_dof[ids, :, 1]
        self.progress[ids] = 0; self.reset[ids] = 0; self.terminate[ids] = 0
        self.contact[ids] = 0; self.dof_force[ids] = 0
        rb, dp = PO.refresh(self.M, self.root[ids], self.jq[ids], self.jw[ids])
        out = self._post(ids, rb, dp, advance=False)
        self.obs[ids] = out["obs"]
        self.amp_buf[ids] = out["amp_obs"].reshape(len(ids), 15, 206)[:, :1]      # history := current step
        if self.traj_flags is not None:                                           # _reset_task AFTER the observations (humanoid_amp_task.py:54-57)
            U = self.traj_rng.random((len(ids), O.TRAJ_RAND_COLS)).astype(F)
            inv = O.traj_reset(self.verts, ids, self.root[ids, 0:3].astype(F), self.root[ids, 7:10].astype(F), U, self.traj_flags, self.traj_pool)
            if self.inverted is None:
                self.inverted = np.zeros(self.n, bool)
            self.inverted[ids] = inv                                                  # task.inverted -> inversion penalty (:62-64)

    def _post(self, ids, rb, dof_pos, advance):
        ds = np.stack([dof_pos, self.jw[ids]], -1).astype(F)
        prog = self.progress[ids] + (1 if advance else 0)
        return O.post_physics_step(rb.astype(F), ds, self.contact[ids].astype(F), self.dof_force[ids].astype(F), prog,
                                   self.verts[ids], self.betas[ids], self.height, self.amp_buf[ids])

    def step(self, noise):
        n = self.n
        self.reset_done()
        obs = self.obs.copy()
        pol = O.policy_forward(obs, self.P, noise=noise)
        tgt = O.action_to_pd_targets(pol["actions"], self.A["pd_offset"], self.A["pd_scale"]).astype(np.float64)
        rb, dp, ct, df = PO.step(self.M, self.cfg, 4, self.root, self.jq, self.jw, tgt, height=self.height)
        self.contact, self.dof_force = ct, df
        all_ids = np.arange(n)
        out = self._post(all_ids, rb, dp, advance=True)
        self.progress += 1
        self.obs = out["obs"]; self.amp_buf = out["amp_obs"].reshape(n, 15, 206)
        self.reset, self.terminate = out["reset"], out["terminate"]
        nv_raw = O.critic_forward(self.obs, self.P)[:, 0]
        amp_r, logit = O.disc_reward(out["amp_obs"], self.D, self.disc_scale)
        # :61-118 (oracle_np.rollout_record; self.state rows: 0 current_rewards, 1 current_lengths, 2 current_combined,
        # 3 discount, 4 game_combined, 5 terminated_flags)
        names = ("current_rewards", "current_lengths", "current_combined", "discount", "game_combined", "terminated_flags")
        d = {k: self.state[i] for i, k in enumerate(names)}
        rows = O.rollout_record(d, out["rew"], self.reset, self.terminate, nv_raw, logit[:, 0], self.inverted, self.v_mean, self.v_var,
                                self.inv_penalty, self.disc_scale, self.gamma, self.step_to_pred)
        for i, k in enumerate(names):
            self.state[i] = d[k]
        r, done, next_values = rows["rewards"], rows["dones"], rows["next_values"]
        values = O.value_unnormalize(pol["value"][:, 0], self.v_mean, self.v_var)
        return dict(obs=obs, actions=pol["actions"], neglogp=pol["neglogp"], mu=pol["mu"], values=values,
                    task_values=pol["task_value"], rewards=r, dones=done, next_values=next_values, amp_rewards=rows["amp_rewards"],
                    next_obs=self.obs, amp_obs=out["amp_obs"], rb=rb, disc_logit=logit)
