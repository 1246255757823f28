"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy, float64, every derivative written out by hand) of one PPO / AMP update:
`AMPValueAgent.calc_gradients` (pacer/pacer/learning/amp_continuous_value.py:276-428) with the losses of
learning/common_agent.py:594-602,657-683, amp_continuous_value.py:430-444 and amp_continuous.py:536-616, the
nn.utils.clip_grad_norm_ / torch.optim.Adam step of common_agent.py:84-87 and the training-mode RunningMeanStd updates
(utils/running_mean_std.py:33-43,86-96).  rl_games 1.1.4 (not in the tree) supplies neglogp / entropy / policy_kl - restated.

Pinned by tests/golden/update_step.npz (the reference's own code under torch autograd, oracle/make_golden.reference_update_step).
The CUDA path (emloco_b200/update.py) follows the same decomposition: every `@` below is one tcgen05 GEMM there.
"""
from __future__ import annotations

import numpy as np

D = np.float64
LOG2PI_HALF = 0.5 * np.log(2.0 * np.pi)


def rms_normalize(x, mean, var, eps=1e-5):
    y = (x.astype(np.float32) - mean.astype(np.float32)) / np.sqrt(var.astype(np.float32) + np.float32(eps))
    return np.clip(y, -5.0, 5.0).astype(D)


def rms_update(mean, var, count, x):
    """running_mean_std.py:33-43 with the reference's type promotion (batch moments in fp32)."""
    x = x.astype(np.float32)
    b = x.shape[0]
    bm, bv = x.mean(0, dtype=np.float32), x.var(0, ddof=1, dtype=np.float32)
    delta = bm.astype(D) - mean
    tot = count + b
    new_mean = mean + delta * b / tot
    m2 = var * count + (bv * np.float32(b)).astype(D) + delta ** 2 * count * b / tot
    return new_mean, m2 / tot, tot


def dropout_mask(u, rate=0.3):
    """amp_models.py:49-90: u [19, Ba, 3] uniform draws -> mask [3, Ba, 3090] (agent, replay, demo)."""
    Ba = u.shape[1]
    m = np.ones((3, Ba, 206), D)
    for j in range(19):
        keep = (u[j] > np.float32(rate)).T.astype(D)                  # [3, Ba]
        m[:, :, 12 + 6 * j:18 + 6 * j] = keep[:, :, None]
        m[:, :, 126 + 3 * j:129 + 3 * j] = keep[:, :, None]
    return np.tile(m, (1, 1, 15))


def _lin(x, W, b):
    return x @ W.T + b


def update_step(sd, batch, stats, cfg, world=1):
    """sd: `a2c_network.*` parameters (numpy); batch / stats: oracle.make_golden.synth_update_batch.  -> dict(losses..., grads,
    new_params, stats_after).  Gradients are those of the total loss; `new_params` after clip-norm + the first Adam step."""
    P = {k: v.astype(D) for k, v in sd.items()}
    B, Ba = batch["obs"].shape[0], batch["amp_obs"].shape[0]
    out = {}
    # ---- input normalisation (training mode: normalise with the current statistics, THEN absorb the batch) ----
    x = rms_normalize(batch["obs"], stats["obs_mean"], stats["obs_var"])
    om, ov, oc = rms_update(stats["obs_mean"], stats["obs_var"], stats["obs_count"], batch["obs"])
    am, av, ac = stats["amp_mean"], stats["amp_var"], stats["amp_count"]
    amps = []
    for k in ("amp_obs", "amp_obs_replay", "amp_obs_demo"):          # three successive calls: each sees the previous update
        amps.append(rms_normalize(batch[k], am, av))
        am, av, ac = rms_update(am, av, ac, batch[k])
    out["stats_after"] = dict(obs_mean=om, obs_var=ov, obs_count=oc, amp_mean=am, amp_var=av, amp_count=ac)
    mask = dropout_mask(batch["dropout_u"], cfg["dropout_rate"])
    xa = np.concatenate([amps[i] * mask[i] for i in range(3)], 0)    # [3 Ba, 3090]: agent, replay, demo

    # ---- forward ----
    xs, xt = x[:, :368], x[:, 368:]
    t1 = np.maximum(_lin(xt, P["_task_mlp.0.weight"], P["_task_mlp.0.bias"]), 0)
    t2 = np.maximum(_lin(t1, P["_task_mlp.2.weight"], P["_task_mlp.2.bias"]), 0)
    ain = np.concatenate([xs, t2], 1)
    a1 = np.maximum(_lin(ain, P["actor_mlp.0.weight"], P["actor_mlp.0.bias"]), 0)
    a2 = np.maximum(_lin(a1, P["actor_mlp.2.weight"], P["actor_mlp.2.bias"]), 0)
    mu = _lin(a2, P["mu.weight"], P["mu.bias"])
    c1 = np.maximum(_lin(ain, P["critic_mlp.0.weight"], P["critic_mlp.0.bias"]), 0)
    c2 = np.maximum(_lin(c1, P["critic_mlp.2.weight"], P["critic_mlp.2.bias"]), 0)
    value = _lin(c2, P["value.weight"], P["value.bias"])[:, 0]
    xv = x[:, 368:398]
    v1 = np.maximum(_lin(xv, P["_task_value_mlp.0.weight"], P["_task_value_mlp.0.bias"]), 0)
    v2 = np.maximum(_lin(v1, P["_task_value_mlp.2.weight"], P["_task_value_mlp.2.bias"]), 0)
    tv = _lin(v2, P["_value_logits.weight"], P["_value_logits.bias"])[:, 0]
    h1 = np.maximum(_lin(xa, P["_disc_mlp.0.weight"], P["_disc_mlp.0.bias"]), 0)
    h2 = np.maximum(_lin(h1, P["_disc_mlp.2.weight"], P["_disc_mlp.2.bias"]), 0)
    logit = _lin(h2, P["_disc_logits.weight"], P["_disc_logits.bias"])[:, 0]
    out.update(mus=mu, values=value, task_values=tv, disc_logit=logit)

    # ---- PPO heads ----
    logstd = P["sigma"]; sigma = np.exp(logstd)
    act = batch["actions"].astype(D)
    neglogp = 0.5 * (((act - mu) / sigma) ** 2).sum(-1) + LOG2PI_HALF * 69 + logstd.sum()
    ratio = np.exp(batch["old_logp_actions"].astype(D) - neglogp)
    A = batch["advantages"].astype(D)
    e = cfg["e_clip"]
    rc = np.clip(ratio, 1 - e, 1 + e)
    s1, s2 = -A * ratio, -A * rc
    a_loss = np.maximum(s1, s2)
    inside = (ratio >= 1 - e) & (ratio <= 1 + e)
    dl_dratio = np.where(inside | (s1 > s2), -A, 0.0)
    ret = batch["returns"].astype(D)[:, 0]
    hi, lo = np.maximum(mu - 1, 0), np.minimum(mu + 1, 0)
    b_loss = (hi ** 2 + lo ** 2).sum(-1)
    out.update(neglogp=neglogp, a_loss=a_loss.mean(), c_loss=((ret - value) ** 2).mean(), tv_loss=((ret - tv) ** 2).mean(),
               b_loss=b_loss.mean(), a_clip_frac=(np.abs(ratio - 1) > e).mean(), entropy=(logstd + 0.5 + LOG2PI_HALF).sum())
    p1m, p1s = batch["mu"].astype(D), batch["sigma"].astype(D)
    out["kl"] = (np.log(p1s / sigma + 1e-5) + (sigma ** 2 + (p1m - mu) ** 2) / (2 * (p1s ** 2 + 1e-5)) - 0.5).sum(-1).mean()
    dneglogp = -ratio * dl_dratio * cfg["actor_coef"] / B
    dmu = dneglogp[:, None] * (-(act - mu) / sigma ** 2) + cfg["bounds_loss_coef"] / B * 2 * (hi + lo)
    dvalue = -2 * (ret - value) * cfg["critic_coef"] / B
    dtv = -2 * (ret - tv) * cfg["tv_coef"] / B

    G = {}

    def back(name, dy, xin, need_dx=True):
        """y = xin W^T + b: accumulates dW, db; returns dx."""
        G[name + ".weight"] = G.get(name + ".weight", 0) + dy.T @ xin
        G[name + ".bias"] = G.get(name + ".bias", 0) + dy.sum(0)
        return dy @ P[name + ".weight"] if need_dx else None

    # actor / critic / task trunk
    da2 = back("mu", dmu, a2) * (a2 > 0)
    da1 = back("actor_mlp.2", da2, a1) * (a1 > 0)
    dain = back("actor_mlp.0", da1, ain)
    dc2 = back("value", dvalue[:, None], c2) * (c2 > 0)
    dc1 = back("critic_mlp.2", dc2, c1) * (c1 > 0)
    dain = dain + back("critic_mlp.0", dc1, ain)
    dt2 = dain[:, 368:] * (t2 > 0)
    dt1 = back("_task_mlp.2", dt2, t1) * (t1 > 0)
    back("_task_mlp.0", dt1, xt, need_dx=False)
    dv2 = back("_value_logits", dtv[:, None], v2) * (v2 > 0)
    dv1 = back("_task_value_mlp.2", dv2, v1) * (v1 > 0)
    back("_task_value_mlp.0", dv1, xv, need_dx=False)

    # ---- discriminator (amp_continuous.py:536-598), everything scaled by disc_coef ----
    dc = cfg["disc_coef"]
    na = 2 * Ba
    la, ld = logit[:na], logit[na:]
    sp = lambda z: np.maximum(z, 0) + np.log1p(np.exp(-np.abs(z)))
    pred = 0.5 * (sp(la).mean() + sp(-ld).mean())
    sig = 1 / (1 + np.exp(-logit))
    dlogit = np.concatenate([0.5 * sig[:na] / na, -0.5 * (1 - sig[na:]) / Ba]) * dc
    w3 = P["_disc_logits.weight"]
    logit_reg = (w3 ** 2).sum()
    wd = (P["_disc_mlp.0.weight"] ** 2).sum() + (P["_disc_mlp.2.weight"] ** 2).sum() + (w3 ** 2).sum()
    # gradient penalty on the demo rows: g = d logit / d (normalised demo obs) = mask * (((w3 * m2) W2 * m1) W1)
    W1, W2 = P["_disc_mlp.0.weight"], P["_disc_mlp.2.weight"]
    m1, m2 = (h1[na:] > 0).astype(D), (h2[na:] > 0).astype(D)
    u2 = m2 * w3                                   # [Ba, 512]
    u1 = (u2 @ W2) * m1                            # [Ba, 1024]
    gx = u1 @ W1                                   # [Ba, 3090]
    g = gx * mask[2]
    gp = (g ** 2).sum(-1).mean()
    disc_loss = dc * (pred + cfg["disc_logit_reg"] * logit_reg + cfg["disc_grad_penalty"] * gp + cfg["disc_weight_decay"] * wd)
    out.update(disc_loss=disc_loss, disc_grad_penalty=gp, disc_logit_loss=logit_reg, disc_agent_acc=(la < 0).mean(), disc_demo_acc=(ld > 0).mean())
    # prediction-loss backward
    dh2 = back("_disc_logits", dlogit[:, None], h2) * (h2 > 0)
    dh1 = back("_disc_mlp.2", dh2, h1) * (h1 > 0)
    back("_disc_mlp.0", dh1, xa, need_dx=False)
    # gradient-penalty backward (second order; ReLU masks are constants)
    cgp = dc * cfg["disc_grad_penalty"]
    e0 = cgp * 2 * g * mask[2] / Ba                # d / d gx
    G["_disc_mlp.0.weight"] += u1.T @ e0
    dv1g = (e0 @ W1.T) * m1                        # d / d (u2 W2)
    G["_disc_mlp.2.weight"] += u2.T @ dv1g
    du2 = dv1g @ W2.T
    G["_disc_logits.weight"] += (du2 * m2).sum(0)[None]
    # regularisers
    G["_disc_logits.weight"] += dc * (cfg["disc_logit_reg"] + cfg["disc_weight_decay"]) * 2 * w3
    G["_disc_mlp.0.weight"] += dc * cfg["disc_weight_decay"] * 2 * W1
    G["_disc_mlp.2.weight"] += dc * cfg["disc_weight_decay"] * 2 * W2
    out["loss"] = (cfg["actor_coef"] * out["a_loss"] + cfg["critic_coef"] * out["c_loss"] - cfg["entropy_coef"] * out["entropy"]
                   + cfg["bounds_loss_coef"] * out["b_loss"] + disc_loss + cfg["tv_coef"] * out["tv_loss"])
    out["grads"] = G

    # ---- clip_grad_norm_ + first Adam step (bias corrections at t = 1) ----
    total = np.sqrt(sum((g_ ** 2).sum() for g_ in G.values()))
    out["total_norm"] = total
    coef = min(cfg["grad_norm"] / (total + 1e-6), 1.0) / world
    new = {}
    for k, g_ in G.items():
        gc = g_ * coef
        m, v = 0.1 * gc, 0.001 * gc ** 2
        new[k] = P[k].reshape(g_.shape) - (cfg["lr"] / 0.1) * m / (np.sqrt(v) / np.sqrt(0.001) + 1e-8)
    out["new_params"] = new
    return out
