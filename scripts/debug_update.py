"""One-off: compares every intermediate of PPOUpdate with torch autograd (retain_grad) to find the first divergence."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_gpu_update import _setup, _dev
from emloco_b200.policy import AMPSeptValueNetwork
from oracle import update_oracle
from oracle.make_golden import UPDATE_CFG as U

B, Ba = int(sys.argv[1]) if len(sys.argv) > 1 else 1000, int(sys.argv[2]) if len(sys.argv) > 2 else 600
LOGSTD = -1.0
up, net, sd, batch, stats = _setup(B, Ba, 5, 7, sigma=LOGSTD)
ref = AMPSeptValueNetwork(); ref.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}); ref = ref.cuda()
rng = np.random.default_rng(1)
x = torch.from_numpy(update_oracle.rms_normalize(batch["obs"], stats["obs_mean"], stats["obs_var"]).astype(np.float32)).cuda()
with torch.no_grad():
    mu0 = ref.mu(ref.actor_mlp(torch.cat([x[:, :368], ref._task_mlp(x[:, 368:])], -1))).cpu().numpy()
sg = np.exp(LOGSTD)
batch["actions"] = (mu0 + sg * rng.normal(0, 1, mu0.shape)).astype(np.float32)
nl0 = 0.5 * (((batch["actions"] - mu0) / sg) ** 2).sum(-1) + 0.5 * np.log(2 * np.pi) * 69 + 69 * LOGSTD
batch["old_logp_actions"] = (nl0 + rng.normal(0, 0.15, B)).astype(np.float32)
d = _dev(batch)
up.forward_backward(d, dropout_u=d["dropout_u"])
torch.cuda.synchronize()
T = {}
def keep(name, t):
    t.retain_grad(); T[name] = t; return t
t1 = keep("t1", torch.relu(ref._task_mlp[0](x[:, 368:]))); t2 = keep("t2", torch.relu(ref._task_mlp[2](t1)))
ain = keep("ain", torch.cat([x[:, :368], t2], -1))
a1 = keep("a1", torch.relu(ref.actor_mlp[0](ain))); a2 = keep("a2", torch.relu(ref.actor_mlp[2](a1))); mu = keep("mu", ref.mu(a2))
c1 = keep("c1", torch.relu(ref.critic_mlp[0](ain))); c2 = keep("c2", torch.relu(ref.critic_mlp[2](c1))); value = keep("value", ref.value(c2))
v1 = keep("v1", torch.relu(ref._task_value_mlp[0](x[:, 368:398]))); v2 = keep("v2", torch.relu(ref._task_value_mlp[2](v1))); tv = keep("tv", ref._value_logits(v2))
sigma = torch.exp(ref.sigma)
neglogp = 0.5 * (((d["actions"] - mu) / sigma) ** 2).sum(-1) + 0.5 * np.log(2 * np.pi) * 69 + ref.sigma.sum()
ratio = torch.exp(d["old_logp_actions"] - neglogp)
a_loss = torch.max(-d["advantages"] * ratio, -d["advantages"] * torch.clamp(ratio, 0.8, 1.2)).mean()
c_loss = ((d["returns"] - value) ** 2).mean(); tv_loss = ((d["returns"] - tv) ** 2).mean()
b_loss = (torch.clamp_min(mu - 1, 0) ** 2 + torch.clamp_max(mu + 1, 0) ** 2).sum(-1).mean()
loss = a_loss + 5 * c_loss + 10 * b_loss + 5 * tv_loss
loss.backward()
torch.cuda.synchronize()
h = up.h
ours_f = dict(t1=up.t1_32, t2=up.t2_32, a1=up.ac1_32[:, :h], c1=up.ac1_32[:, h:], a2=up.a2_32, c2=up.c2_32, mu=up.mu32, value=up.value, tv=up.tv, v1=up.v1_32, v2=up.v2_32)
def rel(a, b):
    a, b = a.detach().float().cpu().numpy().astype(np.float64), b.detach().float().cpu().numpy().astype(np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)), float(np.abs(b).max())
for k, v in ours_f.items():
    print("fwd", k, rel(v, T[k]))
m = lambda g, a: g * (a > 0)
ours_b = dict(mu=up.dmu32, value=up.dvalue.view(B, 1), tv=up.dtv.view(B, 1), a2=m(up.da2_32, up.a2_32), a1=m(up.dac1_32[:, :h], up.ac1_32[:, :h]),
              c1=m(up.dac1_32[:, h:], up.ac1_32[:, h:]), ain=up.dain_32, t1=m(up.dt1_32, up.t1_32), v1=m(up.dv1_32, up.v1_32))
for k, v in ours_b.items():
    gr = T[k].grad * (T[k] > 0) if k in ("a2", "a1", "c1", "t1", "v1") else T[k].grad
    print("bwd", k, rel(v, gr))
# raw (ungated) checks of single products
print("raw da2 vs dmu@Wmu", rel(up.da2_32, up.dmu32 @ ref.mu.weight))
dc2 = (up.dvalue.view(B, 1) * ref.value.weight) * (up.c2_32 > 0)
print("raw dac1[:, h:] vs dc2@Wc2", rel(up.dac1_32[:, h:], dc2 @ ref.critic_mlp[2].weight))
print("critic_mlp.2.bias", rel(up.flat.grad("critic_mlp.2.bias"), dc2.sum(0)), rel(up.flat.grad("critic_mlp.2.bias"), ref.critic_mlp[2].bias.grad))
print("critic_mlp.2.weight vs dc2^T c1", rel(up.flat.grad("critic_mlp.2.weight"), dc2.t() @ up.ac1_32[:, h:]), rel(up.flat.grad("critic_mlp.2.weight"), ref.critic_mlp[2].weight.grad))
print("mu.weight", rel(up.flat.grad("mu.weight"), ref.mu.weight.grad))
