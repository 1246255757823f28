"""GPU: the rollout step (nets + env step + bookkeeping) through the C ABI against the CPU oracle assembly
(oracle/cpu_rollout.py = oracle_np restatements pinned to the reference goldens + the fp64 physics restatement).
Tolerance: north_star's 1e-3 relative for floats (with absolute floors scaled to each quantity), bit-exact masks."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-3


def _pair(n, seed=0, **kw):
    from emloco_b200.model import build_model_arrays, rest_root_height
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    from emloco_b200.synthetic import synthetic_env_state
    from oracle.cpu_rollout import CpuRollout, weights_from_state_dict
    torch.manual_seed(seed)
    net = AMPSeptValueNetwork()
    P, D = weights_from_state_dict(net.state_dict())
    A = build_model_arrays()
    st = synthetic_env_state(n, seed, rest_root_height(A))
    cpu = CpuRollout(A, st, P, D)
    gpu = Rollout(n, seed=seed, net=net, **kw)
    return gpu, cpu


@pytest.mark.parametrize("tc", [False, True])
def test_first_steps_match_cpu_oracle(tc):
    n = 64
    gpu, cpu = _pair(n, seed=3, tensor_cores=tc)
    np.testing.assert_allclose(gpu.sim.obs.cpu().numpy()[:, :398], cpu.obs[:, :398], rtol=RTOL, atol=2e-5)
    rng = np.random.default_rng(0)
    for k in range(2):
        noise = rng.standard_normal((n, 69)).astype(np.float32)
        o = cpu.step(noise)
        gpu.step(k, noise=torch.from_numpy(noise).cuda())
        torch.cuda.synchronize()
        mb = {key: v[k].cpu().numpy() for key, v in gpu.mb.items() if v is not None}
        loose = 1.0 if k == 0 else 20.0       # step 2 starts from states that already differ by fp32-vs-fp64 round-off
        np.testing.assert_allclose(mb["mus"], o["mu"], rtol=RTOL, atol=2e-4 * loose)
        np.testing.assert_allclose(mb["actions"], o["actions"], rtol=RTOL, atol=2e-4 * loose)
        np.testing.assert_allclose(mb["neglogpacs"], o["neglogp"], rtol=RTOL, atol=1e-3 * loose)
        np.testing.assert_allclose(mb["values"][:, 0], o["values"], rtol=RTOL, atol=2e-4 * loose)
        np.testing.assert_allclose(mb["task_values"], o["task_values"], rtol=RTOL, atol=2e-4 * loose)
        np.testing.assert_allclose(gpu.sim.rb_state.view(n, 24, 13).cpu().numpy(), o["rb"], rtol=RTOL, atol=5e-3 * loose)
        np.testing.assert_allclose(mb["rewards"][:, 0], o["rewards"], rtol=5e-3, atol=2e-2 * loose)
        np.testing.assert_array_equal(mb["dones"], o["dones"])
        np.testing.assert_allclose(mb["next_values"][:, 0], o["next_values"], rtol=RTOL, atol=5e-4 * loose)
        np.testing.assert_allclose(mb["amp_rewards"][:, 0], o["amp_rewards"], rtol=RTOL, atol=2e-3 * loose)
        np.testing.assert_allclose(gpu.state.cpu().numpy(), cpu.state, rtol=5e-3, atol=3e-2 * loose)
    gpu.close()


def test_nets_match_oracle_on_identical_obs():
    """The networks alone on random (not simulated) observations: isolates the GEMM path from physics round-off."""
    from emloco_b200.policy import AMPSeptValueNetwork, RolloutNets, RunningMeanStd
    from oracle import oracle_np as O
    from oracle.cpu_rollout import weights_from_state_dict
    torch.manual_seed(1)
    net = AMPSeptValueNetwork().cuda()
    P, D = weights_from_state_dict(net.state_dict())
    rng = np.random.default_rng(2)
    M = 200
    on, an = RunningMeanStd(1422).cuda(), RunningMeanStd(3090).cuda()
    on.running_mean.copy_(torch.from_numpy(rng.normal(0, 0.5, 1422))); on.running_var.copy_(torch.from_numpy(rng.uniform(0.2, 3, 1422)))
    an.running_mean.copy_(torch.from_numpy(rng.normal(0, 0.5, 3090))); an.running_var.copy_(torch.from_numpy(rng.uniform(0.2, 3, 3090)))
    P["mean"], P["var"] = on.running_mean.cpu().numpy(), on.running_var.cpu().numpy()
    D["mean"], D["var"] = an.running_mean.cpu().numpy(), an.running_var.cpu().numpy()
    obs = rng.normal(0, 2, (M, 1422)).astype(np.float32)
    amp = rng.normal(0, 2, (M, 3090)).astype(np.float32)
    noise = rng.standard_normal((M, 69)).astype(np.float32)
    ref = O.policy_forward(obs, P, noise=noise)
    ref_c = O.critic_forward(obs, P)
    ref_r, ref_l = O.disc_reward(amp, D)
    for tc in (False, True):
        nets = RolloutNets(net, on, an, M, tensor_cores=tc)
        res = nets.action_values(torch.from_numpy(obs).cuda(), torch.from_numpy(noise).cuda())
        tol = dict(rtol=RTOL, atol=2e-4)
        np.testing.assert_allclose(res["mus"].cpu().numpy(), ref["mu"], **tol)
        np.testing.assert_allclose(res["values"].cpu().numpy(), ref["value"], **tol)
        np.testing.assert_allclose(res["task_values"].cpu().numpy(), ref["task_value"], **tol)
        np.testing.assert_allclose(res["actions"].cpu().numpy(), ref["actions"], **tol)
        np.testing.assert_allclose(res["neglogpacs"].cpu().numpy(), ref["neglogp"], rtol=RTOL, atol=1e-3)
        np.testing.assert_allclose(nets.critic(torch.from_numpy(obs).cuda()).cpu().numpy(), ref_c, **tol)
        lg = nets.disc_logits(torch.from_numpy(amp).cuda())
        np.testing.assert_allclose(lg.cpu().numpy(), ref_l, rtol=RTOL, atol=1e-3)
        from emloco_b200.policy import disc_reward
        r, comb = disc_reward(lg, torch.ones_like(lg))
        np.testing.assert_allclose(r.cpu().numpy(), ref_r, rtol=RTOL, atol=2e-3)
        np.testing.assert_allclose(comb.cpu().numpy(), 0.5 + 0.5 * ref_r, rtol=RTOL, atol=1e-3)


def test_reset_done_restores_flagged_envs_only():
    from emloco_b200.rollout import Rollout
    n = 32
    R = Rollout(n, seed=5)
    noise = torch.zeros(n, 69, device="cuda")
    for k in range(3):
        R.step(k, noise=noise)
    torch.cuda.synchronize()
    before_rb = R.sim.rb_state.clone(); before_prog = R.sim.progress.clone(); before_amp = R.sim.amp_obs.clone()
    R.sim.reset.zero_(); R.sim.reset[[1, 7]] = 1
    R.sim.reset_done(R.init_root, R.init_dof)
    torch.cuda.synchronize()
    ch = (R.sim.rb_state.view(n, -1) != before_rb.view(n, -1)).any(1).cpu().numpy()
    assert ch.tolist() == [i in (1, 7) for i in range(n)]
    prog = R.sim.progress.cpu().numpy()
    assert prog[1] == 0 and prog[7] == 0 and np.all(np.delete(prog, [1, 7]) == before_prog.cpu().numpy()[0] )
    assert R.sim.reset.sum().item() == 0
    amp = R.sim.amp_obs.cpu().numpy()
    assert np.all(amp[1] == amp[1, :1]) and np.all(amp[7] == amp[7, :1])           # history filled with the current step
    np.testing.assert_array_equal(np.delete(amp, [1, 7], 0), np.delete(before_amp.cpu().numpy(), [1, 7], 0))
    np.testing.assert_allclose(R.sim.root_state[1].cpu().numpy(), R.init_root[1].cpu().numpy(), atol=1e-6)
    R.close()


def test_play_steps_horizon_and_gae_consistency():
    """Full 32-step horizon at a small size: finite outputs, GAE recurrence holds on the stored rows, and the
    post-horizon discriminator pass reproduces the per-step AMP rewards bit for bit (same rows, same weights)."""
    from emloco_b200.rollout import Rollout
    n = 128
    R = Rollout(n, seed=2, horizon=32)
    per_step = []
    for k in range(32):
        R.step(k)
        per_step.append(R.mb["amp_rewards"][k].clone())
    out = R.finish()
    torch.cuda.synchronize()
    for k, v in out.items():
        assert torch.isfinite(v).all(), k
    np.testing.assert_array_equal(torch.stack(per_step).cpu().numpy(), out["amp_rewards"].cpu().numpy())
    d, v, r, nv, adv = (out[k].cpu().numpy().reshape(32, n) for k in ("dones", "values", "rewards", "next_values", "advantages"))
    last = np.zeros(n, np.float32)
    for t in reversed(range(32)):
        last = (r[t] + 0.99 * nv[t] - v[t]) + 0.99 * 0.95 * (1 - d[t]) * last
        np.testing.assert_allclose(adv[t], last, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["returns"].cpu().numpy().reshape(32, n), adv + v, rtol=1e-6, atol=1e-6)
    assert 0 < R.locoval_scores.min().item() and R.locoval_scores.max().item() < 1
    R.close()


def test_merged_schedule_is_bit_identical_to_the_step_by_step_schedule():
    """The merged schedule of the graphed steps (ONE 12-layer launch for the policy pass of step n and the critic / discriminator
    pass of step n-1, bookkeeping beside the physics step, `finish` completing the last step) against the eager step-by-step
    schedule: every experience row, the post-horizon outputs and the simulator state are identical, bit for bit, over several
    horizons with resets and trajectory regeneration."""
    import bench
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    from emloco_b200.synthetic import synthetic_traj_pool
    n, T = 200, 6
    torch.manual_seed(4)
    net = AMPSeptValueNetwork()
    pool = synthetic_traj_pool(64, 0)
    kw = dict(seed=9, net=net, tensor_cores=True, horizon=T, traj_flags=bench.TRAJ_FLAGS, traj_pool=pool, sim_cfg=dict(episode_length=9))
    A = Rollout(n, merged=False, **kw)
    B = Rollout(n, **kw)
    assert B.merged and not A.merged
    for k in range(T):
        A.step(k); B.step(k)            # eager warm-up on both (identical generators -> identical noise)
    A.finish(); B.finish()
    resets = 0
    for rep in range(3):
        for k in range(T):
            A.step(k)
            B.step_graphed(k)
            resets += int(A.sim.reset.sum())
        oa, ob = A.finish(), B.finish_graphed()
        torch.cuda.synchronize()
        for key in ("obses", "actions", "neglogpacs", "mus", "values", "task_values", "rewards", "task_rewards", "dones", "next_values",
                    "amp_rewards", "amp_obs", "flip_obs", "returns", "advantages"):
            np.testing.assert_array_equal(oa[key].cpu().numpy(), ob[key].cpu().numpy(), err_msg=f"{key} rep {rep}")
        np.testing.assert_array_equal(A.state.cpu().numpy(), B.state.cpu().numpy())
    assert resets > 0
    np.testing.assert_array_equal(A.sim.rb_state.cpu().numpy(), B.sim.rb_state.cpu().numpy())
    assert any(isinstance(k, int) for k in B._graphs) and B._pending is None
    # flush() completes an outstanding step on demand
    B.step_graphed(0); A.step(0)
    assert B._pending == 0
    B.flush()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(A.mb["next_values"][0].cpu().numpy(), B.mb["next_values"][0].cpu().numpy())
    np.testing.assert_array_equal(A.mb["amp_rewards"][0].cpu().numpy(), B.mb["amp_rewards"][0].cpu().numpy())
    A.close(); B.close()


def test_graphed_steps_equal_eager_steps():
    """CUDA-graph replay of a step must produce exactly what the eager launches produce (same kernels, same order)."""
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    n = 96
    torch.manual_seed(4)
    net = AMPSeptValueNetwork()
    A = Rollout(n, seed=9, net=net, tensor_cores=True, horizon=4)
    B = Rollout(n, seed=9, net=net, tensor_cores=True, horizon=4)
    for k in range(4):
        A.step(k); B.step(k)            # eager warm-up on both (identical generators -> identical noise)
    for rep in range(2):
        for k in range(4):
            A.step(k)
            B.step_graphed(k)
        oa, ob = A.finish(), B.finish_graphed()
        torch.cuda.synchronize()
        for key in ("obses", "actions", "values", "rewards", "dones", "amp_rewards", "returns", "advantages"):
            np.testing.assert_array_equal(oa[key].cpu().numpy(), ob[key].cpu().numpy(), err_msg=f"{key} rep {rep}")
    np.testing.assert_array_equal(A.sim.rb_state.cpu().numpy(), B.sim.rb_state.cpu().numpy())
    A.close(); B.close()


def test_fused_sinks_equal_separate_copies_and_splits():
    """Post-step sinks (experience rows + normalised bf16 operands written by the post-step / reset kernels) against the
    unfused path (copy + split launches): same experience up to the 1-ulp difference between x*inv_std and x/std."""
    from emloco_b200.policy import AMPSeptValueNetwork, RunningMeanStd
    from emloco_b200.rollout import Rollout
    n = 160
    torch.manual_seed(6)
    net = AMPSeptValueNetwork()
    rng = np.random.default_rng(0)

    def norms():
        on, an = RunningMeanStd(1422), RunningMeanStd(3090)
        on.running_mean.copy_(torch.from_numpy(rng.normal(0, 0.3, 1422))); on.running_var.copy_(torch.from_numpy(rng.uniform(0.3, 2, 1422)))
        an.running_mean.copy_(torch.from_numpy(rng.normal(0, 0.3, 3090))); an.running_var.copy_(torch.from_numpy(rng.uniform(0.3, 2, 3090)))
        return on, an
    rng = np.random.default_rng(0); on_a, an_a = norms()
    rng = np.random.default_rng(0); on_b, an_b = norms()
    A = Rollout(n, seed=3, net=net, tensor_cores=True, horizon=6, fuse_sinks=True, obs_norm=on_a, amp_norm=an_a)
    B = Rollout(n, seed=3, net=net, tensor_cores=True, horizon=6, fuse_sinks=False, obs_norm=on_b, amp_norm=an_b)
    for rep in range(2):
        oa, ob = A.play_steps(), B.play_steps()
        torch.cuda.synchronize()
        # step 0 sees identical states: identical rows; later rows drift by the 1-ulp operand difference fed through the physics
        np.testing.assert_array_equal(oa["obses"][0].cpu().numpy(), ob["obses"][0].cpu().numpy()) if rep == 0 else None
        if rep == 0:
            np.testing.assert_allclose(oa["obses"][..., :398].cpu().numpy(), ob["obses"][..., :398].cpu().numpy(), rtol=1e-3, atol=5e-3)
            np.testing.assert_allclose(oa["amp_obs"].cpu().numpy(), ob["amp_obs"].cpu().numpy(), rtol=1e-3, atol=5e-3)
            np.testing.assert_allclose(oa["flip_obs"][..., :398].cpu().numpy(), ob["flip_obs"][..., :398].cpu().numpy(), rtol=1e-3, atol=5e-3)
            fo, oo = oa["flip_obs"].cpu().numpy(), oa["obses"].cpu().numpy()
            np.testing.assert_array_equal(fo[:-1, :, 369:398:2], -oo[1:, :, 369:398:2])     # row n of flip_obs mirrors the obs AFTER step n
            for key in ("mus", "values", "next_values", "amp_rewards", "returns"):
                np.testing.assert_allclose(oa[key].cpu().numpy(), ob[key].cpu().numpy(), rtol=1e-3, atol=5e-3, err_msg=key)
        assert all(torch.isfinite(v).all() for v in oa.values())
    A.close(); B.close()


def test_vec_env_surface_matches_reference_adapter_contract():
    """RLGPUEnv-shaped adapter (run.py:135-182, vec_task.py:125-134): shapes, dtypes, aliasing and reset semantics."""
    from emloco_b200.vec_env import RLGPUEnv
    n = 48
    env = RLGPUEnv(n, seed=1)
    info = env.get_env_info()
    assert info["observation_space"].shape == (1422,) and info["action_space"].shape == (69,) and info["amp_observation_space"].shape == (3090,)
    assert env.get_number_of_agents() == 1
    obs0 = env.reset()
    assert obs0.shape == (n, 1422) and obs0.is_cuda
    a = torch.zeros(n, 69, device="cuda")
    for _ in range(3):
        obs, rew, done, infos = env.step(a)
    assert obs.shape == (n, 1422) and rew.shape == (n,) and done.shape == (n,) and done.dtype == torch.int64
    assert infos["amp_obs"].shape == (n, 3090) and infos["terminate"].dtype == torch.int64 and infos["reward_raw"].shape == (n, 2)
    assert (env.sim.progress == 3).all()
    env.reset(torch.tensor([0, 5]))
    prog = env.sim.progress.cpu().numpy()
    assert prog[0] == 0 and prog[5] == 0 and (np.delete(prog, [0, 5]) == 3).all()
    w, p, v = env.get_waypoint_traj(), env.get_init_pose(), env.get_init_vel()
    assert w.shape == (n, 13, 2) and p.shape == (n, 24, 3) and v.shape == (n, 2)
    assert float(w[:, 0].abs().max()) == 0 and float(p[:, 0].abs().max()) == 0
    # rl_device = cpu moves obs / rewards / dones like `.to(self.rl_device)` does
    env2 = RLGPUEnv(8, rl_device="cpu")
    o, r, d, _ = env2.step(torch.zeros(8, 69))
    assert not o.is_cuda and not r.is_cuda and not d.is_cuda
    env.close(); env2.close()


def test_gym_shim_tensor_api():
    """The gymapi / gymtorch subset: acquire -> wrap aliases sim memory, PD targets + simulate x2 advance the
    state exactly like one fused emloco_step, indexed state sets re-read the aliases."""
    from emloco_b200 import gym_shim as G
    from emloco_b200.sim import EmlocoSim
    n = 40
    gym = G.acquire_gym()
    sim = gym.create_sim(0, num_envs=n, sim_params=G.SimParams())
    gym.prepare_sim(sim)
    root = G.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
    dof = G.wrap_tensor(gym.acquire_dof_state_tensor(sim))
    rb = G.wrap_tensor(gym.acquire_rigid_body_state_tensor(sim))
    assert root.shape == (n, 13) and dof.shape == (n * 69, 2) and rb.shape == (n * 24, 13)
    assert root.data_ptr() == sim.root_state.data_ptr()                      # alias, not a copy
    root[:, 2] = 1.5; root[:, 6] = 1.0
    ids = torch.arange(n, dtype=torch.int32, device="cuda")
    gym.set_actor_root_state_tensor_indexed(sim, G.unwrap_tensor(root), G.unwrap_tensor(ids), n)
    gym.set_dof_state_tensor_indexed(sim, G.unwrap_tensor(dof), G.unwrap_tensor(ids), n)
    torch.cuda.synchronize()
    assert torch.allclose(rb.view(n, 24, 13)[:, 0, 2], torch.full((n,), 1.5, device="cuda"))
    # reference sequence: set targets, simulate x controlFrequencyInv, fetch  ==  one emloco_step with the matching actions
    ref = EmlocoSim(n)
    ref.root_state.copy_(root); ref.reset_indexed(None)
    act = torch.rand(n, 69, device="cuda") * 0.2 - 0.1
    ref.step(act)
    gym.set_dof_position_target_tensor(sim, G.unwrap_tensor(ref.pd_target.clone()))
    for _ in range(2):
        gym.simulate(sim)
    gym.fetch_results(sim, True)
    gym.refresh_rigid_body_state_tensor(sim)
    np.testing.assert_allclose(rb.cpu().numpy(), ref.rb_state.cpu().numpy(), rtol=1e-4, atol=1e-5)
    with pytest.raises(Exception):
        G.unwrap_tensor(torch.zeros(4, 4, device="cuda").t())
    assert gym.get_asset_dof_count() == 69 and gym.find_actor_rigid_body_handle(sim, name="Head") == 13
    gym.destroy_sim(sim); ref.close()


def test_batched_locoval_filter_equals_batch_of_one_loop():
    """evaluate_jta.py:298-340: S scenes x M modes scored in one launch == the batch-of-1 loop, then the same filter."""
    from emloco_b200.formats import score_and_filter
    from emloco_b200.value_pose_net import ValuePoseNet
    torch.manual_seed(0)
    S, M = 300, 5
    net = ValuePoseNet(True, True).cuda().eval()
    trajs = (torch.randn(S, M, 13, 2, device="cuda") * 0.4).cumsum(2); trajs[:, :, 0] = 0
    pose = torch.randn(S, 24, 3, device="cuda") * 0.3
    vel = torch.randn(S, 2, device="cuda")
    values, keep = score_and_filter(net, trajs, pose, vel, threshold=0.5)
    assert values.shape == (S, M) and keep.shape == (S, M) and keep.any(1).all()
    for s in range(0, S, 37):
        for m in range(M):
            v, _ = net.calc_embodied_motion_loss(trajs[s, m][None].contiguous(), pose[s][None].clone(), vel[s][None].clone())
            assert abs(v.item() - values[s, m].item()) < 2e-4          # batch-of-1 goes to the CUDA-core kernel, the batch to tcgen05
    assert net.mutate_pose is True


def test_value_reuse_is_bit_identical_to_the_second_critic_pass():
    """reuse_values: critic(next obs) taken from the next step's policy pass (non-reset envs), 0 (terminated envs) or a compact
    critic pass over the timed-out envs - every experience tensor must equal the plain two-pass rollout bit for bit.
    Short episodes (episode_length 6) make time-outs frequent so that the compact path is exercised."""
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    n = 200
    torch.manual_seed(8)
    net = AMPSeptValueNetwork()
    with torch.no_grad():
        net.value.bias.fill_(0.3)
    A = Rollout(n, seed=4, net=net, tensor_cores=True, horizon=8, reuse_values=True)
    B = Rollout(n, seed=4, net=net, tensor_cores=True, horizon=8, reuse_values=False)
    for R in (A, B):
        R.sim.close()
    # same sims but with a 6-step episode so that `reset && !terminate` happens every few steps
    from emloco_b200.sim import EmlocoSim
    for R in (A, B):
        R.sim = EmlocoSim(n, episode_length=6)
        R.sim.traj_verts.copy_(torch.from_numpy(__import__("emloco_b200.synthetic", fromlist=["x"]).synthetic_env_state(n, seed=4, root_height=R.sim.rest_height)["verts"]).cuda())
        R.sim.reset.fill_(1)
        if R.fuse:
            R.sim.set_post_sinks(R.nets.post_sinks(obs_copy=R.mb["obses"][R.T]))
        R.sim.reset_done(R.init_root, R.init_dof)
    assert A.reuse_values and not B.reuse_values
    timeouts = 0
    for rep in range(3):
        oa, ob = A.play_steps(), B.play_steps()
        torch.cuda.synchronize()
        d = ob["dones"].cpu().numpy()
        timeouts += int(d.sum())
        for key in ("obses", "actions", "values", "next_values", "rewards", "dones", "amp_rewards", "returns", "advantages"):
            np.testing.assert_array_equal(oa[key].cpu().numpy(), ob[key].cpu().numpy(), err_msg=f"{key} rep {rep}")
    assert timeouts > 50, "the test must exercise resets"
    A.close(); B.close()


def test_host_observation_step_matches_device_step():
    """The vec-env boundary with host buffers (rl_device = cpu, vec_task.py:125-134): sim.obs is overwritten with the
    observations the host hands back (here: the same values after a D2H/H2D round trip) and the policy operands are re-derived
    from it; the split replay (env step | read-back hook | critic, discriminator, bookkeeping) does the same work as one graph."""
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    n, T = 96, 4
    torch.manual_seed(8)
    net = AMPSeptValueNetwork()
    A = Rollout(n, seed=4, net=net, tensor_cores=True, horizon=T)
    B = Rollout(n, seed=4, net=net, tensor_cores=True, horizon=T)
    C_ = Rollout(n, seed=4, net=net, tensor_cores=True, horizon=T)
    g = torch.Generator(device="cuda").manual_seed(1)
    seen = []
    for rep in range(2):                      # rep 0 warms up (eager), rep 1 replays graphs in B and C_
        for k in range(T):
            noise = torch.randn(n, 69, device="cuda", generator=g)
            A.step(k, noise=noise)
            for R, hook in ((B, None), (C_, lambda: seen.append(C_.sim.obs.clone()))):
                h = R.sim.obs.cpu()                                        # D2H
                R.sim.obs.copy_(h.cuda())                                  # H2D
                R.noise.copy_(noise)
                if rep == 0:
                    R.step(k, noise=R.noise, host_obs=True)
                else:
                    R.step_graphed_host_noise(k, after_env_step=hook)
        oa, ob, oc = A.finish(), B.finish(), C_.finish()
        torch.cuda.synchronize()
        for key in ("obses", "actions", "mus", "values", "next_values", "rewards", "amp_rewards", "dones", "returns", "flip_obs"):
            a, b, c = oa[key].cpu().numpy(), ob[key].cpu().numpy(), oc[key].cpu().numpy()
            np.testing.assert_array_equal(b, c, err_msg=key)               # one graph == split graphs
            if key == "dones":
                np.testing.assert_array_equal(a, b)
            else:
                np.testing.assert_allclose(a[0], b[0], rtol=1e-3, atol=1e-3, err_msg=key)      # first step: same state, 1-ulp operand difference
                np.testing.assert_allclose(a, b, rtol=2e-2, atol=5e-2, err_msg=key)            # later steps drift through the physics
    assert len(seen) == T                                                  # the hook ran between the two halves of every step
    np.testing.assert_array_equal(seen[-1].cpu().numpy(), C_.sim.obs.cpu().numpy())   # ... after post_step wrote the observations
    A.close(); B.close(); C_.close()


def test_64_envs_100_steps_lockstep_with_cpu_oracle():
    """BASELINE configs[0] (64 envs, 100 steps): every one of the 100 control steps is checked against the CPU oracle started
    from the GPU's own state of that step (re-synchronised each step, so fp32-vs-fp64 round-off cannot accumulate through the
    chaotic contact dynamics): 6 400 env-steps over the state distribution a rollout actually visits - standing, stumbling,
    fallen, reset.  Tolerances: 1e-3 relative, absolute floors per quantity; masks may differ only where the fp32 and fp64
    contact forces straddle the 50 N fall threshold (budget: 3 of 6 400)."""
    n, steps = 64, 100
    gpu, cpu = _pair(n, seed=5, tensor_cores=True)
    T = gpu.T
    sim = gpu.sim
    rng = np.random.default_rng(1)
    worst = {}
    mask_mismatch = 0
    n_resets = 0

    def upd(key, a, b, atol):
        err = np.abs(a - b) / (atol + RTOL * np.abs(b))
        worst[key] = max(worst.get(key, 0.0), float(err.max()))

    for k in range(steps):
        torch.cuda.synchronize()
        f64 = lambda t: t.detach().cpu().numpy().astype(np.float64)
        cpu.root = f64(sim.root_state).reshape(n, 13).copy()
        cpu.jq = f64(sim.joint_quat).reshape(n, 23, 4).copy()
        cpu.jw = f64(sim.dof_state).reshape(n, 69, 2)[..., 1].copy()
        cpu.progress = sim.progress.cpu().numpy().copy(); cpu.reset = sim.reset.cpu().numpy().copy()
        cpu.terminate = sim.terminate.cpu().numpy().copy()
        # the ring lives in the experience row of the previous step (rows_only sinks); after the initial reset in sim.amp_obs
        ring = sim.amp_obs if k == 0 else gpu.mb["amp_obs"][(k - 1) % T]
        cpu.amp_buf = ring.cpu().numpy().reshape(n, 15, 206).copy()
        cpu.contact = f64(sim.contact).reshape(n, 24, 3).copy(); cpu.dof_force = f64(sim.dof_force).reshape(n, 69).copy()
        cpu.obs = sim.obs.cpu().numpy().copy()
        cpu.state = gpu.state.cpu().numpy().copy()
        n_resets += int(cpu.reset.sum())
        noise = rng.standard_normal((n, 69)).astype(np.float32)
        o = cpu.step(noise)
        slot = k % T
        gpu.step(slot, noise=torch.from_numpy(noise).cuda())
        torch.cuda.synchronize()
        mb = {key: v[slot].cpu().numpy() for key, v in gpu.mb.items() if v is not None}
        upd("mus", mb["mus"], o["mu"], 5e-5); upd("actions", mb["actions"], o["actions"], 5e-5)
        upd("values", mb["values"][:, 0], o["values"], 5e-5)
        rb = sim.rb_state.view(n, 24, 13).cpu().numpy()
        upd("rb_pos", rb[..., 0:3], o["rb"][..., 0:3], 1e-3); upd("rb_rot", rb[..., 3:7], o["rb"][..., 3:7], 1e-3)
        upd("rb_vel", rb[..., 7:13], o["rb"][..., 7:13], 1e-2)
        same = mb["dones"] == o["dones"]
        mask_mismatch += int((~same).sum())
        upd("rewards", mb["rewards"][same, 0], o["rewards"][same], 5e-3)
        upd("next_values", mb["next_values"][same, 0], o["next_values"][same], 1e-4)
        upd("amp_rewards", mb["amp_rewards"][:, 0], o["amp_rewards"], 1e-3)
        upd("self_obs", sim.obs.cpu().numpy()[:, :368], o["next_obs"][:, :368], 5e-3)
        if slot == T - 1:
            gpu.finish()
    print("lockstep worst error / tolerance:", {k_: round(v, 3) for k_, v in worst.items()}, "mask mismatches", mask_mismatch,
          "resets", n_resets)
    assert n_resets > n                       # beyond the initial reset: episodes ended and restarted inside the 100 steps
    assert mask_mismatch <= 3
    for key, v in worst.items():
        assert v <= 1.0, (key, v)
    gpu.close()


def test_rollout_finetunes_locoval_like_the_reference_block():
    """Rollout(finetune=True): after the bookkeeping of every control step LocoVal takes one AdamW step on the envs whose
    game_combined_rewards became non-zero (amp_continuous_value.py:122-146).  A twin rollout without fine-tuning supplies the
    per-step inputs for the fp64 oracle of that block; weights must agree after 12 steps with episodes ending at several
    different steps (and at one step for many envs at once)."""
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    from emloco_b200.value_pose_net import ValuePoseNet
    from oracle import oracle_np as O
    n, K = 96, 12
    torch.manual_seed(11)
    net = AMPSeptValueNetwork()
    torch.manual_seed(12)
    va, vb = ValuePoseNet(True, True, mutate_pose=False), ValuePoseNet(True, True, mutate_pose=False)
    vb.load_state_dict(va.state_dict())
    A = Rollout(n, seed=6, net=net, tensor_cores=True, horizon=K, valuenet=va, finetune=True, traj_flags=4)
    B = Rollout(n, seed=6, net=net, tensor_cores=True, horizon=K, valuenet=vb, finetune=False, traj_flags=4)
    for R in (A, B):
        R.sim.progress[: n // 3] = 160            # these time out inside the horizon (episode length 168)
        R.sim.progress[n // 3: n // 2] = 164
        R.state[1].fill_(140.0)                   # current_lengths: half of the envs cross step_to_pred = 144 at step 4 ...
        R.state[1][n // 2: 3 * n // 4] = 138.0    # ... a quarter at step 6 ...
        R.state[1][: n // 2] = 10.0               # ... and the time-outs end short episodes (done_early) at steps 3 and 7
    sd = {k: v.detach().cpu().numpy().copy() for k, v in vb.state_dict().items()}
    W = {k: [sd[f"_network.{k}.weight"], sd[f"_network.{k}.bias"]] for k in ("fc1", "fc2", "fc3")}
    opt = dict(step=0, m={k: [0.0, 0.0] for k in W}, v={k: [0.0, 0.0] for k in W})
    g = torch.Generator(device="cuda").manual_seed(3)
    used = 0
    for k in range(K):
        noise = torch.randn(n, 69, device="cuda", generator=g)
        B.step(k, noise=noise)
        torch.cuda.synchronize()
        gc = B.state[4].cpu().numpy().copy()
        _, _, _, cnt = O.locoval_finetune_step(W, opt, B.waypoint_traj.cpu().numpy(), B.init_pose.cpu().numpy().copy(),
                                                B.init_vel.cpu().numpy(), gc)
        used += cnt
        B.state[4].zero_()                        # :145
        A.step(k, noise=noise)
    torch.cuda.synchronize()
    assert used >= n and opt["step"] >= 4         # four separate optimiser steps, every env consumed once
    assert int(A.valuenet._ft["step"].item()) == opt["step"]
    assert float(A.state[4].abs().sum()) == 0.0
    got = {k: v.detach().cpu().numpy() for k, v in A.valuenet.state_dict().items()}
    for k in W:
        w, ref = got[f"_network.{k}.weight"], W[k][0]
        if k == "fc1":
            w, ref = np.delete(w, [3, 99], 1), np.delete(ref, [3, 99], 1)      # round-off-noise inputs, see test_gpu_parity
        np.testing.assert_allclose(w, ref, rtol=RTOL, atol=2e-5, err_msg=k)
        np.testing.assert_allclose(got[f"_network.{k}.bias"], W[k][1], rtol=RTOL, atol=2e-5, err_msg=k)
    loss, pred, gt, cnt = A.valuenet.finetune_stats()
    assert cnt == used and 0.0 <= gt <= 1.5 and 0.0 < pred < 1.0
    A.close(); B.close()


def test_gym_shim_env_creation_sequence():
    """The env-creation calls of humanoid.py:643-946 (+ the terrain tri-mesh of ..terrain.py:866-877) recorded into a pending
    sim and realised by prepare_sim: same sim as one built directly from the same gains, start poses and height field."""
    from emloco_b200 import gym_shim as G
    from emloco_b200.mjcf import default_model
    from emloco_b200.model import build_model_arrays
    from emloco_b200.sim import EmlocoSim
    n = 12
    gym = G.acquire_gym()
    sim = gym.create_sim(0, -1, G.SIM_PHYSX, G.SimParams())                        # base_task.py:238
    hs = np.zeros((700, 700), np.int16); hs[520:540, :] = 60                       # a 0.3 m ridge across the patch
    x = np.arange(700) * 0.1
    yy, xx = np.meshgrid(x, x)
    verts = np.stack([xx.flatten(), yy.flatten(), hs.flatten() * 0.005], 1).astype(np.float32)
    tm = G.TriangleMeshParams(); tm.nb_vertices = verts.shape[0]; tm.static_friction = 1.0
    gym.add_triangle_mesh(sim, verts.flatten(), np.zeros(3, np.uint32), tm)
    opt = G.AssetOptions(); opt.angular_damping = 0.01; opt.max_angular_velocity = 100.0
    asset = gym.load_asset(sim, "data/assets/mjcf", "smpl_humanoid.xml", opt)
    pd_scale = asset.model.total_mass / 77.0                                       # humanoid.py:905-910
    for i in range(n):
        env = gym.create_env(sim, G.Vec3(-5, -5, 0), G.Vec3(5, 5, 5), 4)
        pose = G.Transform(); pose.p = G.Vec3(51.0 + 0.3 * i, 52.0 + 0.1 * i, 0.95); pose.r = G.Quat(0, 0, 0, 1)
        h = gym.create_actor(env, asset, pose, "humanoid", i, 0, 0)
        gym.enable_actor_dof_force_sensors(env, h)
        prop = gym.get_asset_dof_properties(asset)
        prop["driveMode"] = G.DOF_MODE_POS
        prop["stiffness"] *= pd_scale; prop["damping"] *= pd_scale
        gym.set_actor_dof_properties(env, h, prop)
    gym.prepare_sim(sim)                                                           # base_task.py:128
    root = G.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
    rb = G.wrap_tensor(gym.acquire_rigid_body_state_tensor(sim))
    assert root.shape == (n, 13)
    np.testing.assert_allclose(root[:, 0].cpu().numpy(), 51.0 + 0.3 * np.arange(n), rtol=1e-6)
    ref = EmlocoSim(n, model_arrays=build_model_arrays(default_model()))           # the default table carries the same mass scaling
    ref.set_height_field(hs)
    ref.root_state.copy_(root); ref.reset_indexed(None)
    np.testing.assert_allclose(np.asarray(sim.real.model_arrays["kp"]), np.asarray(ref.model_arrays["kp"]), rtol=1e-6)
    act = torch.rand(n, 69, device="cuda") * 0.2 - 0.1
    for _ in range(5):
        ref.step(act)
        gym.set_dof_position_target_tensor(sim, G.unwrap_tensor(ref.pd_target.clone()))
        for _ in range(2):
            gym.simulate(sim)
        gym.fetch_results(sim, True)
    np.testing.assert_allclose(rb.cpu().numpy(), ref.rb_state.cpu().numpy(), rtol=1e-4, atol=1e-5)
    assert gym.get_frame_count(sim) == 10 and gym.get_sim_params(sim).substeps == 2
    gym.destroy_sim(sim); ref.close()


def test_rows_only_sinks_give_the_same_experience():
    """rows_only (mirrored observation and AMP ring written once, into the experience rows; each env's ring is shifted out of
    the row its last post-step wrote, or out of sim.amp_obs after a reset) against the mode that also refreshes
    sim.flip_obs / sim.amp_obs: bit-identical experience over two horizons with resets, in a non-consecutive slot order too."""
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    n, T = 96, 6
    torch.manual_seed(2)
    net = AMPSeptValueNetwork()
    outs = []
    for ro in (True, False):
        R = Rollout(n, seed=9, net=net, tensor_cores=True, horizon=T, rows_only=ro)
        R.sim.progress[: n // 2] = 163                     # time-outs inside the first horizon
        g = torch.Generator(device="cuda").manual_seed(5)
        rows = []
        for rep in range(2):
            for k in range(T):
                R.step(k, noise=torch.randn(n, 69, device="cuda", generator=g))
            o = R.finish()
            torch.cuda.synchronize()
            rows.append({k_: o[k_].cpu().numpy().copy() for k_ in ("amp_obs", "flip_obs", "obses", "amp_rewards", "rewards", "dones", "returns")})
        # slots out of order (what bench.py's 8-slot segment loop does): 0, 1, 0, 1 ...
        for k in (0, 1, 0, 1, 2):
            R.step(k, noise=torch.randn(n, 69, device="cuda", generator=g))
        torch.cuda.synchronize()
        rows.append({"amp_obs": R.mb["amp_obs"][:3].cpu().numpy().copy(), "flip_obs": R.mb["flip_obs"][:3].cpu().numpy().copy()})
        if not ro:
            np.testing.assert_array_equal(R.sim.amp_obs.view(n, -1).cpu().numpy(), R.mb["amp_obs"][2].cpu().numpy())
        outs.append(rows)
        R.close()
    assert outs[0][0]["dones"].sum() >= n // 2
    for a, b in zip(*outs):
        for k_ in a:
            np.testing.assert_array_equal(a[k_], b[k_], err_msg=k_)


def test_rollout_from_reference_configuration():
    """the reference's YAML trees + run.py flags -> Rollout, one horizon with fine-tuning and the
    device-side trajectory reset switched on by them."""
    from emloco_b200.formats import kwargs_from_reference_cfg
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    from emloco_b200.synthetic import synthetic_traj_pool
    env_cfg = {"env": dict(numEnvs=48, episodeLength=168, controlFrequencyInv=2, power_coefficient=0.0005, location_coefficient=1,
                           trajSampleTimestep=0.4, stepToPred=144, speedMin=0.0005, speedMax=3.0, accelMax=2.0, sharpTurnProb=0.02,
                           hybridInitProb=0.5, numAMPObsSteps=15, numTrajSamples=15, pdControl=True, terrain=dict(staticFriction=1.0))}
    train_cfg = {"params": dict(
        network=dict(space=dict(continuous=dict(sigma_init=dict(val=-2.9), fixed_sigma=True, learn_sigma=False)),
                     mlp=dict(units=[2048, 1024]), task_mlp=dict(units=[512, 256]), value_mlp=dict(units=[15, 6]), disc=dict(units=[1024, 512])),
        config=dict(horizon_length=4, gamma=0.99, tau=0.95, task_reward_w=0.5, disc_reward_w=0.5, disc_reward_scale=2,
                    inversion_penalty_scale=0.3, normalize_value=True, player=dict(finetune=True)))}
    k = kwargs_from_reference_cfg(env_cfg, train_cfg, dict(real_path=True, adjust_root_vel=True, init_heading=True))
    torch.manual_seed(0)
    R = Rollout(k["num_envs"], net=AMPSeptValueNetwork(**k["net"]), sim_cfg=k["sim"], traj_cfg=k["traj"],
                traj_pool=synthetic_traj_pool(8, 0), tensor_cores=True, **k["rollout"])
    assert R.T == 4 and R.finetune and R.sim.cfg.episode_length == 168
    R.sim.progress.fill_(165)                      # everybody times out inside the horizon
    R.state[1].fill_(100.0)
    v0 = R.sim.traj_verts.clone()
    out = R.play_steps(graphed=False)
    torch.cuda.synchronize()
    assert out["obses"].shape == (4, 48, 1422) and out["returns"].shape == (4, 48, 1) and torch.isfinite(out["returns"]).all()
    assert out["dones"].sum() >= 48 and (R.sim.traj_verts != v0).any()               # resets regenerated trajectories
    assert R.valuenet.finetune_stats()[3] >= 48                                     # done_early episodes fed LocoVal
    R.close()


def test_graphed_rollout_follows_weight_and_normaliser_updates():
    """Captured graphs must not keep rolling out with the weights / statistics of capture time: after an
    optimiser-style in-place parameter update, a load_state_dict, a running-statistics update by in-place copy AND by attribute
    re-assignment (the reference's way, running_mean_std.py:93-96: fresh tensors with _version 0) and a value_mean_std update,
    graphed horizons equal eager horizons bit for bit; in-place updates do not re-capture."""
    from emloco_b200.policy import AMPSeptValueNetwork, RunningMeanStd
    from emloco_b200.rollout import Rollout
    n, T = 96, 3
    mk = lambda: (torch.manual_seed(4), AMPSeptValueNetwork())[1]
    A = Rollout(n, seed=9, net=mk(), tensor_cores=True, horizon=T, traj_flags=0)
    B = Rollout(n, seed=9, net=mk(), tensor_cores=True, horizon=T, traj_flags=0)
    keys = ("obses", "actions", "values", "next_values", "rewards", "dones", "amp_rewards", "returns", "advantages")

    def both(fn):
        for R in (A, B):
            fn(R)

    def horizon(tag):
        oa = A.play_steps(graphed=False); ob = B.play_steps(graphed=True)
        torch.cuda.synchronize()
        for k in keys:
            np.testing.assert_array_equal(oa[k].cpu().numpy(), ob[k].cpu().numpy(), err_msg=f"{k} after {tag}")
        return {k: oa[k].clone() for k in keys}          # the experience rows are reused by the next horizon

    horizon("warm-up"); horizon("capture")
    base = horizon("replay")
    captured = dict(B._graphs)
    g = torch.Generator().manual_seed(1)

    def perturb(R):
        with torch.no_grad():
            for p in R.net.parameters():
                if p.requires_grad:
                    p.add_(0.02 * torch.randn(p.shape, generator=torch.Generator().manual_seed(p.numel())).to(p.device))
    both(perturb)
    o1 = horizon("in-place parameter update")
    assert not np.array_equal(o1["actions"].cpu().numpy(), base["actions"].cpu().numpy())
    assert all(B._graphs.get(k) is captured[k] for k in captured), "an in-place update must not force a re-capture"

    def stats_inplace(R):
        R.obs_norm.running_mean.add_(0.05); R.obs_norm.running_var.mul_(1.3)
        R.amp_norm.running_mean.sub_(0.02); R.amp_norm.running_var.mul_(0.8)
        R.value_norm.running_mean.fill_(0.7); R.value_norm.running_var.fill_(2.5)
    both(stats_inplace)
    o2 = horizon("in-place statistics update")
    assert not np.array_equal(o2["values"].cpu().numpy(), o1["values"].cpu().numpy())
    assert all(B._graphs.get(k) is captured[k] for k in captured)

    def stats_reassign(R):      # twice: the second fresh tensor has the same _version (0) as the first
        for scale in (1.5, 0.6):
            R.obs_norm.running_mean = R.obs_norm.running_mean * scale + 0.01
            R.obs_norm.running_var = R.obs_norm.running_var * scale
            R.value_norm.running_mean = R.value_norm.running_mean + 0.25
            R.value_norm.running_var = R.value_norm.running_var * scale
    both(stats_reassign)
    o3 = horizon("statistics re-assigned")
    assert not np.array_equal(o3["values"].cpu().numpy(), o2["values"].cpu().numpy())

    sd = {k: v + 0.01 for k, v in A.net.state_dict().items()}
    both(lambda R: R.net.load_state_dict(sd))
    horizon("load_state_dict")

    def realloc(R):             # parameter storage replaced: pointers baked into the graphs are stale -> must re-capture
        R.net.mu.bias.data = R.net.mu.bias.data.clone() + 0.05
    both(realloc)
    horizon("parameter re-allocated")
    assert any(B._graphs.get(k) is not captured[k] for k in captured)
    A.close(); B.close()


def test_batched_filter_reference_compat_reproduces_the_sequential_in_place_loop():
    """evaluate_jta.py:298-302 scores prediction and ground truth of every mode on ONE init_pose view that
    ValuePoseNet rotates / zeroes in place, cumulatively.  score_and_filter(reference_compat=True) must give the values of
    that sequential loop (run here through the in-place drop-in, batch of 1, exactly like the reference) in a single launch."""
    from emloco_b200.formats import score_and_filter
    from emloco_b200.value_pose_net import ValuePoseNet
    torch.manual_seed(1)
    S, M = 40, 5
    net = ValuePoseNet(True, True).cuda().eval()
    with torch.no_grad():
        for m in net._network:
            if hasattr(m, "bias"):
                m.bias.uniform_(-0.2, 0.2)
    trajs = (torch.randn(S, M, 13, 2, device="cuda") * 0.4).cumsum(2); trajs[:, :, 0] = 0
    gts = (torch.randn(S, 13, 2, device="cuda") * 0.4).cumsum(1); gts[:, 0] = 0
    pose = torch.randn(S, 24, 3, device="cuda") * 0.3
    vel = torch.randn(S, 2, device="cuda")
    values, keep = score_and_filter(net, trajs, pose.clone(), vel, threshold=0.5, reference_compat=True, gt_trajs=gts)
    plain, _ = score_and_filter(net, trajs, pose.clone(), vel, threshold=0.5)
    seq = torch.zeros(S, M)
    with torch.no_grad():
        for s in range(S):
            p = pose[s].clone()                                        # the scene's init_pose tensor, mutated by every call
            for m in range(M):
                v, _ = net.calc_embodied_motion_loss(trajs[s, m][None].contiguous(), p.unsqueeze(0), vel[s][None])
                net.calc_embodied_motion_loss(gts[s][None].contiguous(), p.unsqueeze(0), vel[s][None])
                seq[s, m] = v.item()
    np.testing.assert_allclose(values.cpu().numpy(), seq.numpy(), rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(plain[:, 0].cpu().numpy(), seq[:, 0].numpy(), rtol=1e-3, atol=2e-4)     # mode 0 sees the original pose
    assert (plain[:, 1:].cpu() - seq[:, 1:]).abs().max() > 1e-2                                        # ... later modes do not
    with pytest.raises(ValueError):
        score_and_filter(net, trajs, pose, vel, reference_compat=True)


def test_benched_configuration_4096_envs_graphed_tensor_core_horizon_in_lockstep_with_cpu_oracle():
    """Parity ON THE CONFIGURATION bench.py TIMES: Rollout(4096, tensor_cores=True, rows_only=True,
    traj_flags=1|2|4, traj_pool=2048 polylines), CUDA-graph replay of every step (`step_graphed`) and of the post-horizon pass
    (`finish_graphed`), parallel branches, deferred trajectory reset.  A full 32-step horizon is replayed; at every step the
    first 256 envs are checked against the CPU oracle started from the GPU's own pre-step state of those envs (re-synchronised
    each step so fp32-vs-fp64 round-off cannot accumulate through the contact dynamics), using the policy noise the graph drew.
    After the horizon: AMP rewards of the post-horizon discriminator pass, combined rewards, GAE advantages and returns."""
    import bench
    from emloco_b200.model import build_model_arrays, rest_root_height
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    from emloco_b200.synthetic import synthetic_env_state, synthetic_traj_pool
    from oracle import oracle_np as O
    from oracle.cpu_rollout import CpuRollout, weights_from_state_dict
    N, n, T = 4096, 256, 32
    torch.manual_seed(0)
    net = AMPSeptValueNetwork()
    gpu = Rollout(N, seed=0, net=net, tensor_cores=True, traj_flags=bench.TRAJ_FLAGS, traj_pool=synthetic_traj_pool(bench.TRAJ_POOL, 0))
    assert gpu.rows_only and gpu.concurrent and gpu._traj_deferred and not gpu.reuse_values      # the bench defaults
    A = build_model_arrays()
    P, D = weights_from_state_dict(net.state_dict())
    st = synthetic_env_state(N, 0, rest_root_height(A))
    cpu = CpuRollout(A, {k: v[:n * 69] if k == "dof" else v[:n] for k, v in st.items()}, P, D)
    sim = gpu.sim
    # as bench.py: three eager steps, then every slot captured during one horizon, then replays
    for k in range(3):
        gpu.step(k)
    gpu.finish()
    for k in range(3, 3 + T):
        gpu.step_graphed(k % T)
        if k % T == T - 1:
            gpu.finish_graphed()
    for k in range(3):                      # slots 0..2 of the next horizon: the checked horizon starts at slot 0 below
        gpu.step_graphed(k)
    for k in range(3, T):
        gpu.step_graphed(k)
    gpu.finish_graphed()
    graphs_before = dict(gpu._graphs)
    assert all(k in graphs_before for k in range(T)) and "finish" in graphs_before

    worst, mask_mismatch, n_resets, n_inv = {}, 0, 0, 0

    def upd(key, a, b, atol):
        err = np.abs(a - b) / (atol + RTOL * np.abs(b))
        worst[key] = max(worst.get(key, 0.0), float(err.max()))

    f64 = lambda t: t.detach().cpu().numpy().astype(np.float64)
    assert gpu.merged, "the bench configuration runs the merged schedule"

    def check_bookkeeping(k, o, cpu_state):
        """Rows the bookkeeping kernel of step k writes - with the merged schedule they are complete one step later (or after finish)."""
        nonlocal mask_mismatch
        mb = {key: gpu.mb[key][k][:n].cpu().numpy() for key in ("dones", "rewards", "next_values", "amp_rewards", "values")}
        upd("values", mb["values"][:, 0], o["values"], 5e-5)
        same = mb["dones"] == o["dones"]
        mask_mismatch += int((~same).sum())
        upd("rewards", mb["rewards"][same, 0], o["rewards"][same], 5e-3)
        upd("next_values", mb["next_values"][same, 0], o["next_values"][same], 1e-4)
        upd("amp_rewards", mb["amp_rewards"][:, 0], o["amp_rewards"], 1e-3)
        upd("state", gpu.state[:, :n].cpu().numpy()[:, same], cpu_state[:, same], 3e-2)

    prev = None
    for k in range(T):
        torch.cuda.synchronize()
        cpu.root = f64(sim.root_state).reshape(N, 13)[:n].copy()
        cpu.jq = f64(sim.joint_quat).reshape(N, 23, 4)[:n].copy()
        cpu.jw = f64(sim.dof_state).reshape(N, 69, 2)[:n, :, 1].copy()
        cpu.progress = sim.progress[:n].cpu().numpy().copy(); cpu.reset = sim.reset[:n].cpu().numpy().copy()
        cpu.terminate = sim.terminate[:n].cpu().numpy().copy()
        ring = gpu.mb["amp_obs"][(k - 1) % T]                       # rows_only: the ring lives in the previous step's experience row
        cpu.amp_buf = ring[:n].cpu().numpy().reshape(n, 15, 206).copy()
        cpu.contact = f64(sim.contact).reshape(N, 24, 3)[:n].copy(); cpu.dof_force = f64(sim.dof_force).reshape(N, 69)[:n].copy()
        cpu.obs = (gpu.mb["obses"][T] if k == 0 else gpu.mb["obses"][k])[:n].cpu().numpy().copy()
        cpu.verts = sim.traj_verts[:n].cpu().numpy().copy()          # reset envs still see their OLD polyline in the reset observation
        n_resets += int(cpu.reset.sum())
        gpu.step_graphed(k)                                          # replay: step k + the outstanding part of step k-1
        torch.cuda.synchronize()
        if prev is not None:
            check_bookkeeping(k - 1, *prev)
        cpu.state = gpu.state[:, :n].cpu().numpy().copy()            # bookkeeping state after step k-1 = before step k
        noise = gpu.noise[:n].cpu().numpy().copy()
        cpu.reset_done()                                             # env_reset(done_indices) with the old polylines ...
        cpu.verts = sim.traj_verts[:n].cpu().numpy().copy()          # ... then _reset_task: the device's regenerated polylines (Philox
        cpu.inverted = gpu.inverted[:n].cpu().numpy().astype(bool)   #     draws; the generator itself is pinned by traj_reset_*.npz)
        n_inv += int(cpu.inverted.sum())
        o = cpu.step(noise)
        prev = (o, cpu.state.copy())
        mb = {key: gpu.mb[key][k][:n].cpu().numpy() for key in ("obses", "mus", "actions", "neglogpacs", "values", "task_values", "amp_obs")}
        upd("obs_in", mb["obses"][:, :398], o["obs"][:, :398], 5e-3)
        upd("mus", mb["mus"], o["mu"], 5e-5); upd("actions", mb["actions"], o["actions"], 5e-5)
        upd("neglogp", mb["neglogpacs"], o["neglogp"], 1e-3)
        upd("task_values", mb["task_values"], o["task_values"], 5e-5)
        rb = sim.rb_state.view(N, 24, 13)[:n].cpu().numpy()
        upd("rb_pos", rb[..., 0:3], o["rb"][..., 0:3], 1e-3); upd("rb_rot", rb[..., 3:7], o["rb"][..., 3:7], 1e-3)
        upd("rb_vel", rb[..., 7:13], o["rb"][..., 7:13], 1e-2)
        upd("self_obs", gpu.mb["obses"][k + 1][:n, :368].cpu().numpy(), o["next_obs"][:, :368], 5e-3)
        upd("amp_row", mb["amp_obs"][:, :206], o["amp_obs"][:, :206], 5e-3)
    out = gpu.finish_graphed()                                        # replay: the outstanding part of the last step + the post-horizon pass
    torch.cuda.synchronize()
    check_bookkeeping(T - 1, *prev)
    assert all(gpu._graphs[k] is graphs_before[k] for k in graphs_before), "the checked horizon must be pure replay"
    g = {k: out[k][:, :n].cpu().numpy() for k in ("amp_obs", "task_rewards", "amp_rewards", "rewards", "dones", "values", "next_values",
                                                   "returns", "advantages")}
    amp_r = np.stack([O.disc_reward(g["amp_obs"][t], cpu.D, 2.0)[0] for t in range(T)])
    upd("post_amp_rewards", g["amp_rewards"], amp_r, 1e-3)
    comb = (np.float32(0.5) * g["task_rewards"] + np.float32(0.5) * amp_r).astype(np.float32)
    upd("combined", g["rewards"], comb, 1e-3)
    adv = O.discount_values(g["dones"], g["values"], g["rewards"], g["next_values"])
    upd("advantages", g["advantages"], adv, 1e-4); upd("returns", g["returns"], adv + g["values"], 1e-4)
    print("benched-config lockstep worst error / tolerance:", {k_: round(v, 3) for k_, v in worst.items()}, "mask mismatches",
          mask_mismatch, "resets", n_resets)
    assert n_resets > 20                      # episodes ended and restarted inside the checked horizon
    assert mask_mismatch <= 3
    for key, v in worst.items():
        assert v <= 1.0, (key, v)
    gpu.close()


def test_step_host_c_abi_entry_equals_the_device_step():
    """emloco_step_host (include/emloco.h: the vec-env call of a host-side user, run.py:148-160): host buffers in and out,
    same results as emloco_step on device tensors - obs, rewards, int64 reset mask and AMP observations bit for bit."""
    import ctypes as C
    from emloco_b200 import _lib
    from emloco_b200.model import rest_root_height
    from emloco_b200.sim import EmlocoSim
    from emloco_b200.synthetic import synthetic_env_state
    n = 80
    sims = [EmlocoSim(n) for _ in range(2)]
    st = synthetic_env_state(n, 3, rest_root_height(sims[0].model_arrays))
    root, dof = torch.from_numpy(st["root"]).cuda(), torch.from_numpy(st["dof"]).cuda()
    for s in sims:
        s.traj_verts.copy_(torch.from_numpy(st["verts"]).cuda())
        s.reset.fill_(1)
        s.reset_done(root, dof)
    rng = np.random.default_rng(0)
    h_obs, h_rew = np.zeros((n, 1422), np.float32), np.zeros(n, np.float32)
    h_reset, h_amp = np.zeros(n, np.int64), np.zeros((n, 3090), np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for k in range(6):
        act = rng.uniform(-0.5, 0.5, (n, 69)).astype(np.float32)
        sims[0].step(torch.from_numpy(act).cuda())
        torch.cuda.synchronize()
        _lib.check(_lib.load().emloco_step_host(sims[1]._h, vp(act), vp(h_obs), vp(h_rew), vp(h_reset), vp(h_amp)), "emloco_step_host")
        np.testing.assert_array_equal(h_obs, sims[0].obs.cpu().numpy(), err_msg=f"obs step {k}")
        np.testing.assert_array_equal(h_rew, sims[0].rew.cpu().numpy())
        np.testing.assert_array_equal(h_reset, sims[0].reset.cpu().numpy())
        np.testing.assert_array_equal(h_amp, sims[0].amp_obs.reshape(n, 3090).cpu().numpy())
    # optional outputs may be NULL; a NULL action pointer is an error, reported through emloco_last_error
    assert _lib.load().emloco_step_host(sims[1]._h, vp(act), None, None, None, None) == 0
    assert _lib.load().emloco_step_host(sims[1]._h, None, vp(h_obs), None, None, None) != 0
    assert b"emloco_step_host" in _lib.load().emloco_last_error()
    for s in sims:
        s.close()


def test_host_pipeline_groups_equal_plain_host_steps():
    """HostRolloutPipeline (the end-to-end path bench.py times): two env groups in flight, host buffers in and out.  Every
    group's results must equal, bit for bit, those of a plain Rollout with the group's seed stepped serially on the same
    host-provided observations and noise."""
    from emloco_b200.host_pipeline import HostRolloutPipeline
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    N, G, T = 192, 2, 4
    mk = lambda: (torch.manual_seed(2), AMPSeptValueNetwork())[1]
    pipe = HostRolloutPipeline(N, groups=G, seed=5, horizon=T, tensor_cores=True, net=mk(), traj_flags=0)
    refs = [Rollout(N // G, seed=5 + 7919 * g, horizon=T, tensor_cores=True, net=mk(), traj_flags=0) for g in range(G)]
    assert pipe.h2d_bytes_per_step == N * (1422 + 69) * 4 and pipe.d2h_bytes_per_step == N * (1422 + 1 + 2 + 69 + 1 + 1) * 4
    for g in range(G):
        pipe.host[g]["noise"].copy_(torch.randn(N // G, 69, generator=torch.Generator().manual_seed(g)))
    feed = [pipe.host[g]["obs_in"].clone() for g in range(G)]
    for k in range(2 * T + 1):                 # eager first steps, graph capture, replay - all compared
        for g in range(G):
            pipe.submit(g)
        for g in range(G):
            h = pipe.wait(g)
            R = refs[g]
            R.sim.obs.copy_(feed[g].cuda()); R.noise.copy_(h["noise"].cuda())
            R.step(k % T, noise=R.noise, host_obs=True)
            if k % T == T - 1:
                R.finish()
            torch.cuda.synchronize()
            np.testing.assert_array_equal(h["obs_in"].numpy(), R.sim.obs.cpu().numpy(), err_msg=f"obs step {k} group {g}")   # swapped: results are the next input
            np.testing.assert_array_equal(h["rew"].numpy(), R.sim.rew.cpu().numpy())
            np.testing.assert_array_equal(h["reset"].numpy(), R.sim.reset.cpu().numpy())
            np.testing.assert_array_equal(h["actions"].numpy(), R.mb["actions"][k % T].cpu().numpy())
            np.testing.assert_array_equal(h["values"].numpy(), R.mb["values"][k % T].cpu().numpy())
            feed[g] = h["obs_in"].clone()
    pipe.close()
    for R in refs:
        R.close()


def test_player_record_kernel_matches_reference_loop_golden_and_oracle():
    """SURVEY component 14: emloco_player_record against the step loop of AMPPlayerContinuousValue.run executed from the
    reference (player.npz, three games played as three parallel envs) and against the oracle on 500 envs x 40 random steps
    with both reward modes."""
    import os
    from conftest import GOLDEN
    from emloco_b200 import _lib
    from emloco_b200.sim import _ptr, _stream
    from oracle import oracle_np as O
    g = np.load(os.path.join(GOLDEN, "player.npz"))
    T = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt).cuda()

    def run(N, steps, gen, plot, stp):
        st = torch.zeros(11, N, device="cuda"); st[2] = 1
        names = ("n", "cr", "coef", "pred", "cr_to_pred", "c_loc", "c_pow", "c_disc", "loc_to_pred", "pow_to_pred", "disc_to_pred")
        ost = {k: np.zeros(N, np.float32) for k in names}; ost["coef"][:] = 1
        res, cnt = torch.zeros(4096, 8, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
        ref = []
        for n in range(steps):
            rew, raw, reset, logit, scores, inv = gen(n)
            d = [T(rew), T(raw), T(reset, torch.int64), T(logit), T(scores), T(inv, torch.uint8)]      # kept alive across the launch
            _lib.check(_lib.load().emloco_player_record(*[_ptr(t) for t in d], _ptr(st), N, _ptr(res), _ptr(cnt), 4096, int(plot), 0.3, 2.0, 0.99,
                                                        stp, -10.0, 100.0, _stream()), "emloco_player_record")
            torch.cuda.synchronize()
            o = O.player_record(ost, rew, raw, reset, logit, scores, inv, plot_val_reward=plot, step_to_pred=stp)
            ref += [(e, o["pred"][j], o["cr_to_pred"][j], o["norm_reward"][j], o["c_loc"][j], o["c_pow"][j], o["c_disc"][j], o["steps"][j], n)
                    for j, e in enumerate(o["env"])]
        torch.cuda.synchronize()
        k = int(cnt.item())
        assert k == len(ref)
        got = res[:k].cpu().numpy()
        np.testing.assert_allclose(st.cpu().numpy(), np.stack([ost[q] for q in names]), rtol=1e-5, atol=1e-6)
        return got, np.array([r[:8] for r in ref], np.float32)

    # the reference's three games as three envs (frozen once over: compare each env's first finished episode)
    G = int(g["n_games"]); lens = [int(g[f"g{i}_steps"]) for i in range(G)]

    def gen_ref(n):
        raw = np.stack([g[f"g{i}_raw"][min(n, lens[i] - 1)] for i in range(G)]).astype(np.float32)
        logit = np.array([g[f"g{i}_logit"][min(n, lens[i] - 1)] for i in range(G)], np.float32)
        reset = np.array([int(n == L - 1) for L in lens], np.int64)
        return raw.sum(1), raw, reset, logit, np.array([g[f"g{i}_score"] for i in range(G)], np.float32), np.zeros(G, np.uint8)
    got, _ = run(G, max(lens), gen_ref, True, int(g["step_to_pred"]))
    first = {}
    for row in got:
        first.setdefault(int(row[0]), row)
    for i in range(G):
        np.testing.assert_allclose(first[i][2], g[f"g{i}_cr_to_pred"], rtol=1e-5)
        np.testing.assert_allclose(first[i][3], g[f"g{i}_norm_reward"], rtol=1e-5)
        np.testing.assert_allclose((first[i][1] - first[i][3]) ** 2, g[f"g{i}_value_loss"], rtol=1e-4)
        np.testing.assert_allclose(first[i][4:7], [g["rewards_loc"][i], g["rewards_pow"][i], g["rewards_disc"][i]], rtol=1e-4)
        assert first[i][7] == lens[i]
    # random play against the oracle, both reward modes
    for plot in (True, False):
        rng = np.random.default_rng(3)
        N = 500

        def gen(n):
            raw = rng.uniform([0, -0.3], [1, 0], (N, 2)).astype(np.float32)
            return (raw.sum(1), raw, (rng.random(N) < 0.07).astype(np.int64), rng.normal(0, 3, N).astype(np.float32),
                    rng.uniform(0, 1, N).astype(np.float32), (rng.random(N) < 0.3).astype(np.uint8))
        got, ref = run(N, 40, gen, plot, 12)
        assert len(got) > 500
        key = lambda a: a[np.lexsort((a[:, 7], a[:, 2], a[:, 0]))]
        np.testing.assert_allclose(key(got), key(ref), rtol=1e-4, atol=1e-5)


def test_player_runs_games_and_reports_the_value_return_correlation():
    """emloco_b200.player.AMPPlayerContinuousValue over a tensor-core Rollout: deterministic actions, finished episodes
    collected on the device, MSE / correlation report as the reference prints it (amp_value_players.py:263-279)."""
    from emloco_b200.player import AMPPlayerContinuousValue
    from emloco_b200.rollout import Rollout
    R = Rollout(256, seed=2, tensor_cores=True, traj_flags=2 | 4)                # --adjust_root_vel --init_heading, random-walk paths
    P = AMPPlayerContinuousValue(R)
    out = P.run(300)
    assert out["games"] >= 300 and np.isfinite(out["value_loss"])
    assert ((out["vals"] > 0) & (out["vals"] < 1)).all()                       # sigmoid outputs of LocoVal
    assert (out["steps"] >= 1).all() and (out["steps"] <= 168).all()
    assert set(np.unique(out["env"])) <= set(range(256))
    # deterministic actions: the sampled action equals the policy mean
    np.testing.assert_array_equal(R.mb["actions"][0].cpu().numpy(), R.mb["mus"][0].cpu().numpy())
    R.close()
