"""GPU: the layer chain (emloco_linear_chain - every dense layer of a network pass in one persistent tcgen05 launch) against the
per-layer launches it replaces (emloco_linear_bf16x3 / _head), which are themselves pinned to fp64 and to the reference golden
`nets.npz` in test_gpu_parity.py.  Same arithmetic in the same order => the comparison is BIT-EXACT."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, dims, seed=0, heads=()):
    """A ReLU MLP dims[0] -> dims[1] -> ... as split operands; heads: indices of layers followed by a single-output layer."""
    from emloco_b200.policy import _Split, split_bf16
    rng = np.random.default_rng(seed)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    x = T(rng.standard_normal((M, dims[0])))
    sx = _Split(M, dims[0], x.device)
    split_bf16(x, sx)
    Ws, bs, hs = [], [], {}
    for i in range(len(dims) - 1):
        w = T(rng.standard_normal((dims[i + 1], dims[i])) / np.sqrt(dims[i]))
        sw = _Split(dims[i + 1], dims[i], x.device)
        split_bf16(w, sw)
        Ws.append(sw); bs.append(T(rng.standard_normal(dims[i + 1]) * 0.1))
        if i in heads:
            hd = torch.nn.Linear(dims[i + 1], 1).cuda()
            hs[i] = hd
    return sx, Ws, bs, hs


def _run_layers(M, dims, sx, Ws, bs, hs, use_chain, order=None, ws=None):
    from emloco_b200.policy import _Split, chain_layer, linear_bf16x3, linear_chain
    dev = sx.hi.device
    acts = [sx] + [_Split(M, d, dev) for d in dims[1:]]
    y32 = [torch.zeros(M, d, device=dev) for d in dims[1:]]
    outs = {i: (torch.zeros(M, 1, device=dev), torch.zeros(M, (dims[i + 1] + 63) // 64, device=dev)) for i in hs}
    if use_chain:
        L = []
        for i in range(len(dims) - 1):
            split_ok = dims[i + 1] % 32 == 0
            L.append(chain_layer(acts[i], Ws[i], bs[i], True, dep=i - 1, y32=y32[i], y16=acts[i + 1] if split_ok else None,
                                 head=(hs[i], outs[i][0], outs[i][1]) if i in hs else None))
        linear_chain(L, order, ws)
    else:
        for i in range(len(dims) - 1):
            split_ok = dims[i + 1] % 32 == 0
            linear_bf16x3(acts[i], Ws[i], bs[i], True, y32=y32[i], y16=acts[i + 1] if split_ok else None, tile=128,
                          head=(hs[i], outs[i][0], outs[i][1]) if i in hs else None)
    torch.cuda.synchronize()
    return y32, {i: o[0] for i, o in outs.items()}, acts


@pytest.mark.parametrize("M,dims,heads", [(300, (200, 192), ()), (64, (70, 96, 64, 69), (1,)), (1000, (1054, 512, 256, 320), (0, 2)),
                                          (4096, (624, 2048, 1024, 69), (1,))])
def test_chain_is_bit_identical_to_the_per_layer_launches(M, dims, heads):
    args = _mk(M, dims, seed=M, heads=heads)
    y_ref, h_ref, a_ref = _run_layers(M, dims, *args, use_chain=False)
    y, h, a = _run_layers(M, dims, *args, use_chain=True)
    for i in range(len(y)):
        assert torch.equal(y[i], y_ref[i]), f"layer {i}"
        assert torch.equal(a[i + 1].hi, a_ref[i + 1].hi) and torch.equal(a[i + 1].lo, a_ref[i + 1].lo)
    for i in h:
        assert torch.equal(h[i], h_ref[i]), f"head after layer {i}"
    # and against fp64 (the per-layer kernels' own bar: ~1e-5 relative)
    from emloco_b200.policy import _Split
    x = (args[0].hi.double() + args[0].lo.double())[:, :dims[0]]
    for i in range(len(dims) - 1):
        w = (args[1][i].hi.double() + args[1][i].lo.double())[:, :dims[i]]
        x = torch.relu(x @ w.T + args[2][i].double())
        np.testing.assert_allclose(y[i].cpu().numpy(), x.cpu().numpy(), rtol=1e-3, atol=2e-4)
        x = (a[i + 1].hi.double() + a[i + 1].lo.double())[:, :dims[i + 1]] if dims[i + 1] % 32 == 0 else x


def test_chain_reuses_its_workspace_and_honours_an_interleaved_order():
    """Repeated launches on one workspace (the kernel leaves it zeroed) with tiles of independent layers interleaved."""
    from emloco_b200.policy import chain_layer, chain_workspace_ints, linear_chain, tiles_of, _Split
    M = 700
    a = _mk(M, (256, 384, 128), seed=1)
    b = _mk(M, (512, 256), seed=2)
    dev = a[0].hi.device
    ya = [torch.zeros(M, 384, device=dev), torch.zeros(M, 128, device=dev)]
    yb = torch.zeros(M, 256, device=dev)
    s1 = _Split(M, 384, dev)

    def layers():
        return [chain_layer(a[0], a[1][0], a[2][0], True, y32=ya[0], y16=s1), chain_layer(b[0], b[1][0], b[2][0], False, y32=yb),
                chain_layer(s1, a[1][1], a[2][1], True, dep=0, y32=ya[1])]
    L = layers()
    t = [tiles_of(l) for l in L]
    order = [(0, 0, t[0]), (1, 0, 5), (2, 0, 3), (1, 5, t[1] - 5), (2, 3, t[2] - 3)]
    ws = torch.zeros(chain_workspace_ints(L), dtype=torch.int32, device=dev)
    linear_chain(L, None, ws)
    torch.cuda.synchronize()
    want = [ya[0].clone(), ya[1].clone(), yb.clone()]
    assert int(ws.abs().sum()) == 0
    for it in range(25):
        for y in (*ya, yb):
            y.zero_()
        linear_chain(layers(), order if it % 2 else None, ws)
        torch.cuda.synchronize()
        assert torch.equal(ya[0], want[0]) and torch.equal(ya[1], want[1]) and torch.equal(yb, want[2]), it
        assert int(ws.abs().sum()) == 0


def test_chain_rejects_orders_that_could_deadlock():
    from emloco_b200 import _lib
    from emloco_b200.policy import chain_layer, linear_chain, tiles_of, _Split
    M = 256
    a = _mk(M, (128, 256, 128), seed=3)
    dev = a[0].hi.device
    s1 = _Split(M, 256, dev)
    y = torch.zeros(M, 128, device=dev)
    L = [chain_layer(a[0], a[1][0], a[2][0], True, y16=s1), chain_layer(s1, a[1][1], a[2][1], True, dep=0, y32=y)]
    t = [tiles_of(l) for l in L]
    with pytest.raises(_lib.EmlocoError, match="before the tiles it depends on"):
        linear_chain(L, [(0, 0, 1), (1, 0, t[1]), (0, 1, t[0] - 1)])
    with pytest.raises(_lib.EmlocoError, match="cover every tile"):
        linear_chain(L, [(0, 0, t[0]), (1, 0, t[1] - 1)])
    with pytest.raises(_lib.EmlocoError, match="twice"):
        linear_chain(L, [(0, 0, t[0]), (0, 0, 1), (1, 0, t[1])])
    with pytest.raises(_lib.EmlocoError, match="earlier layer"):
        linear_chain([chain_layer(s1, a[1][1], a[2][1], True, dep=1, y32=y)])
    linear_chain(L, [(0, 0, t[0]), (1, 0, t[1])])      # a row block's tiles may start as soon as ITS rows are done
    torch.cuda.synchronize()


@pytest.mark.parametrize("M", [64, 4096])
def test_network_passes_through_the_chain_equal_the_per_layer_passes(M):
    """RolloutNets(chain=True): get_action_values and critic(next obs) + discriminator - the two launches of a rollout step -
    give bit-identical mu / value / next value / logit / sampled actions to the per-layer path."""
    from emloco_b200.policy import AMP_OBS, OBS, AMPSeptValueNetwork, RolloutNets, RunningMeanStd
    torch.manual_seed(3)
    net = AMPSeptValueNetwork().cuda()
    on, an = RunningMeanStd(OBS).cuda(), RunningMeanStd(AMP_OBS).cuda()
    on.running_mean.normal_(); on.running_var.uniform_(0.5, 2.0)
    an.running_mean.normal_(); an.running_var.uniform_(0.5, 2.0)
    obs, amp, noise = torch.randn(M, OBS, device="cuda") * 1.5, torch.randn(M, AMP_OBS, device="cuda") * 1.5, torch.randn(M, 69, device="cuda")
    res = {}
    for chain in (False, True):
        nets = RolloutNets(net, on, an, M, tensor_cores=True, concurrent=True, chain=chain)
        nets.sync_weights()
        r = nets.action_values(obs, noise)
        out = {k: r[k].clone() for k in ("mus", "values", "task_values", "actions", "neglogpacs")}
        if chain:
            nv, lg = nets.critic_disc(obs, amp)
        else:
            nv, lg = nets.critic(obs), nets.disc_logits(amp)
        torch.cuda.synchronize()
        out.update(next_values=nv.clone(), logits=lg.clone())
        res[chain] = out
    for k in res[True]:
        assert torch.isfinite(res[True][k]).all()
        assert torch.equal(res[True][k], res[False][k]), k
