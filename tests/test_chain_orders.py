"""CPU: the ticket orders the host hands to emloco_linear_chain (RolloutNets.policy_order / post_order / merged_order) cover every
tile exactly once and never place a tile before the tiles of the layer it depends on for its row block - the two conditions the
library checks before a launch (and that make the persistent kernel deadlock free).  Pure host logic: no GPU, no library call."""
import pytest

from emloco_b200 import _lib
from emloco_b200.policy import RolloutNets, tiles_of


def _layer(M, N, dep):
    L = _lib.ChainLayer()
    L.M, L.N, L.K, L.dep = M, N, 64, dep
    return L


def _check(layers, order):
    tm = [(l.M + 127) // 128 for l in layers]
    tn = [(l.N + 127) // 128 for l in layers]
    seen = [set() for _ in layers]
    claimed = [[0] * m for m in tm]
    for layer, first, count in order:
        assert 0 <= layer < len(layers) and count > 0 and first >= 0 and first + count <= tm[layer] * tn[layer]
        for t in range(first, first + count):
            assert t not in seen[layer], "tile listed twice"
            seen[layer].add(t)
            mb = t // tn[layer]
            d = layers[layer].dep
            if d >= 0:
                assert claimed[d][mb] == tn[d], f"layer {layer} tile {t} ordered before layer {d} finished row block {mb}"
            claimed[layer][mb] += 1
    assert all(len(s) == tm[i] * tn[i] for i, s in enumerate(seen)), "the order does not cover every tile"
    assert len(order) <= 32


def _policy(M, base=0):      # t0 t2 ac1 a2 c2 mu
    return [_layer(M, 512, -1), _layer(M, 256, base), _layer(M, 4096, base + 1), _layer(M, 1024, base + 2), _layer(M, 1024, base + 2),
            _layer(M, 69, base + 3)]


def _next_obs(M, base=0):    # t0 t2 c0 c2 d0 d2
    return [_layer(M, 512, -1), _layer(M, 256, base), _layer(M, 2048, base + 1), _layer(M, 1024, base + 2), _layer(M, 1024, -1),
            _layer(M, 512, base + 4)]


@pytest.mark.parametrize("M", [1, 64, 128, 200, 1000, 2048, 4096, 4097, 32768])
def test_ticket_orders_are_complete_and_dependency_respecting(M):
    tm = (M + 127) // 128
    L = _policy(M)
    _check(L, RolloutNets.policy_order(tm, L))
    L = _next_obs(M)
    _check(L, RolloutNets.post_order(tm, L))
    L = _policy(M) + _next_obs(M, base=6)
    _check(L, RolloutNets.merged_order(tm, L))
    assert sum(tiles_of(l) for l in L) == sum(c for _, _, c in RolloutNets.merged_order(tm, L))


def test_checker_rejects_what_the_library_rejects():
    L = _policy(256)
    good = RolloutNets.policy_order(2, L)
    _check(L, good)
    with pytest.raises(AssertionError):
        _check(L, good[1:] + good[:1])              # the task MLP's first layer after its consumers
    with pytest.raises(AssertionError):
        _check(L, good[:-1])                        # tiles missing
    with pytest.raises(AssertionError):
        _check(L, good + [good[0]])                 # tiles twice
