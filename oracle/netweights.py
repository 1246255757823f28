"""TEST INFRASTRUCTURE ONLY - seeded synthetic parameters of the rollout networks, as numpy arrays keyed like the reference's
`a2c_network.*` state dict (learning/network_builder.py:190-262, amp_network_builder.py:96-116, amp_network_sept_builder.py:113-136,
amp_network_sept_value_builder.py:48-90).  11.2 M parameters are too large for a committed fixture, so the fixtures under
tests/golden/ hold inputs, reference outputs and a checksum of these arrays; both sides regenerate the arrays from the seed
(numpy's PCG64 stream).  Scales follow nn.Linear's default U(-1/sqrt(k), 1/sqrt(k)); biases are NOT zero (the reference zeroes
them at construction - the tests must not be blind to them)."""
from __future__ import annotations

import numpy as np

SHAPES = (("actor_mlp.0", 2048, 624), ("actor_mlp.2", 1024, 2048), ("critic_mlp.0", 2048, 624), ("critic_mlp.2", 1024, 2048),
          ("value", 1, 1024), ("mu", 69, 1024), ("_disc_mlp.0", 1024, 3090), ("_disc_mlp.2", 512, 1024), ("_disc_logits", 1, 512),
          ("_task_mlp.0", 512, 1054), ("_task_mlp.2", 256, 512), ("_task_value_mlp.0", 15, 30), ("_task_value_mlp.2", 6, 15),
          ("_value_logits", 1, 6))


def synth_state_dict(seed=0, sigma=-2.9, mu_gain=1.0):
    rng = np.random.default_rng(seed)
    sd = {}
    for name, n_out, n_in in SHAPES:
        b = 1.0 / np.sqrt(n_in)
        if name in ("_disc_logits", "_value_logits"):
            b = 1.0                                                  # uniform(-1, 1) heads (DISC_LOGIT_INIT_SCALE)
        sd[f"{name}.weight"] = rng.uniform(-b, b, (n_out, n_in)).astype(np.float32)
        sd[f"{name}.bias"] = rng.uniform(-0.1, 0.1, n_out).astype(np.float32)
    sd["mu.weight"] *= np.float32(mu_gain)            # > 1: pushes some action means beyond the +-1 soft bound (bound_loss)
    sd["sigma"] = np.full(69, sigma, np.float32)
    return sd


def checksum(sd):
    return np.array([float(np.asarray(sd[k], np.float64).sum()) for k in sorted(sd)], np.float64)
