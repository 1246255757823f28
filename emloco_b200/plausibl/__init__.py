"""Call surface of the reference's `plausibl/test_value_mlp.py` (the trajectory-only value MLP)."""
from .test_value_mlp import MLP  # noqa: F401
