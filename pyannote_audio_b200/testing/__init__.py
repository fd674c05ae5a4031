"""Test / benchmark fixtures (NOT part of the product path): seeded synthetic weights with the reference's state-dict
keys, a synthetic PLDA and synthetic multi-speaker audio.  No pretrained checkpoint exists offline."""
from . import synthetic  # noqa: F401
