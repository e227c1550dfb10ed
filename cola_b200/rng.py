"""Keyed probe RNG, bit-compatible with the reference's torch backend
(cola/backends/torch_fns.py:154-155, 222-241): SHA-256 key chain + torch.randn under a temporary seed.

The reference draws on the operator's device, so CPU and CUDA runs of the reference itself see different
streams.  `PROBE_DEVICE = "cpu"` makes this module draw on the CPU generator and copy to the GPU, which
reproduces the CPU reference / oracle stream exactly (used by the parity tests); the default (None) draws on
the operator's device like the reference does."""
import hashlib
import logging

import torch

PROBE_DEVICE = None


def sha_hash(n):
    n_bytes = n.to_bytes((n.bit_length() + 7) // 8, "big")
    return int(int.from_bytes(hashlib.sha256(n_bytes).digest(), "big") % (2**32 - 1))


def PRNGKey(x):
    return sha_hash(x)


def next_key(key):
    return sha_hash(key)


def randn(*shape, dtype, device, key=None):
    if key is None:
        logging.warning("Non keyed randn used. To be deprecated soon.")
        key = PRNGKey(0)
    draw_on = PROBE_DEVICE if PROBE_DEVICE is not None else device
    old = torch.random.get_rng_state()
    torch.random.manual_seed(key)
    z = torch.randn(*shape, dtype=dtype, device=draw_on)
    torch.random.set_rng_state(old)
    return z.to(device)
