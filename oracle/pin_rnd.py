"""TEST INFRASTRUCTURE ONLY.  Host-side draw of one step's randomness (shared by pin.py and tests)."""
import numpy as np
import torch


def make_rnd(rs, B, A, pool_n, with_places=False):
    d = dict(
        noise_next=torch.as_tensor(rs.randn(B, A).astype(np.float32)),
        noise_pi=torch.as_tensor(rs.randn(B, A).astype(np.float32)),
        u=float(rs.rand()),
        overlay_ids=rs.randint(0, pool_n, size=B),
        overlay_ids_unused=rs.randint(0, pool_n, size=B),
    )
    if with_places:
        d["places"] = torch.as_tensor(rs.rand(B, 3, 84, 84).astype(np.float32))
    return d
