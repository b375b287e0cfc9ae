"""CPU oracle for the keypoint-SLDS Gibbs sweep.

TEST INFRASTRUCTURE ONLY.  Nothing under ``keypoint_moseq_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs do.

PARITY UNPINNED: the arithmetic of the reference's hot path lives in the
third-party package ``jax-moseq`` (unpinned in /root/reference/setup.cfg:39),
which is absent from /root/reference and not installable here, and the
reference ships no tests or golden vectors.  This oracle restates the published
keypoint-SLDS conditionals (Weinreb et al. 2024) following the reference's call
sites (keypoint_moseq/fitting.py:25, :248, :536-538, :667-673) and the upstream
module structure; every convention that could not be verified is listed in
DESIGN.md ("open points").
"""
from .kpms_oracle import *  # noqa: F401,F403
