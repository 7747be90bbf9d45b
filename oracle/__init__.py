"""oracle/ — TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain PyTorch fp32 functional code, no nn.Module state, no CUDA) of the reference's
arithmetic for the shape-branch denoising hot path, each function citing the reference file:line it
follows.  It exists so that parity can be checked on the GPU box, where /root/reference is absent.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference` legs may import
this package — as the checker or the CPU baseline, never as the product path.  The product
(commonscenes_b200/) never imports it and fails loudly if its CUDA library is missing.

Pinning: the reference has no tests, golden vectors or fixtures for this path, so
the oracle is pinned against the reference's own modules imported from /root/reference in the build
container: `python oracle/validate_against_reference.py` (run at build time, results recorded in
DESIGN.md) and the committed fixtures under tests/golden/ produced by tests/golden/make_golden.py
from those same reference modules.
"""
