#!/usr/bin/env python
"""Run the shared-memory eigensolver on a few small problems; meant to be run under
`compute-sanitizer --tool racecheck --kernel-name kns=jacobi_chol` (shared-memory hazard check)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collaborative_distillation_b200 import ops
for C, rank in ((24, 24), (32, 13), (64, 64), (100, 37), (128, 128)):
    g = torch.Generator().manual_seed(C)
    B = torch.randn(C, rank, generator=g, dtype=torch.float64)
    A = (B @ B.t()).cuda()[None]
    ev, evec, sw = ops.eigh_jacobi(A, [1.0], return_sweeps=True)
    ref = torch.linalg.eigvalsh(A[0].cpu()).flip(0)
    got = ev[0].cpu().sort(descending=True).values
    print("C=%d rank=%d sweeps=%d max|ev-ref|/lmax=%.2e" % (C, rank, int(sw[0]), ((got - ref).abs().max() / ref[0]).item()))
