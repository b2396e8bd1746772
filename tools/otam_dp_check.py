"""Metric tail (cos_sim + OTAM) of the library against the torch-CPU oracle on ordinary and on range-extreme inputs --
every distance ~0 (a video against itself), every distance ~2 (against its negation), mixed -- for the wavefront
formulation the environment selects (SPM_OTAM_DP=log: log domain; default: exponent domain up to T = 16) and the kernel
it selects (SPM_OTAM=stream, SPM_OTAM_TC=0 / SPM_OTAM_TC_MINP, SPM_OTAM_FUSED=0, SPM_OTAM_KC=16).  Prints one line per case and the worst relative error; exit 1 above 1e-5.
Run by tests/test_stages_gpu.py::test_otam_wavefront_formulations_agree in a subprocess per setting (the switches are
read once per process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clip_spm_b200 import ops
from oracle import clipspm_oracle as O

worst = 0.0
for P, W, Q, T, D in ((3, 5, 5, 8, 512), (2, 5, 5, 16, 512), (2, 5, 3, 10, 1024), (300, 5, 5, 8, 512), (300, 5, 5, 16, 512),
                      (2, 4, 2, 20, 512), (1974, 5, 5, 8, 512), (1790, 5, 4, 8, 1024), (1800, 3, 2, 8, 512)):
    g = torch.Generator().manual_seed(P * 31 + T)
    sup = torch.randn(P, W, T, D, generator=g)
    cases = {"random": torch.randn(P, Q, T, D, generator=g),
             "self": sup[:, :1].expand(P, Q, T, D).contiguous(),
             "negated": -sup[:, :1].expand(P, Q, T, D).contiguous(),
             "near": sup[:, :1].expand(P, Q, T, D) + 0.05 * torch.randn(P, Q, T, D, generator=g)}
    for name, tgt in cases.items():
        for single in (False, True):
            if P > 300 and name in ("negated", "near"):
                continue
            # the oracle on a spread of problems: the first and the last ones (a CTA's short last unit) and two in between
            idx = sorted(set(range(min(P, 4))) | set(range(max(P - 4, 0), P)) | {P // 3, P // 2})
            ref = torch.stack([O.otam_distance(sup[p], tgt[p], single) for p in idx])
            out = ops.otam_distance(sup.cuda(), tgt.cuda(), single).cpu()
            assert torch.isfinite(out).all(), (name, P, T)
            err = float((out[idx] - ref).abs().max() / ref.abs().max().clamp_min(1e-3))
            worst = max(worst, err)
            print("P=%d W=%d Q=%d T=%d D=%d %-8s single=%d: rel err %.2e (|ref| max %.3f)" % (P, W, Q, T, D, name, single, err,
                                                                                             float(ref.abs().max())))
print("worst %.3e" % worst)
sys.exit(0 if worst < 1e-5 else 1)
