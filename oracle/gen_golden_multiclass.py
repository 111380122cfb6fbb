"""Golden vectors for the class-aware branches of ``ctdet_decode`` (centerface_ext.py:11-27 with C > 1, :72-77 cat_spec_wh),
generated FROM THE REFERENCE.  The face model has one class, so the bundled images never reach these branches; the inputs are
seeded synthetic head maps (no equal positive scores: torch.topk's order among ties is unspecified, the tie rule is pinned
separately by tests/test_oracle_golden.py::test_topk_tie_rule).

Runs only in the build container (needs /root/reference).  Asserts oracle == reference bit for bit, then writes
tests/golden/multiclass_v1.npz (inputs + the reference's outputs).

    python oracle/gen_golden_multiclass.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

CASES = [  # name, B, C, h, w, K, cat_spec_wh, with reg
    ("c3", 2, 3, 24, 32, 40, False, True),
    ("c3_cat", 2, 3, 24, 32, 40, True, True),
    ("c2_cat_noreg", 3, 2, 40, 40, 100, True, False),
    ("c5_k7", 1, 5, 16, 20, 7, False, False),
    ("c1_cat", 2, 1, 24, 32, 25, True, True),
]


def make_inputs(B, C, h, w, cat, seed):
    g = torch.Generator().manual_seed(seed)
    n = B * C * h * w  # distinct positive scores: a random permutation of an evenly spaced grid in (0.01, 0.99)
    heat = ((torch.randperm(n, generator=g).float() + 1) / (n + 1) * 0.98 + 0.01).view(B, C, h, w)
    wh = torch.rand(B, 2 * C if cat else 2, h, w, generator=g) * 12
    reg = torch.rand(B, 2, h, w, generator=g)
    return heat, wh, reg


def main():
    import centerface_oracle as O
    cwd = os.getcwd()
    os.chdir(REF)
    import model.centernet as mc
    mc.ghost_net = mc.efficientnet_b0
    import centerface_ext as ref_ext
    os.chdir(cwd)
    out = {}
    for i, (name, B, C, h, w, K, cat, with_reg) in enumerate(CASES):
        heat, wh, reg = make_inputs(B, C, h, w, cat, 9000 + i)
        assert heat.flatten().unique().numel() == heat.numel(), "tied scores in a golden input"
        want = ref_ext.ctdet_decode(heat.clone(), wh.clone(), reg.clone() if with_reg else None, cat_spec_wh=cat, K=K)
        got, inds = O.ctdet_decode(heat.clone(), wh.clone(), reg.clone() if with_reg else None, K=K, cat_spec_wh=cat)
        assert torch.equal(got, want), f"{name}: oracle != reference"
        if C > 1:
            assert len(set(want[..., 5].flatten().tolist())) > 1, f"{name}: only one class among the winners"
        out[f"{name}/heat"], out[f"{name}/wh"], out[f"{name}/reg"] = heat.numpy(), wh.numpy(), reg.numpy()
        out[f"{name}/dets"], out[f"{name}/inds"] = want.numpy(), inds.numpy().astype(np.int32)
        out[f"{name}/meta"] = np.array([B, C, h, w, K, int(cat), int(with_reg)], dtype=np.int32)
        print(f"{name}: oracle == reference, classes {sorted(set(want[..., 5].flatten().tolist()))}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "multiclass_v1.npz"), **out)
    print("wrote tests/golden/multiclass_v1.npz")


if __name__ == "__main__":
    main()
