"""Generates the committed golden fixtures by running the UNMODIFIED reference
(/root/reference, via oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container only:  python tests/golden/make_golden.py
Outputs: tests/golden/pc_*.npz (inputs + reference outputs, bit-exact targets) and
         tests/golden/bg_*.npz (input seeds + reference logits / label maps).
The reference ships no golden vectors of its own, so these are the pin.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_loader  # noqa: E402
from panoptic_forecasting_b200 import synthetic  # noqa: E402


def pc_case(m, name, b, t, h, w, dist, seed, ind, is_img=False):
    inp = synthetic.make_pc_inputs(b=b, t=t, h=h, w=w, dist=dist, seed=seed)
    if is_img:
        g = torch.Generator().manual_seed(seed)
        inp["seg"] = torch.randint(0, 256, (b, t, h, w, 3), generator=g, dtype=torch.uint8)
    model = m.build_model(ref_loader.ref_pc_params(ind, is_img or None))
    with torch.no_grad():
        out = model.predict({k: v.clone() for k, v in inp.items()}, {})
    arrs = {"in_" + k: v.numpy() for k, v in inp.items()}
    arrs["in_intrinsics_inv"] = torch.inverse(inp["intrinsics"]).numpy()
    arrs["in_extrinsics_inv"] = torch.inverse(inp["extrinsics"]).numpy()
    arrs["out_seg"] = out["seg"].numpy()
    arrs["out_depth"] = out["depth"].numpy()
    arrs["out_result2d"] = out["result2d"].numpy().astype(np.int32)
    arrs["only_this_ind"] = np.array(-1 if ind is None else ind)
    arrs["is_img"] = np.array(int(is_img))
    np.savez_compressed(os.path.join(HERE, "pc_%s.npz" % name), **arrs)
    print("pc", name, "hit cells", int((out["depth"] >= 0).sum()), "of", out["depth"].numel())


def bg_case(m, name, h, w, fh, fw, seed, mode):
    bg = m.build_model(ref_loader.ref_bg_params(fh, fw)).eval()
    sd = synthetic.make_bg_state_dict(bg.state_dict(), seed=seed)
    if mode == "pc":
        pc = synthetic.make_pc_inputs(1, 3, h, w, "R", seed=seed)
        inp = {"seg": pc["seg"].long(), "depth": pc["depth"].clamp(0.1, 200), "depth_mask": pc["depth_mask"]}
    else:
        inp = synthetic.make_bg_inputs(1, 3, h, w, seed=seed)
    bg.load_state_dict(sd)
    with torch.no_grad():
        q = bg.predict({k: v.clone() for k, v in inp.items()}, {})["orig_size_logits"]
    # centre the class logits so every class wins somewhere (argmax is then a meaningful check)
    bias_shift = q.mean((0, 2, 3))
    sd["model.finalConv.bias"] = sd["model.finalConv.bias"] - bias_shift
    bg.load_state_dict(sd)
    with torch.no_grad():
        out = bg.predict({k: v.clone() for k, v in inp.items()}, {})
    np.savez_compressed(os.path.join(HERE, "bg_%s.npz" % name),
                        h=h, w=w, fh=fh, fw=fw, seed=seed, mode=np.array(mode),
                        bias_shift=bias_shift.numpy(),
                        out_seg=out["seg"].numpy().astype(np.uint8),
                        out_quarter=out["orig_size_logits"].numpy(),
                        out_logits_sample=out["logits"].numpy()[:, :, ::7, ::5])
    print("bg", name, "classes", np.bincount(out["seg"].numpy().ravel(), minlength=11))


def main():
    m = ref_loader.load_reference()
    pc_case(m, "R_ind0", 2, 3, 48, 96, "R", 0, 0)
    pc_case(m, "R_all", 2, 3, 48, 96, "R", 1, None)
    pc_case(m, "U_ind2", 1, 3, 40, 72, "U", 2, 2)
    pc_case(m, "U_all", 2, 2, 40, 72, "U", 3, None)
    pc_case(m, "R_img", 1, 3, 32, 64, "R", 4, 1, is_img=True)
    bg_case(m, "pc64", 64, 128, 64, 128, 0, "pc")
    bg_case(m, "iid64_up", 64, 128, 128, 256, 1, "iid")


if __name__ == "__main__":
    main()
