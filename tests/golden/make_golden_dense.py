"""Golden fixtures for the two BGModel surfaces added in round 2, produced by the UNMODIFIED reference
(/root/reference via oracle/ref_loader.py) on seeded synthetic inputs:

  dense_soft64.npz  `convert2onehot` off (bg_model.py:61-69): float per-class planes in, predict() outputs
  loss_iid64.npz    BGModel.loss (bg_model.py:73-89) in .eval() mode on label inputs: loss + accuracy

Run in the build container only:  python tests/golden/make_golden_dense.py
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


def centred(bg, sd, inp):
    """final-conv bias shifted so that every class wins somewhere (as make_golden.py does)"""
    bg.load_state_dict(sd)
    with torch.no_grad():
        q = bg.predict({k: v.clone() for k, v in inp.items()}, {})["orig_size_logits"]
    shift = q.mean((0, 2, 3))
    sd["model.finalConv.bias"] = sd["model.finalConv.bias"] - shift
    bg.load_state_dict(sd)
    return shift


def main():
    m = ref_loader.load_reference()
    h, w, seed = 64, 128, 7
    # ---- dense planes
    p = ref_loader.ref_bg_params(None, None)
    p["model"]["convert2onehot"] = False
    bg = m.build_model(p).eval()
    sd = synthetic.make_bg_state_dict(bg.state_dict(), seed=seed)
    inp = synthetic.make_bg_dense_inputs(2, 3, h, w, seed=seed)
    shift = centred(bg, sd, inp)
    with torch.no_grad():
        out = bg.predict({k: v.clone() for k, v in inp.items()}, {})
        target = synthetic.make_loss_target(out["seg"], seed=seed)
        ls = bg.loss({k: v.clone() for k, v in inp.items()}, {"seg": target})
    np.savez_compressed(os.path.join(HERE, "dense_soft64.npz"), h=h, w=w, seed=seed, bias_shift=shift.numpy(),
                        out_seg=out["seg"].numpy().astype(np.uint8), out_quarter=out["orig_size_logits"].numpy(),
                        out_logits_sample=out["logits"].numpy()[:, :, ::7, ::5],
                        loss=np.float64(ls["loss"].item()), accuracy=np.float64(ls["accuracy"].item()))
    print("dense classes", np.bincount(out["seg"].numpy().ravel(), minlength=11), "loss", ls["loss"].item(),
          "accuracy", ls["accuracy"].item())
    # ---- loss on label inputs
    bg = m.build_model(ref_loader.ref_bg_params(None, None)).eval()
    sd = synthetic.make_bg_state_dict(bg.state_dict(), seed=seed + 1)
    inp = synthetic.make_bg_inputs(2, 3, h, w, seed=seed + 1)
    shift = centred(bg, sd, inp)
    with torch.no_grad():
        out = bg.predict({k: v.clone() for k, v in inp.items()}, {})
        target = synthetic.make_loss_target(out["seg"], seed=seed + 1)
        ls = bg.loss({k: v.clone() for k, v in inp.items()}, {"seg": target})
    np.savez_compressed(os.path.join(HERE, "loss_iid64.npz"), h=h, w=w, seed=seed + 1, bias_shift=shift.numpy(),
                        out_seg=out["seg"].numpy().astype(np.uint8),
                        loss=np.float64(ls["loss"].item()), accuracy=np.float64(ls["accuracy"].item()))
    print("label-input loss", ls["loss"].item(), "accuracy", ls["accuracy"].item())


if __name__ == "__main__":
    main()
