"""Generates tests/golden/oracle_golden.json: loss, logit samples / checksums and Dice of the CPU oracle on seeded
synthetic inputs (2-D DenseUNet-161 forward = BASELINE config 1 at reduced size; hybrid end2end forward + loss).
The reference itself cannot run here (TensorFlow 1.x absent), so these are the oracle's own outputs, kept to detect
drift of the oracle -- run  python tests/golden/make_golden.py  to regenerate after an intended change."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def compute():
    import h_denseunet_b200 as hdn
    from oracle import hdense_oracle as orc
    from util import Args, perturb_params, synthetic_slab
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    out = {}
    m = hdn.DenseUNet(reduction=0.5, args=Args(b=1, input_size=64))
    perturb_params(m)
    x = np.random.default_rng(1235).normal(0, 60, (1, 64, 64, 3)).astype(np.float32)
    _, feat, logits = orc.forward_2d(m.get_weights_dict(), x, training=False, learn_bn=True)
    lg = logits.numpy()
    out["c1_logits_sample"] = lg[0, ::16, ::16, :].reshape(-1).tolist()
    out["c1_logits_abs_sum"] = float(np.abs(lg).sum())
    out["c1_feature_abs_sum"] = float(np.abs(feat.numpy()).sum())
    m = hdn.dense_rnn_net(Args(b=1, input_size=32, input_cols=8))
    perturb_params(m)
    vol, lab = synthetic_slab(1, 32, 8, seed=1238)
    ctx, logits = orc.forward_hybrid(m.get_weights_dict(), vol, training=False)
    lg = logits.numpy()
    out["c4_logits_sample"] = lg[0, ::8, ::8, ::2, :].reshape(-1).tolist()
    out["c4_logits_abs_sum"] = float(np.abs(lg).sum())
    out["c4_loss"] = float(orc.weighted_crossentropy(torch.as_tensor(lab[..., 0]), logits, crop=True))
    p = torch.softmax(logits, -1).numpy()
    out["c4_dice_liver"] = float(orc.dice(p[..., 1] > 0.5, lab[..., 0] == 1))
    return out


if __name__ == "__main__":
    g = compute()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", len(g), "entries")
