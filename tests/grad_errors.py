"""Parity numbers of one hybrid (end2end) train step at 64x64x8 for a given precision, against the fp32 / fp64 oracle:
logits rel-L2, loss, and the per-tensor gradient rel-L2 distribution (GPU).  A measurement tool of the test suite (it
uses the oracle as the checker): python tests/grad_errors.py mixed bf16x3 bf16 -> profiles/r01c_grad_errors.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


import numpy as np  # noqa: E402
import torch  # noqa: E402

import h_denseunet_b200 as hdn  # noqa: E402
from oracle import hdense_oracle as orc  # noqa: E402
from util import Args, perturb_params, rel_l2, synthetic_slab  # noqa: E402


SIZE = int(os.environ.get("HDN_GE_SIZE", "64"))


def run(prec, oracle):
    a = Args(b=1, input_size=SIZE, input_cols=8)
    m = hdn.dense_rnn_net(a, precision=prec)
    m.dropout = False
    perturb_params(m)
    m.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy])
    vol, lab = synthetic_slab(1, SIZE, 8)
    if not oracle:
        w0 = m.get_weights_dict()
        for dt in (torch.float32, torch.float64):
            ctx, logits = orc.forward_hybrid(w0, vol, training=True, variant="end2end", requires_grad=True, dtype=dt)
            loss = orc.weighted_crossentropy(torch.as_tensor(lab[..., 0]), logits, crop=True)
            oracle[dt] = (orc.grads_of(ctx, loss), float(loss.detach()), logits.detach().numpy())
    got = m.train_on_batch(vol, lab)
    net = m.nets[True]
    g32, l32, lg32 = oracle[torch.float32]
    g64 = oracle[torch.float64][0]
    eg = m.get_grads_dict()
    errs = []
    for k in sorted(p.name for p in m.params.order if p.trainable):
        if g64.get(k) is None or np.abs(g64[k]).max() < 1e-9:
            continue
        errs.append((rel_l2(eg[k], g64[k]), rel_l2(g32[k], g64[k]), k))
    errs.sort(reverse=True)
    e = np.array([x[0] for x in errs])
    return {"precision": prec, "logits_rel_l2": rel_l2(m._logits_to_host(net), lg32), "loss": got, "loss_oracle": l32,
            "loss_rel": abs(got - l32) / abs(l32), "n_tensors": len(errs), "grad_max": float(e.max()),
            "grad_median": float(np.median(e)), "grad_over_5e-3": int((e > 5e-3).sum()), "grad_over_1e-2": int((e > 1e-2).sum()),
            "size": SIZE, "grad_over_1e-3": int((e > 1e-3).sum()),
            "by_family": {fam: [round(float(np.median([x[0] for x in errs if x[2].startswith(fam)])), 5),
                                round(float(max([x[0] for x in errs if x[2].startswith(fam)])), 5)]
                          for fam in ("conv1", "conv2_", "conv3_", "conv4_", "conv5_", "conv_up", "bn_up", "dense167", "3dconv1", "3dconv2_",
                                      "3dconv3_", "3dconv4_", "3dconv5_", "3dconv_up", "3dbn_up", "fianl", "final_bn", "2d3d")
                          if any(x[2].startswith(fam) for x in errs)},
            "worst": [(k, round(a_, 5), round(b_, 5)) for a_, b_, k in errs[:int(os.environ.get("HDN_GE_WORST", "8"))]],
            "passes": sorted(set(p for _, p in net.report))}


if __name__ == "__main__":
    torch.cuda.set_device(0)
    oracle = {}
    for prec in sys.argv[1:] or ["mixed"]:
        print(json.dumps(run(prec, oracle)), flush=True)
