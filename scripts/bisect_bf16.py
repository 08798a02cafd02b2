"""Compare the bf16 (tcgen05) and fp32 (FMA) engines buffer by buffer after one train step (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import h_denseunet_b200 as hdn
from util import Args, perturb_params, rel_l2, synthetic_slab

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ms = {}
vol, lab = synthetic_slab(1, size, 8)
for prec in ("fp32", "bf16"):
    m = hdn.dense_rnn_net(Args(1, size, 8), precision=prec)
    m.dropout = False
    perturb_params(m)
    m.compile(optimizer=hdn.SGD(lr=0.0, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy])
    print(prec, "loss", m.train_on_batch(vol, lab))
    ms[prec] = m
n0, n1 = ms["fp32"].nets[True], ms["bf16"].nets[True]
print("---- forward buffers (rel-L2 bf16 vs fp32), only > 3e-2")
for b0, b1 in zip(n0.buffers, n1.buffers):
    e = rel_l2(b1.data.cpu().numpy(), b0.data.cpu().numpy())
    if e > 3e-2: print("  %-20s %.3e" % (b0.name, e))
print("---- gradient buffers in backward order: rel err, max|g32|, max|g16|  (only rel > 0.5)")
cnt = 0
for b0, b1 in reversed(list(zip(n0.buffers, n1.buffers))):
    if b0.grad is None: continue
    g0, g1 = b0.grad.cpu().numpy(), b1.grad.cpu().numpy()
    e = rel_l2(g1, g0)
    if e > 0.5:
        print("  %-20s %.3e  |g32| %.3e |g16| %.3e" % (b0.name, e, np.abs(g0).max(), np.abs(g1).max())); cnt += 1
    if cnt > 30: break
g0, g1 = ms["fp32"].get_grads_dict(), ms["bf16"].get_grads_dict()
print("---- parameter grads: name, rel, max|g32|, max|g16| (rel > 1)")
cnt = 0
for k in g0:
    if np.abs(g0[k]).max() < 1e-12 and np.abs(g1[k]).max() < 1e-12: continue
    e = rel_l2(g1[k], g0[k])
    if e > 1.0:
        print("  %-28s %.3e  %.3e  %.3e" % (k, e, np.abs(g0[k]).max(), np.abs(g1[k]).max())); cnt += 1
    if cnt > 40: break
# which ops are on which path
print([ (n, p) for n, p in n1.report if not all(p)][:10])
