"""Host-logic tests (no GPU): the engine's program construction / backward planning is driven
with the plain-PyTorch reference backend (tests/torch_backend.py) on CPU tensors and compared
with the oracle (oracle/hdense_oracle.py) -- logits, loss, every parameter gradient, BN moving
statistics and one Nesterov step."""
import numpy as np
import pytest
import torch

import h_denseunet_b200 as hdn
from oracle import hdense_oracle as orc
from torch_backend import TorchBackend
from util import Args, perturb_params, rel_l2, synthetic_slab


def _check_grads(eg, og32, og64, names=None, floor=5e-3):
    """Engine gradients against the fp64 oracle.  Deep training-mode-BN nets amplify fp32 rounding
    (the fp32 oracle itself sits ~1e-2 from the fp64 one on the tiny test shapes), so the bound per
    tensor is max(floor, 8 x the fp32 oracle's own distance from fp64)."""
    bad, n = [], 0
    for k in (names if names is not None else eg.keys()):
        if og64.get(k) is None:
            assert np.abs(eg[k]).max() == 0, "engine has a gradient for %s but the oracle has none" % k
            continue
        if np.abs(og64[k]).max() < 1e-9:
            continue
        n += 1
        e = rel_l2(eg[k], og64[k])
        tol = max(floor, 8.0 * rel_l2(og32[k], og64[k]))
        if e > tol:
            bad.append((k, e, tol))
    assert not bad, "gradient mismatch (name, err, tol): %s" % bad[:10]
    assert n > 50


def _oracle_grads_2d(w0, x, y, skip):
    out = {}
    for dt in (torch.float32, torch.float64):
        ctx, _, logits = orc.forward_2d(w0, x, training=True, learn_bn=True, skip=skip, requires_grad=True, dtype=dt)
        loss = orc.weighted_crossentropy(torch.as_tensor(y), logits, crop=False)
        out[dt] = (ctx, logits.detach(), float(loss), orc.grads_of(ctx, loss))
    return out


@pytest.mark.parametrize("skip", [False, True])
def test_unet2d_train_step_matches_oracle(skip):
    a = Args(b=2, input_size=128)
    m = hdn.DenseUNet(reduction=0.5, args=a, skip=skip, backend=TorchBackend(), device="cpu", precision="fp32")
    m.dropout = False
    perturb_params(m)
    m.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy_2ddense])
    rng = np.random.default_rng(0)
    x = rng.normal(0, 60, (2, 128, 128, 3)).astype(np.float32)
    y = rng.integers(0, 3, (2, 128, 128, 1)).astype(np.int16)
    w0 = m.get_weights_dict()
    o = _oracle_grads_2d(w0, x, y, skip)
    ctx, logits, oloss, og = o[torch.float32]
    loss = m.train_on_batch(x, y)
    assert abs(loss - oloss) <= 1e-4 * abs(oloss)
    net = m.nets[True]
    el = net.outputs["logits"].buf.data.view(2, 128, 128, 3).numpy()
    assert rel_l2(el, logits.numpy()) < 1e-4
    _check_grads(m.get_grads_dict(), og, o[torch.float64][3])
    # BN moving statistics (KNORM:179-185) and the Nesterov update (optimizers.py:172-181)
    w1 = m.get_weights_dict()
    for name, (mean, var, mom) in list(ctx.bn_updates.items())[:20]:
        exp = orc.moving_average_update(w0[name + "/moving_mean"], mean.numpy(), mom)
        assert np.allclose(w1[name + "/moving_mean"], exp, atol=1e-5, rtol=1e-4)
        exp = orc.moving_average_update(w0[name + "/moving_variance"], var.numpy(), mom)
        assert np.allclose(w1[name + "/moving_variance"], exp, atol=1e-5, rtol=1e-4)
    for k in ["conv1/kernel", "conv3_5_x2/kernel", "conv_up2/bias", "conv4_7_x1_scale/gamma", "bn_up3/beta"]:
        g = m.get_grads_dict()[k]
        p1, _ = orc.sgd_nesterov_step(w0[k], g, np.zeros_like(w0[k]))
        assert np.allclose(w1[k], p1, atol=1e-7, rtol=1e-5), k


@pytest.mark.parametrize("variant", ["end2end", "3dpart"])
def test_hybrid_train_step_matches_oracle(variant):
    a = Args(b=1, input_size=64, input_cols=8)     # 3-D block 5 is 2x2x2: smaller makes training-mode BN degenerate
    build = hdn.dense_rnn_net if variant == "end2end" else hdn.denseunet_3d
    m = build(a, backend=TorchBackend(), device="cpu", precision="fp32")
    m.dropout = False
    perturb_params(m)
    m.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy])
    vol, lab = synthetic_slab(1, 64, 8)
    w0 = m.get_weights_dict()
    og = {}
    for dt in (torch.float32, torch.float64):
        ctx, logits = orc.forward_hybrid(w0, vol, training=True, variant=variant, requires_grad=True, dtype=dt)
        oloss = orc.weighted_crossentropy(torch.as_tensor(lab[..., 0]), logits, crop=True)
        og[dt] = orc.grads_of(ctx, oloss)
        if dt == torch.float32:
            oloss32 = float(oloss)
    loss = m.train_on_batch(vol, lab)
    assert abs(loss - oloss32) <= 1e-4 * abs(oloss32)
    trainable = sorted(p.name for p in m.params.order if p.trainable)
    # 3dpart trains every 3-D BN on batch statistics of as few as 8 samples (block 5 is 2x2x2 here):
    # rounding noise reaches a few percent; a wrong formula shows up as O(1)
    _check_grads(m.get_grads_dict(), og[torch.float32], og[torch.float64], names=trainable,
                 floor=5e-2 if variant == "3dpart" else 5e-3)
    if variant == "3dpart":
        assert not any(k.startswith("conv") or k.startswith("bn_up") or k.startswith("dense167") for k in trainable)
    # inference program shares the parameters
    out = m.predict(vol)
    ctx2, l2 = orc.forward_hybrid(m.get_weights_dict(), vol, training=False, variant=variant)
    assert rel_l2(out, l2.numpy()) < 1e-4


def test_hybrid_train_step_with_dropout_matches_oracle():
    """Dropout on (KB:2888, rate 0.3 after fianl_conv, hybridnet.py:415): the forward mask and the backward mask come from the
    same stateless hash of (seed, element index); with that mask injected, the oracle's loss and gradients are the engine's."""
    a = Args(b=1, input_size=64, input_cols=8)
    be = TorchBackend()
    m = hdn.dense_rnn_net(a, backend=be, device="cpu", precision="fp32")
    assert m.dropout
    perturb_params(m)
    m.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy])
    vol, lab = synthetic_slab(1, 64, 8)
    w0 = m.get_weights_dict()
    loss = m.train_on_batch(vol, lab)
    assert len(be.masks) == 1                                  # one dropout site in the end2end graph
    mask = be.masks[0]                                         # (N, S, H, W, C)
    assert abs(float((mask == 0).float().mean()) - 0.3) < 0.02
    om = {"fianl_conv_dropout": mask.permute(0, 4, 2, 3, 1).numpy()}         # oracle layout (B, C, H, W, S)
    og = {}
    for dt in (torch.float32, torch.float64):
        ctx, logits = orc.forward_hybrid(w0, vol, training=True, requires_grad=True, dtype=dt, dropout_masks=om)
        oloss = orc.weighted_crossentropy(torch.as_tensor(lab[..., 0]), logits, crop=True)
        og[dt] = orc.grads_of(ctx, oloss)
        if dt == torch.float32:
            oloss32 = float(oloss)
    assert abs(loss - oloss32) <= 1e-4 * abs(oloss32)
    # and the masked run differs from the unmasked one (the mask is really applied)
    _, l_nomask = orc.forward_hybrid(w0, vol, training=True)
    assert abs(float(orc.weighted_crossentropy(torch.as_tensor(lab[..., 0]), l_nomask, crop=True)) - oloss32) > 1e-4 * abs(oloss32)
    trainable = sorted(p.name for p in m.params.order if p.trainable)
    _check_grads(m.get_grads_dict(), og[torch.float32], og[torch.float64], names=trainable, floor=5e-3)


def test_net3d_matches_oracle():
    a = Args(b=1, input_size=32, input_cols=8)
    m = hdn.DenseNet3D(a, backend=TorchBackend(), device="cpu", precision="fp32")
    perturb_params(m)
    rng = np.random.default_rng(3)
    x = rng.normal(0, 50, (1, 32, 32, 8, 4)).astype(np.float32)
    net = m._net(False)
    m._upload(net, x)
    net.forward()
    f = net.outputs["feature3d"]
    got = torch.relu(f.view.buf.data * f.fold.a + f.fold.b).permute(0, 2, 3, 1, 4).numpy()
    ctx, exp = orc.forward_3d(m.get_weights_dict(), x, training=False)
    assert rel_l2(got, exp.numpy()) < 1e-4
