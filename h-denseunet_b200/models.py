"""Program builders for the H-DenseUNet networks (what the reference's builder files construct
as Keras graphs), emitted as fused ops on an engine.Net.

Reference topology followed line by line:
  2-D DenseUNet-161   hybridnet.py:182-354 (frozen-BN variant), densenet.py:10-193 (trainer),
                      denseunet.py:130-226 (skip-add decoder with `line0`)
  3-D DenseNet        hybridnet.py:11-178, denseunet3d.py:18-190
  hybrid assembly     hybridnet.py:379-423 (end2end), denseunet3d.py:393-439 (3dpart)
Layer / weight names are the reference's, so a weight dict keyed "<layer>/<weight>" round-trips.

BN / trainability matrix, expressed by `Mode`:
  '2d'      train_2ddense.py : everything trainable, every BN in training mode, Dropout(.3) after conv_up4
  'end2end' train_hybrid.py  : 2-D BN frozen-inference, 2-D kernels + Scale trainable; 3-D conv_block BN
                               frozen, 3-D transition BN inference with trainable gamma/beta, the other
                               3-D BNs in training mode; Dropout(.3) after fianl_conv
  '3dpart'                    : whole 2-D branch frozen; every 3-D BN in training mode; Dropout(.1)
"""
from .engine import Act, Src, glorot_uniform, random_normal

EPS_DENSE = 1.1e-5     # hybridnet.py:21


class Mode(object):
    def __init__(self, name):
        assert name in ("2d", "end2end", "3dpart")
        self.name = name
        # 2-D branch
        self.k2d = name != "3dpart"          # conv kernels / biases trainable
        self.bn2d_learn = name == "2d"       # BN uses batch statistics in training
        self.bn2d_train = name == "2d"       # BN gamma/beta trainable
        self.sc2d_train = name != "3dpart"   # Scale trainable
        self.dec_init = random_normal if name == "2d" else glorot_uniform   # densenet.py:71 vs hybridnet.py:236
        # 3-D branch
        self.dense3d_learn = name == "3dpart"      # conv_block / transition BN training mode
        self.dense3d_bn_train = name == "3dpart"   # conv_block BN params trainable (hybridnet.py:27 trainable=False)
        self.trans3d_bn_train = True               # hybridnet.py:87 has no trainable=False
        self.final_drop = {"2d": 0.0, "end2end": 0.3, "3dpart": 0.1}[name]


def _k(nd, k):
    return (1, k, k) if nd == 2 else (k, k, k)


def _p(nd, p):
    return (0, p, p) if nd == 2 else (p, p, p)


def dense_block(net, buf, c0, stage, nb_layers, growth, nd, pre, learn, bn_train, sc_train, k_train):
    """hybridnet.py:330-354 / :46-70 with conv_block (hybridnet.py:264-298 / :11-45) inlined.
    `buf` already holds the block input in channels [0, c0); layer i appends `growth` channels."""
    c = c0
    for i in range(nb_layers):
        base = "%sconv%d_%d" % (pre, stage, i + 1)
        a1 = net.fold(buf.view(0, c), base + "_x1_bn", base + "_x1_scale", EPS_DENSE, learn, 0.99, bn_train, sc_train)
        y1 = net.conv(base + "_x1", [a1], 4 * growth, _k(nd, 1), bias=False, trainable=k_train)
        a2 = net.fold(y1, base + "_x2_bn", base + "_x2_scale", EPS_DENSE, learn, 0.99, bn_train, sc_train)
        net.conv(base + "_x2", [a2], growth, _k(nd, 3), p=_p(nd, 1), bias=False, out=buf.view(c, growth),
                 trainable=k_train)
        c += growth
    return c


def transition_block(net, buf, c, stage, nd, pre, learn, bn_train, sc_train, k_train, out):
    """hybridnet.py:301-327 / :71-97: BN->Scale->ReLU->1x1 conv (C -> C/2) -> average pool 2x2 (x1)."""
    base = "%sconv%d_blk" % (pre, stage)
    a = net.fold(buf.view(0, c), base + "_bn", base + "_scale", EPS_DENSE, learn, 0.99, bn_train, sc_train)
    y = net.conv(base, [a], c // 2, _k(nd, 1), bias=False, trainable=k_train)
    net.avgpool(y, out)


def dense_unet_2d(net, x, mode, skip=False):
    """2-D DenseUNet-161 on x (N,1,H,W,3).  Returns (feature Act (N,1,H,W,64), logits view (N,1,H,W,3))."""
    m = mode
    N, H, W = x.N, x.H, x.W
    nb_layers, growth = [6, 12, 36, 24], 48
    bn = dict(learn=m.bn2d_learn, bn_train=m.bn2d_train, sc_train=m.sc2d_train, k_train=m.k2d)
    box = []
    y = net.conv("conv1", [x], 96, (1, 7, 7), s=(1, 2, 2), p=(0, 3, 3), bias=False, trainable=m.k2d)
    a = net.fold(y, "conv1_bn", "conv1_scale", EPS_DENSE, m.bn2d_learn, 0.99, m.bn2d_train, m.sc2d_train)
    box.append(a)
    c = 96
    h, w = H // 4, W // 4
    bufs = []
    for bi in range(4):
        stage = bi + 2
        cend = c + nb_layers[bi] * growth
        buf = net.buffer("block%d" % stage, N, 1, h, w, cend)
        bufs.append(buf)
        if bi == 0:
            net.maxpool(a, buf.view(0, c), pool_d=False)                       # hybridnet.py:215-216
        else:
            transition_block(net, bufs[bi - 1], cprev, stage - 1, 2, "", out=buf.view(0, c), **bn)
        dense_block(net, buf, c, stage, nb_layers[bi], growth, 2, "", **bn)
        box.append(Act(buf.view(0, cend)))
        cprev = cend
        c = cend // 2
        h, w = h // 2, w // 2
    a = net.fold(bufs[3].view(), "conv5_blk_bn", "conv5_blk_scale", EPS_DENSE, m.bn2d_learn, 0.99, m.bn2d_train,
                 m.sc2d_train)
    box[4] = a
    widths = [768, 384, 96, 96, 64]
    for k in range(5):
        srcs = [Src(a, (1, 2, 2))]
        if skip and k < 4:
            if k == 0:
                l0 = net.conv("line0", [box[3]], 2208, (1, 1, 1), init=random_normal, trainable=m.k2d)   # denseunet.py:190
                srcs = [Src(Act(l0)), srcs[0]]
            else:
                srcs = [Src(box[3 - k]), srcs[0]]                                                       # denseunet.py:197-209
        drop = 0.3 if (k == 4 and m.name == "2d") else 0.0                                              # densenet.py:92
        y = net.conv("conv_up%d" % k, srcs, widths[k], (1, 3, 3), p=(0, 1, 1), init=m.dec_init, trainable=m.k2d,
                     drop_rate=drop)
        a = net.fold(y, "bn_up%d" % k, None, 1e-3, m.bn2d_learn, 0.99, m.bn2d_train)
    logits = net.conv("dense167classifer", [a], 3, (1, 1, 1), init=m.dec_init, trainable=m.k2d)
    return a, logits


def dense_net_3d(net, x, mode):
    """3-D DenseNet (hybridnet.py:98-178) on x (N,S,H,W,4).  Returns the feature Act (N,S,H,W,64).
    Reference axes (H,W,S) <-> engine (D=S,H,W); pooling (2,2,1) leaves D alone."""
    m = mode
    N, S, H, W = x.N, x.D, x.H, x.W
    nb_layers, growth = [3, 4, 12, 8], 32
    bn = dict(learn=m.dense3d_learn, bn_train=m.dense3d_bn_train, sc_train=True, k_train=True)
    y = net.conv("3dconv1", [x], 96, (7, 7, 7), s=(2, 2, 2), p=(3, 3, 3), bias=False)
    a = net.fold(y, "3dconv1_bn", "3dconv1_scale", EPS_DENSE, True, 0.99)
    c = 96
    d, h, w = S // 4, H // 4, W // 4
    bufs = []
    for bi in range(4):
        stage = bi + 2
        cend = c + nb_layers[bi] * growth
        buf = net.buffer("3dblock%d" % stage, N, d, h, w, cend)
        bufs.append(buf)
        if bi == 0:
            net.maxpool(a, buf.view(0, c), pool_d=True)                        # hybridnet.py:128-129
        else:
            tb = dict(bn)
            tb["bn_train"] = m.trans3d_bn_train
            transition_block(net, bufs[bi - 1], cprev, stage - 1, 3, "3d", out=buf.view(0, c), **tb)
        dense_block(net, buf, c, stage, nb_layers[bi], growth, 3, "3d", **bn)
        cprev = cend
        c = cend // 2
        h, w = h // 2, w // 2
    a = net.fold(bufs[3].view(), "3dconv5_blk_bn", "3dconv5_blk_scale", EPS_DENSE, True, 0.99)
    widths = [504, 224, 192, 96, 64]
    ups = [(1, 2, 2), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2)]               # hybridnet.py:151-171 (H,W,S)->(S,H,W)
    for k in range(5):
        y = net.conv("3dconv_up%d" % k, [Src(a, ups[k])], widths[k], (3, 3, 3), p=(1, 1, 1))
        a = net.fold(y, "3dbn_up%d" % k, None, 1e-3, True, 0.99)
    return a


def hybrid_net(net, B, S, H, W, mode):
    """hybridnet.py:379-423 / denseunet3d.py:393-439.  Input 'volumetric_data' (B,S,H,W,1).
    The 2-D network's (B*S,1,H,W,C) outputs are re-read in place as (B,S,H,W,C) volumes."""
    vol = net.input("volumetric_data", B, S, H, W, 1)
    in2d_buf = net.buffer("input2d", B * S, 1, H, W, 4)      # 3 slices + a zero channel: 16-byte pixels
    net.triplets(vol, in2d_buf.view())
    in2d = in2d_buf.view(0, 3)
    feat2d, logits2d = dense_unet_2d(net, in2d, mode)
    in3d = net.buffer("input3d", B, S, H, W, 4).view()
    net.cat4(vol, logits2d.as_nd(B, S), in3d)
    feat3d = dense_net_3d(net, in3d, mode)
    f2 = Act(feat2d.view.as_nd(B, S), feat2d.fold, feat2d.relu)
    y = net.conv("fianl_conv", [Src(feat3d), Src(f2)], 64, (3, 3, 3), p=(1, 1, 1), drop_rate=mode.final_drop)
    a = net.fold(y, "final_bn", None, 1e-3, True, 0.99)
    logits = net.conv("2d3dclassifer", [a], 3, (1, 1, 1))
    net.outputs["logits"] = logits
    net.outputs["logits2d"] = logits2d
    net.outputs["feature2d"] = feat2d
    return logits


def unet2d_net(net, N, H, W, mode, skip=False):
    x = net.input("data", N, 1, H, W, 4).sub(0, 3)            # 3 slices + a zero channel: 16-byte pixels
    feat, logits = dense_unet_2d(net, x, mode, skip)
    net.outputs["logits"] = logits
    net.outputs["feature"] = feat
    return logits


def net3d_only(net, N, S, H, W, mode):
    """The 3-D DenseNet + hybrid head fed directly with a 4-channel volume (BASELINE config 3).
    The 2-D feature source of `fianl_conv` is dropped (single-source add)."""
    x = net.input("input3d", N, S, H, W, 4)
    feat3d = dense_net_3d(net, x, mode)
    y = net.conv("fianl_conv", [Src(feat3d)], 64, (3, 3, 3), p=(1, 1, 1), drop_rate=mode.final_drop)
    a = net.fold(y, "final_bn", None, 1e-3, True, 0.99)
    logits = net.conv("2d3dclassifer", [a], 3, (1, 1, 1))
    net.outputs["logits"] = logits
    net.outputs["feature3d"] = feat3d
    return logits
