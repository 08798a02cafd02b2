"""CPU ORACLE for the H-DenseUNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  Nothing under the product package imports it.

PARITY UNPINNED for conv / batch-norm / pooling / SGD / losses / whole
models: the reference's arithmetic lives in TensorFlow 1.x (requirements.txt:71-72,
`tensorflow==1.5.1`), which is neither vendored under /root/reference nor importable
here, and the reference ships no golden vectors for this path.
What IS pinned (tests/test_oracle.py):
  * upsampling == np.repeat      (Keras-2.0.8/tests/keras/layers/convolutional_test.py:673-681,726-736)
  * zero padding                 (same file :508-628)
  * add / concatenate            (Keras-2.0.8/tests/keras/layers/merge_test.py:13-30,142-177)
  * softmax, relu                (Keras-2.0.8/tests/keras/activations_test.py:53-68,158-164: the upstream test's own
                                  NumPy reference on its standard values)
  * every op against an independent direct-loop NumPy fp64 restatement of the TF op
    semantics (oracle/naive_ops.py) on tiny shapes.

This file restates, in PyTorch-CPU fp32 (fp64 on request), the graph the reference
builders construct and the TF-1.x op semantics the vendored Keras backend maps to:
  topology   hybridnet.py:11-423, densenet.py:10-193, denseunet.py:130-226,
             denseunet3d.py:18-439
  Scale      lib/custom_layers.py:62-69
  losses     loss.py:5-46
  window     lib/funcs.py:4-51
  ops        Keras-2.0.8/keras/backend/tensorflow_backend.py (KB) :1620-1684 (BN),
             :1739-1840 (upsample), :1989-2060 (pad), :3128-3165 / :3277-3314 (conv),
             :3354-3432 (pool), :915-927 (moving average)
  SGD        Keras-2.0.8/keras/optimizers.py:155-186

Tensor conventions at this file's boundary are the reference's: 2D tensors NHWC,
3D tensors (N, H, W, S, C) with S the slice axis; conv kernels HWIO / (kh,kw,ks,I,O).
Parameters are a flat dict  "<layer>/<weight>" -> numpy array, e.g. "conv1/kernel",
"conv1_bn/gamma|beta|moving_mean|moving_variance", "conv1_scale/gamma|beta".
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS_DENSE = 1.1e-5   # hybridnet.py:21,59,82,112,196
EPS_KERAS = 1e-3     # KNORM:59 default (decoder BNs)
CLASS_W = (0.78, 0.65, 8.57)   # loss.py:23


class Ctx:
    """Carries parameters (as torch leaf tensors), run mode and collected BN updates."""

    def __init__(self, params, training, dtype=torch.float32, requires_grad=False,
                 dropout_masks=None):
        self.dtype = dtype
        self.training = training
        self.p = {}
        for k, v in params.items():
            t = torch.tensor(np.asarray(v), dtype=dtype)
            if requires_grad and not k.endswith(("moving_mean", "moving_variance")):
                t.requires_grad_(True)
            self.p[k] = t
        self.bn_updates = {}      # name -> (batch_mean, batch_var) for train-mode BNs
        self.dropout_masks = dropout_masks or {}
        self.taps = {}            # optional named intermediate activations


# ----------------------------------------------------------------------------- ops
def _nd(x):
    return x.dim() - 2


def conv(ctx, x, name, strides=1, padding="valid", use_bias=True):
    """KCONV:157-172 -> KB:3158 / KB:3307 tf.nn.convolution: cross-correlation,
    kernel layout (*k, Cin, Cout).  x is channel-first (N,C,*spatial)."""
    w = ctx.p[name + "/kernel"]
    nd = _nd(x)
    wt = w.permute(nd + 1, nd, *range(nd))        # -> (Cout, Cin, *k)
    b = ctx.p[name + "/bias"] if use_bias else None
    if padding == "same":
        # stride-1 odd kernels only on this path (k=3 or 1): symmetric pad (k-1)/2
        pad = [(k - 1) // 2 for k in w.shape[:nd]]
    else:
        pad = 0
    fn = F.conv2d if nd == 2 else F.conv3d
    return fn(x, wt, b, stride=strides, padding=pad)


def zero_pad(x, p):
    """ZeroPadding2D/3D -> KB:2020 tf.pad with zeros (symmetric p per spatial axis)."""
    nd = _nd(x)
    return F.pad(x, [p, p] * nd)


def bn(ctx, x, name, eps, learn, momentum=0.99):
    """KNORM:126-190.  learn=False <=> the builder passed training=False (frozen).
    Training mode: KB:1635 tf.nn.moments (population variance) + tf.nn.batch_normalization;
    moving stats get the BIASED batch variance (KNORM:179-185, KB:915-927)."""
    g = ctx.p[name + "/gamma"]
    b = ctx.p[name + "/beta"]
    shape = [1, -1] + [1] * _nd(x)
    if learn and ctx.training:
        axes = [0] + list(range(2, x.dim()))
        mean = x.mean(dim=axes)
        var = x.var(dim=axes, unbiased=False)
        ctx.bn_updates[name] = (mean.detach(), var.detach(), momentum)
    else:
        mean = ctx.p[name + "/moving_mean"]
        var = ctx.p[name + "/moving_variance"]
    inv = torch.rsqrt(var + eps) * g
    return x * inv.view(shape) + (b - mean * inv).view(shape)


def scale(ctx, x, name):
    """lib/custom_layers.py:62-69: gamma*x+beta along the channel axis."""
    shape = [1, -1] + [1] * _nd(x)
    return ctx.p[name + "/gamma"].view(shape) * x + ctx.p[name + "/beta"].view(shape)


def upsample(x, size):
    """UpSampling2D/3D: KB:1764-1771 (nearest) / KB:1797-1827 (repeat) == np.repeat per axis."""
    for ax, s in enumerate(size):
        if s != 1:
            x = x.repeat_interleave(s, dim=2 + ax)
    return x


def max_pool(x, k, s):
    fn = F.max_pool2d if _nd(x) == 2 else F.max_pool3d
    return fn(x, k, s)


def avg_pool(x, k):
    fn = F.avg_pool2d if _nd(x) == 2 else F.avg_pool3d
    return fn(x, k, k)


def dropout(ctx, x, name, rate):
    """core.py:72-112 -> KB:2888 tf.nn.dropout: x * mask / keep_prob in training.
    The TF RNG is not reproducible; parity runs inject the mask (or leave it out => identity)."""
    if ctx.training and name in ctx.dropout_masks:
        m = torch.as_tensor(ctx.dropout_masks[name], dtype=x.dtype)
        return x * m / (1.0 - rate)
    return x


# ------------------------------------------------------------------------- 2D net
def conv_block(ctx, x, stage, branch, learn, pre=""):
    """hybridnet.py:264-298 / densenet.py:103-137 (2D), hybridnet.py:11-45 (3D, pre='3d')."""
    base = "%sconv%d_%d" % (pre, stage, branch)
    y = bn(ctx, x, base + "_x1_bn", EPS_DENSE, learn)
    y = relu(scale(ctx, y, base + "_x1_scale"))
    y = conv(ctx, y, base + "_x1", use_bias=False)
    y = bn(ctx, y, base + "_x2_bn", EPS_DENSE, learn)
    y = relu(scale(ctx, y, base + "_x2_scale"))
    y = zero_pad(y, 1)
    return conv(ctx, y, base + "_x2", use_bias=False)


def dense_block(ctx, x, stage, nb_layers, learn, pre=""):
    """hybridnet.py:330-354 / :46-70: concat grows by growth_rate per layer."""
    feat = x
    for i in range(nb_layers):
        y = conv_block(ctx, feat, stage, i + 1, learn, pre)
        feat = torch.cat([feat, y], dim=1)
    return feat


def transition_block(ctx, x, stage, learn, pre=""):
    """hybridnet.py:301-327 (2D: avgpool 2x2) / :71-97 (3D: avgpool (2,2,1))."""
    base = "%sconv%d_blk" % (pre, stage)
    y = bn(ctx, x, base + "_bn", EPS_DENSE, learn)
    y = relu(scale(ctx, y, base + "_scale"))
    y = conv(ctx, y, base, use_bias=False)
    return avg_pool(y, 2 if _nd(y) == 2 else (2, 2, 1))


def dense_unet_2d(ctx, x, learn_bn, skip=False, dropout_rate=0.0):
    """2D DenseUNet-161.  hybridnet.py:182-262 (learn_bn=False, all BN frozen inference),
    densenet.py:10-101 (learn_bn=True, Dropout(.3) after conv_up4),
    denseunet.py:130-226 (skip=True: line0 + add([box[k], up])).
    x: (N,3,H,W) channel-first.  Returns (feature2d, logits2d) channel-first."""
    nb_layers = [6, 12, 36, 24]
    box = []
    y = zero_pad(x, 3)
    y = conv(ctx, y, "conv1", strides=2, use_bias=False)
    y = bn(ctx, y, "conv1_bn", EPS_DENSE, learn_bn)
    y = relu(scale(ctx, y, "conv1_scale"))
    box.append(y)
    y = max_pool(zero_pad(y, 1), 3, 2)
    for bi in range(3):
        stage = bi + 2
        y = dense_block(ctx, y, stage, nb_layers[bi], learn_bn)
        box.append(y)
        y = transition_block(ctx, y, stage, learn_bn)
    y = dense_block(ctx, y, 5, nb_layers[3], learn_bn)
    y = bn(ctx, y, "conv5_blk_bn", EPS_DENSE, learn_bn)
    y = relu(scale(ctx, y, "conv5_blk_scale"))
    box.append(y)
    for k in range(5):
        y = upsample(y, (2, 2))
        if skip and k < 4:
            if k == 0:
                y = conv(ctx, box[3], "line0", padding="same") + y     # denseunet.py:190-191
            else:
                y = box[3 - k] + y                                      # denseunet.py:197,203,209
        y = conv(ctx, y, "conv_up%d" % k, padding="same")
        if k == 4 and dropout_rate:
            y = dropout(ctx, y, "conv_up4_dropout", dropout_rate)      # densenet.py:92
        y = relu(bn(ctx, y, "bn_up%d" % k, EPS_KERAS, learn_bn))
    feat = y
    logits = conv(ctx, feat, "dense167classifer", padding="same")
    return feat, logits


# ------------------------------------------------------------------------- 3D net
def dense_net_3d(ctx, x, learn_dense, learn_other=True):
    """hybridnet.py:98-178 (learn_dense=False: conv_block / transition BNs inference),
    denseunet3d.py:105-190 (learn_dense=True).  x: (N,4,H,W,S).  Returns feature3d.
    The '3dclassifer' head (hybridnet.py:176) is not reachable from the model output."""
    nb_layers = [3, 4, 12, 8]
    y = zero_pad(x, 3)
    y = conv(ctx, y, "3dconv1", strides=2, use_bias=False)
    y = bn(ctx, y, "3dconv1_bn", EPS_DENSE, learn_other)
    y = relu(scale(ctx, y, "3dconv1_scale"))
    y = max_pool(zero_pad(y, 1), 3, 2)
    for bi in range(3):
        stage = bi + 2
        y = dense_block(ctx, y, stage, nb_layers[bi], learn_dense, pre="3d")
        y = transition_block(ctx, y, stage, learn_dense, pre="3d")
    y = dense_block(ctx, y, 5, nb_layers[3], learn_dense, pre="3d")
    y = bn(ctx, y, "3dconv5_blk_bn", EPS_DENSE, learn_other)
    y = relu(scale(ctx, y, "3dconv5_blk_scale"))
    ups = [(2, 2, 1), (2, 2, 1), (2, 2, 1), (2, 2, 2), (2, 2, 2)]
    for k in range(5):
        y = upsample(y, ups[k])
        y = conv(ctx, y, "3dconv_up%d" % k, padding="same")
        y = relu(bn(ctx, y, "3dbn_up%d" % k, EPS_KERAS, learn_other))
    return y


def slice_triplets(vol):
    """hybridnet.py:385-396: (1,H,W,S,1) -> (S,H,W,3); slice s gets [s-1,s,s+1] with the two
    edge slices replicated ([0,0,1] and [S-2,S-1,S-1]).  vol: (B,H,W,S) -> (B*S,3,H,W)."""
    B, H, W, S = vol.shape
    idx = torch.arange(S)
    tri = torch.stack([(idx - 1).clamp(0, S - 1), idx, (idx + 1).clamp(0, S - 1)], dim=1)  # (S,3)
    g = vol[:, :, :, tri]                       # (B,H,W,S,3)
    return g.permute(0, 3, 4, 1, 2).reshape(B * S, 3, H, W)


def hybrid_net(ctx, vol, variant="end2end", dropout_rate=None):
    """hybridnet.py:379-423 (variant 'end2end') / denseunet3d.py:393-439 ('3dpart').
    vol: (B,H,W,S,1) reference layout.  Returns logits (B,H,W,S,3).
    Batch semantics: B independent slabs (the reference is only correct for b=1,
    hybridnet.py:390,395)."""
    B, H, W, S, _ = vol.shape
    v = torch.as_tensor(vol, dtype=ctx.dtype)[..., 0]
    in2d = slice_triplets(v)
    feat2d, log2d = dense_unet_2d(ctx, in2d, learn_bn=False)
    ctx.taps["logits2d"] = log2d
    # (B*S,C,H,W) -> (B,C,H,W,S)      hybridnet.py:359-364,400-406
    f2 = feat2d.reshape(B, S, -1, H, W).permute(0, 2, 3, 4, 1)
    l2 = log2d.reshape(B, S, -1, H, W).permute(0, 2, 3, 4, 1)
    in3d = torch.cat([v.unsqueeze(1), l2 * 250.0], dim=1)            # hybridnet.py:409-411
    feat3d = dense_net_3d(ctx, in3d, learn_dense=(variant == "3dpart"))
    y = feat3d + f2                                                  # hybridnet.py:414
    y = conv(ctx, y, "fianl_conv", padding="same")
    rate = dropout_rate if dropout_rate is not None else (0.3 if variant == "end2end" else 0.1)
    y = dropout(ctx, y, "fianl_conv_dropout", rate)
    y = relu(bn(ctx, y, "final_bn", EPS_KERAS, True))
    y = conv(ctx, y, "2d3dclassifer", padding="same")
    return y.permute(0, 2, 3, 4, 1)                                  # (B,H,W,S,3)


# -------------------------------------------------------------------------- losses
def softmax(x, axis=-1):
    """K.softmax (KB:2708 -> tf.nn.softmax over the last axis).  Pinned by the upstream known-answer test
    Keras-2.0.8/tests/keras/activations_test.py:53-68 (tests/test_oracle.py)."""
    return torch.softmax(torch.as_tensor(x), dim=axis)


def relu(x):
    """K.relu (KB:2671, alpha = 0, no max_value).  Pinned by activations_test.py:158-164."""
    return F.relu(torch.as_tensor(x))


def weighted_crossentropy(y_true, y_pred, crop=True):
    """loss.py:5-25 (crop=True: hybrid, drops first/last slice; `1:7` generalised to `1:S-1`,
    identical for the reference's S=8) and loss.py:27-46 (crop=False: 2D).
    y_pred (...,3) logits channel-last, y_true (...,1) or (...) labels.
    Voxels whose label is not exactly 0,1,2 are excluded from numerator AND denominator."""
    if crop:
        S = y_pred.shape[3]
        y_pred = y_pred[:, :, :, 1:S - 1, :]
        y_true = y_true[:, :, :, 1:S - 1]
    lp = y_pred.reshape(-1, 3)
    yt = torch.as_tensor(y_true, dtype=lp.dtype).reshape(-1)
    sm = softmax(lp, axis=1)
    lg = torch.log(torch.clamp(sm, 1e-10, 1.0))
    w = torch.tensor(CLASS_W, dtype=lp.dtype)
    tot = lp.new_zeros(())
    cnt = 0
    for c in range(3):
        m = yt == float(c)
        tot = tot + w[c] * lg[m, c].sum()
        cnt += int(m.sum())
    return -tot / max(cnt, 1)


# ------------------------------------------------------------------ model wrappers
def _to_cf(x):
    x = torch.as_tensor(x)
    return x.permute(0, x.dim() - 1, *range(1, x.dim() - 1))


def _to_cl(x):
    return x.permute(0, *range(2, x.dim()), 1)


def forward_2d(params, x_nhwc, training=False, learn_bn=True, skip=False, dtype=torch.float32,
               requires_grad=False):
    ctx = Ctx(params, training, dtype, requires_grad)
    feat, logits = dense_unet_2d(ctx, _to_cf(torch.as_tensor(x_nhwc, dtype=dtype)), learn_bn, skip)
    return ctx, _to_cl(feat), _to_cl(logits)


def forward_hybrid(params, vol, training=False, variant="end2end", dtype=torch.float32,
                   requires_grad=False, dropout_masks=None):
    """dropout_masks: {'fianl_conv_dropout': 0/1 array in the internal (B, C, H, W, S) layout} -- see dropout()."""
    ctx = Ctx(params, training, dtype, requires_grad, dropout_masks)
    return ctx, hybrid_net(ctx, vol, variant)


def forward_3d(params, x, training=False, learn_dense=False, dtype=torch.float32, requires_grad=False):
    """3D DenseNet alone on a (N,H,W,S,4) input -> feature3d (N,H,W,S,64)."""
    ctx = Ctx(params, training, dtype, requires_grad)
    f = dense_net_3d(ctx, _to_cf(torch.as_tensor(x, dtype=dtype)), learn_dense)
    return ctx, _to_cl(f)


def grads_of(ctx, loss):
    names = [k for k, t in ctx.p.items() if t.requires_grad]
    gs = torch.autograd.grad(loss, [ctx.p[k] for k in names], allow_unused=True)
    return {k: (g.numpy() if g is not None else None) for k, g in zip(names, gs)}


def sgd_nesterov_step(p, g, m, lr=1e-3, mu=0.9):
    """optimizers.py:172-181: v = mu*m - lr*g; m <- v; p <- p + mu*v - lr*g."""
    v = mu * m - lr * g
    return p + mu * v - lr * g, v


def moving_average_update(mov, batch, momentum):
    """KB:915-927 assign_moving_average, zero_debias=False: mov -= (mov-batch)*(1-momentum)."""
    return mov - (mov - batch) * (1.0 - momentum)


# ------------------------------------------------------------------ sliding window
def window_starts(z, mini_z, maxi_z, cols):
    """lib/funcs.py:12,19-26: z-window start list (py2 integer division) incl. the tail clamp."""
    step = cols // 4
    right = int(min(z, maxi_z + 10) - cols)
    left = max(0, min(mini_z - 5, right))
    out = []
    for c in range(left, right + step, step):
        out.append(z - cols if c > z - cols else c)
    return out


def predict_tumor_inwindow(predict_fn, imgs, num, mini, maxi, size, cols):
    """lib/funcs.py:4-51.  predict_fn: (1,size,size,cols,1) float32 -> logits (1,size,size,cols,num).
    Returns the class num-2 and num-1 probability volumes."""
    x, y, z = imgs.shape
    score = np.zeros((x, y, z, num), np.float32)
    cnt = np.zeros((x, y, z, num), np.int16)
    box = np.zeros((1, size, size, cols, 1), np.float32)
    for c in window_starts(z, mini[2], maxi[2], cols):
        box[0, :, :, :, 0] = imgs[0:size, 0:size, c:c + cols]
        lg = torch.as_tensor(predict_fn(box))
        pm = softmax(lg, axis=-1).numpy()[:, :, :, 1:-1, :]
        score[0:size, 0:size, c + 1:c + cols - 1, :] += pm[0]
        cnt[0:size, 0:size, c + 1:c + cols - 1, :] += 1
    score = score / (cnt + 1e-4)
    return score[..., num - 2], score[..., num - 1]


def dice(a, b):
    a = np.asarray(a, bool)
    b = np.asarray(b, bool)
    s = a.sum() + b.sum()
    return 1.0 if s == 0 else 2.0 * np.logical_and(a, b).sum() / s
