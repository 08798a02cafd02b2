"""Builder functions with the reference's names and signatures."""
from .keras_api import Model


class _Args(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


def DenseUNet(nb_dense_block=4, growth_rate=48, nb_filter=96, reduction=0.0, dropout_rate=0.0, weight_decay=1e-4,
              classes=1000, weights_path=None, args=None, skip=False, **engine_kw):
    """densenet.py:10-101 (train_2ddense.py:178 calls `DenseUNet(reduction=0.5, args=args)`); with
    skip=True the denseunet.py:130-226 topology (line0 + skip adds).  As in the reference the
    architecture arguments are overwritten: nb_filter=96, nb_layers=[6,12,36,24], growth 48
    (densenet.py:40-41); `weight_decay` is accepted and unused (never applied, hybridnet.py:264)."""
    if args is None:
        raise ValueError("DenseUNet needs args with .b and .input_size (densenet.py:34)")
    if reduction != 0.5:
        raise ValueError("the reference only ever builds reduction=0.5 (train_2ddense.py:178)")
    m = Model("unet2d", "denseu161", "2d", args.b, args.input_size, skip=skip, dropout=True, **engine_kw)
    if weights_path is not None:
        m.load_weights(weights_path)
    return m


def dense_rnn_net(args, **engine_kw):
    """hybridnet.py:379-423: H-DenseUNet, end-to-end fine-tuning variant ('auto3d_residual_conv')."""
    return Model("hybrid", "auto3d_residual_conv", "end2end", args.b, args.input_size, args.input_cols,
                 dropout=True, **engine_kw)


def denseunet_3d(args, **engine_kw):
    """denseunet3d.py:393-439: H-DenseUNet with the 2-D branch frozen ('3dpart')."""
    return Model("hybrid", "auto3d_residual_conv", "3dpart", args.b, args.input_size, args.input_cols,
                 dropout=True, **engine_kw)


def DenseNet3D(args, mode="end2end", **engine_kw):
    """hybridnet.py:98-178 + head, fed with a ready 4-channel volume (BASELINE config 3)."""
    return Model("net3d", "densenet3d", mode, args.b, args.input_size, args.input_cols, dropout=True, **engine_kw)


DenseUNet161 = DenseUNet      # north-star spelling
DenseUNet3d = DenseNet3D      # north-star spelling


class Scale(object):
    """lib/custom_layers.py:10-74.  In this engine a Scale layer is never a separate op: its
    gamma/beta are folded with the preceding BatchNormalization into the consumer's load
    (engine.Fold).  The class exists so `from lib.custom_layers import Scale` keeps working and
    carries the reference's constructor signature."""

    def __init__(self, weights=None, axis=-1, momentum=0.9, beta_init="zero", gamma_init="one", **kwargs):
        self.axis, self.momentum, self.initial_weights, self.name = axis, momentum, weights, kwargs.get("name")
