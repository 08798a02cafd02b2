"""hdense-b200: Blackwell-native forward/backward engine for H-DenseUNet.

Public surface (the reference's builder names):
    DenseUNet(...)            densenet.DenseUNet / denseunet.DenseUNet  -> 2-D DenseUNet-161 ('denseu161')
    dense_rnn_net(args)       hybridnet.dense_rnn_net                    -> H-DenseUNet end2end
    denseunet_3d(args)        denseunet3d.denseunet_3d                   -> H-DenseUNet '3dpart'
    DenseUNet161, DenseUNet3d aliases named by the build brief
    SGD, make_parallel, ModelCheckpoint, weighted_crossentropy, weighted_crossentropy_2ddense
    predict_tumor_inwindow    lib.funcs.predict_tumor_inwindow
    augment.DeviceVolumes / augment.CropGenerator   generate_arrays_from_file of the training scripts, on the device
The import name is `h_denseunet_b200` (the directory carries the project's hyphenated name).
"""
from .builders import (DenseUNet, DenseUNet161, DenseUNet3d, DenseNet3D, dense_rnn_net, denseunet_3d, Scale)
from .keras_api import (Model, SGD, ModelCheckpoint, make_parallel, weighted_crossentropy,
                        weighted_crossentropy_2ddense)
from .inference import predict_tumor_inwindow
from .postprocess import postprocess_scores
from . import keras_api, keras_h5, h5lite, preprocessing, augment

__all__ = ["DenseUNet", "DenseUNet161", "DenseUNet3d", "DenseNet3D", "dense_rnn_net", "denseunet_3d", "Scale",
           "Model", "SGD", "ModelCheckpoint", "make_parallel", "weighted_crossentropy",
           "weighted_crossentropy_2ddense", "predict_tumor_inwindow", "postprocess_scores"]
