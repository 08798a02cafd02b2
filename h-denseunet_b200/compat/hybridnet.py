"""Shim for hybridnet.py: dense_rnn_net(args) (hybridnet.py:379)."""
import _root  # noqa: F401
from h_denseunet_b200 import dense_rnn_net, DenseNet3D  # noqa: F401
