import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import _root  # noqa: F401,E402
from h_denseunet_b200 import make_parallel  # noqa: F401,E402   (Keras-2.0.8/keras/utils2/multi_gpu.py:7)
