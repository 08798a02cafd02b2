"""`-m "not gpu"`: the C-ABI library builds for sm_100a, loads, and exports every symbol include/hdn.h declares (no
compute call is made: there is no GPU here); argument validation that needs no device; the drop-in import shims."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "hdn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hdn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import h_denseunet_b200._lib as L
    L.build()
    lib = L.load()
    names = _header_functions()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(L.EXPORTS) == names, (set(L.EXPORTS) ^ set(names))
    assert lib.hdn_version() >= 100


def test_struct_sizes_match_header():
    """ctypes mirrors vs a tiny C program compiled against include/hdn.h."""
    import h_denseunet_b200._lib as L
    code = r'''
#include <stdio.h>
#include "hdn.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(hdn_tensor), sizeof(hdn_src), sizeof(hdn_conv), sizeof(hdn_dgrad_epi),
         sizeof(hdn_pool), sizeof(hdn_bn_fold_t), sizeof(hdn_bn_grad_t));
  return 0;
}'''
    d = os.path.join(ROOT, "tests", "_tmp_cabi")
    os.makedirs(d, exist_ok=True)
    cfile, exe = os.path.join(d, "sz.c"), os.path.join(d, "sz")
    open(cfile, "w").write(code)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe], check=True)
    got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    exp = [C.sizeof(t) for t in (L.Tensor, L.Src, L.Conv, L.DgradEpi, L.Pool, L.BnFold, L.BnGrad)]
    assert got == exp, (got, exp)


def test_bad_descriptors_are_rejected_without_a_device():
    import h_denseunet_b200._lib as L
    lib = L.load()
    c = L.Conv()
    assert lib.hdn_conv_fprop(C.byref(c), None) == -1                 # HDN_ERR_ARG
    assert b"non-positive" in lib.hdn_last_error()
    assert lib.hdn_conv_tc_supported(C.byref(c), 0) == 0
    assert lib.hdn_sgd_nesterov(None, None, None, 0, 0.0, 0.0, 1.0, None) != 0


def test_product_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import h_denseunet_b200 as hdn
    from util import Args
    import numpy as np
    m = hdn.dense_rnn_net(Args(1, 32, 8))
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        m.predict(np.zeros((1, 32, 32, 8, 1), np.float32))


def test_reference_scripts_import_surface():
    """The names train_2ddense.py:10-17, train_hybrid.py:8-15 and test.py:4-11 import resolve to this engine when
    h-denseunet_b200/compat is first on the path."""
    code = '''
import sys
sys.path.insert(0, %r)
from keras.optimizers import SGD
from keras.callbacks import ModelCheckpoint
import keras.backend as K
from keras.utils2.multi_gpu import make_parallel
from hybridnet import dense_rnn_net
from denseunet3d import denseunet_3d
from denseunet import DenseUNet
from loss import weighted_crossentropy, weighted_crossentropy_2ddense
from lib.custom_layers import Scale
from lib.funcs import predict_tumor_inwindow
K.set_image_dim_ordering("tf")
class A: b = 1; input_size = 64; input_cols = 8
m = dense_rnn_net(A)
m.compile(optimizer=SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[weighted_crossentropy])
m2 = DenseUNet(reduction=0.5, args=A)
assert make_parallel(m2, 1, mini_batch=10) is m2
print(m.name, m.count_params(), m2.name, m2.count_params())
''' % os.path.join(ROOT, "h-denseunet_b200", "compat")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    name, n, name2, n2 = r.stdout.split()
    assert name == "auto3d_residual_conv" and name2 == "denseu161"
    assert int(n) == 61444622                                        # 2-D 45.3 M + 3-D and head 16.1 M
