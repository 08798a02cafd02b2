"""Keras-HDF5 weight files: the pure-Python HDF5 subset (h5lite) and the Keras layouts and loaders
on top of it (keras_h5, Model.save_weights / save / load_weights, ModelCheckpoint).

The reader is checked against the one file in the reference tree that the real HDF5 library wrote
(Keras-2.0.8/examples/mymodel.h5, a full `model.save` of an MNIST CNN): group tree and attributes against the committed
golden (tests/golden/h5_mymodel_golden.json, made by tests/golden/make_h5_golden.py) and -- independently of h5lite's own
output -- dataset shapes against the layer configuration stored in the file's `model_config` JSON attribute."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
REAL = "/root/reference/Keras-2.0.8/examples/mymodel.h5"

import h_denseunet_b200 as hdn  # noqa: E402
from h_denseunet_b200 import h5lite, keras_h5  # noqa: E402
from util import Args  # noqa: E402


@pytest.mark.skipif(not os.path.isfile(REAL), reason="the reference tree is not present on this machine")
def test_reader_on_a_file_written_by_the_real_library():
    import make_h5_golden
    got = make_h5_golden.describe(REAL)
    with open(os.path.join(ROOT, "tests", "golden", "h5_mymodel_golden.json")) as fh:
        exp = json.load(fh)
    assert json.loads(json.dumps(got, sort_keys=True)) == exp
    f = h5lite.File(REAL)
    cfg = json.loads(f.attrs["model_config"].decode() if isinstance(f.attrs["model_config"], bytes) else f.attrs["model_config"])
    layers = {l["config"]["name"]: l for l in cfg["config"]}
    mw = f["model_weights"]
    assert [n.decode() for n in mw.attrs["layer_names"]] == [l["config"]["name"] for l in cfg["config"]]
    c1, c2, d2 = layers["conv2d_1"]["config"], layers["conv2d_2"]["config"], layers["dense_2"]["config"]
    assert mw["conv2d_1/conv2d_1/kernel:0"].shape == tuple(c1["kernel_size"]) + (1, c1["filters"])
    assert mw["conv2d_2/conv2d_2/kernel:0"].shape == tuple(c2["kernel_size"]) + (c1["filters"], c2["filters"])
    assert mw["dense_2/dense_2/bias:0"].shape == (d2["units"],)
    k = mw["conv2d_2/conv2d_2/kernel:0"].value()
    assert k.dtype == np.float32 and np.isfinite(k).all() and 0.01 < k.std() < 1.0       # trained weights, not noise bytes


def test_hdf5_round_trip_many_links(tmp_path):
    """> 256 links in one group: two B-tree levels; names with '/' become nested groups; attribute kinds."""
    p = str(tmp_path / "t.h5")
    w = h5lite.Writer(p)
    rng = np.random.default_rng(0)
    names = ["layer_%04d" % i for i in range(700)]
    w.root.attrs["layer_names"] = [n.encode() for n in names]
    w.root.attrs["backend"] = b"tensorflow"
    w.root.attrs["count"] = np.int64(700)
    w.root.attrs["scales"] = np.arange(5, dtype=np.float32)
    vals = {}
    for n in names:
        g = w.root.group(n)
        g.attrs["weight_names"] = [(n + "/kernel:0").encode(), (n + "/bias:0").encode()]
        vals[n + "/kernel:0"] = rng.normal(size=(3, 3, 2, 5)).astype(np.float32)
        vals[n + "/bias:0"] = rng.normal(size=(5,)).astype(np.float32)
        g.dataset(n + "/kernel:0", vals[n + "/kernel:0"])
        g.dataset(n + "/bias:0", vals[n + "/bias:0"])
    w.root.group("empty")
    w.close()
    f = h5lite.File(p)
    assert f.keys() == sorted(names + ["empty"]) and f["empty"].keys() == []
    assert [x.decode() for x in f.attrs["layer_names"]] == names
    assert f.attrs["backend"] == b"tensorflow" and int(f.attrs["count"]) == 700
    assert np.array_equal(f.attrs["scales"], np.arange(5, dtype=np.float32))
    for n in names[::37]:
        g = f[n]
        assert g.keys() == [n] and g[n].keys() == ["bias:0", "kernel:0"]          # sorted, as h5py lists old-style groups
        for wn in g.attrs["weight_names"]:
            assert np.array_equal(g[wn.decode()].value(), vals[wn.decode()])
    with pytest.raises(KeyError):
        f["nope"]
    with open(p, "rb") as fh:
        assert fh.read(8) == h5lite.SIG


def _models():
    a2 = Args(b=1, input_size=32)
    a3 = Args(b=1, input_size=32, input_cols=8)
    return (hdn.DenseUNet(reduction=0.5, args=a2, device="cpu", backend=object()),
            hdn.dense_rnn_net(a3, device="cpu", backend=object()))


def _randomise(m, seed):
    rng = np.random.default_rng(seed)
    for p in m.params.order:
        m.params.set_value(p.name, rng.normal(0, 1, p.shape).astype(np.float32))


def test_keras_layouts_and_the_three_loaders(tmp_path):
    m2, mh = _models()
    _randomise(m2, 1)
    _randomise(mh, 2)
    w2, wh = m2.get_weights_dict(), mh.get_weights_dict()
    # (1) save_weights -> load_weights, same architecture (train_hybrid.py:152, test.py:49)
    p = str(tmp_path / "hybrid.h5")
    mh.save_weights(p)
    f = h5lite.File(p)
    assert f.attrs["keras_version"] == b"2.0.8"
    ln = [x.decode() for x in f.attrs["layer_names"]]
    assert ln[:3] == ["conv1", "conv1_bn", "conv1_scale"] and "fianl_conv" in ln and "2d3dclassifer" in ln
    assert [x.decode() for x in f["conv1_bn"].attrs["weight_names"]] == ["conv1_bn/gamma:0", "conv1_bn/beta:0", "conv1_bn/moving_mean:0",
                                                                           "conv1_bn/moving_variance:0"]
    assert [x.decode() for x in f["conv1_scale"].attrs["weight_names"]] == ["conv1_scale/conv1_scale_gamma:0", "conv1_scale/conv1_scale_beta:0"]
    assert f["3dconv_up4/3dconv_up4/kernel:0"].shape == (3, 3, 3, 96, 64)          # Keras layout (kh, kw, ks, I, O)
    _, mh2 = _models()
    mh2.load_weights(p)
    for k, v in mh2.get_weights_dict().items():
        assert np.array_equal(v, wh[k]), k
    # (2) by_name: the 2-D trainer's file initialises the layers of the same name, the rest keeps its values
    #     (train_2ddense.py:179 loads densenet161_weights_tf.h5 this way)
    p2 = str(tmp_path / "d161.h5")
    m2.save_weights(p2)
    _, mh3 = _models()
    _randomise(mh3, 3)
    before = mh3.get_weights_dict()
    mh3.load_weights(p2, by_name=True)
    after = mh3.get_weights_dict()
    for k in after:
        assert np.array_equal(after[k], w2[k] if k in w2 else before[k]), k
    with pytest.raises(ValueError):
        mh3.load_weights(p2)                                                       # topological load of another architecture
    # (3) by_gpu + two_model + by_flag: a make_parallel'ed 2-D model's checkpoint (one layer group `denseu161`) into the
    #     hybrid model's 2-D branch (train_hybrid.py:146); by_flag=False reads `auto3d_residual_conv`
    p3 = str(tmp_path / "par2d.hdf5")
    keras_h5.write(p3, w2, layout="nested:denseu161", full=True)
    g = h5lite.File(p3)["model_weights/denseu161"]
    assert g["conv1_bn"].keys() == ["beta:0", "gamma:0", "moving_mean:0", "moving_variance:0"]     # the order the swap fixes
    assert g["conv_up0"].keys() == ["bias:0", "kernel:0"]
    _, mh4 = _models()
    _randomise(mh4, 4)
    before = mh4.get_weights_dict()
    mh4.load_weights(p3, by_name=True, by_gpu=True, two_model=True, by_flag=True)
    after = mh4.get_weights_dict()
    for k in after:
        assert np.array_equal(after[k], w2[k] if k in w2 else before[k]), k
    p4 = str(tmp_path / "parh.hdf5")
    keras_h5.write(p4, wh, layout="nested:auto3d_residual_conv", full=True)
    _, mh5 = _models()
    mh5.load_weights(p4, by_name=True, by_gpu=True, two_model=True, by_flag=False)
    for k, v in mh5.get_weights_dict().items():
        assert np.array_equal(v, wh[k]), k
    # (4) by_gpu alone: the tree under `model_1` (topology.py:3199)
    p5 = str(tmp_path / "par1.hdf5")
    keras_h5.write(p5, w2, layout="nested:model_1")
    m2b, _ = _models()
    m2b.load_weights(p5, by_name=True, by_gpu=True)
    for k, v in m2b.get_weights_dict().items():
        assert np.array_equal(v, w2[k]), k


def test_model_checkpoint_file_names_and_optimizer_state(tmp_path):
    """callbacks.py:335-432 with the scripts' template (train_hybrid.py:205): 0-based epoch in the name, a real .hdf5,
    loadable by load_weights, carrying the optimizer state (save_weights_only=False)."""
    m2, _ = _models()
    _randomise(m2, 5)
    m2.global_step = 17
    cb = hdn.ModelCheckpoint(str(tmp_path / "weights.{epoch:02d}-{loss:.2f}.hdf5"), monitor="loss", verbose=0, save_best_only=False,
                             save_weights_only=False, mode="min", period=1)
    cb.set_model(m2)
    cb.on_epoch_end(0, {"loss": 0.5})
    path = str(tmp_path / "weights.00-0.50.hdf5")
    assert os.path.isfile(path) and not os.path.exists(path + ".npz")
    f = h5lite.File(path)
    assert f.keys() == ["model_weights", "optimizer_weights"]
    assert int(f["optimizer_weights/training/SGD/iterations:0"].value()) == 17
    m2b, _ = _models()
    m2b.load_weights(path)
    assert m2b.global_step == 17
    for k, v in m2b.get_weights_dict().items():
        assert np.array_equal(v, m2.get_weights_dict()[k]), k
    # the npz container of round 1 is still readable through its base name
    m2.save_weights(str(tmp_path / "old"))
    m2c, _ = _models()
    m2c.load_weights(str(tmp_path / "old"))
    with pytest.raises(IOError):
        m2c.load_weights(str(tmp_path / "missing.h5"))


def test_loss_history_file(tmp_path):
    os.makedirs(str(tmp_path / "history"))
    h = hdn.keras_api.LossHistory(str(tmp_path))
    h.on_epoch_end(0, {"loss": 0.123456})
    h.on_epoch_end(1, {"loss": 2.0})
    assert open(str(tmp_path / "history" / "lossepoch.txt")).read() == "0.1235\n2.0000\n"
