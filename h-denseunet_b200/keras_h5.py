"""Keras-2.0.8 HDF5 weight files for the engine's parameter dict.

Writing  (Keras-2.0.8/keras/engine/topology.py:2555-2590 `save_weights`, :2845-2872 `save_weights_to_hdf5_group`,
          models.py:56-130 `save_model`):
    root attrs `layer_names`, `backend`, `keras_version`; one group per layer with attr `weight_names` and one dataset
    per weight named after the TF variable -- "<layer>/kernel:0", "<layer>/gamma:0", Scale's
    "<layer>/<layer>_gamma:0" (lib/custom_layers.py:53-54 inside the layer's name scope) -- which h5py turns into
    nested groups "<layer>/<layer>/kernel:0".  `save` puts the same tree under `model_weights` and adds
    `optimizer_weights` (here: the Nesterov momentum arena + the step counter).
    layout="nested:<model>" writes what a make_parallel'ed model saves: ONE layer group named after the inner model
    (`model_1`, `denseu161`, `auto3d_residual_conv`) holding every inner weight (multi_gpu.py:55-69).

Reading  (topology.py:2590-2630 `load_weights` and the reference's three additions, :3171-3330):
    by_name=False           topological: here the layer names of file and model must agree (they do for every
                            checkpoint the reference scripts exchange); otherwise ValueError
    by_name=True            layers found in the file by name, missing ones keep their values (:3107-3168)
    by_name, by_gpu         `load_weights_from_hdf5_group_by_name_mulgpu` (:3171-3243): the tree under `model_1`;
                            weights taken in h5py's key order (sorted) with the first two swapped
    by_name, by_gpu, two_model   `..._twomodelcombine` (:3245-3330): the tree under `denseu161` (by_flag) or
                            `auto3d_residual_conv`; same order rule, swap only for 2 or 4 weights
"""
import numpy as np

from . import h5lite

ORDER = {"kernel": 0, "bias": 1, "gamma": 0, "beta": 1, "moving_mean": 2, "moving_variance": 3}


def layers_of(names):
    """[(layer, [weight, ...])] in first-appearance order, weights in Keras' `layer.weights` order."""
    out, idx = [], {}
    for n in names:
        layer, w = n.rsplit("/", 1)
        if layer not in idx:
            idx[layer] = len(out)
            out.append((layer, []))
        out[idx[layer]][1].append(w)
    return [(l, sorted(ws, key=lambda w: ORDER[w])) for l, ws in out]


def tf_weight_name(layer, w, is_scale):
    return "%s/%s_%s:0" % (layer, layer, w) if is_scale else "%s/%s:0" % (layer, w)


def _is_scale(layer, ws):
    return layer.endswith("_scale") and ws == ["gamma", "beta"]


def write(path, weights, layout="flat", full=False, optimizer=None):
    """weights: {"<layer>/<weight>": array in Keras layout} (Model.get_weights_dict()).  layout "flat" or
    "nested:<inner model name>"; full=True wraps everything in `model_weights` (Model.save)."""
    w = h5lite.Writer(path)
    root = w.root.group("model_weights") if full else w.root
    lay = layers_of(list(weights))
    root.attrs["backend"] = b"tensorflow"
    root.attrs["keras_version"] = b"2.0.8"
    if full:
        w.root.attrs["keras_version"] = b"2.0.8"
        w.root.attrs["backend"] = b"tensorflow"
    if layout == "flat":
        root.attrs["layer_names"] = [l.encode("utf8") for l, _ in lay]
        for layer, ws in lay:
            g = root.group(layer)
            sc = _is_scale(layer, ws)
            names = [tf_weight_name(layer, x, sc) for x in ws]
            g.attrs["weight_names"] = [n.encode("utf8") for n in names]
            for x, n in zip(ws, names):
                g.dataset(n, np.asarray(weights["%s/%s" % (layer, x)], np.float32))
    elif layout.startswith("nested:"):
        inner = layout.split(":", 1)[1]
        root.attrs["layer_names"] = [inner.encode("utf8")]
        g = root.group(inner)
        names = []
        for layer, ws in lay:
            sc = _is_scale(layer, ws)
            for x in ws:
                n = tf_weight_name(layer, x, sc)
                names.append(n)
                g.dataset(n, np.asarray(weights["%s/%s" % (layer, x)], np.float32))
        # (weight_names of ~2000 entries would exceed the 64 KiB attribute limit of this file format, as it does for
        # h5py: the reference's nested loaders never read it -- they walk the group's keys, topology.py:3199,3215)
    else:
        raise ValueError("layout must be 'flat' or 'nested:<name>'")
    if optimizer is not None:
        og = w.root.group("optimizer_weights")
        og.attrs["weight_names"] = [k.encode("utf8") for k in optimizer]
        for k, v in optimizer.items():
            og.dataset(k, np.asarray(v))
    w.close()


def _leaf_to_weight(layer, leaf):
    """'kernel:0' -> 'kernel';  '<layer>_gamma:0' -> 'gamma'."""
    base = leaf.split(":")[0]
    if base.startswith(layer + "_"):
        base = base[len(layer) + 1:]
    return base


def read(path, wanted, by_name=False, by_gpu=False, two_model=False, by_flag=False):
    """Returns ({"<layer>/<weight>": array}, info) for the parameters named in `wanted` that the file provides, following
    the loader the flags select.  `wanted`: iterable of the model's parameter names."""
    f = h5lite.File(path)
    info = {}
    if "layer_names" not in f.attrs and "model_weights" in f:
        if "optimizer_weights" in f:
            og = f["optimizer_weights"]
            wn = og.attrs.get("weight_names")
            wn = [] if wn is None else [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in np.asarray(wn).ravel()]
            info["optimizer"] = {k: og[k].value() for k in wn}
        f = f["model_weights"]
    model_layers = layers_of(list(wanted))
    mdict = dict(model_layers)
    out = {}

    def take(layer, values_in_keras_order, k=0):
        ws = mdict[layer]
        if len(values_in_keras_order) != len(ws):
            raise ValueError('Layer #%d (named "%s") expects %d weight(s), but the saved weights have %d element(s).' % (
                k, layer, len(ws), len(values_in_keras_order)))
        for wname, v in zip(ws, values_in_keras_order):
            out["%s/%s" % (layer, wname)] = v

    if by_name and by_gpu:
        if two_model:
            f = f["denseu161"] if by_flag else f["auto3d_residual_conv"]
        else:
            f = f["model_1"]
        for k, name in enumerate(f.keys()):
            g = f[name]
            if not g.is_group:
                continue
            leaves = g.keys()                              # h5py key order of an old-style group: sorted
            if len(leaves) in (2, 4) or (not two_model and len(leaves) >= 2):
                leaves[0], leaves[1] = leaves[1], leaves[0]
            if name in mdict:
                take(name, [g[x].value() for x in leaves], k)
        return out, info
    layer_names = [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in np.asarray(f.attrs["layer_names"]).ravel()]
    file_layers = []
    for name in layer_names:
        g = f[name]
        wn = g.attrs.get("weight_names")
        wn = [] if wn is None else [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in np.asarray(wn).ravel()]
        if wn:
            file_layers.append((name, g, wn))
    if not by_name:
        # topological loading (topology.py:3047-3105): same number of weighted layers, taken in order.  The engine
        # matches them by name and insists that the two lists agree, which is what "same architecture" means here.
        fl, ml = [n for n, _, _ in file_layers], [l for l, _ in model_layers]
        if len(fl) != len(ml):
            raise ValueError("You are trying to load a weight file containing %d layers into a model with %d layers." % (len(fl), len(ml)))
        if set(fl) != set(ml):
            raise ValueError("weight file and model disagree on layer names (first differences: %s); use by_name=True" % (
                sorted(set(fl) ^ set(ml))[:6],))
    for k, (name, g, wn) in enumerate(file_layers):
        if name in mdict:
            take(name, [g[x].value() for x in wn], k)
    return out, info
