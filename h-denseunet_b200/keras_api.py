"""The slice of the Keras-2.0.8 Model surface that train_2ddense.py / train_hybrid.py / test.py use
(SURVEY.md section 8b, B1), over the engine.  numpy in / numpy out, reference tensor layouts:
2-D (N,H,W,C); 3-D (N,H,W,S,C).

  Model.compile(optimizer=SGD(...), loss=[fn])         Keras-2.0.8/keras/engine/training.py:570
  Model.train_on_batch(x, y) -> loss                   training.py:1715-1765
  Model.predict(x, batch_size, verbose) -> logits      training.py:1659
  Model.fit_generator(gen, steps_per_epoch, epochs, ...)   training.py:1831-2092
  Model.load_weights / save_weights / save             topology.py:2555-2630 (npz container here;
                                                       the HDF5 reader is a SURVEY 8f 'next' row)
  SGD(lr, momentum, nesterov)                          optimizers.py:130-186
"""
import os
import time

import numpy as np
import torch

from . import engine, models


class SGD(object):
    def __init__(self, lr=0.01, momentum=0.0, decay=0.0, nesterov=False, **kwargs):
        if not nesterov:
            raise ValueError("only the Nesterov form used by the reference scripts is implemented "
                             "(train_hybrid.py:150: SGD(lr=1e-3, momentum=0.9, nesterov=True))")
        if decay != 0.0:
            raise ValueError("lr decay is not used by the reference scripts and is not implemented")
        self.lr, self.momentum, self.nesterov, self.decay = float(lr), float(momentum), True, 0.0


def weighted_crossentropy(y_true, y_pred):
    """loss.py:5-25 marker.  Model.compile recognises it; the arithmetic is hdn_wce_accum/grad
    (crop of the first and last slice, `1:7` for the reference's 8 slices)."""
    raise RuntimeError("pass this function to Model.compile(loss=[...]); it is evaluated by the CUDA engine")


def weighted_crossentropy_2ddense(y_true, y_pred):
    """loss.py:27-46 marker (no slice crop)."""
    raise RuntimeError("pass this function to Model.compile(loss=[...]); it is evaluated by the CUDA engine")


weighted_crossentropy.crop = True
weighted_crossentropy_2ddense.crop = False


def is_3d_layer(layer):
    return layer.startswith("3d") or layer in ("fianl_conv", "2d3dclassifer")


class Model(object):
    """A built network.  `kind` in {'unet2d', 'hybrid', 'net3d'}; shapes are the reference's."""

    def __init__(self, kind, name, mode, batch, size, cols=None, skip=False, seed=0, precision=None,
                 backend=None, device=None, dropout=False):
        self.kind, self.name, self.mode = kind, name, models.Mode(mode)
        self.b, self.size, self.cols, self.skip = int(batch), int(size), cols, skip
        self.precision = precision or os.environ.get("HDN_PRECISION", "mixed")
        self.backend = backend
        self.device = device
        self.dropout = dropout
        self.params = engine.ParamStore(seed)
        self.nets = {}
        self.optimizer = None
        self.loss_fn = None
        self.dp = None
        self.stop_training = False
        self._staging = {}
        self._staging_ev = {}
        # register every parameter (host values, Keras default initialisers) without touching a device
        self._build(engine.Net(self.params, "meta", True, "fp32", backend=object(), dropout=False))

    # ---- construction -----------------------------------------------------------------
    def _build(self, net):
        if self.kind == "unet2d":
            models.unet2d_net(net, self.b, self.size, self.size, self.mode, self.skip)
        elif self.kind == "hybrid":
            models.hybrid_net(net, self.b, self.cols, self.size, self.size, self.mode)
        elif self.kind == "net3d":
            models.net3d_only(net, self.b, self.cols, self.size, self.size, self.mode)
        else:
            raise ValueError(self.kind)
        return net

    def _device(self):
        if self.device is not None:
            return torch.device(self.device)
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: this engine has no CPU path")
        return torch.device("cuda", torch.cuda.current_device())

    def _net(self, training):
        net = self.nets.get(training)
        if net is None:
            dev = self._device()
            if self.dp is not None:
                self.dp.realise(self.params, dev)
            net = engine.Net(self.params, dev, training, self.precision, backend=self.backend,
                             dropout=self.dropout and training)
            self._build(net)
            if training:
                net.set_loss(net.outputs["logits"], bool(getattr(self.loss_fn, "crop", self.kind != "unet2d")))
            net.compile()
            self.nets[training] = net
        return net

    # ---- Keras surface ----------------------------------------------------------------
    def compile(self, optimizer=None, loss=None, **kwargs):
        if isinstance(loss, (list, tuple)):
            loss = loss[0]
        if not isinstance(optimizer, SGD):
            raise ValueError("optimizer must be an SGD instance")
        if not hasattr(loss, "crop"):
            raise ValueError("loss must be weighted_crossentropy or weighted_crossentropy_2ddense")
        self.optimizer, self.loss_fn = optimizer, loss

    @property
    def input_shape(self):
        if self.kind == "unet2d":
            return (self.b, self.size, self.size, 3)
        return (self.b, self.size, self.size, self.cols, 4 if self.kind == "net3d" else 1)

    def _upload(self, net, x):
        x = np.asarray(x, dtype=np.float32)
        if tuple(x.shape) != self.input_shape:
            raise ValueError("expected input of shape %s, got %s" % (self.input_shape, tuple(x.shape)))
        buf = list(net.inputs.values())[0]
        self._h2d(x, buf.data, three_d=self.kind != "unet2d")
        return x.nbytes

    def _h2d(self, x, dst, three_d):
        """host (reference layout) -> device NDHWC through a pinned staging buffer.  The staging buffers keep the
        host array's dtype (float32 volumes, int16 label maps -- half the bytes over PCIe and no host-side cast);
        the conversion to the fp32 device tensor happens on the device, inside the layout-changing copy."""
        tdt = torch.int16 if x.dtype == np.int16 else torch.float32
        if tdt == torch.float32 and x.dtype != np.float32:
            x = np.asarray(x, dtype=np.float32)
        key = (x.shape, dst.device, tdt)
        st = self._staging.get(key)
        if st is None:
            pin = dst.device.type == "cuda"
            need_dev = three_d or tdt != torch.float32
            st = (torch.empty(x.shape, dtype=tdt, pin_memory=pin),
                  torch.empty(x.shape, dtype=tdt, device=dst.device) if need_dev else None)
            self._staging[key] = st
        ev = self._staging_ev.get(key)
        if ev is not None:
            ev.synchronize()          # the previous async copy out of this pinned buffer must have finished
        st[0].copy_(torch.from_numpy(np.ascontiguousarray(x)))
        if three_d:    # (N,H,W,S,C) -> (N,S,H,W,C)
            st[1].copy_(st[0], non_blocking=True)
            dst.copy_(st[1].permute(0, 3, 1, 2, 4))
        elif dst.shape[-1] != x.shape[-1] and dst.dim() == 5 and x.ndim == 4:
            # 2-D input (N,H,W,3) into the 4-channel padded buffer (N,1,H,W,4); channel 3 stays zero
            dst[:, 0, :, :, :x.shape[-1]].copy_(st[0], non_blocking=True)
        elif st[1] is not None:
            st[1].copy_(st[0], non_blocking=True)
            dst.view(x.shape).copy_(st[1])
        else:
            dst.view(x.shape).copy_(st[0], non_blocking=True)
        if dst.device.type == "cuda":
            if ev is None:
                ev = self._staging_ev[key] = torch.cuda.Event()
            ev.record()
        return x.nbytes

    def _labels(self, net, y):
        y = np.asarray(y)
        if y.dtype != np.int16:       # the reference feeds int16 label maps (train_hybrid.py:127-132); anything else as float
            y = np.asarray(y, dtype=np.float32)
        lab = net.loss.labels
        if self.kind == "unet2d":
            y = y.reshape(self.b, self.size, self.size)
            return self._h2d(y, lab, three_d=False)
        y = y.reshape(self.b, self.size, self.size, self.cols, 1)
        return self._h2d(y, lab.view(lab.shape + (1,)), three_d=True)

    def train_on_batch(self, x, y, **kwargs):
        if self.optimizer is None:
            raise RuntimeError("You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.")
        net = self._net(True)
        self.h2d_bytes = self._upload(net, x) + self._labels(net, y)
        return self.train_step_device(net)

    def train_step_device(self, net=None):
        """One optimizer step on whatever is resident in the input / label buffers."""
        net = net or self._net(True)
        net.forward()
        net.backward()
        o = self.optimizer
        if self.dp is not None:
            self.dp.step(net, o.lr, o.momentum)
        else:
            ps = self.params
            net.be.sgd(ps.train, ps.grads, ps.moms, ps.n_train, o.lr, o.momentum, 1.0)
        return net.loss.value()

    def predict(self, x, batch_size=None, verbose=0, **kwargs):
        net = self._net(False)
        self._upload(net, x)
        net.forward()
        return self._logits_to_host(net)

    def _logits_to_host(self, net):
        v = net.outputs["logits"]
        t = v.buf.data.view(v.N, v.D, v.H, v.W, 3)
        if self.kind == "unet2d":
            return t.view(v.N, v.H, v.W, 3).cpu().numpy()
        return t.permute(0, 2, 3, 1, 4).contiguous().cpu().numpy()      # (N,S,H,W,3) -> (N,H,W,S,3)

    def fit_generator(self, generator, steps_per_epoch, epochs=1, verbose=1, callbacks=None, max_queue_size=10,
                      workers=1, use_multiprocessing=False, initial_epoch=0, **kwargs):
        callbacks = callbacks or []
        history = {"loss": []}
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        for epoch in range(initial_epoch, epochs):
            losses = []
            t0 = time.time()
            for _ in range(int(steps_per_epoch)):
                x, y = next(generator)[:2]
                losses.append(self.train_on_batch(x, y))
            logs = {"loss": float(np.mean(losses))}
            history["loss"].append(logs["loss"])
            if verbose:
                print("Epoch %d/%d - %.1fs - loss: %.4f" % (epoch + 1, epochs, time.time() - t0, logs["loss"]))
            for cb in callbacks:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        return history

    # ---- weights ------------------------------------------------------------------------
    def get_weights_dict(self):
        """{'<layer>/<weight>': array} in the reference's (Keras) layouts: kernels HWIO / (kh,kw,ks,I,O)."""
        out = {}
        for p in self.params.order:
            v = self.params.get_value(p.name)
            out[p.name] = self._to_keras(p.name, v)
        return out

    def get_grads_dict(self):
        out = {}
        for p in self.params.order:
            g = self.params.get_grad(p.name)
            if g is not None:
                out[p.name] = self._to_keras(p.name, g)
        return out

    @staticmethod
    def _to_keras(name, v):
        layer, wname = name.rsplit("/", 1)
        if wname == "kernel":
            return np.ascontiguousarray(v.transpose(1, 2, 0, 3, 4)) if is_3d_layer(layer) else v[0]
        return v

    @staticmethod
    def _from_keras(name, v):
        layer, wname = name.rsplit("/", 1)
        v = np.asarray(v, dtype=np.float32)
        if wname == "kernel":
            return v.transpose(2, 0, 1, 3, 4) if is_3d_layer(layer) else v[None]
        return v

    def set_weights_dict(self, d, strict=True):
        for k, v in d.items():
            if k not in self.params.params:
                if strict:
                    raise ValueError("unknown weight %s" % k)
                continue
            self.params.set_value(k, self._from_keras(k, v))

    def save_weights(self, path, overwrite=True):
        np.savez(path if path.endswith(".npz") else path + ".npz", **self.get_weights_dict())

    save = save_weights

    def load_weights(self, path, by_name=False, by_gpu=False, two_model=False, by_flag=False):
        """topology.py:2590.  .npz written by save_weights; by_name ignores unknown / missing entries.
        Keras HDF5 files (incl. the by_gpu / two_model / by_flag variants, topology.py:3171-3330)
        need the HDF5 reader listed as the first 'next' row of SURVEY.md 8(f)."""
        if path.endswith((".h5", ".hdf5")):
            raise NotImplementedError("Keras HDF5 weight files are not readable yet (SURVEY.md 8f rank 1); "
                                      "use the .npz written by save_weights")
        with np.load(path if path.endswith(".npz") else path + ".npz") as z:
            self.set_weights_dict({k: z[k] for k in z.files}, strict=not by_name)

    def count_params(self):
        return sum(p.size for p in self.params.order)


class ModelCheckpoint(object):
    """callbacks.py:335-432 (filename template with {epoch} / {loss}); writes the npz container."""

    def __init__(self, filepath, monitor="loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", period=1):
        self.filepath, self.verbose, self.period, self.model = filepath, verbose, period, None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        if (epoch + 1) % self.period == 0:
            path = self.filepath.format(epoch=epoch + 1, **(logs or {}))
            self.model.save_weights(path)
            if self.verbose:
                print("Epoch %05d: saving model to %s" % (epoch + 1, path))


def make_parallel(model, gpu_count, mini_batch=None):
    """Keras-2.0.8/keras/utils2/multi_gpu.py:7-69: data parallel over `gpu_count` devices.
    Here: one process per GPU (torch.distributed), each holding a replica fed `mini_batch`
    samples; gradients are reduced and the SGD update applied by the fused NVLink kernel
    (parallel.DataParallel).  With gpu_count <= 1 or no process group it is the identity."""
    from .parallel import DataParallel
    import torch.distributed as dist
    if gpu_count and gpu_count > 1:
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("make_parallel(gpu_count=%d) needs one process per GPU: launch with "
                               "python -m torch.distributed.run --nproc-per-node %d ..." % (gpu_count, gpu_count))
        if dist.get_world_size() != gpu_count:
            raise ValueError("gpu_count=%d but world size is %d" % (gpu_count, dist.get_world_size()))
        model.dp = DataParallel()
    return model
