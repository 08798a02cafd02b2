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
import contextlib
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
        self.global_step = 0          # optimizer steps over the life of the weights (saved with them; drives the dropout stream)
        self.run_seed = int(os.environ.get("HDN_RUN_SEED", "0"))
        self._staging = {}
        self._staging_ev = {}
        self._staging_used = {}
        self._copy_stream = None
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
                rank = self.dp.rank if self.dp is not None else 0
                net.seed_salt = (self.run_seed * 0x9E3779B1 + (rank + 1) * 0x85EBCA77) & 0xFFFFFFFF
                net.step = self.global_step
            net.compile()
            if getattr(self, "_pending_moms", None) is not None and len(self._pending_moms) == self.params.n_train:
                self.params.moms[:self.params.n_train].copy_(torch.from_numpy(self._pending_moms))   # optimizer state of a loaded checkpoint
                self._pending_moms = None
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
        x = self._check_input(x)
        buf = list(net.inputs.values())[0]
        self._h2d(x, buf.data, three_d=self.kind != "unet2d")
        return x.nbytes

    # ---- host -> device staging ------------------------------------------------------------------------------
    # Two slots of (pinned host buffer, device staging buffer) per array shape.  _stage copies the host array into the
    # pinned buffer and starts the H2D copy on a dedicated copy stream; _unstage makes the compute stream wait for that
    # copy and converts the staged reference layout into the engine's NDHWC buffer (hdn_layout_nhws_to_nshw for the
    # (N,H,W,S,1) volumes and label maps; the int16 -> fp32 conversion of label maps happens there, on the device).
    # fit_generator stages batch k+1 while step k runs, so the upload leaves the critical path.
    def _stage(self, x, dev, slot=0):
        tdt = torch.int16 if x.dtype == np.int16 else torch.float32
        if tdt == torch.float32 and x.dtype != np.float32:
            x = np.asarray(x, dtype=np.float32)
        key = (x.shape, dev, tdt, slot)
        st = self._staging.get(key)
        cuda = dev.type == "cuda"
        if st is None:
            if cuda and self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            # the device staging buffer belongs to the COPY stream's allocator pool: a block taken from the compute stream's
            # pool may still be read by a kernel queued there (a freed temporary), and the H2D copy would overwrite it
            with (torch.cuda.stream(self._copy_stream) if cuda else contextlib.nullcontext()):
                dbuf = torch.empty(x.shape, dtype=tdt, device=dev)
            st = (torch.empty(x.shape, dtype=tdt, pin_memory=cuda), dbuf)
            self._staging[key] = st
        ev = self._staging_ev.get(key)
        if ev is not None:
            ev.synchronize()          # the previous async copy out of this pinned buffer must have finished
        st[0].copy_(torch.from_numpy(np.ascontiguousarray(x)))
        if cuda:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(self._copy_stream):
                used = self._staging_used.get(key)
                if used is not None:
                    self._copy_stream.wait_event(used)     # the compute stream's last reader of this device staging buffer is done
                st[1].copy_(st[0], non_blocking=True)
                if ev is None:
                    ev = self._staging_ev[key] = torch.cuda.Event()
                ev.record(self._copy_stream)
        else:
            st[1].copy_(st[0])
        return key, x.nbytes

    def _unstage(self, key, dst, three_d):
        shape, dev, tdt, _ = key
        src = self._staging[key][1]
        if dev.type == "cuda":
            torch.cuda.current_stream().wait_event(self._staging_ev[key])
        if three_d and shape[-1] == 1 and dev.type == "cuda" and self.backend is None:
            from . import _lib
            N, H, W, S = shape[:4]
            _lib.check(_lib.load().hdn_layout_nhws_to_nshw(src.data_ptr(), dst.data_ptr(), N, H, W, S, 1 if tdt == torch.int16 else 0,
                                                           torch.cuda.current_stream().cuda_stream), "hdn_layout_nhws_to_nshw")
        elif three_d:                 # (N,H,W,S,C) -> (N,S,H,W,C)
            dst.copy_(src.permute(0, 3, 1, 2, 4))
        elif dst.shape[-1] != shape[-1] and dst.dim() == 5 and len(shape) == 4:
            # 2-D input (N,H,W,3) into the 4-channel padded buffer (N,1,H,W,4); channel 3 stays zero
            dst[:, 0, :, :, :shape[-1]].copy_(src)
        else:
            dst.view(shape).copy_(src)
        if dev.type == "cuda":
            used = self._staging_used.get(key)
            if used is None:
                used = self._staging_used[key] = torch.cuda.Event()
            used.record()

    def _h2d(self, x, dst, three_d):
        """host (reference layout) -> device NDHWC through the staging slots (see _stage / _unstage)."""
        key, nbytes = self._stage(np.asarray(x), dst.device)
        self._unstage(key, dst, three_d)
        return nbytes

    def _label_array(self, y):
        y = np.asarray(y)
        if y.dtype != np.int16:       # the reference feeds int16 label maps (train_hybrid.py:127-132); anything else as float
            y = np.asarray(y, dtype=np.float32)
        if self.kind == "unet2d":
            return y.reshape(self.b, self.size, self.size)
        return y.reshape(self.b, self.size, self.size, self.cols, 1)

    def _label_dst(self, net):
        lab = net.loss.labels
        return lab if self.kind == "unet2d" else lab.view(lab.shape + (1,))

    def _labels(self, net, y):
        return self._h2d(self._label_array(y), self._label_dst(net), three_d=self.kind != "unet2d")

    def _check_input(self, x):
        x = np.asarray(x, dtype=np.float32)
        if tuple(x.shape) != self.input_shape:
            raise ValueError("expected input of shape %s, got %s" % (self.input_shape, tuple(x.shape)))
        return x

    def _stage_batch(self, net, x, y, slot):
        """Start the upload of one (x, y) batch into staging slot `slot`; returns the handle _launch_staged takes."""
        if self.dp is not None and len(x) == self.b * self.dp.world:
            x, y = x[self.dp.rank * self.b:(self.dp.rank + 1) * self.b], y[self.dp.rank * self.b:(self.dp.rank + 1) * self.b]
        dev = net.device
        kx, nx = self._stage(self._check_input(x), dev, slot)
        ky, ny = self._stage(self._label_array(y), dev, slot)
        return kx, ky, nx + ny

    def _next_staged(self, net, generator, slot):
        item = next(generator)
        if getattr(item, "is_device_batch", False):       # augment.CropGenerator: the batch already lies on the device
            return item
        return self._stage_batch(net, *item[:2], slot=slot)

    def _consume_device_batch(self, net, batch):
        """A batch produced on the device (augment.DeviceBatch, engine layout): no host copy, one D2D copy each."""
        x_dst = list(net.inputs.values())[0].data
        y_dst = net.loss.labels
        if batch.x.numel() != x_dst.numel() or batch.y.numel() != y_dst.numel():
            raise ValueError("device batch %s / %s does not match the model's input %s / labels %s"
                             % (tuple(batch.x.shape), tuple(batch.y.shape), tuple(x_dst.shape), tuple(y_dst.shape)))
        cur = torch.cuda.current_stream()
        cur.wait_event(batch.ready)
        x_dst.view(batch.x.shape).copy_(batch.x)
        y_dst.view(batch.y.shape).copy_(batch.y)
        used = torch.cuda.Event()
        used.record(cur)
        batch.release(used)
        self.h2d_bytes = 0

    def _launch_staged(self, net, handle):
        """Enqueue one optimizer step on a staged batch (no host synchronisation); the loss is read by the caller."""
        if getattr(handle, "is_device_batch", False):
            self._consume_device_batch(net, handle)
            self._step_kernels(net)
            return
        kx, ky, nbytes = handle
        three_d = self.kind != "unet2d"
        self._unstage(kx, list(net.inputs.values())[0].data, three_d)
        self._unstage(ky, self._label_dst(net), three_d)
        self.h2d_bytes = nbytes
        self._step_kernels(net)

    def train_on_batch(self, x, y, **kwargs):
        if self.optimizer is None:
            raise RuntimeError("You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.")
        net = self._net(True)
        if self.dp is not None and len(x) == self.b * self.dp.world:
            # the reference feeds the merged batch and slices it into towers (multi_gpu.py:20-33); here the tower is this rank
            x, y = x[self.dp.rank * self.b:(self.dp.rank + 1) * self.b], y[self.dp.rank * self.b:(self.dp.rank + 1) * self.b]
        self.h2d_bytes = self._upload(net, x) + self._labels(net, y)
        return self.train_step_device(net)

    def train_step_device(self, net=None):
        """One optimizer step on whatever is resident in the input / label buffers."""
        net = net or self._net(True)
        self._step_kernels(net)
        return net.loss.value()

    def _step_kernels(self, net):
        if self.dp is not None:
            self.dp.begin_step()          # peers have delivered the last update and released this replica's gradients
        net.forward()
        net.backward()
        self.global_step = net.step
        o = self.optimizer
        if self.dp is not None:
            self.dp.step(net, o.lr, o.momentum)
        else:
            ps = self.params
            net.be.sgd(ps.train, ps.grads, ps.moms, ps.n_train, o.lr, o.momentum, 1.0)

    def predict(self, x, batch_size=None, verbose=0, **kwargs):
        net = self._net(False)
        if self.dp is not None:
            self.dp.begin_step()
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
        callbacks = list(callbacks or [])
        if verbose and not any(isinstance(cb, LossHistory) for cb in callbacks):
            callbacks.append(LossHistory())
        history = {"loss": []}
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        if self.optimizer is None:
            raise RuntimeError("You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.")
        net = self._net(True)
        total = (int(epochs) - int(initial_epoch)) * int(steps_per_epoch)
        slot, done = 0, 0
        staged = self._next_staged(net, generator, slot) if total > 0 else None
        for epoch in range(initial_epoch, epochs):
            losses = []
            t0 = time.time()
            for _ in range(int(steps_per_epoch)):
                # step k is enqueued, then batch k+1 is fetched from the generator and staged (pinned copy + H2D on the copy
                # stream) while the GPU works; only then is step k's loss read back (the one host synchronisation per step)
                self._launch_staged(net, staged)
                done += 1
                if done < total:
                    slot ^= 1
                    staged = self._next_staged(net, generator, slot)
                losses.append(net.loss.value())
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
        if self.dp is not None and self.params.realised:
            self.dp.begin_step()          # every peer's shard of the last update has landed
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
        if self.dp is not None and self.dp.ready:
            self.dp._sync_initial(self.params)      # replicas stay identical: rank 0's values win (a collective call)

    @staticmethod
    def _is_h5(path):
        return path.endswith((".h5", ".hdf5", ".hdf"))

    def _optimizer_state(self):
        if not self.params.realised:
            return {"training/SGD/iterations:0": np.asarray(self.global_step, np.int64)}
        return {"training/SGD/iterations:0": np.asarray(self.global_step, np.int64),
                "training/SGD/momentum_arena:0": self.params.moms[:self.params.n_train].detach().cpu().numpy()}

    def save_weights(self, path, overwrite=True):
        """topology.py:2555-2590: a Keras-layout HDF5 file for *.h5 / *.hdf5 paths (keras_h5.write), else the npz container."""
        if not overwrite and os.path.isfile(path):
            raise IOError("%s exists and overwrite=False" % path)
        if self._is_h5(path):
            from . import keras_h5
            keras_h5.write(path, self.get_weights_dict())
        else:
            np.savez(path if path.endswith(".npz") else path + ".npz", **self.get_weights_dict())

    def save(self, path, overwrite=True, include_optimizer=True):
        """models.py:56-130 `save_model`: weights under `model_weights` plus the optimizer state (momentum arena and the
        step counter) under `optimizer_weights`, which is what ModelCheckpoint(save_weights_only=False) writes."""
        if not overwrite and os.path.isfile(path):
            raise IOError("%s exists and overwrite=False" % path)
        opt = self._optimizer_state() if include_optimizer else None
        if self._is_h5(path):
            from . import keras_h5
            keras_h5.write(path, self.get_weights_dict(), full=True, optimizer=opt)
        else:
            d = dict(self.get_weights_dict())
            if opt:
                d.update({"__optimizer__/" + k: v for k, v in opt.items()})
            np.savez(path if path.endswith(".npz") else path + ".npz", **d)

    def _restore_optimizer(self, opt):
        it = opt.get("training/SGD/iterations:0")
        if it is not None:
            self.global_step = int(np.asarray(it))
            for net in self.nets.values():
                net.step = self.global_step
        m = opt.get("training/SGD/momentum_arena:0")
        if m is not None and self.params.realised and len(m) == self.params.n_train:
            self.params.moms[:self.params.n_train].copy_(torch.from_numpy(np.asarray(m, np.float32)))
        elif m is not None:
            self._pending_moms = np.asarray(m, np.float32)

    def load_weights(self, path, by_name=False, by_gpu=False, two_model=False, by_flag=False):
        """topology.py:2590-2630 incl. the reference's by_gpu / two_model / by_flag loaders (keras_h5.read).  A path
        that does not exist falls back to `<path>.npz` (the container of round 1's checkpoints)."""
        if self._is_h5(path) and os.path.isfile(path):
            from . import keras_h5
            got, info = keras_h5.read(path, [p.name for p in self.params.order], by_name=by_name, by_gpu=by_gpu,
                                      two_model=two_model, by_flag=by_flag)
            self.set_weights_dict(got, strict=True)
            if info.get("optimizer"):
                self._restore_optimizer(info["optimizer"])
            return
        npz = path if path.endswith(".npz") else path + ".npz"
        if not os.path.isfile(npz):
            raise IOError("Unable to open file (no such file: %s)" % path)
        with np.load(npz) as z:
            self.set_weights_dict({k: z[k] for k in z.files if not k.startswith("__optimizer__/")}, strict=not by_name)
            opt = {k[len("__optimizer__/"):]: z[k] for k in z.files if k.startswith("__optimizer__/")}
        if opt:
            self._restore_optimizer(opt)

    def count_params(self):
        return sum(p.size for p in self.params.order)


class ModelCheckpoint(object):
    """callbacks.py:335-432: `filepath.format(epoch=epoch, **logs)` with the 0-based epoch (callbacks.py:404), every
    `period` epochs; save_weights_only=False (what the three scripts pass) saves the full model incl. optimizer state."""

    def __init__(self, filepath, monitor="loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", period=1):
        self.filepath, self.monitor, self.verbose, self.period, self.model = filepath, monitor, verbose, period, None
        self.save_best_only, self.save_weights_only = save_best_only, save_weights_only
        self.epochs_since_last_save = 0
        self.best = -np.inf if mode == "max" else np.inf
        self.better = np.greater if mode == "max" else np.less

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        path = self.filepath.format(epoch=epoch, **logs)
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None or not self.better(cur, self.best):
                return
            self.best = cur
        if self.verbose:
            print("Epoch %05d: saving model to %s" % (epoch, path))
        if self.save_weights_only:
            self.model.save_weights(path, overwrite=True)
        else:
            self.model.save(path, overwrite=True)


class LossHistory(object):
    """The reference's ProgbarLogger edit (callbacks.py:311-314): append '%.4f' of the epoch loss to
    <path>/history/lossepoch.txt, path = './Experiments/' (callbacks.py:28).  fit_generator installs it when verbose
    and the directory exists (the reference crashes when it does not)."""

    def __init__(self, path=None):
        self.path = path or os.environ.get("HDN_EXPERIMENTS_PATH", "./Experiments/")

    def on_epoch_end(self, epoch, logs=None):
        d = os.path.join(self.path, "history")
        if os.path.isdir(d) and logs and "loss" in logs:
            with open(os.path.join(d, "lossepoch.txt"), "a") as fh:
                fh.write("%.4f\n" % logs["loss"])


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
        if mini_batch is not None and int(mini_batch) != model.b:
            # multi_gpu.py:20-33: every tower sees `mini_batch` samples (BN batch statistics over those, not over args.b)
            if model.nets:
                raise RuntimeError("make_parallel(mini_batch=%d) after the model was realised with batch %d" % (mini_batch, model.b))
            model.b = int(mini_batch)
        model.dp = DataParallel()
    return model
