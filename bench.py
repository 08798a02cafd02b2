"""bench.py -- CT slabs/s (512x512x48) fwd+bwd of the H-DenseUNet end2end train step on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine (libhdn.so)
    python bench.py --impl reference --gpus N --steps K ...   # the reference graph on the host CPU cores
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py ...)

One "step" = one train_on_batch of hybridnet.dense_rnn_net (train_hybrid.py -arch end2end): forward, weighted
cross-entropy, backward, Nesterov SGD (and for N > 1 the data-parallel gradient exchange) on `--batch` synthetic
slabs per GPU.  `value` times the step with inputs resident in HBM; `e2e` times Model.train_on_batch with HOST
numpy inputs (pinned staging + H2D inside the timed region, loss read back every step).

The reference's arithmetic is TensorFlow-1.x, which cannot be installed here, so the reference arm
and the `cpu_baseline` both time the oracle's PyTorch-CPU restatement of the reference graph (kind "port") on a
bounded sample: one (1, 224, 224, 8) slab -- the reference's own training shape (train_hybrid.py:28-31) -- scaled to
512x512x48-slab units by voxel count; median of >= 5 timed steps after one warm-up step, same thread count in both legs.

--config selects the BASELINE.json configuration (default c4, the headline; the others are kept under profiles/):
    c2  2-D DenseUNet-161 train step (train_2ddense.py), batch 8 x 512x512, training-mode BN + Dropout(.3)
    c3  3-D DenseNet + head fwd+bwd on one 224x224x12 sub-volume (hybridnet.py:98-178)
    c4  H-DenseUNet end2end train step, 512x512x48 slabs, batch 2 per GPU, data parallel 1 -> 8 GPUs
    c5  sliding-window inference of one 512x512x512 volume (test.py / lib/funcs.py:4-51), 253 windows, z-sharded over
        the GPUs, with and without 2-D slice reuse
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stdout carries ONE JSON line: NCCL writes its version banner (any NCCL_DEBUG level >= VERSION) to stdout unless told otherwise
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402
import torch  # noqa: E402

FULL_VOXELS = 512 * 512 * 48

# BASELINE.json configs: metric, unit of one "work unit", model kind
CONFIGS = {
    "c2": {"metric": "2D slices/sec (512x512) DenseUNet-161 train step", "unit": "slices/s"},
    "c3": {"metric": "3D sub-volumes/sec (224x224x12) DenseNet3D fwd+bwd", "unit": "subvolumes/s"},
    "c4": {"metric": "CT slabs/sec (512x512x48) fwd+bwd", "unit": "slabs/s"},
    "c5": {"metric": "CT volumes/sec (512x512x512) sliding-window inference", "unit": "volumes/s"},
}


class Args(object):
    def __init__(self, b, input_size, input_cols):
        self.b, self.input_size, self.input_cols = b, input_size, input_cols


# --------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            try:
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                pass
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "sm_min_mhz": float(min(sm)), "power_w_max": float(max(pw)) if pw else None}


# --------------------------------------------------------------------------------- workloads
def defaults(a):
    """Shape defaults per config (flags override): (batch per GPU, size, cols)."""
    d = {"c2": (8, 512, None), "c3": (1, 224, 12), "c4": (2, 512, 48), "c5": (1, 512, 8)}[a.config]
    a.batch = d[0] if a.batch is None else a.batch
    a.size = d[1] if a.size is None else a.size
    a.cols = d[2] if a.cols is None else a.cols
    # CPU sample of the same workload: the reference's own training shapes (train_hybrid.py:28-31, train_2ddense.py:27)
    a.cpu_size = a.cpu_size or 224
    a.cpu_cols = a.cpu_cols or (12 if a.config == "c3" else 8)
    return a


def build_model(a, precision, seed=0, **kw):
    import h_denseunet_b200 as hdn
    if a.config == "c2":
        return hdn.DenseUNet(reduction=0.5, args=Args(a.batch, a.size, None), precision=precision, seed=seed, **kw)
    if a.config == "c3":
        return hdn.DenseNet3D(Args(a.batch, a.size, a.cols), precision=precision, seed=seed, **kw)
    return hdn.dense_rnn_net(Args(a.batch, a.size, a.cols), precision=precision, seed=seed, **kw)


def host_batch(a, rank):
    """Synthetic host arrays in the reference's layouts."""
    from h_denseunet_b200.synthetic import synthetic_slab
    if a.config == "c2":
        vol, lab = synthetic_slab(1, a.size, max(a.batch + 2, 4), seed=1234 + rank)      # slices of one synthetic volume
        v = vol[0, :, :, :, 0]
        x = np.stack([v[:, :, i:i + 3] for i in range(a.batch)]).astype(np.float32)      # (b, H, W, 3) slice triplets
        y = np.stack([lab[0, :, :, i + 1, :] for i in range(a.batch)]).astype(np.int16)  # (b, H, W, 1)
        return x, y
    vol, lab = synthetic_slab(a.batch, a.size, a.cols, seed=1234 + rank)
    if a.config == "c3":
        rng = np.random.default_rng(7 + rank)
        x = np.concatenate([vol, rng.normal(0, 50, vol.shape[:-1] + (3,)).astype(np.float32)], axis=-1)   # [CT, 250*logits2d]
        return x, lab
    return vol, lab


def units_per_step(a, world):
    if a.config == "c2":
        return a.batch * world * (a.size * a.size) / float(512 * 512)
    if a.config == "c3":
        return a.batch * world * (a.size * a.size * a.cols) / float(224 * 224 * 12)
    return a.batch * world * (a.size * a.size * a.cols) / float(FULL_VOXELS)


def loss_of(a):
    import h_denseunet_b200 as hdn
    return hdn.weighted_crossentropy_2ddense if a.config == "c2" else hdn.weighted_crossentropy


# --------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(a, threads):
    """Returns (fn, sample description, units per call): fn() runs the oracle on a bounded sample of the workload."""
    from oracle import hdense_oracle as orc
    from h_denseunet_b200.synthetic import synthetic_slab
    torch.set_num_threads(threads)
    size, cols = a.cpu_size, a.cpu_cols
    ca = argparse.Namespace(**vars(a))
    ca.size, ca.cols, ca.batch = size, cols, 1
    m = build_model(ca, "fp32", backend=object(), device="cpu")          # parameter dictionary only; no device is touched
    w = m.get_weights_dict()
    mom = {}

    def sgd(g):
        for k, gk in g.items():
            if gk is None:
                continue
            p1, v = orc.sgd_nesterov_step(w[k], gk, mom.get(k, 0.0))
            w[k], mom[k] = p1.astype(np.float32), v

    if a.config == "c2":
        x, y = host_batch(ca, 0)

        def fn():
            ctx, _, logits = orc.forward_2d(w, x, training=True, learn_bn=True, requires_grad=True)
            loss = orc.weighted_crossentropy(torch.as_tensor(y), logits, crop=False)
            sgd(orc.grads_of(ctx, loss))
            return float(loss.detach())
        return fn, "1 slice (1,%d,%d,3) fwd+bwd+SGD, torch-CPU fp32 oracle, scaled by pixel count to 512x512" % (size, size), \
            size * size / float(512 * 512)
    if a.config == "c3":
        x, y = host_batch(ca, 0)

        def fn():
            ctx, feat = orc.forward_3d(w, x, training=True, requires_grad=True)
            loss = (feat * feat).mean()                    # the head (one 3x3x3 + one 1x1x1 conv of ~60) is < 12 % of the FLOPs
            orc.grads_of(ctx, loss)
            return float(loss.detach())
        return fn, "1 sub-volume (1,%d,%d,%d,4) 3-D DenseNet fwd+bwd, torch-CPU fp32 oracle, scaled by voxel count to 224x224x12" % (
            size, size, cols), size * size * cols / float(224 * 224 * 12)
    vol, lab = synthetic_slab(1, size, cols)
    if a.config == "c5":
        def fn():
            with torch.no_grad():
                _, logits = orc.forward_hybrid(w, vol, training=False, variant="end2end")
            return float(logits.mean())
        nwin = 253.0
        return fn, "1 window (1,%d,%d,%d,1) forward, torch-CPU fp32 oracle, scaled by voxel count to 512x512x8 and by the 253 " \
                   "windows of a 512-slice volume" % (size, size, cols), size * size * cols / float(512 * 512 * 8) / nwin

    def fn():
        ctx, logits = orc.forward_hybrid(w, vol, training=True, variant="end2end", requires_grad=True)
        loss = orc.weighted_crossentropy(torch.as_tensor(lab[..., 0]), logits, crop=True)
        sgd(orc.grads_of(ctx, loss))
        return float(loss.detach())

    return fn, "1 slab (1,%d,%d,%d,1) fwd+bwd+SGD, torch-CPU fp32 oracle, scaled by voxel count to 512x512x48" % (
        size, size, cols), size * size * cols / float(FULL_VOXELS)


def cpu_measure(a, threads, reps, budget_s=None):
    """One untimed warm-up call, then `reps` timed calls (fewer if `budget_s` runs out, never fewer than 1): median."""
    fn, sample, units = cpu_step_fn(a, threads)
    out = fn()
    ts = []
    t_all = time.perf_counter()
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_all > budget_s:
            break
    dt = float(np.median(ts))
    return units / dt, dt, sample, len(ts), out


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, a.cpu_threads)
    v, dt, sample, n, out = cpu_measure(a, threads, max(a.steps, 1))
    cfg = CONFIGS[a.config]
    print(json.dumps({
        "impl": "reference", "metric": cfg["metric"], "value": v, "unit": cfg["unit"], "n_gpus": a.gpus, "steps": n,
        "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[a.config] % dict(batch=a.batch, size=a.size, cols=a.cols or 0), "baseline_config": a.config,
                   "shape": [a.size, a.size] + ([a.cols] if a.cols else []), "batch_per_gpu": a.batch, "precision": "fp32 (CPU port of the reference graph)",
                   "sample": "each step = one %dx%dx%s sample of that workload on the host cores, scaled by voxel count" % (a.cpu_size, a.cpu_size, a.cpu_cols),
                   "statistic": "median of %d" % n},
        "cpu_baseline": {"value": v, "unit": cfg["unit"], "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": out}))


# --------------------------------------------------------------------------------- GPU arm
def timed(fn, steps, dist, dev):
    """barrier + synchronize on both sides, CUDA events on the launch stream, max over ranks."""
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    s.record()
    out = None
    for _ in range(steps):
        out = fn()
    e.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = s.elapsed_time(e)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps, out


def profile_step(model, net, steps):
    """Per-kernel-class device time of `steps` steps (CUDA events around every C-ABI call)."""
    be = net.be
    be.prof = []
    for _ in range(steps):
        model.train_step_device(net)
    torch.cuda.synchronize()
    agg, ops = {}, {}
    for ent in be.prof:
        key, flops, s, e, name = ent[:5]
        nbytes = ent[5] if len(ent) > 5 else 0.0
        t = s.elapsed_time(e)
        a = agg.setdefault(key, [0.0, 0.0, 0, 0.0])
        a[0] += t
        a[1] += flops
        a[2] += 1
        a[3] += nbytes
        if flops:
            o = ops.setdefault((key, name), [0.0, 0.0, 0])
            o[0] += t
            o[1] += flops
            o[2] += 1
    # device idle time between consecutive C-ABI calls (end event of call i -> start event of call i+1): tells a
    # launch-bound step (host cannot keep the queue full) from a kernel-bound one.  Diagnostic only: never fatal.
    gaps = {"total_ms": None}
    try:
        total, worst = 0.0, {}
        prof = be.prof
        for i in range(len(prof) - 1):
            g = prof[i][3].elapsed_time(prof[i + 1][2])
            if g > 0:
                total += g
                w = worst.setdefault(prof[i + 1][0], [0.0, 0])
                w[0] += g
                w[1] += 1
        gaps = {"total_ms": round(total / max(steps, 1), 3), "calls": len(prof) // max(steps, 1),
                "before": {k: {"ms": round(v[0] / max(steps, 1), 3), "n": v[1] // max(steps, 1)}
                           for k, v in sorted(worst.items(), key=lambda kv: -kv[1][0])[:6]}}
    except Exception as ex:                                   # noqa: BLE001
        gaps = {"total_ms": None, "error": str(ex)[:120]}
    be.prof = None
    return agg, ops, gaps


WORKLOAD_TEXT = {
    "c2": "2-D DenseUNet-161 train step (densenet.DenseUNet, train_2ddense.py): %(batch)d slice(s)/GPU of %(size)dx%(size)d, training-mode "
          "BN, fwd + weighted CE + bwd + Nesterov SGD",
    "c3": "3-D DenseNet + hybrid head (hybridnet.DenseNet3D): %(batch)d sub-volume(s)/GPU of %(size)dx%(size)dx%(cols)d x 4 channels, "
          "fwd + weighted CE + bwd + Nesterov SGD",
    "c4": "H-DenseUNet end2end train step (hybridnet.dense_rnn_net): %(batch)d slab(s)/GPU of %(size)dx%(size)dx%(cols)d, "
          "fwd + weighted CE + bwd + Nesterov SGD",
}


def load_peaks():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            return json.load(f)
    return {}


def load_traffic():
    """ncu DRAM bytes per launch of the dominant kernels (profiles/r02_traffic.json, written by scripts/ncu_digest.py
    from `ncu --set full` captures of this round's kernels at the headline shape)."""
    tj = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tj):
        with open(tj) as f:
            return json.load(f)
    return {}


def dist_setup(a):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    if world != a.gpus:
        raise RuntimeError("--gpus %d but WORLD_SIZE is %d (launch with torch.distributed.run)" % (a.gpus, world))
    return world, rank, local, dev, dist


def run_gpu(a):
    import h_denseunet_b200 as hdn
    world, rank, local, dev, dist = dist_setup(a)
    peaks = load_peaks()
    cfg = CONFIGS[a.config]

    m = build_model(a, a.precision, seed=0)
    m.dropout = not a.no_dropout
    if world > 1:
        hdn.make_parallel(m, world, mini_batch=a.batch)
    m.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[loss_of(a)])
    x, y = host_batch(a, rank)
    net = m._net(True)
    ntc = sum(1 for _, p in net.report if any(p))

    # device-resident arm: upload once, then time train_step_device
    m.train_on_batch(x, y)
    for _ in range(max(a.warmup - 1, 0)):
        m.train_step_device(net)
    l0, k0 = net.be.launches, net.be.lib.hdn_launch_count()
    with ClockSampler(local) as clk:
        ms, loss = timed(lambda: m.train_step_device(net), a.steps, dist, dev)
    calls = (net.be.launches - l0) // a.steps               # C-ABI calls per step
    launches = net.be.lib.hdn_launch_count() - k0           # kernels libhdn launched inside the timed region (all steps)
    clocks = clk.summary()

    # end-to-end arm: the call the reference scripts make (train_hybrid.py:213, train_2ddense.py:209): fit_generator over a
    # generator of HOST numpy batches.  Every step copies its batch host -> pinned -> device (on the copy stream, while the
    # previous step computes) and reads the step's loss back; both inside the timed region.
    def gen():
        while True:
            yield x, y

    g = gen()
    m.fit_generator(g, steps_per_epoch=2, epochs=1, verbose=0)                    # staging buffers / copy stream warm-up
    ms_e2e, hist = timed(lambda: m.fit_generator(g, steps_per_epoch=a.steps, epochs=1, verbose=0), 1, dist, dev)
    ms_e2e /= a.steps
    h2d = int(m.h2d_bytes)          # bytes staged host -> device per step (fp32 volume + int16 label map)

    units = units_per_step(a, world)
    value = units / (ms * 1e-3)
    e2e = units / (ms_e2e * 1e-3)
    desc = dict(batch=a.batch, size=a.size, cols=a.cols or 0)
    line = {
        "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if a.precision in ("bf16", "bf16x3", "mixed") and ntc else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[a.config] % desc + (" + P2P grad reduce" if world > 1 else ""), "baseline_config": a.config,
                   "shape": [a.size, a.size] + ([a.cols] if a.cols else []), "batch_per_gpu": a.batch, "parallelism": "dp%d" % world,
                   "precision": a.precision, "tc_convs": ntc, "convs": len(net.report), "dropout": m.dropout,
                   "l2": "activations per step (%.1f GiB) exceed the 126 MB L2; no flush needed" % (
                       net.memory_bytes() / 2 ** 30),
                   "e2e_call": "Model.fit_generator(host numpy generator): pinned staging + H2D on a copy stream, loss read back "
                               "every step"},
        "e2e": {"value": e2e, "unit": cfg["unit"], "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 16},
        "gpu_launches": int(launches), "gpu_launches_per_step": int(launches) // max(a.steps, 1), "cabi_calls_per_step": int(calls), "clocks": clocks, "loss": loss, "memory_gib": net.memory_bytes() / 2 ** 30,
    }

    # roofline of the dominant kernel class, measured live with CUDA events (one extra instrumented step; every rank
    # takes part because the data-parallel exchange inside the step is collective)
    agg, ops, gaps = profile_step(m, net, a.profile_steps)
    if rank == 0:
        tot = sum(v[0] for v in agg.values())
        # the dominant kernel = the single most expensive convolution launch of the step (class + layer name)
        (k, opname), (t, fl, n) = max(ops.items(), key=lambda kv: kv[1][0] / kv[1][2])
        t, fl = t / n, fl / n                               # per launch
        sustained = peaks.get("bf16_tflops_sustained")
        peak_tf = sustained if sustained else 1400.0
        if "[simt]" in k:
            # fp32 FMA kernel: the bound that applies is the fp32 CUDA-core peak (148 SMs x 128 lanes x 2 x clock)
            peak_tf = 148 * 128 * 2 * (clocks["sm_max_mhz"] or 1965.0) * 1e6 / 1e12
        ach = fl / (t * 1e-3) / 1e12 if t > 0 else 0.0
        traffic, tsrc = None, None
        tr = load_traffic()
        key = "%s %s" % (k, opname)
        if key in tr.get("per_launch", {}):
            ent = tr["per_launch"][key]
            if ent.get("positions"):                          # captured at the headline shape, or scaled by output positions (said so)
                traffic = ent["dram_bytes"] * (units * FULL_VOXELS / world if a.config == "c4" else ent["positions"]) / ent["positions"] \
                    if a.config == "c4" else None
            tsrc = tr.get("_source")
        line["roofline"] = {"kernel": key, "bound": "tensor" if "[tc" in k else "fp32-fma", "achieved": ach,
                            "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None,
                            "traffic": traffic, "traffic_source": tsrc, "ms_per_launch": t,
                            "share_of_step": t * n / tot if tot else None, "launches": n,
                            "flop_per_launch": fl,
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if "[tc" in k and sustained
                            else ("fallback 1.4 PF" if "[tc" in k else "fp32 FMA nominal at max SM clock")}
        hbm_peak = peaks.get("hbm_gbs") or 6400.0           # measured copy bandwidth (MEASURED_PEAKS.json), else nominal fallback
        # per class: tensor rate of the algorithmic flops AND HBM rate of the algorithmic bytes (fp32 activations moved
        # once) -- the 1x1 layers of the dense blocks are bounded by the second
        line["kernel_classes"] = {kk: {"ms": round(v[0] / a.profile_steps, 3), "tflops": round(v[1] / max(v[0], 1e-9) / 1e9, 2),
                                        "launches": v[2] // a.profile_steps,
                                        "hbm_gbs": round(v[3] / max(v[0], 1e-9) / 1e6, 1),
                                        "hbm_frac": round(v[3] / max(v[0], 1e-9) / 1e6 / hbm_peak, 4)}
                                  for kk, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]}
        line["step_tflops"] = round(sum(v[1] for v in agg.values()) / a.profile_steps / (ms * 1e-3) / 1e12, 1)
        line["launch_gaps"] = gaps
        # CPU baseline on this box's host cores, bounded sample of the same workload (median of 5 after a warm-up)
        if not a.no_cpu:
            threads = min(os.cpu_count() or 1, a.cpu_threads)   # oneDNN convolutions of this size stop scaling past ~32 threads
            v, dt, sample, nrep, _ = cpu_measure(a, threads, 5, budget_s=60.0)
            line["cpu_baseline"] = {"value": v, "unit": cfg["unit"], "cores": threads, "kind": "port", "sample": sample,
                                    "s_per_sample_step": dt, "statistic": "median of %d after 1 warm-up" % nrep}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------- config 5: sliding-window inference
def run_c5(a):
    """One synthetic 512x512xZ volume through predict_tumor_inwindow (lib/funcs.py:4-51 / test.py:48-69): full liver
    mask => windows range(0, Z-6, 2) (253 for Z = 512), z-sharded over the ranks, one all-reduce of the accumulators.
    A "step" is the whole volume.  `value` keeps the volume on the device side of the call as far as the API allows
    (the window stack is cut from a host array, as in the reference); e2e == value's call: host volume in, two host
    probability volumes out."""
    import h_denseunet_b200 as hdn
    from h_denseunet_b200 import inference
    from h_denseunet_b200.synthetic import synthetic_slab
    world, rank, local, dev, dist = dist_setup(a)
    peaks = load_peaks()
    cfg = CONFIGS["c5"]
    Z = a.depth
    vol, _ = synthetic_slab(1, a.size, Z, seed=1239)
    imgs = np.ascontiguousarray(vol[0, :, :, :, 0])                       # (H, W, Z), as test.py hands it over
    mini, maxi = np.array([0, 0, 0]), np.array([a.size - 1, a.size - 1, Z - 1])
    args = Args(1, a.size, a.cols)
    m = hdn.dense_rnn_net(args, precision=a.precision, seed=0)
    out = {}
    starts = inference.window_starts(Z, int(mini[2]), int(maxi[2]), a.cols)
    fwd_tflop = 3.165 * (a.size / 512.0) ** 2 * (a.cols / 8.0)               # 3.165 TFLOP per 512x512x8 window
    f2d = 8.86 / 48.0 * (a.size / 512.0) ** 2                                # 2-D part per slice
    for reuse in (False, True):
        st = {}
        for _ in range(max(a.warmup // 3, 1) if a.steps else 0):            # one untimed volume (programs, staging, clocks)
            hdn.predict_tumor_inwindow(m, imgs, 3, mini, maxi, args, reuse_2d=reuse, stats=st)
        from h_denseunet_b200 import _lib as _hl
        k0 = _hl.load().hdn_launch_count()
        t0 = time.perf_counter()
        with ClockSampler(local) as clk:
            ms, res = timed(lambda: hdn.predict_tumor_inwindow(m, imgs, 3, mini, maxi, args, reuse_2d=reuse, stats=st), a.steps, dist, dev)
        wall = (time.perf_counter() - t0) / max(a.steps, 1)
        klaunch = int(_hl.load().hdn_launch_count() - k0)
        nwin = len(starts)
        naive = nwin * fwd_tflop
        dedup = naive - (nwin * a.cols - st["slices_2d"] * world) * f2d if reuse else naive
        out[reuse] = {"s_per_volume": ms * 1e-3, "wall_s_per_volume": wall, "volumes_per_s": 1e3 / ms, "windows": nwin,
                      "windows_this_rank": st["windows"], "slices_2d_this_rank": st["slices_2d"],
                      "tflop_naive": round(naive, 1), "tflop_evaluated": round(dedup, 1),
                      "tflops_naive_rate": round(naive / (ms * 1e-3), 1), "clocks": clk.summary(),
                      "checksum": [float(res[0].sum()), float(res[1].sum())], "gpu_launches": klaunch}
    if rank == 0:
        best = out[True] if out[True]["s_per_volume"] < out[False]["s_per_volume"] else out[False]
        line = {"metric": cfg["metric"], "value": best["volumes_per_s"], "unit": cfg["unit"], "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": best["s_per_volume"] * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16" if a.precision != "fp32" else "f32", "data": "synthetic",
                "config": {"workload": "sliding-window inference of one %dx%dx%d volume (lib/funcs.py:4-51): %d windows of %d slices, "
                                       "stride %d, z-sharded over %d GPU(s)" % (a.size, a.size, Z, len(starts), a.cols, a.cols // 4, world),
                           "baseline_config": "c5", "precision": a.precision, "parallelism": "z%d" % world,
                           "l2": "one window's activations (>3 GiB) exceed the 126 MB L2; no flush needed"},
                "e2e": {"value": best["volumes_per_s"], "unit": cfg["unit"], "ms_per_step": best["s_per_volume"] * 1e3,
                        "h2d_bytes_per_step": int(len(starts) / world * a.size * a.size * a.cols * 4),
                        "d2h_bytes_per_step": int(a.size * a.size * Z * 2 * 4)},
                "without_slice_reuse": out[False], "with_slice_reuse": out[True],
                "gpu_launches": best["gpu_launches"], "clocks": best["clocks"]}
        sustained = peaks.get("bf16_tflops_sustained") or 1400.0
        line["roofline"] = {"kernel": "whole window forward (231 convolutions)", "bound": "tensor", "achieved": best["tflops_naive_rate"] / world,
                            "peak": sustained, "unit": "TFLOP/s", "frac": best["tflops_naive_rate"] / world / sustained, "traffic": None,
                            "note": "algorithmic (naive, direct-convolution) TFLOP of all windows / wall time, per GPU"}
        if not a.no_cpu:
            threads = min(os.cpu_count() or 1, a.cpu_threads)
            v, dt, sample, nrep, _ = cpu_measure(a, threads, 5, budget_s=60.0)
            line["cpu_baseline"] = {"value": v * (512.0 / Z), "unit": cfg["unit"], "cores": threads, "kind": "port", "sample": sample,
                                    "s_per_sample_step": dt, "statistic": "median of %d after 1 warm-up" % nrep}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default="c4", choices=["c2", "c3", "c4", "c5"], help="BASELINE.json configuration (default: the headline c4)")
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--cols", type=int, default=None)
    ap.add_argument("--depth", type=int, default=512, help="c5: slices of the volume")
    ap.add_argument("--batch", type=int, default=None, help="work units per GPU (c4 north star: 2 slabs; c2: 8 slices)")
    ap.add_argument("--precision", default="mixed", choices=["bf16", "bf16x3", "mixed", "fp32"],
                    help="bf16: tcgen05, operands rounded to bf16; bf16x3: tcgen05, operands split into bf16 head + tail "
                         "(3 MMAs per step, fp32-grade results); mixed: fprop + dgrad bf16x3, wgrad bf16; fp32: FMA parity path")
    ap.add_argument("--no-dropout", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=1)
    ap.add_argument("--cpu-size", type=int, default=None)
    ap.add_argument("--cpu-cols", type=int, default=None)
    ap.add_argument("--cpu-threads", type=int, default=32, help="host threads of the CPU arm (capped at the core count)")
    a = defaults(ap.parse_args())
    if a.impl == "reference":
        run_reference(a)
    elif a.config == "c5":
        run_c5(a)
    else:
        run_gpu(a)


if __name__ == "__main__":
    main()
