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
bounded sample: one (1, 160, 160, 8) slab (the reference trains on 224x224x8, train_hybrid.py:28-31), scaled to
512x512x48-slab units by voxel count.
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "CT slabs/sec (512x512x48) fwd+bwd"
FULL_VOXELS = 512 * 512 * 48


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


# --------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(size, cols, threads):
    """Returns (fn, sample description): fn() runs one fwd+loss+bwd+SGD of the oracle on one slab."""
    from oracle import hdense_oracle as orc
    import h_denseunet_b200 as hdn
    from h_denseunet_b200.synthetic import synthetic_slab
    torch.set_num_threads(threads)
    m = hdn.dense_rnn_net(Args(1, size, cols))          # parameter dictionary only; no device is touched
    w = m.get_weights_dict()
    vol, lab = synthetic_slab(1, size, cols)
    mom = {}

    def fn():
        ctx, logits = orc.forward_hybrid(w, vol, training=True, variant="end2end", requires_grad=True)
        loss = orc.weighted_crossentropy(torch.as_tensor(lab[..., 0]), logits, crop=True)
        g = orc.grads_of(ctx, loss)
        for k, gk in g.items():
            if gk is None:
                continue
            p1, v = orc.sgd_nesterov_step(w[k], gk, mom.get(k, 0.0))
            w[k], mom[k] = p1.astype(np.float32), v
        return float(loss.detach())

    return fn, "1 slab (1,%d,%d,%d,1) fwd+bwd+SGD, torch-CPU fp32 oracle, scaled by voxel count to 512x512x48" % (
        size, size, cols)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, a.cpu_threads)
    size, cols = a.cpu_size, a.cpu_cols
    fn, sample = cpu_step_fn(size, cols, threads)
    warm, _ = cpu_step_fn(32, 8, threads)                   # untimed warm-up on a tiny slab
    for _ in range(max(min(a.warmup, 1), 0)):
        warm()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        loss = fn()
    dt = (time.perf_counter() - t0) / a.steps
    slabs = (size * size * cols) / float(FULL_VOXELS)
    v = slabs / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "slabs/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "hybrid end2end train step (hybridnet.dense_rnn_net), CPU sample %dx%dx%d" % (size, size, cols)},
        "cpu_baseline": {"value": v, "unit": "slabs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "slabs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": loss}))


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


def run_gpu(a):
    import h_denseunet_b200 as hdn
    from h_denseunet_b200.synthetic import synthetic_slab
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

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            peaks = json.load(f)

    m = hdn.dense_rnn_net(Args(a.batch, a.size, a.cols), precision=a.precision, seed=0)
    m.dropout = not a.no_dropout
    if world > 1:
        hdn.make_parallel(m, world, mini_batch=a.batch)
    m.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy])
    vol, lab = synthetic_slab(a.batch, a.size, a.cols, seed=1234 + rank)
    net = m._net(True)
    ntc = sum(1 for _, p in net.report if any(p))

    # device-resident arm: upload once, then time train_step_device
    m.train_on_batch(vol, lab)
    for _ in range(max(a.warmup - 1, 0)):
        m.train_step_device(net)
    l0 = net.be.launches
    with ClockSampler(local) as clk:
        ms, loss = timed(lambda: m.train_step_device(net), a.steps, dist, dev)
    launches = (net.be.launches - l0) // a.steps
    clocks = clk.summary()
    # end-to-end arm: host numpy in, loss out, every step
    m.train_on_batch(vol, lab)
    ms_e2e, loss_e2e = timed(lambda: m.train_on_batch(vol, lab), a.steps, dist, dev)
    h2d = int(m.h2d_bytes)          # bytes Model.train_on_batch copied host -> device this step (fp32 volume + int16 label map)

    slabs_step = a.batch * world * (a.size * a.size * a.cols) / float(FULL_VOXELS)
    value = slabs_step / (ms * 1e-3)
    e2e = slabs_step / (ms_e2e * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": "slabs/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if a.precision in ("bf16", "bf16x3", "mixed") and ntc else "f32", "data": "synthetic",
        "config": {"workload": "H-DenseUNet end2end train step (hybridnet.dense_rnn_net): %d slab(s)/GPU of %dx%dx%d, "
                               "fwd + weighted CE + bwd + Nesterov SGD%s" % (
                                   a.batch, a.size, a.size, a.cols, " + P2P grad reduce" if world > 1 else ""),
                   "slab": [a.size, a.size, a.cols], "batch_per_gpu": a.batch, "parallelism": "dp%d" % world,
                   "precision": a.precision, "tc_convs": ntc, "convs": len(net.report), "dropout": m.dropout,
                   "l2": "activations per step (%.1f GiB) exceed the 126 MB L2; no flush needed" % (
                       net.memory_bytes() / 2 ** 30)},
        "e2e": {"value": e2e, "unit": "slabs/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 16},
        "gpu_launches": int(launches), "clocks": clocks, "loss": loss, "memory_gib": net.memory_bytes() / 2 ** 30,
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
        tj = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tj):
            with open(tj) as f:
                tr = json.load(f)
            what = k.split("[")[0].replace("conv_", "")
            if opname in tr and what in tr[opname]:
                traffic = tr[opname][what] * a.cols * a.batch
                tsrc = tr["_source"]
        line["roofline"] = {"kernel": "%s %s" % (k, opname), "bound": "tensor" if "[tc" in k else "fp32-fma", "achieved": ach,
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
                                  for kk, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]}
        line["launch_gaps"] = gaps
        # CPU baseline on this box's host cores, bounded sample
        if not a.no_cpu:
            threads = min(os.cpu_count() or 1, a.cpu_threads)   # oneDNN convolutions of this size stop scaling past ~32 threads
            fn, sample = cpu_step_fn(a.cpu_size, a.cpu_cols, threads)
            warm, _ = cpu_step_fn(32, 8, threads)           # thread pools / allocator warm-up on a tiny slab, untimed
            warm()
            t0 = time.perf_counter()
            n_rep = 0
            while n_rep < 2 and (time.perf_counter() - t0) < 12.0:      # bounded: ~10-30 s of CPU work
                fn()
                n_rep += 1
            dt = (time.perf_counter() - t0) / n_rep
            v = (a.cpu_size * a.cpu_size * a.cpu_cols) / float(FULL_VOXELS) / dt
            line["cpu_baseline"] = {"value": v, "unit": "slabs/s", "cores": threads, "kind": "port", "sample": sample,
                                    "s_per_sample_step": dt}
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
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--cols", type=int, default=48)
    ap.add_argument("--batch", type=int, default=2, help="slabs per GPU (north star: 2)")
    ap.add_argument("--precision", default="mixed", choices=["bf16", "bf16x3", "mixed", "fp32"],
                    help="bf16: tcgen05, operands rounded to bf16; bf16x3: tcgen05, operands split into bf16 head + tail "
                         "(3 MMAs per step, fp32-grade results); mixed: fprop + dgrad bf16x3, wgrad bf16; fp32: FMA parity path")
    ap.add_argument("--no-dropout", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=1)
    ap.add_argument("--cpu-size", type=int, default=160)
    ap.add_argument("--cpu-cols", type=int, default=8)
    ap.add_argument("--cpu-threads", type=int, default=32, help="host threads of the CPU arm (capped at the core count)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)


if __name__ == "__main__":
    main()
