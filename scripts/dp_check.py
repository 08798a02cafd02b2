"""Multi-GPU check (run under torch.distributed.run, one rank per GPU): the fused NVLink P2P
reduce + Nesterov SGD + parameter push (hdn_dp_reduce_sgd) against the NCCL all-reduce baseline path, and their
timing on the full gradient arena."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import h_denseunet_b200 as hdn
    from h_denseunet_b200.parallel import DataParallel
    from h_denseunet_b200.synthetic import synthetic_slab
    from util import Args, perturb_params

    res = {}
    for impl in ("p2p", "coll"):
        m = hdn.dense_rnn_net(Args(1, 64, 8), precision="fp32", seed=0)
        m.dropout = False
        perturb_params(m)
        m.compile(optimizer=hdn.SGD(lr=1e-3, momentum=0.9, nesterov=True), loss=[hdn.weighted_crossentropy])
        m.dp = DataParallel(impl=impl)
        vol, lab = synthetic_slab(1, 64, 8, seed=1234 + rank)
        losses = [m.train_on_batch(vol, lab) for _ in range(3)]
        w = m.params.train.clone()
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        same = all(torch.equal(ws[0], x) for x in ws)
        res[impl] = (w, losses, same)
        # time the exchange alone on the real arena (61 M parameters)
        net = m.nets[True]
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            m.dp.step(net, 0.0, 0.9)
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / 5], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            n = m.params.n_train
            print("impl %-4s: replicas identical=%s  losses=%s  exchange+SGD %.3f ms for %d params (%.1f GB/s algorithmic grad bytes per GPU)" % (
                impl, same, ["%.5f" % l for l in losses], float(t), n, n * 4 * (world - 1) / world / (float(t) * 1e-3) / 1e9))
    d = float((res["p2p"][0] - res["coll"][0]).abs().max())
    rel = d / float(res["coll"][0].abs().max())
    if rank == 0:
        print("p2p vs NCCL baseline: max |dw| = %.3e (rel %.3e)" % (d, rel))
    assert res["p2p"][2] and res["coll"][2], "replicas diverged"
    assert rel < 1e-4      # three steps of fp32 training with atomics in a different order
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
