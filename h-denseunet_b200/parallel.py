"""Data parallelism over the GPUs of one box: one process per GPU (torch.distributed for the
rendezvous), replicas fed disjoint samples, ONE exchange per step.

Replaces Keras-2.0.8/keras/utils2/multi_gpu.py:7-69 (towers under tf.device + CPU concat, the
gradient sum left to TF's implicit cross-device copies).  Semantics kept: the loss is a mean
over the merged batch (multi_gpu.py:65-69 then training.py:849), so per-replica mean-loss
gradients are AVERAGED; BN batch statistics stay per replica (multi_gpu.py:35-53).

impl 'p2p'  (default on CUDA): parameter and gradient arenas are cudaMalloc'ed, exported with CUDA
            IPC and mapped by every peer.  hdn_dp_reduce_sgd: rank r pulls its 1/world shard of
            all peers' gradients over NVLink, applies the Nesterov update and pushes the new
            parameters into every peer's arena -- reduce-scatter + SGD + all-gather in one kernel.
impl 'coll' (baseline, and the gloo/CPU test path): dist.all_reduce on the gradient arena, then
            the local hdn_sgd_nesterov.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib


class _Raw(object):
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def shard_bounds(n, world, rank):
    """[lo, hi) of rank's shard of an n-element arena; every lo is a multiple of 4 (float4 path)."""
    chunk = ((n + world - 1) // world + 3) // 4 * 4
    lo = min(n, rank * chunk)
    hi = min(n, lo + chunk)
    return lo, hi


class DataParallel(object):
    def __init__(self, impl=None):
        self.impl = impl or os.environ.get("HDN_DP_IMPL", "p2p")
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.ready = False

    def realise(self, params, device):
        if self.ready:
            return
        dev = torch.device(device)
        if self.impl == "p2p" and dev.type != "cuda":
            self.impl = "coll"
        if self.impl == "coll":
            params.realise(dev)
            self._sync_initial(params)
            self.ready = True
            return
        lib = _lib.load()
        holder = {}

        def alloc(n):
            n = max(n, 4)
            if "base" not in holder:
                total = 2 * ((n + 63) // 64 * 64)
                p = C.c_void_p()
                _lib.check(lib.hdn_dev_malloc(C.byref(p), total * 4), "hdn_dev_malloc")
                holder["base"], holder["half"], holder["k"] = p.value, total // 2, 0
            ptr = holder["base"] + holder["k"] * holder["half"] * 4
            holder["k"] += 1
            return torch.as_tensor(_Raw(ptr, n), device=dev)

        params.realise(dev, alloc=alloc)
        self.base, self.half = holder["base"], holder["half"]
        h = (C.c_ubyte * 64)()
        _lib.check(lib.hdn_ipc_get_handle(self.base, h), "hdn_ipc_get_handle")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(h))
        self.peer_base = []
        for r in range(self.world):
            if r == self.rank:
                self.peer_base.append(self.base)
            else:
                q = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                _lib.check(lib.hdn_ipc_open(hb, C.byref(q)), "hdn_ipc_open")
                self.peer_base.append(q.value)
        self.pp = (C.c_void_p * self.world)(*self.peer_base)
        self.pg = (C.c_void_p * self.world)(*[b + self.half * 4 for b in self.peer_base])
        self.lib = lib
        self._sync_initial(params)
        self.ready = True

    def _sync_initial(self, params):
        # replicas start identical: rank 0's values win (the reference shares variables between towers)
        dist.broadcast(params.train, 0)
        dist.broadcast(params.state, 0)

    def step(self, net, lr, mu):
        ps = net.params
        gs = 1.0 / self.world
        if self.impl == "coll":
            dist.all_reduce(ps.grads)
            net.be.sgd(ps.train, ps.grads, ps.moms, ps.n_train, lr, mu, gs)
            return
        lo, hi = shard_bounds(ps.n_train, self.world, self.rank)
        torch.cuda.synchronize()
        dist.barrier()                     # every replica's gradients are complete
        _lib.check(self.lib.hdn_dp_reduce_sgd(self.pp, self.pg, ps.moms.data_ptr(), self.world, self.rank, lo, hi,
                                              lr, mu, gs, torch.cuda.current_stream().cuda_stream),
                   "hdn_dp_reduce_sgd")
        net.be.launches += 1
        torch.cuda.synchronize()
        dist.barrier()                     # every replica holds the updated parameters
