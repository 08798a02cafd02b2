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
            Hand-over between the processes is on the compute stream (sync='flags', default): a row of step flags in
            every peer-mapped arena -- hdn_dp_signal after the backward pass ("my gradients are complete"),
            hdn_dp_wait in front of the exchange, hdn_dp_signal after it ("my shard of everyone's parameters is
            written, I no longer read your gradients"), hdn_dp_wait in front of the next step's first kernel.  No host
            synchronisation, no NCCL call inside a step.  sync='host' (HDN_DP_SYNC=host) keeps the round-1 form
            (cuda synchronize + dist.barrier on both sides); it is what two processes sharing ONE device must use
            (spinning kernels of two contexts time-slice on one GPU).
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
    NFLAG = 64                 # uint32 slots per flag row (two rows: "gradients ready", "exchange done")

    def __init__(self, impl=None, sync=None):
        self.impl = impl or os.environ.get("HDN_DP_IMPL", "p2p")
        self.sync = sync or os.environ.get("HDN_DP_SYNC", "flags")
        self.step_no = 0
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
                _lib.check(lib.hdn_dev_malloc(C.byref(p), (total + 2 * self.NFLAG) * 4), "hdn_dev_malloc")
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
        # flag rows behind the two arenas: [ready: NFLAG][done: NFLAG], zeroed before anyone can signal
        self.flags = torch.as_tensor(_Raw(self.base + 2 * self.half * 4, 2 * self.NFLAG), device=dev).view(torch.int32)
        self.flags.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        self.f_ready = (C.c_void_p * self.world)(*[b + 2 * self.half * 4 for b in self.peer_base])
        self.f_done = (C.c_void_p * self.world)(*[b + 2 * self.half * 4 + self.NFLAG * 4 for b in self.peer_base])
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
        st = torch.cuda.current_stream().cuda_stream
        self.step_no += 1
        if self.sync == "flags":
            lib, me = self.lib, self.base + 2 * self.half * 4
            _lib.check(lib.hdn_dp_signal(self.f_ready, self.world, self.rank, self.step_no, st), "hdn_dp_signal")
            _lib.check(lib.hdn_dp_wait(me, self.world, self.step_no, st), "hdn_dp_wait")       # every replica's gradients are complete
        else:
            torch.cuda.synchronize()
            dist.barrier()
        _lib.check(self.lib.hdn_dp_reduce_sgd(self.pp, self.pg, ps.moms.data_ptr(), self.world, self.rank, lo, hi,
                                              lr, mu, gs, st), "hdn_dp_reduce_sgd")
        net.be.launches += 1
        if self.sync == "flags":
            _lib.check(self.lib.hdn_dp_signal(self.f_done, self.world, self.rank, self.step_no, st), "hdn_dp_signal")
            self.pending = self.step_no    # begin_step() waits for it in front of the next step's first kernel
            net.be.launches += 3
        else:
            torch.cuda.synchronize()
            dist.barrier()                 # every replica holds the updated parameters

    def begin_step(self):
        """In front of a step's first kernel: every peer has pushed its shard of the new parameters into this arena and
        has finished reading this replica's gradients (flags mode; a no-op otherwise)."""
        if self.impl == "p2p" and self.sync == "flags" and getattr(self, "pending", 0):
            _lib.check(self.lib.hdn_dp_wait(self.base + (2 * self.half + self.NFLAG) * 4, self.world, self.pending,
                                            torch.cuda.current_stream().cuda_stream), "hdn_dp_wait")
            self.pending = 0
