"""Data-parallel plumbing (one process per GPU).

The path shards by SAMPLE: every rank runs the full step on its own batch shard; the only exchange is the gradient
all-reduce before the optimiser (DDP semantics, base_model.py:725-737: mean over ranks, bucketed and overlapped with
the backward pass).  Here:

  * the gradient is ONE flat fp32 buffer cut into a few contiguous BUCKETS in reverse registration order (the order
    in which the backward pass finishes them); `GradBuckets` tracks which parameters of a bucket have their gradient
    and hands a finished bucket to the communicator while the backward pass is still running;
  * `Comm` on a GPU is the library's own NCCL communicator (csrc/comm.cu, jg_comm_*): collectives run on its private
    stream, forked from / joined to the compute stream by events, which also works inside a CUDA-graph capture.  The
    128-byte NCCL id travels through the torch.distributed process group the launcher (torchrun) set up — that group
    is used for the rendezvous only.  On the CPU (the gloo tests of the host logic) `Comm` falls back to
    torch.distributed collectives;
  * the SUM is turned into DDP's mean by the 1/world factor folded into the fused optimizer kernel.
"""
import ctypes

import torch
import torch.distributed as dist


def world_size(pg=None):
    return dist.get_world_size(pg) if (dist.is_available() and dist.is_initialized()) else 1


def allreduce_sum_(flat_grad: torch.Tensor, pg=None):
    """In-place SUM all-reduce of the flat gradient buffer through torch.distributed (blocking on the current
    stream); returns the scale (1/world) to apply.  The trainers use `Comm` instead; this stays for the GAN trainers'
    small per-network buffers and for the CPU tests."""
    w = world_size(pg)
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=pg)
    return 1.0 / w


def broadcast_(flat_params: torch.Tensor, pg=None, src=0):
    if world_size(pg) > 1:
        dist.broadcast(flat_params, src=src, group=pg)


def shard_seed(base_seed: int, rank: int) -> int:
    """Per-rank synthetic-data seed (seeds 1234 + rank)."""
    return base_seed + rank


def plan_buckets(offsets, sizes, total, n_buckets=8, min_elems=1 << 20):
    """Cut the flat buffer [0, total) into contiguous buckets, LAST parameters first (the backward pass produces
    gradients in reverse registration order).  offsets / sizes: per-parameter slices in registration order.
    Returns a list of (lo, hi, [param indices]) covering [0, total) exactly, in launch order."""
    n = len(offsets)
    if n == 0:
        return []
    target = max(min_elems, (total + n_buckets - 1) // n_buckets)
    buckets, hi, members = [], total, []
    for i in range(n - 1, -1, -1):
        members.append(i)
        if hi - offsets[i] >= target and i > 0:
            buckets.append((offsets[i], hi, members))
            hi, members = offsets[i], []
    buckets.append((0, hi, members))
    return buckets


class Comm:
    """SUM all-reduce / broadcast of slices of flat buffers for one process group."""

    def __init__(self, pg=None, device=None):
        self.pg = pg
        self.world = world_size(pg)
        self.rank = dist.get_rank(pg) if self.world > 1 else 0
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.handle = None
        self._pending = []
        if self.world > 1 and self.device.type == "cuda":
            from . import lib as L
            lib = L.load()
            ident = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                L._check(lib.jg_comm_unique_id(ident.data_ptr()), "jg_comm_unique_id")
            ident = ident.to(self.device)
            src = dist.get_global_rank(pg, 0) if pg is not None else 0
            dist.broadcast(ident, src=src, group=pg)
            ident = ident.cpu()
            h = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                L._check(lib.jg_comm_init(ident.data_ptr(), self.rank, self.world, ctypes.byref(h)), "jg_comm_init")
            self.handle = h
            self._lib = lib

    # -- collectives ---------------------------------------------------------------------------------------------
    def allreduce_async(self, t: torch.Tensor):
        """In-place SUM of `t` (a contiguous fp32 / bf16 slice) that overlaps whatever the caller launches next."""
        if self.world == 1:
            return
        if self.handle is not None:
            from . import lib as L
            dtype = {torch.float32: 0, torch.bfloat16: 1}[t.dtype]
            L._check(self._lib.jg_comm_allreduce_async(self.handle, t.data_ptr(), t.numel(), dtype, L.stream()),
                     "jg_comm_allreduce_async")
        else:
            self._pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def wait(self):
        """The current stream (GPU) / the caller (CPU) waits for every collective issued so far."""
        if self.world == 1:
            return
        if self.handle is not None:
            from . import lib as L
            L._check(self._lib.jg_comm_wait(self.handle, L.stream()), "jg_comm_wait")
        else:
            for w in self._pending:
                w.wait()
            self._pending = []

    def broadcast(self, t: torch.Tensor, root=0):
        if self.world == 1:
            return
        if self.handle is not None:
            from . import lib as L
            L._check(self._lib.jg_comm_broadcast(self.handle, t.data_ptr(), t.numel() * t.element_size(), root,
                                                 L.stream()), "jg_comm_broadcast")
        else:
            src = dist.get_global_rank(self.pg, root) if self.pg is not None else root
            dist.broadcast(t, src=src, group=self.pg)

    def stats(self):
        """{"collectives": n, "bytes": b, "nccl": version} since creation (library communicator only)."""
        if self.handle is None:
            return None
        c, b, v = ctypes.c_ulonglong(), ctypes.c_ulonglong(), ctypes.c_int()
        self._lib.jg_comm_info(self.handle, None, None, ctypes.byref(c), ctypes.byref(b), ctypes.byref(v))
        return {"collectives": c.value, "bytes": b.value, "nccl": v.value}

    def close(self):
        if self.handle is not None:
            self._lib.jg_comm_destroy(self.handle)
            self.handle = None


class GradBuckets:
    """Readiness bookkeeping for the overlapped exchange.  `ready(i)` is called once per parameter when its gradient
    is final (autograd's post-accumulate hook, or the wgrad kernel's launch for trainer-staged convolution weights);
    a bucket whose parameters are all ready is passed to `on_bucket(bucket_index)` at once.  `finish()` flushes the
    buckets that never completed (parameters without a gradient in this pass), in launch order."""

    def __init__(self, offsets, sizes, total, on_bucket, n_buckets=8, min_elems=1 << 20):
        self.buckets = plan_buckets(offsets, sizes, total, n_buckets, min_elems)
        self.bucket_of = {}
        for b, (_, _, members) in enumerate(self.buckets):
            for i in members:
                self.bucket_of[i] = b
        self.on_bucket = on_bucket
        self.active = False
        self.reset()

    def reset(self):
        self.missing = [set(m) for (_, _, m) in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.late = []

    def begin(self):
        self.reset()
        self.active = True

    def ready(self, i):
        if not self.active:
            return
        b = self.bucket_of[i]
        if self.launched[b]:
            self.late.append(i)  # a second gradient contribution after the bucket left: the caller must not overlap
            return
        self.missing[b].discard(i)
        if not self.missing[b]:
            self.launched[b] = True
            self.on_bucket(b)

    def finish(self):
        self.active = False
        if self.late:
            raise RuntimeError("GradBuckets: parameters %s received a gradient after their bucket was reduced "
                               "(shared parameters?): construct the trainer with overlap_comm=False" % self.late[:4])
        for b in range(len(self.buckets)):
            if not self.launched[b]:
                self.launched[b] = True
                self.on_bucket(b)
