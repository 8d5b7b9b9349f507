"""Palette training inner loop on the B200 kernels.

Mirrors the part of joliGEN that runs every iteration (train.py:268-282):

    model.set_input(data)            palette_model.py:287-556   (device placement of A / B / mask)
    model.optimize_parameters()      base_model.py:1302-1377    (group loop: compute_palette_loss ->
                                     backward -> compute_step (optimizer) -> ema_step)

with the same attribute names (`netG_A`, `loss_G_tot`, `optimizer_G` semantics) but:
  * parameters, gradients, Adam moments and the EMA copy live in FLAT fp32 buffers (the nn.Parameters
    are views), so the gradient all-reduce is one NCCL call and Adam(W)+EMA is one kernel launch;
  * the eps-loss consumes the UNet's NHWC bf16 output directly.
Data parallelism = one process per GPU, batch sharded by the caller, SUM all-reduce of the flat
gradient scaled by 1/world inside the optimizer kernel (DDP's mean, base_model.py:725-737).
"""
import torch

from . import dp
from . import kernels as K
from . import lib as L
from . import nets


class FlatParams:
    """Re-point every parameter of `module` at a slice of one flat fp32 buffer (same for .grad)."""

    def __init__(self, module, late=None):
        """late(name) -> True for parameters whose gradient only becomes final at the very END of the backward pass
        (the timestep-embedding Linears of every ResBlock are evaluated by one batched launch, the label tables and the
        gamma MLP sit in front of the whole net): they are placed FIRST in the flat buffers, i.e. in the LAST gradient
        bucket, so that the other buckets can leave while the backward pass is still running.  state_dict order and
        names are untouched (only .data / .grad are re-pointed)."""
        named = list(module.named_parameters())
        if late is not None:
            named = [(n, p) for n, p in named if late(n)] + [(n, p) for n, p in named if not late(n)]
        self.params = [p for _, p in named]
        self.names = [n for n, _ in named]
        # 64-element (256 B) aligned slices keep every tensor 16-byte aligned for vector / TMA access
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 63) // 64 * 64
        self.total = off
        dev = self.params[0].device
        self.data = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.data[o:o + n].copy_(p.detach().reshape(-1).float())
            p.data = self.data[o:o + n].view(p.shape)
            p.grad = self.grad[o:o + n].view(p.shape)
        nets.invalidate_packed_weights()

    def rebind_grads(self):
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def unflatten(self, flat):
        return {n: flat[o:o + p.numel()].view(p.shape) for n, p, o in zip(self.names, self.params, self.offsets)}


class _WgradSlot:
    __slots__ = ("acc", "layout", "param", "dims", "owner", "index")

    def __init__(self, acc, param, dims, owner):
        self.acc, self.layout, self.param, self.dims, self.owner = acc, None, param, dims, owner
        self.index = None  # position of the parameter in the trainer's flat buffers (set by the trainer)

    @property
    def enabled(self):  # only between WgradStage.begin() and flush(): plain autograd use of the model stays immediate
        return self.owner.active

    def notify(self):
        """called by Conv2dFn.backward right after this slot's wgrad was launched: the gradient is final"""
        cb = self.owner.on_ready
        if cb is not None and self.index is not None:
            cb(self.index)


class WgradStage:
    """Persistent fp32 split-K accumulators for every convolution weight of a model (one flat buffer).  The wgrad
    kernels add their partial tiles into a slot (no per-call memset, no per-call permutation); flush() — ONE launch at
    the end of the backward pass — permutes every slot into the parameter's OIHW .grad (+=) and zeroes it again."""

    def __init__(self, module):
        slots, total = [], 0
        for pack in nets.conv_packs(module):
            w = pack.conv.weight
            w4 = pack._weight4()
            cout, cin, r, s = w4.shape
            if cout % 8 or cin % 8 or not nets.on_device(w):
                continue  # padded channels: Conv2dFn uses its immediate path for these
            slots.append((w, (cout, cin, r * s), total))
            total += (w.numel() + 63) // 64 * 64
        self.buf = torch.zeros(max(total, 1), dtype=torch.float32, device=slots[0][0].device) if slots else None
        self.slots = []
        for w, dims, off in slots:
            slot = _WgradSlot(self.buf[off:off + w.numel()], w, dims, self)
            w._jg_wstage = slot
            self.slots.append(slot)
        self._tables = {}  # subset key -> (table key, WeightTable)
        self.active = False
        self.on_ready = None  # fn(flat parameter index): the trainer's bucket bookkeeping (overlapped all-reduce)

    def begin(self):
        self.active = True

    def end(self):
        self.active = False

    def flush(self, slots=None, tag="all"):
        """Permute-add the accumulators of `slots` (default: all) into their parameters' .grad and zero them: ONE
        launch.  `tag` names the subset for the cached device table (one per gradient bucket)."""
        if slots is None:
            self.active = False
            slots = self.slots
        active = [sl for sl in slots if sl.layout is not None and sl.param.grad is not None]
        if not active:
            return
        key = tuple((id(sl), sl.layout, sl.param.grad.data_ptr()) for sl in active)
        cached = self._tables.get(tag)
        if cached is None or cached[0] != key:
            items = [L.UnpackItem(sl.acc.data_ptr(), sl.param.grad.data_ptr(), sl.dims[0], sl.dims[1], sl.dims[2],
                                  sl.layout) for sl in active]
            cached = (key, K.WeightTable(items, [sl.dims for sl in active], self.buf.device))
            self._tables[tag] = cached
        K.wgrad_unpack_batched(cached[1])


class PaletteTrainer:
    def __init__(self, netG_A, lr=2e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, optim="adamw",
                 ema=True, ema_beta=0.999, iter_size=1, lambda_G=1.0, use_minsnr=False, loss="MSE",
                 device=None, process_group=None, cuda_graph=False, graph_warmup=3, overlap_comm=True,
                 comm_buckets=8, comm_min_bucket=1 << 20, dropout_prob=0.0, num_classes=None):
        if not torch.cuda.is_available():
            raise RuntimeError("joligen_b200.PaletteTrainer needs a CUDA device (there is no CPU path)")
        self.device = torch.device(device if device is not None else "cuda")
        self.netG_A = netG_A.to(self.device)
        self.flat = FlatParams(self.netG_A, late=lambda n: ("emb_layers" in n or n.startswith("cond_embed")
                                                             or "embedder" in n))
        # all conv weights: one batched bf16 re-pack after each optimizer step, one batched wgrad unpack per backward
        self.packset = nets.WeightPackSet(self.netG_A)
        self.wstage = WgradStage(self.netG_A)
        self._loose = [(p, self.flat.grad[o:o + p.numel()].view(p.shape))
                       for p, o in zip(self.flat.params, self.flat.offsets) if not hasattr(p, "_jg_wstage")]
        self.exp_avg = torch.zeros_like(self.flat.data)
        self.exp_avg_sq = torch.zeros_like(self.flat.data)
        self.ema = torch.zeros_like(self.flat.data) if ema else None
        self.ema_started = False
        self.hp = dict(lr=lr, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay, adamw=(optim == "adamw"))
        if optim not in ("adamw", "adam"):
            raise NotImplementedError("optimizer %r (adam / adamw are implemented)" % optim)
        self.ema_beta = ema_beta
        self.iter_size = iter_size
        self.lambda_G = lambda_G
        self.use_minsnr = use_minsnr
        self.l1 = (loss == "L1")
        if loss not in ("MSE", "L1"):
            raise NotImplementedError("alg_palette_loss %r" % loss)
        self.use_ref = getattr(getattr(self.netG_A, "denoise_fn", None), "model_nargs", 2) == 3
        self.ref_A = None
        # class / mask conditioning (alg_diffusion_cond_embed) and its dropout (palette_model.py:565-584: the dropped
        # samples get the highest class = "unconditioned")
        self.conditioning = getattr(getattr(self.netG_A, "denoise_fn", None), "conditioning", "")
        self.dropout_prob = float(dropout_prob)
        self.num_classes = num_classes
        if self.dropout_prob > 0.0 and num_classes is None:
            raise ValueError("PaletteTrainer: dropout_prob > 0 needs num_classes (palette_model.py:148)")
        self.cls = None
        self.pg = process_group
        self.world = dp.world_size(process_group)
        # gradient exchange: the library's own NCCL communicator; buckets of the flat gradient leave as soon as the
        # backward pass has produced them (DDP's overlap, base_model.py:725-737)
        self.comm = dp.Comm(process_group, self.device)
        self.overlap = bool(overlap_comm) and self.world > 1
        self._index = {id(p): i for i, p in enumerate(self.flat.params)}
        for sl in self.wstage.slots:
            sl.index = self._index[id(sl.param)]
        self._sizes = [(p.numel() + 63) // 64 * 64 for p in self.flat.params]
        self.buckets = dp.GradBuckets(self.flat.offsets, self._sizes, self.flat.total, self._reduce_bucket,
                                      n_buckets=comm_buckets, min_elems=comm_min_bucket)
        self._bucket_slots = [[sl for sl in self.wstage.slots if self.buckets.bucket_of[sl.index] == b]
                              for b in range(len(self.buckets.buckets))]
        self._bucket_loose = [[(p, v) for p, v in self._loose if self.buckets.bucket_of[self._index[id(p)]] == b]
                              for b in range(len(self.buckets.buckets))]
        if self.overlap:
            self.wstage.on_ready = self.buckets.ready
            for p, _ in self._loose:
                p.register_post_accumulate_grad_hook(lambda q, i=self._index[id(p)]: self.buckets.ready(i))
        self.niter = 0
        self.step = 0
        self.loss_G_tot = None
        # CUDA-graph replay of the step (static shapes): graph 1 = prologue + forward + loss + backward,
        # [eager NCCL all-reduce of the flat gradient], graph 2 = Adam(W)+EMA + gradient zeroing.
        self.use_graph = bool(cuda_graph)
        if self.use_graph and iter_size != 1:
            raise NotImplementedError("cuda_graph=True requires train_iter_size == 1")
        self.graph_warmup = graph_warmup
        self._graph_fb = None
        self._graph_opt = None
        self._static = None
        self._eager_steps = 0
        self.step_dev = torch.zeros((), dtype=torch.int32, device=self.device)
        self.launches_per_step = 0  # kernels of libjg_b200.so per step (counted on an eager step)
        # weights written through torch after construction (load_state_dict of a checkpoint, base_model.load_networks
        # :957-1103): re-pack the bf16 copies at once — captured graphs never pass through ConvPack.get()
        self.netG_A.register_load_state_dict_post_hook(lambda module, incompatible: self.packset.refresh())

    # -- data -----------------------------------------------------------------------------------
    def set_input(self, data, non_blocking=True):
        """data: {"A": cond image y_t, "B": ground truth, "B_label_mask": int64/float mask [B,1,H,W]} (+ "ref_A": the
        reference image when the denoiser is a UNetGeneratorRefAttn, palette_model.py:373-374, 586-588)"""
        m = data.get("B_label_mask")
        if self.use_ref and "ref_A" not in data:
            raise RuntimeError('PaletteTrainer: this generator needs data["ref_A"]')
        r = data["ref_A"] if self.use_ref else None
        if self._static is not None:
            # graph mode: refill the captured input buffers in place
            st = self._static
            if tuple(data["A"].shape) != tuple(st["A"].shape) or (m is None) != (st["M"] is None):
                raise RuntimeError("cuda_graph=True: input shapes must stay fixed after capture")
            st["A"].copy_(data["A"], non_blocking=non_blocking)
            st["B"].copy_(data["B"], non_blocking=non_blocking)
            if m is not None:
                st["M"].copy_(m, non_blocking=non_blocking)
            if r is not None:
                st["R"].copy_(r, non_blocking=non_blocking)
            if st["C"] is not None:
                st["C"].copy_(data["B_label_cls"], non_blocking=non_blocking)
            return
        self.y_t = data["A"].to(self.device, non_blocking=non_blocking)
        self.gt_image = data["B"].to(self.device, non_blocking=non_blocking)
        self.mask = None if m is None else m.to(self.device, non_blocking=non_blocking)
        self.ref_A = None if r is None else r.to(self.device, non_blocking=non_blocking)
        c = data.get("B_label_cls") if "class" in self.conditioning else None
        if "class" in self.conditioning and c is None:
            raise RuntimeError('PaletteTrainer: conditioning "class" needs data["B_label_cls"]')
        self.cls = None if c is None else c.to(self.device, non_blocking=non_blocking).long()
        self.cond_image = self.y_t

    def broadcast_parameters(self):
        self.comm.broadcast(self.flat.data, root=0)
        self.packset.refresh()

    def comm_stats(self):
        st = self.comm.stats()
        if st is None:
            return None
        st.update(buckets=len(self.buckets.buckets), overlapped=self.overlap,
                  bucket_mb=[round(4e-6 * (hi - lo), 1) for lo, hi, _ in self.buckets.buckets],
                  payload_mb_per_step=round(4e-6 * self.flat.total, 1), dtype="f32",
                  what="library-owned NCCL communicator (jg_comm_*), one all-reduce per gradient bucket on a "
                       "communication stream forked from the backward pass (captured inside the step's CUDA graph)")
        return st

    # -- step -----------------------------------------------------------------------------------
    def compute_palette_loss(self, noise=None, t=None, u=None, drop_u=None):
        mask, cls = self.mask, self.cls
        if self.dropout_prob > 0.0:
            # palette_model.py:565-584; drop_u = the torch.rand(B) draw (explicit in the tests)
            if drop_u is None:
                drop_u = torch.rand(self.gt_image.shape[0], device=self.device)
            if mask is not None:
                mask = K.mask_class_dropout(mask, drop_u, self.dropout_prob, self.num_classes - 1)
            if cls is not None:
                cls = torch.where(drop_u < self.dropout_prob, torch.full_like(cls, self.num_classes - 1), cls)
        kw = {"cls": cls} if cls is not None else {}
        self.loss_G_tot = self.netG_A.forward_loss(self.gt_image, self.cond_image, mask, noise=noise,
                                                   lambda_G=self.lambda_G, use_minsnr=self.use_minsnr, l1=self.l1,
                                                   t=t, u=u, ref=self.ref_A, **kw)
        return self.loss_G_tot

    @staticmethod
    def _fold(pairs):
        """one multi-tensor add of the gradients autograd returned as tensors into their flat-buffer slices"""
        pairs = [(v, p.grad) for p, v in pairs if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if pairs:
            torch._foreach_add_([v for v, _ in pairs], [g.reshape(v.shape) for v, g in pairs])

    def _reduce_bucket(self, b):
        """Bucket b of the flat gradient is final on this rank: unpack its convolution accumulators, fold its small
        gradients, and start its all-reduce on the communication stream (the backward pass keeps running)."""
        self.wstage.flush(self._bucket_slots[b], tag=b)
        self._fold(self._bucket_loose[b])
        lo, hi, _ = self.buckets.buckets[b]
        self.comm.allreduce_async(self.flat.grad[lo:hi])

    def _forward_backward(self, noise=None, t=None, u=None, reduce=True):
        """Forward + loss + backward (+ the gradient exchange across ranks unless reduce=False: an accumulation
        micro-step, base_model.py:1313-1315 no_sync)."""
        self.flat.rebind_grads()
        loss = self.compute_palette_loss(noise=noise, t=t, u=u)
        # Parameters whose gradient arrives as a tensor (norm gains, biases, linears, padded convs: ~250 of them)
        # start the backward WITHOUT a .grad: autograd then just keeps the incoming tensor instead of launching one
        # tiny add kernel per parameter, and a multi-tensor add folds them into the flat gradient buffer.
        for p, _ in self._loose:
            p.grad = None
        overlap = self.overlap and reduce
        self.wstage.begin()
        if overlap:
            self.buckets.begin()
        (loss / self.iter_size).backward()
        self.wstage.end()
        if overlap:
            self.buckets.finish()  # whatever did not complete during the backward pass, in launch order
            self.comm.wait()
        else:
            self.wstage.flush()
            self._fold(self._loose)
            if reduce and self.world > 1:
                self.comm.allreduce_async(self.flat.grad)
                self.comm.wait()
        for p, v in self._loose:
            p.grad = v
        # hand out a graph-free scalar: a retained autograd graph would keep this iteration's
        # AccumulateGrad nodes (and their stream binding) alive across iterations / graph capture
        self.loss_G_tot = loss.detach()
        return self.loss_G_tot

    def eager_step(self):
        """One full step launched kernel by kernel (no graph replay), on the current inputs: profiling entry point."""
        loss = self._forward_backward()
        self._optimizer_step()
        return loss

    def reduced_gradient(self, noise=None, t=None, u=None):
        """Forward + backward + the gradient exchange, without the optimizer: returns a copy of the flat fp32 gradient
        SUMMED over the ranks (divide by the world size for DDP's mean) and leaves the gradient buffer zeroed.
        Diagnostic / test entry point (tests/test_gpu_multi.py)."""
        self._forward_backward(noise=noise, t=t, u=u)
        g = self.flat.grad.clone()
        self.flat.grad.zero_()
        return g

    def _optimizer_step(self):
        grad_scale = 1.0 / self.world
        self.step += 1
        K.adamw_ema_step(self.flat.data, self.flat.grad, self.exp_avg, self.exp_avg_sq, self.ema, step=self.step,
                         step_dev=self.step_dev, grad_scale=grad_scale, ema_beta=self.ema_beta,
                         ema_init=not self.ema_started, **self.hp)
        self.ema_started = True
        self.flat.grad.zero_()
        self.packset.refresh()

    def _capture(self):
        """Capture the step into CUDA graphs (called once, after `graph_warmup` eager steps)."""
        self._static = {"A": self.y_t.clone(), "B": self.gt_image.clone(),
                        "M": None if self.mask is None else self.mask.clone(),
                        "R": None if self.ref_A is None else self.ref_A.clone(),
                        "C": None if self.cls is None else self.cls.clone()}
        self.y_t = self.cond_image = self._static["A"]
        self.gt_image = self._static["B"]
        self.mask = self._static["M"]
        self.ref_A = self._static["R"]
        self.cls = self._static["C"]
        self.loss_G_tot = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        self._graph_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_fb):
            self._static_loss = self._forward_backward()
        self._graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_opt):
            self._optimizer_step()
        # the capture itself did not execute: undo its host-side bookkeeping
        self.step -= 1

    def optimize_parameters(self, noise=None, t=None, u=None):
        self.niter += 1
        explicit = noise is not None or t is not None or u is not None
        if self.use_graph and not explicit:
            if self._graph_fb is None and self._eager_steps >= self.graph_warmup:
                self._capture()
            if self._graph_fb is not None:
                self._graph_fb.replay()  # forward + backward + the bucketed all-reduce (a forked branch of the graph)
                self._graph_opt.replay()
                self.step += 1
                from . import lib as L
                L.launch_count[0] += self.launches_per_step  # the replayed graphs hold the same kernels
                self.loss_G_tot = self._static_loss
                return self._static_loss
        from . import lib as L
        n0 = L.launch_count[0]
        last = self.niter % self.iter_size == 0
        loss = self._forward_backward(noise=noise, t=t, u=u, reduce=last)
        if last:
            self._optimizer_step()
        self._eager_steps += 1
        self.launches_per_step = L.launch_count[0] - n0
        return loss

    # -- state ----------------------------------------------------------------------------------
    def state_dict(self):
        """Everything a resumed run needs beyond netG_A.state_dict() (base_model.save_networks :824-868 stores the
        nets and the optimizers): Adam moments, step counters and the EMA copy, keyed by parameter name."""
        sd = {"step": self.step, "niter": self.niter, "ema_started": self.ema_started,
              "exp_avg": {k: v.clone() for k, v in self.flat.unflatten(self.exp_avg).items()},
              "exp_avg_sq": {k: v.clone() for k, v in self.flat.unflatten(self.exp_avg_sq).items()}}
        if self.ema is not None:
            sd["ema"] = {k: v.clone() for k, v in self.flat.unflatten(self.ema).items()}
        return sd

    def load_state_dict(self, sd):
        """Inverse of state_dict(); the live weights are restored separately by netG_A.load_state_dict (which
        re-packs the bf16 copies through the hook installed in __init__)."""
        for name, flat in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq), ("ema", self.ema)):
            if flat is None or name not in sd:
                continue
            views = self.flat.unflatten(flat)
            missing = set(views) - set(sd[name])
            if missing:
                raise KeyError("PaletteTrainer.load_state_dict: %s lacks %s" % (name, sorted(missing)[:3]))
            for k, v in views.items():
                v.copy_(sd[name][k])
        self.step = int(sd["step"])
        self.niter = int(sd.get("niter", self.step * self.iter_size))
        self.ema_started = bool(sd.get("ema_started", self.step > 0))
        self.step_dev.fill_(self.step)  # the device-side counter the captured optimizer graph increments
        self.packset.refresh()

    def ema_state_dict(self):
        """state_dict of netG_A_ema (base_model.py:1284-1297) — parameters from the flat EMA buffer,
        buffers copied from the live net."""
        sd = {k: v.clone() for k, v in self.netG_A.state_dict().items()}
        if self.ema is not None and self.ema_started:
            for k, v in self.flat.unflatten(self.ema).items():
                sd[k] = v.clone()
        return sd
