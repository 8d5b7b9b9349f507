"""B2BModel.optimize_parameters for the b2b video backbone (/root/reference/models/b2b_model.py: set_input,
compute_b2b_loss :1081-1168 without the perceptual terms, compute_step / ema_step of base_model.py:1250-1297):
flow-matching forward of nets_jit.B2BGenerator + masked pseudo-Huber loss + backward + ONE fused AdamW(+EMA) launch over
flat fp32 buffers.  Mirrors the reference's quirk that `set_requires_grad` makes the "fixed" sin-cos `pos_embed`
trainable (pinned by tests/golden/b2b_plumbing.pt)."""
import torch

from . import dp
from . import kernels as K
from . import nets
from .trainer import FlatParams


class B2BTrainer:
    def __init__(self, net, lr=1e-4, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.0, optim="adamw", ema=True,
                 ema_beta=0.999, lambda_G=1.0, use_cond=False, device="cuda", train_pos_embed=True, cuda_graph=False,
                 graph_warmup=2, process_group=None):
        if not torch.cuda.is_available():
            raise RuntimeError("B2BTrainer needs a CUDA device (the B200 kernels have no CPU fallback)")
        if optim not in ("adamw", "adam"):
            raise NotImplementedError("optimizer %r (adam / adamw are implemented)" % optim)
        self.device = torch.device(device)
        self.net = net.to(self.device)
        if train_pos_embed:
            self.net.b2b_model.pos_embed.requires_grad_(True)
        self.flat = FlatParams(self.net)
        # data parallel (DDP's mean gradient, base_model.py:725-737): the flat gradient is summed over the ranks on the
        # library's communicator (dp.Comm: NCCL on its own stream, capturable; torch.distributed on the CPU) and 1/world
        # is folded into the optimizer kernel.  One clip's gradient is 0.6 GB of fp32 for JiTVid-B/16: one exchange.
        self.pg = process_group
        self.world = dp.world_size(process_group)
        self.comm = dp.Comm(process_group, self.device)
        self.exp_avg = torch.zeros_like(self.flat.data)
        self.exp_avg_sq = torch.zeros_like(self.flat.data)
        self.ema = torch.zeros_like(self.flat.data) if ema else None
        self.ema_started = False
        self.ema_beta = ema_beta
        self.hp = dict(lr=lr, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay, adamw=(optim == "adamw"))
        self.step = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.lambda_G = lambda_G
        self.use_cond = use_cond
        self.loss_G_tot = None
        # CUDA-graph replay of the whole step (implicit random draws only): one clip per GPU is ~700 small launches
        self.use_graph = bool(cuda_graph)
        self.graph_warmup = int(graph_warmup)
        self._graph = None
        self._eager_steps = 0
        self._static = None
        self.launches_per_step = 0

    def broadcast_parameters(self):
        """rank 0's weights everywhere (what DistributedDataParallel does at construction)"""
        self.comm.broadcast(self.flat.data, root=0)
        nets.invalidate_packed_weights()

    def set_input(self, data):
        """data["B"] clip [B, F, 3, H, W], data["B_label_mask"] [B, F, 1, H, W], data["A"] the conditioning clip."""
        dev = self.device
        gt = data["B"].to(dev, non_blocking=True).float()
        mask = data["B_label_mask"].to(dev, non_blocking=True).float()
        cond = data["A"].to(dev, non_blocking=True).float() if self.use_cond else None
        if self._static is not None:   # the captured graph reads these buffers
            self._static["B"].copy_(gt)
            self._static["M"].copy_(mask)
            if cond is not None:
                self._static["A"].copy_(cond)
            return
        self.gt, self.mask, self.cond = gt, mask, cond
        self.label = torch.zeros(self.gt.shape[0], dtype=torch.long, device=dev)

    def _step(self, t_base=None, e=None):
        self.flat.rebind_grads()
        loss = self.net.forward_loss(self.gt, self.mask, self.cond, self.label, t_base=t_base, e=e, lambda_G=self.lambda_G)
        loss.backward()
        if self.world > 1:
            self.comm.allreduce_async(self.flat.grad)
            self.comm.wait()
        self.step += 1
        K.adamw_ema_step(self.flat.data, self.flat.grad, self.exp_avg, self.exp_avg_sq, self.ema, step=self.step,
                         step_dev=self.step_dev, grad_scale=1.0 / self.world, ema_beta=self.ema_beta,
                         ema_init=not self.ema_started, **self.hp)
        self.ema_started = True
        self.flat.grad.zero_()
        nets.invalidate_packed_weights()   # the fused optimizer wrote the masters through raw pointers
        return loss.detach()

    def _capture(self):
        self._static = {"B": self.gt.clone(), "M": self.mask.clone(), "A": None if self.cond is None else self.cond.clone()}
        self.gt, self.mask, self.cond = self._static["B"], self._static["M"], self._static["A"]
        import gc
        gc.collect()
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self._step()
        self.step -= 1   # the capture did not execute

    def optimize_parameters(self, t_base=None, e=None):
        from . import lib as L
        explicit = t_base is not None or e is not None
        if self.use_graph and not explicit:
            if self._graph is None and self._eager_steps >= self.graph_warmup:
                self._capture()
            if self._graph is not None:
                self._graph.replay()
                self.step += 1
                L.launch_count[0] += self.launches_per_step
                self.loss_G_tot = self._static_loss
                return self._static_loss
        n0 = L.launch_count[0]
        self.loss_G_tot = self._step(t_base, e)
        self._eager_steps += 1
        self.launches_per_step = L.launch_count[0] - n0
        return self.loss_G_tot

    def eager_step(self):
        self.loss_G_tot = self._step()
        return self.loss_G_tot

    def params(self):
        return self.flat.unflatten(self.flat.data)

    def state_dict(self):
        """What a resumed run needs beyond net.state_dict(): Adam moments, step counter, EMA copy (by parameter name)."""
        sd = {"step": self.step, "ema_started": self.ema_started,
              "exp_avg": {k: v.clone() for k, v in self.flat.unflatten(self.exp_avg).items()},
              "exp_avg_sq": {k: v.clone() for k, v in self.flat.unflatten(self.exp_avg_sq).items()}}
        if self.ema is not None:
            sd["ema"] = {k: v.clone() for k, v in self.flat.unflatten(self.ema).items()}
        return sd

    def load_state_dict(self, sd):
        for name, flat in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq), ("ema", self.ema)):
            if flat is None or name not in sd:
                continue
            views = self.flat.unflatten(flat)
            missing = set(views) - set(sd[name])
            if missing:
                raise KeyError("B2BTrainer.load_state_dict: %s lacks %s" % (name, sorted(missing)[:3]))
            for k, v in views.items():
                v.copy_(sd[name][k])
        self.step = int(sd["step"])
        self.ema_started = bool(sd.get("ema_started", self.step > 0))
        self.step_dev.fill_(self.step)
        nets.invalidate_packed_weights()

    def ema_state_dict(self):
        return self.flat.unflatten(self.ema) if self.ema is not None else None
