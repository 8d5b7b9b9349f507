"""GAN (CUT-style) training inner loop for one generator + one PatchGAN discriminator on the B200 kernels.

Mirrors the two NetworkGroups that `BaseModel.optimize_parameters` iterates for `cut_model`
(models/cut_model.py:406-437; loss wiring models/base_gan_model.py:382-419, 457-503;
models/modules/loss.py:288-313):

    G group: fake_B = netG_A(real_A); loss_G_GAN = lambda_gan * GANLoss(netD(fake_B), True, relu=False)
             backward (D frozen via set_requires_grad) -> optimizer_G (+ EMA)
    D group: loss_D = 0.5 * (GANLoss(netD(real_B), True) + GANLoss(netD(fake_B.detach()), False))
             backward -> optimizer_D

The contrastive PatchNCE terms of CUT (`compute_G_loss_cut`) are a "next" row and are not
part of this step.  Parameters / gradients / Adam moments are flat fp32 buffers per network; the gradient
all-reduce is one SUM per group; Adam(+EMA) is one fused launch per network.
"""
import torch

from . import dp
from . import kernels as K
from . import nets
from . import ops
from .trainer import FlatParams


class _FlatAdam:
    def __init__(self, module, lr, beta1, beta2, eps, weight_decay, adamw, ema_beta=None):
        self.flat = FlatParams(module)
        self.m = torch.zeros_like(self.flat.data)
        self.v = torch.zeros_like(self.flat.data)
        self.ema = torch.zeros_like(self.flat.data) if ema_beta is not None else None
        self.ema_beta = ema_beta if ema_beta is not None else 0.0
        self.hp = dict(lr=lr, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay, adamw=adamw)
        self.step = 0
        # the step count lives on the device too, so that a CUDA graph holding this update stays valid when replayed
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.flat.data.device)

    def apply(self, pg):
        scale = dp.allreduce_sum_(self.flat.grad, pg)
        self.step += 1
        K.adamw_ema_step(self.flat.data, self.flat.grad, self.m, self.v, self.ema, step=self.step, step_dev=self.step_dev,
                         grad_scale=scale, ema_beta=self.ema_beta, ema_init=(self.step == 1), **self.hp)
        self.flat.grad.zero_()
        nets.invalidate_packed_weights()

    # -- checkpoint state beyond the net's own state_dict (base_model.save_networks :824-868 stores the nets AND the
    #    optimizers): Adam moments, step counters, the EMA copy, keyed by parameter name ---------------------------------
    def state_dict(self):
        sd = {"step": self.step, "exp_avg": {k: v.clone() for k, v in self.flat.unflatten(self.m).items()},
              "exp_avg_sq": {k: v.clone() for k, v in self.flat.unflatten(self.v).items()}}
        if self.ema is not None:
            sd["ema"] = {k: v.clone() for k, v in self.flat.unflatten(self.ema).items()}
        return sd

    def load_state_dict(self, sd):
        for name, flat in (("exp_avg", self.m), ("exp_avg_sq", self.v), ("ema", self.ema)):
            if flat is None or name not in sd:
                continue
            views = self.flat.unflatten(flat)
            missing = set(views) - set(sd[name])
            if missing:
                raise KeyError("load_state_dict: %s lacks %s" % (name, sorted(missing)[:3]))
            for k, v in views.items():
                v.copy_(sd[name][k])
        self.step = int(sd["step"])
        self.step_dev.fill_(self.step)
        nets.invalidate_packed_weights()


class GanTrainer:
    def __init__(self, netG_A, netD_B, gan_mode="lsgan", lambda_gan=1.0, G_lr=2e-4, D_lr=1e-4, beta1=0.9, beta2=0.999,
                 eps=1e-8, weight_decay=0.0, optim="adam", G_ema_beta=None, device=None, process_group=None):
        if not torch.cuda.is_available():
            raise RuntimeError("joligen_b200.GanTrainer needs a CUDA device (there is no CPU path)")
        from .nets_gan import GANLoss
        self.device = torch.device(device if device is not None else "cuda")
        self.netG_A = netG_A.to(self.device)
        self.netD_B = netD_B.to(self.device)
        self.crit = GANLoss(gan_mode)
        self.lambda_gan = lambda_gan
        adamw = optim == "adamw"
        self.optG = _FlatAdam(self.netG_A, G_lr, beta1, beta2, eps, weight_decay, adamw, G_ema_beta)
        self.optD = _FlatAdam(self.netD_B, D_lr, beta1, beta2, eps, weight_decay, adamw)
        self.pg = process_group
        self.niter = 0
        self.loss_G_tot = self.loss_D_tot = None

    def set_input(self, data, non_blocking=True):
        """data: {"A": source-domain images, "B": target-domain images} NCHW fp32 in [-1, 1]"""
        self.real_A = data["A"].to(self.device, non_blocking=non_blocking)
        self.real_B = data["B"].to(self.device, non_blocking=non_blocking)

    @staticmethod
    def set_requires_grad(net, flag):
        for p in net.parameters():
            p.requires_grad = flag

    def optimize_parameters(self):
        self.niter += 1
        a = ops.to_nhwc(self.real_A)
        b = ops.to_nhwc(self.real_B)
        # ---- G group
        self.optG.flat.rebind_grads()
        self.set_requires_grad(self.netD_B, False)
        fake = self.netG_A.forward_nhwc(a)
        self.fake_B = fake
        pred_fake = self.netD_B.forward_nhwc(fake)
        loss_G = self.lambda_gan * self.crit.forward_nhwc(pred_fake, True, relu=False)
        loss_G.backward()
        self.optG.apply(self.pg)
        self.loss_G_tot = loss_G.detach()
        # ---- D group
        self.set_requires_grad(self.netD_B, True)
        self.optD.flat.rebind_grads()
        pred_real = self.netD_B.forward_nhwc(b)
        pred_fake = self.netD_B.forward_nhwc(fake.detach())
        loss_D = 0.5 * (self.crit.forward_nhwc(pred_real, True) + self.crit.forward_nhwc(pred_fake, False))
        loss_D.backward()
        self.optD.apply(self.pg)
        self.loss_D_tot = loss_D.detach()
        self.fake_B = fake.detach()   # the value only: the generator's graph is not kept across steps
        return self.loss_G_tot, self.loss_D_tot

    def state_dict(self):
        """optimizer-side state of both groups (the nets' weights travel in their own state_dicts)"""
        return {"niter": self.niter, "G": self.optG.state_dict(), "D": self.optD.state_dict()}

    def load_state_dict(self, sd):
        self.niter = int(sd["niter"])
        self.optG.load_state_dict(sd["G"])
        self.optD.load_state_dict(sd["D"])
