"""The CUT training step (BASELINE.json config 3) on the B200 kernels: one `optimize_parameters()` of
`models/cut_model.py` — the (G_A, F) group, then the D group (cut_model.py:406-437):

    forward_cut (:608-640)         fake = G(cat(real_A, real_B));  fake_B = fake[:B], idt_B = fake[B:]
    compute_G_loss_GAN             lambda_GAN * GANLoss(D(fake_B), True, relu=False)   (base_gan_model.py:467-503)
    compute_G_loss_cut (:715-845)  NCE(real_A -> fake_B) and, with --alg_cut_nce_idt, NCE(real_B -> idt_B), averaged:
                                   calculate_feats (:848-887) = encoder features of both images at --alg_cut_nce_layers,
                                   the SAME random positions for keys and queries, PatchSampleF MLP, PatchNCE loss;
                                   the layer sum is divided by the number of REQUESTED layers (:892, 909)
    compute_D_loss                 0.5 * (GANLoss(D(real_B), True) + GANLoss(D(fake_B.detach()), False))

--alg_cut_nce_loss patchnce or monce (the example's default: Sinkhorn-weighted negatives, csrc/nce.cu).
Parameters / gradients / Adam moments are flat fp32 buffers per network (G, F, D), one SUM all-reduce per group.
Checked on a B200 against the reference's own control path for both losses
(tests/test_gpu_widen_cut.py::test_cut_trainer_vs_reference_plumbing, cut_plumbing*.pt).
"""
import torch

from . import nets_cut
from . import ops
from .trainer_gan import _FlatAdam


class CutTrainer:
    def __init__(self, netG_A, netF, netD_B, nce_layers=(0, 4, 8, 12, 16), num_patches=256, nce_T=0.07, lambda_NCE=1.0,
                 nce_idt=True, nce_loss="patchnce", nce_includes_all_negatives_from_minibatch=False, gan_mode="lsgan",
                 lambda_gan=1.0, G_lr=2e-4, D_lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, optim="adam",
                 device=None, process_group=None, cuda_graph=False, graph_warmup=2):
        if not torch.cuda.is_available():
            raise RuntimeError("joligen_b200.CutTrainer needs a CUDA device (there is no CPU path)")
        if nce_loss not in ("patchnce", "monce"):
            raise NotImplementedError("B200 CutTrainer: --alg_cut_nce_loss %r (patchnce / monce)" % nce_loss)
        if not netF.mlp_init:
            raise RuntimeError("CutTrainer: call netF.data_dependent_initialize(netG_A.get_feats(x, nce_layers)) first "
                               "(cut_model.data_dependent_initialize, :504-538)")
        from types import SimpleNamespace
        from .nets_gan import GANLoss
        self.device = torch.device(device if device is not None else "cuda")
        self.netG_A = netG_A.to(self.device)
        self.netF = netF.to(self.device)
        self.netD_B = netD_B.to(self.device)
        self.crit = GANLoss(gan_mode)
        self.crit_nce = (nets_cut.MoNCELoss if nce_loss == "monce" else nets_cut.PatchNCELoss)(SimpleNamespace(
            alg_cut_nce_T=nce_T, alg_cut_num_patches=num_patches,
            alg_cut_nce_includes_all_negatives_from_minibatch=nce_includes_all_negatives_from_minibatch))
        self.nce_layers = list(nce_layers)
        self.num_patches, self.lambda_NCE, self.nce_idt, self.lambda_gan = num_patches, lambda_NCE, nce_idt, lambda_gan
        adamw = optim == "adamw"
        self.optG = _FlatAdam(self.netG_A, G_lr, beta1, beta2, eps, weight_decay, adamw)
        self.optF = _FlatAdam(self.netF, G_lr, beta1, beta2, eps, weight_decay, adamw)  # optimizer_F uses train_G_lr
        self.optD = _FlatAdam(self.netD_B, D_lr, beta1, beta2, eps, weight_decay, adamw)
        self.pg = process_group
        self.niter = 0
        self.loss_G_tot = self.loss_G_GAN = self.loss_G_NCE = self.loss_G_NCE_Y = self.loss_D_tot = None
        # CUDA-graph replay of the whole step (both optimizer groups): ~3 900 small launches per step are launch-bound
        # when issued from Python.  EXPERIMENTAL, off by default: the one capture attempted on a B200 ended in
        # cudaErrorStreamCaptureImplicit (profiles/r02_cut_graph_capture_error.log) — the stored, un-detached losses kept
        # the warm-up's autograd graph and its default-stream AccumulateGrad nodes alive; they are detached now, but the
        # capture has NOT been re-run on hardware since.  The eager path is the tested one.  Single process only.
        self.use_graph = bool(cuda_graph) and process_group is None
        self.graph_warmup = int(graph_warmup)
        self._graph = None
        self._static = None
        self._eager_steps = 0
        self.launches_per_step = 0

    def set_input(self, data, non_blocking=True):
        """data: {"A": source-domain images, "B": target-domain images} NCHW fp32 in [-1, 1]"""
        a = data["A"].to(self.device, non_blocking=non_blocking)
        b = data["B"].to(self.device, non_blocking=non_blocking)
        if self._static is not None:   # the captured graph reads these buffers
            self._static["A"].copy_(a)
            self._static["B"].copy_(b)
            return
        self.real_A, self.real_B = a, b

    @staticmethod
    def set_requires_grad(net, flag):
        for p in net.parameters():
            p.requires_grad = flag

    def _feats(self, x_nhwc):
        """ResnetGenerator.get_feats on an NHWC tensor: the encoder runs again on x (as the reference's does)."""
        _, feats = self.netG_A._runner.run(self.netG_A.encoder.model, x_nhwc, collect=set(self.nce_layers))
        return [f for _, f in feats]

    def _nce(self, src, tgt, patch_ids=None):
        feat_q, feat_k = self._feats(tgt), self._feats(src)
        k_pool, ids = self.netF.forward_nhwc(feat_k, self.num_patches, patch_ids)
        q_pool, _ = self.netF.forward_nhwc(feat_q, self.num_patches, ids)
        total = 0.0
        b = src.shape[0]
        for fq, fk in zip(q_pool, k_pool):
            total = total + (self.crit_nce(feat_q=fq, feat_k=fk, current_batch=b) * self.lambda_NCE).mean()
        return total / len(self.nce_layers)

    def _capture(self):
        self._static = {"A": self.real_A.clone(), "B": self.real_B.clone()}
        self.real_A, self.real_B = self._static["A"], self._static["B"]
        import gc
        gc.collect()
        torch.cuda.synchronize()
        # one more real step on the side stream the capture will use: autograd's gradient accumulators then belong to that
        # stream (on the legacy default stream they would make it wait on the capturing stream, which CUDA refuses)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step(None, None)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=side):
            self._static_losses = self._step(None, None)
        for o in (self.optG, self.optF, self.optD):
            o.step -= 1   # the capture did not execute
        self.niter -= 1

    def optimize_parameters(self, patch_ids_A=None, patch_ids_B=None):
        """patch_ids_*: optional explicit positions (one LongTensor per NCE layer) instead of torch.randperm draws."""
        from . import lib as L
        if self.use_graph and patch_ids_A is None and patch_ids_B is None:
            if self._graph is None and self._eager_steps >= self.graph_warmup:
                self._capture()
            if self._graph is not None:
                self._graph.replay()
                self.niter += 1
                for o in (self.optG, self.optF, self.optD):
                    o.step += 1
                L.launch_count[0] += self.launches_per_step
                self.loss_G_tot, self.loss_D_tot = self._static_losses
                return self._static_losses
        n0 = L.launch_count[0]
        out = self._step(patch_ids_A, patch_ids_B)
        self._eager_steps += 1
        self.launches_per_step = L.launch_count[0] - n0
        return out

    def eager_step(self):
        return self._step(None, None)

    def state_dict(self):
        """optimizer-side state of the three groups (the nets' weights travel in their own state_dicts)"""
        return {"niter": self.niter, "G": self.optG.state_dict(), "F": self.optF.state_dict(), "D": self.optD.state_dict()}

    def load_state_dict(self, sd):
        if self._graph is not None:
            raise RuntimeError("CutTrainer.load_state_dict after the CUDA graph was captured")
        self.niter = int(sd["niter"])
        for key, opt in (("G", self.optG), ("F", self.optF), ("D", self.optD)):
            opt.load_state_dict(sd[key])

    def _step(self, patch_ids_A=None, patch_ids_B=None):
        self.niter += 1
        a = ops.to_nhwc(self.real_A)
        b = ops.to_nhwc(self.real_B)
        n = a.shape[0]
        # ---- (G_A, F) group
        self.optG.flat.rebind_grads()
        self.optF.flat.rebind_grads()
        self.set_requires_grad(self.netD_B, False)
        real = torch.cat([a, b], dim=0) if self.nce_idt else a
        fake = self.netG_A.forward_nhwc(real)
        fake_B = fake[:n]
        self.fake_B = fake_B
        self.loss_G_GAN = self.lambda_gan * self.crit.forward_nhwc(self.netD_B.forward_nhwc(fake_B), True, relu=False)
        self.loss_G_NCE = self._nce(a, fake_B, patch_ids_A)
        if self.nce_idt:
            self.loss_G_NCE_Y = self._nce(b, fake[n:], patch_ids_B)
            nce_both = (self.loss_G_NCE + self.loss_G_NCE_Y) * 0.5
        else:
            self.loss_G_NCE_Y = torch.zeros((), device=self.device)
            nce_both = self.loss_G_NCE
        loss_G = self.loss_G_GAN + nce_both
        loss_G.backward()
        self.optG.apply(self.pg)
        self.optF.apply(self.pg)
        # keep VALUES only: a stored loss with its graph keeps the whole iteration's autograd graph alive — memory, and
        # the AccumulateGrad nodes of the eager warm-up (made on the default stream) would be reused inside a later
        # CUDA-graph capture (cudaErrorStreamCaptureImplicit, profiles/r02_cut_graph_capture_error.log)
        self.loss_G_tot = loss_G.detach()
        self.loss_G_GAN, self.loss_G_NCE = self.loss_G_GAN.detach(), self.loss_G_NCE.detach()
        self.loss_G_NCE_Y = self.loss_G_NCE_Y.detach()
        # ---- D group
        self.set_requires_grad(self.netD_B, True)
        self.optD.flat.rebind_grads()
        pred_real = self.netD_B.forward_nhwc(b)
        pred_fake = self.netD_B.forward_nhwc(fake_B.detach())
        loss_D = 0.5 * (self.crit.forward_nhwc(pred_real, True) + self.crit.forward_nhwc(pred_fake, False))
        loss_D.backward()
        self.optD.apply(self.pg)
        self.loss_D_tot = loss_D.detach()
        self.fake_B = fake_B.detach()
        return self.loss_G_tot, self.loss_D_tot
