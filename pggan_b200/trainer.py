"""The PGGAN D+G training step and the scale/alpha schedule around it — mirror of
reference pggan/model.py (ProgressiveGAN.train_step :206-255, schedule :131-204) without the data
loading, checkpoint and wandb glue (out of scope).

step variants:
  "L"  the reference's LIVE step: D loss = BCE(real,1)+BCE(fake,0)+R1 (pggan/loss.py:16-27)
  "W"  the WGAN-GP step of the commented-out caller (pggan/loss.py:29-51), wired as
       zero_grad -> GP.backward -> loss_D.backward -> Adam, drift reported as a value.
Both run D forward three times (real, fake, interpolates/R1 graph), G forward twice, and the double
backward through D.  The G step freezes D's parameters (the reference computes and discards those
gradients, SURVEY.md D15 — results are identical).
"""
from __future__ import annotations

import types
from typing import Optional

import torch
import torch.distributed as dist

from . import checkpoint, ops
from .dist import GradSync, broadcast_parameters
from .loss import WGANGPLoss
from .nets import Discriminator, Generator
from .optim import FlatGrads, FusedAdam

# reference configs.yaml values on the hot path
DEFAULT_ARGS = dict(
    lr_G=1e-4, lr_D=1e-5, beta1=0, beta2=0.99, W_adv=1, W_gp=10, W_drift_D=0.001, batch_per_gpu=16,
    max_step=2000000, latent_dim=512, input_dim=3, output_dim=3, init_bias_to_zero=True,
    depths=[512, 512, 512, 512, 256, 128, 64], max_step_at_scale=[10000, 20000, 40000, 80000, 80000, 80000, 80000,
                                                                  80000, 80000],
    alpha_jump_start=[-1, 2000, 4000, 10000, 10000, 10000, 10000, 10000, 10000],
    alpha_jump_interval=[0, 100, 100, 100, 100, 100, 100, 100, 100],
    alpha_jump_Ntimes=[0, 100, 200, 400, 400, 400, 400, 400, 400],
    LReLU_slope=0.2, generator_last_activation=None, apply_pixel_norm=True, apply_minibatch_norm=True,
    equalized_lr=True, decision_layer_size=1, isMaster=False,
)
# the paper's CelebA-HQ widths for the 1024^2 configuration
DEPTHS_1024 = [512, 512, 512, 512, 256, 128, 64, 32, 16]


def make_args(**overrides):
    d = dict(DEFAULT_ARGS)
    d.update(overrides)
    return types.SimpleNamespace(**d)


class ProgressiveGAN:
    def __init__(self, args=None, gpu=0, variant: str = "W", data_parallel: Optional[bool] = None,
                 bucket_bytes: int = 32 << 20):
        self.args = args if args is not None else make_args()
        self.gpu = gpu
        self.device = torch.device("cuda", gpu) if isinstance(gpu, int) else torch.device(gpu)
        self.variant = variant
        self.scale_index = 0
        self.alpha_index = 0
        self.alpha_jump_value = 0
        self.next_scale_jump_step = self.args.max_step_at_scale[0]
        self.next_alpha_jump_step = self.args.alpha_jump_start[0]
        if data_parallel is None:
            data_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.data_parallel = data_parallel
        self.world = dist.get_world_size() if data_parallel else 1
        # static loss scale for the half-precision gradient tensors: both losses (and the penalty's own backward) are
        # multiplied by it, the fused Adam kernel divides it out together with the 1/world of the gradient mean
        self.loss_scale = 1024.0 if ops._REAL[ops._fmt(ops.G)] == torch.float16 else 1.0
        self.bucket_bytes = bucket_bytes
        self.initialize_models()
        self.set_multi_GPU()
        self.set_optimizers()
        self.set_loss_collector()

    # ------------------------------------------------------------------ construction (model.py:18-47)
    def initialize_models(self):
        a = self.args
        self.G = Generator(a.latent_dim, a.depths[0], a.init_bias_to_zero, a.LReLU_slope, a.apply_pixel_norm,
                           a.generator_last_activation, a.output_dim, a.equalized_lr).to(self.device)
        self.D = Discriminator(a.depths[0], a.init_bias_to_zero, a.LReLU_slope, a.decision_layer_size,
                               a.apply_minibatch_norm, a.input_dim, a.equalized_lr).to(self.device)
        self.G.train()
        self.D.train()

    def set_multi_GPU(self):
        """intended semantics of reference lib/model.py:74-79: replicas start equal"""
        if self.data_parallel:
            broadcast_parameters(self.G)
            broadcast_parameters(self.D)

    def set_optimizers(self):
        """reference lib/model.py:95-97; re-created (moments reset) on every scale jump — SURVEY.md D14"""
        a = self.args
        for attr in ("sync_G", "sync_D"):
            if getattr(self, attr, None) is not None:
                getattr(self, attr).remove()
        for p in list(self.G.parameters()) + list(self.D.parameters()):
            p.grad = None
        gG, gD = FlatGrads(self.G.live_parameters()), FlatGrads(self.D.live_parameters())
        scale = 1.0 / (self.world * self.loss_scale)
        self.opt_G = FusedAdam(gG, lr=a.lr_G, betas=(float(a.beta1), float(a.beta2)), grad_scale=scale,
                               all_params=self.G.parameters())
        self.opt_D = FusedAdam(gD, lr=a.lr_D, betas=(float(a.beta1), float(a.beta2)), grad_scale=scale,
                               all_params=self.D.parameters())
        self.sync_G = GradSync(gG, self.bucket_bytes) if self.data_parallel else None
        self.sync_D = GradSync(gD, self.bucket_bytes) if self.data_parallel else None

    def set_loss_collector(self):
        self._loss_collector = WGANGPLoss(self.args)

    @property
    def loss_collector(self):
        return self._loss_collector

    # ------------------------------------------------------------------ checkpoints (model.py:50-101)
    def save_checkpoint(self, global_step):
        """reference pggan/model.py:50-67: {G,D}_{step}.pt and _latest.pt under {save_root}/{run_id}/ckpt"""
        ckpt_dict = {
            "args": dict(vars(self.args)),
            "global_step": global_step,
            "alpha_G": self.G.alpha,
            "alpha_D": self.D.alpha,
            "alpha_index": self.alpha_index,
            "alpha_jump_value": self.alpha_jump_value,
            "next_alpha_jump_step": self.next_alpha_jump_step,
            "scale_index": self.scale_index,
            "next_scale_jump_step": self.next_scale_jump_step,
        }
        checkpoint.save_checkpoint(self.G, self.opt_G, name="G", ckpt_dict=ckpt_dict)
        checkpoint.save_checkpoint(self.D, self.opt_D, name="D", ckpt_dict=ckpt_dict)

    def load_checkpoint(self):
        """reference pggan/model.py:70-101: schedule variables, add_block replay, new optimizers, weights, moments.
        Works for checkpoints written by this class and by the reference (torch.optim.Adam state)."""
        G_ckpt_dict, D_ckpt_dict = checkpoint.load_checkpoint(self.args, name="G"), \
            checkpoint.load_checkpoint(self.args, name="D")
        for k, v in G_ckpt_dict["args"].items():
            if k not in ("ckpt_id", "ckpt_step", "save_root", "run_id"):
                setattr(self.args, k, v)
        self.global_step = G_ckpt_dict["global_step"]
        self.alpha_index = G_ckpt_dict["alpha_index"]
        self.alpha_jump_value = G_ckpt_dict["alpha_jump_value"]
        self.next_alpha_jump_step = G_ckpt_dict["next_alpha_jump_step"]
        self.scale_index = G_ckpt_dict["scale_index"]
        self.next_scale_jump_step = G_ckpt_dict["next_scale_jump_step"]
        checkpoint.grow_to_checkpoint(self.G, self.D, self.args.depths, self.scale_index)
        self.G.to(self.device)
        self.D.to(self.device)
        self.G.alpha = G_ckpt_dict["alpha_G"]
        self.D.alpha = G_ckpt_dict["alpha_D"]
        self.reset_solver()
        self.G.load_state_dict(G_ckpt_dict["model"], strict=False)
        self.opt_G.load_state_dict(G_ckpt_dict["optimizer"])
        self.D.load_state_dict(D_ckpt_dict["model"], strict=False)
        self.opt_D.load_state_dict(D_ckpt_dict["optimizer"])
        ops.clear_pack_cache()
        self.opt_G.sync_shadows()
        self.opt_D.sync_shadows()
        return self.global_step

    # --------------------------------------------------------------------- schedule (model.py:131-204)
    def reset_solver(self):
        self.set_optimizers()

    def reset_alpha(self, global_step):
        self.G.alpha = 0
        self.D.alpha = 0
        self.alpha_index = 0
        self.next_alpha_jump_step = global_step + self.args.alpha_jump_start[self.scale_index]
        self.alpha_jump_value = 1 / self.args.alpha_jump_Ntimes[self.scale_index]

    def change_scale(self, global_step):
        self.scale_index += 1
        self.next_scale_jump_step += self.args.max_step_at_scale[self.scale_index]
        self.G.add_block(self.args.depths[self.scale_index])
        self.D.add_block(self.args.depths[self.scale_index])
        self.G.to(self.device)
        self.D.to(self.device)
        self.set_multi_GPU()            # new blocks are random per rank: re-broadcast
        self.reset_solver()
        self.reset_alpha(global_step)

    def change_alpha(self, global_step):
        self.alpha_index += 1
        self.G.alpha += self.alpha_jump_value
        self.D.alpha += self.alpha_jump_value
        self.G.alpha = round(self.G.alpha, 4)
        self.D.alpha = round(self.D.alpha, 4)
        if self.alpha_index == self.args.alpha_jump_Ntimes[self.scale_index]:
            self.next_alpha_jump_step = 0
        else:
            self.next_alpha_jump_step = global_step + self.args.alpha_jump_interval[self.scale_index]

    def check_jump(self, global_step):
        if self.next_scale_jump_step == global_step:
            self.change_scale(global_step)
        if self.next_alpha_jump_step == global_step:
            self.change_alpha(global_step)

    def set_scale(self, scale_index: int, alpha: float):
        """test/bench helper: grow to `scale_index` and set alpha directly"""
        while self.scale_index < scale_index:
            self.scale_index += 1
            self.G.add_block(self.args.depths[self.scale_index])
            self.D.add_block(self.args.depths[self.scale_index])
        self.set_multi_GPU()
        self.reset_solver()
        self.G.alpha = alpha
        self.D.alpha = alpha

    # ------------------------------------------------------------------- the step (model.py:206-255)
    def train_step(self, img_real, latent_D=None, latent_G=None, eps=None):
        """img_real: [B,3,R,R] fp32 in [-1,1] on the device.  latent_D/latent_G/eps: optional
        pre-drawn inputs (device tensors); when None they are drawn on the CPU generator and copied,
        exactly like the reference (pggan/model.py:226,244; pggan/loss.py:71-73)."""
        a = self.args
        lc = self.loss_collector
        n_samples = len(img_real)
        nvtx = torch.cuda.nvtx

        # ---- Train D ----
        nvtx.range_push("D forward (real, fake)")
        if self.scale_index:
            img_real = ops.ImageFade.apply(img_real, None, self.D.alpha_ref())      # model.py:217-221
        if latent_D is None:
            latent_D = torch.randn(n_samples, a.latent_dim).to(self.device)
        if self.variant == "W" and n_samples % 4 == 0:
            # The WGAN-GP variant needs no gradient w.r.t. the real images in these two passes, and
            # minibatch-stddev groups (4 consecutive samples) never straddle the real/fake boundary, so
            # D(real) and D(fake) run as ONE pass over 2B images: identical results, larger GEMMs.
            img_real = img_real.detach()
            img_fake = self.G(latent_D).detach()
            pred = self.D(torch.cat([img_real, img_fake], dim=0))
            pred_real, pred_fake = pred[:n_samples], pred[n_samples:]
        else:
            img_real = img_real.detach().requires_grad_()
            pred_real = self.D(img_real)
            img_fake = self.G(latent_D).detach()
            pred_fake = self.D(img_fake)
        D_dict = {"img_real": img_real, "img_fake": img_fake, "pred_real": pred_real, "pred_fake": pred_fake}
        nvtx.range_pop()
        nvtx.range_push("D backward (gradient penalty double backward, loss) + Adam")
        self.opt_D.zero_grad()
        with ops.grad_sink():          # weight / bias gradients are accumulated straight into the flat .grad buffer
            if self.variant == "L":
                loss_D = lc.get_loss_D(D_dict)
            else:
                lc.get_gradient_penalty(D_dict, self.D, backward=True, eps=eps, lazy=True,
                                        loss_scale=self.loss_scale)                  # accumulates into D.grad
                loss_D = lc.get_loss_D_wgangp_terms(D_dict)
            if self.sync_D is not None:
                self.sync_D.arm()
            (loss_D * self.loss_scale if self.loss_scale != 1.0 else loss_D).backward()
        nvtx.range_pop()

        # ---- Train G ----
        nvtx.range_push("G step (G forward, D forward, backward, Adam)")
        if latent_G is None:
            latent_G = torch.randn(n_samples, a.latent_dim).to(self.device)
        # G's forward does not read D: under data parallelism it runs while the tail of D's gradient all-reduce is still
        # on the wire; D's Adam step (which needs the reduced gradients) follows it.  Same arithmetic, same order of
        # parameter updates as pggan/model.py:238-253.
        img_fake = self.G(latent_G)
        if self.sync_D is not None:
            self.sync_D.finish()
        self.opt_D.step()
        d_params = [p for p in self.D.parameters() if p.requires_grad]
        for p in d_params:
            p.requires_grad_(False)
        try:
            pred_fake, _ = self.D(img_fake, True)
            loss_G = lc.get_loss_G({"pred_fake": pred_fake})
            self.opt_G.zero_grad()
            with ops.grad_sink():
                if self.sync_G is not None:
                    self.sync_G.arm()
                (loss_G * self.loss_scale if self.loss_scale != 1.0 else loss_G).backward()
                if self.sync_G is not None:
                    self.sync_G.finish()
        finally:
            for p in d_params:
                p.requires_grad_(True)
        self.opt_G.step()
        nvtx.range_pop()
        return [img_real, img_fake]

    # ------------------------------------------------------------------------- CUDA graph
    def _snapshot(self):
        """everything train_step mutates: parameters, Adam moments and step counters (capture's warm-up restores it)"""
        return ([p.detach().clone() for p in self.G.parameters()], [p.detach().clone() for p in self.D.parameters()],
                [(o.exp_avg.clone(), o.exp_avg_sq.clone(), o.step_count.clone()) for o in (self.opt_G, self.opt_D)])

    def _restore(self, snap):
        g, d, opts = snap
        with torch.no_grad():
            for p, v in zip(self.G.parameters(), g):
                p.copy_(v)
            for p, v in zip(self.D.parameters(), d):
                p.copy_(v)
            for o, (m, v, st) in zip((self.opt_G, self.opt_D), opts):
                o.exp_avg.copy_(m)
                o.exp_avg_sq.copy_(v)
                o.step_count.copy_(st)
        ops.clear_pack_cache()
        self.opt_G.sync_shadows()          # the copy_ above bumped the parameters' versions
        self.opt_D.sync_shadows()

    def capture(self, batch: int, resolution: Optional[int] = None, warmup: int = 3):
        """Capture one whole train_step (3 D + 2 G passes, double backward, both Adam updates, and under data
        parallelism the bucketed all-reduces) into a CUDA graph and return `replay(img, latent_D, latent_G, eps)`,
        which copies the inputs into the static buffers and launches the graph: ~900 kernel launches per step cost no
        CPU time.  alpha is a DEVICE scalar the kernels read at run time (G.alpha / D.alpha are re-uploaded by replay()
        whenever the floats changed), so change_alpha() needs no new capture; change_scale() does (new layers, new
        optimizers).  The warm-up steps run on noise and are UNDONE (parameters, Adam moments and step counters are
        restored), so capturing has no effect on training."""
        a = self.args
        res = resolution or 4 * 2 ** self.scale_index
        dev = self.device
        static = {
            "img": torch.zeros(batch, a.input_dim, res, res, device=dev),
            "z_d": torch.zeros(batch, a.latent_dim, device=dev),
            "z_g": torch.zeros(batch, a.latent_dim, device=dev),
            "eps": torch.zeros(batch, 1, device=dev),
        }
        static["img"].uniform_(-1, 1)
        static["z_d"].normal_()
        static["z_g"].normal_()
        static["eps"].uniform_()
        self.G.alpha_ref()
        self.D.alpha_ref()
        snap = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):           # allocator / lazily-built state / DP contribution counts settle
                self.train_step(static["img"], latent_D=static["z_d"], latent_G=static["z_g"], eps=static["eps"])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore(snap)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.train_step(static["img"], latent_D=static["z_d"], latent_G=static["z_g"], eps=static["eps"])
        # the capture itself recorded (did not run) one step; nothing to undo
        self._graph, self._graph_static, self._graph_out = graph, static, out
        self._graph_scale = self.scale_index

        def replay(img=None, latent_D=None, latent_G=None, eps=None):
            if self.scale_index != self._graph_scale:
                raise RuntimeError("the networks grew since this graph was captured: capture() again after change_scale")
            for key, val in (("img", img), ("z_d", latent_D), ("z_g", latent_G), ("eps", eps)):
                if val is not None:
                    static[key].copy_(val, non_blocking=True)
            self.G.alpha_ref()                # upload alpha if check_jump() changed it (one 4-byte fill, else nothing)
            self.D.alpha_ref()
            graph.replay()
            return out

        return replay

    def validation(self):
        pass
