"""VQVAE LightningModule surface (reference: vqvae/model.py) on top of the libvqgan_b200 kernels.

Same constructor (`VQVAE(image_size, ae_conf, q_conf, l_conf, t_conf, init_cb, load_loss)`, model.py:25-26), same
sub-module / parameter / buffer names, same construction order (quantizer -> encoder -> decoder -> criterion ->
init_codebook, so seeded initialisation matches the reference), same YAML-derived dict schema, same hooks and
inference API.  Reference defects are handled as stated in each method's docstring.
"""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import nn

from . import ops
from .lightning_shim import LightningModule
from .modules.abstract_modules.base_autoencoder import BaseVQVAE
from .modules.autoencoder import Conv2d, Decoder, Encoder, GroupNorm
from .modules.loss.loss import VQLPIPS, VQLPIPSWithDiscriminator
from .modules.vector_quantizers import EMAVectorQuantizer, VectorQuantizer
from .optim import FusedAdamW
from .schedulers import CosineScheduler, LinearCosineScheduler, LinearScheduler

try:
    from .modules.vector_quantizers import EntropyVectorQuantizer, GumbelVectorQuantizer
except ImportError:          # pragma: no cover
    EntropyVectorQuantizer = GumbelVectorQuantizer = None


class MSELoss(nn.Module):
    """torch.nn.MSELoss() replacement used when l_conf is None (model.py:136-137): mean((a-b)^2) by vqb_diff_sums."""

    def forward(self, recon: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return ops.mse_l1(recon, target)[0]


class VQVAE(BaseVQVAE, LightningModule):

    def __init__(self, image_size: int, ae_conf: dict, q_conf: dict, l_conf: Optional[dict], t_conf: Optional[dict],
                 init_cb: bool = True, load_loss: bool = True, fix_param_groups: bool = False, pretrained_lpips: bool = True):
        """Arguments as in the reference (model.py:25-78).  `fix_param_groups=True` repairs defect B2 (53 encoder
        tensors never reach the optimizer because of relative-name collisions); the default replicates it.
        `pretrained_lpips=False` keeps the seeded random LPIPS weights (offline parity / benchmark runs)."""
        LightningModule.__init__(self)
        BaseVQVAE.__init__(self, image_size=image_size)
        self.t_conf = t_conf
        self.fix_param_groups = fix_param_groups

        self.cb_size = q_conf['num_embeddings']
        self.latent_dim = q_conf['embedding_dim']
        self.reinit_every_n_epochs = q_conf.get('reinit_every_n_epochs')
        qtype, qp = q_conf['type'], q_conf.get('params') or {}
        self.kl_warmup_epochs = self.temp_decay_epochs = self.temp_final = None
        if qtype == 'standard':
            self.quantizer = VectorQuantizer(self.cb_size, self.latent_dim, float(qp['commitment_cost']))
        elif qtype == 'ema':
            self.quantizer = EMAVectorQuantizer(self.cb_size, self.latent_dim, float(qp['commitment_cost']),
                                                float(qp['decay']), float(qp['epsilon']))
        elif qtype == 'gumbel':
            if GumbelVectorQuantizer is None:
                raise NotImplementedError('gumbel quantizer kernels are not built yet')
            self.quantizer = GumbelVectorQuantizer(self.cb_size, self.latent_dim, bool(qp['straight_through']),
                                                   float(qp['temp']), float(qp['kl_cost']))
            self.kl_warmup_epochs = qp.get('kl_warmup_epochs')
            self.temp_decay_epochs = qp.get('temp_decay_epochs')
            self.temp_final = qp.get('temp_final')
        elif qtype == 'entropy':
            if EntropyVectorQuantizer is None:
                raise NotImplementedError('entropy quantizer kernels are not built yet')
            self.quantizer = EntropyVectorQuantizer(self.cb_size, self.latent_dim, float(qp['ent_loss_ratio']),
                                                    float(qp['ent_temperature']), str(qp['ent_loss_type']),
                                                    float(qp['commitment_cost']))
        else:
            raise ValueError(f'unrecognized quantizer: {q_conf["type"]}')

        channels = ae_conf['channels']
        num_res_blocks = ae_conf['num_res_blocks']
        channel_multipliers = tuple(ae_conf['channel_multipliers'])
        final_conv_channels = self.cb_size if qtype == 'gumbel' else self.latent_dim
        self.encoder = Encoder(channels, num_res_blocks, channel_multipliers, final_conv_channels)
        self.decoder = Decoder(channels, num_res_blocks, channel_multipliers, self.latent_dim)

        if load_loss:
            if l_conf is None:
                self.criterion = MSELoss()
            elif l_conf.get('adversarial_params') is None:
                self.criterion = VQLPIPS(l_conf['l1_weight'], l_conf['l2_weight'], l_conf['perc_weight'],
                                         net_type=l_conf.get('lpips_net', 'alex'), pretrained_lpips=pretrained_lpips)
            else:
                self.criterion = VQLPIPSWithDiscriminator(image_size, l_conf['l1_weight'], l_conf['l2_weight'],
                                                          l_conf['perc_weight'], l_conf['adversarial_params'],
                                                          pretrained_lpips=pretrained_lpips)
        else:
            self.criterion = None

        if init_cb:
            self.quantizer.init_codebook()

    # ---------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor):
        """x [B,3,H,W] in [-1,1] -> (reconstructions [B,3,H,W], quantizer loss, used indices [B,S])  (model.py:151-161)"""
        z = self.encoder(x)
        quantized, used_indices, e_loss = self.quantizer(z)
        x_recon = self.decoder(quantized)
        return x_recon, e_loss, used_indices

    # ---- schedules (model.py:163-230) --------------------------------------------------------------------
    def on_train_start(self):
        lr = float(self.t_conf['lr'])
        nb = self.trainer.num_training_batches
        wu, dc = self.t_conf.get('warmup_epochs'), self.t_conf.get('decay_epochs')
        if wu is not None and dc is not None:
            self.scheduler = LinearCosineScheduler(0, dc * nb, lr, lr / 2., wu * nb)
        elif wu is not None:
            self.scheduler = LinearScheduler(0, wu * nb, 1e-20, lr)
        elif dc is not None:
            self.scheduler = CosineScheduler(0, dc * nb, lr, lr / 2.)
        if GumbelVectorQuantizer is not None and isinstance(self.quantizer, GumbelVectorQuantizer):
            temp, kl = self.quantizer.get_consts()
            if self.kl_warmup_epochs is not None:
                self.quantizer.kl_warmup = CosineScheduler(0, int(self.kl_warmup_epochs * nb), 0.0, kl)
            if self.temp_decay_epochs is not None and self.temp_final is not None:
                self.quantizer.temp_decay = CosineScheduler(0, int(self.temp_decay_epochs * nb), temp, self.temp_final)

    def on_train_batch_start(self, _: Any, batch_index: int):
        current_step = (self.current_epoch * self.trainer.num_training_batches) + batch_index
        step_lr = self.scheduler.step(current_step) if self.scheduler is not None else float(self.t_conf['lr'])
        for optimizer in self.trainer.optimizers:
            for g in optimizer.param_groups:
                g['lr'] = step_lr
        if GumbelVectorQuantizer is not None and isinstance(self.quantizer, GumbelVectorQuantizer):
            this_temp, this_kl = self.quantizer.get_consts()
            if self.quantizer.kl_warmup is not None:
                this_kl = self.quantizer.kl_warmup.step(current_step)
            if self.quantizer.temp_decay is not None:
                this_temp = self.quantizer.temp_decay.step(current_step)
            self.quantizer.set_consts(this_temp, this_kl)
        else:
            this_temp, this_kl = 0.0, 0.0
        self.log('gumbel_quantizer/temperature', this_temp, sync_dist=True)
        self.log('gumbel_quantizer/kl_constant', this_kl, sync_dist=True)

    # ---- the hot loop body (model.py:232-295) ---------------------------------------------------------
    def training_step(self, batch: Any, batch_index: int):
        """model.py:232-295.  Branch A = VQGAN (two optimizers, manual optimisation), B = LPIPS only, C = plain MSE.
        Defect B1 (the reference returns an unbound `loss` outside branch A) is fixed by returning ae_loss; logged
        scalars stay on the device (the reference does 7 .item() host syncs per step)."""
        images = self.preprocess_batch(batch[0] if isinstance(batch, tuple) else batch, training=True)
        x_recon, q_loss, used_indices = self.forward(images)
        zero = torch.zeros(1, device=images.device)
        if isinstance(self.criterion, VQLPIPSWithDiscriminator):
            ae_opt, disc_opt = self.optimizers()
            ae_opt.zero_grad()
            res = self.criterion.forward_autoencoder(q_loss, images, x_recon, self.current_epoch,
                                                     last_layer=self.decoder.conv_out.weight)
            ae_loss, l1_loss, l2_loss, p_loss, g_loss, g_weight = res
            self.manual_backward(ae_loss, ae_opt)
            ae_opt.step()
            step = (self.current_epoch * self.trainer.num_training_batches) + batch_index
            loss, d_loss, r1_penalty = self.criterion.forward_discriminator(images, x_recon, self.current_epoch, step)
            if loss is not None:
                disc_opt.zero_grad()
                self.manual_backward(loss, disc_opt)
                disc_opt.step()
        elif isinstance(self.criterion, VQLPIPS):
            ae_loss, l1_loss, l2_loss, p_loss = self.criterion(q_loss, images, x_recon)
            g_loss, d_loss, g_weight, r1_penalty = zero, zero, 0., 0.
        else:
            l2_loss = self.criterion(x_recon, images)
            l1_loss, g_loss, p_loss, d_loss, g_weight, r1_penalty = zero, zero, zero, zero, 0., 0.
            ae_loss = q_loss + l2_loss
        self.log('g_weight', g_weight)
        self.log('r1_penalty', r1_penalty.detach() if torch.is_tensor(r1_penalty) else r1_penalty)
        self.log('train/loss', ae_loss.detach())
        self.log('train/l1_loss', l1_loss.detach())
        self.log('train/l2_loss', l2_loss.detach())
        self.log('train/quant_loss', q_loss.detach())
        self.log('train/perc_loss', p_loss.detach())
        self.log('train/gen_loss', g_loss.detach())
        self.log('train/disc_loss', d_loss.detach())
        # per-batch code usage (model.py:289-293; defect B3 -- only the last batch is kept -- replicated)
        self.train_epoch_usage_count = ops.code_histogram(used_indices, self.cb_size)
        return ae_loss

    def graph_variant(self, batch_index: int):
        """What makes two training steps DIFFERENT sequences of kernel launches (the CUDA-graph trainer keeps one captured graph
        per value): whether the discriminator has started and whether this step carries the R1 penalty (loss.py:144-150)."""
        c = self.criterion
        if isinstance(c, VQLPIPSWithDiscriminator):
            started = self.current_epoch >= c.adversarial_start_epoch
            step = (self.current_epoch * self.trainer.num_training_batches) + batch_index
            r1 = started and c.r1_regularization_cost is not None and step % c.r1_regularization_every == 0
            return (started, r1)
        return ()

    def on_train_epoch_end(self):
        if (self.reinit_every_n_epochs is not None and self.current_epoch % self.reinit_every_n_epochs == 0
                and self.current_epoch > 0):
            # under data parallelism the usage counts are summed over ranks and rank 0's draw is broadcast (base_quantizer.py)
            self.quantizer.reinit_unused_codes(self.quantizer.get_codebook_usage(
                self.quantizer.reduce_usage(self.train_epoch_usage_count.float()))[0])
        self.train_epoch_usage_count = None

    def on_train_end(self):
        if self.scheduler is not None:          # defect B10 guarded
            self.scheduler.destroy()

    @torch.no_grad()
    def validation_step(self, batch: Any, batch_index: int):
        """model.py:309-356: the three criterion branches of training_step without the optimizers -- VQGAN (forward_autoencoder +
        forward_discriminator; no adaptive weight and no R1 outside training, loss.py:127,147), LPIPS only, plain MSE -- and the
        seven validation/* scalars; usage counts are kept for on_validation_epoch_end (the reference's `else + used_indices`
        keeps the LAST batch only, defect B3, replicated)."""
        images = self.preprocess_batch(batch[0] if isinstance(batch, tuple) else batch)
        x_recon, q_loss, used_indices = self.forward(images)
        zero = torch.zeros(1, device=images.device)
        if isinstance(self.criterion, VQLPIPSWithDiscriminator):
            res = self.criterion.forward_autoencoder(q_loss, images, x_recon, self.current_epoch,
                                                     last_layer=self.decoder.conv_out.weight)
            loss, l1_loss, l2_loss, p_loss, g_loss, _ = res
            step = (self.current_epoch * self.trainer.num_training_batches) + batch_index
            _, d_loss, _ = self.criterion.forward_discriminator(images, x_recon, self.current_epoch, step)
        elif isinstance(self.criterion, VQLPIPS):
            loss, l1_loss, l2_loss, p_loss = self.criterion(q_loss, images, x_recon)
            g_loss, d_loss = zero, zero
        else:
            l2_loss = self.criterion(x_recon, images)
            l1_loss, g_loss, p_loss, d_loss = zero, zero, zero, zero
            loss = q_loss + l2_loss
        self.log('validation/loss', loss)
        self.log('validation/l1_loss', l1_loss)
        self.log('validation/l2_loss', l2_loss)
        self.log('validation/quant_loss', q_loss)
        self.log('validation/perc_loss', p_loss)
        self.log('validation/gen_loss', g_loss)
        self.log('validation/disc_loss', d_loss)
        self.val_epoch_usage_count = torch.bincount(used_indices.view(-1), minlength=self.cb_size)
        return loss

    def on_validation_epoch_end(self):
        if self.val_epoch_usage_count is not None:
            _, perplexity, cb_usage = self.quantizer.get_codebook_usage(self.val_epoch_usage_count.float())
            self.log('val_metrics/used_codebook', cb_usage)
            self.log('val_metrics/perplexity', perplexity)
        self.val_epoch_usage_count = None

    # ---- test-time evaluation (model.py:491-562) --------------------------------------------------------------
    def on_test_epoch_start(self):
        """MSE and PSNR are accumulated with the library's own reduction kernel (vqb_diff_sums, fp64 sums), following
        torchmetrics' definitions (MeanSquaredError: sum of squared errors / elements; PeakSignalNoiseRatio with
        data_range=None: 10 log10((max - min of all targets)^2 / MSE)).  SSIM is the library's own kernel (vqb_ssim_sums:
        StructuralSimilarityIndexMeasure() defaults -- Gaussian 11 x 11 window, sigma 1.5, per-batch data range, mean over images;
        torchmetrics' published algorithm, un-pinned: the package is absent here).  rFID needs torchmetrics AND its pretrained
        Inception weights: it is delegated exactly as in the reference when both are available, otherwise not logged.
        Codebook usage ACCUMULATES over the epoch (the reference's `else + used_indices` keeps the last batch only)."""
        dev = next(self.parameters()).device
        self._test_sse = torch.zeros((), dtype=torch.float64, device=dev)
        self._test_elems = 0
        self._test_min = torch.zeros((), device=dev)          # torchmetrics PeakSignalNoiseRatio(data_range=None) starts both at 0
        self._test_max = torch.zeros((), device=dev)
        self.test_usage_count = None
        self._test_ssim_sum = torch.zeros((), dtype=torch.float64, device=dev)
        self._test_images = 0
        self.test_rfid = None
        try:
            from torchmetrics.image.fid import FrechetInceptionDistance
            self.test_rfid = FrechetInceptionDistance().to(dev)
        except Exception as e:           # not installed (or its Inception weights are not downloadable)
            import warnings
            warnings.warn(f'torchmetrics rFID unavailable ({type(e).__name__}: {e}); MSE, PSNR, SSIM and codebook usage are logged')
            self.test_rfid = None

    @torch.no_grad()
    def test_step(self, images: Any, _: int = 0):
        images = images[0] if isinstance(images, tuple) else images
        reconstructions, _q, used_indices = self.forward(self.preprocess_batch(images))
        reconstructions = self.preprocess_visualization(reconstructions)              # NCHW fp32 in [0,1]
        target = images.float() / 255.0 if images.dtype == torch.uint8 else images.float()
        target = target.contiguous()
        counts = torch.bincount(used_indices.view(-1), minlength=self.cb_size)
        self.test_usage_count = counts if self.test_usage_count is None else self.test_usage_count + counts
        self._test_sse += ops.diff_sums(reconstructions, target)[0]                   # the kernel's fp64 sum of squared errors
        self._test_elems += target.numel()
        lo, hi = torch.aminmax(target)
        self._test_min = torch.minimum(self._test_min, lo)
        self._test_max = torch.maximum(self._test_max, hi)
        self._test_ssim_sum += ops.ssim_per_image(reconstructions, target).sum()
        self._test_images += target.shape[0]
        if self.test_rfid is not None:
            # torchvision ConvertImageDtype(torch.uint8) on float images (model.py:543-545): x * (255 + 1 - 1e-3), truncated
            to_u8 = lambda t: t.mul(255.999).to(torch.uint8)
            self.test_rfid.update(to_u8(reconstructions), real=False)
            self.test_rfid.update(to_u8(target), real=True)

    def on_test_epoch_end(self):
        mse = (self._test_sse / max(self._test_elems, 1)).float()
        self.log('mse', mse)
        data_range = (self._test_max - self._test_min).double()
        self.log('psnr', (10.0 * torch.log10(data_range * data_range / self._test_sse * max(self._test_elems, 1))).float())
        self.log('ssim', (self._test_ssim_sum / max(self._test_images, 1)).float())
        if self.test_rfid is not None:
            self.log('rfid', self.test_rfid.compute())
        if self.test_usage_count is not None:
            _, perplexity, cb_usage = self.quantizer.get_codebook_usage(self.test_usage_count.float())
            self.log('used_codebook', cb_usage)
            self.log('perplexity', perplexity)

    # ---- optimizer (model.py:372-440) -----------------------------------------------------------------
    def configure_optimizers(self):
        """AdamW with two groups (Conv2d weights decay; biases, Embedding and GroupNorm weights do not), fused over flat
        buffers.  Defect B2: the reference keys its parameter dict by names RELATIVE to encoder / decoder / quantizer, so
        a decoder tensor silently replaces the encoder tensor of the same relative name; replicated unless
        fix_param_groups=True."""
        lr = float(self.t_conf['lr'])
        betas = [float(b) for b in self.t_conf['betas']]
        eps = float(self.t_conf['eps'])
        weight_decay = float(self.t_conf['weight_decay'])

        decay, no_decay, param_dict = set(), set(), {}
        for prefix, sub in (('encoder', self.encoder), ('decoder', self.decoder), ('quantizer', self.quantizer)):
            rename = (lambda n: f'{prefix}.{n}') if self.fix_param_groups else (lambda n: n)
            for mn, m in sub.named_modules():
                for pn, _ in m.named_parameters():
                    fpn = rename('%s.%s' % (mn, pn) if mn else pn)
                    if pn.endswith('bias'):
                        no_decay.add(fpn)
                    elif pn.endswith('weight') and isinstance(m, nn.Conv2d):
                        decay.add(fpn)
                    elif pn.endswith('weight') and isinstance(m, (nn.Embedding, GroupNorm)):
                        no_decay.add(fpn)
            for pn, p in sub.named_parameters():
                param_dict[rename(pn)] = p          # later sub-modules overwrite equal relative names (B2)
        assert len(decay & no_decay) == 0
        assert len(param_dict.keys() - (decay | no_decay)) == 0
        groups = [
            {'params': [param_dict[pn] for pn in sorted(decay)], 'weight_decay': weight_decay},
            {'params': [param_dict[pn] for pn in sorted(no_decay)], 'weight_decay': 0.0},
        ]
        # flat-buffer layout = expected order of gradient completion (reverse of construction = reverse of the forward pass):
        # decoder.conv_out first, encoder.conv_in last, so that gradient buckets can be all-reduced during backward
        rank = {id(p): i for i, p in enumerate(reversed(list(self.parameters())))}
        ae_optimizer = FusedAdamW(groups, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, layout_rank=rank)
        if isinstance(self.criterion, VQLPIPSWithDiscriminator):
            # every discriminator tensor decays (model.py:431-433); manual optimisation with two optimizers (:436-438)
            disc_optimizer = FusedAdamW(list(self.criterion.discriminator.parameters()), lr=lr, betas=betas, eps=eps,
                                        weight_decay=weight_decay, layout_rank=rank)
            self.automatic_optimization = False
            return [ae_optimizer, disc_optimizer], []
        return ae_optimizer

    # ---- reference checkpoints (vqvae/train.py:106-111, vqvae/evaluate.py:48-49) ---------------------------------
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict: bool = True, **kwargs):
        """pl.LightningModule.load_from_checkpoint for the reference's Lightning `.ckpt` files: a torch pickle holding
        `state_dict` (keys `encoder.* / decoder.* / quantizer.* [/ criterion.*]`, the names and shapes this package keeps) and,
        when the reference saved them, `hyper_parameters`.  Constructor arguments come from **kwargs exactly as in the reference
        call sites (`strict=False, image_size=..., ae_conf=..., q_conf=..., l_conf=..., t_conf=..., init_cb=False,
        load_loss=False`), falling back to the checkpoint's hyper_parameters.  With strict=False, keys of sub-modules that were
        not built (e.g. `criterion.*` under load_loss=False) are ignored, as Lightning does."""
        ckpt = torch.load(checkpoint_path, map_location=map_location or 'cpu', weights_only=False)
        hp = dict(ckpt.get('hyper_parameters') or {})
        hp.update(kwargs)
        missing = [k for k in ('image_size', 'ae_conf', 'q_conf') if k not in hp]
        if missing:
            raise ValueError(f'load_from_checkpoint: missing constructor arguments {missing} (not in the checkpoint either)')
        hp.setdefault('l_conf', None)
        hp.setdefault('t_conf', None)
        hp.setdefault('init_cb', False)
        model = cls(**hp)
        state = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
        result = model.load_state_dict(state, strict=strict)
        if not strict and result.missing_keys:
            import warnings
            warnings.warn(f'load_from_checkpoint: {len(result.missing_keys)} tensors keep their initial values '
                          f'(first: {result.missing_keys[0]})')
        model._checkpoint_extras = {k: ckpt[k] for k in ('epoch', 'global_step', 'optimizer_states') if k in ckpt}
        if map_location is not None and map_location != 'cpu':
            model = model.to(map_location)
        ops.bump_weights_epoch()
        return model

    # ---- two-stage-model API (model.py:458-489) -------------------------------------------------------------
    @torch.no_grad()
    def get_tokens(self, images: torch.Tensor) -> torch.Tensor:
        """images [B,3,H,W] in [0,1] -> codebook indices [B,S]"""
        return self.quantizer.vec_to_codes(self.encoder(self.preprocess_batch(images)))

    @torch.no_grad()
    def quantize(self, images: torch.Tensor) -> torch.Tensor:
        """images [B,3,H,W] in [0,1] -> quantized latents [B,S,D]"""
        q = self.quantizer(self.encoder(self.preprocess_batch(images)))[0]          # [B,D,h,w], physically NHWC
        b, d, h, w = q.shape
        return q.permute(0, 2, 3, 1).reshape(b, h * w, d)

    @torch.no_grad()
    def reconstruct(self, images: torch.Tensor) -> torch.Tensor:
        """images [B,3,H,W] in [0,1] -> reconstructions [B,3,H,W] in [0,1]"""
        return self.preprocess_visualization(self(self.preprocess_batch(images))[0])

    @torch.no_grad()
    def reconstruct_from_tokens(self, tokens: torch.Tensor) -> torch.Tensor:
        """tokens [B,S] -> images [B,3,H,W] in [0,1].  Defect B4 (the reference feeds [B,S,D] to a Conv2d decoder)
        is fixed: S must be a square h*w and the vectors are laid out as a [B,D,h,w] latent."""
        b, s = tokens.shape
        h = int(round(s ** 0.5))
        if h * h != s:
            raise ValueError('token sequence length must be a perfect square')
        vec = self.quantizer.codes_to_vec(tokens)                                     # [B,S,D]
        latent = vec.reshape(b, h, h, self.latent_dim).permute(0, 3, 1, 2)          # logical NCHW, physical NHWC
        return self.preprocess_visualization(self.decoder(latent))
