"""`Diffuser` and `EDM` — drop-in for the hot-path half of src/tinyedm/edm.py.

Kept: constructor keywords (mirrored as attributes), `training_step(batch, batch_idx)`,
`validation_step`, `forward(noisy_image, sigma, class_label)`, `predict_step`, `configure_optimizers`,
`get_lr_scheduler`, the `conditional` / `num_classes` properties and the parameter registration order
(diffuser, embedding, denoiser, u, train_mse, val_mse — edm.py:128-151).
Changed underneath: Diffuser is one fused kernel (noise scaling + add, edm.py:84-93); the loss is one fused
reduction incl. lambda(sigma), exp(-u) and mean(u) (edm.py:212-219, metric.py:8-18); the optimiser is the fused
Adam(+EMA) kernel of tinyedm_b200.optim. When `lightning` is installed EDM derives from
`lightning.LightningModule` exactly like the reference; without it (this image) it derives from `nn.Module` and
`self.log` / `self.lr_schedulers` degrade to no-ops so the same step code runs under any plain training loop.
`load_from_checkpoint` / `save_config` read and write the reference's checkpoint layout (tinyedm_b200/utils.py); the
Lightning callback wiring around them is out of scope.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .metric import WeightedMeanSquaredError
from .networks import UncertaintyNet
from .ops import F32

try:  # pragma: no cover - lightning is not part of this image
    import lightning as L
    _Base = L.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _Base = nn.Module
    HAVE_LIGHTNING = False


class Diffuser(nn.Module):
    """ln(sigma) ~ N(P_mean, P_std); returns (clean + sigma * n, sigma) — edm.py:64-96.

    Both normal draws are made inside ONE kernel (tedm_diffuse_philox, counter-based Philox4x32-10 keyed by
    (seed, step, element)) instead of two torch Philox launches plus three elementwise launches. `seed` is taken from
    torch's default generator at first use (so `torch.manual_seed` makes a run reproducible, as with the reference's
    `torch.randn`), XORed with the process rank; `step` counts the calls. Both live in a device tensor the kernel reads, so
    inside a captured CUDA graph every replay draws fresh noise and `seed()` reaches the replays. When an `EDM` owns this diffuser, the same kernel also emits the
    Denoiser's input block (c_in * noisy, ones channel, 3x3 patch gather: networks.py:578-587) and hands it over on the
    returned tensor (`noisy._tedm_xcol`), so the image never makes a second trip through HBM before conv_in.
    """

    def __init__(self, P_mean: float, P_std: float) -> None:
        super().__init__()
        self.P_mean = P_mean
        self.P_std = P_std
        self._rng: Tensor | None = None                # device int64 {seed, step}: read by the kernel, so graph replays follow it
        self._pending: tuple[int, int] | None = None   # seed() before the first CUDA call
        self._fuse_sigma_data: float | None = None     # set by EDM: the sigma_data of the denoiser that consumes `noisy`

    def seed(self, seed: int, step: int = 0) -> None:
        """Explicit (seed, step) of the draw stream; `step` counts the calls made so far."""
        if self._rng is not None:
            self._rng.copy_(torch.tensor([int(seed) & (2 ** 63 - 1), int(step)], dtype=torch.int64))
        else:
            self._pending = (int(seed) & (2 ** 63 - 1), int(step))

    def rng_state(self) -> tuple[int, int]:
        """(seed, calls made so far)."""
        if self._rng is None:
            return self._pending if self._pending is not None else (0, 0)
        s = self._rng.tolist()
        return int(s[0]), int(s[1])

    def _state(self, device) -> Tensor:
        if self._rng is None or self._rng.device != device:
            if self._rng is not None:
                seed, step = self.rng_state()
            elif self._pending is not None:
                seed, step = self._pending
            else:
                rank = 0
                if torch.distributed.is_available() and torch.distributed.is_initialized():
                    rank = torch.distributed.get_rank()
                seed = (int(torch.randint(0, 2 ** 62, (1,)).item()) ^ (rank * 0x9E3779B97F4A7C15)) & (2 ** 63 - 1)
                step = 0
            self._rng = torch.tensor([seed, step], dtype=torch.int64).to(device)
        return self._rng

    @torch.no_grad()
    def forward(self, clean_image: Tensor) -> tuple[Tensor, Tensor]:
        if not clean_image.is_cuda:
            raise RuntimeError("tinyedm_b200.Diffuser runs on CUDA (sm_100a) only; there is no CPU fallback")
        ops.ensure_device(clean_image.device)
        clean = ops.check(clean_image.float().contiguous(), F32, "clean_image")
        if clean.dim() != 4:
            raise RuntimeError("tinyedm_b200.Diffuser expects (B, C, H, W) images")
        state = self._state(clean.device)
        state[1:2] += 1
        sd = self._fuse_sigma_data if 9 * (clean.shape[1] + 1) <= 64 else None
        noisy, sigma, xcol = ops.diffuse_philox(clean, state, float(self.P_mean), float(self.P_std), sd)
        if xcol is not None:
            noisy._tedm_xcol = (xcol, float(sd), sigma.data_ptr())    # picked up by DenoiserEngine.forward
        return noisy, sigma

    def draws(self, batch: int, n: int, device) -> tuple[Tensor, Tensor]:
        """(epsilon (B,), noise (B, n)) of the MOST RECENT call, regenerated from (seed, step): for tests and debugging."""
        return ops.philox_normal_draws(self._state(torch.device(device)), batch, n)

    def extra_repr(self) -> str:
        return f"P_mean={self.P_mean}, P_std={self.P_std}"


class EDM(_Base):
    def __init__(self, *, diffuser, embedding, denoiser, use_ema: bool, use_uncertainty: bool, steady_steps: int,
                 rampup_steps: int, scheduler_interval: str, sigma_data: float | None = None, lr: float = 1e-4,
                 betas: tuple[float, float] = (0.9, 0.999), ema_length: float | None = None,
                 validate_original_weights: bool = False, every_n_steps: int = 1, cpu_offload: bool = False) -> None:
        super().__init__()
        assert getattr(embedding, "fourier_dim", None) is not None, "Embedding must have an fourier_dim attribute."
        if use_ema and ema_length is None:
            raise ValueError("ema_length must be specified when use_ema is True.")
        self.diffuser = diffuser
        self.embedding = embedding
        self.denoiser = denoiser
        self.use_ema = use_ema
        self.use_uncertainty = use_uncertainty
        self.steady_steps = steady_steps
        self.rampup_steps = rampup_steps
        self.scheduler_interval = scheduler_interval
        self.betas = betas
        self.ema_length = ema_length
        self.validate_original_weights = validate_original_weights
        self.every_n_steps = every_n_steps
        self.cpu_offload = cpu_offload
        self.u = UncertaintyNet(embedding.fourier_dim, embedding.fourier_dim) if use_uncertainty else None
        self.sigma_data = sigma_data if sigma_data is not None else denoiser.sigma_data
        self.lr = lr
        self.train_mse = WeightedMeanSquaredError()
        self.val_mse = WeightedMeanSquaredError()
        self.solver = None
        if isinstance(diffuser, Diffuser) and hasattr(denoiser, "sigma_data"):
            diffuser._fuse_sigma_data = float(denoiser.sigma_data)      # the input block rides in the diffuser's kernel
        if HAVE_LIGHTNING:  # pragma: no cover - the reference's checkpoints carry the deinstantiate tree (edm.py:154-157)
            self.hparams.update(self.save_config())

    # ---- Lightning shims (no-ops without Lightning) ----
    if not HAVE_LIGHTNING:
        def log(self, *args, **kwargs) -> None:  # noqa: D401
            return None

        def lr_schedulers(self):
            return None

    def _log_metric(self, name: str, metric, **kw) -> None:
        """The reference logs the torchmetrics object itself; this metric is a plain module, so log its running value."""
        if HAVE_LIGHTNING:  # pragma: no cover
            self.log(name, metric.compute().reshape(()), **kw)

    def _log_lr(self) -> None:
        sched = self.lr_schedulers()
        if sched is not None:
            self.log("learning_rate", sched.get_last_lr()[0])

    # ---- the hot path ----
    def training_step(self, batch, batch_idx):
        """edm.py:205-236."""
        clean_image, class_label = batch
        class_label = class_label if self.conditional else None
        noisy_image, sigma = self.diffuser(clean_image)
        fourier_embedding, embedding = self.embedding(sigma, class_label)
        denoised_image = self.denoiser(noisy_image, sigma, embedding)
        uncertainty = self.u(fourier_embedding).flatten() if self.u is not None else None
        loss = self.train_mse.edm_loss(denoised_image, clean_image, sigma, self.sigma_data, uncertainty)
        self._log_metric("train_loss", self.train_mse, prog_bar=True)
        if uncertainty is not None:
            self.log("uncertainty", uncertainty.detach().mean())
        self._log_lr()
        return loss

    def validation_step(self, batch, batch_idx):
        """edm.py:238-248."""
        clean_image, class_label = batch
        class_label = class_label if self.conditional else None
        noisy_image, sigma = self.diffuser(clean_image)
        _, embedding = self.embedding(sigma, class_label)
        denoised_image = self.denoiser(noisy_image, sigma, embedding)
        loss = self.val_mse.edm_loss(denoised_image, clean_image, sigma, self.sigma_data)
        self._log_metric("val_loss", self.val_mse)
        return loss

    def forward(self, noisy_image: Tensor, sigma: Tensor, class_label: Tensor | None = None) -> Tensor:
        """edm.py:280-286 — what the sampler calls."""
        class_label = class_label if self.conditional else None
        _, embedding = self.embedding(sigma, class_label)
        return self.denoiser(noisy_image, sigma, embedding)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, *, map_location=None, load_ema: bool = False, **kwargs):
        """edm.py:159-194: rebuilds the module from the checkpoint's `hyper_parameters` tree and loads its weights
        (or the EMA weights kept in `optimizer_states[0]["ema"]`). Reads the reference's own checkpoints."""
        from .utils import load_reference_checkpoint
        return load_reference_checkpoint(checkpoint_path, load_ema=load_ema, map_location=map_location)

    def save_config(self) -> dict:
        """edm.py:154-157: the `deinstantiate` tree of this module (what the reference stores as `hyper_parameters`)."""
        from .utils import deinstantiate
        return deinstantiate(self)

    def invalidate_weights(self) -> None:
        """Call after writing parameters through `.data` (which bumps no version counter): the cached normalised weights
        of the embedding, the denoiser and the uncertainty head are rebuilt at the next forward."""
        from .engine import bump_weights_epoch
        bump_weights_epoch()

    def prepare_weights(self, device) -> None:
        """Refreshes the cached normalised weights of the embedding and the denoiser (no-op while they are current);
        lets `DeterministicSolver` replay its CUDA graph after the parameters were updated or swapped with the EMA."""
        self.embedding.prepare_weights(device)
        self.denoiser.prepare_weights(device)

    def predict_step(self, batch, batch_idx: int, dataloader_idx: int | None = None):
        """edm.py:288-295."""
        x0, class_label = batch
        class_label = class_label if self.conditional else None
        return self.solver.solve(self, x0, class_label)

    @property
    def num_classes(self) -> int | None:
        return self.embedding.num_classes

    @property
    def conditional(self) -> bool:
        return self.num_classes is not None

    # ---- optimiser (edm.py:250-266, :305-317) ----
    def configure_optimizers(self):
        from .optim import FusedAdamEMA
        optimizer = FusedAdamEMA(self.parameters(), lr=self.lr, betas=self.betas,
                                 ema_length=self.ema_length if self.use_ema else None, every_n_steps=self.every_n_steps)
        scheduler = self.get_lr_scheduler(optimizer, self.rampup_steps, self.steady_steps)
        return {"optimizer": optimizer,
                "lr_scheduler": {"scheduler": scheduler, "interval": self.scheduler_interval, "frequency": 1}}

    @staticmethod
    def lr_factor(step: int, rampup_steps: int, steady_steps: int) -> float:
        """Linear ramp-up from 1e-8, plateau, then 1/sqrt decay (edm.py:306-317)."""
        if step < rampup_steps:
            return 1e-8 + (1.0 - 1e-8) * step / rampup_steps
        if step < rampup_steps + steady_steps:
            return 1.0
        return float(1 / np.sqrt(1 + (step - rampup_steps - steady_steps) / steady_steps))

    @staticmethod
    def get_lr_scheduler(optimizer, rampup_steps, steady_steps):
        return torch.optim.lr_scheduler.LambdaLR(optimizer, lambda s: EDM.lr_factor(s, rampup_steps, steady_steps))
