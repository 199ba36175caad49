"""Architecture constants of the reference's shipped configs (experiments/conf/{cifar10,mnist}.yaml:20-56) as plain
keyword dicts for `Denoiser(**...)` / `Embedding(**...)`. The reference builds these through Hydra `_target_` trees."""

CIFAR10 = dict(
    denoiser=dict(
        in_channels=3, out_channels=3, sigma_data=0.5, embedding_dim=256, num_heads=4, dropout_rate=0.13,
        encoder_block_types=("Enc", "Enc", "EncD", "EncA", "EncA", "EncD", "EncA", "EncA"),
        decoder_block_types=("DecA", "Dec", "DecA", "DecA", "DecA", "DecU", "DecA", "DecA", "DecA", "DecU", "Dec", "Dec", "Dec"),
        encoder_out_channels=(256,) * 8, decoder_out_channels=(256,) * 13,
        skip_connections=(False, False, True, True, True, False, True, True, True, False, True, True, True)),
    embedding=dict(fourier_dim=64, embedding_dim=256, num_classes=None),
    diffuser=dict(P_mean=-1.2, P_std=1.2),
    edm=dict(use_ema=True, ema_length=0.13, use_uncertainty=False, lr=0.02, steady_steps=200, rampup_steps=200,
             scheduler_interval="epoch"),
    image=(3, 32, 32), batch=256)

MNIST = dict(
    denoiser=dict(
        in_channels=1, out_channels=1, sigma_data=0.5, embedding_dim=256, num_heads=4, dropout_rate=0.1,
        encoder_block_types=("Enc", "Enc", "Enc", "EncD", "EncA", "EncA", "EncA", "EncD", "EncA", "EncA", "EncA"),
        decoder_block_types=("DecA", "Dec", "DecA", "DecA", "DecA", "DecA", "DecU", "DecA", "DecA", "DecA", "DecA", "DecU",
                             "Dec", "Dec", "Dec", "Dec"),
        encoder_out_channels=(128,) * 4 + (256,) * 4 + (512,) * 3,
        decoder_out_channels=(512,) * 7 + (256,) * 5 + (128,) * 4,
        skip_connections=(False, False, True, True, True, True, False, True, True, True, True, False, True, True, True, True)),
    embedding=dict(fourier_dim=64, embedding_dim=256, num_classes=10),
    diffuser=dict(P_mean=-1.2, P_std=1.2),
    edm=dict(use_ema=False, ema_length=0.1, use_uncertainty=False, lr=0.01, steady_steps=500, rampup_steps=500,
             scheduler_interval="epoch"),
    image=(1, 28, 28), batch=128)


IMAGENET = dict(  # experiments/conf/imagenet.yaml:20-51; the architecture lists are the Denoiser defaults (networks.py:332-432)
    denoiser=dict(in_channels=4, out_channels=4, sigma_data=0.5, embedding_dim=768, num_heads=4, dropout_rate=0.0),
    embedding=dict(fourier_dim=192, embedding_dim=768, num_classes=1000),
    diffuser=dict(P_mean=-0.4, P_std=1.0),
    edm=dict(use_ema=True, ema_length=0.13, use_uncertainty=False, lr=0.01, steady_steps=70000, rampup_steps=2000,
             scheduler_interval="step"),
    image=(4, 64, 64), batch=176)


def build_edm(cfg: dict, *, num_classes="cfg", dropout_rate=None, use_uncertainty=None):
    """EDM module for one of the dicts above (keyword overrides for the benchmark / tests)."""
    from . import EDM, Denoiser, Diffuser, Embedding
    dk = dict(cfg["denoiser"])
    if dropout_rate is not None:
        dk["dropout_rate"] = dropout_rate
    ek = dict(cfg["embedding"])
    if num_classes != "cfg":
        ek["num_classes"] = num_classes
    mk = dict(cfg["edm"])
    if use_uncertainty is not None:
        mk["use_uncertainty"] = use_uncertainty
    return EDM(diffuser=Diffuser(**cfg["diffuser"]), embedding=Embedding(**ek), denoiser=Denoiser(**dk), **mk)
