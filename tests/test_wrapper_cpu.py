"""`DenoiserWrapper` (networks.py:608-646): unused by every shipped config, kept importable and equal to
the reference's — checked against the reference's own class (bytecode in oracle/_ref) when that is built, and against the
EDM preconditioning formulas (networks.py:637-641) always. Plain tensor arithmetic: runs on the CPU."""
import torch

import tinyedm_b200 as T
from oracle import ref_loader


class _Net(torch.nn.Module):
    """Stands in for `net(c_in * x, c_noise, embedding)`: uses all three arguments."""

    def forward(self, x, c_noise, emb):
        y = torch.tanh(x) * (1.0 + c_noise.view(-1, 1, 1, 1))
        return y if emb is None else y + emb.mean(dim=1).view(-1, 1, 1, 1)


def test_denoiser_wrapper_matches_the_reference_class_and_the_formulas():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5, 3, 8, 8, generator=g)
    sigma = torch.exp(torch.randn(5, generator=g))
    emb = torch.randn(5, 16, generator=g)
    sd = 0.5
    ours = T.DenoiserWrapper(_Net(), sd)
    assert ours.sigma_data == sd
    for e in (emb, None):
        for s in (sigma, sigma.view(-1, 1, 1, 1)):
            D = ours(x, s, e)
            sv = sigma.view(-1, 1, 1, 1)
            c_skip = sd ** 2 / (sv ** 2 + sd ** 2)
            c_out = sv * sd / (sv ** 2 + sd ** 2).sqrt()
            c_in = 1 / (sd ** 2 + sv ** 2).sqrt()
            want = c_skip * x + c_out * _Net()(c_in * x, (sv.log() / 4).flatten(), e)
            assert torch.allclose(D, want, rtol=1e-6, atol=1e-6)
            ref = ref_loader.load()
            if ref is not None:
                D_ref = ref.networks.DenoiserWrapper(_Net(), sd)(x, s, e)
                assert torch.allclose(D, D_ref, rtol=1e-6, atol=1e-6)
