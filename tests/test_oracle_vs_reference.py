"""CPU, build container only: the oracle restatement against the LIVE unmodified reference, stage by stage with
amplified inputs (end-to-end parity alone is blind to e.g. a wrong GELU variant in corr_mlp)."""
import sys

import pytest
import torch

from cases import O
from cotracker_b200.synthetic import random_queries, seeded_state_dict, texture_video


@pytest.fixture(scope="module")
def ref(reference_path):
    sys.path.insert(0, reference_path)
    from cotracker.models.build_cotracker import build_cotracker
    sd = seeded_state_dict(2024, offline=True, window_len=60, head_gain=10.0, vis_gain=100.0)
    m = build_cotracker(None, offline=True, window_len=60).eval()
    m.load_state_dict(sd)
    return m, sd


def test_updateformer_amplified(ref):
    m, sd = ref
    x = torch.randn(1, 33, 7, 1110, generator=torch.Generator().manual_seed(1)) * 2
    with torch.no_grad():
        want = m.updateformer(x)
        got = O.updateformer(sd, x)
    assert float((got - want).abs().max()) < 1e-4 * max(1.0, float(want.abs().max()))


def test_corr_volume_and_corr_mlp_amplified(ref):
    m, sd = ref
    g = torch.Generator().manual_seed(2)
    T, N, H, W = 3, 11, 12, 16
    fm = torch.randn(1, T, 128, H, W, generator=g)
    coords = torch.rand(1 * T, N, 2, generator=g) * torch.tensor([W + 4.0, H + 4.0]) - 2.0   # some outside
    sup = torch.randn(1, 49, N, 128, generator=g)
    with torch.no_grad():
        feat = m.get_correlation_feat(fm, coords)
        s = sup.view(1, 1, 7, 7, N, 128).squeeze(1).permute(0, 3, 1, 2, 4)
        want = torch.einsum("btnhwc,bnijc->btnhwij", feat, s).reshape(T, N, 2401)
        got = O.correlation_volume(fm[0], sup[0], coords)
        assert float((got - want).abs().max()) < 1e-4
        big = want * 10                                            # out of the regime where erf == tanh GELU
        assert torch.allclose(O.mlp(sd, "corr_mlp", big, "none"), m.corr_mlp(big), atol=1e-4)
        assert not torch.allclose(O.mlp(sd, "corr_mlp", big, "tanh"), m.corr_mlp(big), atol=1e-4)


def test_support_features(ref):
    m, sd = ref
    g = torch.Generator().manual_seed(3)
    T, N, H, W = 4, 9, 12, 16
    fm = torch.randn(1, T, 128, H, W, generator=g)
    qf = torch.randint(0, T, (1, N), generator=g)
    qc = torch.rand(1, N, 2, generator=g) * torch.tensor([W - 1.0, H - 1.0])
    with torch.no_grad():
        _, want = m.get_track_feat(fm, qf, qc, support_radius=3)
        got = O.support_features(fm[0], qf[0], qc[0])
    assert float((got - want[0]).abs().max()) < 1e-5


def test_posenc_and_time_embedding(ref, reference_path):
    m, sd = ref
    from cotracker.models.core.cotracker.cotracker3_online import posenc
    x = torch.randn(5, 3, 4) * 0.1
    assert torch.equal(O.posenc(x), posenc(x, 0, 10))
    for t in (60, 16, 7):
        assert torch.allclose(O.time_embedding(sd, t), m.interpolate_time_embed(torch.zeros(1, t, 1110), t), atol=0)


def test_encoder_and_pyramid(ref):
    m, sd = ref
    v = texture_video(2, 64, 96, seed=5)[0] / 255 * 2 - 1
    with torch.no_grad():
        assert float((O.encoder(sd, v) - m.fnet(v)).abs().max()) < 1e-4


def test_offline_forward_live(ref):
    m, sd = ref
    video = texture_video(5, 64, 96, seed=6)
    q = random_queries(9, 5, 64, 96, seed=7)
    with torch.no_grad():
        wc, wv, wq, _ = m(video, q, iters=3)
        gc, gv, gq = O.offline_forward(sd, video, q, iters=3)
    assert float((gc - wc).abs().max()) < 1e-4 and float((gv - wv).abs().max()) < 1e-5
    assert float((wc[0, -1] - wc[0, 0]).abs().max()) > 0.5   # tracks move
