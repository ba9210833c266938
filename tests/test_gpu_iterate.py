"""-m gpu: the whole hot path through the public class / C-ABI against the oracle and the reference goldens.

Bar (BASELINE.json north_star): loss within 1e-3 relative of the reference's own PyTorch path on identical
inputs/seeds.  The CUDA path stores activations in bf16 (fp32 accumulate); the oracle's `sim_bf16` mode models those
rounding points, so the comparison against it is much tighter (3e-4) and isolates real bugs from quantisation.
"""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import st_oracle as O  # noqa: E402

GOLD = Path(__file__).resolve().parent / 'golden'
_spec = importlib.util.spec_from_file_location('make_golden', GOLD / 'make_golden.py')
MG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MG)


@pytest.fixture(scope='module')
def G():
    import gpu_util as g
    return g


@pytest.mark.parametrize('H,W,pooling', [(48, 64, 'max'), (56, 80, 'average'), (72, 72, 'l2'), (45, 34, 'max'),
                                         (16, 16, 'max')])
def test_targets_loss_and_gradient(G, vgg_weights, H, W, pooling):
    st = G.make_st(pooling, vgg_weights)
    cimg = O.to_tensor(O.synth_image(1, 16, W, H))
    simg = O.to_tensor(O.synth_image(2, 32, max(W - 8, 16), max(H - 4, 16)))
    m = st.model
    m.ensure_workspace([(H, W), tuple(simg.shape[2:])])
    ct = m.content_features(cimg.to(G.DEV))
    means, srms = m.style_stats(simg.to(G.DEV))
    torch.cuda.synchronize()
    acts_s = O.vgg_forward(simg, vgg_weights, pooling, 29)
    for li, layer in enumerate(O.STYLE_LAYERS):
        om, osrm = O.style_stats(acts_s[layer])
        assert G.rel_err(means[li], om) < 1e-2
        assert G.rel_err(srms[li], osrm) < 1e-2
    acts_c = O.vgg_forward(cimg, vgg_weights, pooling, 22)
    assert G.rel_err(G.nchw(ct), acts_c[22]) < 3e-2
    m.set_targets(H, W, ct, 0.015, means, srms, st.style_weights, 2.0)
    torch.manual_seed(0)
    img = (cimg + 0.05 * torch.randn_like(cimg)).clamp(0, 1)
    st.image = img.to(G.DEV).contiguous()
    terms, grad = st.loss_and_grad()
    for sim, tol_loss, tol_cos in ((False, 1e-3, 0.995), (True, 3e-4, 0.9995)):
        a_s = O.vgg_forward(simg, vgg_weights, pooling, 29, sim)
        a_c = O.vgg_forward(cimg, vgg_weights, pooling, 22, sim)
        tg = O.ScaleTargets(a_c[22], [O.StyleTarget.build(*O.style_stats(a_s[layer])) for layer in O.STYLE_LAYERS],
                            0.015, 2.0)
        det = {}
        ol, og = O.loss_and_grad(img, vgg_weights, tg, pooling, sim_bf16=sim, detail=det)
        assert abs(terms[0].item() - float(ol)) / float(ol) < tol_loss
        np.testing.assert_allclose(terms[1:8].numpy(), det['terms'], rtol=3e-2, atol=2e-6)
        cos = F.cosine_similarity(grad.cpu().flatten(), og.flatten(), dim=0).item()
        assert cos > tol_cos


def test_loss_parity_at_512(G, vgg_weights):
    """The loss bar must hold where the accumulation chains are long: at 512^2 every style term's Gram runs over
    262 144 pixels and the ill-conditioned (eps = 1e-4) covariances amplify any bias of the tensor-core accumulation
    (a mirrored / long-chain W2 engine was -1.3e-3 here while passing every small case)."""
    H = W = 512
    st = G.make_st('max', vgg_weights)
    cimg = O.to_tensor(O.synth_image(1, 16, W, H))
    simg = O.to_tensor(O.synth_image(2, 32, W - 8, H - 4))
    m = st.model
    m.ensure_workspace([(H, W), tuple(simg.shape[2:])])
    ct = m.content_features(cimg.to(G.DEV))
    means, srms = m.style_stats(simg.to(G.DEV))
    m.set_targets(H, W, ct, 0.015, means, srms, st.style_weights, 2.0)
    torch.manual_seed(0)
    img = (cimg + 0.05 * torch.randn_like(cimg)).clamp(0, 1)
    st.image = img.to(G.DEV).contiguous()
    terms, _ = st.loss_and_grad()
    a_s = O.vgg_forward(simg, vgg_weights, 'max', 29)
    a_c = O.vgg_forward(cimg, vgg_weights, 'max', 22)
    tg = O.ScaleTargets(a_c[22], [O.StyleTarget.build(*O.style_stats(a_s[layer])) for layer in O.STYLE_LAYERS],
                        0.015, 2.0)
    det = {}
    ol, _ = O.loss_and_grad(img, vgg_weights, tg, 'max', detail=det)
    assert abs(terms[0].item() - float(ol)) / float(ol) < 1e-3   # north_star bar; measured 1.5e-4
    # the two large style terms (relu1_1, relu2_1 = 2/3 of the loss) individually
    np.testing.assert_allclose(terms[2:4].numpy(), det['terms'][1:3], rtol=1e-3)


@pytest.mark.parametrize('name', ['max_64x48_single', 'avg_80x56_two_styles', 'l2_72x72_single', 'max_pyramid_32_64',
                                  'max_128_noise_tv'])
def test_stylize_matches_reference_golden(G, vgg_weights, name):
    """Public API end to end (pyramid, Adam warm start with carried step, EMA) vs the UNMODIFIED reference's trace."""
    gold = np.load(GOLD / f'{name}.npz')
    content, styles, pooling, kw = MG.build_case(name)
    st = G.make_st(pooling, vgg_weights)
    trace = []
    out = st.stylize(content, styles, callback=lambda it: trace.append((it.loss, it.w, it.h, it.i)), **kw)
    assert len(trace) == len(gold['losses'])
    np.testing.assert_array_equal(np.array([t[1:] for t in trace]), gold['sizes'][:, :3])
    losses = np.array([t[0] for t in trace])
    # first iteration of the first scale: identical inputs -> the 1e-3 bar of the north star
    assert abs(losses[0] - gold['losses'][0]) / gold['losses'][0] < 1e-3
    # later iterations: trajectories of a bf16 and an fp32 optimiser drift slowly
    np.testing.assert_allclose(losses, gold['losses'], rtol=5e-3)
    # same uint8 truncation as get_image() on both sides; the residual is trajectory drift of a sign-like optimiser
    img = np.asarray(out, dtype=np.float32).transpose(2, 0, 1) / 255
    gold_q = np.floor(gold['final_image'] * 255) / 255
    assert np.abs(img - gold_q).mean() < 6e-3


def test_iterate_state_update_matches_oracle(G, vgg_weights):
    """One stb_iterate: Adam moments, clamped image and EMA vs the oracle fed with the native gradient's oracle twin."""
    H, W = 48, 64
    content, style = O.synth_image(1, 16, W, H), O.synth_image(2, 32, 56, 40)
    st = G.make_st('max', vgg_weights)
    tr = []
    st.stylize(content, [style], min_scale=64, end_scale=64, initial_iterations=1, callback=lambda it: tr.append(it.loss))
    tg, _ = O.make_targets(content, [style], [1.0], 64, vgg_weights, 'max', 0.015, 2.0, sim_bf16=True)
    s0 = O.IterState.fresh(O.to_tensor(content))
    loss = O.iterate(s0, vgg_weights, tg, 'max', sim_bf16=True)
    assert abs(tr[0] - loss) / loss < 3e-4
    # after one Adam step every pixel moved by lr * sign(g) (bias-corrected); stylize() then copies the
    # bias-corrected EMA back into the image (ST:496-497): compare both with the oracle's EMA
    ema_native = st.average.get().cpu()
    assert (ema_native - s0.ema_get()).abs().mean() < 1e-3   # sign flips only where |g| ~ 0
    assert (st.image.cpu() - s0.ema_get()).abs().mean() < 1e-3


def test_errors_are_reported(G, vgg_weights):
    st = G.make_st('max', vgg_weights)
    m = st.model
    m.ensure_workspace([(32, 32)])
    with pytest.raises(ValueError):  # ST:82-83
        m.style_stats(torch.zeros(1, 3, 8, 40, device=G.DEV))
    with pytest.raises(RuntimeError):  # iterate before set_targets
        st.image = torch.zeros(1, 3, 32, 32, device=G.DEV)
        st.loss_and_grad()
    with pytest.raises(ValueError):
        st.stylize(O.synth_image(1, 8, 32, 32), [O.synth_image(2, 8, 32, 32)], style_weights=[1, 2])
    with pytest.raises(ValueError):
        st.stylize(O.synth_image(1, 8, 32, 32), [O.synth_image(2, 8, 32, 32)], init='bogus')


def test_full_size_properties(G, vgg_weights):
    """2048x2048 (BASELINE.json config 3) is too big for the CPU oracle in a test, so check size-independent
    properties: finite decreasing loss, image stays in [0,1], EMA bias correction, determinism of the loss."""
    size = 2048
    content, style = O.synth_image(1, 16, size, size), O.synth_image(2, 32, size, size)
    st = G.make_st('max', vgg_weights)
    tr = []
    st.stylize(content, [style], min_scale=size, end_scale=size, initial_iterations=8, callback=lambda it: tr.append(it.loss))
    assert all(np.isfinite(tr)) and tr[-1] < tr[0]
    assert float(st.image.min()) >= 0.0 and float(st.image.max()) <= 1.0
    st2 = G.make_st('max', vgg_weights)
    tr2 = []
    st2.stylize(content, [style], min_scale=size, end_scale=size, initial_iterations=2, callback=lambda it: tr2.append(it.loss))
    np.testing.assert_allclose(tr2, tr[:2], rtol=1e-5)  # split-K / reductions are order-deterministic
    # translation-of-scale property: the first-iteration loss at 2048^2 of a low-frequency pair is close to the
    # 256^2 loss of the same pair evaluated by the ORACLE (statistics of smooth fields are resolution independent
    # only loosely, so this is a sanity bound, not a parity claim)
    assert 0.0 < tr[0] < 10.0


def test_two_gpu_banded_equals_single_gpu():
    """Spatial tiling over 2 GPUs (NCCL): loss trace and result vs the single-GPU run (tools/dist_check.py)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29533', str(root / 'tools' / 'dist_check.py'),
                        '384', '512', '5'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
