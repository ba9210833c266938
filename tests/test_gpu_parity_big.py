"""-m gpu: loss / gradient parity against the UNMODIFIED reference at the BASELINE.json sizes.

The fixtures tests/golden/big_<S>.npz hold what the reference's own modules (VGGFeatures + ContentLossMSE +
StyleLossW2 + TVLoss under autograd, fp32 on the CPU; ST:20-234, SQ:9-55) produce for the S x S case of
tests/golden/make_big_parity.py: the seven loss terms of ST:455, their total, and three views of d loss / d image.
Here the same case is rebuilt from its seeds and evaluated by the native closure (`stb_iterate_ex`, apply_update=0).

Bar (BASELINE.json north_star): 1e-3 relative on the loss.  It is asserted on the total AND on every style term --
the W2 terms are cancellations (tr(St + S - 2 sqrt(.))) that amplify accumulation bias with the pixel count, so they
are the ones that move with size (1.5e-4 at 512^2 ... 5e-4 at 2048^2 in round 1).
"""
import importlib.util
import json
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).resolve().parent / 'golden'
_spec = importlib.util.spec_from_file_location('make_big_parity', GOLD / 'make_big_parity.py')
MB = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MB)

TERM_NAMES = ['content', 'relu1_1', 'relu2_1', 'relu3_1', 'relu4_1', 'relu5_1', 'tv']
LOSS_TOL = 1e-3          # north_star
# every term individually, relative to the term itself.  relu1_1 / relu2_1 carry 94 % of the style weight (256 + 64 of
# 341) and hold the bar; the deep taps (weights 16, 4, 1 of 341) sit on bf16 features that went through 5 / 9 / 13
# bf16-stored layers and are averaged over few pixels (relu5_1: (S/16)^2), their noise floor is ~1.3e-3 (measured at
# 256^2, where accumulation length plays no role).  The content MSE is a small difference of two independently rounded
# bf16 feature maps: the rounding noise adds its variance to it (0.7e-3 at 256^2 ... 5.8e-3 at 4096^2 of a term that
# is 1.5-3.5 % of the loss).  TV is fp32 on the raw image.
TERM_TOL = dict(content=8e-3, relu1_1=1e-3, relu2_1=1e-3, relu3_1=2.5e-3, relu4_1=3e-3, relu5_1=3.5e-3, tv=1e-5)
# ... and that gap IS the operand quantisation BASELINE.json configs[2] prescribes ("bf16 conv / fp32 accumulate"): the
# fixtures also hold the same case evaluated by the oracle's `sim_bf16` mode (reference algorithm + bf16-rounded conv
# weights + bf16-stored activations / feature gradients, nothing else changed; tests/golden/make_big_parity.py sim).
# Against THAT every native term must agree to a few 1e-4 -- a kernel bug cannot hide behind the bf16 allowance.
# (measured 256^2 ... 4096^2: content <= 9e-4, relu1_1 <= 6.5e-4, relu2_1 <= 3.3e-4, relu3_1 <= 1.8e-4, relu4_1 <= 1.9e-4,
# relu5_1 <= 2.3e-4 except 8.5e-4 at 256^2 where it averages 256 pixels; total <= 4.0e-4)
SIM_TERM_TOL = dict(content=1.5e-3, relu1_1=9e-4, relu2_1=7e-4, relu3_1=6e-4, relu4_1=6e-4, relu5_1=1.2e-3, tv=1e-5)
SIM_LOSS_TOL = 6e-4
RECORD = Path(__file__).resolve().parent.parent / 'gpurun_out'


@pytest.fixture(scope='module')
def G():
    import gpu_util as g
    return g


def native_eval(G, vgg_weights, size):
    content, style, img = MB.big_case(size)
    from oracle import st_oracle as O  # to_tensor only (input conversion, identical to TF.to_tensor)
    st = G.make_st('max', vgg_weights)
    m = st.model
    m.ensure_workspace([(size, size)])
    ct = m.content_features(O.to_tensor(content).to(G.DEV))
    means, srms = m.style_stats(O.to_tensor(style).to(G.DEV))
    m.set_targets(size, size, ct, MB.CONTENT_WEIGHT, means, srms, st.style_weights, MB.TV_WEIGHT)
    st.image = img.to(G.DEV).contiguous()
    terms, grad = st.loss_and_grad()
    m.release_workspace()
    return terms.double().numpy(), grad


@pytest.mark.parametrize('size', [256, 512, 1024, 2048, 4096])
def test_loss_terms_and_gradient_match_the_reference(G, vgg_weights, size):
    gold_path = GOLD / f'big_{size}.npz'
    if not gold_path.exists():
        pytest.skip(f'{gold_path.name} not generated')
    free, _ = torch.cuda.mem_get_info()
    if free < 1100 * size * size + (2 << 30):   # ~1.05 KiB of workspace per pixel (4.1 GiB at 2048^2) + slack
        pytest.skip('not enough free HBM for this size')
    gold = np.load(gold_path)
    terms, grad = native_eval(G, vgg_weights, size)
    total_err = abs(terms[0] - gold['total']) / gold['total']
    term_err = np.abs(terms[1:8] - gold['terms']) / np.abs(gold['terms'])
    norm, pooled, crop = MB.grad_views(grad.cpu())
    cos_pooled = F.cosine_similarity(torch.from_numpy(pooled).flatten().double(),
                                     torch.from_numpy(gold['grad_pooled']).flatten().double(), dim=0).item()
    cos_crop = F.cosine_similarity(torch.from_numpy(crop).flatten().double(),
                                   torch.from_numpy(gold['grad_crop']).flatten().double(), dim=0).item()
    norm_err = abs(norm - gold['grad_norm']) / gold['grad_norm']
    rec = dict(size=size, total=float(terms[0]), total_ref=float(gold['total']), total_rel_err=float(total_err),
               term_rel_err=dict(zip(TERM_NAMES, map(float, term_err))), grad_norm_rel_err=float(norm_err),
               grad_cos_block_means=cos_pooled, grad_cos_crop=cos_crop)
    if 'terms_sim' in gold.files:
        sim_err = np.abs(terms[1:8] - gold['terms_sim']) / np.abs(gold['terms_sim'])
        rec['vs_sim_bf16_oracle'] = dict(total=float(abs(terms[0] - gold['total_sim']) / gold['total_sim']),
                                         terms=dict(zip(TERM_NAMES, map(float, sim_err))))
    print(json.dumps(rec))
    if RECORD.is_dir():
        with open(RECORD / 'parity_big.jsonl', 'a') as f:
            f.write(json.dumps(rec) + '\n')
    assert total_err < LOSS_TOL, rec
    for name, e in zip(TERM_NAMES, term_err):
        assert e < TERM_TOL[name], (name, rec)
    if 'terms_sim' in gold.files:
        assert rec['vs_sim_bf16_oracle']['total'] < SIM_LOSS_TOL, rec
        for name, e in zip(TERM_NAMES, sim_err):
            assert e < SIM_TERM_TOL[name], (name, rec)
    # gradient: bf16 activations / feature gradients move individual pixels by a few %, not the direction
    assert norm_err < 2e-2, rec
    assert cos_pooled > 0.999 and cos_crop > 0.995, rec
