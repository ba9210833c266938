"""Generate the committed full-size parity fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

    python tests/golden/make_big_parity.py 1024 2048 [4096]

For each size S the reference's own modules (VGGFeatures, ContentLossMSE, StyleLossW2, TVLoss, Scale, LayerApply,
SumLoss; ST:20-234) are assembled exactly as `stylize()` does for one S x S scale (ST:416-455) and evaluated with
autograd, in fp32 on the CPU, on a perturbed iterate.  Stored per size (tests/golden/big_<S>.npz, a few hundred KB):
the seven loss terms in ST:455 order, their python-sum total, and three views of d loss / d image (its L2 norm, its
64 x 64 block means, one full-resolution 64 x 64 crop).  Inputs are regenerated from seeds by `big_case(S)` -- the
test on the GPU box (tests/test_gpu_parity_big.py) calls the same function, so nothing but outputs is committed.

Fixture: VGG-19 conv weights = oracle.st_oracle.make_vgg_weights(1234) served in place of the ImageNet checkpoint
download; torch 2.11.0+cu128 / torchvision 0.26.0, CPU.
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import st_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
CONTENT_WEIGHT, TV_WEIGHT = 0.015, 2.0
CROP = 64


def big_case(size):
    """(content PIL, style PIL, iterate tensor [1,3,S,S]) of the S x S parity case; deterministic everywhere (numpy
    PCG64 + PIL bicubic)."""
    content, style = O.synth_image(1, 16, size, size), O.synth_image(2, 32, size, size)
    cimg = O.to_tensor(content)
    noise = np.random.default_rng(1000 + size).standard_normal((1, 3, size, size), dtype=np.float32)
    img = (cimg + 0.05 * torch.from_numpy(noise)).clamp(0, 1).contiguous()
    return content, style, img


def grad_views(grad):
    """norm, block means [3,64,64], crop [3,64,64] of a [1,3,S,S] gradient (float64 statistics)."""
    g = grad.detach().double()[0]
    s = g.shape[-1]
    b = s // 64
    pooled = g[:, :b * 64, :b * 64].reshape(3, 64, b, 64, b).mean(dim=(2, 4))
    o = s // 3
    return float(g.norm()), pooled.numpy(), g[:, o:o + CROP, o:o + CROP].numpy().copy()


def main():
    from oracle import reference_harness as R
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048]
    weights = O.make_vgg_weights(1234)
    ref = R.import_reference()
    with R.patched_checkpoint(weights):
        st = ref.StyleTransfer(devices=['cpu'], pooling='max')
    for size in sizes:
        t0 = time.time()
        content, style, img = big_case(size)
        with torch.no_grad():
            feats_c = st.model(O.to_tensor(content), layers=[22])                     # ST:425
            ctarget = feats_c[22].clone()
            del feats_c
            sfeats = st.model(O.to_tensor(style), layers=st.style_layers)             # ST:440
            targets = {layer: ref.StyleLossW2.get_target(sfeats[layer]) for layer in st.style_layers}
            del sfeats
        losses = [ref.Scale(ref.LayerApply(ref.ContentLossMSE(ctarget), 22), CONTENT_WEIGHT)]
        for layer, w in zip(st.style_layers, st.style_weights):
            losses.append(ref.Scale(ref.LayerApply(ref.StyleLossW2(targets[layer]), layer), w))
        losses.append(ref.Scale(ref.LayerApply(ref.TVLoss(), 'input'), TV_WEIGHT))
        x = img.clone().requires_grad_()
        feats = st.model(x)
        terms = [l(feats) for l in losses]
        total = sum(terms)                                                             # ST:208
        total.backward()
        norm, pooled, crop = grad_views(x.grad)
        np.savez_compressed(OUT / f'big_{size}.npz', terms=np.array([float(t) for t in terms], dtype=np.float64),
                            total=np.float64(float(total)), grad_norm=np.float64(norm),
                            grad_pooled=pooled.astype(np.float32), grad_crop=crop.astype(np.float32))
        print(size, 'terms', [f'{float(t):.6g}' for t in terms], 'total', float(total), 'gnorm', norm,
              f'{time.time() - t0:.0f}s', flush=True)
        del feats, terms, total, x, losses, targets, ctarget


def main_sim():
    """python tests/golden/make_big_parity.py sim 256 512 ...: adds `terms_sim` / `total_sim` to big_<S>.npz -- the SAME
    case evaluated by oracle/st_oracle.py in its `sim_bf16` mode, i.e. the reference algorithm with the operand
    quantisation BASELINE.json configs[2] prescribes for the CUDA path ("bf16 conv / fp32 accumulate": bf16-rounded
    conv weights, activations and feature gradients stored in bf16, everything else fp32).  The GPU test holds the
    native terms to these within a few 1e-4: what separates a native style term from the fp32 reference (up to 3e-3 on
    the deep taps) is then that declared quantisation and nothing else."""
    sizes = [int(a) for a in sys.argv[2:]]
    weights = O.make_vgg_weights(1234)
    for size in sizes:
        t0 = time.time()
        content, style, img = big_case(size)
        with torch.no_grad():
            a_s = O.vgg_forward(O.to_tensor(style), weights, 'max', 29, True)
            targets = [O.StyleTarget.build(*O.style_stats(a_s[layer])) for layer in O.STYLE_LAYERS]
            del a_s
            a_c = O.vgg_forward(O.to_tensor(content), weights, 'max', 22, True)
            tg = O.ScaleTargets(a_c[22], targets, CONTENT_WEIGHT, TV_WEIGHT)
            del a_c
            det = {}
            loss, _ = O.loss_and_grad(img, weights, tg, 'max', sim_bf16=True, detail=det)
        path = OUT / f'big_{size}.npz'
        old = dict(np.load(path))
        old['terms_sim'] = np.array(det['terms'], dtype=np.float64)
        old['total_sim'] = np.float64(float(loss))
        np.savez_compressed(path, **old)
        rel = (old['terms_sim'] - old['terms']) / old['terms']
        print(size, 'sim terms', [f'{t:.6g}' for t in det['terms']], 'vs reference', [f'{r:+.1e}' for r in rel],
              f'{time.time() - t0:.0f}s', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'sim':
        main_sim()
    else:
        main()
