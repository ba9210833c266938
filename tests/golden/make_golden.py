"""Generate the committed golden vectors by running the UNMODIFIED reference (/root/reference) on CPU.

Run once in the build container:  python tests/golden/make_golden.py
Writes tests/golden/*.npz (small: inputs are regenerated from seeds, only outputs are stored).  These pin
oracle/st_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA path (tests/test_gpu_*.py).

Fixture: VGG-19 conv weights = oracle.st_oracle.make_vgg_weights(1234) served in place of the ImageNet checkpoint
download (no network); images = low-frequency synthetic fields.  torch 2.11.0+cu128 /
torchvision 0.26.0, CPU, `torch.manual_seed(0)`.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_harness as R  # noqa: E402
from oracle import st_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent

# name -> (content (seed, base, w, h), styles [(seed, base, w, h)], pooling, stylize kwargs)
CASES = {
    'max_64x48_single': dict(content=(1, 16, 64, 48), styles=[(2, 32, 56, 40)], pooling='max',
                             kw=dict(min_scale=64, end_scale=64, initial_iterations=6)),
    'avg_80x56_two_styles': dict(content=(1, 16, 80, 56), styles=[(2, 32, 72, 60), (3, 32, 50, 80)], pooling='average',
                                 kw=dict(min_scale=80, end_scale=80, initial_iterations=4, style_weights=[3, 1])),
    'l2_72x72_single': dict(content=(4, 16, 72, 72), styles=[(5, 32, 64, 64)], pooling='l2',
                            kw=dict(min_scale=72, end_scale=72, initial_iterations=3)),
    'max_pyramid_32_64': dict(content=(1, 16, 64, 48), styles=[(2, 32, 56, 40)], pooling='max',
                              kw=dict(min_scale=32, end_scale=64, iterations=3, initial_iterations=4)),
    'max_128_noise_tv': dict(content=(6, 128, 128, 96), styles=[(7, 64, 100, 128)], pooling='max',
                             kw=dict(min_scale=128, end_scale=128, initial_iterations=3, tv_weight=5.0,
                                     content_weight=0.05, style_scale_fac=0.75)),
    # BASELINE.json configs[1]: 512 x 512 end_scale, default multi-scale (128, 181, 256, 362, 512), iteration counts
    # cut from 1000 + 4 x 500 to 20 + 4 x 10
    'max_pyramid_128_512': dict(content=(1, 16, 512, 512), styles=[(2, 32, 512, 512)], pooling='max',
                                kw=dict(min_scale=128, end_scale=512, iterations=10, initial_iterations=20)),
}
QUANTISED_FINAL = {'max_pyramid_128_512'}  # final image stored as floor(x * 255) uint8 (what get_image() returns)


def build_case(name):
    case = CASES[name]
    content = O.synth_image(*case['content'])
    styles = [O.synth_image(*s) for s in case['styles']]
    return content, styles, case['pooling'], case['kw']


def main():
    weights = O.make_vgg_weights(1234)
    only = set(sys.argv[1:])
    for name in CASES:
        if only and name not in only:
            continue
        content, styles, pooling, kw = build_case(name)
        torch.manual_seed(0)
        image, trace, st = R.run_reference(content, styles, weights, pooling=pooling, seed=0, **kw)
        losses = np.array([t['loss'] for t in trace], dtype=np.float64)
        sizes = np.array([(t['w'], t['h'], t['i'], t['i_max']) for t in trace], dtype=np.int32)
        final = image.numpy().astype(np.float32)
        if name in QUANTISED_FINAL:
            final = np.floor(final * 255).astype(np.uint8)
        np.savez_compressed(OUT / f'{name}.npz', losses=losses, sizes=sizes, final_image=final)
        print(name, losses[:3], '...', losses[-1], image.shape, flush=True)

    if only and 'max_64x48_internals' not in only:
        return
    # single-iteration internals for the first case: per-term losses and d loss/d image from reference autograd
    ref = R.import_reference()
    content, styles, pooling, kw = build_case('max_64x48_single')
    torch.manual_seed(0)
    with R.patched_checkpoint(weights):
        st = ref.StyleTransfer(devices=['cpu'], pooling=pooling)
    captured = {}

    def cb(it):
        pass

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        st.stylize(content, styles, callback=cb, min_scale=64, end_scale=64, initial_iterations=0)
    # rebuild the criterion exactly as stylize() does, on a perturbed image
    from PIL import Image
    cw, ch = 64, 48
    cimg = ref.TF.to_tensor(content.resize((cw, ch), Image.BICUBIC))[None]
    g = torch.Generator().manual_seed(123)
    img = (cimg + 0.05 * torch.randn(cimg.shape, generator=g)).clamp(0, 1).requires_grad_()
    feats_c = st.model(cimg, layers=[22])
    sfeats = st.model(ref.TF.to_tensor(styles[0].resize(ref.size_to_fit(styles[0].size, 64), Image.BICUBIC))[None],
                      layers=st.style_layers)
    losses = [ref.Scale(ref.LayerApply(ref.ContentLossMSE(feats_c[22]), 22), 0.015)]
    for layer, w in zip(st.style_layers, st.style_weights):
        tgt = ref.StyleLossW2.get_target(sfeats[layer])
        losses.append(ref.Scale(ref.LayerApply(ref.StyleLossW2(tgt), layer), w))
    losses.append(ref.Scale(ref.LayerApply(ref.TVLoss(), 'input'), 2.0))
    feats = st.model(img)
    terms = [l(feats) for l in losses]
    total = sum(terms)
    total.backward()
    np.savez_compressed(OUT / 'max_64x48_internals.npz', image=img.detach().numpy(),
                        terms=np.array([float(t) for t in terms]), total=float(total), grad=img.grad.numpy(),
                        tap1=feats[1].detach().numpy()[:, :8], tap29=feats[29].detach().numpy()[:, :8])
    print('internals', [float(t) for t in terms], float(total))


if __name__ == '__main__':
    main()
