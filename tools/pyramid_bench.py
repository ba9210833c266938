"""BASELINE.json configs[1]: the default pyramid to end_scale=512 (scales 128,181,256,362,512; 1000 + 4 x 500 Adam
iterations), end to end through stylize(): native build vs the unmodified reference on its own CUDA path.
Usage: python tools/pyramid_bench.py [native|reference|both]"""
import contextlib
import io
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import st_oracle as O  # noqa: E402  (fixture generator only)

which = sys.argv[1] if len(sys.argv) > 1 else 'both'
wts = O.make_vgg_weights(1234)
content, style = O.synth_image(1, 16, 512, 512), O.synth_image(2, 32, 512, 512)
kw = dict(min_scale=128, end_scale=512, iterations=500, initial_iterations=1000)
out = {}
if which in ('native', 'both'):
    import style_transfer_b200 as stb
    st = stb.StyleTransfer(devices=['cuda:0'], pooling='max', vgg_weights=wts)
    with contextlib.redirect_stdout(io.StringIO()):
        st.stylize(content, [style], min_scale=128, end_scale=128, initial_iterations=5, callback=lambda it: None)
    tr = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        st.stylize(content, [style], callback=lambda it: tr.append(it.loss), **kw)
    torch.cuda.synchronize()
    out['native'] = dict(seconds=time.perf_counter() - t0, iterations=len(tr), final_loss=tr[-1])
if which in ('reference', 'both'):
    from oracle import reference_harness as RH
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, trace, _ = RH.run_reference(content, [style], wts, devices=('cuda:0',), **kw)
    torch.cuda.synchronize()
    out['reference_cuda'] = dict(seconds=time.perf_counter() - t0, iterations=len(trace), final_loss=trace[-1]['loss'])
if len(out) == 2:
    out['speedup'] = out['reference_cuda']['seconds'] / out['native']['seconds']
print(json.dumps(out))
