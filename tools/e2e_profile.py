"""Where does the end-to-end stylize() wall time go?  cProfile of one 2048^2 single-scale call (after a warm call)."""
import cProfile
import contextlib
import io
import pstats
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import style_transfer_b200 as stb  # noqa: E402
from oracle import st_oracle as O  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
its = int(sys.argv[2]) if len(sys.argv) > 2 else 20
wts = O.make_vgg_weights(1234)
content, style = O.synth_image(1, 16, size, size), O.synth_image(2, 32, size, size)
st = stb.StyleTransfer(devices=['cuda:0'], pooling='max', vgg_weights=wts)
kw = dict(min_scale=size, end_scale=size)
with contextlib.redirect_stdout(io.StringIO()):
    st.stylize(content, [style], initial_iterations=2, callback=lambda it: None, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    out = st.stylize(content, [style], initial_iterations=its, callback=lambda it: None, **kw)
torch.cuda.synchronize()
pr.disable()
t = time.perf_counter() - t0
print(f'total {t * 1000:.1f} ms for {its} iterations -> {its / t:.1f} it/s')
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28)
print(s.getvalue()[:6000])
