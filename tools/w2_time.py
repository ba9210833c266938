"""Device time of the W2 chain (and the other kernel classes) per iteration at a small image size -- the chain's cost is
resolution independent.  STB_W2_CHAIN=0 selects one launch per round instead of the persistent chain kernel."""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import style_transfer_b200 as stb  # noqa: E402
from style_transfer_b200 import _lib  # noqa: E402
from oracle import st_oracle as O  # noqa: E402  (weights/images fixture only)

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device('cuda:0')
wts = O.make_vgg_weights(1234)
st = stb.StyleTransfer(devices=['cuda:0'], pooling='max', vgg_weights=wts)
m = st.model
m.ensure_workspace([(size, size)])
cimg = O.to_tensor(O.synth_image(1, 16, size, size)).to(dev)
simg = O.to_tensor(O.synth_image(2, 32, size, size)).to(dev)
ct = m.content_features(cimg)
means, srms = m.style_stats(simg)
m.set_targets(size, size, ct, 0.015, means, srms, st.style_weights, 2.0)
st.image = cimg.clone()
st.average = stb.style_transfer.EMA(st.image, 0.99)
ea, eas = torch.zeros_like(st.image), torch.zeros_like(st.image)
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
for i in range(5):
    st._iterate(ea, eas, i + 1, 0.02, 0.99, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(50):
    st._iterate(ea, eas, i + 6, 0.02, 0.99, True)
e1.record()
torch.cuda.synchronize()
print(f'size {size}: graphed iteration {e0.elapsed_time(e1) / 50:.4f} ms, loss {float(st._loss_host[0]):.6f}')
_lib.check(m.lib.stb_profile_enable(m.ctx, 1))
for i in range(20):
    st._iterate(ea, eas, i + 56, 0.02, 0.99, True)
torch.cuda.synchronize()
ms, cnt = (ctypes.c_float * 10)(), (ctypes.c_int * 10)()
_lib.check(m.lib.stb_profile_read(m.ctx, ms, cnt, 10))
names = ['conv0_fwd_tv', 'conv_fwd', 'pool_fwd', 'gram', 'sse', 'w2', 'conv_bwd', 'pool_bwd', 'conv0_bwd_adam', 'finalize']
print({n: round(ms[i] / 20, 4) for i, n in enumerate(names)})
