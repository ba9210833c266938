"""Where does a style-term discrepancy come from?  Native (mean, srm) of one image vs the fp32/fp64 CPU oracle, and
the W2 term evaluated by the ORACLE chain from either set of statistics."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import style_transfer_b200 as stb  # noqa: E402
from oracle import st_oracle as O  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device('cuda:0')
wts = O.make_vgg_weights(1234)
st = stb.StyleTransfer(devices=['cuda:0'], pooling='max', vgg_weights=wts)
m = st.model
m.ensure_workspace([(size, size)])
torch.manual_seed(0)
cimg = O.to_tensor(O.synth_image(1, 16, size, size))
simg = O.to_tensor(O.synth_image(2, 32, size - 8, size - 4))
img = (cimg + 0.05 * torch.randn_like(cimg)).clamp(0, 1)
nm, ns = m.style_stats(img.to(dev))
tm, ts = m.style_stats(simg.to(dev))
torch.cuda.synchronize()
acts = O.vgg_forward(img, wts, 'max', 6)
acts_b = O.vgg_forward(img, wts, 'max', 6, True)
acts_s = O.vgg_forward(simg, wts, 'max', 6)
for li, layer in enumerate((1, 6)):
    f = acts[layer]
    c = f.shape[1]
    om, os_ = O.style_stats(f.double())
    bm, bs = O.style_stats(acts_b[layer].double())
    nmu, nsr = nm[li].cpu().double(), ns[li].cpu().double()
    cov_o = os_ - torch.outer(om, om)
    cov_b = bs - torch.outer(bm, bm)
    cov_n = nsr - torch.outer(nmu, nmu)
    sc = cov_o.diagonal().mean()
    print(f'L{layer}: mean rel err native {float((nmu-om).abs().max()/om.abs().max()):.2e} (bf16sim {float((bm-om).abs().max()/om.abs().max()):.2e})  '
          f'srm rel {float((nsr-os_).abs().max()/os_.abs().max()):.2e} (bf16sim {float((bs-os_).abs().max()/os_.abs().max()):.2e})  '
          f'cov max err/mean diag {float((cov_n-cov_o).abs().max()/sc):.2e} (bf16sim {float((cov_b-cov_o).abs().max()/sc):.2e})  '
          f'tr(cov) rel {float((cov_n.trace()-cov_o.trace())/cov_o.trace()):+.2e} (bf16sim {float((cov_b.trace()-cov_o.trace())/cov_o.trace()):+.2e})')
    tmu, tsr = O.style_stats(acts_s[layer].double())
    tg = O.StyleTarget.build(tmu, tsr)
    tg_n = O.StyleTarget.build(tm[li].cpu().double(), ts[li].cpu().double())

    def term(mean, srm, t):
        eye = torch.eye(c, dtype=torch.float64)
        cov = srm - torch.outer(mean, mean) + eye * 1e-4
        r = O.sqrtm_ns(t.cov_sqrt @ cov @ t.cov_sqrt, 12)
        return float(((mean - t.mean) ** 2).mean() + torch.diagonal(t.cov + cov - 2 * r).mean())
    print(f'    W2 term (fp64 chain): oracle stats {term(om, os_, tg):.7f}  bf16sim stats {term(bm, bs, tg):.7f}  '
          f'native image stats + oracle target {term(nmu, nsr, tg):.7f}  native both {term(nmu, nsr, tg_n):.7f}')

# ---- the native W2 engine alone, fed with the fp64 oracle statistics (cast to fp32)
import ctypes
from style_transfer_b200 import _lib
lib = _lib.load_test()
P = lambda t: ctypes.c_void_p(t.data_ptr())
for li, layer in enumerate((1, 6)):
    f = acts_b[layer]
    c = f.shape[1]
    npix = float(f.shape[2] * f.shape[3])
    om, os_ = O.style_stats(f.double())
    tmu, tsr = O.style_stats(O.vgg_forward(simg, wts, 'max', 6, True)[layer].double())
    wsb = lib.stb_test_w2_workspace_bytes()
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    loss = torch.zeros(1, device=dev); gs = torch.empty(c, c, device=dev); gmu = torch.empty(c, device=dev)
    cs = torch.empty(c, c, device=dev)
    Sraw = (os_ * npix).float().to(dev); sums = (om * npix).float().to(dev)
    mt_d, st_d = tmu.float().to(dev), tsr.float().to(dev)
    _lib.check(lib.stb_test_w2(P(mt_d), P(st_d), P(Sraw), P(sums), c, ctypes.c_float(npix), ctypes.c_float(1.0), P(ws),
                               ctypes.c_size_t(wsb), P(loss), P(gs), P(gmu), P(cs), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    tg = O.StyleTarget.build(tmu, tsr)
    eye = torch.eye(c, dtype=torch.float64)
    cov = os_ - torch.outer(om, om) + eye * 1e-4
    r = O.sqrtm_ns(tg.cov_sqrt @ cov @ tg.cov_sqrt, 12)
    l64 = float(((om - tg.mean) ** 2).mean() + torch.diagonal(tg.cov + cov - 2 * r).mean())
    print(f'L{layer}: native W2 engine loss {float(loss):.7f} vs fp64 chain {l64:.7f}  rel {(float(loss) - l64) / l64:+.2e}   '
          f'cov_sqrt max err / max {float((cs.cpu().double() - tg.cov_sqrt).abs().max() / tg.cov_sqrt.abs().max()):.2e}')
