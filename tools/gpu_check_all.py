"""GPU bring-up: every kernel against the oracle, then the whole iteration, then a timing.  Diagnostic script
(prints, does not assert) -- the pytest versions live in tests/test_gpu_*.py."""
import ctypes
import math
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import style_transfer_b200 as stb  # noqa: E402
from style_transfer_b200 import _lib  # noqa: E402
from oracle import st_oracle as O  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')
lib = _lib.load_test()
P = _lib.ptr
S = _lib.cur_stream
ALL_OK = True


def rep(name, got, ref, tol, scale=None):
    global ALL_OK
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    err = (got - ref).abs().max().item()
    den = (ref.abs().max().item() if scale is None else scale) + 1e-30
    bad = int((~torch.isfinite(got)).sum())
    ok = bad == 0 and err / den < tol
    ALL_OK &= ok
    print(f'[{"OK " if ok else "BAD"}] {name}: max_err={err:.3e} rel={err / den:.3e} (tol {tol:g}) nonfinite={bad}',
          flush=True)
    return ok


def t_conv0(H, W):
    g = torch.Generator().manual_seed(H * 1000 + W)
    img = torch.rand(1, 3, H, W, generator=g)
    w = O.make_vgg_weights(7)
    w0, b0 = w[0]
    out = torch.empty(H, W, 64, dtype=torch.bfloat16, device=dev)
    gtv = torch.empty(3, H, W, device=dev)
    nparts = ((W + 63) // 64) * H
    parts = torch.zeros(nparts, device=dev)
    n = ctypes.c_int()
    tvw = 2.0
    img_d, w0_d, b0_d = img.to(dev), w0.to(dev), b0.to(dev)
    _lib.check(lib.stb_test_conv0_fwd(P(img_d), P(w0_d), P(b0_d), P(out), H, W, tvw, P(gtv), P(parts),
                                      ctypes.byref(n), S()))
    torch.cuda.synchronize()
    acts = O.vgg_forward(img, w, 'max', 1)
    rep(f'conv0 fwd {H}x{W}', out.float().permute(2, 0, 1)[None], acts[1], 6e-3)
    tvl, tvg = O.tv_loss_and_grad(img.double())
    rep(f'tv loss {H}x{W}', parts.sum().reshape(1), tvl.reshape(1).float(), 1e-5)
    rep(f'tv grad {H}x{W}', gtv[None], (tvg * tvw).float(), 1e-5)
    # conv0 bwd
    g0 = (torch.randn(H, W, 64, generator=g) * (torch.rand(H, W, 64, generator=g) > 0.5)).bfloat16()
    grad = torch.empty(1, 3, H, W, device=dev)
    g0_d = g0.to(dev)
    _lib.check(lib.stb_test_conv0_bwd(P(g0_d), P(w0_d), None, P(grad), H, W, S()))
    torch.cuda.synchronize()
    gp = F.conv_transpose2d(g0.float().permute(2, 0, 1)[None].double(), w0.double())
    gp[:, :, 1, :] += gp[:, :, 0, :]
    gp[:, :, -2, :] += gp[:, :, -1, :]
    gp[:, :, :, 1] += gp[:, :, :, 0]
    gp[:, :, :, -2] += gp[:, :, :, -1]
    ref = gp[:, :, 1:-1, 1:-1] / torch.tensor(O.NORM_STD, dtype=torch.float64).view(1, 3, 1, 1)
    rep(f'conv0 bwd {H}x{W}', grad, ref.float(), 1e-5)


def t_pool(H, W, C):
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(H, W, C, generator=g)).bfloat16()
    x[::3, ::2] = 0  # zero windows / ties
    go = torch.randn(H // 2, W // 2, C, generator=g).bfloat16()
    x_d, go_d = x.to(dev), go.to(dev)
    for name, code in (('max', 0), ('average', 1), ('l2', 2)):  # pool forward is fused into the conv (tests/)
        xin = x.float().permute(2, 0, 1)[None]
        gin = torch.full((H, W, C), float('nan'), dtype=torch.bfloat16, device=dev)
        _lib.check(lib.stb_test_pool_bwd(code, P(go_d), P(x_d), P(gin), H, W, C, S()))
        gref = O.pool_bwd(go.float().permute(2, 0, 1)[None], xin, name) * (xin > 0)
        rep(f'pool bwd {name} {H}x{W}x{C}', gin.float().permute(2, 0, 1)[None], gref, 5e-3)


def t_gram(Pn, C):
    g = torch.Generator().manual_seed(Pn + C)
    f = torch.relu(torch.randn(Pn, C, generator=g)).bfloat16().to(dev)
    nf = lib.stb_test_gram_partials_floats(Pn, C)
    ws = torch.empty(nf, device=dev)
    Sr = torch.empty(C, C, device=dev)
    sm = torch.empty(C, device=dev)
    _lib.check(lib.stb_test_gram(P(f), Pn, C, P(ws), nf, P(Sr), P(sm), S()))
    torch.cuda.synchronize()
    fd = f.double()
    rep(f'gram P={Pn} C={C}', Sr, (fd.t() @ fd).float(), 2e-5)
    rep(f'sums P={Pn} C={C}', sm, fd.sum(0).float(), 2e-5)


def t_w2(C):
    g = torch.Generator().manual_seed(C)
    n = 4 * C

    def moments(scale):
        f = torch.relu(torch.randn(C, n, generator=g) * scale + 0.2)
        return f.mean(1), (f @ f.t()) / n

    mt, st = moments(1.0)
    mc, sc = moments(1.2)
    npix, weight = float(n), 0.37
    wsb = lib.stb_test_w2_workspace_bytes()
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    loss = torch.zeros(1, device=dev)
    gs = torch.empty(C, C, device=dev)
    gmu = torch.empty(C, device=dev)
    cs = torch.empty(C, C, device=dev)
    Sraw = (sc * npix).to(dev)
    sums = (mc * npix).to(dev)
    mt_d, st_d = mt.to(dev), st.to(dev)
    _lib.check(lib.stb_test_w2(P(mt_d), P(st_d), P(Sraw), P(sums), C, npix, weight, P(ws), wsb, P(loss),
                               P(gs), P(gmu), P(cs), S()))
    torch.cuda.synchronize()
    for dt, tag, tol in ((torch.float32, 'f32', 2e-3), (torch.float64, 'f64', 5e-3)):
        tgt = O.StyleTarget.build(mt.to(dt), st.to(dt))
        # oracle expects an activation; feed moments directly through the same formulas
        eye = torch.eye(C, dtype=dt)
        cov = sc.to(dt) - torch.outer(mc.to(dt), mc.to(dt)) + eye * 1e-4
        md = ((mc.to(dt) - tgt.mean) ** 2).mean()
        r = O.sqrtm_ns(tgt.cov_sqrt @ cov @ tgt.cov_sqrt, 12)
        l = (md + torch.diagonal(tgt.cov + cov - 2 * r).mean()) * weight
        g_m = O.sqrtm_ns_lyap_backward(r, eye * (-2.0 * weight / C), 12)
        g_cov = tgt.cov_sqrt.t() @ g_m @ tgt.cov_sqrt.t() + eye * (weight / C)
        gsr = g_cov + g_cov.t()
        gmr = 2.0 * weight * (mc.to(dt) - tgt.mean) / C - gsr @ mc.to(dt)
        rep(f'w2 C={C} csqrt vs {tag}', cs, tgt.cov_sqrt, tol)
        rep(f'w2 C={C} loss  vs {tag}', loss, l.reshape(1), tol)
        rep(f'w2 C={C} Gs    vs {tag}', gs, gsr, tol * 5)
        rep(f'w2 C={C} gmu   vs {tag}', gmu * npix, gmr, tol * 5)


def make_st(pooling, wts):
    return stb.StyleTransfer(devices=['cuda:0'], pooling=pooling, vgg_weights=wts)


def t_full(H, W, pooling, wts, sims=(False, True)):
    st = make_st(pooling, wts)
    content = O.synth_image(1, 16, W, H)
    style = O.synth_image(2, 32, W - 8, H - 4)
    m = st.model
    m.ensure_workspace([(H, W), (H - 4, W - 8)])
    cimg = O.to_tensor(content)
    simg = O.to_tensor(style)
    ct = m.content_features(cimg.to(dev))
    means, srms = m.style_stats(simg.to(dev))
    torch.cuda.synchronize()
    # oracle targets
    acts_s = O.vgg_forward(simg, wts, pooling, 29)
    for li, layer in enumerate(O.STYLE_LAYERS):
        om, osrm = O.style_stats(acts_s[layer])
        rep(f'[{pooling} {H}x{W}] style mean L{layer}', means[li], om, 2e-2)
        rep(f'[{pooling} {H}x{W}] style srm  L{layer}', srms[li], osrm, 2e-2)
    acts_c = O.vgg_forward(cimg, wts, pooling, 22)
    rep(f'[{pooling} {H}x{W}] content feat', ct.float().permute(2, 0, 1)[None], acts_c[22], 3e-2)
    m.set_targets(H, W, ct, 0.015, means, srms, st.style_weights, 2.0)
    torch.manual_seed(0)
    img = (cimg + 0.05 * torch.randn_like(cimg)).clamp(0, 1)
    st.image = img.to(dev).contiguous()
    terms, grad = st.loss_and_grad()
    # oracle with the oracle's own targets (fp32) and with bf16 storage model
    style_t = [O.StyleTarget.build(*O.style_stats(acts_s[layer])) for layer in O.STYLE_LAYERS]
    tg = O.ScaleTargets(acts_c[22], style_t, 0.015, 2.0)
    for sim in sims:
        if sim:
            a_s = O.vgg_forward(simg, wts, pooling, 29, True)
            a_c = O.vgg_forward(cimg, wts, pooling, 22, True)
            tg = O.ScaleTargets(a_c[22], [O.StyleTarget.build(*O.style_stats(a_s[layer])) for layer in O.STYLE_LAYERS],
                                0.015, 2.0)
        det = {}
        ol, og = O.loss_and_grad(img, wts, tg, pooling, sim_bf16=sim, detail=det)
        tag = 'bf16sim' if sim else 'fp32'
        print(f'   native terms {[f"{v:.6f}" for v in terms.tolist()]}')
        print(f'   oracle terms({tag}) loss={float(ol):.6f} {[f"{v:.6f}" for v in det["terms"]]}')
        rep(f'[{pooling} {H}x{W}] loss vs oracle {tag}', terms[0:1], ol.reshape(1), 1e-3 if not sim else 3e-4)
        cos = F.cosine_similarity(grad.cpu().flatten(), og.flatten(), dim=0).item()
        rel = ((grad.cpu() - og).norm() / og.norm()).item()
        print(f'   grad cos={cos:.6f} relL2={rel:.4f} ({tag})', flush=True)
    return st


def t_stylize(wts, pooling='max'):
    st = make_st(pooling, wts)
    c = O.synth_image(1, 16, 64, 48)
    s = [O.synth_image(2, 32, 56, 40), O.synth_image(3, 32, 40, 60)]
    kw = dict(min_scale=32, end_scale=64, iterations=4, initial_iterations=6, style_weights=[3, 1])
    tr = []
    st.stylize(c, s, callback=lambda it: tr.append(it.loss), **kw)
    tr2 = []
    O.stylize(c, s, wts, pooling=pooling, callback=lambda si, i, l, s_: tr2.append(l), **kw)
    print('   native', [f'{v:.5f}' for v in tr])
    print('   oracle', [f'{v:.5f}' for v in tr2])
    rel = max(abs(a - b) / abs(b) for a, b in zip(tr, tr2))
    global ALL_OK
    ok = rel < 5e-3
    ALL_OK &= ok
    print(f'[{"OK " if ok else "BAD"}] stylize multi-scale loss trace max rel diff {rel:.3e}', flush=True)


def t_bench(size, wts, iters=10):
    st = make_st('max', wts)
    m = st.model
    H = W = size
    m.ensure_workspace([(H, W)])
    cimg = O.to_tensor(O.synth_image(1, 16, W, H)).to(dev)
    simg = O.to_tensor(O.synth_image(2, 32, W, H)).to(dev)
    ct = m.content_features(cimg)
    means, srms = m.style_stats(simg)
    m.set_targets(H, W, ct, 0.015, means, srms, st.style_weights, 2.0)
    st.image = cimg.clone()
    st.average = stb.style_transfer.EMA(st.image, 0.99)
    ea, eas = torch.zeros_like(st.image), torch.zeros_like(st.image)
    for i in range(3):
        st._iterate(ea, eas, i + 1, 0.02, 0.99, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        st._iterate(ea, eas, i + 4, 0.02, 0.99, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f'bench {size}x{size}: {ms:.3f} ms/iter  {1000 / ms:.2f} it/s  loss={st._loss_host[0].item():.6f} '
          f'conv-TFLOP/s={1.444608e6 * size * size / ms / 1e9:.1f}  mem={torch.cuda.max_memory_allocated() / 2**30:.2f} GiB',
          flush=True)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1:] or ['kern', 'full', 'stylize', 'bench']
    wts = O.make_vgg_weights(1234)
    if 'kern' in which:
        t_conv0(16, 16)
        t_conv0(37, 70)
        t_conv0(64, 130)
        t_pool(16, 16, 64)
        t_pool(37, 21, 128)
        t_gram(16, 512)
        t_gram(1000, 64)
        t_gram(4096, 128)
        t_gram(3001, 256)
        t_gram(5000, 512)
        t_w2(64)
        t_w2(256)
        t_w2(512)
    if 'full' in which:
        t_full(48, 64, 'max', wts)
        t_full(56, 80, 'average', wts)
        t_full(72, 72, 'l2', wts)
    if 'big' in which:  # parity of the loss terms against the fp32 CPU oracle at larger sizes (minutes of CPU time)
        t_full(512, 512, 'max', wts)
        t_full(1024, 1024, 'max', wts)
    if 'huge' in which:  # the BASELINE.json size itself against the fp32 oracle (about two minutes of host time)
        t_full(2048, 2048, 'max', wts, sims=(False,))
    if 'stylize' in which:
        t_stylize(wts)
    if 'bench' in which:
        for sz in (256, 512, 1024, 2048):
            t_bench(sz, wts)
    print('ALL', 'PASS' if ALL_OK else 'FAIL')
