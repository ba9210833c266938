"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A CPU restatement of the per-iteration hot path of the reference's ``StyleTransfer.stylize()``
(/root/reference/style_transfer/style_transfer.py = "ST", sqrtm.py = "SQ"), written as an *explicit schedule*:
manual forward, manual backward formulas (no autograd), manual Adam / clamp / EMA.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module, and
only as the checker / CPU baseline.  The product (style-transfer-pytorch_b200) never imports it.

Parity pin: the reference ships no tests or golden vectors, so this oracle is pinned against
outputs of the *unmodified reference itself*, run in the build container by ``tests/golden/make_golden.py`` with a
seeded synthetic VGG-19 (``make_vgg_weights``) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them.

Third-party arithmetic the reference reaches (not under /root/reference; effective pin = this image's
torch 2.11.0 / torchvision 0.26.0): conv2d, max/avg/LP pool, matmul, Adam.  Their published algorithms are restated
here; ``torch.nn.functional.conv2d`` / ``conv_transpose2d`` are used as the dense-contraction primitive (the same
oneDNN kernels the reference's CPU path reaches), and are themselves cross-checked against the plain-C loops of
``oracle/st_oracle_c.c`` in ``tests/test_oracle_c.py``.
"""
from __future__ import annotations

import copy
import math
import warnings
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

# torchvision vgg19().features[:30] topology (torchvision/models/vgg.py:73-94, cfg "E"); ST:35.
CONV_IDX = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28]
CONV_CH = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256), (256, 256), (256, 512),
           (512, 512), (512, 512), (512, 512), (512, 512)]
POOL_IDX = [4, 9, 18, 27]
NUM_LAYERS = 30
STYLE_LAYERS = [1, 6, 11, 20, 29]      # ST:317
CONTENT_LAYERS = [22]                  # ST:316
STYLE_LAYER_WEIGHTS = [w / 341 for w in (256, 64, 16, 4, 1)]  # ST:320-322
NORM_MEAN = (0.485, 0.456, 0.406)      # ST:30-31
NORM_STD = (0.229, 0.224, 0.225)
POOL_SCALE = {'max': 1.0, 'average': 2.0, 'l2': 0.78}  # ST:22


def layer_kind(i: int) -> str:
    if i in CONV_IDX:
        return 'conv'
    if i in POOL_IDX:
        return 'pool'
    return 'relu'


def make_vgg_weights(seed: int = 1234):
    """Deterministic synthetic VGG-19 conv weights (He init, small biases) from numpy's PCG64, so the same
    fixture can be regenerated anywhere (no network for the ImageNet checkpoint)."""
    rng = np.random.default_rng(seed)
    out = []
    for cin, cout in CONV_CH:
        w = rng.standard_normal((cout, cin, 3, 3), dtype=np.float32) * np.float32(math.sqrt(2.0 / (9 * cin)))
        b = rng.standard_normal((cout,), dtype=np.float32) * np.float32(0.05)
        out.append((torch.from_numpy(w), torch.from_numpy(b)))
    return out


def synth_image(seed: int, base: int, w: int, h: int):
    """Low-frequency synthetic PIL image."""
    from PIL import Image
    arr = np.random.default_rng(seed).integers(0, 256, (base, base, 3), dtype=np.uint8)
    return Image.fromarray(arr).resize((w, h), Image.BICUBIC)


# ----------------------------------------------------------------------------------------------- rounding model
def _q(x: torch.Tensor, sim_bf16: bool) -> torch.Tensor:
    """Optional model of the CUDA path's storage rounding (activations / feature gradients kept in bf16)."""
    return x.bfloat16().to(x.dtype) if sim_bf16 else x


# ----------------------------------------------------------------------------------------------- pooling
def pool_fwd(x, pooling):
    """2x2 stride-2 floor-mode pooling incl. the reference's Scale wrapper (ST:21-22, 41-46)."""
    if pooling == 'max':
        return F.max_pool2d(x, 2)
    if pooling == 'average':
        return F.avg_pool2d(x, 2) * 2.0
    if pooling == 'l2':
        # torch/nn/functional.py lp_pool2d: avg_pool(x^2) -> sign*relu(abs) -> *k -> pow(1/2)
        s = F.avg_pool2d(x * x, 2) * 4.0
        return s.sqrt() * 0.78
    raise ValueError(pooling)


def pool_bwd(g, x, pooling):
    """Manual backward of pool_fwd w.r.t. x.  Max: first maximum in window scan order (0,0),(0,1),(1,0),(1,1)
    wins (ATen max_pool2d_with_indices); rows/cols dropped by floor mode get zero gradient."""
    n, c, h, w = x.shape
    ho, wo = h // 2, w // 2
    gx = torch.zeros_like(x)
    xs = [x[:, :, i:2 * ho:2, j:2 * wo:2] for i in (0, 1) for j in (0, 1)]
    if pooling == 'max':
        m = torch.maximum(torch.maximum(xs[0], xs[1]), torch.maximum(xs[2], xs[3]))
        taken = torch.zeros_like(m, dtype=torch.bool)
        k = 0
        for i in (0, 1):
            for j in (0, 1):
                sel = (xs[k] == m) & ~taken
                gx[:, :, i:2 * ho:2, j:2 * wo:2] = torch.where(sel, g, torch.zeros_like(g))
                taken |= sel
                k += 1
    elif pooling == 'average':
        for i in (0, 1):
            for j in (0, 1):
                gx[:, :, i:2 * ho:2, j:2 * wo:2] = g * 0.5
    elif pooling == 'l2':
        s = (xs[0] ** 2 + xs[1] ** 2 + xs[2] ** 2 + xs[3] ** 2).sqrt()
        inv = torch.where(s > 0, 0.78 / s, torch.zeros_like(s))
        k = 0
        for i in (0, 1):
            for j in (0, 1):
                gx[:, :, i:2 * ho:2, j:2 * wo:2] = g * xs[k] * inv
                k += 1
    else:
        raise ValueError(pooling)
    return gx


# ----------------------------------------------------------------------------------------------- VGG trunk
def vgg_forward(image, weights, pooling='max', last_layer=29, sim_bf16=False):
    """ST:78-90.  Returns list `acts` with acts[i] = output of features[i] (ReLU is in place in the reference, so
    the conv slot holds nothing separately: acts[conv] is None and acts[relu] is the post-ReLU tensor)."""
    h, w = image.shape[2:]
    min_size = 1
    for layer in (4, 9, 18, 27, 36):  # ST:61-69
        if last_layer < layer:
            break
        min_size *= 2
    if min(h, w) < min_size:
        raise ValueError(f'Input is {h}x{w} but must be at least {min_size}x{min_size}')
    mean = torch.tensor(NORM_MEAN, dtype=image.dtype).view(1, 3, 1, 1)
    std = torch.tensor(NORM_STD, dtype=image.dtype).view(1, 3, 1, 1)
    x = (image - mean) / std  # ST:85 (sub then div)
    acts = [None] * (last_layer + 1)
    ci = 0
    for i in range(last_layer + 1):
        kind = layer_kind(i)
        if kind == 'conv':
            wt, b = weights[ci]
            wt, b = wt.to(x.dtype), b.to(x.dtype)
            if sim_bf16:
                wt = _q(wt, True)  # every conv (conv0 included) takes bf16 weights on the tensor cores
            if i == 0:
                x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='replicate'), wt, b)  # ST:39, 52-59
            else:
                x = F.conv2d(x, wt, b, padding=1)
            ci += 1
        elif kind == 'relu':
            x = _q(torch.relu(x), sim_bf16)
            acts[i] = x
        else:
            x = _q(pool_fwd(x, pooling), sim_bf16)
            acts[i] = x
    return acts


def vgg_backward(tap_grads, acts, weights, pooling='max', sim_bf16=False):
    """Manual autograd of vgg_forward down to the (un-normalised) image: conv dgrad only (weights frozen, ST:49),
    ReLU mask from the saved *output*, pool backward, tap gradients added where they were read (ST:475)."""
    last = max(tap_grads)
    g = None
    ci = sum(1 for c in CONV_IDX if c <= last) - 1
    for i in range(last, -1, -1):
        kind = layer_kind(i)
        if i in tap_grads:
            g = tap_grads[i] if g is None else g + tap_grads[i]
        if kind == 'relu':
            g = _q(g * (acts[i] > 0), sim_bf16)
        elif kind == 'pool':
            g = _q(pool_bwd(g, acts[i - 1], pooling), sim_bf16)
        else:
            wt = weights[ci][0].to(g.dtype)
            if i == 0:
                gp = F.conv_transpose2d(g, _q(wt, sim_bf16))  # gradient on the replicate-padded grid (H+2, W+2)
                # fold the pad ring back onto the border pixels (adjoint of replicate padding)
                gp[:, :, 1, :] += gp[:, :, 0, :]
                gp[:, :, -2, :] += gp[:, :, -1, :]
                gp[:, :, :, 1] += gp[:, :, :, 0]
                gp[:, :, :, -2] += gp[:, :, :, -1]
                g = gp[:, :, 1:-1, 1:-1]
            else:
                if sim_bf16:
                    wt = _q(wt, True)
                g = F.conv_transpose2d(g, wt, padding=1)
            ci -= 1
    std = torch.tensor(NORM_STD, dtype=g.dtype).view(1, 3, 1, 1)
    return g / std


# ----------------------------------------------------------------------------------------------- sqrtm (SQ:9-47)
def sqrtm_ns(a, num_iters=12):
    """Newton-Schulz matrix square root, SQ:9-25."""
    norm_a = a.pow(2).sum().sqrt()
    y = a / norm_a
    n = a.shape[-1]
    eye3 = torch.eye(n, dtype=a.dtype) * 3
    z = torch.eye(n, dtype=a.dtype)
    for _ in range(num_iters):
        t = (eye3 - z @ y) / 2
        y = y @ t
        z = t @ z
    return y * norm_a.sqrt()


def sqrtm_ns_lyap_backward(z, grad_output, num_iters=12):
    """Iterative Lyapunov-equation backward of the square root, SQ:36-47 (uses only the saved output z)."""
    norm_z = z.pow(2).sum().sqrt()
    a = z / norm_z
    n = z.shape[-1]
    eye3 = torch.eye(n, dtype=z.dtype) * 3
    q = grad_output / norm_z
    for i in range(num_iters):
        eye_a_a = eye3 - a @ a
        q = (q @ eye_a_a - a.t() @ (a.t() @ q - q @ a)) / 2
        if i < num_iters - 1:
            a = a @ eye_a_a / 2
    return q / 2


# ----------------------------------------------------------------------------------------------- losses
def style_stats(feat):
    """ST:163-168: mean [C] and second raw moment [C,C] of a [1,C,H,W] activation."""
    c = feat.shape[1]
    f = feat.reshape(c, -1)
    n = f.shape[1]
    return f.mean(1), (f @ f.t()) / n


@dataclass
class StyleTarget:
    """ST:152-160: buffers of StyleLossW2 built from the blended (mean, srm) target."""
    mean: torch.Tensor
    cov: torch.Tensor
    cov_sqrt: torch.Tensor

    @staticmethod
    def build(mean, srm, eps=1e-4):
        cov = srm - torch.outer(mean, mean) + torch.eye(mean.numel(), dtype=mean.dtype) * eps
        return StyleTarget(mean, cov, sqrtm_ns(cov, 12))


def w2_loss_and_grad(feat, tgt: StyleTarget, eps=1e-4):
    """ST:175-181 forward and its manual backward w.r.t. the activation `feat` [1,C,H,W] (unit upstream grad)."""
    c = feat.shape[1]
    f = feat.reshape(c, -1)
    n = f.shape[1]
    mean = f.mean(1)
    srm = (f @ f.t()) / n
    eye = torch.eye(c, dtype=feat.dtype)
    cov = srm - torch.outer(mean, mean) + eye * eps
    mean_diff = ((mean - tgt.mean) ** 2).mean()
    m = tgt.cov_sqrt @ cov @ tgt.cov_sqrt
    r = sqrtm_ns(m, 12)
    cov_diff = torch.diagonal(tgt.cov + cov - 2 * r).mean()
    loss = mean_diff + cov_diff
    # backward
    g_r = eye * (-2.0 / c)
    g_m = sqrtm_ns_lyap_backward(r, g_r, 12)
    g_cov = tgt.cov_sqrt.t() @ g_m @ tgt.cov_sqrt.t() + eye / c
    g_mean = 2.0 * (mean - tgt.mean) / c - (g_cov + g_cov.t()) @ mean
    g_f = ((g_cov + g_cov.t()) @ f) / n + g_mean[:, None] / n
    return loss, g_f.reshape(feat.shape), dict(mean=mean, srm=srm, cov=cov, sqrt=r, g_cov=g_cov, g_mean=g_mean)


def content_loss_and_grad(feat, target):
    """ST:119-126 nn.MSELoss (mean) and its gradient."""
    d = feat - target
    return (d * d).mean(), 2.0 * d / d.numel()


def tv_loss_and_grad(x):
    """ST:184-195: nine-point L2 TV on the raw image, with its manual gradient."""
    p = F.pad(x, (1, 1, 1, 1), mode='replicate')
    s1, s2 = slice(1, -1), slice(2, None)
    s3, s4 = slice(None, -1), slice(1, None)
    e1 = p[..., s1, s2] - p[..., s1, s1]
    e2 = p[..., s2, s1] - p[..., s1, s1]
    e3 = p[..., s4, s4] - p[..., s3, s3]
    e4 = p[..., s4, s3] - p[..., s3, s4]
    loss = 2 * (e1.pow(2).mean() / 3 + e2.pow(2).mean() / 3 + e3.pow(2).mean() / 12 + e4.pow(2).mean() / 12)
    gp = torch.zeros_like(p)
    k1 = 2 * 2 / (3 * e1.numel())
    k3 = 2 * 2 / (12 * e3.numel())
    gp[..., s1, s2] += k1 * e1
    gp[..., s1, s1] -= k1 * e1
    gp[..., s2, s1] += k1 * e2
    gp[..., s1, s1] -= k1 * e2
    gp[..., s4, s4] += k3 * e3
    gp[..., s3, s3] -= k3 * e3
    gp[..., s4, s3] += k3 * e4
    gp[..., s3, s4] -= k3 * e4
    gp[..., 1, :] += gp[..., 0, :]
    gp[..., -2, :] += gp[..., -1, :]
    gp[..., :, 1] += gp[..., :, 0]
    gp[..., :, -2] += gp[..., :, -1]
    return loss, gp[..., 1:-1, 1:-1]


# ----------------------------------------------------------------------------------------------- one iteration
@dataclass
class ScaleTargets:
    content_target: torch.Tensor
    style: list  # [StyleTarget] per STYLE_LAYERS
    content_weight: float = 0.015
    tv_weight: float = 2.0


@dataclass
class IterState:
    image: torch.Tensor
    exp_avg: torch.Tensor
    exp_avg_sq: torch.Tensor
    step: int
    ema_value: torch.Tensor
    ema_accum: float

    @staticmethod
    def fresh(image, avg_decay=0.99):
        st = IterState(image.clone(), torch.zeros_like(image), torch.zeros_like(image), 0, torch.zeros_like(image), 1.0)
        st.ema_update(avg_decay)  # ST:245: ctor primes the average once
        return st

    def ema_update(self, decay):
        self.ema_accum *= decay
        self.ema_value = self.ema_value * decay + (1 - decay) * self.image

    def ema_get(self):
        return self.ema_value / (1 - self.ema_accum)


def loss_and_grad(image, weights, tg: ScaleTargets, pooling='max', sim_bf16=False, detail=None):
    """closure() of ST:472-476: loss (sum in ST:455 order) and d loss / d image."""
    acts = vgg_forward(image, weights, pooling, 29, sim_bf16)
    terms = []
    tap_grads = {}
    cl, cg = content_loss_and_grad(acts[22], tg.content_target)
    terms.append(cl * tg.content_weight)
    tap_grads[22] = cg * tg.content_weight
    for layer, lw, st in zip(STYLE_LAYERS, STYLE_LAYER_WEIGHTS, tg.style):
        sl, sg, _ = w2_loss_and_grad(acts[layer], st)
        terms.append(sl * lw)
        tap_grads[layer] = _q(sg * lw, sim_bf16) if layer != 22 else sg * lw
    tvl, tvg = tv_loss_and_grad(image)
    terms.append(tvl * tg.tv_weight)
    if sim_bf16:
        tap_grads[22] = tap_grads[22]  # content term is added in fp32 inside the dgrad epilogue
    g = vgg_backward(tap_grads, acts, weights, pooling, sim_bf16) + tvg * tg.tv_weight
    loss = terms[0]
    for t in terms[1:]:
        loss = loss + t  # python sum, left to right (ST:208)
    if detail is not None:
        detail['terms'] = [float(t) for t in terms]
        detail['acts'] = acts
    return loss, g


def iterate(st: IterState, weights, tg: ScaleTargets, pooling='max', lr=0.02, betas=(0.9, 0.99), adam_eps=1e-8,
            avg_decay=0.99, sim_bf16=False, detail=None):
    """One pass of ST:480-486 (Adam branch).  Returns the pre-update loss like `opt.step(closure)`."""
    loss, g = loss_and_grad(st.image, weights, tg, pooling, sim_bf16, detail)
    b1, b2 = betas
    st.step += 1
    st.exp_avg = st.exp_avg + (g - st.exp_avg) * (1 - b1)          # lerp_  (torch/optim/adam.py:413-546)
    st.exp_avg_sq = st.exp_avg_sq * b2 + (1 - b2) * g * g
    bc1 = 1 - b1 ** st.step
    bc2 = 1 - b2 ** st.step
    denom = st.exp_avg_sq.sqrt() / math.sqrt(bc2) + adam_eps
    st.image = (st.image - (lr / bc1) * st.exp_avg / denom).clamp(0, 1)  # ST:483-485
    st.ema_update(avg_decay)                                             # ST:486
    if detail is not None:
        detail['grad'] = g
    return float(loss)


# ----------------------------------------------------------------------------------------------- per-scale setup
def size_to_fit(size, max_dim, scale_up=False):  # ST:256-265
    w, h = size
    if not scale_up and max(h, w) <= max_dim:
        return w, h
    new_w, new_h = max_dim, max_dim
    if h > w:
        new_w = round(max_dim * w / h)
    else:
        new_h = round(max_dim * h / w)
    return new_w, new_h


def gen_scales(start, end):  # ST:268-276
    scale, i, scales = end, 0, set()
    while scale >= start:
        scales.add(scale)
        i += 1
        scale = round(end / pow(2, i / 2))
    return sorted(scales)


def _interp(x, size, mode):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', UserWarning)
        return F.interpolate(x, size, mode=mode)


def to_tensor(pil_image):
    arr = np.asarray(pil_image.convert('RGB'), dtype=np.uint8)
    return torch.from_numpy(arr.copy()).permute(2, 0, 1).float().div(255)[None].contiguous()


def make_targets(content_pil, style_pils, style_weights, scale, weights, pooling, content_weight, tv_weight,
                 style_scale_fac=1.0, style_size=None, sim_bf16=False):
    """ST:416-453 for one scale."""
    from PIL import Image
    cw, ch = size_to_fit(content_pil.size, scale, scale_up=True)
    content = to_tensor(content_pil.resize((cw, ch), Image.BICUBIC))
    ctarget = vgg_forward(content, weights, pooling, 22, sim_bf16)[22]
    acc = {}
    for i, im in enumerate(style_pils):
        if style_size is None:
            sw, sh = size_to_fit(im.size, round(scale * style_scale_fac))
        else:
            sw, sh = size_to_fit(im.size, style_size)
        acts = vgg_forward(to_tensor(im.resize((sw, sh), Image.BICUBIC)), weights, pooling, 29, sim_bf16)
        for layer in STYLE_LAYERS:
            mean, srm = style_stats(acts[layer])
            mean, srm = mean * style_weights[i], srm * style_weights[i]
            if layer not in acc:
                acc[layer] = [mean, srm]
            else:
                acc[layer][0] += mean
                acc[layer][1] += srm
    style = [StyleTarget.build(*acc[layer]) for layer in STYLE_LAYERS]
    return ScaleTargets(ctarget, style, content_weight, tv_weight), (cw, ch)


def stylize(content_pil, style_pils, weights, *, style_weights=None, content_weight=0.015, tv_weight=2.0,
            min_scale=128, end_scale=512, iterations=500, initial_iterations=1000, step_size=0.02, avg_decay=0.99,
            pooling='max', style_scale_fac=1.0, style_size=None, callback=None, sim_bf16=False):
    """Restatement of ST:349-499 for init='content', optimizer='adam'.  callback(scale_index, i, loss, state)."""
    from PIL import Image
    min_scale = min(min_scale, end_scale)
    if style_weights is None:
        style_weights = [1 / len(style_pils)] * len(style_pils)
    else:
        ws = sum(abs(w) for w in style_weights)
        style_weights = [w / ws for w in style_weights]
    if len(style_pils) != len(style_weights):
        raise ValueError('style_images and style_weights must have the same length')
    scales = gen_scales(min_scale, end_scale)
    cw, ch = size_to_fit(content_pil.size, scales[0], scale_up=True)
    image = to_tensor(content_pil.resize((cw, ch), Image.BICUBIC))
    st = None
    for si, scale in enumerate(scales):
        tg, (cw, ch) = make_targets(content_pil, style_pils, style_weights, scale, weights, pooling, content_weight,
                                    tv_weight, style_scale_fac, style_size, sim_bf16)
        image = _interp(image, (ch, cw), 'bicubic').clamp(0, 1)  # ST:420
        new = IterState.fresh(image, avg_decay)
        if st is not None:  # ST:285-295, 460-462: warm start, step counter carried over
            new.exp_avg = _interp(st.exp_avg, (ch, cw), 'bicubic')
            new.exp_avg_sq = _interp(st.exp_avg_sq, (ch, cw), 'bilinear').relu_()
            new.step = st.step
        st = new
        its = initial_iterations if si == 0 else iterations  # ST:478
        for i in range(1, its + 1):
            loss = iterate(st, weights, tg, pooling, lr=step_size, avg_decay=avg_decay, sim_bf16=sim_bf16)
            if callback is not None:
                callback(si, i, loss, st)
        image = st.ema_get()  # ST:496-497
        st.image = image
    return st.ema_get()[0].clamp(0, 1)  # ST:335-336
