"""Shared helpers of the -m gpu parity tests (all calls go through the C-ABI of libstb200.so)."""
import ctypes

import torch

import style_transfer_b200 as stb
from style_transfer_b200 import _lib

DEV = torch.device('cuda:0')


def lib():
    """libstb200_test.so: kernel-level hooks (include/stb200_test.h).  The product library is reached through the
    `style_transfer_b200` package (gpu_util.make_st / st.model.lib)."""
    return _lib.load_test()


P = _lib.ptr
S = _lib.cur_stream


def check(rc):
    _lib.check(rc, test_lib=True)


def rel_err(got, ref):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert torch.isfinite(got).all(), 'non-finite values in the CUDA result'
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-30)).item()


def nchw(x_hwc):
    return x_hwc.float().permute(2, 0, 1)[None]


def pack(w, bwd):
    co, ci = w.shape[:2]
    out = torch.empty(9 * co * ci, dtype=torch.bfloat16, device=DEV)
    check(lib().stb_pack_weights(P(w), P(out), co, ci, int(bwd), S()))
    return out


def pixel_gemm(H, W, Cin, Cout, C2, mode, A=None, Bw=None, A2=None, a2_row0=0, a2_rows=0, B2=None, bias=None,
               mask=None, ctarget=None, cscale=0.0, row_lo=0, row_hi=1 << 30):
    out = torch.full((H, W, Cout), float('nan'), dtype=torch.bfloat16, device=DEV)
    check(lib().stb_test_pixel_gemm(H, W, Cin, Cout, C2, mode, P(A), P(Bw), P(A2), a2_row0, a2_rows, P(B2), P(out),
                                    P(bias), P(mask), P(ctarget), ctypes.c_float(cscale), row_lo, row_hi, S()))
    torch.cuda.synchronize()
    return out


def make_st(pooling, wts):
    return stb.StyleTransfer(devices=['cuda:0'], pooling=pooling, vgg_weights=wts)
