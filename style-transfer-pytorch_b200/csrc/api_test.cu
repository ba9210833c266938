// Kernel-level C-ABI entry points used by tests/ (each wraps exactly one launcher so that the parity tests can
// check every kernel against the oracle in isolation).  Declared in include/stb200.h under "kernel test hooks".
#include "kernels.h"
#include "stb200.h"

using namespace stb;

extern "C" {

const char* stb_last_error(void) { return last_error_string().c_str(); }

int stb_pack_weights(const float* w_oihw, void* out_bf16, int Cout, int Cin, int bwd, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return bwd ? pack_weights_bwd(w_oihw, static_cast<bf16*>(out_bf16), Cout, Cin, s)
             : pack_weights_fwd(w_oihw, static_cast<bf16*>(out_bf16), Cout, Cin, s);
}

int stb_test_pixel_gemm(int H, int W, int Cin, int Cout, int C2, int mode, const void* A, const void* Bw,
                        const void* A2, int a2_row0, int a2_rows, const void* B2, void* out, const float* bias,
                        const void* mask_src, const void* ctarget, float cscale, int row_lo, int row_hi,
                        void* stream) {
  PixelGemmArgs a;
  a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.C2 = C2; a.mode = mode;
  a.A = static_cast<const bf16*>(A);
  a.Bw = static_cast<const bf16*>(Bw);
  a.A2 = static_cast<const bf16*>(A2);
  a.a2_row0 = a2_row0; a.a2_rows = a2_rows;
  a.B2 = static_cast<const bf16*>(B2);
  a.out = static_cast<bf16*>(out);
  a.bias = bias;
  a.mask_src = static_cast<const bf16*>(mask_src);
  a.ctarget = static_cast<const bf16*>(ctarget);
  a.cscale = cscale;
  a.row_lo = row_lo; a.row_hi = row_hi;
  return launch_pixel_gemm(a, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
