// Internal launcher prototypes (all stream-ordered, no host sync).  Layout conventions:
//   activations / feature gradients : NHWC bf16, i.e. [H][W][C] with C contiguous ("pixel-major")
//   image, Adam moments, EMA        : the reference's own NCHW fp32 [1,3,H,W] torch tensors
//   conv weights                    : packed bf16 [tap][N][K] (K contiguous), see pack_weights_*
#pragma once
#include "host_util.h"

namespace stb {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- tcgen05 implicit-GEMM "pixel GEMM"
// out[p][n] = epilogue( sum_{tap,k} A[p + off(tap)][k] * B[tap][n][k]  +  sum_k A2[p][k] * B2[n][k] )
//   fwd  (mode 0): + bias[n], ReLU                       (VGG conv 3x3 + bias + ReLU; ST:86-89 -> torchvision vgg.py)
//   bwd  (mode 1): + bias[n] (rows in [row_lo,row_hi)), + cscale*(y - ctarget), * (y > 0)
//                                                         (conv dgrad + tap-gradient GEMM + ReLU mask; autograd of ST:475)
struct PixelGemmArgs {
  int H = 0, W = 0;
  int Cin = 0;    // main 3x3 source channels (multiple of 64) or 0
  int Cout = 0;   // output channels (multiple of 64)
  int C2 = 0;     // second (1x1) source channels (multiple of 64) or 0
  int mode = 0;
  const bf16* A = nullptr;        // [H][W][Cin]
  const bf16* Bw = nullptr;       // [9][Cout][Cin]
  const bf16* A2 = nullptr;       // [H2][W][C2]  (rows a2_row0 .. a2_row0+H2 of the output grid)
  int a2_row0 = 0, a2_rows = 0;   // row window in which the second source contributes (multi-GPU own rows)
  const bf16* B2 = nullptr;       // [Cout][C2]
  bf16* out = nullptr;            // [H][W][Cout]
  const float* bias = nullptr;    // [Cout] or null
  const bf16* mask_src = nullptr; // bwd: [H][W][Cout]
  const bf16* ctarget = nullptr;  // bwd, optional: [H][W][Cout]
  float cscale = 0.f;
  int row_lo = 0, row_hi = 1 << 30;  // rows where bias (bwd) / content term apply
};
int launch_pixel_gemm(const PixelGemmArgs& a, cudaStream_t stream);

// fp32 OIHW [Cout][Cin][3][3] -> bf16 [9][Cout][Cin] (fwd) / [9][Cin][Cout] with 180-degree rotated taps (dgrad)
int pack_weights_fwd(const float* w, bf16* out, int Cout, int Cin, cudaStream_t s);
int pack_weights_bwd(const float* w, bf16* out, int Cout, int Cin, cudaStream_t s);

}  // namespace stb
