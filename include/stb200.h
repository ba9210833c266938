/* libstb200 -- C ABI of the B200-native stylize() hot path.
 *
 * The reference (crowsonkb/style-transfer-pytorch) has no FFI layer: its hot path is the Python loop body of
 * StyleTransfer.stylize() (style_transfer/style_transfer.py:472-486, "ST" below).  This header is the seam a
 * maintainer binds directly beneath that class (see INTEGRATION.md for the ctypes stub).  Every entry point is
 * stream-ordered, borrows caller-owned device pointers (torch tensors) for the duration of the call, never throws,
 * never exits; it returns 0 on success or a negative STB_ERR_* code, with stb_last_error() giving the message.
 *
 * Data layouts
 *   image / exp_avg / exp_avg_sq / ema : fp32 NCHW [1,3,H,W]   (exactly the reference's tensors, ST:420-421, 457-463)
 *   activations inside the workspace   : bf16 NHWC
 *   style statistics                   : fp32, mean [C], second raw moment [C,C] row-major (ST:163-168)
 */
#ifndef STB200_H_
#define STB200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define STB_API __attribute__((visibility("default")))
#else
#define STB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define STB_OK 0
#define STB_ERR_INVALID (-1)   /* bad argument              -> ValueError   (ST:83, ST:331, ST:373, ST:405, ST:467) */
#define STB_ERR_CUDA (-2)      /* CUDA runtime/driver error -> RuntimeError                                          */
#define STB_ERR_WORKSPACE (-3) /* workspace unbound / small -> RuntimeError                                          */
#define STB_ERR_STATE (-4)     /* call-order violation      -> RuntimeError                                          */

#define STB_POOL_MAX 0     /* nn.MaxPool2d(2)                 ST:21 */
#define STB_POOL_AVERAGE 1 /* Scale(nn.AvgPool2d(2), 2.0)     ST:21-22, 41-46 */
#define STB_POOL_L2 2      /* Scale(nn.LPPool2d(2, 2), 0.78)  ST:21-22, 41-46 */

#define STB_NUM_CONVS 13      /* VGG-19 features[:30]: convs at 0,2,5,7,10,12,14,16,19,21,23,25,28 */
#define STB_NUM_STYLE_TAPS 5  /* ReLU outputs 1,6,11,20,29 (ST:317) */

typedef struct stb_ctx stb_ctx;

/* Thread-local message for the last failing call on this thread. */
STB_API const char* stb_last_error(void);

/* ------------------------------------------------------------------ kernel test hooks (used by tests/ only) */
STB_API int stb_pack_weights(const float* w_oihw, void* out_bf16, int Cout, int Cin, int bwd, void* stream);
STB_API int stb_test_pixel_gemm(int H, int W, int Cin, int Cout, int C2, int mode, const void* A, const void* Bw,
                                const void* A2, int a2_row0, int a2_rows, const void* B2, void* out,
                                const float* bias, const void* mask_src, const void* ctarget, float cscale,
                                int row_lo, int row_hi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STB200_H_ */
