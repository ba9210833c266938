/* libstb200_test -- kernel-level test hooks (NOT part of the product library).
 *
 * libstb200_test.so is built from the same translation units as libstb200.so plus csrc/api_test.cu; each hook wraps
 * exactly one internal launcher so that tests/test_gpu_kernels.py can check every kernel against the oracle in
 * isolation.  The product library (include/stb200.h) exports none of these.
 */
#ifndef STB200_TEST_H_
#define STB200_TEST_H_

#include "stb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* message of the last failing hook on this thread (the test library keeps its own buffer) */
STB_API const char* stb_test_last_error(void);
STB_API int stb_pack_weights(const float* w_oihw, void* out_bf16, int Cout, int Cin, int bwd, void* stream);
STB_API int stb_test_pixel_gemm(int H, int W, int Cin, int Cout, int C2, int mode, const void* A, const void* Bw,
                                const void* A2, int a2_row0, int a2_rows, const void* B2, void* out,
                                const float* bias, const void* mask_src, const void* ctarget, float cscale,
                                int row_lo, int row_hi, void* stream);
STB_API int stb_test_conv0_fwd(const float* img, const float* w0, const float* b0, void* out_bf16, int H, int W,
                               float tv_weight, float* gtv, float* tv_partials, int* n_partials, void* stream);
STB_API int stb_test_conv0_bwd(const void* g0_bf16, const float* w0, const float* gtv, float* grad_out, int H, int W,
                               void* stream);
/* conv 3x3 + bias + ReLU with the 2x2 pool fused into its epilogue (the product's only pool-forward path):
 * out [H][W][Cout] and pool_out [H/2][W/2][Cout], both bf16 NHWC. */
STB_API int stb_test_conv_pool(int H, int W, int Cin, int Cout, const void* A, const void* Bw, const float* bias,
                               void* out, void* pool_out, int pooling, void* stream);
/* pool backward (+ ReLU mask of the pool input y): gin [H][W][C] from gout [H/2][W/2][C]. */
STB_API int stb_test_pool_bwd(int pooling, const void* gout, const void* y, void* gin, int H, int W, int C,
                              void* stream);
STB_API int stb_test_gram(const void* F_bf16, long P, int C, float* partials_ws, size_t partials_floats,
                          float* S_raw, float* sums, void* stream);
STB_API size_t stb_test_gram_partials_floats(long P, int C);
STB_API int stb_test_w2(const float* mean_t, const float* srm_t, const float* S_raw, const float* sums, int C,
                        float npix, float weight, void* ws, size_t ws_bytes, float* loss_out, float* gs_out,
                        float* gmu_out, float* csqrt_out, void* stream);
STB_API size_t stb_test_w2_workspace_bytes(void);

#ifdef __cplusplus
}
#endif
#endif /* STB200_TEST_H_ */
