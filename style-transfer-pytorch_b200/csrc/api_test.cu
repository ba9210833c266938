// Kernel-level C-ABI entry points used by tests/ (each wraps exactly one launcher so that the parity tests can
// check every kernel against the oracle in isolation).  Declared in include/stb200_test.h; linked ONLY into
// libstb200_test.so -- the product library libstb200.so does not contain this file.
#include "kernels.h"
#include "stb200_test.h"

using namespace stb;

extern "C" {

const char* stb_test_last_error(void) { return last_error_string().c_str(); }

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

extern "C" {

int stb_test_conv0_fwd(const float* img, const float* w0, const float* b0, void* out_bf16, int H, int W,
                       float tv_weight, float* gtv, float* tv_partials, int* n_partials, void* stream) {
  // product path of conv0: TV kernel and the fused im2col + tcgen05 GEMM + bias + ReLU kernel
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (gtv != nullptr) STB_TRY(launch_tv(img, H, W, 0, H, H, tv_weight, gtv, tv_partials, n_partials, s));
  bf16* wp = nullptr;
  STB_CUDA_CHECK(cudaMalloc(&wp, 64 * 64 * 2));
  int rc = pack_weights_conv0_fwd(w0, wp, s);
  if (rc == 0) rc = launch_conv0_fwd(img, wp, b0, static_cast<bf16*>(out_bf16), H, W, s);
  cudaStreamSynchronize(s);
  cudaFree(wp);
  return rc;
}

int stb_test_conv0_bwd(const void* g0_bf16, const float* w0, const float* gtv, float* grad_out, int H, int W,
                       void* stream) {
  // product path: interior via the tcgen05 1x1 GEMM + col2im kernel with the image-space epilogue, borders in SIMT
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bf16* wp = nullptr;
  STB_CUDA_CHECK(cudaMalloc(&wp, 32 * 64 * 2));
  int rc = pack_weights_conv0_bwd(w0, wp, s);
  if (rc == 0)
    rc = launch_conv0_bwd_interior(static_cast<const bf16*>(g0_bf16), wp, gtv, nullptr, nullptr, nullptr, nullptr,
                                   grad_out, H, W, nullptr, 0, s);
  if (rc == 0)
    rc = launch_conv0_bwd_adam(static_cast<const bf16*>(g0_bf16), true, w0, gtv, nullptr, nullptr, nullptr, nullptr,
                               grad_out, H, W, nullptr, 0, s);
  cudaStreamSynchronize(s);
  cudaFree(wp);
  return rc;
}

int stb_test_conv_pool(int H, int W, int Cin, int Cout, const void* A, const void* Bw, const float* bias, void* out,
                       void* pool_out, int pooling, void* stream) {
  PixelGemmArgs a;
  a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.mode = 0;
  a.A = static_cast<const bf16*>(A);
  a.Bw = static_cast<const bf16*>(Bw);
  a.out = static_cast<bf16*>(out);
  a.bias = bias;
  a.pool_out = static_cast<bf16*>(pool_out);
  a.pooling = pooling;
  return launch_pixel_gemm(a, static_cast<cudaStream_t>(stream));
}

int stb_test_pool_bwd(int pooling, const void* gout, const void* y, void* gin, int H, int W, int C, void* stream) {
  return launch_pool_bwd(pooling, static_cast<const bf16*>(gout), static_cast<const bf16*>(y), static_cast<bf16*>(gin),
                         H, W, C, static_cast<cudaStream_t>(stream));
}

size_t stb_test_gram_partials_floats(long P, int C) { return gram_partials_floats(P, C); }

int stb_test_gram(const void* F_bf16, long P, int C, float* partials_ws, size_t partials_floats, float* S_raw,
                  float* sums, void* stream) {
  STB_CHECK(partials_floats >= gram_partials_floats(P, C), STB_ERR_WORKSPACE, "gram partials workspace too small");
  return launch_gram(static_cast<const bf16*>(F_bf16), P, C, partials_ws, partials_floats, S_raw, sums,
                     static_cast<cudaStream_t>(stream));
}

size_t stb_test_w2_workspace_bytes(void) { return W2Engine::workspace_bytes(); }

// Runs the W2 engine with the given problem placed in the slot of matching size (other slots get a benign
// identity problem).  Outputs: weighted loss, Gs = G + G^T (fp32, NOT divided by npix), gmu (not divided),
// and the target's sqrtm(cov_t).
int stb_test_w2(const float* mean_t, const float* srm_t, const float* S_raw, const float* sums, int C, float npix,
                float weight, void* ws, size_t ws_bytes, float* loss_out, float* gs_out, float* gmu_out,
                float* csqrt_out, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static W2Engine eng;  // test-only
  const int ns[5] = {64, 128, 256, 512, 512};
  STB_TRY(eng.init(ws, ws_bytes, ns));
  int slot = -1;
  for (int l = 0; l < 5; ++l)
    if (ns[l] == C) { slot = l; break; }
  STB_CHECK(slot >= 0, STB_ERR_INVALID, "C must be one of 64,128,256,512");
  // benign content for the other slots: srm = I, mean = 0, S_raw = I, sums = 0, npix = 1
  for (int l = 0; l < 5; ++l) {
    W2Layer& L = eng.host_layers[l];
    const int n = L.n;
    std::vector<float> eye((size_t)n * n, 0.f), zero(n, 0.f);
    for (int i = 0; i < n; ++i) eye[(size_t)i * n + i] = 1.f;
    L.weight = 1.f; L.npix = 1.f;
    L.S_raw = L.X1;   // scratch matrices that the forward does not touch before reading S_raw
    L.sums = L.gmu_bias;
    STB_CUDA_CHECK(cudaMemcpyAsync(L.srm_t, eye.data(), eye.size() * 4, cudaMemcpyHostToDevice, s));
    STB_CUDA_CHECK(cudaMemcpyAsync(L.S_raw, eye.data(), eye.size() * 4, cudaMemcpyHostToDevice, s));
    STB_CUDA_CHECK(cudaMemcpyAsync(L.mean_t, zero.data(), n * 4, cudaMemcpyHostToDevice, s));
    STB_CUDA_CHECK(cudaMemcpyAsync(L.sums, zero.data(), n * 4, cudaMemcpyHostToDevice, s));
    STB_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  W2Layer& T = eng.host_layers[slot];
  T.weight = weight; T.npix = npix;
  STB_CUDA_CHECK(cudaMemcpyAsync(T.srm_t, srm_t, (size_t)C * C * 4, cudaMemcpyDeviceToDevice, s));
  STB_CUDA_CHECK(cudaMemcpyAsync(T.mean_t, mean_t, (size_t)C * 4, cudaMemcpyDeviceToDevice, s));
  STB_CUDA_CHECK(cudaMemcpyAsync(T.S_raw, S_raw, (size_t)C * C * 4, cudaMemcpyDeviceToDevice, s));
  STB_CUDA_CHECK(cudaMemcpyAsync(T.sums, sums, (size_t)C * 4, cudaMemcpyDeviceToDevice, s));
  STB_TRY(eng.upload_layers(s));
  STB_TRY(eng.build_targets(s));
  STB_TRY(eng.forward_backward(T.scal + 32, s));
  STB_CUDA_CHECK(cudaMemcpyAsync(loss_out, T.scal + W2S_LOSS, 4, cudaMemcpyDeviceToDevice, s));
  STB_CUDA_CHECK(cudaMemcpyAsync(gs_out, T.Gs, (size_t)C * C * 4, cudaMemcpyDeviceToDevice, s));
  STB_TRY(W2Engine::read_matrix(csqrt_out, T.P, C, s));
  // gmu_bias holds gmu / npix
  STB_CUDA_CHECK(cudaMemcpyAsync(gmu_out, T.gmu_bias, (size_t)C * 4, cudaMemcpyDeviceToDevice, s));
  return STB_OK;
}

}  // extern "C"
