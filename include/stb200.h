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

#include <stddef.h>
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

/* ------------------------------------------------------------------ context (replaces ST:324-333)
 * conv_w[i] / conv_b[i]: DEVICE fp32 pointers to the 13 conv weights (OIHW) and biases of vgg19().features[:30] in
 * layer order.  The context packs them once into its own bf16 tensor-core layouts (the only memory it owns). */
STB_API int stb_ctx_create(int device, int pooling, const float* const* conv_w, const float* const* conv_b,
                           void* stream, stb_ctx** out);
STB_API void stb_ctx_destroy(stb_ctx* ctx);

/* ------------------------------------------------------------------ workspace (caller/torch owns all memory;
 * mirrors the reference's per-scale allocations, ST:413-414, 469-470).  Bytes needed to run any entry point on
 * an H x W image; bind a 1 KiB-aligned device block of at least the maximum over the sizes that will be used.
 * Rebinding invalidates the targets. */
STB_API int stb_workspace_bytes(stb_ctx* ctx, int H, int W, size_t* bytes);
STB_API int stb_bind_workspace(stb_ctx* ctx, void* ptr, size_t bytes, void* stream);

/* ------------------------------------------------------------------ target extraction (no-grad VGG forward)
 * stb_style_stats      = self.model(style, layers=style_layers) + StyleLossW2.get_target   (ST:440-443, 163-168)
 *                        mean_out[l]: [C_l] fp32, srm_out[l]: [C_l,C_l] fp32, l over taps 1,6,11,20,29
 * stb_content_features = self.model(content, layers=[22])[22]                               (ST:425)
 *                        target_out: bf16 NHWC [H/8][W/8][512]
 * `img` is a device fp32 NCHW [1,3,H,W] tensor with values in [0,1]; H,W >= 16 (ST:61-69, 82-83). */
STB_API int stb_style_stats(stb_ctx* ctx, const float* img, int H, int W, float* const* mean_out,
                            float* const* srm_out, void* stream);
STB_API int stb_content_features(stb_ctx* ctx, const float* img, int H, int W, void* target_out_bf16, void* stream);

/* ------------------------------------------------------------------ per-scale loss state (ST:426-455)
 * mean_t/srm_t: the style-weight-blended target moments (ST:443-450); the library forms cov = srm - mean mean^T
 * + eps I and cov_sqrt = sqrtm_ns(cov, 12) (ST:152-160).  style_w: the five layer weights (ST:320-322). */
STB_API int stb_set_targets(stb_ctx* ctx, int H, int W, const void* content_target_bf16, float content_weight,
                            const float* const* mean_t, const float* const* srm_t, const float* style_w,
                            float tv_weight, float eps, void* stream);

/* ------------------------------------------------------------------ THE HOT PATH: one iteration of ST:480-486
 * forward + losses + backward + Adam(lr, betas, eps; bias correction with `step` = 1-based count carried across
 * scales, ST:287-295/461-462) + clamp_(0,1) + EMA value update, all stream-ordered with no host sync.
 * loss_out_host8 (pinned host, optional) receives asynchronously {loss, content, style1..5, tv} of the
 * PRE-update image (what `opt.step(closure)` returns, ST:481). */
STB_API int stb_iterate(stb_ctx* ctx, float* img, float* exp_avg, float* exp_avg_sq, float* ema, int64_t step,
                        float lr, float beta1, float beta2, float adam_eps, float ema_decay, float* loss_out_host8,
                        void* stream);
/* closure-only variant (apply_update = 0): loss and d loss/d image (grad_out fp32 NCHW), e.g. for an L-BFGS host. */
STB_API int stb_iterate_ex(stb_ctx* ctx, float* img, float* exp_avg, float* exp_avg_sq, float* ema, int64_t step,
                           float lr, float beta1, float beta2, float adam_eps, float ema_decay, int apply_update,
                           float* grad_out, float* loss_out_host8, void* stream);

/* ------------------------------------------------------------------ spatial tiling across GPUs
 * A context may work on a horizontal band (plus halo aprons) of a taller image: H passed to the other calls is the
 * LOCAL height, rows [own_row0, own_row0+own_rows) (multiples of 16) are the band's own rows, H_global the full
 * height.  Only own rows enter the statistics, losses and tap gradients.  Per iteration the host runs
 *   stb_iterate_fwd  -> all-reduce(sum) of the stats block over the ranks (one NCCL call, ~2.4 MB)
 *   stb_iterate_bwd  -> grad_out = d loss / d (local image) incl. contributions to the halo rows
 *   [exchange + add the halo rows of grad_out with the neighbouring bands]
 *   stb_adam_update  on the own rows, then refresh the halo rows of the image from the neighbours.
 * With a band set, stb_style_stats returns RAW sums over the own rows (all-reduce, then divide by the global count). */
STB_API int stb_set_band(stb_ctx* ctx, int enabled, int H_global, int own_row0, int own_rows);
STB_API int stb_stats_block(stb_ctx* ctx, int H, int W, float** dev_ptr, size_t* n_floats);
STB_API int stb_iterate_fwd(stb_ctx* ctx, const float* img, void* stream);
STB_API int stb_iterate_bwd(stb_ctx* ctx, float* img, float* grad_out, float* loss_out_host8, void* stream);
STB_API int stb_adam_update(float* img, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int H, int W,
                            int row0, int rows, int64_t step, float lr, float beta1, float beta2, float adam_eps,
                            float ema_decay, void* stream);

/* ------------------------------------------------------------------ sync-free loss read-back (ST:487-493)
 * host_ring: PINNED host memory, slots x 16 floats (NULL switches it off).  Each updating iteration stores its eight
 * terms into slot (step % slots) and then the step as the slot's int32 stamp (word 8), from the loss kernel itself,
 * BEFORE the backward pass starts: the host polls the stamp -- the per-iteration callback needs no stream sync and
 * overlaps the rest of the iteration. */
STB_API int stb_set_loss_ring(stb_ctx* ctx, float* host_ring, int slots);

/* ------------------------------------------------------------------ per-scale warm start (ST:285-295, 420)
 * out[1,C,Ho,Wo] = F.interpolate(in[1,C,H,W], (Ho,Wo), mode, align_corners=False) on the device, fp32.
 * mode: 0 bilinear, 1 bicubic (A = -0.75).  post: 0 none, 1 relu (exp_avg_sq, ST:293), 2 clamp to [0,1] (image). */
STB_API int stb_resize(const float* in, int C, int H, int W, float* out, int Ho, int Wo, int mode, int post,
                       void* stream);

/* ------------------------------------------------------------------ tiled iteration, exchanges inside the library
 * (csrc/comm.cu).  Each rank owns a MAILBOX (iteration stamps, its statistics block, its image gradient, its first /
 * last 80 updated rows) that the peers map with CUDA IPC and read over NVLink; stb_iterate_banded is then the whole
 * iteration of a band -- halo pull, forward, all-reduce of the statistics, backward, seam reduce of the gradient fused
 * with Adam + clamp + EMA -- as ONE stream-ordered sequence (one CUDA graph), with no host call between its phases.
 *   stb_comm_create        allocate the own mailbox for bands up to max_h_local x max_W (same numbers on every rank);
 *                          ipc_handle_out64: 64-byte cudaIpcMemHandle_t to ship to the peers; mailbox_out: the pointer
 *   stb_comm_connect_ipc   handles: world x 64 bytes in rank order (one process per GPU)
 *   stb_comm_connect_local mailboxes[world]: device pointers of contexts living in THIS process (tests / emulation)
 *   stb_comm_set_geometry  per scale: local height, first own row, own rows of this band; the neighbours' local
 *                          heights and the first bottom-apron row of the upper one (all in their local coordinates)
 *   stb_comm_reset         zero the iteration stamps; the host barriers over all ranks before AND after
 * A peer that does not show up within 30 s makes the waiting kernel trap (CUDA error), it never hangs. */
STB_API int stb_comm_create(stb_ctx* ctx, int rank, int world, int max_h_local, int max_W, void* ipc_handle_out64,
                            void** mailbox_out);
STB_API int stb_comm_connect_ipc(stb_ctx* ctx, const void* handles);
STB_API int stb_comm_connect_local(stb_ctx* ctx, void* const* mailboxes);
STB_API int stb_comm_disconnect(stb_ctx* ctx);  /* unmap the peers' mailboxes (host barrier, then re-create) */
/* per-layer-halo mode (default of tiled runs): the band computes only its own rows of every layer and pulls the one row
 * above / below them from the neighbours' workspaces, which therefore are cudaMalloc blocks of the library mapped by
 * the neighbours (CUDA IPC).  stb_comm_alloc_workspace allocates, zeroes and binds (as stb_bind_workspace does) and
 * returns the 64-byte IPC handle / the pointer; stb_comm_connect_ws_* takes every rank's handle (world x 64 bytes) or
 * pointer in rank order; stb_comm_set_geometry(..., halo_rows = 1) then selects that mode for a scale (0: the band
 * recomputes 80-row aprons instead, no per-layer exchange -- cheaper when the layers are small);
 * stb_comm_release_workspace(unmap_only = 1) unmaps the neighbours, (0) also frees the own block (host barrier in
 * between). */
STB_API int stb_comm_alloc_workspace(stb_ctx* ctx, size_t bytes, void* ipc_handle_out64, void** ptr_out, void* stream);
STB_API int stb_comm_connect_ws_ipc(stb_ctx* ctx, const void* handles);
STB_API int stb_comm_connect_ws_local(stb_ctx* ctx, void* const* pointers);
STB_API int stb_comm_release_workspace(stb_ctx* ctx, int unmap_only);
STB_API int stb_comm_set_geometry(stb_ctx* ctx, int W, int h_local, int own0, int own_rows, int up_h_local,
                                  int up_apron_row0, int dn_h_local, int halo_rows);
STB_API int stb_comm_reset(stb_ctx* ctx, void* stream);
STB_API int stb_iterate_banded(stb_ctx* ctx, float* img, float* exp_avg, float* exp_avg_sq, float* ema, int64_t step,
                               float lr, float beta1, float beta2, float adam_eps, float ema_decay,
                               float* loss_out_host8, void* stream);
/* 1: iterations replay as CUDA graphs, 2: enabled but nothing captured yet, 0: eager launches (note_out says why). */
STB_API int stb_graph_status(stb_ctx* ctx, char* note_out, size_t note_bytes);
/* Kernel launches issued through graph replays so far, counted from the captured graphs: number of replays, kernels
 * they launched, kernel nodes per graph slot {stb_iterate, stb_iterate_fwd, stb_iterate_bwd, stb_iterate_banded}. */
STB_API int stb_launch_count(stb_ctx* ctx, int64_t* graph_replays, int64_t* kernels_replayed, int* kernels_per_graph4);

/* ------------------------------------------------------------------ measurement (bench.py roofline leg)
 * CUDA-event timing per kernel class on the launching stream; classes in order: conv0_fwd_tv, conv_fwd, pool_fwd,
 * gram, sse, w2, conv_bwd, pool_bwd, conv0_bwd_adam, finalize (STB_PROF_CLASSES entries). */
#define STB_PROF_CLASSES 10
STB_API int stb_profile_enable(stb_ctx* ctx, int enable);
STB_API int stb_profile_read(stb_ctx* ctx, float* ms_out, int* count_out, int n_classes);

/* ------------------------------------------------------------------ diagnostics
 * stb_debug_w2_trace (with STB_W2_TRACE=1 in the environment): per-round %globaltimer stamps of CTA 0 of the W2 chain
 * kernel, 8 words per round for up to 128 rounds (tools/w2_trace.py prints the timeline). */
STB_API int stb_debug_w2_trace(stb_ctx* ctx, unsigned long long* host_out, size_t words, int* rounds_out);
/*
 * Copy of an internal activation (post-ReLU output of conv `conv_index`, bf16 NHWC) of the last forward. */
STB_API int stb_debug_activation(stb_ctx* ctx, int H, int W, int conv_index, void* out_bf16, size_t out_bytes,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STB200_H_ */
