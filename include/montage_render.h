/*
 * montage_render.h -- C ABI of libmontage_render.so (B200 / sm_100a).
 *
 * The drop-in boundary for ONE hot path of uchidalab/docker-montage-gan: the global GAN's
 * analytic renderer = warp every RGBA layer by its 2x3 placement, then alpha-over composite
 * back to front; forward and backward.  Plain pointers and sizes only: no torch types, no
 * C++ in the signatures, so the host side can be bound from ctypes / cffi / pybind / cgo.
 *
 * Reference interfaces replaced (paths relative to /root/reference/montage_gan):
 *   - the warp inside STNv2c.forward / STNv2b.forward      fukuwarai/networks.py:247-258, 217-226
 *     (torch.nn.functional.affine_grid + grid_sample, bilinear, zeros, align_corners=False)
 *   - alpha_composite_pytorch (default, non-premultiplied)   custom_utils/image_utils.py:112-163
 *   - normalize_zero1 / normalize_minus11                     custom_utils/image_utils.py:184-195
 *   - the chain at MontageGANLoss.run_global_D                custom/loss_aio.py:245-257 (:251)
 *   - the call signature of Renderer*.forward                 diff_rendering/networks.py:36-44
 *   - convert_translate_to_2x3                                custom_utils/image_utils.py:316-335
 *   - make_batch_for_pos_estimator (pad to canvas + stack)    custom_utils/image_utils.py:216-243
 * The reference binds its own native ops with pybind11 torch extensions
 * (torch_utils/ops/bias_act.cpp:94-97, upfirdn2d.cpp:98-101); INTEGRATION.md shows the ctypes
 * stub a maintainer adds instead.
 *
 * Conventions
 *   - Tensors: x [B,L,4,H,W] (RGBA, alpha = channel 3), theta [B,L,2,3] float32 mapping OUTPUT
 *     normalised coords to INPUT normalised coords (align_corners=False), out [B,4,H,W].
 *   - Layer 0 is the back (image_utils.py:142-146).
 *   - dtype: element type of x / out / grad_out / grad_x.  All arithmetic is fp32 in registers.
 *   - range_mode MGR_RANGE_M11: x and out in [-1,1]; the kernel applies the STNv2c "+1 ...
 *     -1" workaround and normalize_zero1 / normalize_minus11 internally.  MGR_RANGE_01: x and
 *     out in [0,1] (STNv2b / random_position / bare alpha_composite_pytorch).
 *   - Where the composited alpha is exactly 0 every gradient is defined as 0 (the reference
 *     produces NaN there: 0/0 in a_over_b, image_utils.py:128-133).
 *   - All device pointers belong to the CURRENT CUDA device; every call is asynchronous on
 *     `stream` (a cudaStream_t passed as void*), never synchronises, never allocates device
 *     memory, and keeps no pointer after returning: graph-capturable.
 *   - Return value: 0 on success, otherwise a non-zero code (MGR_ERR_* or a cudaError_t
 *     offset by MGR_ERR_CUDA_BASE); mgr_last_error() returns a thread-local message.
 *   - x_strides: element strides of x for [B,L,4,H,W]; NULL means contiguous.  The innermost
 *     (W) stride must be 1 for the tiled kernels' vector loads; other strides are free
 *     (e.g. a [B,L*4,H,W] view or a batch slice).
 */
#ifndef MONTAGE_RENDER_H_
#define MONTAGE_RENDER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGR_ABI_VERSION 4

enum { MGR_F32 = 0, MGR_BF16 = 1, MGR_F16 = 2 };
enum { MGR_RANGE_M11 = 0, MGR_RANGE_01 = 1 };
enum { MGR_NEED_GRAD_X = 1, MGR_NEED_GRAD_THETA = 2 };
enum {
  MGR_OK = 0,
  MGR_ERR_INVALID_ARGUMENT = 1,
  MGR_ERR_UNSUPPORTED = 2,
  MGR_ERR_WORKSPACE_TOO_SMALL = 3,
  MGR_ERR_CUDA_BASE = 1000
};

/* ABI / build identification. */
int mgr_abi_version(void);
const char* mgr_build_info(void);
/* Thread-local, NUL-terminated description of the last non-zero return on this thread. */
const char* mgr_last_error(void);
/* Number of CUDA kernels this library has launched in this process (all threads, all calls);
 * bench.py differences it around the timed region to report "gpu_launches". */
long long mgr_kernel_launch_count(void);
/* Testing / A-B timing only: 0 = automatic kernel selection (default), 1 = force the general
 * direct-gather kernels even where the tiled shared-memory kernels apply, 2 = tiled kernels but
 * without the pure-translation stencil kernels, 3 = round 1's two-barrier tiled kernels instead of the
 * warp-specialised ones, 4 = the staged stencil kernels instead of the ones on TMA box copies (renderer and materialised warp).  Process-wide. */
int mgr_set_debug_path(int path);

/*
 * Bytes of the optional "saved alpha" buffer [B,L,H,W] the forward can fill for the backward:
 * the warped alpha sample of every (layer, pixel), fp32 for MGR_F32 tensors, fp16 otherwise.
 * (What autograd would keep alive in the reference is 13x larger: the grid, the warped layers and
 * every a_over_b intermediate.)  The buffer ends with one int per sample -- "all placements of the sample
 * are pure translations" -- which the forward writes first, so that every CTA of the two kernels that
 * share a batch learns with one load whether a sample is its own.
 */
size_t mgr_saved_alpha_bytes(int B, int L, int H, int W, int dtype);

/*
 * Fused warp + composite, forward.
 *   replaces: fukuwarai/networks.py:250-257 + custom/loss_aio.py:251 (theta != NULL)
 *             custom/loss_aio.py:251 alone on the real branch :313-320 (theta == NULL)
 *   out          written completely
 *   saved_alpha  NULL (inference), or mgr_saved_alpha_bytes(...) bytes written completely when
 *                theta != NULL; pass it to mgr_render_backward to enable the atomics-free backward
 */
int mgr_render_forward(const void* x, const int64_t* x_strides, const float* theta, void* out,
                       void* saved_alpha, int B, int L, int H, int W, int dtype, int range_mode,
                       void* stream);

/*
 * Bytes of scratch mgr_render_backward needs for this problem (0 is possible).
 */
size_t mgr_render_backward_workspace_bytes(int B, int L, int H, int W, int dtype, int has_theta,
                                           int flags);
/*
 * The same for ONE given tensor: knows from the pointer and the strides which kernels the call will take (the tiled
 * two-pass kernels need about 8 bytes per layer-pixel, the scatter fallback for 16-bit tensors 16), so it is what a
 * caller that allocates per call should ask; the function above is the upper bound over all layouts.
 */
size_t mgr_render_backward_workspace_bytes_for(const void* x, const int64_t* x_strides, int has_theta,
                                               int has_saved_alpha, int B, int L, int H, int W, int dtype,
                                               int flags);

/*
 * Fused warp + composite, backward (the autograd of the chain above).
 *   out         saved forward result [B,4,H,W] (same dtype), read-only
 *   grad_out    [B,4,H,W] contiguous
 *   saved_alpha what the forward wrote, or NULL (then the general scatter kernels run)
 *   grad_x      [B,L,4,H,W] contiguous, written completely (no pre-zeroing needed); may be NULL
 *               when flags lacks MGR_NEED_GRAD_X
 *   grad_theta  [B,L,2,3] float32, written completely; may be NULL when theta is NULL or flags
 *               lacks MGR_NEED_GRAD_THETA
 *   workspace   device scratch of at least mgr_render_backward_workspace_bytes(...) bytes,
 *               256-byte aligned; contents undefined on entry and exit
 */
int mgr_render_backward(const void* x, const int64_t* x_strides, const float* theta,
                        const void* out, const void* grad_out, const void* saved_alpha,
                        void* grad_x, float* grad_theta, void* workspace, size_t workspace_bytes,
                        int B, int L, int H, int W, int dtype, int range_mode, int flags,
                        void* stream);

/*
 * Materialised warp of every layer: out[b,l] = warp(x[b,l], theta[b,l]), [B,L,4,H,W] contiguous.
 *   replaces: the warp lines of STNv2c.forward (fukuwarai/networks.py:250-257, MGR_RANGE_M11: the
 *   "+1, grid_sample, -1" form) and of STNv2b.forward / random_position (networks.py:219-225,
 *   custom_utils/image_utils.py:289-294, MGR_RANGE_01).  Kept for the callers that consume warped
 *   layers themselves (EMA snapshots, metrics); the training path uses mgr_render_forward instead.
 */
int mgr_warp_forward(const void* x, const int64_t* x_strides, const float* theta, void* out, int B,
                     int L, int H, int W, int dtype, int range_mode, void* stream);
size_t mgr_warp_backward_workspace_bytes(int B, int L, int H, int W, int dtype, int flags);
/* grad_out [B,L,4,H,W] contiguous; grad_x / grad_theta written completely (see mgr_render_backward). */
int mgr_warp_backward(const void* x, const int64_t* x_strides, const float* theta,
                      const void* grad_out, void* grad_x, float* grad_theta, void* workspace,
                      size_t workspace_bytes, int B, int L, int H, int W, int dtype, int range_mode,
                      int flags, void* stream);

/*
 * translation [n,2] (dx, dy) -> theta [n,2,3] = [[1,0,dx],[0,1,dy]], one launch.
 *   replaces: convert_translate_to_2x3 (custom_utils/image_utils.py:316-335), a B*L Python loop that
 *   builds one torch.tensor(device=...) per layer.  Backward is a slice (grad_theta[..., 2]).
 */
int mgr_translation_to_theta(const float* translation, float* theta, long long n, void* stream);

/*
 * Centre-pad one local generator's output src [B,4,h,w] into layer l of dst [B,L,4,H,W]
 * (constant pad_value outside), one launch per layer.
 *   replaces: pad_256 + make_batch_for_pos_estimator (custom_utils/image_utils.py:216-243), a
 *   per-sample F.pad loop followed by stack + transpose + contiguous.  src_strides: element strides
 *   [b,c,h,w] or NULL for contiguous.  Backward is a crop (a view).
 */
int mgr_pad_stack_layer(const void* src, const int64_t* src_strides, void* dst, int B, int L, int l,
                        int h, int w, int H, int W, float pad_value, int dtype, void* stream);

/*
 * Forward-mode derivative of the composite (theta == NULL path): out_tangent [B,4,H,W] = d out along
 * `tangent` [B,L,4,H,W] (contiguous) at x.  This is the double backward of the composite w.r.t. grad_out,
 * which the global discriminator's R1 penalty on the real layers needs (custom/loss_aio.py:327-338; the
 * reference gets it from stock autograd, its own plugins provide it by hand: torch_utils/ops/bias_act.py:198-226).
 */
int mgr_composite_jvp(const void* x, const int64_t* x_strides, const void* tangent, void* out_tangent,
                      int B, int L, int H, int W, int dtype, int range_mode, void* stream);

/*
 * Ragged stacks: the L layers of a sample are separate tensors of their native sizes -- what
 * the local generators emit (training/dataset_aio.py:30-83: 256x256, 160x224, 96x160, 64x96, ...) -- each centred
 * on the H x W canvas, instead of one [B,L,4,H,W] tensor padded with -1 by make_batch_for_pos_estimator
 * (custom_utils/image_utils.py:216-243).  Texels outside a layer's rectangle read as the padding value
 * (transparent black), exactly as if the padded canvas had been built, but are never stored, loaded or
 * differentiated; tiles whose taps miss the rectangle skip the layer.
 *   layers[l]       layer l of every sample: ptr -> [B,4,h,w] (element strides sb, sc, sh; column stride 1) placed
 *                   with its top-left texel at (left, top) of the canvas.  HOST array of L entries, read during the call.
 *   grads[l]        where grad of layers[l] goes, same rectangle, written once (no zero-fill needed)
 * Requirements (else MGR_ERR_UNSUPPORTED): theta given, 2 <= L <= 32, W, every w and every left multiples of 4,
 * strides multiples of 4 elements (grads: 2), 4-element-aligned base pointers, rectangles inside the canvas.
 * saved_alpha / workspace: as for mgr_render_forward / mgr_render_backward with the same B, L, H, W.
 */
typedef struct MgrLayer {
  void* ptr;
  int64_t sb, sc, sh;
  int h, w, top, left;
} MgrLayer;
int mgr_render_forward_ragged(const MgrLayer* layers, const float* theta, void* out, void* saved_alpha,
                              int B, int L, int H, int W, int dtype, int range_mode, void* stream);
int mgr_render_backward_ragged(const MgrLayer* layers, const float* theta, const void* out, const void* grad_out,
                               const void* saved_alpha, const MgrLayer* grads, float* grad_theta, void* workspace,
                               size_t workspace_bytes, int B, int L, int H, int W, int dtype, int range_mode,
                               int flags, void* stream);

/*
 * Non-differentiable 8-bit composite, bit-exact with the reference's Pillow path
 * (custom_utils/image_utils.py:74-96 alpha_composite: ToPILImage -> Image.alpha_composite per layer -> ToTensor;
 * callers custom/loss_aio.py:351,362, custom/training_loop_aio.py:531,765,775, metrics/metric_utils.py:233,304).
 *   x        [B,L,4,H,W] as in mgr_render_forward (any dtype; MGR_RANGE_M11 applies normalize_zero1 first,
 *            image_utils.py:184-187), layer 0 = back
 *   out_f32  [B,4,H,W] fp32 = byte / 255 (what the reference returns), or NULL
 *   out_u8   [B,4,H,W] uint8 canvas bytes (for PNG snapshots without a second pass), or NULL -- not both NULL
 * Bytes are trunc(v * 255) of the fp32 value; values outside [0,1] saturate (the reference's cast wraps there).
 */
int mgr_composite_u8(const void* x, const int64_t* x_strides, float* out_f32, unsigned char* out_u8,
                     int B, int L, int H, int W, int dtype, int range_mode, void* stream);

/*
 * AugmentPipe's geometric execution block (training/augment.py:306-342): reflect pad by the margins,
 * x2 upsample with the sym6 low-pass, affine bilinear resampling onto a 2(H+6) x 2(W+6) grid, low-pass + x2 decimation
 * + crop back to H x W.  fp32, contiguous [B,C,H,W].  The caller supplies what the reference computes on the host:
 *   mx0, my0, mx1, my1  the reflect padding (augment.py:311-322), each in [0, size - 1]
 *   theta [B,2,3]       device pointer: the matrices the reference hands to affine_grid (augment.py:326-338)
 * (montage_gan_b200.augment.geometric_warp derives both from G_inv).  The backward is the adjoint chain (the block is
 * linear in the images): grad_images is written completely.  workspace: mgr_augment_geom_workspace_bytes(...) bytes.
 */
size_t mgr_augment_geom_workspace_bytes(int B, int C, int H, int W, int mx0, int my0, int mx1, int my1);
int mgr_augment_geom_forward(const float* images, const float* theta, float* out, void* workspace, size_t workspace_bytes,
                             int B, int C, int H, int W, int mx0, int my0, int mx1, int my1, void* stream);
int mgr_augment_geom_backward(const float* grad_out, const float* theta, float* grad_images, void* workspace,
                              size_t workspace_bytes, int B, int C, int H, int W, int mx0, int my0, int mx1, int my1,
                              void* stream);

/*
 * End to end with HOST buffers: out, grad_x, grad_theta = fwd+bwd(x, theta, grad_out), everything in
 * (preferably pinned) host memory, laid out exactly like the device tensors.  The batch is cut
 * into chunks of chunk_B samples that flow through two device slots on three streams (H2D copy,
 * kernels on `stream`, D2H copy) so PCIe traffic in both directions overlaps the kernels.  The
 * call is asynchronous: `stream` completes when the last result byte has reached the host.
 * d_workspace: device scratch of mgr_render_host_workspace_bytes(chunk_B, ...) bytes.
 *   This is the call bench.py times for its "e2e" figure; in the reference the same data path is
 *   .to(device) + the chain of custom/loss_aio.py:238-257 + .cpu().
 */
size_t mgr_render_host_workspace_bytes(int chunk_B, int L, int H, int W, int dtype);
int mgr_render_fwd_bwd_host(const void* h_x, const float* h_theta, const void* h_grad_out, void* h_out,
                            void* h_grad_x, float* h_grad_theta, void* d_workspace,
                            size_t d_workspace_bytes, int chunk_B, int B, int L, int H, int W,
                            int dtype, int range_mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MONTAGE_RENDER_H_ */
