/* cdan_b200.h — C ABI of the B200-native CDAN forward (libcdan_b200.so).
 *
 * The reference (danielluca00/Multi-Degradation-Image-Enhancement) is pure Python/PyTorch and has no FFI of its
 * own; the seam this library sits behind is the eval-mode `nn.Module.__call__` of `models.cdan.CDAN`
 * (reference models/cdan.py:171-176, called from models/model.py:252,342) and `models.cbam.CBAM`
 * (models/cbam.py:91-95), plus the post-processing table (utils/postprocessing_factory.py:19-41) and the PSNR/SSIM
 * items of the metrics pipeline (utils/metrics_factory.py:74-94).  Each entry point below names the reference
 * interface it replaces.  INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 on success, non-zero on error; the message is available from
 * cdan_last_error() (thread-local).  All pointers are plain device pointers unless the name says `_host`.
 * Calls are stream-ordered on the `stream` argument (a cudaStream_t passed as void*; NULL = default stream) and
 * never synchronise the host, except the `_host` variants.  The caller owns inputs/outputs; the plan owns packed
 * weights and workspace.  A plan may be used by one host thread at a time.  There is no CPU fallback.
 */
#ifndef CDAN_B200_H_
#define CDAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CDAN_API __attribute__((visibility("default")))
#else
#define CDAN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define CDAN_DTYPE_F32 0  /* fp32 storage, fp32 FFMA accumulate (CUDA cores) — matches the reference to ~1e-5 */
#define CDAN_DTYPE_BF16 1 /* bf16 storage + tcgen05 bf16 MMA, fp32 accumulate in TMEM — the throughput path      */

typedef struct cdan_plan cdan_plan;

CDAN_API const char* cdan_last_error(void);
/* Library / build identification, e.g. "cdan_b200 0.1 sm_100a". */
CDAN_API const char* cdan_version(void);

/* ---- plan life cycle (replaces: CDAN().to(device).eval(), reference models/model.py:29,232) */
CDAN_API int cdan_plan_create(int device, int dtype, cdan_plan** plan_out);
CDAN_API int cdan_plan_destroy(cdan_plan* plan);

/* Load the 236-entry state_dict (reference weight-file format, models/base.py:52-55, models/model.py:231).
 * keys[i] is the state_dict key, ptrs[i] a host or device pointer to `numels[i]` fp32 values (entries whose key ends
 * in "num_batches_tracked" are ignored and may be omitted).  The library folds eval-mode BatchNorm into the
 * following/preceding convolution where algebraically possible, flips/transposes ConvTranspose2d kernels, pads
 * channels and packs the result for its kernels; it keeps its own copy. */
CDAN_API int cdan_plan_load_weights(cdan_plan* plan, int n, const char* const* keys, const void* const* ptrs,
                           const int64_t* numels);

/* Options: "host_chunk" = largest sub-batch of the host-buffer copy/compute pipeline (default 0 = auto: 16 x 1080p of pixels; the
 *                        schedule ramps up to it and down again, see cdan_forward_host);
 *          "fd_fused"  = 1 (default) final dense block as one fused kernel on bf16 plans, 0 = layer by layer (also materialises
 *                        the "dec.final_in" stage tap);
 *          "conv_impl" = 0 auto (tcgen05 where supported), 1 force CUDA-core path;
 *          "profile"   = 1 brackets every launch group with CUDA events on the launching stream. */
CDAN_API int cdan_plan_set_option(cdan_plan* plan, const char* name, int value);

/* Bytes of workspace the plan will hold for an [N,3,H,W] forward (H and W must be multiples of 8). */
CDAN_API int cdan_workspace_bytes(cdan_plan* plan, int N, int H, int W, size_t* bytes_out);

/* CDAN.forward in eval mode (reference models/cdan.py:171-176).  x, y: fp32 NCHW contiguous [N,3,H,W] on the plan's
 * device.  H % 8 != 0 or W % 8 != 0 is an error (the reference fails with a size mismatch there too). */
CDAN_API int cdan_forward(cdan_plan* plan, void* stream, const float* x, float* y, int N, int H, int W);
/* Same through HOST buffers (pinned memory recommended): the batch is processed in sub-batches — a geometric ramp from about one
 * 1080p image of pixels up to "host_chunk" images and down again, so that the first H2D and the last D2H copy are short and every
 * other copy hides behind a neighbouring forward; the H2D copy of the next and the D2H copy of the previous sub-batch overlap the
 * forward of the current one on separate streams.  Results do not depend on the schedule.  Returns after the last D2H copy. */
CDAN_API int cdan_forward_host(cdan_plan* plan, const float* x_host, float* y_host, int N, int H, int W);
/* The reference's real data path end to end (uint8 image -> /255 -> forward -> x255 -> uint8; data/dataset.py:86-92 with
 * `A.Normalize(0,1,255)` + `ToTensorV2`, models/model.py:80-83): x_host, y_host are interleaved uint8 [N,H,W,3] HOST buffers.
 * Per sub-batch the uint8 pixels cross PCIe (a quarter of the fp32 bytes each way), are normalised on the device
 * (x = u8 * float32(1/255)), run through cdan_forward and are quantised on the device exactly as cdan_quantize_u8. */
CDAN_API int cdan_forward_host_u8(cdan_plan* plan, const unsigned char* x_host, unsigned char* y_host, int N, int H, int W);

/* ---- spatial row tiling of very large images (BASELINE config C5: one 4K image over the GPUs of a box; the cross-row operators
 *      are the 3x3 (transposed) convolutions, bilinear x2 and SpatialGate 7x7 of reference models/cdan.py:70-159, models/cbam.py:72-82,
 *      the global pooling of ChannelGate is models/cbam.py:41,44).  One plan per band; band r of `nbands` owns the rows
 *      cdan_band_rows reports (boundaries at multiples of 8 rows) and computes on them plus `halo` rows above and below
 *      (multiple of 8, >= 24).  Halo rows of intermediate tensors are refreshed from the neighbouring bands only where the
 *      schedule needs it (7 exchanges per forward at halo 24); ChannelGate statistics are all-reduced (SUM, MAX).
 *      Transports: NCCL (one process per GPU, ncclSend/ncclRecv/ncclAllReduce on the caller's stream; libnccl.so.2 is loaded
 *      at run time) or in-process (all bands in one process, one host thread per band calling cdan_forward_band
 *      concurrently; used to test the schedule on a single GPU). */
typedef struct cdan_band_group cdan_band_group;
CDAN_API int cdan_band_group_create(int nbands, cdan_band_group** group_out);
CDAN_API int cdan_band_group_destroy(cdan_band_group* group);
CDAN_API int cdan_plan_band_attach_local(cdan_plan* plan, cdan_band_group* group, int rank);
/* 128-byte NCCL unique id: create on rank 0, distribute with any host mechanism, pass to every rank's attach. */
CDAN_API int cdan_band_nccl_unique_id(void* id_out, size_t len);
CDAN_API int cdan_plan_band_attach_nccl(cdan_plan* plan, int rank, int nranks, const void* id, size_t len);
CDAN_API int cdan_plan_band_detach(cdan_plan* plan);
/* rows_out = {owned begin, owned end, extended begin, extended end} of band `rank` (rows of the full image). */
CDAN_API int cdan_band_rows(int H, int nbands, int rank, int halo, int rows_out[4]);
/* Forward of this plan's band.  x_ext / y_ext: fp32 NCHW [N,3,Hext,W] holding rows [extended begin, extended end) of the
 * [N,3,H,W] image; on return the OWNED rows of y_ext equal the untiled forward's (halo rows of y_ext are scratch).
 * All bands must make the call (collective). */
CDAN_API int cdan_forward_band(cdan_plan* plan, void* stream, const float* x_ext, float* y_ext, int N, int H, int W, int halo);
/* Counters of the most recent cdan_forward_band: {halo exchanges, halo bytes received, all-reduces}. */
CDAN_API int cdan_band_stats(cdan_plan* plan, long long out[3]);

/* Read an intermediate tensor of the most recent cdan_forward as fp32 NCHW (per-stage parity tests).
 * Names: "enc.out1","enc.out2","enc.out3" (max-pooled ConvBlock outputs = skip connections), "enc.dense1".."enc.dense3",
 * "enc.conv4","bottleneck","dec.bn1".."dec.bn4" (relu(bn(convT))), "dec.gated1".."dec.gated3" (cbam_i(.)*dense),
 * "dec.final_in" (3 channels).  shape_out receives {N,C,H,W}; pass dst=NULL to query the shape only. */
CDAN_API int cdan_stage_read(cdan_plan* plan, void* stream, const char* name, float* dst, int64_t shape_out[4]);

/* Number of kernels launched by the most recent cdan_forward on this plan. */
CDAN_API int cdan_last_launch_count(cdan_plan* plan);

/* With option "profile"=1: drain the recorded spans into text lines "label total_ms count\n" (labels:
 * "conv|<state_dict prefix>|<umma_tma|umma_pro|simt>", "cbam|C<channels>", "glue|up_add") and reset. Synchronises. */
CDAN_API int cdan_profile_read(cdan_plan* plan, char* buf, size_t buflen);

/* ---- single operators (unit tests; all tensors fp32 NCHW on the device, converted internally to the dtype's
 *      NHWC layout).  impl: 0 auto, 1 CUDA-core, 2 tcgen05. */
/* y = [maxpool2x2]( [relu]( conv_ks(  [relu(pre_scale*x+pre_shift)]  , w) + bias ) ), w is [Cout,Cin,ks,ks]
 * (reference ConvBlock models/cdan.py:8-19, dense layers :41-53). */
CDAN_API int cdan_op_conv2d(int dtype, int impl, void* stream, const float* x, int N, int Cin, int H, int W, const float* w,
                   const float* bias, int Cout, int ks, const float* pre_scale, const float* pre_shift, int relu,
                   int pool, float* y);
/* y = CBAM(x) [* mul]   (reference models/cbam.py:84-95; decoder use models/cdan.py:131-133).
 * w1 [C/16,C], b1 [C/16], w2 [C,C/16], b2 [C], w7 [1,2,7,7]; bn = {weight,bias,running_mean,running_var} of the
 * SpatialGate BatchNorm2d(1) (4 floats, host pointer). */
CDAN_API int cdan_op_cbam(int dtype, void* stream, const float* x, int N, int C, int H, int W, const float* w1, const float* b1,
                 const float* w2, const float* b2, const float* w7, const float* bn_host4, const float* mul, float* y);
/* y = (up ? bilinear_x2(a) : a) + skip   (reference models/cdan.py:130,137-138). a: [N,C,H,W]; y: [N,C,H*(1+up),W*(1+up)]. */
CDAN_API int cdan_op_upsample_add(int dtype, void* stream, const float* a, const float* skip, int N, int C, int H, int W, int up,
                         float* y);

/* ---- tail (reference utils/post_processing.py:5-77 via utils/postprocessing_factory.py:10-15).
 * op: 0 enhance_contrast(contrast_factor), 1 enhance_color(saturation_factor), 2 sharpen(strength),
 * 3 soft_denoise(sigma).  x,y: fp32 [N,3,H,W]; in-place (x == y) allowed for ops 0 and 1. */
CDAN_API int cdan_postprocess(void* stream, int op, float arg, const float* x, float* y, int N, int H, int W);
/* Output quantisation of Model._save_batch_outputs (reference models/model.py:80-83, `(img*255).clip(0,255).astype(uint8)`
 * after `permute(1,2,0)`): x fp32 [N,3,H,W] (device) -> y uint8 [N,H,W,3] (device), truncation toward zero.  H*W % 4 == 0.
 * Stream-ordered; a quarter of the bytes of x then cross PCIe. */
CDAN_API int cdan_quantize_u8(void* stream, const float* x, unsigned char* y, int N, int H, int W);
/* Input side (reference data/dataset.py:86-92 with utils/transforms_factory.py:50-86: `A.Resize` = cv2.resize INTER_LINEAR
 * on the uint8 image, `A.Normalize(mean 0, std 1, max_pixel_value 255)`, `ToTensorV2`): src uint8 [N,Hs,Ws,3] (device) ->
 * dst fp32 [N,3,Hd,Wd] (device).  Bit-exact with cv2.resize for 8-bit images (OpenCV's 11-bit fixed-point bilinear);
 * the normalisation multiplies by float32(1/255).  Synchronises the stream. */
CDAN_API int cdan_resize_normalize_u8(void* stream, const unsigned char* src, int N, int Hs, int Ws, float* dst, int Hd,
                                      int Wd);
/* PSNR and SSIM with torchmetrics' default settings (reference utils/metrics_factory.py:74-94); result_host[0] =
 * PSNR (dB), result_host[1] = SSIM.  Synchronises the stream. */
CDAN_API int cdan_psnr_ssim(void* stream, const float* pred, const float* target, int N, int C, int H, int W,
                   float* result_host2);

#ifdef __cplusplus
}
#endif
#endif /* CDAN_B200_H_ */
