/*
 * clasfv_b200.h - C ABI of libclasfv_b200.so: CLAS-FV full-video inference hot path on B200 (sm_100a).
 *
 * The reference (yc015/fully-automated-multi-heartbeat-echocardiography-video-segmentation-and-
 * motion-tracking) is pure Python and has no FFI layer; its boundary for this path is a set of
 * Python call signatures.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference checkout).  The Python drop-ins in
 * fully-automated-...-tracking_b200/src/ bind these with ctypes.
 *
 * Conventions
 *   - extern "C", plain ints / pointers only.  Every function returns 0 on success and a non-zero
 *     CLASFV_E* code on failure; clasfv_last_error() gives the message.  No C++ exception crosses.
 *   - The caller owns every I/O buffer.  Pointers named *_dev are CUDA device pointers on the
 *     handle's device (e.g. torch tensor.data_ptr()); pointers named *_host are host memory.
 *   - The library owns the packed weights and its activation workspace inside the handle.
 *   - All device work is enqueued on the caller's stream (cudaStream_t passed as void*; NULL =
 *     the legacy default stream) and is asynchronous; the library never synchronises the device
 *     except inside clasfv_finalize() and when the workspace has to grow.
 *   - One handle per (device, stream); a handle is not thread-safe, different handles are independent.
 *   - Tensors are contiguous in the reference's layouts: video (3,T,H,W); clips (N,3,T,H,W);
 *     seg / prob (N,2,T,H,W); motion (N,4,T,H,W) = [fwd x, fwd y, bwd x, bwd y], tanh units
 *     (x pixels = value * W/2, y pixels = value * H/2).
 */
#ifndef CLASFV_B200_H_
#define CLASFV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLASFV_ABI_VERSION 2

/* element types of activation-sized buffers / arithmetic mode of the network */
#define CLASFV_F32  0   /* fp32 storage, fp32 CUDA-core arithmetic (reference-tolerance mode)        */
#define CLASFV_BF16 1   /* bf16 storage, tcgen05 tensor-core arithmetic with fp32 accumulation       */
#define CLASFV_F16  2   /* fp16 storage (11 significant bits instead of 8; conversions saturate), the same
                           tcgen05 kernels at the same rate, fp32 accumulation                          */

/* error codes */
#define CLASFV_OK          0
#define CLASFV_EINVAL      1   /* bad argument (shape, dtype, null pointer, unsorted clip starts ...)  */
#define CLASFV_ESTATE      2   /* call order: tensors missing at finalize, forward before finalize     */
#define CLASFV_ECUDA       3   /* a CUDA runtime / driver call failed                                  */
#define CLASFV_ENOMEM      4
#define CLASFV_EUNSUPPORTED 5  /* device is not sm_100 / no CUDA device                                */

/* forward() output selection */
#define CLASFV_OUT_LOGITS 0    /* seg = raw 2-class logits (what the reference forward returns)        */
#define CLASFV_OUT_PROB   1    /* seg = softmax over the class axis, fused into the head kernel        */
#define CLASFV_OUT_LVPROB 2    /* seg = (N,1,T,H,W): the LV class probability only (what warp-and-fuse reads) */

typedef struct clasfv_handle clasfv_handle;

int clasfv_abi_version(void);

/* Message of the last failing call on this thread ("" if none). Never NULL. */
const char* clasfv_last_error(void);

/* ---- network lifetime -------------------------------------------------------------------------
 * Replaces R2plus1D_18_MotionNet.__init__ + load_state_dict (src/model/R2plus1D_18_MotionNet.py:11-24,
 * motion_segment.py:69-76): create a handle, hand over each state_dict tensor under its reference
 * key (without any "module." prefix; fp32, host memory, reference shape), then finalize, which folds
 * every BatchNorm (inference statistics, eps 1e-5) into the preceding convolution, pads channel
 * counts to multiples of 16, packs K-major per filter tap and uploads.  Unused keys (fc.*,
 * num_batches_tracked) are accepted and ignored.  finalize may be called again after new
 * set_tensor calls (e.g. load_state_dict) or to switch precision. */
int  clasfv_create(int device, clasfv_handle** out);
void clasfv_destroy(clasfv_handle* h);
int  clasfv_set_tensor(clasfv_handle* h, const char* key, const float* data_host, const int64_t* shape, int ndim);
int  clasfv_finalize(clasfv_handle* h, int precision /* CLASFV_F32 | CLASFV_BF16 | CLASFV_F16 */);

/* ---- network forward --------------------------------------------------------------------------
 * Replaces R2plus1D_18_MotionNet.forward (src/model/R2plus1D_18_MotionNet.py:26-71) and, with
 * CLASFV_OUT_PROB, the F.softmax(seg, 1) that follows it at src/fuse_utils.py:60.
 *
 * Input clip n, channel c, frame t lives at  x_dev + clip_offset_host[n] + c*channel_stride + t*H*W
 * (fp32 elements).  clip_offset_host == NULL means a dense (N,3,T,H,W) tensor
 * (offset n*3*T*H*W, channel_stride must then be T*H*W or 0 for "dense").  Passing offsets lets
 * overlapping stride-1 windows of one resident video (3,Tv,H,W) be used in place without copies:
 * offset = start_n*H*W, channel_stride = Tv*H*W.
 * T % 8 == 0, H % 16 == 0, W % 16 == 0 (the decoder's scale-factor upsampling must land on T,H,W).
 * out_dtype selects the element type of seg_dev / motion_dev (CLASFV_F32, CLASFV_BF16 or CLASFV_F16). */
int clasfv_forward(clasfv_handle* h, const float* x_dev, const int64_t* clip_offset_host, int64_t channel_stride,
                   int n, int t, int height, int width, int out_kind, int out_dtype,
                   void* seg_dev, void* motion_dev, void* stream);

/* Options of a handle (no reference counterpart):
 *   "sub_batch"    clips per internal batch of clasfv_forward (default 32): a call with more clips is processed
 *                  batch by batch inside one workspace
 *   "dense_video"  0/1 (default 1).  When the clips of a call are equally spaced windows (1..8 frames apart) of one
 *                  resident video, bf16 tensor-core mode, N >= 4, T >= 16: the stem and layer1 are evaluated once over
 *                  the union of the frames, and per clip only the <= 5 frames next to each clip edge (the ones the
 *                  clip's zero padding reaches through 5 temporal convolutions) are recomputed.  Outputs are
 *                  bit-identical to the per-clip schedule; pass every window of the video in ONE call to benefit.
 *   "umma_pair"    0/1 (default 1).  Convolutions with >= 128 input and output channels and more than one filter tap run on
 *                  CTA pairs (tcgen05.mma.cta_group::2: two SMs of a TPC share one M = 256 MMA, each staging half of the
 *                  filter rows).  Outputs are bit-identical either way (the K order does not depend on the tiling);
 *                  the switch exists for measurements and tests. */
int clasfv_set_option(clasfv_handle* h, const char* name, int value);

/* Stage timing of clasfv_forward with CUDA events recorded on the caller's stream (measurement support,
 * no reference counterpart).  Between begin and end every forward records an event after each stage of each
 * internal batch; end waits for them and returns the summed milliseconds of the four stages
 *   [0] stem 1x7x7   [1] trunk convolutions (stem 3x1x1 + layer1..4, the tcgen05 kernel in bf16 mode)
 *   [2] decoder lateral 1x1x1 projections   [3] fused decoder head
 * and the number of forward calls covered. */
int clasfv_profile_begin(clasfv_handle* h);
int clasfv_profile_end(clasfv_handle* h, float* stage_ms_host /* [4] */, int* calls_host);
/* GFLOP (2 x true, unpadded MACs) the convolutions of each stage performed since clasfv_profile_begin. */
int clasfv_profile_gflop(clasfv_handle* h, double* stage_gflop_host /* [4] */);

/* Largest workspace (bytes) the handle currently holds; informational. */
int64_t clasfv_workspace_bytes(const clasfv_handle* h);
/* Kernels launched through this handle since clasfv_create (every forward / fusion / ingest kernel; measurement support:
 * bench.py reports the difference over its timed region as gpu_launches). */
int64_t clasfv_launch_count(const clasfv_handle* h);

/* ---- video ingest -----------------------------------------------------------------------------
 * Replaces the float pipeline of motion_segment.py:96-106 after the cv2 decode: frames_dev (T,H0,W0,3) uint8 (bgr != 0:
 * bytes are B,G,R as cv2 delivers them, else R,G,B) -> video_dev (3,T,height,width) fp32 =
 * zeroone_normalizer(F.interpolate(float(video), size=(T,height,width), mode="trilinear", align_corners=True))
 * (src/echonet_dataset.py:38-50: per channel x -= min; x /= max after the shift). */
int clasfv_ingest_u8(clasfv_handle* h, const uint8_t* frames_dev, int t, int height0, int width0, int bgr,
                     float* video_dev, int height, int width, void* stream);

/* ---- warp primitive ---------------------------------------------------------------------------
 * Replaces generate_2dmotion_field (src/transform_utils.py:14-34) + its call site
 * F.grid_sample(src, grid, align_corners=False, mode="bilinear", padding_mode="border")
 * (src/clasfv_losses.py:86-87,112-113; src/visualization_utils.py:123-129).
 * src (N,C,H,W), flow (N,2,H,W) [x, y] in normalised units, out (N,C,H,W); all fp32 device.
 *   out(i,j) = bilinear src( clamp(i*H/(H-1) - 1/2 + flow_y*H/2), clamp(j*W/(W-1) - 1/2 + flow_x*W/2) ) */
int clasfv_warp(const float* src_dev, const float* flow_dev, float* out_dev, int n, int c, int height, int width,
                void* stream);
/* The same with the interpolation mode of the call site selectable: the reference's label / image propagation
 * (apply_sequence_deformation, src/visualization_utils.py:106-128) calls grid_sample with mode="nearest" for labels
 * (clamp to the border, then round half to even) and "bilinear" for images. */
#define CLASFV_WARP_BILINEAR 0
#define CLASFV_WARP_NEAREST  1
int clasfv_warp_mode(const float* src_dev, const float* flow_dev, float* out_dev, int n, int c, int height, int width,
                     int mode, void* stream);
/* The sampling grid itself, (N,H,W,2) fp32 = what generate_2dmotion_field returns. */
int clasfv_motion_field(const float* flow_dev, float* grid_dev, int n, int height, int width, void* stream);

/* ---- warp + fuse (north-star operator F2; specified by oracle/fuse_ref.py:warp_fuse) -----------
 * prob_dev (n,prob_planes,L,H,W), motion_dev (n,4,L,H,W) of element type dtype; prob_planes = 2 (background, LV:
 * the softmax of the segmentation logits) or 1 (the LV probability alone, CLASFV_OUT_LVPROB); only the LV plane is
 * read.  Clip c covers global frames clip_start_host[c] + t (ascending starts).  Every clip frame votes its LV
 * probability on its own frame, and - warped along its forward / backward flow - on the next / previous frame; hops
 * leaving the clip are dropped unless edge_hops != 0; votes outside [0, t_out) are dropped.
 * Outputs (any may be NULL except acc_dev):
 *   acc_dev  (t_out,2,H,W) fp32  plane 1 = sum of the LV votes, plane 0 = votes - plane 1 (the background sum: class
 *                                 probabilities and bilinear weights sum to one). accumulate != 0 adds to the existing
 *                                 content (fusing a video clip-batch by clip-batch), else overwrites.
 *   cnt_dev  (t_out) int32       votes per frame (same accumulate rule)
 *   mask_dev (t_out,H,W) uint8   argmax over the two sums after this call (ties -> 0)
 *   area_dev (t_out) int32       LV pixel count of mask per frame (the EF size trace, fuse_utils.py:106) */
int clasfv_warp_fuse(clasfv_handle* h, const void* prob_dev, int prob_planes, const void* motion_dev, int dtype,
                     const int32_t* clip_start_host, int n_clips, int clip_len, int t_out, int height, int width,
                     int edge_hops, int accumulate, float* acc_dev, int32_t* cnt_dev, uint8_t* mask_dev,
                     int32_t* area_dev, void* stream);

/* mask = argmax over the two class sums of acc_dev (t,2,H,W) (ties -> 0) and its per-frame LV area: the last step
 * of clasfv_warp_fuse on its own, for sums that were completed by adding other ranks' partial sums
 * (one long video split by clip range across GPUs). */
int clasfv_finalize_mask(const float* acc_dev, int t, int height, int width, uint8_t* mask_dev, int32_t* area_dev,
                         void* stream);

/* ---- reference-exact fusion (F1) ---------------------------------------------------------------
 * Replaces the device-independent body of segment_a_video_with_fusion (src/fuse_utils.py:36-102)
 * and divide_to_consecutive_clips (src/fuse_utils.py:16-33).
 *
 * Shift k (k < n_shifts) starts at frame shift_start_host[k], is shift_len_host[k] frames long and
 * is cut into shift_nclips_host[k] consecutive clips after (if 32*nclips != len) a linear temporal
 * resample with align_corners=False; its clips are numbered from shift_clip_base_host[k]. */

/* video_dev (3,T,H,W) fp32 -> clips_dev (total_clips,3,clip_len,H,W) fp32 (fuse_utils.py:19-33). */
int clasfv_build_shift_clips(clasfv_handle* h, const float* video_dev, int t, int height, int width, int clip_len,
                             int n_shifts, const int32_t* shift_start_host, const int32_t* shift_len_host,
                             const int32_t* shift_nclips_host, const int32_t* shift_clip_base_host,
                             float* clips_dev, void* stream);

/* prob_dev (total_clips,2,clip_len,H,W) of element type dtype -> mask_dev (t,H,W) uint8:
 * per shift resample the softmax back to shift_len frames (align_corners=False), argmax (tie -> 0),
 * then frame 0 = shift 0 and frame i = majority over shifts k < min(i, n_shifts) with
 * i - k*step >= 0 (tie -> 0) (fuse_utils.py:70-98).  area_dev as in clasfv_warp_fuse. */
int clasfv_fuse_shift_votes(clasfv_handle* h, const void* prob_dev, int dtype, int t, int height, int width,
                            int clip_len, int step, int n_shifts, const int32_t* shift_len_host,
                            const int32_t* shift_nclips_host, const int32_t* shift_clip_base_host,
                            uint8_t* mask_dev, int32_t* area_dev, void* stream);

/* Linear temporal resample, align_corners=False, of a (C,L_in,HW) fp32 array to (C,L_out,HW)
 * (the F.interpolate calls at src/fuse_utils.py:21-23 and :74-76 with unchanged H,W). */
int clasfv_temporal_resample(const float* in_dev, float* out_dev, int channels, int l_in, int l_out, int64_t hw,
                             void* stream);

/* ---- single layer (test / debugging surface) ---------------------------------------------------
 * One Conv3d (+ optional folded BN, residual, ReLU) on channels-last activations, through the same
 * kernels the network uses.  x_dev (N,T,H,W,Cin) and out_dev (N,To,Ho,Wo,Cout) of element type
 * dtype (out_f32 != 0 forces fp32 output); w_host (Cout,Cin,kt,kh,kw) fp32 reference layout;
 * scale_host / shift_host (Cout) fp32 or NULL (per-channel affine applied after the convolution);
 * residual_dev (same shape/type as out) or NULL.  engine: 0 = CUDA-core kernel, 1 = tcgen05 kernel
 * (dtype must be CLASFV_BF16 or CLASFV_F16, Cin % 16 == 0, Cout % 16 == 0). */
int clasfv_conv3d(clasfv_handle* h, const void* x_dev, int dtype, int n, int t, int height, int width, int cin,
                  const float* w_host, const float* scale_host, const float* shift_host, int cout,
                  int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                  const void* residual_dev, int relu, int engine, int out_f32, void* out_dev, void* stream);

/* The fused tensor-core decoder head on its own (same test surface idea): the four laterally projected 64-channel
 * maps g_l, channels-last fp16 (N,T,H/2^(l+1),W/2^(l+1),64), l = 0..3, all at the output's frame rate, through the
 * head of the finalized network (comb_1 bias + ReLU, comb_2 + ReLU, segmentation / motion heads of
 * src/model/R2plus1D_18_MotionNet.py:55-69, bilinear align_corners=True interpolation of :41-49 inside the first GEMM).
 * The handle must be finalized in a tensor-core precision (CLASFV_BF16 or CLASFV_F16). */
int clasfv_decoder_head(clasfv_handle* h, const void* g0_dev, const void* g1_dev, const void* g2_dev, const void* g3_dev,
                        int n, int t, int height, int width, int out_kind, int out_dtype,
                        void* seg_dev, void* motion_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLASFV_B200_H_ */
