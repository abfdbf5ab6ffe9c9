// internal.h - shared declarations between the translation units of libclasfv_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include <utility>

#include "../../include/clasfv_b200.h"

namespace clasfv {

void set_error(const char* fmt, ...);

#define CLASFV_CUDA(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::clasfv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CLASFV_ECUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define CLASFV_REQUIRE(cond, ...)                                                                 \
  do {                                                                                            \
    if (!(cond)) { ::clasfv::set_error(__VA_ARGS__); return CLASFV_EINVAL; }                      \
  } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Opt a kernel in to the full 227 KB of dynamic shared memory once per device (the attribute is a maximum, so one value
// serves every launch); a driver call per launch is ~170 calls per video step otherwise.
cudaError_t allow_max_dynamic_smem_impl(const void* kernel);     // api.cu: once per (kernel, device)
template <typename Kernel>
inline cudaError_t allow_max_dynamic_smem(Kernel kernel) { return allow_max_dynamic_smem_impl(reinterpret_cast<const void*>(kernel)); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Launch with programmatic stream serialisation: the kernel may begin before its predecessor in the stream has finished
// and MUST execute griddepcontrol.wait (ptx::griddep_wait) before it touches anything a predecessor reads or writes.
// Off unless CLASFV_PDL=1 (it measured no gain, see pdl_enabled in api.cu); without the attribute the griddepcontrol
// instructions in the kernels are no-ops.
bool pdl_enabled();                                                  // api.cu
// `cluster` > 1 launches thread-block clusters of that many CTAs along x (the grid must be a multiple of it).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------------------------------------
constexpr int CLASFV_MAX_FSEL = 64;      // longest virtual clip of a frame-selected convolution

// One convolution on channels-last (N,T,H,W,C) activations.
struct ConvShape {
  int n, ti, hi, wi, cin;      // cin, cout: stored (padded) channel counts
  int to, ho, wo, cout;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
};

struct ConvArgs {
  ConvShape s;
  const void* in;         // (N,Ti,Hi,Wi,Cin)  element type = act_dtype
  const void* in2;        // optional second source of a two-source 1x1x1 convolution: out = in*W[0] + in2*W[1]
  int64_t in_batch_stride;  // elements between consecutive clips of `in` (0 = dense Ti*Hi*Wi*Cin); smaller than a
                            // clip for overlapping temporal windows of one resident video
  const void* weight;     // [tap][Cout][Cin]  element type = act_dtype (fp32 or bf16), BN scale folded
  const float* bias;      // [Cout] fp32 or nullptr
  const void* residual;   // (N,To,Ho,Wo,Cout) element type = out type, or nullptr
  void* out;              // (N,To,Ho,Wo,Cout)
  int act_dtype;          // CLASFV_F32 | CLASFV_BF16 | CLASFV_F16: type of in / weight
  int out_f32;            // 1: out (and residual) are fp32 regardless of act_dtype
  int out_f16;            // 1: out (and residual) are fp16 regardless of act_dtype (tcgen05 path: the decoder's lateral maps)
  int relu;
  double macs_per_pos;    // true (unpadded) MACs per output position, for the profiler's performed-FLOP count
  // ---- optional, tcgen05 path only (zero-initialised by make_conv): ragged clip geometry of the dense-video trunk
  int64_t out_batch_stride;   // elements between consecutive clips of `out` (0 = dense To*Ho*Wo*Cout)
  int no_pair;                // 1: never run this convolution on CTA pairs (cta_group::2); option "umma_pair" of the handle
  // Time-segmented input of a 3x1x1 stride-1 pad-1 convolution (seg.on): the clip the convolution sees is VIRTUAL.
  // Its input frame v (v = -1 .. to, the convolution's own zero padding included) is frame v + a_toff of source A
  // (`in`, a_t frames per clip, in_batch_stride) when v < split, else frame v + b_toff of source B (b, b_t frames per
  // clip, b_batch_stride); a frame index outside its source's [0, *_t) reads as zeros.  The output has `to` frames.
  struct TimeSeg {
    int on, to, split, a_t, a_toff, b_t, b_toff;
    const void* b; int64_t b_batch_stride;
  } seg;
  // Segmented residual: output frame ot adds frame ot + a_toff of `residual` (res.a_batch_stride) when ot < split,
  // else frame ot + b_toff of res.b (res.b_batch_stride).  res.on == 0: residual is dense like `out`.
  struct ResSeg {
    int on, split, a_toff, b_toff;
    int64_t a_batch_stride; const void* b; int64_t b_batch_stride;
  } res;
  // Frame-selected input of a convolution without temporal extent (kt == 1, pt == 0; fsel.on): the clip the convolution
  // sees is VIRTUAL, s.ti frames long.  Its frame f is frame fsel.idx[f] of `in` (fsel.src[f] == 0; fsel.a_t frames per clip,
  // in_batch_stride) or of fsel.b (fsel.src[f] == 1; fsel.b_t frames per clip, fsel.b_batch_stride).  Applies to `in` only
  // (not to `in2`).  The dense-video trunk reads a clip's layer-1 output this way: edge frames from the clip's own
  // buffer, interior frames in place from the shared video-level map (no assembled per-clip copy).
  struct FrameSel {
    int on, a_t, b_t;
    const void* b; int64_t b_batch_stride;
    int8_t src[CLASFV_MAX_FSEL]; int16_t idx[CLASFV_MAX_FSEL];
  } fsel;
};

// CUDA-core implicit GEMM (both storage types).  conv_simt.cu
int launch_conv_simt(const ConvArgs& a, cudaStream_t stream);
// tcgen05 / TMEM / TMA implicit GEMM, bf16 only.  conv_umma.cu
int launch_conv_umma(const ConvArgs& a, int num_sms, cudaStream_t stream);
int umma_selftest_supported();
// cuTensorMapEncodeTiled for a 16-bit tensor, SWIZZLE_128B, zero fill out of bounds (conv_umma.cu)
int encode_tmap_16bit(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool fp16);

// Stem 1x7x7 stride (1,2,2) pad (0,3,3) convolution from the planar fp32 input.  conv_simt.cu
struct StemArgs {
  const float* x;              // planar fp32 input
  const int64_t* clip_offset;  // device [N] element offsets
  int64_t channel_stride;
  int n, t, h, w;              // input dims (output is (N,T,H/2,W/2,48))
  const float* weight;         // [147][48] fp32 (tap-major: (c*7+kh)*7+kw), BN folded, channels 45..47 zero
  const float* bias;           // [48]
  void* out;                   // (N,T,H/2,W/2,out_channels); channels 45.. are written as zero
  int out_channels;
  int out_dtype;
};
int launch_stem(const StemArgs& a, cudaStream_t stream);

// Decoder head: 4-level trilinear (align_corners=True) gather-sum of the laterally projected feature
// maps + bias + ReLU + 64x64 + ReLU + 6x64 heads + softmax / tanh.  decoder.cu
struct HeadArgs {
  const void* g[4];            // (N,Tl,Hl,Wl,64) fp32 (g_dtype F32) or fp16, levels 1/2 (T/1), 1/4 (T/2), 1/8 (T/4), 1/16 (T/8)
  int g_dtype;
  int tl[4], hl[4], wl[4];
  int n, t, h, w;
  const float* b1;             // [64]   folded comb_1 bias + BN1
  const float* w2;             // [64][64] folded comb_2 * BN2 scale, row = output channel
  // tensor-core head, dense-video schedule: level 0 of clip c, frame t is g[0][c][t] for t < g0_lo, g[0][c][t - (g0_hi - g0_lo)]
  // for t >= g0_hi (g[0] then holds tl[0] = g0_lo + t - g0_hi edge frames per clip), and frame c * g0_step + t of the
  // video-level map g0_video (g0_video_t frames) in between.  g0_video == nullptr: g[0] holds whole clips.
  const void* g0_video; int g0_lo, g0_hi, g0_step, g0_video_t;
  const void* a_tab;           // tensor-core head: interpolation matrices of the frame geometry (launch_head_table)
  int tail_f16;                // tensor-core head: comb_2 / head operands in fp16 (1) or bf16 (0)
  const float* b2;             // [64]
  const float* wh;             // [6][64]  rows 0-1 segmentation head, 2-5 motion head
  const float* bh;             // [6]
  void* seg; void* motion;     // (N,2,T,H,W), (N,4,T,H,W)
  int out_dtype;               // CLASFV_F32 | CLASFV_BF16 | CLASFV_F16
  int out_kind;                // CLASFV_OUT_LOGITS | CLASFV_OUT_PROB
};
int launch_head(const HeadArgs& a, cudaStream_t stream);        // CUDA-core head (fp32 g)
// tcgen05 head (fp16 g, every level at the output's frame rate), decoder_umma.cu
int launch_head_umma(const HeadArgs& a, int num_sms, cudaStream_t stream);
size_t head_table_bytes(const HeadArgs& a);                              // 0: geometry not supported
int launch_head_table(const HeadArgs& a, void* tab, cudaStream_t stream);  // fills a.h x a.w's interpolation matrices (once per geometry)
// (n,tl,hl,wl,64) fp16 -> (n,t,hl,wl,64) fp16, linear along T, align_corners=True (pre-pass of the tcgen05 head)
int launch_temporal_upsample_f16(const void* in, void* out, int n, int tl, int t, int hl, int wl, cudaStream_t stream);

// fusion.cu
int launch_ingest_u8(const uint8_t* frames, int t, int h0, int w0, int bgr, float* out, int h, int w, uint32_t* minmax_dev, cudaStream_t s);
int launch_warp(const float* src, const float* flow, float* out, int n, int c, int h, int w, int nearest, cudaStream_t s);
int launch_motion_field(const float* flow, float* grid, int n, int h, int w, cudaStream_t s);
struct WarpFuseArgs {
  const void* prob; const void* motion; int dtype;
  int prob_planes;             // class planes per clip in `prob`: 2 (background, LV) or 1 (LV only); the LV plane is the last
  const int32_t* clip_start;   // device [n_clips]
  const int32_t* frame_lo;     // device [t_out]   first candidate clip for the frame
  const int32_t* frame_hi;     // device [t_out]   one past the last candidate clip
  int n_clips, clip_len, t_out, h, w, edge_hops, accumulate;
  float* acc; int32_t* cnt; uint8_t* mask; int32_t* area;
};
int launch_warp_fuse(const WarpFuseArgs& a, cudaStream_t s);
struct ShiftTable {            // device arrays [n_shifts]
  const int32_t* start; const int32_t* len; const int32_t* nclips; const int32_t* clip_base;
};
int launch_build_shift_clips(const float* video, int t, int h, int w, int clip_len, int n_shifts, int total_clips,
                             const int32_t* clip_shift /*device [total_clips]*/, ShiftTable tab, float* clips,
                             cudaStream_t s);
int launch_fuse_shift_votes(const void* prob, int dtype, int t, int h, int w, int clip_len, int step, int n_shifts,
                            ShiftTable tab, uint8_t* mask, int32_t* area, cudaStream_t s);
int launch_finalize_mask(const float* acc, int t, int h, int w, uint8_t* mask, int32_t* area, cudaStream_t s);
int launch_temporal_resample(const float* in, float* out, int channels, int l_in, int l_out, int64_t hw, cudaStream_t s);

}  // namespace clasfv
