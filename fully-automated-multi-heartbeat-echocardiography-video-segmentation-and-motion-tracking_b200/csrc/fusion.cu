// fusion.cu - bandwidth-bound kernels either side of the network: the warp primitive, the
// warp-and-fuse operator (F2), and the reference-exact shifted-pass fusion (F1).
//
// Reference semantics reproduced here (paths relative to the reference checkout):
//   generate_2dmotion_field          src/transform_utils.py:14-34   base grid linspace(-1,1,S) (an
//       align_corners=True grid) consumed by grid_sample(align_corners=False, bilinear, border):
//       zero flow is NOT the identity, x_src = j*W/(W-1) - 1/2.
//   divide_to_consecutive_clips      src/fuse_utils.py:16-33        temporal resample, align_corners=False
//   segment_a_video_with_fusion      src/fuse_utils.py:70-98        resample back, argmax, per-frame vote
#include "internal.h"
#include "umma_ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace clasfv {
namespace {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(__ldg(p)); }

// torch.linspace(-1, 1, steps)[idx] in fp32 (symmetric evaluation, as ATen does)
__device__ __forceinline__ float linspace_pm1(int idx, int steps) {
  if (steps <= 1) return -1.f;
  const float step = 2.f / (float)(steps - 1);
  return idx < steps / 2 ? __fmaf_rn(step, (float)idx, -1.f) : __fmaf_rn(-step, (float)(steps - idx - 1), 1.f);
}

// Every kernel in this file evaluates the sampling arithmetic through the helpers below with explicit rounding intrinsics.
// Left to the compiler, `a*b + c*d` may be contracted into fma(a, b, rn(c*d)) or fma(c, d, rn(a*b)) - it chose differently
// in different kernels, which made the direct-gather and the staged warp-fuse kernels disagree in the last bit.
// grid_sample's unnormalise for align_corners=False: ((g + 1) * size - 1) / 2.  ATen rounds rn(rn(g + 1) * size - 1) (one fused
// multiply-add) and halves it; halving is exact, so rn(x) / 2 == rn(x / 2) and the halving folds into the constants:
// rn(rn(g + 1) * (size / 2) - 1/2) is the same fp32 number with one instruction fewer (size / 2 is exact).
__device__ __forceinline__ float unnormalize_half(float g, float half_size) {
  return __fmaf_rn(__fadd_rn(g, 1.f), half_size, -0.5f);
}
__device__ __forceinline__ float unnormalize(float g, float size) { return unnormalize_half(g, size * 0.5f); }
// nw*a + ne*b + sw*c + se*d, accumulated left to right, one rounding per step
__device__ __forceinline__ float tap_sum(float a, float nw, float b, float ne, float c, float sw, float d, float se) {
  return __fmaf_rn(d, se, __fmaf_rn(c, sw, __fmaf_rn(b, ne, __fmul_rn(a, nw))));
}

struct Bilinear {
  int x0, y0;                 // north-west corner
  float nw, ne, sw, se;       // weights; out-of-range corners carry weight of an in-range clamp below
  bool x1ok, y1ok;
};

// grid_sample(align_corners=False, padding_mode="border", mode="bilinear") source location for
// output pixel (i, j) displaced by the normalised flow (fx, fy).
__device__ __forceinline__ Bilinear bilinear_setup(int i, int j, float fx, float fy, int h, int w) {
  const float gx = linspace_pm1(j, w) + fx, gy = linspace_pm1(i, h) + fy;
  float ix = unnormalize(gx, (float)w);
  float iy = unnormalize(gy, (float)h);
  ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  Bilinear b;
  b.x0 = (int)fx0; b.y0 = (int)fy0;
  const float x1 = fx0 + 1.f, y1 = fy0 + 1.f;
  b.nw = __fmul_rn(x1 - ix, y1 - iy);
  b.ne = __fmul_rn(ix - fx0, y1 - iy);
  b.sw = __fmul_rn(x1 - ix, iy - fy0);
  b.se = __fmul_rn(ix - fx0, iy - fy0);
  b.x1ok = b.x0 + 1 < w; b.y1ok = b.y0 + 1 < h;
  return b;
}

// same, with the base grid coordinates (linspace values of the pixel) already known
__device__ __forceinline__ Bilinear bilinear_setup_base(float bx, float by, float fx, float fy, int h, int w) {
  const float gx = bx + fx, gy = by + fy;
  float ix = unnormalize(gx, (float)w);
  float iy = unnormalize(gy, (float)h);
  ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  Bilinear b;
  b.x0 = (int)fx0; b.y0 = (int)fy0;
  const float x1 = fx0 + 1.f, y1 = fy0 + 1.f;
  b.nw = __fmul_rn(x1 - ix, y1 - iy);
  b.ne = __fmul_rn(ix - fx0, y1 - iy);
  b.sw = __fmul_rn(x1 - ix, iy - fy0);
  b.se = __fmul_rn(ix - fx0, iy - fy0);
  b.x1ok = b.x0 + 1 < w; b.y1ok = b.y0 + 1 < h;
  return b;
}

template <typename T>
__device__ __forceinline__ float bilinear_fetch(const T* __restrict__ plane, const Bilinear& b, int w) {
  const T* p = plane + (int64_t)b.y0 * w + b.x0;
  float v = __fmul_rn(ldf<T>(p), b.nw);
  if (b.x1ok) v = __fmaf_rn(ldf<T>(p + 1), b.ne, v);
  if (b.y1ok) v = __fmaf_rn(ldf<T>(p + w), b.sw, v);
  if (b.x1ok && b.y1ok) v = __fmaf_rn(ldf<T>(p + w + 1), b.se, v);
  return v;
}

__global__ void warp_kernel(const float* __restrict__ src, const float* __restrict__ flow, float* __restrict__ out,
                            int c, int h, int w, int nearest) {
  const int n = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= h * w) return;
  const int i = pix / w, j = pix % w;
  const int64_t hw = (int64_t)h * w;
  const float fx = __ldg(flow + (int64_t)n * 2 * hw + pix), fy = __ldg(flow + ((int64_t)n * 2 + 1) * hw + pix);
  if (nearest) {
    // grid_sample(mode="nearest", padding_mode="border", align_corners=False): clamp, then round half to even (nearbyint)
    const float gx = linspace_pm1(j, w) + fx, gy = linspace_pm1(i, h) + fy;
    float ix = unnormalize(gx, (float)w), iy = unnormalize(gy, (float)h);
    ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
    const int xs = (int)rintf(ix), ys = (int)rintf(iy);
    for (int k = 0; k < c; ++k) out[((int64_t)n * c + k) * hw + pix] = __ldg(src + ((int64_t)n * c + k) * hw + (int64_t)ys * w + xs);
    return;
  }
  const Bilinear b = bilinear_setup(i, j, fx, fy, h, w);
  for (int k = 0; k < c; ++k) out[((int64_t)n * c + k) * hw + pix] = bilinear_fetch<float>(src + ((int64_t)n * c + k) * hw, b, w);
}

__global__ void motion_field_kernel(const float* __restrict__ flow, float* __restrict__ grid, int h, int w) {
  const int n = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= h * w) return;
  const int i = pix / w, j = pix % w;
  const int64_t hw = (int64_t)h * w;
  float2 g;
  g.x = linspace_pm1(j, w) + __ldg(flow + (int64_t)n * 2 * hw + pix);
  g.y = linspace_pm1(i, h) + __ldg(flow + ((int64_t)n * 2 + 1) * hw + pix);
  reinterpret_cast<float2*>(grid)[(int64_t)n * hw + pix] = g;
}

// ------------------------------------------------------------------------------------------- ingest
// motion_segment.py:80-106 after the cv2 decode: uint8 frames (T,H0,W0,3) -> float, F.interpolate(size=(T,h,w),
// mode="trilinear", align_corners=True) (the frame count is unchanged, so this is a per-frame bilinear resize) ->
// zeroone_normalizer (echonet_dataset.py:38-50): per channel x -= min; x /= max-after-the-shift.
// Pass 1 resizes, writes the planar (3,T,h,w) fp32 video and reduces min / max per channel; pass 2 normalises in place.
// Values are >= 0, so the unsigned bit pattern of a float orders like the float: atomicMin / atomicMax on uint32.
struct LerpAC { int i0, i1; float l0, l1; };
__device__ __forceinline__ LerpAC lerp_align_corners(int dst, int in_size, int out_size) {
  LerpAC r;
  const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  const float src = scale * (float)dst;
  r.i0 = min((int)src, in_size - 1);
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

__global__ void __launch_bounds__(256) ingest_resize_kernel(const uint8_t* __restrict__ frames, int t, int h0, int w0, int bgr,
                                                            float* __restrict__ out, int h, int w, uint32_t* __restrict__ minmax) {
  const int ti = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  const bool valid = pix < h * w;
  float v[3] = {0.f, 0.f, 0.f};
  if (valid) {
    const int i = pix / w, j = pix % w;
    const LerpAC ly = lerp_align_corners(i, h0, h), lx = lerp_align_corners(j, w0, w);
    const uint8_t* f = frames + (int64_t)ti * h0 * w0 * 3;
    const uint8_t* p00 = f + ((int64_t)ly.i0 * w0 + lx.i0) * 3; const uint8_t* p01 = f + ((int64_t)ly.i0 * w0 + lx.i1) * 3;
    const uint8_t* p10 = f + ((int64_t)ly.i1 * w0 + lx.i0) * 3; const uint8_t* p11 = f + ((int64_t)ly.i1 * w0 + lx.i1) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int cs = bgr ? 2 - c : c;         // output channel c reads byte cs of the pixel
      // the nesting and rounding of ATen's linear upsampling: h0*(w0*a + w1*b) + h1*(w0*c + w1*d), no fused multiply-add
      const float top = __fadd_rn(__fmul_rn(lx.l0, (float)p00[cs]), __fmul_rn(lx.l1, (float)p01[cs]));
      const float bot = __fadd_rn(__fmul_rn(lx.l0, (float)p10[cs]), __fmul_rn(lx.l1, (float)p11[cs]));
      v[c] = __fadd_rn(__fmul_rn(ly.l0, top), __fmul_rn(ly.l1, bot));
      out[((int64_t)c * t + ti) * h * w + pix] = v[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float lo = valid ? v[c] : 3.0e38f, hi = valid ? v[c] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(minmax + 2 * c, __float_as_uint(lo)); atomicMax(minmax + 2 * c + 1, __float_as_uint(hi)); }
  }
}

__global__ void __launch_bounds__(256) ingest_normalize_kernel(float* __restrict__ out, int64_t per_channel, const uint32_t* __restrict__ minmax) {
  const int c = blockIdx.y;
  const float lo = __uint_as_float(minmax[2 * c]);
  const float range = __fsub_rn(__uint_as_float(minmax[2 * c + 1]), lo);       // the max after the shift
  float* o = out + (int64_t)c * per_channel;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < per_channel; i += (int64_t)gridDim.x * 256)
    o[i] = __fdiv_rn(__fsub_rn(o[i], lo), range);
}

// ------------------------------------------------------------------------------------------- F2
constexpr int WF_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(WF_THREADS) warp_fuse_kernel(const WarpFuseArgs a) {
  const int g = blockIdx.y;
  const int pix = blockIdx.x * WF_THREADS + threadIdx.x;
  const int hw = a.h * a.w;
  const bool valid = pix < hw;
  const int i = valid ? pix / a.w : 0, j = valid ? pix % a.w : 0;
  const float base_x = linspace_pm1(j, a.w), base_y = linspace_pm1(i, a.h);
  const int L = a.clip_len;
  // the LV plane of clip c (the last of its prob_planes class planes)
  const T* __restrict__ prob = static_cast<const T*>(a.prob) + (int64_t)(a.prob_planes - 1) * L * hw;
  const T* __restrict__ mot = static_cast<const T*>(a.motion);
  float* acc = a.acc + (int64_t)g * 2 * hw;
  float a1 = 0.f;          // LV votes of this call; what the accumulators already hold is added at the end
  int votes = 0;
  const int lo = __ldg(a.frame_lo + g), hi = __ldg(a.frame_hi + g);
  for (int c = lo; c < hi; ++c) {
    const int t = g - __ldg(a.clip_start + c);
    const T* pc = prob + (int64_t)c * a.prob_planes * L * hw;
    const T* mc = mot + (int64_t)c * 4 * L * hw;
    if (t >= 0 && t < L) {                                   // direct vote
      ++votes;
      if (valid) a1 += ldf<T>(pc + (int64_t)t * hw + pix);
    }
    int ts = t - 1;                                          // forward hop: frame ts -> ts + 1 == t
    if (ts >= 0 && ts < L && (a.edge_hops || ts + 1 < L)) {
      ++votes;
      if (valid) {
        const float fx = ldf<T>(mc + (int64_t)ts * hw + pix), fy = ldf<T>(mc + (int64_t)(L + ts) * hw + pix);
        const Bilinear b = bilinear_setup_base(base_x, base_y, fx, fy, a.h, a.w);
        a1 += bilinear_fetch<T>(pc + (int64_t)ts * hw, b, a.w);
      }
    }
    ts = t + 1;                                              // backward hop: frame ts -> ts - 1 == t
    if (ts >= 0 && ts < L && (a.edge_hops || ts >= 1)) {
      ++votes;
      if (valid) {
        const float fx = ldf<T>(mc + (int64_t)(2 * L + ts) * hw + pix), fy = ldf<T>(mc + (int64_t)(3 * L + ts) * hw + pix);
        const Bilinear b = bilinear_setup_base(base_x, base_y, fx, fy, a.h, a.w);
        a1 += bilinear_fetch<T>(pc + (int64_t)ts * hw, b, a.w);
      }
    }
  }
  // background sum = votes - LV sum (class probabilities and bilinear weights each sum to one): only the LV plane is read
  float a0 = __fsub_rn((float)votes, a1);
  if (a.accumulate && valid) { a0 = __fadd_rn(acc[pix], a0); a1 = __fadd_rn(acc[hw + pix], a1); }
  const bool lv = valid && (a1 > a0);
  if (valid) {
    acc[pix] = a0; acc[hw + pix] = a1;
    if (a.mask) a.mask[(int64_t)g * hw + pix] = lv ? 1 : 0;
  }
  if (a.area) {
    const int cnt = __syncthreads_count(lv);
    if (threadIdx.x == 0 && cnt) atomicAdd(a.area + g, cnt);
  }
  if (a.cnt && blockIdx.x == 0 && threadIdx.x == 0) a.cnt[g] = (a.accumulate ? a.cnt[g] : 0) + votes;
}

// ------------------------------------------------------------------------------------------- F2, staged
// Both kernels read ONLY the LV plane of every clip frame (north-star: "each clip's per-frame LV softmax
// is warped"): the LV class sum is accumulated and the background sum is votes - LV sum, because the two class
// probabilities and the four bilinear weights each sum to one.  That is 5 planes per clip frame instead of 6 and half the
// taps and the staging; at 224 x 224 a 16-bit LV plane is 100 KB, so two ring units fit and the staged kernel runs there too.
//
// The same operator with the gather sources staged in shared memory.  warp_fuse_kernel above is bound by the
// L1/LSU pipe (16 scattered global loads per clip and pixel); here a CTA owns one output frame g and one slice of
// its pixels, and for every hop (clip c, source frame ts) landing on g a producer warp bulk-copies the two class
// planes prob[c, :, ts] (contiguous H*W elements each) into a ring of shared-memory units (mbarrier full / empty).
// Consumer threads own fixed pixels for the whole CTA lifetime: running sums live in registers, flow and direct
// votes are read from global memory as coalesced loads, and the eight bilinear taps of a hop are LDS.  The
// order of additions per pixel is exactly warp_fuse_kernel's (clip by clip: direct, forward hop, backward hop), so
// both kernels give bit-identical sums.  No float atomics, no intermediate warped volume.
constexpr int WS_MAX_UNITS = 8;
// CTA shape (measured in round 1 on the two-plane kernel, config 3, 256 clips, ms at zero / 4 px flow; re-measured in round 2 on
// the LV-plane kernel: 512x7 0.63 / 0.66 fp32, 0.66 / 0.68 bf16; 768x5 the same; 1024x4 0.74 / 0.75 fp32, 0.86 / 0.88 bf16).  An item is a pair of adjacent pixels, loaded with one
// instruction (8 bytes fp32, 4 bytes bf16).  512 threads x 7 pairs (128 registers, no spills worth the name, 14 independent
// tap chains per thread) is the best shape for both element types:
//   fp32   512x7 0.80 / 0.96    384x9 0.85 / 0.98    768x5 0.94 / 1.09
//   bf16   512x7 0.82 / 0.92    768x5 1.10 / 1.17    1024x4 1.14 / 1.22 (64-register cap: spills and constant reloads)
// Also measured and rejected for fp32: single pixels per item, so that the lanes' taps are one word apart instead of two
// (no 2-way bank conflict under a smooth flow): twice the global load instructions cost more than the conflicts
// (1.05-1.18 ms in round 1; again on the flow-staged kernel, 14 single-pixel items per thread: 0.50-0.58 ms against 0.43-0.51,
// profiles/r02u_warp_fuse_fp32_items.jsonl - twice the LDS / FADD instructions for the flows and the sums).
constexpr int WS_THREADS_PER_CTA = 512, WS_ITEMS = 7;
constexpr int WS_PP = 2;

template <typename T> struct Item;
template <> struct Item<float> {
  using Raw = float2;
  static __device__ __forceinline__ Raw ld_raw(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
  static __device__ __forceinline__ void unpack(const Raw& o, float (&v)[2]) { v[0] = o.x; v[1] = o.y; }
  static __device__ __forceinline__ void ld(const float* p, float (&v)[2]) { unpack(ld_raw(p), v); }
  static __device__ __forceinline__ void lds(const float* p, float (&v)[2]) { unpack(*reinterpret_cast<const float2*>(p), v); }
  static __device__ __forceinline__ void ld_acc(const float* p, float (&v)[2]) {
    const float2 o = *reinterpret_cast<const float2*>(p); v[0] = o.x; v[1] = o.y;
  }
  static __device__ __forceinline__ void st_acc(float* p, const float (&v)[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
  static __device__ __forceinline__ void st_mask(uint8_t* p, const int (&m)[2]) {
    *reinterpret_cast<uchar2*>(p) = make_uchar2((unsigned char)m[0], (unsigned char)m[1]);
  }
};
template <> struct Item<__nv_bfloat16> {
  using Raw = uint32_t;
  static __device__ __forceinline__ Raw ld_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[2]) {
    v[0] = __uint_as_float(r << 16); v[1] = __uint_as_float(r & 0xffff0000u);
  }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[2]) { unpack(ld_raw(p), v); }
  static __device__ __forceinline__ void lds(const __nv_bfloat16* p, float (&v)[2]) { unpack(*reinterpret_cast<const uint32_t*>(p), v); }
  static __device__ __forceinline__ void ld_acc(const float* p, float (&v)[2]) {
    const float2 o = *reinterpret_cast<const float2*>(p); v[0] = o.x; v[1] = o.y;
  }
  static __device__ __forceinline__ void st_acc(float* p, const float (&v)[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
  static __device__ __forceinline__ void st_mask(uint8_t* p, const int (&m)[2]) {
    *reinterpret_cast<uchar2*>(p) = make_uchar2((unsigned char)m[0], (unsigned char)m[1]);
  }
};
template <> struct Item<__half> {
  using Raw = uint32_t;
  static __device__ __forceinline__ Raw ld_raw(const __half* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[2]) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&r)); v[0] = f.x; v[1] = f.y;
  }
  static __device__ __forceinline__ void ld(const __half* p, float (&v)[2]) { unpack(ld_raw(p), v); }
  static __device__ __forceinline__ void lds(const __half* p, float (&v)[2]) { unpack(*reinterpret_cast<const uint32_t*>(p), v); }
  static __device__ __forceinline__ void ld_acc(const float* p, float (&v)[2]) {
    const float2 o = *reinterpret_cast<const float2*>(p); v[0] = o.x; v[1] = o.y;
  }
  static __device__ __forceinline__ void st_acc(float* p, const float (&v)[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
  static __device__ __forceinline__ void st_mask(uint8_t* p, const int (&m)[2]) {
    *reinterpret_cast<uchar2*>(p) = make_uchar2((unsigned char)m[0], (unsigned char)m[1]);
  }
};
template <typename T> __device__ __forceinline__ float lds_val(const T* p);
template <> __device__ __forceinline__ float lds_val<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float lds_val<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float lds_val<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short*>(p)) << 16);
}
// Branch-free bilinear taps for the staged kernel.  grid_sample(padding_mode="border") clamps the source coordinate to
// [0, S-1] BEFORE splitting it, so a corner that falls outside the plane always has weight exactly 0 (x0 == W-1 implies
// ix == W-1).  Instead of skipping that corner the north-west corner is clamped to (H-2, W-2): at ix == W-1 the pair
// (W-2, W-1) then gets the weights (0, 1) where grid_sample uses (W-1, -) with (1, 0).  The four-term sum
// nw*a + ne*b + sw*c + se*d keeps its order, a zero-weight term adds exactly 0, and the non-zero products are the same
// two numbers in the same order, so the result is bit-identical to warp_fuse_kernel's conditional taps - with no
// select, no per-corner offset and the +1 neighbours at immediate offsets.  (x0+1) - ix is written 1 - (ix - x0): ix - x0
// is exact, so both are one rounding of the same real number.
struct Taps { int p; float nw, ne, sw, se; };
struct TapGeom {                      // loop invariants of taps_setup, converted once per thread
  float wf, whalf, hhalf, wm1, hm1, wm2, hm2;
  __device__ __forceinline__ TapGeom(int h, int w)
      : wf((float)w), whalf(0.5f * (float)w), hhalf(0.5f * (float)h), wm1((float)(w - 1)), hm1((float)(h - 1)), wm2((float)(w - 2)), hm2((float)(h - 2)) {}
};
// CLAMP_NW = false: the plane is followed by W + 2 zero elements (flow-staged ring units), so the north-west corner may sit
// on the last row / column as grid_sample has it - the corners past the plane have weight exactly 0 and read zeros or
// the next row's first pixel (a finite probability): the same sum again, two instructions fewer per tap set.
template <bool CLAMP_NW>
__device__ __forceinline__ Taps taps_setup(float bx, float by, float fx, float fy, const TapGeom& g) {
  const float gx = bx + fx, gy = by + fy;
  float ix = unnormalize_half(gx, g.whalf);
  float iy = unnormalize_half(gy, g.hhalf);
  ix = fminf(g.wm1, fmaxf(ix, 0.f));
  iy = fminf(g.hm1, fmaxf(iy, 0.f));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  if (CLAMP_NW) { fx0 = fminf(fx0, g.wm2); fy0 = fminf(fy0, g.hm2); }
  const float wx1 = ix - fx0, wy1 = iy - fy0;
  const float wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  Taps t;
  t.nw = __fmul_rn(wx0, wy0);
  t.ne = __fmul_rn(wx1, wy0);
  t.sw = __fmul_rn(wx0, wy1);
  t.se = __fmul_rn(wx1, wy1);
  t.p = (int)fmaf(fy0, g.wf, fx0);            // y0*W + x0: small integers, exact in fp32
  return t;
}
template <typename T>
__device__ __forceinline__ float taps_fetch(const T* plane, const Taps& t, int w) {
  const T* p = plane + t.p;
  return tap_sum(lds_val<T>(p), t.nw, lds_val<T>(p + 1), t.ne, lds_val<T>(p + w), t.sw, lds_val<T>(p + w + 1), t.se);
}

// FLOWS = true (round 2, the default whenever two such units fit): a ring unit also carries the CTA's pixel slice of the hop's
// two flow planes, [LV plane][W + 2 zeros][flow x slice][flow y slice], all three filled by the producer's bulk copies.  With
// the flows read from global memory by the consumers, 47 % of a warp's latency per instruction was the wait on those loads
// (ncu, profiles/r02_wf_ncu.csv: long scoreboard 3.3 of 7.0 cycles per issue; every warp of the CTA issued them at the
// same point - right after being released by the same barrier); now DRAM latency is hidden by the depth of the ring, the
// consumers' flow reads are conflict-free LDS, the 28 registers that held a hop's flows are gone, and the only global loads left
// to the consumers - the direct votes - are requested one clip ahead.
template <typename T, int WS_THREADS, int ITEMS, bool FLOWS>
__global__ void __launch_bounds__(WS_THREADS, 1) warp_fuse_staged_kernel(const WarpFuseArgs a, int n_units, int slice_pix, uint32_t unit_bytes,
                                                                          uint32_t pad_bytes, uint32_t slice_stride) {
  extern __shared__ __align__(128) uint8_t ws_smem[];
  using namespace ptx;
  using Raw = typename Item<T>::Raw;
  constexpr int PP = WS_PP, WS_CONSUMERS = WS_THREADS - 32;   // warp 0 = producer
  const int g = blockIdx.y;
  const int hw = a.h * a.w;
  const int L = a.clip_len;
  const uint32_t plane_bytes = (uint32_t)hw * (uint32_t)sizeof(T);
  const uint32_t sbase = smem_u32(ws_smem);
  const uint32_t bar0 = sbase + (uint32_t)n_units * unit_bytes;         // full[u] at bar0 + 8u, empty[u] at bar0 + 8(n_units + u)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* __restrict__ prob = static_cast<const T*>(a.prob) + (int64_t)(a.prob_planes - 1) * L * hw;     // LV planes
  const T* __restrict__ mot = static_cast<const T*>(a.motion);
  const int64_t clip_elems = (int64_t)a.prob_planes * L * hw;
  const int lo = __ldg(a.frame_lo + g), hi = __ldg(a.frame_hi + g);
  const int pix0 = blockIdx.x * slice_pix;
  const int pix_end = min(hw, pix0 + slice_pix);

  if (threadIdx.x == 0) {
    for (int u = 0; u < n_units; ++u) { mbar_init(bar0 + 8u * u, 1); mbar_init(bar0 + 8u * (n_units + u), WS_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (FLOWS) {          // the zero elements behind every unit's plane (never written by a bulk copy)
    const int words = (int)(pad_bytes >> 2);
    for (int i = threadIdx.x; i < n_units * words; i += WS_THREADS)
      *reinterpret_cast<uint32_t*>(ws_smem + (size_t)(i / words) * unit_bytes + plane_bytes + 4u * (uint32_t)(i % words)) = 0u;
  }
  __syncthreads();

  if (warp == 0) {
    // ---------------------------------------------------------------- producer: the bulk copies of every hop
    if (lane == 0) {
      const uint32_t slice_bytes = (uint32_t)(pix_end - pix0) * (uint32_t)sizeof(T);
      const uint32_t tx = plane_bytes + (FLOWS ? 2u * slice_bytes : 0u);
      int item = 0;
      for (int c = lo; c < hi; ++c) {
        const int t = g - __ldg(a.clip_start + c);
        const T* pc = prob + (int64_t)c * clip_elems;
        const T* mc = mot + (int64_t)c * 4 * L * hw + pix0;
        for (int hop = 0; hop < 2; ++hop) {
          const int ts = hop == 0 ? t - 1 : t + 1;
          const bool on = ts >= 0 && ts < L && (a.edge_hops || (hop == 0 ? ts + 1 < L : ts >= 1));
          if (!on) continue;
          const int u = item % n_units; const uint32_t ph = (uint32_t)(item / n_units) & 1u;
          mbar_wait(bar0 + 8u * (n_units + u), ph ^ 1u);
          mbar_arrive_expect_tx(bar0 + 8u * u, tx);
          const uint32_t dst = sbase + (uint32_t)u * unit_bytes;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(pc + (int64_t)ts * hw), "r"(plane_bytes), "r"(bar0 + 8u * u) : "memory");
          if (FLOWS) {
            const uint32_t fdst = dst + plane_bytes + pad_bytes;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(fdst), "l"(mc + (int64_t)((hop == 0 ? 0 : 2 * L) + ts) * hw), "r"(slice_bytes), "r"(bar0 + 8u * u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(fdst + slice_stride), "l"(mc + (int64_t)((hop == 0 ? L : 3 * L) + ts) * hw), "r"(slice_bytes), "r"(bar0 + 8u * u) : "memory");
          }
          ++item;
        }
      }
    }
    // the producer warp also owns the vote count of the frame
    if (lane == 1 && a.cnt && blockIdx.x == 0) {
      int votes = 0;
      for (int c = lo; c < hi; ++c) {
        const int t = g - __ldg(a.clip_start + c);
        if (t >= 0 && t < L) ++votes;
        int ts = t - 1;
        if (ts >= 0 && ts < L && (a.edge_hops || ts + 1 < L)) ++votes;
        ts = t + 1;
        if (ts >= 0 && ts < L && (a.edge_hops || ts >= 1)) ++votes;
      }
      a.cnt[g] = (a.accumulate ? a.cnt[g] : 0) + votes;
    }
  } else {
    // ---------------------------------------------------------------- consumers: fixed items (pixel pairs), sums in registers
    const int ctid = threadIdx.x - 32;
    const int p0 = pix0 + PP * ctid;                       // item k of this thread starts at pixel p0 + k * PP * WS_CONSUMERS
    constexpr int ITEM_STEP = PP * WS_CONSUMERS;
    // items past the slice end are simply not there: n_items is the same for all but the last lanes of a slice
    const int n_items = p0 < pix_end ? min(ITEMS, (pix_end - p0 + ITEM_STEP - 1) / ITEM_STEP) : 0;
    float bx[ITEMS][PP], by[ITEMS];
    float s1[ITEMS][PP];                                      // LV votes of this call
    float* acc = a.acc + (int64_t)g * 2 * hw;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const int p = p0 + k * ITEM_STEP;
      const bool on = k < n_items;
      const int i = on ? p / a.w : 0, j = on ? p % a.w : 0;     // PP == 2: W is even, a pair never straddles rows
      by[k] = linspace_pm1(i, a.h);
#pragma unroll
      for (int q = 0; q < PP; ++q) { bx[k][q] = linspace_pm1(j + q < a.w ? j + q : j, a.w); s1[k][q] = 0.f; }
    }
    int votes = 0;
    const TapGeom geom(a.h, a.w);
    int item = 0;
    // direct votes of clip c: requested while clip c - 1 is being processed (FLOWS), consumed first thing in clip c's turn
    Raw dv[ITEMS];
    auto request_direct = [&](int c) {
      const int t = g - __ldg(a.clip_start + c);
      if (t < 0 || t >= L) return;
      const T* d1 = prob + (int64_t)c * clip_elems + (int64_t)t * hw + p0;
#pragma unroll
      for (int k = 0; k < ITEMS; ++k)
        if (k < n_items) dv[k] = Item<T>::ld_raw(d1 + k * ITEM_STEP);
    };
    if (FLOWS && lo < hi) request_direct(lo);
    for (int c = lo; c < hi; ++c) {
      const int t = g - __ldg(a.clip_start + c);
      if (t >= 0 && t < L) {                                   // direct vote
        ++votes;
        if (!FLOWS) request_direct(c);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
          if (k >= n_items) break;
          float v1[PP];
          Item<T>::unpack(dv[k], v1);
#pragma unroll
          for (int q = 0; q < PP; ++q) s1[k][q] += v1[q];
        }
      }
      if (FLOWS && c + 1 < hi) request_direct(c + 1);
      const T* mc = mot + (int64_t)c * 4 * L * hw + p0;
#pragma unroll
      for (int hop = 0; hop < 2; ++hop) {
        const int ts = hop == 0 ? t - 1 : t + 1;
        const bool on = ts >= 0 && ts < L && (a.edge_hops || (hop == 0 ? ts + 1 < L : ts >= 1));
        if (!on) continue;
        ++votes;
        const int u = item % n_units; const uint32_t ph = (uint32_t)(item / n_units) & 1u;
        const T* u1 = reinterpret_cast<const T*>(ws_smem + (size_t)u * unit_bytes);
        if (FLOWS) {
          const T* ufx = reinterpret_cast<const T*>(ws_smem + (size_t)u * unit_bytes + plane_bytes + pad_bytes) + PP * ctid;
          const T* ufy = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(ufx) + slice_stride);
          mbar_wait(bar0 + 8u * u, ph);
#pragma unroll
          for (int k = 0; k < ITEMS; ++k) {
            if (k >= n_items) break;
            float fx[PP], fy[PP];
            Item<T>::lds(ufx + k * ITEM_STEP, fx); Item<T>::lds(ufy + k * ITEM_STEP, fy);
#pragma unroll
            for (int q = 0; q < PP; ++q) {
              const Taps ta = taps_setup<false>(bx[k][q], by[k], fx[q], fy[q], geom);
              s1[k][q] += taps_fetch<T>(u1, ta, a.w);
            }
          }
        } else {
          const T* fxp = mc + (int64_t)((hop == 0 ? 0 : 2 * L) + ts) * hw;
          const T* fyp = mc + (int64_t)((hop == 0 ? L : 3 * L) + ts) * hw;
          float fx[ITEMS][PP], fy[ITEMS][PP];
#pragma unroll
          for (int k = 0; k < ITEMS; ++k)
            if (k < n_items) { Item<T>::ld(fxp + k * ITEM_STEP, fx[k]); Item<T>::ld(fyp + k * ITEM_STEP, fy[k]); }
          mbar_wait(bar0 + 8u * u, ph);
#pragma unroll
          for (int k = 0; k < ITEMS; ++k) {
            if (k >= n_items) break;
#pragma unroll
            for (int q = 0; q < PP; ++q) {
              const Taps ta = taps_setup<true>(bx[k][q], by[k], fx[k][q], fy[k][q], geom);
              s1[k][q] += taps_fetch<T>(u1, ta, a.w);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8u * (n_units + u));
        ++item;
      }
    }
    int lv_count = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      if (k >= n_items) break;
      const int p = p0 + k * ITEM_STEP;
      // background sum = votes - LV sum; previous content added last (the same expressions as warp_fuse_kernel: same bits)
      float s0[PP];
#pragma unroll
      for (int q = 0; q < PP; ++q) s0[q] = __fsub_rn((float)votes, s1[k][q]);
      if (a.accumulate) {
        float o0[PP], o1[PP];
        Item<T>::ld_acc(acc + p, o0); Item<T>::ld_acc(acc + hw + p, o1);
#pragma unroll
        for (int q = 0; q < PP; ++q) { s0[q] = __fadd_rn(o0[q], s0[q]); s1[k][q] = __fadd_rn(o1[q], s1[k][q]); }
      }
      Item<T>::st_acc(acc + p, s0);
      Item<T>::st_acc(acc + hw + p, s1[k]);
      int m[PP];
#pragma unroll
      for (int q = 0; q < PP; ++q) { m[q] = s1[k][q] > s0[q] ? 1 : 0; lv_count += m[q]; }
      if (a.mask) Item<T>::st_mask(a.mask + (int64_t)g * hw + p, m);
    }
    if (a.area) {
      lv_count = __reduce_add_sync(0xffffffffu, lv_count);
      if (lane == 0 && lv_count) atomicAdd(a.area + g, lv_count);
    }
  }
}


// ------------------------------------------------------------------------------------------- F1
struct Lerp { int i0, i1; float l0, l1; };
// PyTorch linear resample, align_corners=False: src = max(scale*(dst+0.5)-0.5, 0)
__device__ __forceinline__ Lerp lerp_tap(int dst, int in_size, int out_size) {
  Lerp r;
  const float scale = (float)in_size / (float)out_size;
  const float src = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  r.i0 = min((int)src, in_size - 1);
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = fminf(fmaxf(src - (float)r.i0, 0.f), 1.f);
  r.l0 = 1.f - r.l1;
  return r;
}

__global__ void build_shift_clips_kernel(const float* __restrict__ video, int t_video, int64_t hw4, int clip_len,
                                         const int32_t* __restrict__ clip_shift, ShiftTable tab,
                                         float* __restrict__ clips) {
  const int clip = blockIdx.z, tt = blockIdx.y;
  const int64_t p4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p4 >= hw4) return;
  const int k = __ldg(clip_shift + clip);
  const int start = __ldg(tab.start + k), len = __ldg(tab.len + k), lr = clip_len * __ldg(tab.nclips + k);
  const int f = (clip - __ldg(tab.clip_base + k)) * clip_len + tt;
  Lerp r;
  if (lr == len) { r.i0 = r.i1 = f; r.l0 = 1.f; r.l1 = 0.f; }
  else r = lerp_tap(f, len, lr);
  const float4* v = reinterpret_cast<const float4*>(video);
  float4* o = reinterpret_cast<float4*>(clips);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 x0 = __ldg(v + ((int64_t)c * t_video + start + r.i0) * hw4 + p4);
    float4 y = x0;
    if (lr != len) {
      const float4 x1 = __ldg(v + ((int64_t)c * t_video + start + r.i1) * hw4 + p4);
      y.x = r.l0 * x0.x + r.l1 * x1.x; y.y = r.l0 * x0.y + r.l1 * x1.y;
      y.z = r.l0 * x0.z + r.l1 * x1.z; y.w = r.l0 * x0.w + r.l1 * x1.w;
    }
    o[(((int64_t)clip * 3 + c) * clip_len + tt) * hw4 + p4] = y;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fuse_shift_votes_kernel(const T* __restrict__ prob, int t_video, int hw, int clip_len,
                                                               int step, int n_shifts, ShiftTable tab,
                                                               uint8_t* __restrict__ mask, int32_t* __restrict__ area) {
  const int i = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = pix < hw;
  int ones = 0, voters = 0;
  const int kmax = i == 0 ? 1 : min(i, n_shifts);
  for (int k = 0; k < kmax; ++k) {
    const int j = i - k * step;
    if (j < 0) break;
    const int len = __ldg(tab.len + k), lr = clip_len * __ldg(tab.nclips + k), cb = __ldg(tab.clip_base + k);
    if (j >= (lr == len ? lr : len)) continue;   // the reference raises IndexError here; the host rejects such plans
    ++voters;
    if (!valid) continue;
    float p0, p1;
    if (lr == len) {
      const T* pc = prob + ((int64_t)(cb + j / clip_len) * 2 * clip_len + j % clip_len) * hw + pix;
      p0 = ldf<T>(pc); p1 = ldf<T>(pc + (int64_t)clip_len * hw);
    } else {
      const Lerp r = lerp_tap(j, lr, len);
      const T* pa = prob + ((int64_t)(cb + r.i0 / clip_len) * 2 * clip_len + r.i0 % clip_len) * hw + pix;
      const T* pb = prob + ((int64_t)(cb + r.i1 / clip_len) * 2 * clip_len + r.i1 % clip_len) * hw + pix;
      p0 = r.l0 * ldf<T>(pa) + r.l1 * ldf<T>(pb);
      p1 = r.l0 * ldf<T>(pa + (int64_t)clip_len * hw) + r.l1 * ldf<T>(pb + (int64_t)clip_len * hw);
    }
    ones += p1 > p0 ? 1 : 0;
  }
  const bool lv = valid && (2 * ones > voters);
  if (valid) mask[(int64_t)i * hw + pix] = lv ? 1 : 0;
  if (area) {
    const int cnt = __syncthreads_count(lv);
    if (threadIdx.x == 0 && cnt) atomicAdd(area + i, cnt);
  }
}

// mask / LV area from already fused class sums (after a multi-GPU halo exchange added the neighbours' votes)
__global__ void finalize_mask_kernel(const float* __restrict__ acc, int hw, uint8_t* __restrict__ mask, int32_t* __restrict__ area) {
  const int g = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = pix < hw;
  const bool lv = valid && acc[((int64_t)g * 2 + 1) * hw + pix] > acc[(int64_t)g * 2 * hw + pix];
  if (valid && mask) mask[(int64_t)g * hw + pix] = lv ? 1 : 0;
  if (area) {
    const int cnt = __syncthreads_count(lv);
    if (threadIdx.x == 0 && cnt) atomicAdd(area + g, cnt);
  }
}

__global__ void temporal_resample_kernel(const float* __restrict__ in, float* __restrict__ out, int l_in, int l_out, int64_t hw) {
  const int d = blockIdx.y, c = blockIdx.z;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hw) return;
  const Lerp r = lerp_tap(d, l_in, l_out);
  const float* base = in + (int64_t)c * l_in * hw;
  out[((int64_t)c * l_out + d) * hw + p] = r.l0 * __ldg(base + (int64_t)r.i0 * hw + p) + r.l1 * __ldg(base + (int64_t)r.i1 * hw + p);
}

}  // namespace

int launch_ingest_u8(const uint8_t* frames, int t, int h0, int w0, int bgr, float* out, int h, int w, uint32_t* minmax_dev, cudaStream_t s) {
  static const uint32_t init[6] = {0x7f7fffffu, 0u, 0x7f7fffffu, 0u, 0x7f7fffffu, 0u};      // {+FLT_MAX, 0} per channel
  CLASFV_CUDA(cudaMemcpyAsync(minmax_dev, init, sizeof(init), cudaMemcpyHostToDevice, s));
  ingest_resize_kernel<<<dim3((unsigned)cdiv((int64_t)h * w, 256), (unsigned)t), 256, 0, s>>>(frames, t, h0, w0, bgr, out, h, w, minmax_dev);
  CLASFV_CUDA(cudaGetLastError());
  const int64_t per_channel = (int64_t)t * h * w;
  ingest_normalize_kernel<<<dim3((unsigned)std::min<int64_t>(cdiv(per_channel, 256), 1184), 3u), 256, 0, s>>>(out, per_channel, minmax_dev);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

int launch_warp(const float* src, const float* flow, float* out, int n, int c, int h, int w, int nearest, cudaStream_t s) {
  dim3 grid((unsigned)cdiv((int64_t)h * w, 256), (unsigned)n);
  warp_kernel<<<grid, 256, 0, s>>>(src, flow, out, c, h, w, nearest);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

int launch_motion_field(const float* flow, float* grid_out, int n, int h, int w, cudaStream_t s) {
  dim3 grid((unsigned)cdiv((int64_t)h * w, 256), (unsigned)n);
  motion_field_kernel<<<grid, 256, 0, s>>>(flow, grid_out, h, w);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

struct StagedPlan { int units, slices, slice_pix; uint32_t unit_bytes, pad_bytes, slice_stride; size_t smem; };

template <typename T, int THREADS, int ITEMS, bool FLOWS>
static cudaError_t launch_staged(const WarpFuseArgs& a, const StagedPlan& pl, cudaStream_t s) {
  cudaError_t e = allow_max_dynamic_smem(warp_fuse_staged_kernel<T, THREADS, ITEMS, FLOWS>);
  if (e != cudaSuccess) return e;
  warp_fuse_staged_kernel<T, THREADS, ITEMS, FLOWS><<<dim3((unsigned)pl.slices, (unsigned)a.t_out), THREADS, pl.smem, s>>>(
      a, pl.units, pl.slice_pix, pl.unit_bytes, pl.pad_bytes, pl.slice_stride);
  return cudaSuccess;
}
template <int THREADS, int ITEMS, bool FLOWS>
static cudaError_t launch_staged_typed(const WarpFuseArgs& a, const StagedPlan& pl, cudaStream_t s) {
  return a.dtype == CLASFV_F32   ? launch_staged<float, THREADS, ITEMS, FLOWS>(a, pl, s)
         : a.dtype == CLASFV_F16 ? launch_staged<__half, THREADS, ITEMS, FLOWS>(a, pl, s)
                                 : launch_staged<__nv_bfloat16, THREADS, ITEMS, FLOWS>(a, pl, s);
}

// Ring geometry of the staged kernel for a CTA of `threads` threads x `items` pixel pairs.  `flows`: the units also carry the
// slice's flow planes (and W + 2 zeros behind the LV plane).  Slices are multiples of 8 pixels, so that every bulk copy starts
// and ends on 16 bytes.  Returns false when fewer than two units fit.
static bool plan_staged(const WarpFuseArgs& a, int threads, int items, bool flows, StagedPlan* pl) {
  const int64_t hw = (int64_t)a.h * a.w;
  const size_t es = a.dtype == CLASFV_F32 ? 4 : 2;
  const size_t budget = 227 * 1024 - 256;
  const int per_cta = (threads - 32) * WS_PP * items;
  pl->slices = (int)cdiv(hw, per_cta);
  pl->slice_pix = (int)(cdiv(cdiv(hw, pl->slices), 8) * 8);
  if (pl->slice_pix > per_cta) { pl->slices += 1; pl->slice_pix = (int)(cdiv(cdiv(hw, pl->slices), 8) * 8); }
  if (pl->slice_pix > per_cta) return false;
  pl->slices = (int)cdiv(hw, pl->slice_pix);
  pl->pad_bytes = flows ? (uint32_t)(cdiv((a.w + 2) * es, 16) * 16) : 0u;
  pl->slice_stride = flows ? (uint32_t)(pl->slice_pix * es) : 0u;
  pl->unit_bytes = (uint32_t)(hw * es) + pl->pad_bytes + 2u * pl->slice_stride;
  pl->units = (int)std::min<size_t>((budget - 16 * WS_MAX_UNITS) / pl->unit_bytes, (size_t)WS_MAX_UNITS);
  pl->smem = (size_t)pl->units * pl->unit_bytes + 16 * (size_t)pl->units;
  return pl->units >= 2;
}

int launch_warp_fuse(const WarpFuseArgs& a, cudaStream_t s) {
  if (a.area) CLASFV_CUDA(cudaMemsetAsync(a.area, 0, sizeof(int32_t) * a.t_out, s));
  // staged kernel: needs at least two ring units in shared memory, an even width (pixel pairs) and 16-byte aligned planes
  // (bulk copies); otherwise the direct-gather kernel runs.  Units that also carry the flows are preferred (112 x 112: four
  // 49 KB units in a 16-bit type, two 99 KB units in fp32); a 16-bit 224 x 224 plane leaves room for LV-only units.
  {
    const int64_t hw = (int64_t)a.h * a.w;
    const size_t es = a.dtype == CLASFV_F32 ? 4 : 2;
    static const bool no_staged = getenv("CLASFV_WARP_FUSE_DIRECT") != nullptr;
    static const bool no_flows = getenv("CLASFV_WARP_FUSE_NO_FLOW_STAGING") != nullptr;
    const bool aligned = (hw * es) % 16 == 0 && a.w % 2 == 0 && ((uintptr_t)a.prob % 16) == 0 && ((uintptr_t)a.motion % 16) == 0 &&
                         ((uintptr_t)a.acc % 8) == 0 && (!a.mask || ((uintptr_t)a.mask % 2) == 0);
    if (aligned && a.h >= 2 && !no_staged) {        // (w >= 2 follows from the even width)
      StagedPlan pl;
      cudaError_t e = cudaErrorInvalidValue; bool launched = false;
      // (measured on config 3, ms at 0 / 4 px flow, profiles/r02w_warp_fuse_variants.jsonl: 512 x 7 with the fewest slices
      // 0.43 / 0.50 fp32, 0.45 / 0.49 bf16; 1 024 threads x 4 pairs 0.46 / 0.51, 0.49 / 0.51; three slices 0.45 / 0.50, 0.47 / 0.50;
      // four slices 0.50 / 0.55, 0.51 / 0.54; flows from global memory, same build: 0.56 / 0.59, 0.57 / 0.59)
      if (!no_flows && plan_staged(a, WS_THREADS_PER_CTA, WS_ITEMS, true, &pl)) { e = launch_staged_typed<WS_THREADS_PER_CTA, WS_ITEMS, true>(a, pl, s); launched = true; }
      else if (plan_staged(a, WS_THREADS_PER_CTA, WS_ITEMS, false, &pl)) { e = launch_staged_typed<WS_THREADS_PER_CTA, WS_ITEMS, false>(a, pl, s); launched = true; }
      if (launched) {
        CLASFV_CUDA(e);
        CLASFV_CUDA(cudaGetLastError());
        return CLASFV_OK;
      }
    }
  }
  dim3 grid((unsigned)cdiv((int64_t)a.h * a.w, WF_THREADS), (unsigned)a.t_out);
  if (a.dtype == CLASFV_F32) warp_fuse_kernel<float><<<grid, WF_THREADS, 0, s>>>(a);
  else if (a.dtype == CLASFV_F16) warp_fuse_kernel<__half><<<grid, WF_THREADS, 0, s>>>(a);
  else warp_fuse_kernel<__nv_bfloat16><<<grid, WF_THREADS, 0, s>>>(a);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

int launch_build_shift_clips(const float* video, int t, int h, int w, int clip_len, int n_shifts, int total_clips,
                             const int32_t* clip_shift, ShiftTable tab, float* clips, cudaStream_t s) {
  (void)n_shifts;
  const int64_t hw4 = (int64_t)h * w / 4;
  dim3 grid((unsigned)cdiv(hw4, 128), (unsigned)clip_len, (unsigned)total_clips);
  build_shift_clips_kernel<<<grid, 128, 0, s>>>(video, t, hw4, clip_len, clip_shift, tab, clips);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

int launch_fuse_shift_votes(const void* prob, int dtype, int t, int h, int w, int clip_len, int step, int n_shifts,
                            ShiftTable tab, uint8_t* mask, int32_t* area, cudaStream_t s) {
  if (area) CLASFV_CUDA(cudaMemsetAsync(area, 0, sizeof(int32_t) * t, s));
  dim3 grid((unsigned)cdiv((int64_t)h * w, 256), (unsigned)t);
  if (dtype == CLASFV_F32)
    fuse_shift_votes_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(prob), t, h * w, clip_len, step, n_shifts, tab, mask, area);
  else if (dtype == CLASFV_F16)
    fuse_shift_votes_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(prob), t, h * w, clip_len, step, n_shifts, tab, mask, area);
  else
    fuse_shift_votes_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(prob), t, h * w, clip_len, step, n_shifts, tab, mask, area);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

int launch_finalize_mask(const float* acc, int t, int h, int w, uint8_t* mask, int32_t* area, cudaStream_t s) {
  if (area) CLASFV_CUDA(cudaMemsetAsync(area, 0, sizeof(int32_t) * t, s));
  dim3 grid((unsigned)cdiv((int64_t)h * w, 256), (unsigned)t);
  finalize_mask_kernel<<<grid, 256, 0, s>>>(acc, h * w, mask, area);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

int launch_temporal_resample(const float* in, float* out, int channels, int l_in, int l_out, int64_t hw, cudaStream_t s) {
  dim3 grid((unsigned)cdiv(hw, 256), (unsigned)l_out, (unsigned)channels);
  temporal_resample_kernel<<<grid, 256, 0, s>>>(in, out, l_in, l_out, hw);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

}  // namespace clasfv
