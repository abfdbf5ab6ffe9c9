// conv_simt.cu - CUDA-core convolution kernels on channels-last activations.
//
//  * conv_ndhwc_simt: generic Conv3d as an implicit GEMM (rows = output positions, cols = output
//    channels, K = taps x input channels), fp32 FFMA with fp32 accumulation.  It is the arithmetic of
//    the CLASFV_F32 mode (softmax tolerance 1e-4 against the reference needs true fp32 products, which
//    tcgen05 does not offer) and, instantiated on bf16 storage, the A/B partner the tcgen05 kernel is
//    checked against (same rounding points, different execution unit).
//  * stem_conv_kernel: the 1x7x7 stride (1,2,2) stem convolution (torchvision R2Plus1dStem, the trunk
//    the reference builds at src/model/R2plus1D_18_MotionNet.py:13,29).  Cin = 3 is useless as a GEMM K
//    dimension, so the stem reads the planar fp32 clip directly (no layout pass) and emits
//    channels-last activations for everything downstream.
#include "internal.h"
#include "umma_ptx.cuh"

#include <algorithm>

namespace clasfv {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, SIMT_THREADS = 256;

template <typename T> struct Vec8;  // 8 consecutive channels
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
template <> struct Vec8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
};

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, const float (&v)[4]) {
  uint2 w;
  w.x = ptx::cvt_f16x2_sat(v[0], v[1]); w.y = ptx::cvt_f16x2_sat(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = w;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 r; r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

struct SimtParams {
  ConvShape s;
  const void* in; const void* in2; const void* weight; const float* bias; const void* residual; void* out;
  int64_t in_batch_stride;
  int relu; int m_total;
};

template <typename InT, typename OutT>
__global__ void __launch_bounds__(SIMT_THREADS) conv_ndhwc_simt(const SimtParams p) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const ConvShape& s = p.s;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const InT* __restrict__ in = static_cast<const InT*>(p.in);
  const InT* __restrict__ wt = static_cast<const InT*>(p.weight);

  // A-operand gather role: one output position (row) and one 8-channel half of the K chunk
  const int a_row = tid & (BM - 1), a_half = tid >> 7;
  const int m = m0 + a_row;
  const bool row_ok = m < p.m_total;
  int ti0 = 0, hi0 = 0, wi0 = 0;
  int64_t in_n = 0;
  if (row_ok) {
    int r = m;
    const int wo = r % s.wo; r /= s.wo;
    const int ho = r % s.ho; r /= s.ho;
    const int to = r % s.to; r /= s.to;
    ti0 = to * s.st - s.pt; hi0 = ho * s.sh - s.ph; wi0 = wo * s.sw - s.pw;
    in_n = (int64_t)r * p.in_batch_stride;      // in positions (elements / cin)
  }
  // B-operand role (threads 0..127): output channel row and 8-channel half
  const int b_row = tid & (BN - 1), b_half = (tid >> 6) & 1;
  const bool b_active = tid < 2 * BN;
  const bool b_ok = b_active && (n0 + b_row) < s.cout;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int kchunks = s.cin / BK;
  const InT* __restrict__ in2 = static_cast<const InT*>(p.in2);
  const int ntaps = s.kt * s.kh * s.kw * (in2 ? 2 : 1);   // two-source 1x1x1: tap 1 reads the second tensor
  const int ksteps = ntaps * kchunks;

  Vec8<InT> ra, rb;
  auto fetch = [&](int ks) {
    const int tap = ks / kchunks, c0 = (ks - tap * kchunks) * BK;
    const InT* src = in;
    int sp = tap;
    if (in2 && tap == 1) { src = in2; sp = 0; }
    const int dw = sp % s.kw, dh = (sp / s.kw) % s.kh, dt = sp / (s.kw * s.kh);
    const int ti = ti0 + dt, hi = hi0 + dh, wi = wi0 + dw;
    const bool ok = row_ok && (unsigned)ti < (unsigned)s.ti && (unsigned)hi < (unsigned)s.hi && (unsigned)wi < (unsigned)s.wi;
    if (ok) ra.load(src + ((in_n + ((int64_t)ti * s.hi + hi) * s.wi + wi) * s.cin + c0 + a_half * 8));
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) ra.v[i] = 0.f;
    }
    if (b_ok) rb.load(wt + (((int64_t)tap * s.cout + n0 + b_row) * s.cin + c0 + b_half * 8));
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) rb.v[i] = 0.f;
    }
  };

  fetch(0);
  for (int ks = 0; ks < ksteps; ++ks) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) As[a_half * 8 + i][a_row] = ra.v[i];
    if (b_active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) Bs[b_half * 8 + i][b_row] = rb.v[i];
    }
    __syncthreads();
    if (ks + 1 < ksteps) fetch(ks + 1);
    // Blocked summation: the 16 products of this K chunk are summed on their own and added to the running
    // total once, so the rounding error grows with K/16 + 16 instead of K (K reaches 10 368 in layer4).
    float part[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
  }

  const int col = n0 + tx * 4;
  if (col >= s.cout) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) load4<float>(p.bias + col, bias);
  OutT* __restrict__ out = static_cast<OutT*>(p.out);
  const OutT* __restrict__ res = static_cast<const OutT*>(p.residual);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty * 8 + i;
    if (row >= p.m_total) break;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    const int64_t off = (int64_t)row * s.cout + col;
    if (res) {
      float r[4]; load4<OutT>(res + off, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += r[j];
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    store4<OutT>(out + off, v);
  }
}

// ------------------------------------------------------------------------------------------ stem
constexpr int STEM_TH = 8, STEM_TW = 16, STEM_CO = 48, STEM_TAPS = 147;
constexpr int STEM_IH = 2 * STEM_TH + 5, STEM_IW = 2 * STEM_TW + 5;

template <typename OutT>
__global__ void __launch_bounds__(STEM_TH * STEM_TW) stem_conv_kernel(const StemArgs a) {
  __shared__ __align__(16) float wsm[STEM_TAPS * STEM_CO];
  __shared__ float tile[3][STEM_IH][STEM_IW];
  const int tid = threadIdx.x;
  const int ho_n = a.h / 2, wo_n = a.w / 2;
  const int nt = blockIdx.z, n = nt / a.t, t = nt % a.t;
  const int oh0 = blockIdx.y * STEM_TH, ow0 = blockIdx.x * STEM_TW;

  for (int i = tid; i < STEM_TAPS * STEM_CO / 4; i += blockDim.x)
    reinterpret_cast<float4*>(wsm)[i] = __ldg(reinterpret_cast<const float4*>(a.weight) + i);
  const float* __restrict__ base = a.x + a.clip_offset[n] + (int64_t)t * a.h * a.w;
  const int ih0 = 2 * oh0 - 3, iw0 = 2 * ow0 - 3;
  for (int i = tid; i < 3 * STEM_IH * STEM_IW; i += blockDim.x) {
    const int c = i / (STEM_IH * STEM_IW), r = i % (STEM_IH * STEM_IW);
    const int y = r / STEM_IW, x = r % STEM_IW;
    const int ih = ih0 + y, iw = iw0 + x;
    float v = 0.f;
    if ((unsigned)ih < (unsigned)a.h && (unsigned)iw < (unsigned)a.w) v = __ldg(base + c * a.channel_stride + (int64_t)ih * a.w + iw);
    tile[c][y][x] = v;
  }
  __syncthreads();

  const int ohl = tid / STEM_TW, owl = tid % STEM_TW;
  float acc[STEM_CO];
#pragma unroll
  for (int i = 0; i < STEM_CO; ++i) acc[i] = 0.f;
  for (int c = 0; c < 3; ++c)
    for (int kh = 0; kh < 7; ++kh) {
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const float v = tile[c][2 * ohl + kh][2 * owl + kw];
        const float4* wr = reinterpret_cast<const float4*>(wsm + ((c * 7 + kh) * 7 + kw) * STEM_CO);
#pragma unroll
        for (int q = 0; q < STEM_CO / 4; ++q) {
          const float4 w4 = wr[q];
          acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
        }
      }
    }
  const int oh = oh0 + ohl, ow = ow0 + owl;
  if (oh >= ho_n || ow >= wo_n) return;
  OutT* out = static_cast<OutT*>(a.out) + ((((int64_t)n * a.t + t) * ho_n + oh) * wo_n + ow) * a.out_channels;
#pragma unroll
  for (int q = 0; q < STEM_CO / 4; ++q) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(acc[4 * q + j] + __ldg(a.bias + 4 * q + j), 0.f);
    store4<OutT>(out + 4 * q, v);
  }
  const float zero[4] = {0.f, 0.f, 0.f, 0.f};   // channel padding of the stored activation
  for (int q = STEM_CO / 4; q < a.out_channels / 4; ++q) store4<OutT>(out + 4 * q, zero);
}

}  // namespace

int launch_conv_simt(const ConvArgs& a, cudaStream_t stream) {
  const ConvShape& s = a.s;
  CLASFV_REQUIRE(s.cin % BK == 0 && s.cout % 4 == 0, "conv_simt: cin %% 16 and cout %% 4 required (cin=%d cout=%d)", s.cin, s.cout);
  CLASFV_REQUIRE(!a.seg.on && !a.res.on && !a.out_batch_stride, "conv_simt: ragged clip geometry (dense-video trunk) is a tcgen05-path feature");
  CLASFV_REQUIRE(a.act_dtype != CLASFV_F16 && !a.out_f16, "conv_simt: fp16 storage is a tcgen05-path feature");
  SimtParams p;
  p.s = s; p.in = a.in; p.in2 = a.in2; p.weight = a.weight; p.bias = a.bias; p.residual = a.residual; p.out = a.out; p.relu = a.relu;
  CLASFV_REQUIRE(!a.in2 || (s.kt * s.kh * s.kw == 1 && s.st == 1 && s.sh == 1 && s.sw == 1), "conv_simt: two-source mode is 1x1x1 only");
  const int64_t dense = (int64_t)s.ti * s.hi * s.wi;
  CLASFV_REQUIRE(a.in_batch_stride % s.cin == 0, "conv_simt: batch stride must be a whole number of positions");
  p.in_batch_stride = a.in_batch_stride ? a.in_batch_stride / s.cin : dense;
  const int64_t m_total = (int64_t)s.n * s.to * s.ho * s.wo;
  CLASFV_REQUIRE(m_total > 0 && m_total < (1ll << 31), "conv_simt: bad row count");
  p.m_total = (int)m_total;
  dim3 grid((unsigned)cdiv(m_total, BM), (unsigned)cdiv(s.cout, BN));
  if (a.act_dtype == CLASFV_F32) {
    conv_ndhwc_simt<float, float><<<grid, SIMT_THREADS, 0, stream>>>(p);
  } else if (a.out_f32) {
    conv_ndhwc_simt<__nv_bfloat16, float><<<grid, SIMT_THREADS, 0, stream>>>(p);
  } else {
    conv_ndhwc_simt<__nv_bfloat16, __nv_bfloat16><<<grid, SIMT_THREADS, 0, stream>>>(p);
  }
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

int launch_stem(const StemArgs& a, cudaStream_t stream) {
  CLASFV_REQUIRE(a.out_channels >= STEM_CO && a.out_channels % 4 == 0, "stem: bad output channel count %d", a.out_channels);
  dim3 grid((unsigned)cdiv(a.w / 2, STEM_TW), (unsigned)cdiv(a.h / 2, STEM_TH), (unsigned)(a.n * a.t));
  if (a.out_dtype == CLASFV_F32) stem_conv_kernel<float><<<grid, STEM_TH * STEM_TW, 0, stream>>>(a);
  else if (a.out_dtype == CLASFV_F16) stem_conv_kernel<__half><<<grid, STEM_TH * STEM_TW, 0, stream>>>(a);
  else stem_conv_kernel<__nv_bfloat16><<<grid, STEM_TH * STEM_TW, 0, stream>>>(a);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

}  // namespace clasfv
