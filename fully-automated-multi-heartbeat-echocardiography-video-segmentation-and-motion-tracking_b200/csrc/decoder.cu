// decoder.cu - the CLAS-FV decoder after commuting comb_1_layer with the up-sampling.
//
// Reference (src/model/R2plus1D_18_MotionNet.py:41-69): five trilinear (align_corners=True) up-samplings
// to full resolution, a 1024-channel concat (1.64 GB fp32 per clip), 1x1x1 conv 1024->64 + BN + ReLU,
// 1x1x1 conv 64->64 + BN + ReLU, then the 2-channel segmentation head and the 4-channel tanh motion head.
// A 1x1x1 convolution and a linear interpolation commute, so the 1024->64 projection is applied to each
// feature map at its native resolution (api.cu runs those as ordinary convolutions; 0.96 GMAC instead of
// 26.3 GMAC per clip) and this kernel does the rest in one pass per output row:
//   phase 1  the T- and H-interpolated rows of the four projected maps are built in shared memory
//            (interpolation is separable: 4 corners per low-res column instead of 8 per output voxel),
//   phase 2  one thread per output voxel: W-interpolate and sum the four levels, + bias, ReLU,
//            64x64 (comb_2 with BN folded), ReLU, the 6x64 heads, softmax / tanh, and six planar stores.
// Nothing between the lateral projections and the six output planes touches HBM.
#include "internal.h"

namespace clasfv {
namespace {

constexpr int HEAD_THREADS = 128;
constexpr int HC = 64;        // decoder width

struct AxisTap { int i0, i1; float l0, l1; };

// PyTorch upsample, align_corners=True: src = dst * (in-1)/(out-1)
__device__ __forceinline__ AxisTap axis_tap(int dst, int in_size, int out_size) {
  AxisTap a;
  const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  const float src = scale * (float)dst;
  a.i0 = min((int)src, in_size - 1);
  a.i1 = a.i0 + (a.i0 < in_size - 1 ? 1 : 0);
  a.l1 = src - (float)a.i0;
  a.l0 = 1.f - a.l1;
  return a;
}

template <typename OutT> __device__ __forceinline__ void put(OutT* p, float v);
template <> __device__ __forceinline__ void put<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void put<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ void put<__half>(__half* p, float v) { *p = __float2half_rn(v); }

template <typename OutT>
__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const HeadArgs a, int rowbuf_floats) {
  extern __shared__ __align__(16) float smem[];
  float* w2s = smem;                       // [64][64]
  float* whs = w2s + HC * HC;              // [6][64]
  float* b1s = whs + 6 * HC;               // [64]
  float* b2s = b1s + HC;                   // [64]
  float* bhs = b2s + HC;                   // [8]
  float* rows = bhs + 8;                   // 4 levels x [64][wl_pad]
  const int tid = threadIdx.x;
  const int h = blockIdx.x, t = blockIdx.y, n = blockIdx.z;

  for (int i = tid; i < HC * HC / 4; i += HEAD_THREADS) reinterpret_cast<float4*>(w2s)[i] = __ldg(reinterpret_cast<const float4*>(a.w2) + i);
  for (int i = tid; i < 6 * HC; i += HEAD_THREADS) whs[i] = __ldg(a.wh + i);
  if (tid < HC) { b1s[tid] = __ldg(a.b1 + tid); b2s[tid] = __ldg(a.b2 + tid); }
  if (tid < 6) bhs[tid] = __ldg(a.bh + tid);

  // phase 1: rows[l][c][x] = sum over the (t,h) corners of level l
  int row_off[4], wl_pad[4];
  {
    int off = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) { wl_pad[l] = a.wl[l] | 1; row_off[l] = off; off += HC * wl_pad[l]; }
  }
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const AxisTap at = axis_tap(t, a.tl[l], a.t), ah = axis_tap(h, a.hl[l], a.h);
    const float* __restrict__ g = static_cast<const float*>(a.g[l]) + (int64_t)n * a.tl[l] * a.hl[l] * a.wl[l] * HC;
    const int64_t r00 = ((int64_t)at.i0 * a.hl[l] + ah.i0) * a.wl[l], r01 = ((int64_t)at.i0 * a.hl[l] + ah.i1) * a.wl[l];
    const int64_t r10 = ((int64_t)at.i1 * a.hl[l] + ah.i0) * a.wl[l], r11 = ((int64_t)at.i1 * a.hl[l] + ah.i1) * a.wl[l];
    const float w00 = at.l0 * ah.l0, w01 = at.l0 * ah.l1, w10 = at.l1 * ah.l0, w11 = at.l1 * ah.l1;
    float* dst = rows + row_off[l];
    const int total = a.wl[l] * (HC / 4);
    for (int i = tid; i < total; i += HEAD_THREADS) {
      const int x = i / (HC / 4), c4 = i % (HC / 4);
      const float4 v00 = __ldg(reinterpret_cast<const float4*>(g + (r00 + x) * HC) + c4);
      const float4 v01 = __ldg(reinterpret_cast<const float4*>(g + (r01 + x) * HC) + c4);
      const float4 v10 = __ldg(reinterpret_cast<const float4*>(g + (r10 + x) * HC) + c4);
      const float4 v11 = __ldg(reinterpret_cast<const float4*>(g + (r11 + x) * HC) + c4);
      // same nesting as the reference's trilinear kernel: t outside, h inside
      const float o0 = at.l0 * (ah.l0 * v00.x + ah.l1 * v01.x) + at.l1 * (ah.l0 * v10.x + ah.l1 * v11.x);
      const float o1 = at.l0 * (ah.l0 * v00.y + ah.l1 * v01.y) + at.l1 * (ah.l0 * v10.y + ah.l1 * v11.y);
      const float o2 = at.l0 * (ah.l0 * v00.z + ah.l1 * v01.z) + at.l1 * (ah.l0 * v10.z + ah.l1 * v11.z);
      const float o3 = at.l0 * (ah.l0 * v00.w + ah.l1 * v01.w) + at.l1 * (ah.l0 * v10.w + ah.l1 * v11.w);
      (void)w00; (void)w01; (void)w10; (void)w11;
      dst[(4 * c4 + 0) * wl_pad[l] + x] = o0;
      dst[(4 * c4 + 1) * wl_pad[l] + x] = o1;
      dst[(4 * c4 + 2) * wl_pad[l] + x] = o2;
      dst[(4 * c4 + 3) * wl_pad[l] + x] = o3;
    }
  }
  __syncthreads();

  // phase 2: one output voxel per thread
  const int64_t plane = (int64_t)a.h * a.w;
  for (int w = tid; w < a.w; w += HEAD_THREADS) {
    float f[HC];
#pragma unroll
    for (int c = 0; c < HC; ++c) f[c] = b1s[c];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const AxisTap aw = axis_tap(w, a.wl[l], a.w);
      const float* r = rows + row_off[l];
      const int pad = wl_pad[l];
#pragma unroll
      for (int c = 0; c < HC; ++c) f[c] += aw.l0 * r[c * pad + aw.i0] + aw.l1 * r[c * pad + aw.i1];
    }
#pragma unroll
    for (int c = 0; c < HC; ++c) f[c] = fmaxf(f[c], 0.f);
    float o[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = bhs[k];
    for (int j = 0; j < HC; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(w2s + j * HC);
      float s0 = b2s[j], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int q = 0; q < HC / 4; ++q) {
        const float4 w4 = wr[q];
        s0 = fmaf(w4.x, f[4 * q + 0], s0); s1 = fmaf(w4.y, f[4 * q + 1], s1);
        s2 = fmaf(w4.z, f[4 * q + 2], s2); s3 = fmaf(w4.w, f[4 * q + 3], s3);
      }
      const float hj = fmaxf((s0 + s1) + (s2 + s3), 0.f);
#pragma unroll
      for (int k = 0; k < 6; ++k) o[k] = fmaf(whs[k * HC + j], hj, o[k]);
    }
    float s0 = o[0], s1 = o[1];
    const bool lv_only = a.out_kind == CLASFV_OUT_LVPROB;
    if (a.out_kind == CLASFV_OUT_PROB || lv_only) {
      const float mx = fmaxf(s0, s1);
      const float e0 = expf(s0 - mx), e1 = expf(s1 - mx);
      const float inv = 1.f / (e0 + e1);
      s0 = e0 * inv; s1 = e1 * inv;
    }
    const int64_t pix = (int64_t)h * a.w + w;
    OutT* seg = static_cast<OutT*>(a.seg) + ((int64_t)n * (lv_only ? 1 : 2) * a.t + t) * plane + pix;
    if (lv_only) {
      put<OutT>(seg, s1);
    } else {
      put<OutT>(seg, s0);
      put<OutT>(seg + (int64_t)a.t * plane, s1);
    }
    OutT* mot = static_cast<OutT*>(a.motion) + ((int64_t)n * 4 * a.t + t) * plane + pix;
#pragma unroll
    for (int k = 0; k < 4; ++k) put<OutT>(mot + (int64_t)k * a.t * plane, tanhf(o[2 + k]));
  }
}

}  // namespace

int launch_head(const HeadArgs& a, cudaStream_t stream) {
  CLASFV_REQUIRE(a.g_dtype == CLASFV_F32, "head: the CUDA-core head reads fp32 lateral maps");
  int rowbuf = 0;
  for (int l = 0; l < 4; ++l) rowbuf += HC * (a.wl[l] | 1);
  const size_t smem = (size_t)(HC * HC + 6 * HC + 2 * HC + 8 + rowbuf) * sizeof(float);
  CLASFV_REQUIRE(smem <= 200 * 1024, "head: frame too wide for the row buffers (W=%d)", a.w);
  dim3 grid((unsigned)a.h, (unsigned)a.t, (unsigned)a.n);
  if (a.out_dtype == CLASFV_F32) {
    CLASFV_CUDA(allow_max_dynamic_smem(head_kernel<float>));
    head_kernel<float><<<grid, HEAD_THREADS, smem, stream>>>(a, rowbuf);
  } else if (a.out_dtype == CLASFV_F16) {
    CLASFV_CUDA(allow_max_dynamic_smem(head_kernel<__half>));
    head_kernel<__half><<<grid, HEAD_THREADS, smem, stream>>>(a, rowbuf);
  } else {
    CLASFV_CUDA(allow_max_dynamic_smem(head_kernel<__nv_bfloat16>));
    head_kernel<__nv_bfloat16><<<grid, HEAD_THREADS, smem, stream>>>(a, rowbuf);
  }
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

}  // namespace clasfv
