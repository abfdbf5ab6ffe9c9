// decoder_umma.cu - the fused decoder head of decoder.cu with its 64x64 convolution on the tensor cores.
//
// One CTA (128 threads) produces one output row (n, t, h) of up to 128 voxels:
//   phase 1  the T- and H-interpolated rows of the four laterally projected maps (bf16 in HBM) are
//            built in shared memory as fp32 (trilinear, align_corners=True, reference
//            src/model/R2plus1D_18_MotionNet.py:41-49); corners with zero weight are not read
//   phase 2  thread (voxel group of 4, channel quarter) W-interpolates and sums the four levels, adds
//            the folded comb_1 bias, ReLU, converts to bf16 and writes its part of the 128 x 64 A tile
//            directly in the K-major SWIZZLE_128B layout tcgen05 consumes (adjacent voxels share their
//            low-resolution taps, so each tap vector is read from shared memory once per group)
//   MMA 1    one elected thread issues 4 x tcgen05.mma (M=128, N=64, K=16): D = A * W2^T into TMEM
//   mid      thread = voxel: tcgen05.ld its 64 accumulators, + folded comb_2 bias, ReLU, bf16, back into the
//            (now free) A tile
//   MMA 2    4 x tcgen05.mma (M=128, N=16, K=16): the 6x64 segmentation + motion heads (10 zero rows)
//   epilogue thread = voxel: tcgen05.ld 8 accumulators, + head bias, softmax / tanh, six coalesced planar stores
// The kernel is instruction-bound, not memory-bound: the interpolated rows are kept in bf16 and the heads run on
// the tensor core to cut CUDA-core instructions per row and to fit four CTAs per SM.
// Nothing between the lateral projections and the six output planes touches HBM.
#include "internal.h"
#include "umma_ptx.cuh"

namespace clasfv {
namespace {

using namespace ptx;

constexpr int HU_THREADS = 128;
constexpr int HC = 64;
constexpr int ROW_PITCH = HC + 8;       // bf16 per low-res column: 144 B keeps 16-byte alignment and spreads banks

struct AxisTap { int i0, i1; float l0, l1; };
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

__device__ __forceinline__ float4 ld_bf16x4(const __nv_bfloat16* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ uint32_t pack_relu_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(fmaxf(a, 0.f), fmaxf(b, 0.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <typename OutT> __device__ __forceinline__ void put(OutT* p, float v);
template <> __device__ __forceinline__ void put<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void put<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct HeadSmem {
  // byte offsets from the 1024-aligned base
  static constexpr uint32_t A = 0;                 // 128 x 128 B : relu(h1) for MMA 1, then relu(h2) for MMA 2
  static constexpr uint32_t B = 16384;             // 64 x 128 B  : W2
  static constexpr uint32_t B3 = 24576;            // 16 x 128 B  : heads (rows 0-1 seg, 2-5 motion, 6-15 zero)
  static constexpr uint32_t B1 = B3 + 2048;        // [64] fp32
  static constexpr uint32_t B2 = B1 + 256;         // [64]
  static constexpr uint32_t BH = B2 + 256;         // [8]
  static constexpr uint32_t BAR = BH + 32;         // mbarrier MMA 1
  static constexpr uint32_t BAR3 = BAR + 8;        // mbarrier MMA 2
  static constexpr uint32_t TMEM = BAR3 + 8;       // tmem base
  static constexpr uint32_t ROWS = 27648;          // 4 levels x [wl][ROW_PITCH] bf16
};

__device__ __forceinline__ void bf16x8_to_f32(const uint4& r, float (&f)[8]) {
  f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
  f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
}

template <typename OutT>
__global__ void __launch_bounds__(HU_THREADS, 4) head_umma_kernel(const HeadArgs a, int w_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  float* b1s = reinterpret_cast<float*>(sm + HeadSmem::B1);
  float* b2s = reinterpret_cast<float*>(sm + HeadSmem::B2);
  float* bhs = reinterpret_cast<float*>(sm + HeadSmem::BH);
  __nv_bfloat16* rows = reinterpret_cast<__nv_bfloat16*>(sm + HeadSmem::ROWS);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + HeadSmem::TMEM);
  const uint32_t bar = sbase + HeadSmem::BAR, bar3 = sbase + HeadSmem::BAR3;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x / w_tiles, w_base = (blockIdx.x % w_tiles) * 128;
  const int t = blockIdx.y, n = blockIdx.z;

  // ---- one-time setup
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + HeadSmem::TMEM), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar3, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // W2 (64 out x 64 in, bf16) into the K-major swizzled B tile: 512 chunks of 16 bytes
  for (int i = tid; i < 512; i += HU_THREADS) {
    const int row = i >> 3, chunk = i & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.w2_bf16 + row * HC) + chunk);
    *reinterpret_cast<uint4*>(sm + HeadSmem::B + sw128_offset(row, chunk)) = v;
  }
  // heads (6 x 64 fp32 -> bf16) into the K-major swizzled B3 tile, rows 6..15 zero: 128 chunks of 16 bytes
  {
    const int row = tid >> 3, chunk = tid & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < 6) {
      const float* src = a.wh + row * HC + chunk * 8;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(__ldg(src + 0), __ldg(src + 1)), h1 = __floats2bfloat162_rn(__ldg(src + 2), __ldg(src + 3));
      __nv_bfloat162 h2 = __floats2bfloat162_rn(__ldg(src + 4), __ldg(src + 5)), h3 = __floats2bfloat162_rn(__ldg(src + 6), __ldg(src + 7));
      v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
      v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
    }
    *reinterpret_cast<uint4*>(sm + HeadSmem::B3 + sw128_offset(row, chunk)) = v;
  }
  if (tid < HC) { b1s[tid] = __ldg(a.b1 + tid); b2s[tid] = __ldg(a.b2 + tid); }
  if (tid < 6) bhs[tid] = __ldg(a.bh + tid);

  // ---- phase 1: T/H-interpolated rows, rows[off_l + x*ROW_PITCH + c]
  int row_off[4];
  {
    int off = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) { row_off[l] = off; off += a.wl[l] * ROW_PITCH; }
  }
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const AxisTap at = axis_tap(t, a.tl[l], a.t), ah = axis_tap(h, a.hl[l], a.h);
    const __nv_bfloat16* __restrict__ g = static_cast<const __nv_bfloat16*>(a.g[l]) + (int64_t)n * a.tl[l] * a.hl[l] * a.wl[l] * HC;
    const int64_t r00 = ((int64_t)at.i0 * a.hl[l] + ah.i0) * a.wl[l], r01 = ((int64_t)at.i0 * a.hl[l] + ah.i1) * a.wl[l];
    const int64_t r10 = ((int64_t)at.i1 * a.hl[l] + ah.i0) * a.wl[l], r11 = ((int64_t)at.i1 * a.hl[l] + ah.i1) * a.wl[l];
    const float w00 = at.l0 * ah.l0, w01 = at.l0 * ah.l1, w10 = at.l1 * ah.l0, w11 = at.l1 * ah.l1;
    __nv_bfloat16* dst = rows + row_off[l];
    const int total = a.wl[l] * (HC / 4);
    for (int i = tid; i < total; i += HU_THREADS) {
      const int x = i / (HC / 4), c4 = i % (HC / 4);
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (w00 != 0.f) fma4(o, w00, ld_bf16x4(g + (r00 + x) * HC + 4 * c4));
      if (w01 != 0.f) fma4(o, w01, ld_bf16x4(g + (r01 + x) * HC + 4 * c4));
      if (w10 != 0.f) fma4(o, w10, ld_bf16x4(g + (r10 + x) * HC + 4 * c4));
      if (w11 != 0.f) fma4(o, w11, ld_bf16x4(g + (r11 + x) * HC + 4 * c4));
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(o.x, o.y), p1 = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk; pk.x = *reinterpret_cast<const uint32_t*>(&p0); pk.y = *reinterpret_cast<const uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(dst + x * ROW_PITCH + 4 * c4) = pk;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- phase 2: A tile.  thread = (voxel group vg of 4 voxels, channel quarter cq of 16 channels)
  {
    const int vg = tid >> 2, cq = tid & 3;
    float4 f[4][4];                         // [voxel][4 x float4 = 16 channels]
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int q = 0; q < 4; ++q) f[v][q] = *reinterpret_cast<const float4*>(b1s + cq * 16 + 4 * q);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const __nv_bfloat16* r = rows + row_off[l] + cq * 16;
      int cached = -1;
      float4 tv[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int w = min(w_base + vg * 4 + v, a.w - 1);
        const AxisTap aw = axis_tap(w, a.wl[l], a.w);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const int x = side ? aw.i1 : aw.i0;
          const float wt = side ? aw.l1 : aw.l0;
          if (wt == 0.f) continue;
          if (x != cached) {
            const uint4 lo = *reinterpret_cast<const uint4*>(r + x * ROW_PITCH), hi = *reinterpret_cast<const uint4*>(r + x * ROW_PITCH + 8);
            float f8[8];
            bf16x8_to_f32(lo, f8);
            tv[0] = make_float4(f8[0], f8[1], f8[2], f8[3]); tv[1] = make_float4(f8[4], f8[5], f8[6], f8[7]);
            bf16x8_to_f32(hi, f8);
            tv[2] = make_float4(f8[0], f8[1], f8[2], f8[3]); tv[3] = make_float4(f8[4], f8[5], f8[6], f8[7]);
            cached = x;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) fma4(f[v][q], wt, tv[q]);
        }
      }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint32_t row = (uint32_t)(vg * 4 + v);
      uint4 c0, c1;
      c0.x = pack_relu_bf16(f[v][0].x, f[v][0].y); c0.y = pack_relu_bf16(f[v][0].z, f[v][0].w);
      c0.z = pack_relu_bf16(f[v][1].x, f[v][1].y); c0.w = pack_relu_bf16(f[v][1].z, f[v][1].w);
      c1.x = pack_relu_bf16(f[v][2].x, f[v][2].y); c1.y = pack_relu_bf16(f[v][2].z, f[v][2].w);
      c1.z = pack_relu_bf16(f[v][3].x, f[v][3].y); c1.w = pack_relu_bf16(f[v][3].z, f[v][3].w);
      *reinterpret_cast<uint4*>(sm + HeadSmem::A + sw128_offset(row, (uint32_t)(2 * cq))) = c0;
      *reinterpret_cast<uint4*>(sm + HeadSmem::A + sw128_offset(row, (uint32_t)(2 * cq + 1))) = c1;
    }
  }
  fence_async_smem();          // generic-proxy writes of A and B -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();

  // ---- MMA 1: D[128 x 64] = A[128 x 64] * W2[64 x 64]^T
  if (tid == 0) {
    tc_fence_after();
    const uint64_t da = smem_desc_sw128(sbase + HeadSmem::A), db = smem_desc_sw128(sbase + HeadSmem::B);
    const uint32_t idesc = idesc_bf16_f32(128, 64);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
    tc_commit(bar);
  }

  // ---- mid: thread = voxel (TMEM lane = A row): h2 = relu(D + b2) -> bf16 -> A tile (MMA 1 has finished reading it)
  mbar_wait(bar, 0);
  tc_fence_after();
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  {
    uint32_t acc[2][16];
    tc_ld16(taddr, acc[0]);
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      tc_wait_ld();
      if (ci < 3) tc_ld16(taddr + (uint32_t)(16 * (ci + 1)), acc[(ci + 1) & 1]);
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        pk[j] = pack_relu_bf16(__uint_as_float(acc[ci & 1][2 * j]) + b2s[16 * ci + 2 * j], __uint_as_float(acc[ci & 1][2 * j + 1]) + b2s[16 * ci + 2 * j + 1]);
      *reinterpret_cast<uint4*>(sm + HeadSmem::A + sw128_offset((uint32_t)tid, (uint32_t)(2 * ci))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(sm + HeadSmem::A + sw128_offset((uint32_t)tid, (uint32_t)(2 * ci + 1))) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();            // every thread has drained its accumulators: the same TMEM columns take the head outputs

  // ---- MMA 2: D3[128 x 16] = relu(h2)[128 x 64] * Wh[16 x 64]^T
  if (tid == 0) {
    tc_fence_after();
    const uint64_t da = smem_desc_sw128(sbase + HeadSmem::A), db = smem_desc_sw128(sbase + HeadSmem::B3);
    const uint32_t idesc = idesc_bf16_f32(128, 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
    tc_commit(bar3);
  }

  // ---- epilogue
  mbar_wait(bar3, 0);
  tc_fence_after();
  {
    uint32_t r8[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r8[0]), "=r"(r8[1]), "=r"(r8[2]), "=r"(r8[3]), "=r"(r8[4]), "=r"(r8[5]), "=r"(r8[6]), "=r"(r8[7]) : "r"(taddr));
    tc_wait_ld();
    float o[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = __uint_as_float(r8[k]) + bhs[k];
    const int w = w_base + tid;
    if (w < a.w) {
      float s0 = o[0], s1 = o[1];
      if (a.out_kind == CLASFV_OUT_PROB) {
        const float mx = fmaxf(s0, s1);
        const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
        const float inv = 1.f / (e0 + e1);
        s0 = e0 * inv; s1 = e1 * inv;
      }
      const int64_t plane = (int64_t)a.h * a.w;
      const int64_t pix = (int64_t)h * a.w + w;
      OutT* seg = static_cast<OutT*>(a.seg) + ((int64_t)n * 2 * a.t + t) * plane + pix;
      put<OutT>(seg, s0);
      put<OutT>(seg + (int64_t)a.t * plane, s1);
      OutT* mot = static_cast<OutT*>(a.motion) + ((int64_t)n * 4 * a.t + t) * plane + pix;
#pragma unroll
      for (int k = 0; k < 4; ++k) put<OutT>(mot + (int64_t)k * a.t * plane, tanhf(o[2 + k]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

}  // namespace

int launch_head_umma(const HeadArgs& a, cudaStream_t stream) {
  CLASFV_REQUIRE(a.g_dtype == CLASFV_BF16 && a.w2_bf16, "head_umma: bf16 lateral maps and bf16 W2 required");
  int rowbuf = 0;
  for (int l = 0; l < 4; ++l) rowbuf += a.wl[l] * ROW_PITCH;
  const size_t smem = 1024 + HeadSmem::ROWS + (size_t)rowbuf * sizeof(__nv_bfloat16);
  CLASFV_REQUIRE(smem <= 200 * 1024, "head_umma: frame too wide for the row buffers (W=%d)", a.w);
  const int w_tiles = (a.w + 127) / 128;
  dim3 grid((unsigned)(a.h * w_tiles), (unsigned)a.t, (unsigned)a.n);
  if (a.out_dtype == CLASFV_F32) {
    CLASFV_CUDA(cudaFuncSetAttribute(head_umma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_umma_kernel<float><<<grid, HU_THREADS, smem, stream>>>(a, w_tiles);
  } else {
    CLASFV_CUDA(cudaFuncSetAttribute(head_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_umma_kernel<__nv_bfloat16><<<grid, HU_THREADS, smem, stream>>>(a, w_tiles);
  }
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

}  // namespace clasfv
