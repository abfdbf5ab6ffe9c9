// decoder_umma.cu - the fused decoder head of decoder.cu with its 64x64 convolution on the tensor cores.
//
// A persistent CTA (512 threads) produces one output row (n, t, h) of up to 128 voxels at a time:
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

#include <algorithm>

namespace clasfv {
namespace {

using namespace ptx;

constexpr int HU_THREADS = 512;
constexpr int HC = 64;
constexpr int ROW_PITCH = HC + 4;       // floats per low-res column: 272 B keeps 16-byte alignment and spreads banks

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
  static constexpr uint32_t BAR_RAW = BH + 32;     // mbarrier: raw corner rows landed
  static constexpr uint32_t BAR = BAR_RAW + 8;     // mbarrier MMA 1
  static constexpr uint32_t BAR3 = BAR + 8;        // mbarrier MMA 2
  static constexpr uint32_t TMEM = BAR3 + 8;       // tmem base
  static constexpr uint32_t PLAN = 27264;          // 2 x TilePlan (double-buffered: written one row ahead by thread 0)
  static constexpr uint32_t WTAP = 27648;          // [4 levels][128 voxels] WTap: the W-axis taps, computed once per CTA
  static constexpr uint32_t ROWS = WTAP + 4 * 128 * 16;   // 4 levels x [wl][ROW_PITCH] fp32, then the raw corner rows (bf16)
};

struct __align__(16) WTap { int i0, i1; float l0, l1; };

__device__ __forceinline__ void bf16x4_to_f32(const uint2& r, float (&f)[4]) {
  f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
}

struct TilePlan {                 // one output row (n, t, h): which corner rows feed it and with what weight
  int n, t, h, w_base;
  float wgt[4][4];                // [level][corner = 2*tc + hc]
  int ti[4][2], hi[4][2];         // corner indices per level
};

// Persistent, 512 threads: every CTA loops over output rows (n, t, h).  The kernel is latency-bound, not
// bandwidth-bound (few thousand instructions per row, each phase a dependent chain), so a row's work is spread
// over 16 warps and the raw bf16 corner rows of the NEXT row are fetched with 1-D bulk copies
// (cp.async.bulk -> mbarrier) while the current row is interpolated, multiplied and written out.
template <typename OutT>
__global__ void __launch_bounds__(HU_THREADS, 2) head_umma_kernel(const HeadArgs a, int w_tiles, int total_tiles, int rows_elems) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  float* b1s = reinterpret_cast<float*>(sm + HeadSmem::B1);
  float* b2s = reinterpret_cast<float*>(sm + HeadSmem::B2);
  float* bhs = reinterpret_cast<float*>(sm + HeadSmem::BH);
  TilePlan* plans = reinterpret_cast<TilePlan*>(sm + HeadSmem::PLAN);
  WTap* wtap = reinterpret_cast<WTap*>(sm + HeadSmem::WTAP);
  float* rows = reinterpret_cast<float*>(sm + HeadSmem::ROWS);
  const __nv_bfloat16* raw = reinterpret_cast<const __nv_bfloat16*>(rows + rows_elems);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + HeadSmem::TMEM);
  const uint32_t bar_raw = sbase + HeadSmem::BAR_RAW, bar = sbase + HeadSmem::BAR, bar3 = sbase + HeadSmem::BAR3;
  const uint32_t raw_base = sbase + HeadSmem::ROWS + (uint32_t)rows_elems * 4u;
  const int tid = threadIdx.x, warp = tid >> 5;

  int row_off[4], raw_off[4];             // rows: floats; raw: bf16 elements
  {
    int off = 0, roff = 0;
#pragma unroll
    // a level whose temporal size equals the output's is sampled at exact frames: only its 2 h-corners exist
    for (int l = 0; l < 4; ++l) { row_off[l] = off; off += a.wl[l] * ROW_PITCH; raw_off[l] = roff; roff += (a.tl[l] == a.t ? 2 : 4) * a.wl[l] * HC; }
  }
  // thread 0 only: plan a row into shared memory and launch the bulk copies of its corner rows
  auto plan_and_fetch = [&](int tile, TilePlan* pl) {
    int r = tile;
    const int wt = r % w_tiles; r /= w_tiles;
    pl->h = r % a.h; r /= a.h;
    pl->t = r % a.t; pl->n = r / a.t;
    pl->w_base = wt * 128;
    uint32_t bytes = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const AxisTap at = axis_tap(pl->t, a.tl[l], a.t), ah = axis_tap(pl->h, a.hl[l], a.h);
      pl->ti[l][0] = at.i0; pl->ti[l][1] = at.i1; pl->hi[l][0] = ah.i0; pl->hi[l][1] = ah.i1;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float wg = ((c >> 1) ? at.l1 : at.l0) * ((c & 1) ? ah.l1 : ah.l0);
        pl->wgt[l][c] = wg;
        if (wg != 0.f) bytes += (uint32_t)a.wl[l] * HC * 2;
      }
    }
    fence_async_smem();
    mbar_arrive_expect_tx(bar_raw, bytes);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(a.g[l]) + (int64_t)pl->n * a.tl[l] * a.hl[l] * a.wl[l] * HC;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (pl->wgt[l][c] == 0.f) continue;
        const __nv_bfloat16* src = g + ((int64_t)pl->ti[l][c >> 1] * a.hl[l] + pl->hi[l][c & 1]) * a.wl[l] * HC;
        const uint32_t dst = raw_base + (uint32_t)((raw_off[l] + c * a.wl[l] * HC) * 2);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"((uint32_t)a.wl[l] * HC * 2), "r"(bar_raw) : "memory");
      }
    }
  };

  // ---- one-time setup
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + HeadSmem::TMEM), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar_raw, 1);
    mbar_init(bar, 1);
    mbar_init(bar3, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    // W2 (64 out x 64 in, bf16) into the K-major swizzled B tile: 512 chunks of 16 bytes, one per thread
    const int row = tid >> 3, chunk = tid & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.w2_bf16 + row * HC) + chunk);
    *reinterpret_cast<uint4*>(sm + HeadSmem::B + sw128_offset(row, chunk)) = v;
  }
  if (tid < 128) {
    // heads (6 x 64 fp32 -> bf16) into the K-major swizzled B3 tile, rows 6..15 zero: 128 chunks of 16 bytes
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
  auto fill_wtaps = [&](int w_base) {       // 4 levels x 128 voxels = 512 entries, one per thread
    const int l = tid >> 7, v = tid & 127;
    const AxisTap aw = axis_tap(min(w_base + v, a.w - 1), a.wl[l], a.w);
    WTap tp; tp.i0 = aw.i0; tp.i1 = aw.i1; tp.l0 = aw.l0; tp.l1 = aw.l1;
    wtap[tid] = tp;
  };
  int wtap_base = 0;
  fill_wtaps(0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint64_t desc_a = smem_desc_sw128(sbase + HeadSmem::A), desc_b = smem_desc_sw128(sbase + HeadSmem::B),
                 desc_b3 = smem_desc_sw128(sbase + HeadSmem::B3);
  const uint32_t idesc1 = idesc_bf16_f32(128, 64), idesc2 = idesc_bf16_f32(128, 16);

  int tile = blockIdx.x;
  if (tile < total_tiles && tid == 0) plan_and_fetch(tile, &plans[0]);
  __syncthreads();
  uint32_t phase = 0;
  for (; tile < total_tiles; tile += gridDim.x, phase ^= 1u) {
    const TilePlan& p = plans[phase];
    if (p.w_base != wtap_base) {            // only for frames wider than 128 voxels (uniform branch)
      __syncthreads();
      wtap_base = p.w_base;
      fill_wtaps(wtap_base);
      __syncthreads();
    }
    // ---- phase 1: T/H-interpolated rows (fp32) from the raw corner rows, rows[off_l + x*ROW_PITCH + c]
    mbar_wait(bar_raw, phase);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      float wgt[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) wgt[c] = p.wgt[l][c];
      const __nv_bfloat16* src = raw + raw_off[l];
      float* dst = rows + row_off[l];
      const int total = a.wl[l] * (HC / 4);
      for (int i = tid; i < total; i += HU_THREADS) {
        const int x = i >> 4, c4 = i & 15;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (wgt[c] == 0.f) continue;
          float f[4];
          bf16x4_to_f32(*reinterpret_cast<const uint2*>(src + ((size_t)c * a.wl[l] + x) * HC + 4 * c4), f);
          o.x = fmaf(wgt[c], f[0], o.x); o.y = fmaf(wgt[c], f[1], o.y); o.z = fmaf(wgt[c], f[2], o.z); o.w = fmaf(wgt[c], f[3], o.w);
        }
        *reinterpret_cast<float4*>(dst + x * ROW_PITCH + 4 * c4) = o;
      }
    }
    tc_fence_before();          // (the epilogue of the previous row read TMEM)
    __syncthreads();
    // the raw buffer is free: plan the next row and fetch its corner rows behind the rest of this row
    if (tid == 0 && tile + (int)gridDim.x < total_tiles) plan_and_fetch(tile + gridDim.x, &plans[phase ^ 1u]);

    // ---- phase 2: A tile.  thread = (voxel group vg of 4 voxels, channel group c4 of 4 channels)
    {
      const int vg = tid >> 4, c4 = tid & 15;
      float4 f[4];
      const float4 bias = *reinterpret_cast<const float4*>(b1s + 4 * c4);
#pragma unroll
      for (int v = 0; v < 4; ++v) f[v] = bias;
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float* r = rows + row_off[l] + 4 * c4;
        int cached = -1;
        float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const WTap tp = wtap[l * 128 + vg * 4 + v];
          if (tp.i0 != cached) { tv = *reinterpret_cast<const float4*>(r + tp.i0 * ROW_PITCH); cached = tp.i0; }
          fma4(f[v], tp.l0, tv);
          if (tp.l1 != 0.f) {
            if (tp.i1 != cached) { tv = *reinterpret_cast<const float4*>(r + tp.i1 * ROW_PITCH); cached = tp.i1; }
            fma4(f[v], tp.l1, tv);
          }
        }
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const uint32_t row = (uint32_t)(vg * 4 + v);
        uint2 pk;
        pk.x = pack_relu_bf16(f[v].x, f[v].y); pk.y = pack_relu_bf16(f[v].z, f[v].w);
        *reinterpret_cast<uint2*>(sm + HeadSmem::A + sw128_offset(row, (uint32_t)(c4 >> 1)) + (uint32_t)((c4 & 1) * 8)) = pk;
      }
    }
    fence_async_smem();          // generic-proxy writes of A (and B) -> visible to the tensor core's async proxy
    tc_fence_before();
    __syncthreads();

    // ---- MMA 1: D[128 x 64] = A[128 x 64] * W2[64 x 64]^T
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base, desc_a + (uint64_t)(2 * k), desc_b + (uint64_t)(2 * k), idesc1, k > 0 ? 1u : 0u);
      tc_commit(bar);
    }

    // ---- mid: warp = (TMEM lane quarter q, 16-column chunk): h2 = relu(D + b2) -> bf16 -> A tile
    mbar_wait(bar, phase);
    tc_fence_after();
    const int q = warp & 3, chunk = warp >> 2;
    const int vrow = q * 32 + (tid & 31);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    {
      uint32_t acc[16];
      tc_ld16(taddr + (uint32_t)(16 * chunk), acc);
      tc_wait_ld();
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        pk[j] = pack_relu_bf16(__uint_as_float(acc[2 * j]) + b2s[16 * chunk + 2 * j], __uint_as_float(acc[2 * j + 1]) + b2s[16 * chunk + 2 * j + 1]);
      *reinterpret_cast<uint4*>(sm + HeadSmem::A + sw128_offset((uint32_t)vrow, (uint32_t)(2 * chunk))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(sm + HeadSmem::A + sw128_offset((uint32_t)vrow, (uint32_t)(2 * chunk + 1))) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();            // every warp has drained its accumulators: the same TMEM columns take the head outputs

    // ---- MMA 2: D3[128 x 16] = relu(h2)[128 x 64] * Wh[16 x 64]^T
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_bf16(tmem_base, desc_a + (uint64_t)(2 * k), desc_b3 + (uint64_t)(2 * k), idesc2, k > 0 ? 1u : 0u);
      tc_commit(bar3);
    }

    // ---- epilogue: the first warp of every lane quarter, thread = voxel
    if (chunk == 0) {
      mbar_wait(bar3, phase);
      tc_fence_after();
      uint32_t r8[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r8[0]), "=r"(r8[1]), "=r"(r8[2]), "=r"(r8[3]), "=r"(r8[4]), "=r"(r8[5]), "=r"(r8[6]), "=r"(r8[7]) : "r"(taddr));
      tc_wait_ld();
      float o[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) o[k] = __uint_as_float(r8[k]) + bhs[k];
      const int w = p.w_base + vrow;
      if (w < a.w) {
        float s0 = o[0], s1 = o[1];
        if (a.out_kind == CLASFV_OUT_PROB) {
          const float mx = fmaxf(s0, s1);
          const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
          const float inv = 1.f / (e0 + e1);
          s0 = e0 * inv; s1 = e1 * inv;
        }
        const int64_t plane = (int64_t)a.h * a.w;
        const int64_t pix = (int64_t)p.h * a.w + w;
        OutT* seg = static_cast<OutT*>(a.seg) + ((int64_t)p.n * 2 * a.t + p.t) * plane + pix;
        put<OutT>(seg, s0);
        put<OutT>(seg + (int64_t)a.t * plane, s1);
        OutT* mot = static_cast<OutT*>(a.motion) + ((int64_t)p.n * 4 * a.t + p.t) * plane + pix;
#pragma unroll
        for (int k = 0; k < 4; ++k) put<OutT>(mot + (int64_t)k * a.t * plane, tanhf(o[2 + k]));
      }
    }
    // the other warps run ahead into the next row's phase 1; MMA 1 of that row is issued only after two more
    // __syncthreads, by which time the epilogue warps above have drained D3
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
  int rowbuf = 0, rawelems = 0;
  for (int l = 0; l < 4; ++l) { rowbuf += a.wl[l] * ROW_PITCH; rawelems += (a.tl[l] == a.t ? 2 : 4) * a.wl[l] * HC; }
  const size_t smem = 1024 + HeadSmem::ROWS + (size_t)rowbuf * sizeof(float) + (size_t)rawelems * sizeof(__nv_bfloat16);
  static_assert(sizeof(TilePlan) * 2 <= HeadSmem::WTAP - HeadSmem::PLAN, "plan slots overflow");
  static_assert(HeadSmem::TMEM + 4 <= HeadSmem::PLAN, "smem header overlap");
  CLASFV_REQUIRE(smem <= 227 * 1024, "head_umma: frame too wide for the row buffers (W=%d)", a.w);
  const int w_tiles = (a.w + 127) / 128;
  const int64_t total = (int64_t)a.n * a.t * a.h * w_tiles;
  CLASFV_REQUIRE(total < (1ll << 31), "head_umma: too many rows");
  int dev = 0, sms = 0;
  CLASFV_CUDA(cudaGetDevice(&dev));
  CLASFV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int per_sm = smem <= 113 * 1024 ? 2 : 1;
  const int grid = (int)std::min<int64_t>(total, (int64_t)sms * per_sm);
  if (a.out_dtype == CLASFV_F32) {
    CLASFV_CUDA(cudaFuncSetAttribute(head_umma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_umma_kernel<float><<<grid, HU_THREADS, smem, stream>>>(a, w_tiles, (int)total, rowbuf);
  } else {
    CLASFV_CUDA(cudaFuncSetAttribute(head_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_umma_kernel<__nv_bfloat16><<<grid, HU_THREADS, smem, stream>>>(a, w_tiles, (int)total, rowbuf);
  }
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

}  // namespace clasfv
