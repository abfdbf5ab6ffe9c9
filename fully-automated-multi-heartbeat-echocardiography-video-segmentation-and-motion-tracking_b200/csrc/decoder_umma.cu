// decoder_umma.cu - the fused decoder head (upsample x4 levels + comb_1 bias/ReLU + comb_2 + heads + softmax/tanh)
// with everything that is a contraction on the tensor cores, as a warp-specialised pipeline.  sm_100a only.
//
// Reference: src/model/R2plus1D_18_MotionNet.py:41-71 (trilinear align_corners=True upsampling of the five feature
// maps, cat, comb_1 + BN + ReLU, comb_2 + BN + ReLU, segmentation / motion heads, tanh) and src/fuse_utils.py:60
// (softmax).  comb_1 has been commuted with the upsampling (api.cu): the inputs here are the four laterally
// projected 64-channel maps g_l at 1/2, 1/4, 1/8, 1/16 resolution.
//
// One output row (n, t, h) of up to 128 voxels is one "tile".  For that row
//   R_l[x, c] = sum over the <= 4 (T,H) corners  wT * wH * g_l[n, t_i, h_i, x, c]        (CUDA cores, "phase 1", -> fp16)
//   h1[v, c]  = relu( sum_l sum_x Wmat_l[v, x] * R_l[x, c] + b1[c] )                      (MMA 0)
//   h2[v, c]  = relu( sum_k h1[v, k] * W2[c, k] + b2[c] )                                 (MMA 1)
//   o[v, j]   = sum_k h2[v, k] * Wh[j, k] + bh[j]   -> softmax / tanh -> 6 planar stores   (MMA 2)
// MMA 0 is the W-axis interpolation written as a GEMM: its A operand is the constant interpolation matrix of the
// row [128 voxels x K], K = the low-resolution columns of the four levels side by side (+2 columns of ones that
// pick up b1 as two extra rows of R), split into a bf16 high and low part so the weights carry ~16 mantissa bits;
// its B operand is R, built per row in shared memory directly in the MN-major SWIZZLE_128B layout (K rows of 64
// channels = 128 bytes).  ReLU + bf16 conversion of an accumulator is one cvt.rn.relu.bf16x2.f32 per two values.
//
// Warp roles (768 threads, persistent, one CTA per SM); every hand-over is an mbarrier, every buffer is at least doubled.
// Phase 1 is the throughput-limiting role (clock64 timeline), so it gets every warp the epilogues can spare:
//   warp 0           producer: plans the row (corner indices / weights), bulk-copies its raw corner rows (bf16)
//   warp 1           MMA issuer (one lane): per step MMA0(i), MMA1(i-1), MMA2(i-2)
//   warps 4-7        epilogue 0 of row i (acc0 -> relu -> bf16 -> tensor memory = A of MMA 1) and epilogue 2 of row i-2
//                    (acc2 + bh -> softmax / tanh -> global); one warp per TMEM lane quarter
//   warps 8-11       epilogue 1 (acc1 -> relu -> bf16 -> tensor memory = A of MMA 2; comb_2's bias rides in K)
//   warps 2,3,12-23  phase 1: raw corner rows -> R (two groups of 7 warps on alternate rows)
#include "internal.h"
#include "umma_ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <cuda_fp16.h>

namespace clasfv {
namespace {

using namespace ptx;

constexpr int HU_THREADS = 768;
constexpr int HC = 64;
constexpr int P1_WARPS = 14;
constexpr int P1_THREADS = P1_WARPS * 32;
constexpr int MAX_WT = 4;                 // w tiles of 128 voxels (W <= 512)
constexpr int NRAW = 4;                   // raw corner-row stages wanted (as many as fit shared memory, at least 2)
constexpr int NR = 3;                     // R stages (phase 1 runs up to two rows ahead of MMA 0; 2 when 3 do not fit)
constexpr int KSLABS = 2;                 // K of MMA 0 is always 2 slabs of 64 (interpolation columns + 2 bias rows, zero padded)

struct AxisTap { int i0, i1; float l0, l1; };
__host__ __device__ __forceinline__ AxisTap axis_tap(int dst, int in_size, int out_size) {
  AxisTap a;
  const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  const float src = scale * (float)dst;
  a.i0 = (int)src < in_size - 1 ? (int)src : in_size - 1;
  a.i1 = a.i0 + (a.i0 < in_size - 1 ? 1 : 0);
  a.l1 = src - (float)a.i0;
  a.l0 = 1.f - a.l1;
  return a;
}

template <typename OutT> __device__ __forceinline__ void put(OutT* p, float v);
template <> __device__ __forceinline__ void put<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void put<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// {low 16 bits = bf16(relu(a)), high 16 bits = bf16(relu(b))}: a is the element at the lower address
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

// fp16 pair, saturating to the largest finite value instead of overflowing to infinity
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float a, float b) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

// tcgen05.mma with the A operand in tensor memory (lane = row, two 16-bit K elements per 32-bit column)
__device__ __forceinline__ void tc_mma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// compiler-level fence: values produced by an asynchronous tcgen05.ld may not be consumed before tcgen05.wait::ld
__device__ __forceinline__ void reg_fence16(uint32_t (&r)[16]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}

// Wait used by every role except the MMA issuer: back off between polls so that twenty-odd waiting warps do not
// take issue slots from the one thread that feeds the tensor core.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done) {
      __nanosleep(40);
      if (spin > (1u << 22)) { printf("clasfv head_umma: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
  }
}

struct HeadGeom {                 // host-computed layout of the K axis of MMA 0 and of a raw corner-row stage
  int w_tiles, ktot, kpad, nslab;
  int koff[4];                    // first K row of level l
  int nxmax[4];                   // K rows reserved for level l
  int xlo[MAX_WT][4], nx[MAX_WT][4];   // low-resolution columns [xlo, xlo+nx) the voxels of w tile wt touch
  int raw_off[4];                 // byte offset of level l inside a raw stage: corner slots of nxmax columns each (4, or 2
                                  // when the level is at the output's temporal resolution and only its H corners exist)
  int raw_stage_bytes;
  int total_rows;                 // n * t * h
  int nraw;                       // raw stages actually used (2 or NRAW)
  int nr;                         // R stages actually used (2 or NR)
  int wt_vox;                     // voxels per w tile (128, or fewer when the tile's interpolation columns would exceed K)
};

struct TilePlan { float wgt[4][4]; };   // weight of each (T,H) corner per level, [level][2*tc + hc]; 0 = corner not fetched

struct Smem {                     // byte offsets from the 1024-aligned base
  static constexpr uint32_t W2B = 0;                      // 2 slabs of 64 x 128 B: W2 (K 0..63), then K 64..79 = {b2 hi, b2 lo, 0...}
  static constexpr uint32_t WHB = W2B + 16384;            // 16 x 128 B   heads, K-major (rows 6..15 zero)
  static constexpr uint32_t BH = WHB + 2048;              // [8] fp32
  static constexpr uint32_t PLAN = BH + 32;               // NRAW x TilePlan
  static constexpr uint32_t BARS = PLAN + NRAW * 64;      // 14 groups of up to 4 mbarriers
  static constexpr uint32_t TMEM = BARS + 8 * 56;
  static constexpr uint32_t WA = 20480;                   // interpolation matrix (fp16), staging only: KSLABS slabs of 16 KB
  // then: R (nr stages x KSLABS x 8 KB), raw (nraw stages x raw_stage_bytes)
};
static_assert(Smem::TMEM + 4 <= Smem::WA, "head smem header overflow");
static_assert(sizeof(TilePlan) == 64, "plan slot size");

enum Bar {  // index of the first mbarrier of each group (one per buffer stage, up to 4 stages)
  RAW_FULL = 0, RAW_EMPTY = 4, R_FULL = 8, R_EMPTY = 12, ACC0_FULL = 16, ACC0_EMPTY = 20, A1_FULL = 24, A1_EMPTY = 28,
  ACC1_FULL = 32, ACC1_EMPTY = 36, A2_FULL = 40, ACC2_FULL = 44, ACC2_EMPTY = 48, A2_EMPTY = 52,
};

template <typename OutT>
__global__ void __launch_bounds__(HU_THREADS, 1) head_umma_kernel(const HeadArgs a, const HeadGeom g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(sm);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* bhs = reinterpret_cast<float*>(sm + Smem::BH);
  TilePlan* plans = reinterpret_cast<TilePlan*>(sm + Smem::PLAN);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + Smem::TMEM);
  auto bar = [&](int which, int s) { return sbase + Smem::BARS + 8u * (uint32_t)(which + s); };
  const uint32_t wa_bytes = (uint32_t)KSLABS * 16384u;
  const uint32_t r_stage_bytes = (uint32_t)g.nslab * 8192u;
  const uint32_t r_off = Smem::WA + wa_bytes;
  const uint32_t raw_off0 = r_off + (uint32_t)g.nr * r_stage_bytes;

  const int wt = blockIdx.x % g.w_tiles;
  const int row0 = blockIdx.x / g.w_tiles, row_step = gridDim.x / g.w_tiles;
  const int w_base = wt * g.wt_vox;
  const int my_rows = row0 < g.total_rows ? (g.total_rows - row0 + row_step - 1) / row_step : 0;

  // ------------------------------------------------------------------ one-time setup (all threads)
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + Smem::TMEM), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < NRAW; ++s) { mbar_init(bar(RAW_FULL, s), 1); mbar_init(bar(RAW_EMPTY, s), P1_WARPS / 2); }
    for (int s = 0; s < NR; ++s) { mbar_init(bar(R_FULL, s), P1_WARPS / 2); mbar_init(bar(R_EMPTY, s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(ACC0_FULL, s), 1); mbar_init(bar(ACC0_EMPTY, s), 4);
      mbar_init(bar(A1_FULL, s), 4); mbar_init(bar(A1_EMPTY, s), 1); mbar_init(bar(A2_EMPTY, s), 1);
      mbar_init(bar(ACC1_FULL, s), 1); mbar_init(bar(ACC1_EMPTY, s), 4);
      mbar_init(bar(A2_FULL, s), 4); mbar_init(bar(ACC2_FULL, s), 1); mbar_init(bar(ACC2_EMPTY, s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the interpolation matrix, the R stages (unused K rows / columns must be exact zeros, not stale NaNs) and the
  // raw stages (phase 1 reads the slot of an unfetched corner with weight 0)
  for (uint32_t i = (uint32_t)tid * 16u; i < wa_bytes + (uint32_t)g.nr * r_stage_bytes + (uint32_t)(g.nraw * g.raw_stage_bytes); i += HU_THREADS * 16u)
    *reinterpret_cast<uint4*>(sm + Smem::WA + i) = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 512) {
    // W2 (64 out x 64 in, bf16) into the K-major swizzled B tile: 512 chunks of 16 bytes
    const int row = tid >> 3, chunk = tid & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.w2_bf16 + row * HC) + chunk);
    *reinterpret_cast<uint4*>(sm + Smem::W2B + sw128_offset(row, chunk)) = v;
    // second K slab of W2: K 64, 65 carry the folded comb_2 bias (hi + lo bf16), picked up by two columns of ones in the
    // A tile, so that epilogue 1 is a pure ReLU + convert like epilogue 0
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if (chunk == 0) {
      const float b = __ldg(a.b2 + row);
      const __nv_bfloat16 hi = __float2bfloat16_rn(b), lo = __float2bfloat16_rn(b - __bfloat162float(hi));
      x.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
    }
    *reinterpret_cast<uint4*>(sm + Smem::W2B + 8192u + sw128_offset(row, chunk)) = x;
  } else if (tid < 640) {
    // heads (6 x 64 fp32 -> bf16) into the K-major swizzled tile, rows 6..15 zero
    const int i = tid - 512, row = i >> 3, chunk = i & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < 6) {
      const float* src = a.wh + row * HC + chunk * 8;
      v.x = cvt_bf16x2(__ldg(src + 0), __ldg(src + 1)); v.y = cvt_bf16x2(__ldg(src + 2), __ldg(src + 3));
      v.z = cvt_bf16x2(__ldg(src + 4), __ldg(src + 5)); v.w = cvt_bf16x2(__ldg(src + 6), __ldg(src + 7));
    }
    *reinterpret_cast<uint4*>(sm + Smem::WHB + sw128_offset(row, chunk)) = v;
  } else if (tid < 646) {
    bhs[tid - 640] = __ldg(a.bh + tid - 640);
  }
  __syncthreads();
  {
    // interpolation matrix: element (voxel v, K column k) of slab k/64, hi and lo parts
    // fp16 (11-bit significand, weights in [0,1]): 8x finer than the bf16 rounding of everything downstream
    auto put_w = [&](int v, int k, float w) {
      const uint32_t o = (uint32_t)(k >> 6) * 16384u + sw128_offset((uint32_t)v, (uint32_t)((k & 63) >> 3)) + (uint32_t)(k & 7) * 2u;
      *reinterpret_cast<__half*>(sm + Smem::WA + o) = __float2half_rn(w);
    };
    for (int i = tid; i < 4 * 128; i += HU_THREADS) {
      const int l = i >> 7, v = i & 127;
      if (v < g.wt_vox && w_base + v < a.w) {
        const AxisTap aw = axis_tap(w_base + v, a.wl[l], a.w);
        put_w(v, g.koff[l] + aw.i0 - g.xlo[wt][l], aw.l0);
        if (aw.l1 != 0.f) put_w(v, g.koff[l] + aw.i1 - g.xlo[wt][l], aw.l1);
      }
    }
    if (tid >= 512 && tid < 640) { put_w(tid - 512, g.ktot - 2, 1.f); put_w(tid - 512, g.ktot - 1, 1.f); }
    // the two constant rows of R: b1 = hi + lo (both stages)
    if (tid >= 640 && tid < 640 + HC) {
      const int c = tid - 640;
      const float b = __ldg(a.b1 + c);
      const __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
      for (int s = 0; s < g.nr; ++s) {
        const uint32_t base = r_off + (uint32_t)s * r_stage_bytes;
        *reinterpret_cast<__half*>(sm + base + sw128_offset((uint32_t)(g.ktot - 2), (uint32_t)(c >> 3)) + (uint32_t)(c & 7) * 2u) = hi;
        *reinterpret_cast<__half*>(sm + base + sw128_offset((uint32_t)(g.ktot - 1), (uint32_t)(c >> 3)) + (uint32_t)(c & 7) * 2u) = lo;
      }
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc0 = tmem_base, acc1 = tmem_base + 128u, acc2 = tmem_base + 256u;   // 2 x 64, 2 x 64, 2 x 16 columns
  const uint32_t tmem_wa = tmem_base + 288u;                                           // 64 columns: the interpolation matrix
  // relu(h1): 2 stages x 40 columns (32 of data + 8 for K 64..79 = {1, 1, 0, ...}: the bias columns); relu(h2): 2 x 32
  const uint32_t tmem_a1 = tmem_base + 352u, tmem_a2 = tmem_base + 432u;
  if (warp >= 4 && warp < 8) {
    // The interpolation matrix is the A operand of every MMA 0 and never changes: it lives in tensor memory
    // (128 lanes x 64 columns, two fp16 K elements per column), so MMA 0 reads no A bytes from shared memory.
    const int v = (warp & 3) * 32 + lane;
#pragma unroll
    for (int c16 = 0; c16 < 4; ++c16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = 2 * (c16 * 16 + j);
        r[j] = *reinterpret_cast<const uint32_t*>(sm + Smem::WA + (uint32_t)(k >> 6) * 16384u + sw128_offset((uint32_t)v, (uint32_t)((k & 63) >> 3)) + (uint32_t)(k & 7) * 2u);
      }
      tc_st16(tmem_wa + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(c16 * 16), r);
    }
    // constant bias columns of both relu(h1) stages: K 64 and 65 = 1.0 (bf16 0x3F80), K 66..79 = 0
#pragma unroll
    for (int st = 0; st < 2; ++st)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %2, %2, %2, %2, %2, %2};"
                   ::"r"(tmem_a1 + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(st * 40 + 32)), "r"(0x3F803F80u), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // rows of this CTA: row0, row0 + row_step, ...; (n, t, h) advance by a fixed carry-propagated step (no divisions
  // in the per-row paths: the producer's planning latency is on the critical path of the whole pipeline)
  struct RowIter {
    int n, t, h, dn, dt, dh, T, H;
    __device__ __forceinline__ void next() {
      h += dh; int c = h >= H ? 1 : 0; h -= c * H;
      t += dt + c; c = t >= T ? 1 : 0; t -= c * T;
      n += dn + c;
    }
  };
  RowIter it0;
  it0.T = a.t; it0.H = a.h;
  it0.h = row0 % a.h; it0.t = (row0 / a.h) % a.t; it0.n = row0 / (a.h * a.t);
  it0.dh = row_step % a.h; it0.dt = (row_step / a.h) % a.t; it0.dn = row_step / (a.h * a.t);

  if (warp == 0) {
    // ================================================================ producer (lane = level * 4 + corner, 16 lanes)
    const int l = (lane >> 2) & 3, c = lane & 3;
    const bool has_slot = lane < 16 && (c < 2 || a.tl[l] != a.t);
    const __nv_bfloat16* gbase = static_cast<const __nv_bfloat16*>(a.g[l]) + (int64_t)g.xlo[wt][l] * HC;
    const int64_t clip_elems = (int64_t)a.tl[l] * a.hl[l] * a.wl[l] * HC, row_elems = (int64_t)a.wl[l] * HC;
    const uint32_t my_bytes = (uint32_t)g.nx[wt][l] * 128u;
    const uint32_t my_dst = sbase + raw_off0 + (uint32_t)g.raw_off[l] + (uint32_t)(c * g.nxmax[l]) * 128u;
    RowIter it = it0;
    for (int i = 0; i < my_rows; ++i, it.next()) {
      const int s = i % g.nraw; const uint32_t ph = (uint32_t)(i / g.nraw) & 1u;
      const AxisTap at = axis_tap(it.t, a.tl[l], a.t), ah = axis_tap(it.h, a.hl[l], a.h);
      const float wg = ((c >> 1) ? at.l1 : at.l0) * ((c & 1) ? ah.l1 : ah.l0);
      const bool fetch = has_slot && wg != 0.f;
      const __nv_bfloat16* src = gbase + it.n * clip_elems + ((int64_t)((c >> 1) ? at.i1 : at.i0) * a.hl[l] + ((c & 1) ? ah.i1 : ah.i0)) * row_elems;
      const uint32_t bytes = __reduce_add_sync(0xffffffffu, fetch ? my_bytes : 0u);
      mbar_wait(bar(RAW_EMPTY, s), ph ^ 1u);
      if (lane < 16) plans[s].wgt[l][c] = fetch ? wg : 0.f;
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar(RAW_FULL, s), bytes);
      __syncwarp();
      if (fetch)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(my_dst + (uint32_t)s * (uint32_t)g.raw_stage_bytes), "l"(src), "r"(my_bytes), "r"(bar(RAW_FULL, s)) : "memory");
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    // The whole warp runs this loop (uniform control flow, so descriptors live in uniform registers); only the
    // tcgen05 instructions themselves are issued by one elected lane.
    // MMA 0: A = fp16 interpolation weights, B = R in fp16 (formats 0; the hardware rejects mixed f16 x bf16 operands),
    // MN-major B, fp32 accumulate
    const uint32_t idesc0 = (idesc_bf16_f32(128, 64) & ~((7u << 7) | (7u << 10))) | (1u << 16);
    const uint32_t idesc1 = idesc_bf16_f32(128, 64), idesc2 = idesc_bf16_f32(128, 16);
    const uint64_t desc_w2 = smem_desc_sw128(sbase + Smem::W2B), desc_wh = smem_desc_sw128(sbase + Smem::WHB);
    const uint64_t desc_wa = smem_desc_sw128(sbase + Smem::WA), desc_r = smem_desc_sw128(sbase + r_off);
    for (int k = 0; k < my_rows + 2; ++k) {
      if (k < my_rows) {
        const int s = k & 1; const uint32_t ph = (uint32_t)(k >> 1) & 1u;
        const int rs = k % g.nr; const uint32_t rph = (uint32_t)(k / g.nr) & 1u;
        mbar_wait(bar(R_FULL, rs), rph);
        mbar_wait(bar(ACC0_EMPTY, s), ph ^ 1u);
        tc_fence_after();
        // K is always 2 slabs = 128 (zero rows / columns beyond ktot): 16 instructions with constant descriptor
        // offsets.  MN-major B: 16 K rows of 128 bytes per instruction = two 8-row swizzle atoms (SBO = 1024), so K
        // advances by 2048 bytes = 128 descriptor units.
        const uint64_t db0 = desc_r + (uint64_t)(rs * (int)(KSLABS * 8192 / 16));
        const uint32_t d0 = acc0 + (uint32_t)(s * 64);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < KSLABS * 4; ++kk)
            tc_mma_ts_f16(d0, tmem_wa + (uint32_t)(kk * 8), db0 + (uint64_t)(kk * 128), idesc0, kk ? 1u : 0u);
          tc_commit(bar(R_EMPTY, rs));
          tc_commit(bar(ACC0_FULL, s));
        }
        __syncwarp();
      }
      if (k >= 1 && k <= my_rows) {
        const int j = k - 1, s = j & 1; const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        mbar_wait(bar(A1_FULL, s), ph);
        mbar_wait(bar(ACC1_EMPTY, s), ph ^ 1u);
        tc_fence_after();
        const uint32_t ta = tmem_a1 + (uint32_t)(s * 40);
        const uint32_t d1 = acc1 + (uint32_t)(s * 64);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_ts_f16(d1, ta + (uint32_t)(8 * kk), desc_w2 + (uint64_t)(2 * kk), idesc1, kk > 0 ? 1u : 0u);
          tc_mma_ts_f16(d1, ta + 32u, desc_w2 + (uint64_t)(8192 / 16), idesc1, 1u);      // + b2 (bias slab)
          tc_commit(bar(A1_EMPTY, s));
          tc_commit(bar(ACC1_FULL, s));
        }
        __syncwarp();
      }
      if (k >= 2) {
        const int j = k - 2, s = j & 1; const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        mbar_wait(bar(A2_FULL, s), ph);
        mbar_wait(bar(ACC2_EMPTY, s), ph ^ 1u);
        tc_fence_after();
        const uint32_t ta = tmem_a2 + (uint32_t)(s * 32);
        const uint32_t d2 = acc2 + (uint32_t)(s * 16);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_ts_f16(d2, ta + (uint32_t)(8 * kk), desc_wh + (uint64_t)(2 * kk), idesc2, kk > 0 ? 1u : 0u);
          tc_commit(bar(A2_EMPTY, s));
          tc_commit(bar(ACC2_FULL, s));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ================================================================ epilogues (warp % 4 = TMEM lane quarter)
    // Two groups of four warps: warps 4-7 run epilogue 0 of row i and epilogue 2 of row i-2, warps 8-11 epilogue 1
    // (measured: one group for all three is slower, three groups leave phase 1 four warps short).  Epilogues 0 and 1 are
    // the same code: accumulator -> ReLU -> bf16 -> the tensor-memory A tile of the next GEMM (biases ride in K).
    const int role = (warp - 4) >> 2, q = warp & 3;
    const int vrow = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    auto relu_to_tmem = [&](uint32_t taddr, uint32_t aaddr) {
      uint32_t v0[16], v1[16];
      tc_ld16(taddr, v0);
      tc_ld16(taddr + 16u, v1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tc_wait_ld(); reg_fence16(v0); reg_fence16(v1);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          pk[j] = cvt_relu_bf16x2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1]));
          pk[8 + j] = cvt_relu_bf16x2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1]));
        }
        if (half == 0) { tc_ld16(taddr + 32u, v0); tc_ld16(taddr + 48u, v1); }
        tc_st16(aaddr + (uint32_t)(16 * half), pk);       // 32 channels = 16 packed columns
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
    };
    const int64_t plane = (int64_t)a.h * a.w;
    const int w = w_base + vrow;
    RowIter it2 = it0;
    // epilogue 2 of row i (rows are visited in order; it2 follows)
    auto epilogue2 = [&](int i) {
      RowIter& it = it2;
      const int s = i & 1; const uint32_t ph = (uint32_t)(i >> 1) & 1u;
        const int n = it.n, t = it.t, h = it.h;
        mbar_wait_sleep(bar(ACC2_FULL, s), ph);
        tc_fence_after();
        uint32_t r8[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r8[0]), "=r"(r8[1]), "=r"(r8[2]), "=r"(r8[3]), "=r"(r8[4]), "=r"(r8[5]), "=r"(r8[6]), "=r"(r8[7])
                     : "r"(acc2 + lane_addr + (uint32_t)(s * 16)));
        tc_wait_ld();
        asm volatile("" : "+r"(r8[0]), "+r"(r8[1]), "+r"(r8[2]), "+r"(r8[3]), "+r"(r8[4]), "+r"(r8[5]), "+r"(r8[6]), "+r"(r8[7]));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(ACC2_EMPTY, s));
        float o[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) o[k] = __uint_as_float(r8[k]) + bhs[k];
        if (vrow < g.wt_vox && w < a.w) {
          float s0 = o[0], s1 = o[1];
          if (a.out_kind == CLASFV_OUT_PROB) {
            const float mx = fmaxf(s0, s1);
            const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
            const float inv = 1.f / (e0 + e1);
            s0 = e0 * inv; s1 = e1 * inv;
          }
          const int64_t pix = (int64_t)h * a.w + w;
          OutT* seg = static_cast<OutT*>(a.seg) + ((int64_t)n * 2 * a.t + t) * plane + pix;
          put<OutT>(seg, s0);
          put<OutT>(seg + (int64_t)a.t * plane, s1);
          OutT* mot = static_cast<OutT*>(a.motion) + ((int64_t)n * 4 * a.t + t) * plane + pix;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float th;     // MUFU.TANH: |error| ~ 2^-11, below the bf16 resolution of everything upstream in this mode
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(o[2 + k]));
            put<OutT>(mot + (int64_t)k * a.t * plane, th);
          }
        }
      it.next();
    };
    if (role == 0) {
      for (int k = 0; k < my_rows + 2; ++k) {
        if (k < my_rows) {
          const int s = k & 1; const uint32_t ph = (uint32_t)(k >> 1) & 1u;
          mbar_wait_sleep(bar(ACC0_FULL, s), ph);
          mbar_wait_sleep(bar(A1_EMPTY, s), ph ^ 1u);     // MMA 1 of row k-2 has finished reading this A tile
          tc_fence_after();
          relu_to_tmem(acc0 + lane_addr + (uint32_t)(s * 64), tmem_a1 + lane_addr + (uint32_t)(s * 40));
          if (lane == 0) { mbar_arrive(bar(ACC0_EMPTY, s)); mbar_arrive(bar(A1_FULL, s)); }
        }
        if (k >= 2) epilogue2(k - 2);                     // the same warps drain the head accumulators two rows behind
      }
    } else {
      for (int j = 0; j < my_rows; ++j) {
        const int s = j & 1; const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        mbar_wait_sleep(bar(ACC1_FULL, s), ph);
        mbar_wait_sleep(bar(A2_EMPTY, s), ph ^ 1u);       // MMA 2 of row j-2 has finished reading this A tile
        tc_fence_after();
        relu_to_tmem(acc1 + lane_addr + (uint32_t)(s * 64), tmem_a2 + lane_addr + (uint32_t)(s * 32));
        if (lane == 0) { mbar_arrive(bar(ACC1_EMPTY, s)); mbar_arrive(bar(A2_FULL, s)); }
      }
    }
  } else {
    // ================================================================ phase 1: R = (T,H)-interpolated rows, bf16, MN-major swizzled
    const int ptid = (warp < 4 ? warp - 2 : warp - 10) * 32 + lane;       // warps 2,3,12..23 -> 0 .. P1_THREADS-1
    // Two groups of 7 warps take alternate rows: the role is latency-bound (wait -> loads -> FMAs -> stores -> proxy fence
    // -> arrive is one dependent chain per row, ~3 000 clk in the clock64 timeline, at < 2 of 4 issue slots used), so two
    // rows' chains in flight beat ten warps on one row.  One flat item list over the four levels (item = one low-resolution
    // column x 8 channels), P1_ITEMS per thread and pass; all loads of a pass are issued before any arithmetic.  Corners 0/1
    // are read unconditionally (a corner that was not fetched has weight 0 and its slot holds finite stale data: the
    // stages are zero-filled at start).
    constexpr int P1_ITEMS = 2, P1_PASSES = 3, P1_SLOTS = P1_ITEMS * P1_PASSES, GT = P1_THREADS / 2;
    const int grp = ptid / GT, gtid = ptid % GT;
    const int q1 = g.nx[wt][0] * 8, q2 = q1 + g.nx[wt][1] * 8, q3 = q2 + g.nx[wt][2] * 8, n_items = q3 + g.nx[wt][3] * 8;
    // item -> (level, column, chunk) is the same for every row: source / destination offsets are computed once
    uint32_t i_src[P1_SLOTS], i_cs[P1_SLOTS], i_dst[P1_SLOTS], i_woff[P1_SLOTS];
    bool i_on[P1_SLOTS], i_four[P1_SLOTS];
#pragma unroll
    for (int u = 0; u < P1_SLOTS; ++u) {
      const int q = gtid + u * GT;
      i_on[u] = q < n_items;
      const int qq = i_on[u] ? q : 0;
      const int l = (qq >= q1 ? 1 : 0) + (qq >= q2 ? 1 : 0) + (qq >= q3 ? 1 : 0);
      const int it = qq - (l == 0 ? 0 : l == 1 ? q1 : l == 2 ? q2 : q3);
      i_src[u] = (uint32_t)(l == 0 ? g.raw_off[0] : l == 1 ? g.raw_off[1] : l == 2 ? g.raw_off[2] : g.raw_off[3]) + (uint32_t)it * 16u;
      i_cs[u] = (uint32_t)(l == 0 ? g.nxmax[0] : l == 1 ? g.nxmax[1] : l == 2 ? g.nxmax[2] : g.nxmax[3]) * 128u;
      i_dst[u] = sw128_offset((uint32_t)((l == 0 ? g.koff[0] : l == 1 ? g.koff[1] : l == 2 ? g.koff[2] : g.koff[3]) + (it >> 3)), (uint32_t)(it & 7));
      i_woff[u] = (uint32_t)l * 16u;
      i_four[u] = (l == 0 ? a.tl[0] : l == 1 ? a.tl[1] : l == 2 ? a.tl[2] : a.tl[3]) != a.t;   // the level has T corners
    }
    for (int i = grp; i < my_rows; i += 2) {
      const int s = i % g.nr; const uint32_t ph = (uint32_t)(i / g.nr) & 1u;
      const int rs = i % g.nraw; const uint32_t rph = (uint32_t)(i / g.nraw) & 1u;
      mbar_wait(bar(RAW_FULL, rs), rph);
      mbar_wait(bar(R_EMPTY, s), ph ^ 1u);               // MMA 0 of row i-nr has finished reading this R stage
      const uint8_t* plw = reinterpret_cast<const uint8_t*>(&plans[rs]);
      const uint8_t* stage = sm + raw_off0 + (uint32_t)rs * (uint32_t)g.raw_stage_bytes;
      uint8_t* rst = sm + r_off + (uint32_t)s * r_stage_bytes;
#pragma unroll
      for (int pass = 0; pass < P1_PASSES; ++pass) {
        uint4 ra[P1_ITEMS], rb[P1_ITEMS];
        float4 wq[P1_ITEMS];
#pragma unroll
        for (int v = 0; v < P1_ITEMS; ++v) {
          const int u = pass * P1_ITEMS + v;
          if (!i_on[u]) continue;
          wq[v] = *reinterpret_cast<const float4*>(plw + i_woff[u]);
          ra[v] = *reinterpret_cast<const uint4*>(stage + i_src[u]);
          rb[v] = *reinterpret_cast<const uint4*>(stage + i_src[u] + i_cs[u]);
        }
#pragma unroll
        for (int v = 0; v < P1_ITEMS; ++v) {
          const int u = pass * P1_ITEMS + v;
          if (!i_on[u]) continue;
          const uint32_t xa[4] = {ra[v].x, ra[v].y, ra[v].z, ra[v].w}, xb[4] = {rb[v].x, rb[v].y, rb[v].z, rb[v].w};
          float lo[4], hi[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            lo[e] = wq[v].x * __uint_as_float(xa[e] << 16); hi[e] = wq[v].x * __uint_as_float(xa[e] & 0xffff0000u);
            lo[e] = fmaf(wq[v].y, __uint_as_float(xb[e] << 16), lo[e]); hi[e] = fmaf(wq[v].y, __uint_as_float(xb[e] & 0xffff0000u), hi[e]);
          }
          if (i_four[u]) {
            const uint4 rc = *reinterpret_cast<const uint4*>(stage + i_src[u] + 2u * i_cs[u]);
            const uint4 rd = *reinterpret_cast<const uint4*>(stage + i_src[u] + 3u * i_cs[u]);
            const uint32_t xc[4] = {rc.x, rc.y, rc.z, rc.w}, xd[4] = {rd.x, rd.y, rd.z, rd.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              lo[e] = fmaf(wq[v].z, __uint_as_float(xc[e] << 16), lo[e]); hi[e] = fmaf(wq[v].z, __uint_as_float(xc[e] & 0xffff0000u), hi[e]);
              lo[e] = fmaf(wq[v].w, __uint_as_float(xd[e] << 16), lo[e]); hi[e] = fmaf(wq[v].w, __uint_as_float(xd[e] & 0xffff0000u), hi[e]);
            }
          }
          *reinterpret_cast<uint4*>(rst + i_dst[u]) =
              make_uint4(cvt_f16x2_sat(lo[0], hi[0]), cvt_f16x2_sat(lo[1], hi[1]), cvt_f16x2_sat(lo[2], hi[2]), cvt_f16x2_sat(lo[3], hi[3]));
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar(R_FULL, s)); mbar_arrive(bar(RAW_EMPTY, rs)); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Temporal pre-pass: levels at reduced temporal resolution are interpolated along T once per output frame
// (trilinear is separable; align_corners=True), so that the head only ever needs the two H corner rows of a level.
__global__ void __launch_bounds__(256) temporal_upsample_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int tl, int t,
                                                                int64_t frame16) {
  const int to = blockIdx.y; const int64_t n = blockIdx.z;
  const AxisTap at = axis_tap(to, tl, t);
  const uint4* a0 = in + (n * tl + at.i0) * frame16;
  const uint4* a1 = in + (n * tl + at.i1) * frame16;
  uint4* o = out + (n * t + to) * frame16;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < frame16; i += (int64_t)gridDim.x * 256) {
    const uint4 x = __ldg(a0 + i);
    if (at.l1 == 0.f) { o[i] = x; continue; }
    const uint4 y = __ldg(a1 + i);
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
    uint32_t r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float lo = at.l0 * __uint_as_float(xs[e] << 16) + at.l1 * __uint_as_float(ys[e] << 16);
      const float hi = at.l0 * __uint_as_float(xs[e] & 0xffff0000u) + at.l1 * __uint_as_float(ys[e] & 0xffff0000u);
      r[e] = cvt_bf16x2(lo, hi);
    }
    o[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

}  // namespace

int launch_temporal_upsample_bf16(const void* in, void* out, int n, int tl, int t, int hl, int wl, cudaStream_t stream) {
  const int64_t frame16 = (int64_t)hl * wl * HC * 2 / 16;
  const int bx = (int)std::min<int64_t>(cdiv(frame16, 256), 32);
  temporal_upsample_kernel<<<dim3((unsigned)bx, (unsigned)t, (unsigned)n), 256, 0, stream>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), tl, t, frame16);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

int launch_head_umma(const HeadArgs& a, cudaStream_t stream) {
  CLASFV_REQUIRE(a.g_dtype == CLASFV_BF16 && a.w2_bf16, "head_umma: bf16 lateral maps and bf16 W2 required");
  HeadGeom g;
  memset(&g, 0, sizeof(g));
  // w tile: 128 voxels unless the low-resolution columns they touch (plus the 2 bias rows) exceed the fixed K
  for (g.wt_vox = 128; g.wt_vox >= 16; g.wt_vox -= 16) {
    g.w_tiles = (a.w + g.wt_vox - 1) / g.wt_vox;
    if (g.w_tiles > MAX_WT) break;
    for (int l = 0; l < 4; ++l) g.nxmax[l] = 0;
    for (int l = 0; l < 4; ++l)
      for (int wt = 0; wt < g.w_tiles; ++wt) {
        const int v0 = wt * g.wt_vox, v1 = std::min(a.w, v0 + g.wt_vox) - 1;
        const AxisTap t0 = axis_tap(v0, a.wl[l], a.w), t1 = axis_tap(v1, a.wl[l], a.w);
        g.xlo[wt][l] = t0.i0; g.nx[wt][l] = t1.i1 - t0.i0 + 1;
        g.nxmax[l] = std::max(g.nxmax[l], g.nx[wt][l]);
      }
    if (g.nxmax[0] + g.nxmax[1] + g.nxmax[2] + g.nxmax[3] + 2 <= KSLABS * 64) break;
  }
  CLASFV_REQUIRE(g.wt_vox >= 16 && g.w_tiles <= MAX_WT, "head_umma: frame width %d is not supported", a.w);
  int k = 0, raw = 0;
  for (int l = 0; l < 4; ++l) { g.koff[l] = k; k += g.nxmax[l]; g.raw_off[l] = raw; raw += (a.tl[l] == a.t ? 2 : 4) * g.nxmax[l] * 128; }
  g.ktot = k + 2;                               // + the two bias rows
  g.kpad = round_up(g.ktot, 16);
  g.nslab = KSLABS;
  g.raw_stage_bytes = raw;
  CLASFV_REQUIRE(g.kpad <= KSLABS * 64, "head_umma: interpolation K too large (%d)", g.kpad);
  const int64_t total = (int64_t)a.n * a.t * a.h;
  CLASFV_REQUIRE(total < (1ll << 31), "head_umma: too many rows");
  g.total_rows = (int)total;
  // three R stages matter more than three raw stages (ncu: phase 1 otherwise idles a third of the time on R_EMPTY)
  const size_t base_bytes = 1024 + Smem::WA + (size_t)KSLABS * 16384, r_stage = (size_t)KSLABS * 8192, limit = 227 * 1024;
  // The row rate is (bulk-copy latency + phase 1 + hand-offs) / raw stages (clock64 timeline): raw stages
  // first, a third R stage only if it is free.
  g.nr = 2;
  for (g.nraw = NRAW; g.nraw > 2 && base_bytes + 2 * r_stage + (size_t)g.nraw * g.raw_stage_bytes > limit; --g.nraw) {}
  if (base_bytes + NR * r_stage + (size_t)g.nraw * g.raw_stage_bytes <= limit) g.nr = NR;
  const size_t smem = base_bytes + g.nr * r_stage + (size_t)g.nraw * g.raw_stage_bytes;
  CLASFV_REQUIRE(smem <= 227 * 1024, "head_umma: shared memory overflow (%zu bytes, W=%d)", smem, a.w);
  int dev = 0, sms = 0;
  CLASFV_CUDA(cudaGetDevice(&dev));
  CLASFV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = (int)std::min<int64_t>(total * g.w_tiles, (int64_t)(sms / g.w_tiles) * g.w_tiles);
  grid = std::max(grid / g.w_tiles, 1) * g.w_tiles;
  if (a.out_dtype == CLASFV_F32) {
    CLASFV_CUDA(allow_max_dynamic_smem(head_umma_kernel<float>));
    head_umma_kernel<float><<<grid, HU_THREADS, smem, stream>>>(a, g);
  } else {
    CLASFV_CUDA(allow_max_dynamic_smem(head_umma_kernel<__nv_bfloat16>));
    head_umma_kernel<__nv_bfloat16><<<grid, HU_THREADS, smem, stream>>>(a, g);
  }
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

}  // namespace clasfv
