// conv_umma.cu - Conv3d on channels-last bf16 activations as an implicit GEMM on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA.  sm_100a only.
//
// GEMM view of one convolution (all factorised convolutions of the R(2+1)D trunk: 1x3x3 spatial,
// 3x1x1 temporal, 1x1x1 strided downsample, and the decoder's 1x1x1 lateral projections):
//   D[pos, co] = sum over taps, ci of  X[pos shifted by tap, ci] * W[tap][co][ci]
//   M = output positions, tiled as a (bw x bh x bt x bn) box of the (Wo, Ho, To, N) output grid, 128 rows
//   N = output channels, one tile of BN <= 256 columns (multiple of 16)
//   K = taps x input channels, streamed in 64-channel (128-byte) slabs
// There is no im2col buffer: for every tap the A slab is ONE 5-D TMA box load of the input tensor at
// the tile origin shifted by the tap offset; TMA's out-of-bounds zero fill implements the zero padding.
// Strided convolutions read through per-parity views of the input (a tensor map whose base is offset
// by the tap's parity and whose strides are multiplied by the conv stride), so they are plain box loads
// too.  Both operands land in the canonical K-major SWIZZLE_128B layout that tcgen05 consumes directly.
//
// Warp roles (192 threads, persistent over tiles, one CTA per SM):
//   warp 0    TMA producer         : NSTAGE-deep ring of {A slab, B slab}, mbarrier full/empty
//   warp 1    MMA issuer           : one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x <=4 per slab,
//                                    tcgen05.commit releases the slab / publishes the accumulator
//   warps 2-5 epilogue             : tcgen05.ld the fp32 accumulator (lane = row), + folded-BN bias,
//                                    + residual, ReLU, convert, store channels-last
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include "internal.h"

#include <cstdio>
#include <cstring>

namespace clasfv {
namespace {

constexpr int TILE_M = 128;
constexpr int SLAB_K = 64;          // channels per K slab (128 bytes of bf16 = one swizzle row)
constexpr int UMMA_K = 16;
constexpr int UMMA_THREADS = 192;
constexpr int MAX_TAPS = 27;
constexpr int MAX_VIEWS = 4;
constexpr int A_SLAB_BYTES = TILE_M * SLAB_K * 2;   // 16 KiB

struct UmmaParams {
  CUtensorMap tmap_a[MAX_VIEWS];
  CUtensorMap tmap_b;
  int ntaps, kslabs, k16_last;
  int8_t tap_view[MAX_TAPS], tap_dw[MAX_TAPS], tap_dh[MAX_TAPS], tap_dt[MAX_TAPS];
  int tiles_w, tiles_h, tiles_t, tiles_b;   // M tiling of the output grid
  int tiles_n;                              // N tiling
  int bw, bh, bt, bb;                       // box extents (rows = bw*bh*bt*bb <= 128)
  int bn;                                   // N tile width
  int n, to, ho, wo, cout;
  int nstages, b_slab_bytes, tx_bytes, tmem_cols;
  uint32_t idesc;
  const float* bias; const void* residual; void* out;
  int relu, out_f32;
};

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && spin > (1u << 24)) { printf("clasfv conv_umma: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}

// ------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(UMMA_THREADS, 1) conv_umma_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A slabs][B slabs][barriers][tmem ptr]; base re-aligned to 1024 for the swizzle atoms
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + (uint32_t)p.nstages * A_SLAB_BYTES;
  const uint32_t bar_base = b_base + (uint32_t)p.nstages * (uint32_t)p.b_slab_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(p.nstages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * p.nstages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * p.nstages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * p.nstages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_t * p.tiles_b;
  const int total_tiles = m_tiles * p.tiles_n;

  if (warp == 0 && lane == 0) {
    for (int v = 0; v < MAX_VIEWS; ++v) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmap_a[v]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmap_b) : "memory");
    for (int s = 0; s < p.nstages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.tiles_n;
        int mt = tile / p.tiles_n;
        const int w0 = (mt % p.tiles_w) * p.bw; mt /= p.tiles_w;
        const int h0 = (mt % p.tiles_h) * p.bh; mt /= p.tiles_h;
        const int t0 = (mt % p.tiles_t) * p.bt; mt /= p.tiles_t;
        const int b0 = mt * p.bb;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const CUtensorMap* map = &p.tmap_a[p.tap_view[tap]];
          const int cw = w0 + p.tap_dw[tap], ch = h0 + p.tap_dh[tap], ct = t0 + p.tap_dt[tap];
          for (int ks = 0; ks < p.kslabs; ++ks) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(stage), (uint32_t)p.tx_bytes);
            tma_load_5d(a_base + (uint32_t)stage * A_SLAB_BYTES, map, full_bar(stage), ks * SLAB_K, cw, ch, ct, b0);
            tma_load_3d(b_base + (uint32_t)stage * (uint32_t)p.b_slab_bytes, &p.tmap_b, full_bar(stage), ks * SLAB_K, n_tile * p.bn, tap);
            if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1; const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.bn);
        uint32_t accumulate = 0;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int ks = 0; ks < p.kslabs; ++ks) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint64_t da = smem_desc_sw128(a_base + (uint32_t)stage * A_SLAB_BYTES);
            const uint64_t db = smem_desc_sw128(b_base + (uint32_t)stage * (uint32_t)p.b_slab_bytes);
            const int nk = (ks == p.kslabs - 1) ? p.k16_last : SLAB_K / UMMA_K;
            for (int k = 0; k < nk; ++k) {
              // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (>>4) address field
              tc_mma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), p.idesc, accumulate);
              accumulate = 1;
            }
            tc_commit(empty_bar(stage));          // slab reusable once these MMAs have read it
            if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
          }
        }
        tc_commit(tfull_bar(as));                 // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1; const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const int n_tile = tile % p.tiles_n;
      int mt = tile / p.tiles_n;
      const int w0 = (mt % p.tiles_w) * p.bw; mt /= p.tiles_w;
      const int h0 = (mt % p.tiles_h) * p.bh; mt /= p.tiles_h;
      const int t0 = (mt % p.tiles_t) * p.bt; mt /= p.tiles_t;
      const int b0 = mt * p.bb;
      int r = row;
      const int iw = r % p.bw; r /= p.bw;
      const int ih = r % p.bh; r /= p.bh;
      const int itt = r % p.bt; r /= p.bt;
      const int ib = r;
      const int ow = w0 + iw, oh = h0 + ih, ot = t0 + itt, ob = b0 + ib;
      const bool valid = ib < p.bb && ow < p.wo && oh < p.ho && ot < p.to && ob < p.n;
      const int c0 = n_tile * p.bn;
      const int64_t off = ((((int64_t)ob * p.to + ot) * p.ho + oh) * p.wo + ow) * p.cout + c0;

      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.bn);
      const int ncols = min(p.bn, p.cout - c0);   // ragged last N tile: columns past Cout are never stored
      for (int cc = 0; cc < ncols; cc += 16) {
        uint32_t acc[16];
        tc_ld16(taddr + (uint32_t)cc, acc);
        tc_wait_ld();
        if (valid) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(acc[i]);
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + cc) + i);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          }
          if (p.out_f32) {
            float* o = static_cast<float*>(p.out) + off + cc;
            if (p.residual) {
              const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + off + cc);
#pragma unroll
              for (int i = 0; i < 4; ++i) { const float4 r4 = rp[i]; v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w; }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
            __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + off + cc;
            if (p.residual) {
              const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + off + cc);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const uint4 r4 = rp[i];
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(hh[j]); v[8 * i + 2 * j] += f.x; v[8 * i + 2 * j + 1] += f.y; }
              }
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint4 w4;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * i + 0], v[8 * i + 1]);
              __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
              __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
              w4.x = *reinterpret_cast<uint32_t*>(&h0); w4.y = *reinterpret_cast<uint32_t*>(&h1);
              w4.z = *reinterpret_cast<uint32_t*>(&h2); w4.w = *reinterpret_cast<uint32_t*>(&h3);
              reinterpret_cast<uint4*>(o)[i] = w4;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));   // 4 arrivals (one per epilogue warp) free the accumulator
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

int encode_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return CLASFV_ECUDA; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
    return CLASFV_ECUDA;
  }
  return CLASFV_OK;
}

// floor division / modulo for the tap -> (parity, shift) split of strided convolutions
inline void split_offset(int off, int stride, int* q, int* r) {
  int qq = off / stride, rr = off % stride;
  if (rr < 0) { rr += stride; --qq; }
  *q = qq; *r = rr;
}

// Choose the (bw,bh,bt,bb) box of <= 128 output positions that wastes the fewest MMA rows.
void choose_box(int wo, int ho, int to, int n, int* bw, int* bh, int* bt, int* bb) {
  double best = -1.0; int best_rows = 0;
  for (int w = 1; w <= wo && w <= TILE_M; ++w)
    for (int h = 1; h <= ho && w * h <= TILE_M; ++h)
      for (int t = 1; t <= to && w * h * t <= TILE_M; ++t) {
        int b = TILE_M / (w * h * t);
        if (b > n) b = n;
        if (b < 1) continue;
        const int64_t tiles = cdiv(wo, w) * cdiv(ho, h) * cdiv(to, t) * cdiv(n, b);
        const double eff = (double)wo * ho * to * n / ((double)tiles * TILE_M);
        const int rows = w * h * t * b;
        // prefer higher efficiency, then fuller boxes, then wider rows (longer contiguous runs)
        if (eff > best + 1e-9 || (eff > best - 1e-9 && (rows > best_rows || (rows == best_rows && w > *bw)))) {
          best = eff; best_rows = rows; *bw = w; *bh = h; *bt = t; *bb = b;
        }
      }
}

}  // namespace

int umma_selftest_supported() { return get_encode_fn() != nullptr; }

int launch_conv_umma(const ConvArgs& a, int num_sms, cudaStream_t stream) {
  const ConvShape& s = a.s;
  CLASFV_REQUIRE(a.act_dtype == CLASFV_BF16, "conv_umma: bf16 activations only");
  CLASFV_REQUIRE(s.cin % 16 == 0 && s.cout % 16 == 0, "conv_umma: channel counts must be multiples of 16 (cin=%d cout=%d)", s.cin, s.cout);
  const int ntaps = s.kt * s.kh * s.kw;
  CLASFV_REQUIRE(ntaps <= MAX_TAPS, "conv_umma: too many filter taps (%d)", ntaps);
  CLASFV_REQUIRE(((uintptr_t)a.in & 15) == 0 && ((uintptr_t)a.weight & 15) == 0 && ((uintptr_t)a.out & 15) == 0, "conv_umma: pointers must be 16-byte aligned");

  UmmaParams p;
  memset(&p, 0, sizeof(p));
  // ---- N tiling
  // equal tiles of a multiple of 16 columns when the channel count allows it, else a ragged last tile
  const int min_tiles = (int)cdiv(s.cout, 256);
  int bn = 0;
  for (int nt = min_tiles; nt <= min_tiles + 2 && !bn; ++nt)
    if (s.cout % nt == 0 && (s.cout / nt) % 16 == 0) bn = s.cout / nt;
  if (!bn) bn = round_up((int)cdiv(s.cout, min_tiles), 16);
  CLASFV_REQUIRE(bn >= 16 && bn <= 256, "conv_umma: N tile overflow");
  p.bn = bn; p.tiles_n = (int)cdiv(s.cout, bn);
  // ---- M tiling
  choose_box(s.wo, s.ho, s.to, s.n, &p.bw, &p.bh, &p.bt, &p.bb);
  p.tiles_w = (int)cdiv(s.wo, p.bw); p.tiles_h = (int)cdiv(s.ho, p.bh); p.tiles_t = (int)cdiv(s.to, p.bt); p.tiles_b = (int)cdiv(s.n, p.bb);
  p.n = s.n; p.to = s.to; p.ho = s.ho; p.wo = s.wo; p.cout = s.cout;
  // ---- K
  p.ntaps = ntaps;
  p.kslabs = (int)cdiv(s.cin, SLAB_K);
  p.k16_last = (s.cin - (p.kslabs - 1) * SLAB_K) / UMMA_K;
  // ---- taps -> (parity view, coordinate shift)
  int view_key[MAX_VIEWS]; int nviews = 0;
  int view_rt[MAX_VIEWS], view_rh[MAX_VIEWS], view_rw[MAX_VIEWS];
  for (int tap = 0; tap < ntaps; ++tap) {
    const int kw = tap % s.kw, kh = (tap / s.kw) % s.kh, kt = tap / (s.kw * s.kh);
    int qt, rt, qh, rh, qw, rw;
    split_offset(kt - s.pt, s.st, &qt, &rt);
    split_offset(kh - s.ph, s.sh, &qh, &rh);
    split_offset(kw - s.pw, s.sw, &qw, &rw);
    const int key = (rt * 8 + rh) * 8 + rw;
    int v = -1;
    for (int i = 0; i < nviews; ++i) if (view_key[i] == key) v = i;
    if (v < 0) {
      CLASFV_REQUIRE(nviews < MAX_VIEWS, "conv_umma: more than %d stride parities", MAX_VIEWS);
      v = nviews++; view_key[v] = key; view_rt[v] = rt; view_rh[v] = rh; view_rw[v] = rw;
    }
    p.tap_view[tap] = (int8_t)v; p.tap_dw[tap] = (int8_t)qw; p.tap_dh[tap] = (int8_t)qh; p.tap_dt[tap] = (int8_t)qt;
  }
  const uint32_t boxa[5] = {(uint32_t)SLAB_K, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bt, (uint32_t)p.bb};
  for (int v = 0; v < MAX_VIEWS; ++v) {
    const int vv = v < nviews ? v : 0;          // unused slots alias view 0 so that prefetch.tensormap is harmless
    const int rt = view_rt[vv], rh = view_rh[vv], rw = view_rw[vv];
    const uint64_t dims[5] = {(uint64_t)s.cin, (uint64_t)cdiv(s.wi - rw, s.sw), (uint64_t)cdiv(s.hi - rh, s.sh),
                              (uint64_t)cdiv(s.ti - rt, s.st), (uint64_t)s.n};
    const uint64_t e = 2;
    const uint64_t strides[4] = {(uint64_t)s.sw * s.cin * e, (uint64_t)s.sh * s.wi * s.cin * e,
                                 (uint64_t)s.st * s.hi * s.wi * s.cin * e, (uint64_t)s.ti * s.hi * s.wi * s.cin * e};
    char* basep = (char*)a.in + (((int64_t)rt * s.hi + rh) * s.wi + rw) * s.cin * (int64_t)e;
    int rc = encode_map(&p.tmap_a[v], basep, 5, dims, strides, boxa);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)s.cin, (uint64_t)s.cout, (uint64_t)ntaps};
    const uint64_t strides[2] = {(uint64_t)s.cin * 2, (uint64_t)s.cout * s.cin * 2};
    const uint32_t box[3] = {(uint32_t)SLAB_K, (uint32_t)bn, 1};
    int rc = encode_map(&p.tmap_b, const_cast<void*>(a.weight), 3, dims, strides, box);
    if (rc) return rc;
  }
  // ---- pipeline sizing
  p.b_slab_bytes = bn * SLAB_K * 2;
  const int rows_a = p.bw * p.bh * p.bt * p.bb;
  p.tx_bytes = rows_a * SLAB_K * 2 + p.b_slab_bytes;
  const int stage_bytes = A_SLAB_BYTES + p.b_slab_bytes;
  int nstages = (200 * 1024) / stage_bytes;
  if (nstages > 6) nstages = 6;
  CLASFV_REQUIRE(nstages >= 2, "conv_umma: tile does not fit shared memory");
  p.nstages = nstages;
  int cols = 32;
  while (cols < 2 * bn) cols *= 2;
  p.tmem_cols = cols;
  // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3, M>>4 (cute::UMMA::InstrDescriptor)
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  p.bias = a.bias; p.residual = a.residual; p.out = a.out; p.relu = a.relu; p.out_f32 = a.out_f32;

  const size_t smem = (size_t)nstages * stage_bytes + 1024 /*align*/ + 8 * (2 * nstages + 4) + 16;
  CLASFV_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_t * p.tiles_b * p.tiles_n;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  conv_umma_kernel<<<grid, UMMA_THREADS, smem, stream>>>(p);
  CLASFV_CUDA(cudaGetLastError());
  return CLASFV_OK;
}

}  // namespace clasfv
