// conv_tc.cu -- dense convolution as implicit GEMM on the 5th-generation tensor cores (sm_100a).
//
// Replaces the ATen/cuDNN convolution behind layers.Conv2d (reference layers/misc.py:30-43, call
// sites modeling/backbone/resnet.py:258-353, fpn.py:35-36, rpn/rpn.py:87-93, make_layers.py:53,99) and
// the FrozenBatchNorm2d / ReLU / residual-add passes that follow it (layers/batch_norm.py:27-31,
// resnet.py:324-344), fused into the epilogue.
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[cout, tap, cin]
//
// GEMM view: M = output pixels (tile = th x tw = 128 pixels of one image), N = Cout tile (<= 256),
// K = taps x Cin in blocks of 64 channels of one filter tap.
//   * A (activations, NHWC bf16): one 4-D TMA box {64 ch, tw, th, 1} per (tap, channel block), its
//     (w, h) corner shifted by the tap offset.  Out-of-image rows/columns are zero-filled by the TMA
//     unit, so padding and partial tiles cost nothing and no im2col matrix exists anywhere.
//     The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle == the canonical
//     K-major SWIZZLE_128B operand layout of tcgen05.mma.
//   * B (weights, [Cout, taps, Cin] bf16): 3-D TMA box {64 ch, 1 tap, BN}.
//   * accumulators: fp32 in TMEM, two stages of BN columns, so the epilogue of tile i overlaps the
//     main loop of tile i+1.
//   * warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
//     tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld 32x32b -> scale/bias/residual/ReLU/mask
//     -> bf16/fp32 -> 128-bit global stores).  smem ring of 4-8 stages guarded by mbarriers;
//     tcgen05.commit releases ring slots and publishes accumulators.
//   * persistent: grid = min(#tiles, 148 SMs), static round-robin tile order with the Cout tile
//     fastest so co-resident CTAs share activation tiles through L2.
// The same kernel serves forward, data-gradient (flipped/transposed weights prepared by
// conv_prepare_dgrad_weights_kernel, output optionally strided for stride-2 1x1 layers) and, through
// the degenerate H = 1 view, the fully-connected layers of the ROI heads.
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>

#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mrb {

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// smem -> global tile store (bulk async group) and the fences around it
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 = 1024 B (8 rows x 128 B)
//   [46,48) version = 1 (Blackwell) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// ---- CTA-pair (cta_group::2) variants: two SMs of a cluster cooperate on one 256-row MMA; each CTA stages its own 128-row A
// tile and HALF of the B tile, the leader CTA (cluster rank 0) issues the MMAs and its commits are multicast to both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {   // same offset in CTA `rank` of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {     // arrives on `bar` of BOTH CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------- kernel
constexpr int kEpiWarps = 8;                  // epilogue warps: 2 per TMEM lane quadrant (latency hiding: each SMSP
                                              // gets 2 epilogue warps; a lone warp per scheduler stalls on every LDTM/LDS/SHFL)
constexpr int kCtrlWarps = 3;                 // warp 0: TMA producer, warp 1: MMA issuer, warp 2: epilogue DMA (tile-buffer loads/stores)
constexpr int kConvThreads = 32 * kCtrlWarps + 32 * kEpiWarps;
constexpr int kWgradThreads = 192;
constexpr int kTileM = 128;
constexpr int kBlockK = 64;                 // channels per k-block (128 B of bf16)
constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KB

struct ConvArgs {
  int tiles_total, tiles_n, tiles_w, tiles_h;  // tiles_total = batch * tiles_h * tiles_w * tiles_n
  int th, tw, Ho, Wo;
  int cin_blocks, kh, kw, pad_h, pad_w;
  int cout, bn, stages, relu, out_f32;
  long long out_n, out_h, out_w;  // output strides in elements (channel stride 1)
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;   // same indexing as out (bf16) unless res_up2
  int res_up2;                     // residual is [N, ceil(Ho/2), ceil(Wo/2), Cout]: read at (h>>1, w>>1) == nearest 2x upsample
  long long res_n, res_h, res_w;   // its strides in elements
  const __nv_bfloat16* relu_mask;  // same indexing as out: result zeroed where mask <= 0
  void* out;
  int tma_epi;                     // epilogue through TMA tile buffers (see conv_tc_kernel); else per-thread global access
  int tile_bufs;                   // 1, or 2 when both a residual and a mask tile are staged
  int tile_sets;                   // 2 or 3 buffer sets in a ring across tiles (store drain / input prefetch overlap)
  int tile_rows;                   // th * tw <= 128 rows of the M = 128 tile carry pixels (the rest is never stored)
  int grouped;                     // block-diagonal 64-channel super-groups: A channel offset = n_tile * 64, BN = 64
  int tiles_m;                     // batch * tiles_h * tiles_w; with CTA pairs tiles_total counts PAIRS of M tiles x tiles_n
};

// CTA pairs: `tile` indexes (pair of M tiles, N tile); CTA `rank` of the pair owns M tile 2 * pair + rank.  An odd tail gets a
// tile beyond the image: its loads are zero-filled and its stores dropped by the TMA unit / the validity test.
template <bool TWO>
__device__ __forceinline__ void tile_coords_t(const ConvArgs& a, int tile, int rank, int& n_tile, int& img, int& h0, int& w0) {
  n_tile = tile % a.tiles_n;
  int m = tile / a.tiles_n;
  if (TWO) {
    m = 2 * m + rank;
    if (m >= a.tiles_m) {
      img = 0; h0 = a.tiles_h * a.th; w0 = 0;
      return;
    }
  }
  const int wt = m % a.tiles_w; m /= a.tiles_w;
  const int ht = m % a.tiles_h;
  img = m / a.tiles_h;
  h0 = ht * a.th;
  w0 = wt * a.tw;
}

__device__ __forceinline__ void tile_coords(const ConvArgs& a, int tile, int& n_tile, int& img, int& h0, int& w0) {
  n_tile = tile % a.tiles_n;
  int m = tile / a.tiles_n;
  const int wt = m % a.tiles_w; m /= a.tiles_w;
  const int ht = m % a.tiles_h;
  img = m / a.tiles_h;
  h0 = ht * a.th;
  w0 = wt * a.tw;
}

// Slow-path epilogue for one 32-column chunk: fp32 outputs and ragged Cout tails (see conv_tc_kernel).
__device__ __noinline__ void epilogue_chunk32_slow(const ConvArgs& a, uint32_t taddr, int c0, bool valid, long long pix,
                                                   long long rpix, const float* sc, const float* bi, bool affine) {
  uint32_t v[32];
  tmem_ld32(taddr, v);
  tmem_ld_wait();
  if (!valid) return;
  const long long o = pix + c0;
  const long long ro = rpix + c0;
  const bool full = (c0 + 32 <= a.cout) && ((a.cout & 7) == 0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {          // 8 channels per step, all indices static
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[q * 8 + j]);
    if (affine) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[q * 8 + j], bi[q * 8 + j]);
    }
    if (full) {
      if (a.residual) {
        const uint4 rr = __ldg(reinterpret_cast<const uint4*>(a.residual + ro + q * 8));
        const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __bfloat1622float2(r2[j]);
          f[2 * j] += t.x; f[2 * j + 1] += t.y;
        }
      }
      if (a.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (a.relu_mask) {
        const uint4 rr = __ldg(reinterpret_cast<const uint4*>(a.relu_mask + o + q * 8));
        const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __bfloat1622float2(r2[j]);
          if (!(t.x > 0.f)) f[2 * j] = 0.f;
          if (!(t.y > 0.f)) f[2 * j + 1] = 0.f;
        }
      }
      if (a.out_f32) {
        float* dst = reinterpret_cast<float*>(a.out) + o + q * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
      } else {
        uint4 pk;
        __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) p2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.out) + o + q * 8) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + q * 8 + j;
        if (c < a.cout) {
          float x = f[j];
          if (a.residual) x += __bfloat162float(a.residual[ro + q * 8 + j]);
          if (a.relu) x = fmaxf(x, 0.f);
          if (a.relu_mask && !(__bfloat162float(a.relu_mask[o + q * 8 + j]) > 0.f)) x = 0.f;
          if (a.out_f32) reinterpret_cast<float*>(a.out)[o + q * 8 + j] = x;
          else reinterpret_cast<__nv_bfloat16*>(a.out)[o + q * 8 + j] = __float2bfloat16_rn(x);
        }
      }
    }
  }
}

template <bool TWO>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel_t(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
               const __grid_constant__ CUtensorMap map_mask, const ConvArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;            // CTA pair: 0 = leader (issues the MMAs)
  const uint32_t b_rows = TWO ? (uint32_t)a.bn >> 1 : (uint32_t)a.bn;   // rows of the B tile this CTA stages
  const uint32_t stage_bytes = kABytes + b_rows * 128u;
  const int tile0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tstep = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // [ring: stages x (A 16 KB + B)] [tile buffers: tile_bufs x ceil(BN/64) x 16 KB (TMA epilogue only)] [barriers] ...
  const uint32_t tile_base = smem_base + (uint32_t)a.stages * stage_bytes;
  const uint32_t tile_buf_bytes = (uint32_t)((a.bn + 63) >> 6) * kABytes;
  const uint32_t bar_base = tile_base + (a.tma_epi ? (uint32_t)(a.tile_bufs * a.tile_sets) * tile_buf_bytes : 0u);
  // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2], tile[3] (inputs of a set landed),
  // done[3] (all epilogue threads finished a set); then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (a.stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * a.stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * a.stages + 2 + s); };
  auto tile_bar = [&](int b) { return bar_base + 8u * (2 * a.stages + 4 + b); };
  auto done_bar = [&](int b) { return bar_base + 8u * (2 * a.stages + 7 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * a.stages + 10);   // keeps what follows 16-byte aligned
  const uint32_t a_tx = (uint32_t)a.tile_rows * 128u;               // bytes one A / tile-buffer box really transfers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_blocks = a.kh * a.kw * a.cin_blocks;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * a.bn) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), (TWO ? 2 : 1) * 32 * kEpiWarps); }
    for (int b = 0; b < 3; ++b) { mbar_init(tile_bar(b), 1); mbar_init(done_bar(b), 32 * kEpiWarps); }
    fence_barrier_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (a.tma_epi) {
      tma_prefetch_desc(&map_out);
      if (a.residual) tma_prefetch_desc(&map_res);
      if (a.relu_mask) tma_prefetch_desc(&map_mask);
    }
  }
  if (warp == 1) {
    if (TWO) tmem_alloc_2cta(tmem_slot, tmem_cols);
    else tmem_alloc(tmem_slot, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();       // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no
  // global memory and may overlap the tail of the previous kernel in the stream; from here on we need its results
  // (and it must be done reading any buffer we are about to overwrite).  Our own dependents may start their
  // prologue right away: they block at the same point until this grid has fully completed.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int s = 0; uint32_t phase = 0;
      for (int tile = tile0; tile < a.tiles_total; tile += tstep) {
        int n_tile, img, h0, w0;
        tile_coords_t<TWO>(a, tile, (int)rank, n_tile, img, h0, w0);
        for (int kb = 0; kb < k_blocks; ++kb) {
          const int tap = kb / a.cin_blocks, cb = kb - tap * a.cin_blocks;
          const int r = tap / a.kw, q = tap - r * a.kw;
          mbar_wait(empty_bar(s), phase ^ 1u);
          const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
          if (TWO) {
            // both CTAs' boxes complete on the LEADER's full barrier; the leader announces the bytes of the pair
            if (rank == 0) mbar_expect_tx(full_bar(s), 2u * (a_tx + b_rows * 128u));
            const uint32_t lead_bar = mapa_shared(full_bar(s), 0);
            tma_load_4d_2cta(sa, &map_a, lead_bar, cb * kBlockK, w0 + q - a.pad_w, h0 + r - a.pad_h, img);
            tma_load_3d_2cta(sa + kABytes, &map_b, lead_bar, cb * kBlockK, tap, n_tile * a.bn + (int)(rank * b_rows));
          } else {
            mbar_expect_tx(full_bar(s), a_tx + (uint32_t)a.bn * 128u);
            tma_load_4d(sa, &map_a, full_bar(s), (a.grouped ? n_tile * 64 : 0) + cb * kBlockK, w0 + q - a.pad_w, h0 + r - a.pad_h, img);
            tma_load_3d(sa + kABytes, &map_b, full_bar(s), cb * kBlockK, tap, n_tile * a.bn);
          }
          if (++s == a.stages) { s = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0 && rank == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
      // a,b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29) (M = 256 over the CTA pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.bn >> 3) << 17) |
                             ((uint32_t)((TWO ? 2 * kTileM : kTileM) >> 4) << 24);
      int s = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = tile0; tile < a.tiles_total; tile += tstep) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.bn);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(s), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
          const uint64_t da = umma_desc_k_sw128(sa), db = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {  // +32 B per K=16 step inside the 128 B swizzle span
            if (TWO) umma_bf16_2cta(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
            else umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          }
          if (TWO) {
            umma_commit_2cta(empty_bar(s));                         // frees the slot in both CTAs
            if (kb == k_blocks - 1) umma_commit_2cta(tfull_bar(acc));   // publishes both CTAs' accumulators
          } else {
            umma_commit(empty_bar(s));
            if (kb == k_blocks - 1) umma_commit(tfull_bar(acc));
          }
          if (++s == a.stages) { s = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp == 2) {
    // ================================ epilogue DMA ================================
    // One thread owns all bulk traffic of the TMA epilogue: it stages the residual / ReLU-mask tile of upcoming tiles
    // into the buffer-set ring (or just releases a set), and stores finished tiles.  The epilogue warps never wait for
    // a store to drain and never meet at a CTA-wide barrier: they hand a finished set over through done_bar.
    if (lane == 0 && a.tma_epi) {
      const bool has_res = a.residual != nullptr, has_mask = a.relu_mask != nullptr;
      const uint32_t set_bytes = (uint32_t)a.tile_bufs * tile_buf_bytes;          // one buffer set (T [+ M])
      const uint32_t m_off = (has_res && has_mask) ? tile_buf_bytes : 0u;         // the mask shares T when it is alone
      auto arm = [&](int tile, int b) {  // stage the inputs of `tile` into set b (or just release it)
        int n_tile, img, h0, w0;
        tile_coords_t<TWO>(a, tile, (int)rank, n_tile, img, h0, w0);
        const int cols = min(a.bn, a.cout - n_tile * a.bn), nb = (cols + 63) >> 6;
        const uint32_t bt = tile_base + (uint32_t)b * set_bytes;
        if (has_res || has_mask) {
          mbar_expect_tx(tile_bar(b), (uint32_t)nb * a_tx * (uint32_t)((has_res ? 1 : 0) + (has_mask ? 1 : 0)));
          for (int x = 0; x < nb; ++x) {
            if (has_res) tma_load_4d(bt + (uint32_t)x * kABytes, &map_res, tile_bar(b), n_tile * a.bn + x * 64, w0, h0, img);
            if (has_mask) tma_load_4d(bt + m_off + (uint32_t)x * kABytes, &map_mask, tile_bar(b), n_tile * a.bn + x * 64, w0, h0, img);
          }
        } else {
          mbar_arrive(tile_bar(b));
        }
      };
      const int sets = a.tile_sets;
      {
        int t = tile0;
        for (int j = 0; j < sets && t < a.tiles_total; ++j, t += tstep) arm(t, j);
      }
      int sidx = 0; uint32_t dphase = 0;
      for (int tile = tile0; tile < a.tiles_total; tile += tstep) {
        int n_tile, img, h0, w0;
        tile_coords_t<TWO>(a, tile, (int)rank, n_tile, img, h0, w0);
        mbar_wait(done_bar(sidx), dphase);           // every epilogue thread wrote (and proxy-fenced) its row of set sidx
        const uint32_t buf_t = tile_base + (uint32_t)sidx * set_bytes;
        const int cols = min(a.bn, a.cout - n_tile * a.bn), nb = (cols + 63) >> 6;
        for (int x = 0; x < nb; ++x) tma_store_4d(&map_out, buf_t + (uint32_t)x * kABytes, n_tile * a.bn + x * 64, w0, h0, img);
        bulk_commit();
        const long long nxt = (long long)tile + (long long)sets * tstep;
        if (nxt < a.tiles_total) {
          bulk_wait_read0();                         // the set has been read out: refill / release it
          arm((int)nxt, sidx);
        }
        if (++sidx == sets) { sidx = 0; dphase ^= 1u; }
      }
      bulk_wait_read0();                             // shared memory must outlive the last store's read
    }
  } else {
    // ================================ epilogue ================================
    // Everything below indexes registers with compile-time constants only: the 32-column chunk must never
    // spill to local memory (with ~200 KB of shared memory carved out, L1 is tiny and local traffic runs at
    // L2 latency).
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read (hardware: warp id % 4)
    const int grp = (warp - kCtrlWarps) >> 2;  // which of the kEpiWarps/4 warps of this quadrant: takes every
    constexpr int kGroups = kEpiWarps / 4;     // kGroups-th 32-column chunk
    const int row = quad * 32 + lane;          // accumulator row == pixel within the tile
    const int et = threadIdx.x - 32 * kCtrlWarps;   // 0..32*kEpiWarps-1 among the epilogue threads
    const int hh = row / a.tw, ww = row - hh * a.tw;
    float* s_aff = reinterpret_cast<float*>(smem_raw + (tmem_slot - smem_u32(smem_raw)) + 16);  // [2 acc][2][256]
    uint4* stg4 = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(s_aff) + 4096 + (warp - kCtrlWarps) * 2048);  // [32 rows][4 x 16 B]
    const bool has_scale = a.scale != nullptr, has_bias = a.bias != nullptr;
    const bool cout8 = (a.cout & 7) == 0;
    int acc = 0; uint32_t acc_phase = 0;
    if (a.tma_epi) {
      // ---- TMA epilogue.  The output tile lives in a swizzled shared-memory tile buffer T ([BN/64 boxes][128 rows]
      // [128 B], the layout TMA produces/consumes with SWIZZLE_128B): the residual (or, without one, the ReLU-backward
      // mask) of the tile is TMA-loaded INTO it ahead of time by the DMA warp, every thread updates its own row in
      // place and hands the set back (done_bar); the DMA warp TMA-stores it.  All global traffic of the epilogue is bulk
      // and asynchronous -- with per-thread loads a 1x1 layer with a residual ran at ~1.4 TB/s (too few bytes in flight
      // per SM) -- and the warps run tile after tile without a CTA-wide barrier (round 1 met at two named barriers per
      // tile and lane 0 waited for every store to drain: 2.4 us per 128x128 tile on the 1x1 layers).
      const bool has_res = a.residual != nullptr, has_mask = a.relu_mask != nullptr;
      const uint32_t set_bytes = (uint32_t)a.tile_bufs * tile_buf_bytes;
      const uint32_t m_off = (has_res && has_mask) ? tile_buf_bytes : 0u;
      const uint32_t row_off = (uint32_t)row * 128u, row_sw = (uint32_t)(row & 7);
      const int sets = a.tile_sets;
      int sidx = 0; uint32_t sphase = 0;
      for (int tile = tile0; tile < a.tiles_total; tile += tstep) {
        const int n_tile = tile % a.tiles_n;
        const uint32_t buf_t = tile_base + (uint32_t)sidx * set_bytes, buf_m = buf_t + m_off;
        mbar_wait(tile_bar(sidx), sphase);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * a.bn);
        for (int col = grp * 32; col < a.bn; col += 32 * kGroups) {
          const int cg0 = n_tile * a.bn + col;
          if (cg0 >= a.cout) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)col, v);
          tmem_ld_wait();
          const uint32_t box = (uint32_t)(col >> 6) * kABytes + row_off;
          const uint32_t u0 = (uint32_t)(col & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[q * 8 + j]);
            const int cg = cg0 + q * 8;                      // Cout % 8 == 0: an 8-group is entirely in or out
            if ((has_scale || has_bias) && cg < a.cout) {
              // per-channel affine straight from global memory (uniform address: one L1-resident line per warp)
              float4 s0 = make_float4(1.f, 1.f, 1.f, 1.f), s1 = s0, b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (has_scale) { s0 = __ldg(reinterpret_cast<const float4*>(a.scale + cg)); s1 = __ldg(reinterpret_cast<const float4*>(a.scale + cg + 4)); }
              if (has_bias) { b0 = __ldg(reinterpret_cast<const float4*>(a.bias + cg)); b1 = __ldg(reinterpret_cast<const float4*>(a.bias + cg + 4)); }
              f[0] = fmaf(f[0], s0.x, b0.x); f[1] = fmaf(f[1], s0.y, b0.y); f[2] = fmaf(f[2], s0.z, b0.z); f[3] = fmaf(f[3], s0.w, b0.w);
              f[4] = fmaf(f[4], s1.x, b1.x); f[5] = fmaf(f[5], s1.y, b1.y); f[6] = fmaf(f[6], s1.z, b1.z); f[7] = fmaf(f[7], s1.w, b1.w);
            }
            const uint32_t slot = box + (((u0 + (uint32_t)q) ^ row_sw) << 4);     // this thread's 16 B of its row
            if (has_res) {
              const uint4 rr = lds128(buf_t + slot);
              const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 t = __bfloat1622float2(r2[j]);
                f[2 * j] += t.x; f[2 * j + 1] += t.y;
              }
            }
            if (a.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (has_mask) {
              const uint4 mk = lds128(buf_m + slot);
              const __nv_bfloat162* m2 = reinterpret_cast<const __nv_bfloat162*>(&mk);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 t = __bfloat1622float2(m2[j]);
                if (!(t.x > 0.f)) f[2 * j] = 0.f;
                if (!(t.y > 0.f)) f[2 * j + 1] = 0.f;
              }
            }
            uint4 pk;
            __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) p2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            sts128(buf_t + slot, pk);
          }
        }
        tc_fence_before();
        if (TWO) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));     // the leader's MMA waits for both CTAs' epilogues
        else mbar_arrive(tempty_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        fence_proxy_async_smem();            // this thread's tile-buffer writes -> visible to the TMA (async proxy)
        mbar_arrive(done_bar(sidx));
        if (++sidx == sets) { sidx = 0; sphase ^= 1u; }
      }
    }
    for (int tile = tile0; tile < a.tiles_total && !a.tma_epi; tile += tstep) {
      int n_tile, img, h0, w0;
      tile_coords_t<TWO>(a, tile, (int)rank, n_tile, img, h0, w0);
      const int h = h0 + hh, w = w0 + ww;
      const bool valid = (row < a.tile_rows) && (h < a.Ho) && (w < a.Wo);
      const long long pix = (long long)img * a.out_n + (long long)h * a.out_h + (long long)w * a.out_w;
      const long long rpix = a.res_up2 ? (long long)img * a.res_n + (long long)(h >> 1) * a.res_h + (long long)(w >> 1) * a.res_w : pix;
      float* sc = s_aff + acc * 512;
      float* bi = sc + 256;
      if (has_scale || has_bias) {
        for (int c = et; c < a.bn; c += 32 * kEpiWarps) {
          const int cg = n_tile * a.bn + c;
          sc[c] = (has_scale && cg < a.cout) ? __ldg(a.scale + cg) : 1.f;
          bi[c] = (has_bias && cg < a.cout) ? __ldg(a.bias + cg) : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * a.bn);
      const unsigned vmask = __ballot_sync(0xffffffffu, valid);
      const bool fast = !a.out_f32 && cout8;
      for (int col = grp * 32; col < a.bn; col += 32 * kGroups) {
        const int c0 = n_tile * a.bn + col;
        if (c0 >= a.cout) break;  // warp-uniform
        if (fast && c0 + 32 <= a.cout) {
          // ---- 32 channels (64 B per pixel) per step.  Global traffic is staged through a per-warp 2 KB smem block
          // so that every warp-level load/store moves whole 32-byte sectors of 8 pixel rows (4 lanes x 16 B per row)
          // instead of 32 scattered 16-byte pieces; slot (row, seg) lives at seg ^ ((row >> 1) & 3): conflict-free
          // both for the thread-private accesses (row == lane) and for the coalesced ones.
          if (a.residual) {
            const long long my = rpix + c0;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r = it * 8 + (lane >> 2), sg = lane & 3;
              const long long off = __shfl_sync(0xffffffffu, my, r);
              if ((vmask >> r) & 1u) stg4[r * 4 + (sg ^ ((r >> 1) & 3))] = __ldg(reinterpret_cast<const uint4*>(a.residual + off + sg * 8));
            }
            __syncwarp();
          }
          {
            uint32_t v[32];
            tmem_ld32(t_row + (uint32_t)col, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[q * 8 + j]);
              if (has_scale || has_bias) {
                const float* scq = sc + col + q * 8;
                const float* biq = bi + col + q * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], scq[j], biq[j]);
              }
              uint4* slot = stg4 + lane * 4 + (q ^ ((lane >> 1) & 3));   // this thread's private 16 B
              if (a.residual) {
                const uint4 rr = *slot;
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 t = __bfloat1622float2(r2[j]);
                  f[2 * j] += t.x; f[2 * j + 1] += t.y;
                }
              }
              if (a.relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
              }
              uint4 pk;
              __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int j = 0; j < 4; ++j) p2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              *slot = pk;
            }
          }
          __syncwarp();
          {
            // coalesced write-out; the ReLU-backward mask (dgrad) is applied here, on the same (row, segment) the lane
            // stores, straight from its coalesced global read
            const long long my = pix + c0;
            __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(a.out);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r = it * 8 + (lane >> 2), sg = lane & 3;
              const long long off = __shfl_sync(0xffffffffu, my, r);
              if ((vmask >> r) & 1u) {
                uint4 val = stg4[r * 4 + (sg ^ ((r >> 1) & 3))];
                if (a.relu_mask) {
                  const uint4 mk = __ldg(reinterpret_cast<const uint4*>(a.relu_mask + off + sg * 8));
                  const __nv_bfloat162* m2 = reinterpret_cast<const __nv_bfloat162*>(&mk);
                  __nv_bfloat162* v2 = reinterpret_cast<__nv_bfloat162*>(&val);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 t = __bfloat1622float2(m2[j]);
                    float2 x = __bfloat1622float2(v2[j]);
                    if (!(t.x > 0.f)) x.x = 0.f;
                    if (!(t.y > 0.f)) x.y = 0.f;
                    v2[j] = __floats2bfloat162_rn(x.x, x.y);
                  }
                }
                *reinterpret_cast<uint4*>(outp + off + sg * 8) = val;
              }
            }
          }
          __syncwarp();
          continue;
        }
        // ---- fallback: per-thread 16-byte accesses (fp32 output, ragged Cout tails).
        // Kept out of line: it is rare and would otherwise double the instruction footprint of this loop.
        __syncwarp();
        epilogue_chunk32_slow(a, t_row + (uint32_t)col, c0, valid, pix, rpix, sc + col, bi + col, has_scale || has_bias);
      }
      tc_fence_before();
      if (TWO) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
      else mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();       // the leader's MMAs read the peer's shared memory: nobody leaves before both are done
  tc_fence_after();
  if (warp == 1) {
    if (TWO) tmem_dealloc_2cta(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}



// ------------------------------------------------------------------------------- weight gradient
// dW[co][tap][ci] = sum over pixels G[pixel, co] * X[pixel + tap, ci]      (fp32 result, split-K + red.add)
// GEMM view: M = Cout tile (128), N = Cin tile (<= 256), K = output pixels in blocks of 64 (a th x tw
// rectangle of one image).  Both operands are "MN-major" for the tensor core (channels are contiguous,
// pixels are the reduction axis):
//   A = G^T : two TMA boxes {64 co, tw, th, 1}  -> 2 x [64 pixel rows x 128 B], SWIZZLE_128B
//   B = X^T : BN/64 boxes   {64 ci, tw, th, 1}, corner shifted by the tap (zero-filled outside the image)
// UMMA descriptors: MN-major SWIZZLE_128B, LBO = 8 KB (distance between 64-channel groups),
// SBO = 1 KB (8 pixel rows); one K=16 step advances the start address by 16 rows = 2 KB.
// Work item = (co tile, ci tile, tap, K split); CTAs walk items round-robin; the accumulator tile is
// flushed with red.global.add.f32 (the reduction over splits) into the zero-initialised dW.
constexpr int kWgBlockK = 64;                       // pixels per k-block
constexpr int kWgGroupBytes = kWgBlockK * 128;      // one 64-channel group of one k-block: 8 KB

struct WgradArgs {
  int items_total, co_tiles, ci_tiles, taps, splits;
  int kh, kw, pad_h, pad_w;
  int th, tw, tiles_w, tiles_h, batch;   // pixel tiling of the OUTPUT (G) plane
  int kblocks_total, kblocks_per_split;
  int cin, cout, bn, stages;
  int grouped;         // block-diagonal super-groups: the Cin tile of a work item is the Cout tile's own 128 channels,
                       // dW is [cout][taps][128] (the caller extracts the diagonal blocks)
  float* dw;
  const float* scale;  // optional per-Cout factor (frozen-BN scale of the forward epilogue)
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  // [0,14) start>>4 | [16,30) LBO>>4 = 8 KB | [32,46) SBO>>4 = 1 KB | version 1 | SWIZZLE_128B
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(kWgGroupBytes >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void wg_item(const WgradArgs& a, int item, int& co_t, int& ci_t, int& tap, int& split) {
  split = item % a.splits; item /= a.splits;
  ci_t = item % a.ci_tiles; item /= a.ci_tiles;
  tap = item % a.taps;
  co_t = item / a.taps;
}

__global__ void __launch_bounds__(kWgradThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x, const WgradArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_groups = a.bn / 64;
  const uint32_t a_bytes = 2u * kWgGroupBytes;
  const uint32_t stage_bytes = a_bytes + (uint32_t)b_groups * kWgGroupBytes;
  const uint32_t bar_base = smem_base + (uint32_t)a.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (a.stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * a.stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * a.stages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * a.stages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * a.bn) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 128); }
    fence_barrier_init();
    tma_prefetch_desc(&map_g);
    tma_prefetch_desc(&map_x);
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no
  // global memory and may overlap the tail of the previous kernel in the stream; from here on we need its results
  // (and it must be done reading any buffer we are about to overwrite).  Our own dependents may start their
  // prologue right away: they block at the same point until this grid has fully completed.
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < a.items_total; item += gridDim.x) {
        int co_t, ci_t, tap, split;
        wg_item(a, item, co_t, ci_t, tap, split);
        const int r = tap / a.kw, q = tap - r * a.kw;
        const int kb0 = split * a.kblocks_per_split, kb1 = min(a.kblocks_total, kb0 + a.kblocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          int m = kb;
          const int wt = m % a.tiles_w; m /= a.tiles_w;
          const int ht = m % a.tiles_h;
          const int img = m / a.tiles_h;
          const int h0 = ht * a.th, w0 = wt * a.tw;
          mbar_wait(empty_bar(s), phase ^ 1u);
          mbar_expect_tx(full_bar(s), stage_bytes);
          const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
          tma_load_4d(sa, &map_g, full_bar(s), co_t * 128, w0, h0, img);
          tma_load_4d(sa + kWgGroupBytes, &map_g, full_bar(s), co_t * 128 + 64, w0, h0, img);
          for (int gi = 0; gi < b_groups; ++gi)
            tma_load_4d(sa + a_bytes + (uint32_t)gi * kWgGroupBytes, &map_x, full_bar(s), (a.grouped ? co_t * 128 : ci_t * a.bn) + gi * 64,
                        w0 + q - a.pad_w, h0 + r - a.pad_h, img);
          if (++s == a.stages) { s = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // a_major = b_major = 1 (MN-major): bits 15 and 16
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(a.bn >> 3) << 17) |
                             ((uint32_t)(kTileM >> 4) << 24);
      int s = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < a.items_total; item += gridDim.x) {
        int co_t, ci_t, tap, split;
        wg_item(a, item, co_t, ci_t, tap, split);
        const int kb0 = split * a.kblocks_per_split, kb1 = min(a.kblocks_total, kb0 + a.kblocks_per_split);
        if (kb1 <= kb0) continue;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.bn);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(s), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)s * stage_bytes;
          const uint64_t da = umma_desc_mn_sw128(sa), db = umma_desc_mn_sw128(sa + a_bytes);
#pragma unroll
          for (int k = 0; k < kWgBlockK / 16; ++k)  // 16 pixel rows = 2 KB per step
            umma_bf16(d_tmem, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc, (kb > kb0 || k) ? 1u : 0u);
          umma_commit(empty_bar(s));
          if (kb == kb1 - 1) umma_commit(tfull_bar(acc));
          if (++s == a.stages) { s = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < a.items_total; item += gridDim.x) {
      int co_t, ci_t, tap, split;
      wg_item(a, item, co_t, ci_t, tap, split);
      const int kb0 = split * a.kblocks_per_split, kb1 = min(a.kblocks_total, kb0 + a.kblocks_per_split);
      if (kb1 <= kb0) continue;
      const int co = co_t * 128 + row;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * a.bn);
      const int row_len = a.grouped ? 128 : a.cin;        // elements per (co, tap) row of dW
      float* __restrict__ dst = a.dw + ((size_t)co * a.taps + tap) * row_len;
      for (int col = 0; col < a.bn; col += 32) {
        const int c0 = ci_t * a.bn + col;
        if (c0 >= row_len) break;
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(t_row + (uint32_t)col, v);
        tmem_ld_wait();
        if (co >= a.cout) continue;
        const int nvalid = min(32, row_len - c0);
        const float sco = a.scale ? __ldg(a.scale + co) : 1.f;
        if (nvalid == 32 && (row_len & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            atomicAdd(reinterpret_cast<float4*>(dst + c0 + 4 * j),
                      make_float4(sco * __uint_as_float(v[4 * j]), sco * __uint_as_float(v[4 * j + 1]),
                                  sco * __uint_as_float(v[4 * j + 2]), sco * __uint_as_float(v[4 * j + 3])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)  // static indices only: v[] must stay in registers
            if (j < nvalid) atomicAdd(dst + c0 + j, sco * __uint_as_float(v[j]));
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------- weight prep
// dgrad weights: Wd[cin][kh-1-r][kw-1-q][cout] = W[cout][r][q][cin] * (scale ? scale[cout] : 1)
__global__ void conv_prepare_dgrad_weights_kernel(const __nv_bfloat16* __restrict__ w, const float* __restrict__ scale,
                                                  __nv_bfloat16* __restrict__ wd, int cout, int taps, int cin) {
  const int total = cout * taps * cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    // i indexes the destination [cin][tap'][cout] so that writes are coalesced
    const int co = i % cout;
    const int tp = (i / cout) % taps;
    const int ci = i / cout / taps;
    float v = __bfloat162float(w[((size_t)co * taps + (taps - 1 - tp)) * cin + ci]);
    if (scale) v *= scale[co];
    wd[i] = __float2bfloat16_rn(v);
  }
}

// one launch for many layers (descriptors travel as kernel parameters: no device-side table, no H2D copy)
constexpr int kPrepBatch = 40;
struct PrepBatch {
  const __nv_bfloat16* w[kPrepBatch];
  const float* scale[kPrepBatch];
  __nv_bfloat16* wd[kPrepBatch];
  int cout[kPrepBatch], taps[kPrepBatch], cin[kPrepBatch];
};
// Tiled version for the common case (Cin % 8 == 0 and Cout % 8 == 0): per tap a [Cout x Cin] -> [Cin x Cout]
// transpose through a 64 x 64 shared-memory tile, 16-byte global accesses on both sides.
struct PrepTiles {
  PrepBatch b;
  int tile_start[kPrepBatch + 1];   // exclusive prefix sum of taps * ceil(cout/64) * ceil(cin/64) per layer
};
__global__ void __launch_bounds__(256)
conv_prepare_dgrad_weights_tiled_kernel(PrepTiles t) {
  __shared__ __nv_bfloat16 tile[64][64 + 8];
  int l = 0;
  const int tidx = blockIdx.x;
  while (tidx >= t.tile_start[l + 1]) ++l;
  const int cout = t.b.cout[l], taps = t.b.taps[l], cin = t.b.cin[l];
  int rem = tidx - t.tile_start[l];
  const int ci_tiles = (cin + 63) >> 6, co_tiles = (cout + 63) >> 6;
  const int ci_t = rem % ci_tiles; rem /= ci_tiles;
  const int co_t = rem % co_tiles;
  const int tap = rem / co_tiles;
  const __nv_bfloat16* __restrict__ w = t.b.w[l];
  const float* __restrict__ scale = t.b.scale[l];
  __nv_bfloat16* __restrict__ wd = t.b.wd[l];
  const int chunk = threadIdx.x & 7, row = threadIdx.x >> 3;      // 8 x 16-byte chunks per 64-element row, 32 rows per pass
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int co = co_t * 64 + pass * 32 + row, ci = ci_t * 64 + chunk * 8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (co < cout && ci < cin) v = __ldg(reinterpret_cast<const uint4*>(w + ((size_t)co * taps + (taps - 1 - tap)) * cin + ci));
    *reinterpret_cast<uint4*>(&tile[pass * 32 + row][chunk * 8]) = v;
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int ci_l = pass * 32 + row, ci = ci_t * 64 + ci_l, co = co_t * 64 + chunk * 8;
    if (ci < cin && co < cout) {
      __align__(16) __nv_bfloat16 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float f = __bfloat162float(tile[chunk * 8 + j][ci_l]);
        if (scale) f *= __ldg(scale + co + j);
        o[j] = __float2bfloat16_rn(f);
      }
      *reinterpret_cast<uint4*>(wd + ((size_t)ci * taps + tap) * cout + co) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

__global__ void __launch_bounds__(256)
conv_prepare_dgrad_weights_batched_kernel(PrepBatch b) {
  const int l = blockIdx.y;
  const int cout = b.cout[l], taps = b.taps[l], cin = b.cin[l];
  const int total = cout * taps * cin;
  const __nv_bfloat16* __restrict__ w = b.w[l];
  const float* __restrict__ scale = b.scale[l];
  __nv_bfloat16* __restrict__ wd = b.wd[l];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i % cout;
    const int tp = (i / cout) % taps;
    const int ci = i / cout / taps;
    float v = __bfloat162float(w[((size_t)co * taps + (taps - 1 - tp)) * cin + ci]);
    if (scale) v *= scale[co];
    wd[i] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  });
  return fn;
}

static int encode_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                       const cuuint32_t* box) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return MRB_ERR_DRIVER;
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MRB_OK : MRB_ERR_BAD_ARG;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait in the kernels).  MRB_NO_PDL=1 in the
// environment falls back to a plain launch (A/B switch for measurements).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, int cluster, Args... args) {
  static const bool no_pdl = [] { const char* e = getenv("MRB_NO_PDL"); return e && e[0] == '1'; }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster > 1) {                 // thread-block cluster of `cluster` CTAs along x (the CTA pairs of conv_tc_kernel_t<true>)
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize, once per (kernel slot, device ordinal)
template <typename K>
static cudaError_t ensure_max_smem(K kernel, int slot) {
  static std::mutex mu;
  static bool done[4][64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  if (dev >= 0 && dev < 64 && done[slot][dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess && dev >= 0 && dev < 64) done[slot][dev] = true;
  return e;
}

struct ConvPlan {
  // logical GEMM-side geometry after the 1x1 flattening
  int batch, Hin, Win;        // TMA view of the input (already subsampled for stride 2)
  long long in_w, in_h, in_n; // input strides in elements
  int Ho, Wo;
  long long out_n, out_h, out_w;
};

// Launch one implicit-GEMM convolution.  `x` is the TMA-visible input [batch][Hin][Win][cin] with the
// given element strides; `w` is [cout][taps][cin].
static int conv_launch(const ConvPlan& pl, const void* x, const void* w, int cin, int cout, int kh, int kw, int pad_h, int pad_w,
                       const float* scale, const float* bias, const void* residual, const void* relu_mask, void* out,
                       int relu, int out_f32, cudaStream_t stream, int res_up2 = 0, int res_hh = 0, int res_ww = 0, bool grouped = false) {
  // grouped: block-diagonal 64-channel super-groups (cin == cout, % 64 == 0); `w` is [cout][taps][64]
  if (grouped && (cin != cout || (cin % 64) != 0)) return MRB_ERR_UNSUPPORTED;
  if (cin % 8 || ((uintptr_t)x & 15) || ((uintptr_t)w & 15) || ((uintptr_t)out & 15)) return MRB_ERR_UNSUPPORTED;
  if ((pl.in_w * 2) % 16 || (pl.in_h * 2) % 16 || (pl.in_n * 2) % 16) return MRB_ERR_UNSUPPORTED;
  if (residual && ((uintptr_t)residual & 15)) return MRB_ERR_UNSUPPORTED;
  if (relu_mask && ((uintptr_t)relu_mask & 15)) return MRB_ERR_UNSUPPORTED;
  // ---- plan: tile rectangle (th x tw <= 128 pixels of one image = the rows of the M = 128 tile that carry data),
  // N tile, epilogue mode.  The persistent grid runs ceil(tiles / 148) waves, so the plan that wins is usually the
  // one whose tile count lands just under a multiple of 148 -- a 10 x 12 tile (120 of 128 rows used) gives res4's
  // 50 x 84 planes 140 tiles = one wave where 4 x 32 gave 156 = two.  Cost = waves x per-tile time, per-tile time =
  // max(main loop, epilogue) with the main loop bound by the MMA issue rate or by the L2->smem operand stream.
  const int k_blocks_tile = kh * kw * (grouped ? 1 : ceil_div(cin, kBlockK));
  const bool tma_ok = !out_f32 && (cout % 8) == 0 && !res_up2 && ((pl.out_w | pl.out_h | pl.out_n) % 8) == 0;
  int th = 1, tw = 128, bn = 0;
  bool tma_epi = false;
  {
    int bn_max = (cout + 15) / 16 * 16;
    if (bn_max > 256) bn_max = 256;
    int cands[3] = {bn_max, 0, 0}, ncand = 1;
    if (bn_max == 256) { cands[1] = 128; cands[2] = 64; ncand = 3; }
    else if (bn_max == 128) { cands[1] = 64; ncand = 2; }
    if (grouped) { cands[0] = 64; ncand = 1; }
    double best_cost = -1;
    for (int t_h = 1; t_h <= 128; ++t_h) {
      if (pl.Ho == 1 && t_h > 1) break;
      int t_w = 128 / t_h;
      if (t_w > pl.Wo) t_w = pl.Wo > 0 ? pl.Wo : 1;
      if (t_h > pl.Ho) break;
      if (t_w > 256) t_w = 256;
      const int rows = t_h * t_w;
      const long long m_tiles = (long long)pl.batch * ceil_div(pl.Ho, t_h) * ceil_div(pl.Wo, t_w);
      // halo overhead of the A stream (re-fetched rows/columns of neighbouring tiles come from L2)
      const double halo = (double)(t_h + kh - 1) * (t_w + kw - 1) / (double)rows;
      for (int ci = 0; ci < ncand; ++ci) {
        const int cand = cands[ci];
        for (int mode = 0; mode < 2; ++mode) {        // 0: per-thread epilogue, 1: TMA tile-buffer epilogue
          if (mode == 1 && (!tma_ok || cand > 128)) continue;
          if (mode == 0 && tma_ok && cand <= 128 && !grouped) continue;      // narrow tiles: the TMA epilogue always wins
          const long long tiles = m_tiles * ceil_div(cout, cand);
          const double waves = (double)ceil_div(tiles, kNumSMs);
          const double mma = k_blocks_tile * (cand >= 256 ? 520.0 : cand >= 128 ? 270.0 : cand >= 64 ? 200.0 : 160.0);
          const double load = k_blocks_tile * (rows * 128.0 * (0.5 + 0.5 * halo) + cand * 128.0) / 96.0;
          const double epi = mode == 1 ? 500.0 + 4.0 * cand + (residual || relu_mask ? 300.0 : 0.0)
                                       : 600.0 + 9.0 * cand + (residual || relu_mask ? 4.0 * cand : 0.0);
          double t_tile = mma > load ? mma : load;
          if (epi > t_tile) t_tile = epi;
          // partially filled tiles waste MMA rows and L2->smem bytes: mild preference for full 128-row rectangles
          // (profiles/sweep_conv_r2_b.json: 8x16 beats 10x12 on the 200x336 planes at equal wave count)
          const double cost = waves * (t_tile + 150.0) * (1.0 + 0.02 * (halo - 1.0)) * (1.0 + 0.10 * (128 - rows) / 128.0);
          if (best_cost < 0 || cost < best_cost * 0.995) { best_cost = cost; th = t_h; tw = t_w; bn = cand; tma_epi = mode == 1; }
        }
      }
    }
  }
  // Long-K narrow tiles without epilogue inputs: the per-thread epilogue needs 16 KB of staging instead of 64 KB of tile
  // buffers, i.e. 6 ring stages instead of 4, and the main loop of these tiles is bound by (operand latency) / (ring depth):
  // measured 17.3 vs 19.4 us (res4 3x3), 10.6 vs 12.2 us (1x1 1024->256), profiles/sweep_conv_r2_e.json.  With a residual or
  // a ReLU mask to stage the TMA epilogue stays (per-thread loads of those were the 1.4 TB/s case of round 1).
  if (tma_epi && bn <= 128 && !residual && !relu_mask && k_blocks_tile >= 8 && !grouped) tma_epi = false;
  int sets_req = 0;
  if (const char* e = getenv("MRB_CONV_TILE")) {     // "th,tw,bn,epi[,sets]": measurement override (tools/bench_conv.py sweeps)
    int v[5] = {0, 0, 0, -1, 0};
    if (sscanf(e, "%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4]) >= 3 && v[0] > 0 && v[1] > 0 && v[0] * v[1] <= 128 &&
        v[2] >= 16 && v[2] <= 256 && v[2] % 16 == 0 && !(pl.Ho == 1 && v[0] > 1) && !grouped) {
      th = v[0]; tw = v[1]; bn = v[2];
      if (v[3] >= 0) tma_epi = v[3] && tma_ok && bn <= 128;
      else tma_epi = tma_ok && bn <= 128;
      sets_req = v[4];
    }
  }
  const int tile_rows = th * tw;
  int tile_bufs = (tma_epi && residual && relu_mask) ? 2 : 1;
  // buffer-set ring of the TMA epilogue: 3 sets when the tiles have inputs to prefetch (residual / mask) and the operand
  // ring keeps >= 3 stages, else 2
  int tile_sets = 2;
  const size_t fixed0 = 8 * (2 * 8 + 10) + 16 + 4096 + 1024;
  if (tma_epi) {
    const size_t set_b = (size_t)tile_bufs * ceil_div(bn, 64) * kABytes;
    auto stages_with = [&](int sets) { return (int)(((long long)227 * 1024 - (long long)fixed0 - (long long)(sets * set_b)) / (kABytes + bn * 128)); };
    if ((residual || relu_mask) && stages_with(3) >= 3) tile_sets = 3;
    if (sets_req >= 1 && sets_req <= 3) tile_sets = sets_req;
    if (stages_with(tile_sets) < 2) {
      if (tile_sets == 3 && stages_with(2) >= 2) tile_sets = 2;
      else { tma_epi = false; tile_bufs = 1; }
    }
  }
  ConvArgs a;
  a.th = th; a.tw = tw; a.Ho = pl.Ho; a.Wo = pl.Wo;
  a.tiles_h = ceil_div(pl.Ho, th); a.tiles_w = ceil_div(pl.Wo, tw); a.tiles_n = ceil_div(cout, bn);
  const long long tiles_m = (long long)pl.batch * a.tiles_h * a.tiles_w;
  const long long tiles = tiles_m * a.tiles_n;
  if (tiles <= 0 || tiles >= (1ll << 31)) return tiles == 0 ? MRB_OK : MRB_ERR_UNSUPPORTED;
  a.tiles_total = (int)tiles;
  a.tiles_m = (int)tiles_m;
  // CTA pairs (cta_group::2): two SMs share one B tile (each stages half of it) for a 256-row MMA: 2/3 of the L2 -> SM operand
  // bytes per FLOP at BN = 256.  MRB_CONV_2CTA=1 enables it for every launch it supports, =0 disables it.
  static const int two_mode = [] { const char* e = getenv("MRB_CONV_2CTA"); return e ? atoi(e) : 0; }();
  const bool two = two_mode == 1 && !grouped && bn >= 32 && (bn % 16) == 0 && tiles_m >= 2;
  if (two) a.tiles_total = (int)(((tiles_m + 1) / 2) * a.tiles_n);
  a.cin_blocks = grouped ? 1 : ceil_div(cin, kBlockK); a.kh = kh; a.kw = kw; a.pad_h = pad_h; a.pad_w = pad_w;
  a.cout = cout; a.bn = bn; a.relu = relu; a.out_f32 = out_f32;
  a.out_n = pl.out_n; a.out_h = pl.out_h; a.out_w = pl.out_w;
  a.scale = scale; a.bias = bias; a.residual = (const __nv_bfloat16*)residual; a.relu_mask = (const __nv_bfloat16*)relu_mask;
  a.out = out;
  a.res_up2 = res_up2;
  a.res_w = cout; a.res_h = (long long)res_ww * cout; a.res_n = (long long)res_hh * res_ww * cout;
  a.tma_epi = tma_epi ? 1 : 0;
  a.tile_bufs = tile_bufs;
  a.tile_sets = tile_sets;
  a.tile_rows = tile_rows;
  a.grouped = grouped ? 1 : 0;
  const uint32_t stage_bytes = kABytes + (two ? bn / 2 : bn) * 128;
  // fixed part: barriers + TMEM slot + scale/bias (4 KB) + alignment slack, plus either the tile buffers or the
  // per-warp staging blocks of the per-thread epilogue
  const size_t epi_bytes = tma_epi ? (size_t)tile_sets * tile_bufs * ceil_div(bn, 64) * kABytes : (size_t)2048 * kEpiWarps;
  const size_t fixed = fixed0 + epi_bytes;
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return MRB_ERR_UNSUPPORTED;
  a.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed;

  CUtensorMap map_a, map_b, map_out, map_res, map_mask;
  if (tma_epi) {
    // output-side tensors [batch][Ho][Wo][cout] with the plan's element strides; box = one 64-channel slice of a tile
    cuuint64_t dims[4] = {(cuuint64_t)cout, (cuuint64_t)pl.Wo, (cuuint64_t)pl.Ho, (cuuint64_t)pl.batch};
    cuuint64_t strides[3] = {(cuuint64_t)pl.out_w * 2, (cuuint64_t)pl.out_h * 2, (cuuint64_t)pl.out_n * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)tw, (cuuint32_t)th, 1};
    int rc = encode_bf16(&map_out, out, 4, dims, strides, box);
    if (rc) return rc;
    map_res = map_out; map_mask = map_out;
    if (residual && (rc = encode_bf16(&map_res, residual, 4, dims, strides, box))) return rc;
    if (relu_mask && (rc = encode_bf16(&map_mask, relu_mask, 4, dims, strides, box))) return rc;
  } else {
    memset(&map_out, 0, sizeof(map_out)); map_res = map_out; map_mask = map_out;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)pl.Win, (cuuint64_t)pl.Hin, (cuuint64_t)pl.batch};
    cuuint64_t strides[3] = {(cuuint64_t)pl.in_w * 2, (cuuint64_t)pl.in_h * 2, (cuuint64_t)pl.in_n * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)tw, (cuuint32_t)th, 1};
    int rc = encode_bf16(&map_a, x, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const int taps = kh * kw;
    const int wk = grouped ? 64 : cin;          // K extent of one weight row
    cuuint64_t dims[3] = {(cuuint64_t)wk, (cuuint64_t)taps, (cuuint64_t)cout};
    cuuint64_t strides[2] = {(cuuint64_t)wk * 2, (cuuint64_t)taps * wk * 2};
    cuuint32_t box[3] = {(cuuint32_t)kBlockK, 1, (cuuint32_t)(two ? bn / 2 : bn)};
    int rc = encode_bf16(&map_b, w, 3, dims, strides, box);
    if (rc) return rc;
  }
  // the attribute is per device: one process may drive several GPUs (cuda:0 then cuda:1)
  if (two) {
    MRB_CUDA_TRY(ensure_max_smem(conv_tc_kernel_t<true>, 2));
    const int pairs = a.tiles_total < kNumSMs / 2 ? a.tiles_total : kNumSMs / 2;
    MRB_CUDA_TRY(launch_pdl(conv_tc_kernel_t<true>, 2 * pairs, kConvThreads, smem, stream, 2, map_a, map_b, map_out, map_res, map_mask, a));
    return MRB_OK;
  }
  MRB_CUDA_TRY(ensure_max_smem(conv_tc_kernel_t<false>, 0));
  const int grid = a.tiles_total < kNumSMs ? a.tiles_total : kNumSMs;
  MRB_CUDA_TRY(launch_pdl(conv_tc_kernel_t<false>, grid, kConvThreads, smem, stream, 1, map_a, map_b, map_out, map_res, map_mask, a));
  return MRB_OK;
}

static inline int pad_w_of(const mrb_conv_params* p) { return (p->flags & MRB_CONV_PAD_W) ? p->pad_w : p->pad; }
static inline bool has_pitch(const long long* v) { return v[0] || v[1] || v[2]; }

static int conv_check(const mrb_conv_params* p) {
  if (!p) return MRB_ERR_BAD_ARG;
  if (p->batch < 0 || p->height <= 0 || p->width <= 0 || p->cin <= 0 || p->cout <= 0 || p->kh <= 0 || p->kw <= 0 ||
      p->stride <= 0 || p->pad < 0)
    return MRB_ERR_BAD_ARG;
  if (p->stride != 1 && !(p->stride == 2 && p->kh == 1 && p->kw == 1 && p->pad == 0 && pad_w_of(p) == 0)) return MRB_ERR_UNSUPPORTED;
  if ((p->flags & MRB_CONV_PAD_W) && p->pad_w < 0) return MRB_ERR_BAD_ARG;
  for (int i = 0; i < 3; ++i)
    if (p->x_pitch[i] < 0 || p->y_pitch[i] < 0) return MRB_ERR_BAD_ARG;
  if ((has_pitch(p->x_pitch) || has_pitch(p->y_pitch)) && p->stride != 1) return MRB_ERR_UNSUPPORTED;
  return MRB_OK;
}

}  // namespace mrb
using namespace mrb;

static int conv2d_fwd_impl(const mrb_conv_params* p, const void* input, const void* weight, const float* scale,
                           const float* bias, const void* residual, int residual_up2, void* output, mrb_stream_t stream);

MRB_API int mrb_conv2d_fwd(const mrb_conv_params* p, const void* input, const void* weight, const float* scale,
                           const float* bias, const void* residual, void* output, mrb_stream_t stream) {
  return conv2d_fwd_impl(p, input, weight, scale, bias, residual, 0, output, stream);
}

MRB_API int mrb_conv2d_fwd_up2(const mrb_conv_params* p, const void* input, const void* weight, const float* scale,
                               const float* bias, const void* residual_half, void* output, mrb_stream_t stream) {
  return conv2d_fwd_impl(p, input, weight, scale, bias, residual_half, 1, output, stream);
}

static int conv2d_fwd_impl(const mrb_conv_params* p, const void* input, const void* weight, const float* scale,
                           const float* bias, const void* residual, int residual_up2, void* output, mrb_stream_t stream) {
  int rc = conv_check(p);
  if (rc) return rc;
  if (p->batch == 0) return MRB_OK;
  if (!input || !weight || !output) return MRB_ERR_BAD_ARG;
  const int pad_w = pad_w_of(p);
  int Ho = (p->height + 2 * p->pad - p->kh) / p->stride + 1, Wo = (p->width + 2 * pad_w - p->kw) / p->stride + 1;
  if (Ho <= 0 || Wo <= 0) return MRB_ERR_BAD_ARG;
  if (p->out_h > 0 || p->out_w > 0) {
    if (p->out_h <= 0 || p->out_w <= 0 || p->out_h > Ho || p->out_w > Wo || p->stride != 1) return MRB_ERR_BAD_ARG;
    Ho = p->out_h; Wo = p->out_w;
  }
  ConvPlan pl;
  const long long C = p->cin, Co = p->cout;
  if (residual_up2 && (!residual || p->stride != 1)) return MRB_ERR_BAD_ARG;
  const bool pitched = has_pitch(p->x_pitch) || has_pitch(p->y_pitch);
  if (pitched && residual_up2) return MRB_ERR_UNSUPPORTED;
  if (p->kh == 1 && p->kw == 1 && p->stride == 1 && p->pad == 0 && pad_w == 0 && p->out_h == 0 && !residual_up2 && !pitched) {
    // pure GEMM: all pixels of the batch on one axis, zero tile waste
    pl.batch = 1; pl.Hin = 1; pl.Win = p->batch * p->height * p->width;
    pl.in_w = C; pl.in_h = (long long)pl.Win * C; pl.in_n = pl.in_h;
    pl.Ho = 1; pl.Wo = pl.Win;
    pl.out_w = Co; pl.out_h = (long long)pl.Wo * Co; pl.out_n = pl.out_h;
  } else {
    pl.batch = p->batch; pl.Hin = (p->stride == 2 ? Ho : p->height); pl.Win = (p->stride == 2 ? Wo : p->width);
    pl.in_w = C * p->stride; pl.in_h = (long long)p->width * C * p->stride; pl.in_n = (long long)p->height * p->width * C;
    pl.Ho = Ho; pl.Wo = Wo;
    pl.out_w = Co; pl.out_h = (long long)Wo * Co; pl.out_n = (long long)Ho * Wo * Co;
    if (has_pitch(p->x_pitch)) { pl.in_n = p->x_pitch[0]; pl.in_h = p->x_pitch[1]; pl.in_w = p->x_pitch[2]; }
    if (has_pitch(p->y_pitch)) { pl.out_n = p->y_pitch[0]; pl.out_h = p->y_pitch[1]; pl.out_w = p->y_pitch[2]; }
  }
  const bool grouped = (p->flags & MRB_CONV_GROUPED64) != 0;
  if (grouped && (p->stride != 1 || residual_up2)) return MRB_ERR_UNSUPPORTED;
  return conv_launch(pl, input, weight, p->cin, p->cout, p->kh, p->kw, p->pad, pad_w, scale, bias, residual, nullptr, output,
                     p->relu, p->out_dtype == MRB_F32, (cudaStream_t)stream, residual_up2, (Ho + 1) / 2, (Wo + 1) / 2, grouped);
}

MRB_API size_t mrb_conv2d_dgrad_workspace_bytes(const mrb_conv_params* p) {
  if (conv_check(p)) return 0;
  return ((size_t)p->cout * p->kh * p->kw * p->cin * 2 + 255) & ~(size_t)255;
}

MRB_API int mrb_conv2d_prepare_dgrad_weights(int num_layers, const void* const* weights_host, const float* const* scales_host,
                                             void* const* prepared_host, const int* couts_host, const int* taps_host,
                                             const int* cins_host, mrb_stream_t stream_) {
  if (num_layers < 0 || (num_layers && (!weights_host || !prepared_host || !couts_host || !taps_host || !cins_host)))
    return MRB_ERR_BAD_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  for (int l0 = 0; l0 < num_layers; l0 += kPrepBatch) {
    PrepTiles pt;
    PrepBatch& b = pt.b;
    const int n = min(kPrepBatch, num_layers - l0);
    int max_total = 0;
    bool tiled = true;
    long long tiles = 0;
    for (int i = 0; i < kPrepBatch; ++i) {
      const int l = l0 + (i < n ? i : 0);
      if (!weights_host[l] || !prepared_host[l] || couts_host[l] <= 0 || taps_host[l] <= 0 || cins_host[l] <= 0) return MRB_ERR_BAD_ARG;
      b.w[i] = (const __nv_bfloat16*)weights_host[l];
      b.scale[i] = scales_host ? scales_host[l] : nullptr;
      b.wd[i] = (__nv_bfloat16*)prepared_host[l];
      b.cout[i] = couts_host[l]; b.taps[i] = taps_host[l]; b.cin[i] = cins_host[l];
      max_total = max(max_total, couts_host[l] * taps_host[l] * cins_host[l]);
      if ((couts_host[l] | cins_host[l]) & 7 || ((uintptr_t)weights_host[l] | (uintptr_t)prepared_host[l]) & 15) tiled = false;
      if (i < n) {
        pt.tile_start[i] = (int)tiles;
        tiles += (long long)taps_host[l] * ceil_div(couts_host[l], 64) * ceil_div(cins_host[l], 64);
      }
    }
    for (int i = n; i <= kPrepBatch; ++i) pt.tile_start[i] = (int)tiles;
    if (tiled && tiles > 0 && tiles < (1ll << 30)) {
      conv_prepare_dgrad_weights_tiled_kernel<<<(unsigned)tiles, 256, 0, stream>>>(pt);
      MRB_LAUNCH_CHECK();
      continue;
    }
    dim3 grid(min(ceil_div(max_total, 256 * 4), 64), n);
    conv_prepare_dgrad_weights_batched_kernel<<<grid, 256, 0, stream>>>(b);
    MRB_LAUNCH_CHECK();
  }
  return MRB_OK;
}

static int conv2d_dgrad_impl(const mrb_conv_params* p, const void* grad_output, const __nv_bfloat16* wd, const void* add,
                             const void* relu_mask, void* grad_input, cudaStream_t stream);

MRB_API int mrb_conv2d_dgrad_prepared(const mrb_conv_params* p, const void* grad_output, const void* prepared_weight,
                                      const void* add, const void* relu_mask, void* grad_input, mrb_stream_t stream) {
  int rc = conv_check(p);
  if (rc) return rc;
  if (p->batch == 0) return MRB_OK;
  if (!grad_output || !prepared_weight || !grad_input) return MRB_ERR_BAD_ARG;
  if (p->out_h || p->out_w) return MRB_ERR_UNSUPPORTED;
  return conv2d_dgrad_impl(p, grad_output, (const __nv_bfloat16*)prepared_weight, add, relu_mask, grad_input, (cudaStream_t)stream);
}

MRB_API int mrb_conv2d_dgrad(const mrb_conv_params* p, const void* grad_output, const void* weight, const float* scale,
                             const void* add, const void* relu_mask, void* grad_input, void* workspace,
                             size_t workspace_bytes, mrb_stream_t stream_) {
  int rc = conv_check(p);
  if (rc) return rc;
  if (p->batch == 0) return MRB_OK;
  if (!grad_output || !weight || !grad_input || !workspace) return MRB_ERR_BAD_ARG;
  if (p->out_h || p->out_w) return MRB_ERR_UNSUPPORTED;
  if (p->flags & MRB_CONV_GROUPED64) return MRB_ERR_UNSUPPORTED;   // grouped: use mrb_conv2d_dgrad_prepared
  if (workspace_bytes < mrb_conv2d_dgrad_workspace_bytes(p)) return MRB_ERR_WORKSPACE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int taps = p->kh * p->kw;
  __nv_bfloat16* wd = (__nv_bfloat16*)workspace;
  {
    const int total = p->cout * taps * p->cin;
    conv_prepare_dgrad_weights_kernel<<<grid_for(total, 256, 8, 4), 256, 0, stream>>>((const __nv_bfloat16*)weight, scale, wd,
                                                                                     p->cout, taps, p->cin);
    MRB_LAUNCH_CHECK();
  }
  return conv2d_dgrad_impl(p, grad_output, wd, add, relu_mask, grad_input, stream);
}

static int conv2d_dgrad_impl(const mrb_conv_params* p, const void* grad_output, const __nv_bfloat16* wd, const void* add,
                             const void* relu_mask, void* grad_input, cudaStream_t stream) {
  const int pad_w = pad_w_of(p);
  const int Ho = (p->height + 2 * p->pad - p->kh) / p->stride + 1, Wo = (p->width + 2 * pad_w - p->kw) / p->stride + 1;
  const bool pitched = has_pitch(p->x_pitch) || has_pitch(p->y_pitch);
  // dgrad == forward conv of grad_output [N,Ho,Wo,Cout] with Wd [Cin][taps][Cout], pad' = k - 1 - pad
  ConvPlan pl;
  const long long Ci = p->cin, Co = p->cout;
  if ((p->flags & MRB_CONV_GROUPED64) && p->stride != 1) return MRB_ERR_UNSUPPORTED;
  if (p->stride == 2) {
    // 1x1 stride 2: grad_input[2h, 2w] = Wd . grad_output[h, w]; every other position is zero
    // `add` must be grad_input itself (accumulate a second stride-2 branch in place, e.g. conv1 + downsample of a
    // stage's first bottleneck): then the zero fill is skipped.  The mask only needs the even positions.
    if (add && add != grad_input) return MRB_ERR_UNSUPPORTED;
    if (add && p->out_dtype == MRB_F32) return MRB_ERR_UNSUPPORTED;
    if (!add) {
      const size_t bytes = (size_t)p->batch * p->height * p->width * p->cin * (p->out_dtype == MRB_F32 ? 4 : 2);
      MRB_CUDA_TRY(cudaMemsetAsync(grad_input, 0, bytes, stream));
    }
    pl.batch = p->batch; pl.Hin = Ho; pl.Win = Wo;
    pl.in_w = Co; pl.in_h = (long long)Wo * Co; pl.in_n = (long long)Ho * Wo * Co;
    pl.Ho = Ho; pl.Wo = Wo;
    pl.out_w = 2 * Ci; pl.out_h = 2ll * p->width * Ci; pl.out_n = (long long)p->height * p->width * Ci;
    return conv_launch(pl, grad_output, wd, p->cout, p->cin, 1, 1, 0, 0, nullptr, nullptr, add, relu_mask, grad_input, 0,
                       p->out_dtype == MRB_F32, stream);
  }
  if (p->kh == 1 && p->kw == 1 && p->pad == 0 && pad_w == 0 && !pitched) {
    pl.batch = 1; pl.Hin = 1; pl.Win = p->batch * Ho * Wo;
    pl.in_w = Co; pl.in_h = (long long)pl.Win * Co; pl.in_n = pl.in_h;
    pl.Ho = 1; pl.Wo = pl.Win;
    pl.out_w = Ci; pl.out_h = (long long)pl.Wo * Ci; pl.out_n = pl.out_h;
  } else {
    pl.batch = p->batch; pl.Hin = Ho; pl.Win = Wo;
    pl.in_w = Co; pl.in_h = (long long)Wo * Co; pl.in_n = (long long)Ho * Wo * Co;
    pl.Ho = p->height; pl.Wo = p->width;
    pl.out_w = Ci; pl.out_h = (long long)p->width * Ci; pl.out_n = (long long)p->height * p->width * Ci;
    if (has_pitch(p->y_pitch)) { pl.in_n = p->y_pitch[0]; pl.in_h = p->y_pitch[1]; pl.in_w = p->y_pitch[2]; }
    if (has_pitch(p->x_pitch)) { pl.out_n = p->x_pitch[0]; pl.out_h = p->x_pitch[1]; pl.out_w = p->x_pitch[2]; }
  }
  return conv_launch(pl, grad_output, wd, p->cout, p->cin, p->kh, p->kw, p->kh - 1 - p->pad, p->kw - 1 - pad_w, nullptr, nullptr, add, relu_mask,
                     grad_input, 0, p->out_dtype == MRB_F32, stream, 0, 0, 0, (p->flags & MRB_CONV_GROUPED64) != 0);
}

static int conv_wgrad_impl(const mrb_conv_params* p, const void* input, const void* grad_output, const float* scale,
                           float* grad_weight, bool accumulate, mrb_stream_t stream_) {
  int rc = conv_check(p);
  if (rc) return rc;
  if (p->out_h || p->out_w) return MRB_ERR_UNSUPPORTED;
  if (!grad_weight) return MRB_ERR_BAD_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int taps = p->kh * p->kw;
  const bool grouped = (p->flags & MRB_CONV_GROUPED64) != 0;
  if (grouped && (p->cin != p->cout || p->cout % 128 || p->stride != 1)) return MRB_ERR_UNSUPPORTED;
  if (!accumulate) MRB_CUDA_TRY(cudaMemsetAsync(grad_weight, 0, (size_t)p->cout * taps * (grouped ? 128 : p->cin) * sizeof(float), stream));
  if (p->batch == 0) return MRB_OK;
  if (!input || !grad_output) return MRB_ERR_BAD_ARG;
  if (p->cin % 8 || p->cout % 8 || ((uintptr_t)input & 15) || ((uintptr_t)grad_output & 15) || ((uintptr_t)grad_weight & 15))
    return MRB_ERR_UNSUPPORTED;
  const int pad_w = pad_w_of(p);
  const bool pitched = has_pitch(p->x_pitch) || has_pitch(p->y_pitch);
  const int Ho = (p->height + 2 * p->pad - p->kh) / p->stride + 1, Wo = (p->width + 2 * pad_w - p->kw) / p->stride + 1;
  if (Ho <= 0 || Wo <= 0) return MRB_ERR_BAD_ARG;
  // geometry of the pixel (reduction) axis; 1x1 stride-1 layers flatten the batch onto one axis
  int batch = p->batch, gh = Ho, gw = Wo, xh = p->stride == 2 ? Ho : p->height, xw = p->stride == 2 ? Wo : p->width;
  long long g_w = p->cout, g_h = (long long)Wo * p->cout, g_n = (long long)Ho * Wo * p->cout;
  long long x_w = (long long)p->cin * p->stride, x_h = (long long)p->width * p->cin * p->stride,
            x_n = (long long)p->height * p->width * p->cin;
  if (p->kh == 1 && p->kw == 1 && p->stride == 1 && p->pad == 0 && pad_w == 0 && !pitched) {
    gw = xw = p->batch * Ho * Wo; gh = xh = 1; batch = 1;
    g_h = g_n = (long long)gw * p->cout; x_h = x_n = (long long)xw * p->cin;
  }
  if (has_pitch(p->x_pitch)) { x_n = p->x_pitch[0]; x_h = p->x_pitch[1]; x_w = p->x_pitch[2]; }
  if (has_pitch(p->y_pitch)) { g_n = p->y_pitch[0]; g_h = p->y_pitch[1]; g_w = p->y_pitch[2]; }
  if ((x_w | x_h | x_n | g_w | g_h | g_n) & 7) return MRB_ERR_UNSUPPORTED;   // TMA strides: multiples of 16 B
  WgradArgs a;
  int best_th = 1, best_tw = 64;
  long long best = -1;
  const int th_order[7] = {8, 4, 16, 2, 32, 1, 64};
  for (int i = 0; i < 7; ++i) {
    const int th = th_order[i], tw = 64 / th;
    const long long cost = (long long)ceil_div(gh, th) * th * ceil_div(gw, tw) * tw;
    if (best < 0 || cost < best) { best = cost; best_th = th; best_tw = tw; }
  }
  a.th = best_th; a.tw = best_tw; a.tiles_h = ceil_div(gh, a.th); a.tiles_w = ceil_div(gw, a.tw); a.batch = batch;
  const long long kblocks = (long long)batch * a.tiles_h * a.tiles_w;
  if (kblocks >= (1ll << 31)) return MRB_ERR_UNSUPPORTED;
  a.kblocks_total = (int)kblocks;
  a.kh = p->kh; a.kw = p->kw; a.pad_h = p->pad; a.pad_w = pad_w; a.taps = taps;
  a.cin = p->cin; a.cout = p->cout;
  a.bn = ceil_div(p->cin, 64) * 64;
  if (a.bn > 256) a.bn = 256;
  a.grouped = grouped ? 1 : 0;
  a.co_tiles = ceil_div(p->cout, 128); a.ci_tiles = ceil_div(p->cin, a.bn);
  if (grouped) { a.bn = 128; a.ci_tiles = 1; }
  const int out_tiles = a.co_tiles * a.ci_tiles * taps;
  // one wave of equally sized work items (items <= #SMs): every extra split costs a 128 x BN fp32 red.add flush
  int splits = kNumSMs / out_tiles;
  const int max_splits = ceil_div(a.kblocks_total, 8);     // at least 8 k-blocks (512 pixels) per item
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.kblocks_per_split = ceil_div(a.kblocks_total, splits);
  a.splits = ceil_div(a.kblocks_total, a.kblocks_per_split);
  a.items_total = out_tiles * a.splits;
  a.dw = grad_weight;
  a.scale = scale;
  const uint32_t stage_bytes = (2 + a.bn / 64) * kWgGroupBytes;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > 8) stages = 8;
  a.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 8 * (2 * stages + 4) + 16 + 1024;
  CUtensorMap map_g, map_x;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p->cout, (cuuint64_t)gw, (cuuint64_t)gh, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)g_w * 2, (cuuint64_t)g_h * 2, (cuuint64_t)g_n * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)a.tw, (cuuint32_t)a.th, 1};
    rc = encode_bf16(&map_g, grad_output, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)p->cin, (cuuint64_t)xw, (cuuint64_t)xh, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)x_w * 2, (cuuint64_t)x_h * 2, (cuuint64_t)x_n * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)a.tw, (cuuint32_t)a.th, 1};
    rc = encode_bf16(&map_x, input, 4, dims, strides, box);
    if (rc) return rc;
  }
  MRB_CUDA_TRY(ensure_max_smem(conv_wgrad_tc_kernel, 1));
  const int grid = a.items_total < kNumSMs ? a.items_total : kNumSMs;
  MRB_CUDA_TRY(launch_pdl(conv_wgrad_tc_kernel, grid, kWgradThreads, smem, stream, 1, map_g, map_x, a));
  return MRB_OK;
}

MRB_API int mrb_conv2d_wgrad(const mrb_conv_params* p, const void* input, const void* grad_output, const float* scale,
                             float* grad_weight, mrb_stream_t stream) {
  return conv_wgrad_impl(p, input, grad_output, scale, grad_weight, false, stream);
}
MRB_API int mrb_conv2d_wgrad_accumulate(const mrb_conv_params* p, const void* input, const void* grad_output,
                                        const float* scale, float* grad_weight, mrb_stream_t stream) {
  return conv_wgrad_impl(p, input, grad_output, scale, grad_weight, true, stream);
}

// ------------------------------------------------------------------------------- bias gradient
// db[c] = sum over pixels of g[pixel, c] for an NHWC bf16 gradient: each thread owns 2 adjacent channels
// (one bf16x2 word), a CTA walks a slab of pixels with fully coalesced rows, one red.add per channel pair.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const __nv_bfloat16* __restrict__ g, float* __restrict__ db, long long pixels, int C, int rows_per_cta) {
  const int pairs = C >> 1;
  const long long p0 = (long long)blockIdx.x * rows_per_cta;
  const long long p1 = p0 + rows_per_cta < pixels ? p0 + rows_per_cta : pixels;
  for (int c2 = threadIdx.x; c2 < pairs; c2 += 256) {
    float sx = 0.f, sy = 0.f;
    const __nv_bfloat162* col = reinterpret_cast<const __nv_bfloat162*>(g) + c2;
#pragma unroll 4
    for (long long p = p0; p < p1; ++p) {
      const float2 t = __bfloat1622float2(col[p * pairs]);
      sx += t.x; sy += t.y;
    }
    atomicAdd(db + 2 * c2, sx);
    atomicAdd(db + 2 * c2 + 1, sy);
  }
}

// Vector version (C % 8 == 0, C <= 2048): a thread owns 8 adjacent channels (one 16-byte word) of every rpi-th pixel of
// the CTA's slab, C/8 threads cover a pixel row, so the whole block streams full lines; partial sums meet in shared
// memory and leave as one red.add per channel per CTA.
__global__ void __launch_bounds__(256)
bias_grad_vec_kernel(const uint4* __restrict__ g, float* __restrict__ db, long long pixels, int C8, int rows_per_cta) {
  __shared__ float part[256 * 8];
  const int rpi = 256 / C8;                              // pixel rows per block iteration
  const int c8 = threadIdx.x % C8, r = threadIdx.x / C8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long p0 = (long long)blockIdx.x * rows_per_cta;
  const long long p1 = p0 + rows_per_cta < pixels ? p0 + rows_per_cta : pixels;
  if (r < rpi) {
#pragma unroll 4
    for (long long p = p0 + r; p < p1; p += rpi) {
      const uint4 v = __ldcs(g + p * C8 + c8);
      const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(v2[j]);
        acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < C8 * 8; c += 256) {
    float s = 0.f;
    for (int rr = 0; rr < rpi; ++rr) s += part[(rr * C8 + (c >> 3)) * 8 + (c & 7)];
    atomicAdd(db + c, s);
  }
}

static int bias_grad_impl(const void* grad_bf16_nhwc, float* grad_bias, long long pixels, int channels, bool accumulate,
                          mrb_stream_t stream_) {
  if (pixels < 0 || channels <= 0 || (channels & 1)) return MRB_ERR_BAD_ARG;
  if (!grad_bias) return MRB_ERR_BAD_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!accumulate) MRB_CUDA_TRY(cudaMemsetAsync(grad_bias, 0, sizeof(float) * channels, stream));
  if (pixels == 0) return MRB_OK;
  if (!grad_bf16_nhwc || ((uintptr_t)grad_bf16_nhwc & 3)) return MRB_ERR_BAD_ARG;
  if ((channels & 7) == 0 && channels <= 2048 && ((uintptr_t)grad_bf16_nhwc & 15) == 0) {
    const int C8 = channels >> 3, rpi = 256 / C8;
    long long ctas = (pixels + (long long)rpi * 8 - 1) / ((long long)rpi * 8);
    const long long cap = (long long)kNumSMs * 4;
    if (ctas > cap) ctas = cap;
    const int rows = (int)((pixels + ctas - 1) / ctas);
    ctas = (pixels + rows - 1) / rows;
    bias_grad_vec_kernel<<<(unsigned)ctas, 256, 0, stream>>>((const uint4*)grad_bf16_nhwc, grad_bias, pixels, C8, rows);
    MRB_LAUNCH_CHECK();
    return MRB_OK;
  }
  long long ctas = (pixels + 63) / 64;
  const long long cap = (long long)kNumSMs * 8;
  if (ctas > cap) ctas = cap;
  const int rows = (int)((pixels + ctas - 1) / ctas);
  ctas = (pixels + rows - 1) / rows;
  bias_grad_kernel<<<(unsigned)ctas, 256, 0, stream>>>((const __nv_bfloat16*)grad_bf16_nhwc, grad_bias, pixels, channels, rows);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_bias_grad(const void* grad_bf16_nhwc, float* grad_bias, long long pixels, int channels, mrb_stream_t stream) {
  return bias_grad_impl(grad_bf16_nhwc, grad_bias, pixels, channels, false, stream);
}
MRB_API int mrb_bias_grad_accumulate(const void* grad_bf16_nhwc, float* grad_bias, long long pixels, int channels,
                                     mrb_stream_t stream) {
  return bias_grad_impl(grad_bf16_nhwc, grad_bias, pixels, channels, true, stream);
}
