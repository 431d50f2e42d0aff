// nms.cu -- greedy IoU non-maximum suppression, fully on device, for sm_100a.
//
// Replaces nms_cpu_kernel (reference csrc/cpu/nms_cpu.cpp:5-65 -- the parity target) and
// nms_kernel/nms_cuda (csrc/cuda/nms.cu:23-131).  Differences from the reference CUDA path, all
// deliberate:
//   * suppression test is `IoU >= thr` and the IoU is computed with round-to-nearest
//     sub/add/mul/div and no FMA, in the CPU kernel's operation order (nms_cpu.cpp:49-60), so
//     the kept set is bit-exact with the CPU reference (the reference CUDA kernel uses `>` and
//     lets nvcc contract, nms.cu:13-21,60);
//   * no blocking D2H copy of the N x N/64 mask and no host scan (nms.cu:100-123): one CTA per
//     problem resolves the greedy walk as a fixed point over kept / removed bit sets in smem (a box
//     is removed once a suppressor is known kept, kept once all its suppressors are known removed);
//     same unique result as the serial walk in a handful of rounds instead of n/64 dependent steps;
//   * only the upper triangle of mask tiles is computed (the reference computes and ships all);
//   * the descending-score order is produced on device by a rank (counting) sort on a total-order
//     integer key with index tie-break == a stable descending sort;
//   * many (image, level) problems are processed by one launch sequence (mrb_nms_batched).
// Kept indices are emitted ascending by original index (at::nonzero, nms_cpu.cpp:64).
#include "common.cuh"

namespace mrb {

constexpr int kMaxProblems = 64;
constexpr int kRankThreads = 128;
constexpr int kScanThreads = 1024;

struct NmsBatch {
  int num;
  int off[kMaxProblems + 1];        // row offsets of each problem in boxes/scores/keep
  unsigned long long ws[kMaxProblems];  // byte offset of each problem's workspace slice
};

struct NmsWs {
  float4* boxes;          // [n] sorted xyxy
  float* areas;           // [n]
  int* order;             // [n] sorted position -> original index
  unsigned char* flags;   // [n] keep flag per ORIGINAL index
  unsigned long long* mask;  // [n][col_blocks]
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t nms_ws_bytes(int n) {
  const size_t cb = (size_t)(n + 63) / 64;
  return align16((size_t)n * 16) + align16((size_t)n * 4) + align16((size_t)n * 4) + align16((size_t)n) +
         align16((size_t)n * cb * 8);
}

__host__ __device__ inline NmsWs nms_ws_carve(void* base, int n) {
  NmsWs w;
  unsigned char* p = (unsigned char*)base;
  w.boxes = (float4*)p; p += align16((size_t)n * 16);
  w.areas = (float*)p; p += align16((size_t)n * 4);
  w.order = (int*)p; p += align16((size_t)n * 4);
  w.flags = p; p += align16((size_t)n);
  w.mask = (unsigned long long*)p;
  return w;
}

// total-order key: larger float -> larger unsigned; -0.0 == +0.0; NaN sorts first (largest),
// matching torch.sort(descending=True)
__device__ __forceinline__ unsigned score_key(float s) {
  s = s + 0.0f;  // -0.0 -> +0.0
  unsigned u = __float_as_uint(s);
  if (s != s) return 0xffffffffu;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// --- 1. rank sort + gather ----------------------------------------------------------------
__global__ void __launch_bounds__(kRankThreads)
nms_rank_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, NmsBatch nb,
                unsigned char* __restrict__ ws_base) {
  const int p = blockIdx.y;
  const int n = nb.off[p + 1] - nb.off[p];
  if ((int)(blockIdx.x * kRankThreads) >= n) return;
  const float* __restrict__ sc = scores + nb.off[p];
  const float4* __restrict__ bx = reinterpret_cast<const float4*>(boxes) + nb.off[p];
  const NmsWs w = nms_ws_carve(ws_base + nb.ws[p], n);
  __shared__ unsigned tile[kRankThreads];
  const int i = blockIdx.x * kRankThreads + threadIdx.x;
  const unsigned ki = (i < n) ? score_key(sc[i]) : 0u;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += kRankThreads) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    tile[threadIdx.x] = (j < n) ? score_key(sc[j]) : 0u;
    __syncthreads();
    const int lim = min(kRankThreads, n - j0);
#pragma unroll 8
    for (int t = 0; t < lim; ++t) {
      const unsigned kj = tile[t];
      rank += (kj > ki) || (kj == ki && (j0 + t) < i);
    }
  }
  if (i < n) {
    const float4 b = bx[i];
    w.boxes[rank] = b;
    // nms_cpu.cpp:22: (x2 - x1 + 1) * (y2 - y1 + 1)
    w.areas[rank] = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
    w.order[rank] = i;
    w.flags[i] = 0;
  }
}

// rows already in descending-score order (the RPN's sorted top-k): rank == index, the O(n^2) rank sort is a plain gather
__global__ void __launch_bounds__(kRankThreads)
nms_presorted_kernel(const float* __restrict__ boxes, NmsBatch nb, unsigned char* __restrict__ ws_base) {
  const int p = blockIdx.y;
  const int n = nb.off[p + 1] - nb.off[p];
  const int i = blockIdx.x * kRankThreads + threadIdx.x;
  if (i >= n) return;
  const NmsWs w = nms_ws_carve(ws_base + nb.ws[p], n);
  const float4 b = (reinterpret_cast<const float4*>(boxes) + nb.off[p])[i];
  w.boxes[i] = b;
  w.areas[i] = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
  w.order[i] = i;
  w.flags[i] = 0;
}

// --- 2. upper-triangular suppression mask ----------------------------------------------------
__device__ __forceinline__ bool suppresses(const float4 a, float a_area, const float4 b, float b_area, float thr) {
  // nms_cpu.cpp:49-60
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
  const float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
  // disjoint boxes: inter == +0, so the quotient is +-0 or NaN and `>= thr` is false for every thr > 0 -- skip the division
  // (most pairs of an RPN problem; the kernel is bound by the divisions otherwise)
  if (thr > 0.f && (w == 0.f || h == 0.f)) return false;
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a_area, b_area), inter));
  return ovr >= thr;
}

__global__ void __launch_bounds__(64)
nms_mask_kernel(NmsBatch nb, unsigned char* __restrict__ ws_base, float thr) {
  // Suppressor matrix: word (c, row_b) of `mask` holds, for the box of rank c, the boxes r of rank block row_b with r < c
  // that suppress it (IoU >= thr).  suppresses() is symmetric bit for bit (max/min and a commutative sum), so the tile is
  // evaluated from the column box's side; only tiles on or above the diagonal exist.
  const int p = blockIdx.z;
  const int n = nb.off[p + 1] - nb.off[p];
  const int col_blocks = (n + 63) >> 6;
  const int row_b = blockIdx.y, col_b = blockIdx.x;
  if (row_b >= col_blocks || col_b >= col_blocks || col_b < row_b) return;
  const NmsWs w = nms_ws_carve(ws_base + nb.ws[p], n);
  __shared__ float4 rb[64];
  __shared__ float ra[64];
  const int col_size = min(n - col_b * 64, 64);
  const int row_size = min(n - row_b * 64, 64);
  if ((int)threadIdx.x < row_size) {
    rb[threadIdx.x] = w.boxes[row_b * 64 + threadIdx.x];
    ra[threadIdx.x] = w.areas[row_b * 64 + threadIdx.x];
  }
  __syncthreads();
  if ((int)threadIdx.x < col_size) {
    const int c = col_b * 64 + threadIdx.x;
    const float4 a = w.boxes[c];
    const float aa = w.areas[c];
    unsigned long long t = 0;
    const int end = (row_b == col_b) ? (int)threadIdx.x : row_size;   // rows of lower rank only
    for (int i = 0; i < end; ++i)
      if (suppresses(rb[i], ra[i], a, aa, thr)) t |= 1ull << i;
    w.mask[(size_t)c * col_blocks + row_b] = t;
  }
}

// --- 3. greedy resolution + ascending compaction, one CTA per problem --------------------------
// The greedy scan "walk the boxes by descending score, keep a box unless a kept one suppresses it" has a unique
// solution, reached here as a fixed point instead of a serial walk: a box is REMOVED as soon as one of its suppressors is
// known kept, KEPT as soon as all of its suppressors are known removed (a box without suppressors at once).  Every round
// decides at least the undecided box of lowest rank, typical inputs need a handful of rounds (the length of the longest
// keep/remove dependency chain), and a round costs one L2 latency instead of the serial walk's two per 64 boxes.
__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(NmsBatch nb, unsigned char* __restrict__ ws_base, long long* __restrict__ keep,
                int* __restrict__ num_keep) {
  extern __shared__ unsigned long long sm_scan[];  // kept[col_blocks], removed[col_blocks], nz[n] (col_blocks <= 64)
  __shared__ int s_warp_tot[kScanThreads / 32];
  __shared__ int s_base;
  const int p = blockIdx.x;
  const int n = nb.off[p + 1] - nb.off[p];
  const int col_blocks = (n + 63) >> 6;
  const NmsWs w = nms_ws_carve(ws_base + nb.ws[p], n);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long* kept = sm_scan;
  unsigned long long* removed = sm_scan + col_blocks;
  unsigned long long* nz = sm_scan + 2 * col_blocks;
  const bool use_nz = col_blocks <= 64;
  // the bit sets only grow (atomicOr) while a round reads them: any mix of old and new words yields decisions that are valid
  volatile unsigned long long* vkept = kept;
  volatile unsigned long long* vremoved = removed;
  for (int t = tid; t < 2 * col_blocks; t += kScanThreads) sm_scan[t] = 0;
  __syncthreads();
  // round 0: which suppressor words of every box are non-zero; boxes without suppressors are kept
  for (int i = tid; i < n; i += kScanThreads) {
    const int wd_end = (i >> 6) + 1;
    const unsigned long long* row = w.mask + (size_t)i * col_blocks;
    unsigned long long m = 0;
    bool any = false;
#pragma unroll 4
    for (int wd = 0; wd < wd_end; ++wd) {
      const unsigned long long t = row[wd];
      if (t) {
        any = true;
        if (use_nz) m |= 1ull << wd;
      }
    }
    if (use_nz) nz[i] = m;
    if (!any) atomicOr(&kept[i >> 6], 1ull << (i & 63));
  }
  for (;;) {
    __syncthreads();
    int undecided = 0;
    for (int i = tid; i < n; i += kScanThreads) {
      const int word = i >> 6;
      const unsigned long long bit = 1ull << (i & 63);
      if ((vkept[word] | vremoved[word]) & bit) continue;
      const unsigned long long* row = w.mask + (size_t)i * col_blocks;
      bool any_kept = false, any_open = false;
      if (use_nz) {
        unsigned long long m = nz[i];
        while (m) {
          const int wd = __ffsll((long long)m) - 1;
          m &= m - 1;
          const unsigned long long t = row[wd];
          const unsigned long long k = vkept[wd], r = vremoved[wd];
          any_kept |= (t & k) != 0;
          any_open |= (t & ~k & ~r) != 0;
        }
      } else {
        for (int wd = 0; wd <= word; ++wd) {
          const unsigned long long t = row[wd];
          const unsigned long long k = vkept[wd], r = vremoved[wd];
          any_kept |= (t & k) != 0;
          any_open |= (t & ~k & ~r) != 0;
        }
      }
      if (any_kept) atomicOr(&removed[word], bit);
      else if (!any_open) atomicOr(&kept[word], bit);
      else ++undecided;
    }
    if (__syncthreads_count(undecided) == 0) break;
  }
  for (int i = tid; i < n; i += kScanThreads)
    if ((kept[i >> 6] >> (i & 63)) & 1ull) w.flags[w.order[i]] = 1;
  __syncthreads();

  // ascending-index compaction of flags[0..n)
  if (tid == 0) s_base = 0;
  __syncthreads();
  long long* __restrict__ out = keep + nb.off[p];
  for (int base = 0; base < n; base += kScanThreads) {
    const int i = base + tid;
    const int f = (i < n) ? (int)w.flags[i] : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp_tot[warp] = __popc(bal);
    __syncthreads();
    int warp_off = 0, tot = 0;
    for (int k = 0; k < kScanThreads / 32; ++k) {
      const int v = s_warp_tot[k];
      if (k < warp) warp_off += v;
      tot += v;
    }
    const int b0 = s_base;
    if (f) out[b0 + warp_off + in_warp] = i;
    __syncthreads();
    if (tid == 0) s_base = b0 + tot;
    __syncthreads();
  }
  if (tid == 0) num_keep[p] = s_base;
}

static int nms_run(const float* boxes, const float* scores, const NmsBatch& nb, float thr, long long* keep,
                   int* num_keep, void* ws, cudaStream_t stream, bool presorted = false) {
  int max_n = 0;
  for (int p = 0; p < nb.num; ++p) max_n = max(max_n, nb.off[p + 1] - nb.off[p]);
  if (max_n == 0) {
    return (int)cudaMemsetAsync(num_keep, 0, sizeof(int) * nb.num, stream);
  }
  const int max_cb = (max_n + 63) / 64;
  if ((size_t)max_cb * 16 > 160 * 1024) return MRB_ERR_UNSUPPORTED;
  {
    dim3 grid(ceil_div(max_n, kRankThreads), nb.num);
    if (presorted)
      nms_presorted_kernel<<<grid, kRankThreads, 0, stream>>>(boxes, nb, (unsigned char*)ws);
    else
      nms_rank_kernel<<<grid, kRankThreads, 0, stream>>>(boxes, scores, nb, (unsigned char*)ws);
    MRB_LAUNCH_CHECK();
  }
  {
    if (max_cb > 65535) return MRB_ERR_UNSUPPORTED;
    dim3 grid(max_cb, max_cb, nb.num);
    nms_mask_kernel<<<grid, 64, 0, stream>>>(nb, (unsigned char*)ws, thr);
    MRB_LAUNCH_CHECK();
  }
  {
    const size_t smem = (size_t)max_cb * 16 + (max_cb <= 64 ? (size_t)max_n * 8 : 0);
    if (smem > 48 * 1024)
      MRB_CUDA_TRY(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_scan_kernel<<<nb.num, kScanThreads, smem, stream>>>(nb, (unsigned char*)ws, keep, num_keep);
    MRB_LAUNCH_CHECK();
  }
  return MRB_OK;
}

}  // namespace mrb
using namespace mrb;

MRB_API size_t mrb_nms_workspace_bytes(int n) { return n > 0 ? nms_ws_bytes(n) : 16; }

MRB_API size_t mrb_nms_batched_workspace_bytes(const int* offsets_host, int num_problems) {
  size_t tot = 16;
  for (int p = 0; p < num_problems; ++p) tot += nms_ws_bytes(offsets_host[p + 1] - offsets_host[p]);
  return tot;
}

static int nms_batched_impl(const float* boxes, const float* scores, const int* offsets_host, int num_problems,
                            float threshold, int64_t* keep, int32_t* num_keep, void* workspace,
                            size_t workspace_bytes, mrb_stream_t stream, bool presorted) {
  if (num_problems < 0 || !offsets_host) return MRB_ERR_BAD_ARG;
  if (num_problems == 0) return MRB_OK;
  if (!num_keep) return MRB_ERR_BAD_ARG;
  if (((uintptr_t)workspace & 15) || ((uintptr_t)boxes & 15)) return MRB_ERR_BAD_ARG;
  for (int p0 = 0; p0 < num_problems; p0 += kMaxProblems) {
    NmsBatch nb;
    nb.num = min(kMaxProblems, num_problems - p0);
    size_t off = 0;
    for (int p = 0; p < nb.num; ++p) {
      nb.off[p] = offsets_host[p0 + p];
      const int n = offsets_host[p0 + p + 1] - offsets_host[p0 + p];
      if (n < 0) return MRB_ERR_BAD_ARG;
      nb.ws[p] = off;
      off += nms_ws_bytes(n);
    }
    nb.off[nb.num] = offsets_host[p0 + nb.num];
    if (off > workspace_bytes) return MRB_ERR_WORKSPACE;
    if (off > 0 && (!boxes || !scores || !keep || !workspace)) return MRB_ERR_BAD_ARG;
    int rc = nms_run(boxes, scores, nb, threshold, (long long*)keep, num_keep + p0, workspace, (cudaStream_t)stream, presorted);
    if (rc) return rc;
  }
  return MRB_OK;
}

MRB_API int mrb_nms_batched(const float* boxes, const float* scores, const int* offsets_host, int num_problems,
                            float threshold, int64_t* keep, int32_t* num_keep, void* workspace,
                            size_t workspace_bytes, mrb_stream_t stream) {
  return nms_batched_impl(boxes, scores, offsets_host, num_problems, threshold, keep, num_keep, workspace, workspace_bytes, stream, false);
}

MRB_API int mrb_nms_batched_presorted(const float* boxes, const int* offsets_host, int num_problems, float threshold, int64_t* keep,
                                      int32_t* num_keep, void* workspace, size_t workspace_bytes, mrb_stream_t stream) {
  return nms_batched_impl(boxes, boxes /* unused */, offsets_host, num_problems, threshold, keep, num_keep, workspace,
                          workspace_bytes, stream, true);
}

MRB_API int mrb_nms(const float* boxes, const float* scores, int n, float threshold, int64_t* keep,
                    int32_t* num_keep, void* workspace, size_t workspace_bytes, mrb_stream_t stream) {
  if (n < 0) return MRB_ERR_BAD_ARG;
  const int offs[2] = {0, n};
  return mrb_nms_batched(boxes, scores, offs, 1, threshold, keep, num_keep, workspace, workspace_bytes, stream);
}
