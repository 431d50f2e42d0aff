// detect_glue.cu -- the detection "glue" between the tensor-core phases of a train step, as a handful of launches.
//
// In the reference these stages are Python over BoxList objects, a few hundred tiny elementwise launches and several
// host synchronisations per step:
//   1. mrb_rpn_decode      RPNPostProcessor.forward_for_single_feature_map after the top-k (modeling/rpn/inference.py:
//                          91-111): gather the k best anchors' deltas, BoxCoder.decode (modeling/box_coder.py:52-95),
//                          clip_to_image (structures/bounding_box.py:198-212), sigmoid of the k logits.
//   2. mrb_rpn_collect     the part after NMS (inference.py:116-123): first post_nms_top_n survivors of every
//                          (image, level) problem, select_over_all_levels (inference.py:154-181, per batch in training /
//                          per image otherwise) and add_gt_proposals (inference.py:53-74), as fixed-shape
//                          [N, W + Gmax] rows with a validity flag.
//   3. mrb_roi_assign_sample  FastRCNNLossComputation.match_targets_to_proposals / prepare_targets / subsample
//                          (modeling/roi_heads/box_head/loss.py:41-118): IoU against the ground truth
//                          (structures/boxlist_ops.py:53-89), Matcher (modeling/matcher.py:42-81, no low-quality pass),
//                          BalancedPositiveNegativeSampler (balanced_positive_negative_sampler.py:19-68) driven by
//                          caller-supplied iid keys (the `limit` smallest keys == randperm[:limit]), BoxCoder.encode
//                          (box_coder.py:22-50) of the sampled rows, and keep_only_positive_boxes
//                          (mask_head/mask_head.py:11-32) as a positives-first list for the mask branch.
//   4. mrb_rpn_anchor_match   RPNLossComputation.match_targets_to_anchors / prepare_targets (modeling/rpn/loss.py:40-90):
//                          IoU of every anchor with the ground truth, Matcher with allow_low_quality_matches
//                          (matcher.py:83-112), visibility / between-threshold discards -> label per anchor.
// Arithmetic follows the reference's operation order with round-to-nearest mul/add/div and no FMA contraction, so the
// results equal the PyTorch formulation run on the same GPU bit for bit (tests/test_glue_gpu.py).
// Selection problems (top-n over scores, n smallest keys) are solved by an 8-bit radix select on a total-order integer
// key with ties broken by ascending index; the reference's torch.topk / randperm leave tie order unspecified.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mrb {

constexpr int kGlueThreads = 1024;
constexpr int kMaxLevels = 8;

// larger float -> larger unsigned (total order; -0.0 == +0.0)
__device__ __forceinline__ unsigned glue_key(float s) {
  s = s + 0.0f;
  const unsigned u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// exclusive prefix sum of one int per thread over the CTA (kGlueThreads threads); *total = sum.  `ws`: 33 ints of smem.
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* ws) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  __syncthreads();  // ws may still be read from a previous call
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = ws[lane];
    int si = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, si, d);
      if (lane >= d) si += t;
    }
    ws[lane] = si - s;  // exclusive offset of every warp
    if (lane == 31) ws[32] = si;
  }
  __syncthreads();
  *total = ws[32];
  return ws[warp] + inc - v;
}

// Radix select over the keys {key(i) : present(i), i in [0, n)} of ALL CTAs of the cluster (one CTA = one slice):
// finds the threshold T such that taking every key > T and the first `quota` keys == T (in slice-major, index order)
// yields the `want` largest.  If fewer than `want` keys are present, T = 0 and quota = 0 with everything present > T
// (keys of present elements must be > 0).  `hist`, `ghist`: 256 ints of smem each; `sh`: 4 ints.
template <bool kCluster = true, class F>
__device__ void radix_select(F key_of, int n, int want, unsigned* T_out, int* quota_out, int* hist, int* ghist, int* sh) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned nranks = kCluster ? cluster.num_blocks() : 1u;
  unsigned prefix = 0, mask = 0;
  int remaining = want;
  bool short_of = false;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int t = threadIdx.x; t < 256; t += blockDim.x) hist[t] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      unsigned k;
      if (key_of(i, &k) && (k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
    }
    if (kCluster) {
      cluster.sync();
      for (int t = threadIdx.x; t < 256; t += blockDim.x) {
        int s = 0;
        for (unsigned r = 0; r < nranks; ++r) s += cluster.map_shared_rank(hist, r)[t];
        ghist[t] = s;
      }
      cluster.sync();  // everybody has read every hist before the next pass clears it; also orders ghist for thread 0
    } else {           // single CTA: the local histogram is the global one
      __syncthreads();
      for (int t = threadIdx.x; t < 256; t += blockDim.x) ghist[t] = hist[t];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      int acc = 0, b = 255;
      for (; b >= 0; --b) {
        if (acc + ghist[b] >= remaining) break;
        acc += ghist[b];
      }
      if (b < 0) {  // fewer than `want` candidates in total
        sh[0] = -1;
      } else {
        sh[0] = b;
        sh[1] = remaining - acc;
      }
    }
    __syncthreads();
    if (sh[0] < 0) {
      short_of = true;
      __syncthreads();
      break;
    }
    prefix |= (unsigned)sh[0] << shift;
    mask |= 255u << shift;
    remaining = sh[1];
    __syncthreads();
  }
  if (short_of) {
    *T_out = 0u;
    *quota_out = 0;
  } else {
    *T_out = prefix;
    *quota_out = remaining;
  }
}

// ------------------------------------------------------------------------------------------ 1. decode + clip + sigmoid
__global__ void __launch_bounds__(256)
rpn_decode_kernel(const float* __restrict__ logits, const float4* __restrict__ deltas, const float4* __restrict__ anchors,
                  const int64_t* __restrict__ idx, const float* __restrict__ im_w, const float* __restrict__ im_h,
                  float4* __restrict__ boxes, float* __restrict__ scores, int N, int A, int k, float wx, float wy, float ww,
                  float wh, float clip, int packed_apl, int packed_ld) {   // wx..wh: RECIPROCALS of the box-coder weights
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * k) return;
  const int i = t / k;
  int64_t a = idx[t];
  a = a < 0 ? 0 : (a >= A ? A - 1 : a);
  float4 d;
  float z;
  if (packed_apl > 0) {
    // head output [N, A / apl locations, apl logits + 4 apl deltas] read in place (`logits` is its base)
    const int pix = (int)(a / packed_apl), q = (int)(a - (int64_t)pix * packed_apl);
    const float* row = logits + ((size_t)i * (A / packed_apl) + pix) * (size_t)packed_ld;
    z = row[q];
    const float* dp = row + packed_apl + 4 * q;
    d = make_float4(dp[0], dp[1], dp[2], dp[3]);
  } else {
    d = deltas[(size_t)i * A + a];
    z = logits[(size_t)i * A + a];
  }
  const float4 an = anchors[a];
  // box_coder.py:62-93 (TO_REMOVE = 1)
  const float w = __fadd_rn(__fsub_rn(an.z, an.x), 1.f), h = __fadd_rn(__fsub_rn(an.w, an.y), 1.f);
  const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
  // rel_codes / w with a Python scalar w: ATen multiplies by the reciprocal (computed in fp32 on the host)
  const float dx = __fmul_rn(d.x, wx), dy = __fmul_rn(d.y, wy);
  const float dw = fminf(__fmul_rn(d.z, ww), clip), dh = fminf(__fmul_rn(d.w, wh), clip);
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  float x2 = __fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.f), y2 = __fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 1.f);
  // bounding_box.py:198-206: clamp_(min=0, max=size-1)
  const float lx = __fsub_rn(im_w[i], 1.f), ly = __fsub_rn(im_h[i], 1.f);
  x1 = fminf(fmaxf(x1, 0.f), lx);
  y1 = fminf(fmaxf(y1, 0.f), ly);
  x2 = fminf(fmaxf(x2, 0.f), lx);
  y2 = fminf(fmaxf(y2, 0.f), ly);
  boxes[t] = make_float4(x1, y1, x2, y2);
  scores[t] = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-z)));   // inference.py:88 (.sigmoid())
}

// ------------------------------------------------------------------------------------------ 1b. top-k + decode in one launch
// objectness.topk(pre_nms_top_n, sorted=True) (inference.py:91-95) fused with the decode: one thread-block cluster per image
// (its CTAs split the level's anchors), keys cached in shared memory, radix select of the k-th logit over the cluster
// (histograms summed through distributed shared memory), the k selected (key, anchor) pairs gathered into the leader CTA's
// shared memory in slice order, sorted there (bitonic, descending score, ascending anchor among equal logits) and decoded.
struct TopkArgs {
  int N, A, k, apl, ld, slice;      // images, anchors of the level, k, anchors per location, floats per location, anchors per CTA
  float wx, wy, ww, wh, clip;       // reciprocals of the box-coder weights
};

__device__ __forceinline__ void decode_one(const float4 d, const float4 an, float lx, float ly, const TopkArgs& a, float4* out) {
  const float w = __fadd_rn(__fsub_rn(an.z, an.x), 1.f), h = __fadd_rn(__fsub_rn(an.w, an.y), 1.f);
  const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
  const float dx = __fmul_rn(d.x, a.wx), dy = __fmul_rn(d.y, a.wy);
  const float dw = fminf(__fmul_rn(d.z, a.ww), a.clip), dh = fminf(__fmul_rn(d.w, a.wh), a.clip);
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  float x2 = __fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.f), y2 = __fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 1.f);
  x1 = fminf(fmaxf(x1, 0.f), lx);
  y1 = fminf(fmaxf(y1, 0.f), ly);
  x2 = fminf(fmaxf(x2, 0.f), lx);
  y2 = fminf(fmaxf(y2, 0.f), ly);
  *out = make_float4(x1, y1, x2, y2);
}

__global__ void __launch_bounds__(kGlueThreads)
rpn_topk_decode_kernel(const float* __restrict__ head_out, const float4* __restrict__ anchors, const float* __restrict__ im_w,
                       const float* __restrict__ im_h, float4* __restrict__ boxes, float* __restrict__ scores, TopkArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int hist[256], ghist[256], sh[4], scan_ws[33], cnt_sh[2];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned cs = cluster.num_blocks(), rank = cluster.block_rank();
  const int i = blockIdx.x / cs;                                   // image
  unsigned* keys = (unsigned*)smem_raw;                            // [slice]
  int p2 = 1;
  while (p2 < a.k) p2 <<= 1;
  unsigned long long* sortbuf = (unsigned long long*)(smem_raw + (((size_t)a.slice * 4 + 15) & ~(size_t)15));   // [p2], used in the leader
  const int a_lo = min((int)rank * a.slice, a.A), a_hi = min(a_lo + a.slice, a.A), n = a_hi - a_lo;
  const float* __restrict__ img = head_out + (size_t)i * (a.A / a.apl) * a.ld;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int g = a_lo + t, pix = g / a.apl, q = g - pix * a.apl;
    unsigned key = glue_key(img[(size_t)pix * a.ld + q]);
    keys[t] = key == 0 ? 1u : key;
  }
  __syncthreads();
  unsigned T;
  int quota;
  auto key_of = [&](int t, unsigned* k) { *k = keys[t]; return true; };
  radix_select(key_of, n, a.k, &T, &quota, hist, ghist, sh);
  // selected elements of this slice, in anchor order; ties at the cut are taken in (slice, anchor) order
  const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
  const int t_lo = min((int)threadIdx.x * per, n), t_hi = min(t_lo + per, n);
  int my_ties = 0;
  for (int t = t_lo; t < t_hi; ++t) my_ties += (keys[t] == T);
  int tie_tot;
  int tr = block_excl_scan(my_ties, &tie_tot, scan_ws);
  if (threadIdx.x == 0) cnt_sh[0] = tie_tot;
  cluster.sync();
  int ties_before = 0;
  for (unsigned r = 0; r < rank; ++r) ties_before += cluster.map_shared_rank(cnt_sh, r)[0];
  tr += ties_before;
  int my_sel = 0;
  for (int t = t_lo; t < t_hi; ++t) {
    const bool is_tie = keys[t] == T;
    const bool sel = keys[t] > T || (is_tie && tr < quota);
    tr += is_tie;
    my_sel += sel;
    if (!sel) keys[t] = 0;          // 0 = not selected from here on
  }
  int sel_tot;
  int pos = block_excl_scan(my_sel, &sel_tot, scan_ws);
  if (threadIdx.x == 0) cnt_sh[1] = sel_tot;
  cluster.sync();
  int base = 0;
  for (unsigned r = 0; r < rank; ++r) base += cluster.map_shared_rank(cnt_sh, r)[1];
  unsigned long long* lead = cluster.map_shared_rank(sortbuf, 0);
  pos += base;
  for (int t = t_lo; t < t_hi; ++t)
    if (keys[t]) {
      if (pos < p2) lead[pos] = ((unsigned long long)keys[t] << 32) | (unsigned)(0x7fffffff - (a_lo + t));
      ++pos;
    }
  if (rank == 0) {
    int total = 0;      // == min(k, A)
    for (unsigned r = 0; r < cs; ++r) total += cluster.map_shared_rank(cnt_sh, r)[1];
    if (threadIdx.x == 0) sh[3] = total;
  }
  cluster.sync();       // all pairs have landed in the leader's shared memory
  if (rank != 0) return;
  const int total = sh[3];
  for (int t = total + threadIdx.x; t < p2; t += blockDim.x) sortbuf[t] = 0ull;
  __syncthreads();
  for (int size = 2; size <= p2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (p2 >> 1); t += blockDim.x) {
        const int lo = (t / stride) * (stride << 1) + (t % stride), hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long x = sortbuf[lo], y = sortbuf[hi];
        if ((x < y) == desc) {
          sortbuf[lo] = y;
          sortbuf[hi] = x;
        }
      }
      __syncthreads();
    }
  }
  const float lx = __fsub_rn(im_w[i], 1.f), ly = __fsub_rn(im_h[i], 1.f);
  for (int t = threadIdx.x; t < a.k; t += blockDim.x) {
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = -1.f;
    if (t < total) {
      const int g = 0x7fffffff - (int)(unsigned)(sortbuf[t] & 0xffffffffull);
      const int pix = g / a.apl, q = g - pix * a.apl;
      const float* row = img + (size_t)pix * a.ld;
      const float* dp = row + a.apl + 4 * q;
      decode_one(make_float4(dp[0], dp[1], dp[2], dp[3]), anchors[g], lx, ly, a, &bx);
      sc = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-row[q])));
    }
    boxes[(size_t)i * a.k + t] = bx;
    scores[(size_t)i * a.k + t] = sc;
  }
}

// ------------------------------------------------------------------------------------------ 2. post-NMS selection
struct CollectArgs {
  int L, N, C, post_n, topn, W, gmax, per_batch, sorted;
  int k[kMaxLevels];       // rows per (image, level) problem
  int m[kMaxLevels];       // slots per level = min(k, post_n)
  int slot0[kMaxLevels];   // first slot of the level in an image's candidate list
  int row0[kMaxLevels];    // first row of the level's [N, k] block in boxes / scores / keep
};

__global__ void __launch_bounds__(kGlueThreads)
rpn_collect_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ keep,
                   const int32_t* __restrict__ counts, const float4* __restrict__ gt, const int32_t* __restrict__ gt_count,
                   float4* __restrict__ out_b, float* __restrict__ out_s, unsigned char* __restrict__ out_v, CollectArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int hist[256], ghist[256], sh[4], scan_ws[33], tie_cnt_sh;
  unsigned* keys = (unsigned*)smem_raw;                 // [C]  0 = invalid slot
  int* rows = (int*)(smem_raw + (size_t)a.C * 4);       // [C]  row in boxes / scores
  unsigned long long* sortbuf = (unsigned long long*)(smem_raw + (size_t)a.C * 8);   // [pow2 >= W] (sorted mode)
  cg::cluster_group cluster = cg::this_cluster();
  const int i = blockIdx.x;
  // candidate slots of this image
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    int l = 0;
    while (l + 1 < a.L && c >= a.slot0[l + 1]) ++l;
    const int j = c - a.slot0[l];
    const int cnt = min(counts[l * a.N + i], a.post_n);
    unsigned key = 0;
    int row = 0;
    if (j < cnt) {
      const int base = a.row0[l] + i * a.k[l];
      int64_t r = keep[base + j];
      r = r < 0 ? 0 : (r >= a.k[l] ? a.k[l] - 1 : r);
      row = base + (int)r;
      key = glue_key(scores[row]);
      if (key == 0) key = 1;
    }
    keys[c] = key;
    rows[c] = row;
  }
  __syncthreads();
  unsigned T;
  int quota;
  auto key_of = [&](int c, unsigned* k) { *k = keys[c]; return *k != 0; };
  radix_select(key_of, a.C, a.topn, &T, &quota, hist, ghist, sh);
  // ties at the threshold are taken in (image, slot) order: count those of the images before this one
  int ties_before = 0;
  if (a.per_batch && cluster.num_blocks() > 1) {
    int mine = 0;
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) mine += (keys[c] == T && T != 0);
    int tot;
    block_excl_scan(mine, &tot, scan_ws);
    if (threadIdx.x == 0) tie_cnt_sh = tot;
    cluster.sync();
    for (unsigned r = 0; r < cluster.block_rank(); ++r) ties_before += *cluster.map_shared_rank(&tie_cnt_sh, r);
    cluster.sync();
  }
  const size_t out_w = (size_t)a.W + a.gmax;
  float4* ob = out_b + (size_t)i * out_w;
  float* os = out_s + (size_t)i * out_w;
  unsigned char* ov = out_v + (size_t)i * out_w;
  // every thread owns a contiguous run of slots (slot order == output order): two CTA-wide scans in total
  const int per = (a.C + (int)blockDim.x - 1) / (int)blockDim.x;
  const int c_lo = min((int)threadIdx.x * per, a.C), c_hi = min(c_lo + per, a.C);
  int my_ties = 0;
  for (int c = c_lo; c < c_hi; ++c) my_ties += (keys[c] != 0 && keys[c] == T);
  int tie_tot;
  int tie_rank = ties_before + block_excl_scan(my_ties, &tie_tot, scan_ws);
  int my_sel = 0;
  for (int c = c_lo; c < c_hi; ++c) {
    const unsigned key = keys[c];
    const bool is_tie = key != 0 && key == T;
    const bool sel = key != 0 && (key > T || (is_tie && tie_rank < quota));
    tie_rank += is_tie;
    my_sel += sel;
    rows[c] = sel ? rows[c] : -1;       // rows doubles as the selection flag from here on
  }
  int n_sel;
  int pos = block_excl_scan(my_sel, &n_sel, scan_ws);
  for (int c = c_lo; c < c_hi; ++c) {
    const int row = rows[c];
    if (row < 0) continue;
    if (a.sorted) {
      sortbuf[pos] = ((unsigned long long)keys[c] << 32) | (unsigned)(0x7fffffff - c);   // descending: key, then ascending slot
    } else {
      ob[pos] = boxes[row];
      os[pos] = scores[row];
      ov[pos] = 1;
    }
    ++pos;
  }
  if (a.sorted) {
    int p2 = 1;
    while (p2 < a.W) p2 <<= 1;
    for (int t = n_sel + threadIdx.x; t < p2; t += blockDim.x) sortbuf[t] = 0ull;
    __syncthreads();
    for (int size = 2; size <= p2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = threadIdx.x; t < (p2 >> 1); t += blockDim.x) {
          const int lo = (t / stride) * (stride << 1) + (t % stride), hi = lo + stride;
          const bool desc = ((lo & size) == 0);
          const unsigned long long x = sortbuf[lo], y = sortbuf[hi];
          if ((x < y) == desc) {
            sortbuf[lo] = y;
            sortbuf[hi] = x;
          }
        }
        __syncthreads();
      }
    }
    for (int t = threadIdx.x; t < n_sel; t += blockDim.x) {
      const int c = 0x7fffffff - (int)(unsigned)(sortbuf[t] & 0xffffffffull);
      const int row = rows[c];
      ob[t] = boxes[row];
      os[t] = scores[row];
      ov[t] = 1;
    }
  }
  for (int t = n_sel + threadIdx.x; t < a.W; t += blockDim.x) {
    ob[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    os[t] = -1.f;
    ov[t] = 0;
  }
  // add_gt_proposals (inference.py:53-74): the ground-truth boxes with objectness 1
  const int g_n = a.gmax > 0 ? gt_count[i] : 0;
  for (int g = threadIdx.x; g < a.gmax; g += blockDim.x) {
    const bool ok = g < g_n;
    ob[a.W + g] = ok ? gt[(size_t)i * a.gmax + g] : make_float4(0.f, 0.f, 0.f, 0.f);
    os[a.W + g] = ok ? 1.f : 0.f;
    ov[a.W + g] = ok ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------ IoU (boxlist_ops.py:53-89)
__device__ __forceinline__ float box_area1(const float4 b) {   // bounding_box.py:214-226, TO_REMOVE = 1
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

__device__ __forceinline__ float box_iou1(const float4 g, float g_area, const float4 p, float p_area) {
  const float lx = fmaxf(g.x, p.x), ly = fmaxf(g.y, p.y), rx = fminf(g.z, p.z), ry = fminf(g.w, p.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(rx, lx), 1.f), 0.f), h = fmaxf(__fadd_rn(__fsub_rn(ry, ly), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(g_area, p_area), inter));
}

// ------------------------------------------------------------------------------------------ 3. ROI-head assign + sample
struct AssignArgs {
  int N, P, gmax, S, pos_cap, mask_m;
  float fg_thr, bg_thr, wx, wy, ww, wh;
};

__global__ void __launch_bounds__(kGlueThreads)
roi_assign_sample_kernel(const float4* __restrict__ boxes, const unsigned char* __restrict__ valid, const float* __restrict__ rnd,
                         const float4* __restrict__ gt, const int64_t* __restrict__ gt_labels, const int32_t* __restrict__ gt_count,
                         float* __restrict__ out_rois, int64_t* __restrict__ out_labels, float4* __restrict__ out_reg,
                         int64_t* __restrict__ out_gidx, float* __restrict__ m_rois, int64_t* __restrict__ m_labels,
                         float* __restrict__ m_w, int64_t* __restrict__ m_gidx, int64_t* __restrict__ out_index, AssignArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int hist[256], ghist[256], sh[4], scan_ws[33];
  unsigned* keys = (unsigned*)smem_raw;          // [P] inverted random key (larger = earlier in the permutation)
  int* lab = (int*)(smem_raw + (size_t)a.P * 4);     // [P] label: -1 ignore, 0 background, >0 class
  int* mid = (int*)(smem_raw + (size_t)a.P * 8);     // [P] matched gt (clamped at 0)
  int* sel_of = (int*)(smem_raw + (size_t)a.P * 12); // [S] sampled row of every output slot
  const int i = blockIdx.x;
  const int G = gt_count[i];
  const float4* bx = boxes + (size_t)i * a.P;
  const float4* gb = gt + (size_t)i * a.gmax;
  for (int p = threadIdx.x; p < a.P; p += blockDim.x) {
    const float4 b = bx[p];
    const float pa = box_area1(b);
    float best = -1.f;
    int bi = 0;
    for (int g = 0; g < G; ++g) {
      const float4 q = gb[g];
      const float v = box_iou1(q, box_area1(q), b, pa);
      if (v > best) {   // first maximum (torch.max over dim 0)
        best = v;
        bi = g;
      }
    }
    // matcher.py:60-81 (allow_low_quality_matches=False); box_head/loss.py:64-75
    int l;
    if (G == 0 || best < a.bg_thr) l = 0;
    else if (best < a.fg_thr) l = -1;
    else l = (int)gt_labels[(size_t)i * a.gmax + bi];
    if (!valid[(size_t)i * a.P + p]) l = -1;
    lab[p] = l;
    mid[p] = (G == 0 || best < a.fg_thr) ? 0 : bi;
    unsigned k = ~glue_key(rnd[(size_t)i * a.P + p]);
    keys[p] = k == 0 ? 1u : k;
  }
  __syncthreads();
  // positives: the pos_cap smallest keys among label > 0
  unsigned Tp, Tn;
  int qp, qn;
  {
    auto key_of = [&](int p, unsigned* k) { *k = keys[p]; return lab[p] > 0; };
    radix_select<false>(key_of, a.P, a.pos_cap, &Tp, &qp, hist, ghist, sh);
  }
  // every thread owns a contiguous run of proposals (index order == output order): one CTA-wide scan per quantity
  const int per = (a.P + (int)blockDim.x - 1) / (int)blockDim.x;
  const int p_lo = min((int)threadIdx.x * per, a.P), p_hi = min(p_lo + per, a.P);
  int n_pos;
  {
    int my_ties = 0;
    for (int p = p_lo; p < p_hi; ++p) my_ties += (lab[p] > 0 && keys[p] == Tp);
    int tie_tot;
    int tr = block_excl_scan(my_ties, &tie_tot, scan_ws);
    int my_sel = 0;
    for (int p = p_lo; p < p_hi; ++p) {
      if (lab[p] <= 0) continue;
      const bool is_tie = keys[p] == Tp;
      const bool sel = keys[p] > Tp || (is_tie && tr < qp);
      tr += is_tie;
      if (sel) ++my_sel;
      else lab[p] = -2 - lab[p];     // unsampled positive: remember the class as -(2 + class)
    }
    block_excl_scan(my_sel, &n_pos, scan_ws);
  }
  __syncthreads();
  const int neg_want = a.S - n_pos;
  {
    auto key_of = [&](int p, unsigned* k) { *k = keys[p]; return lab[p] == 0; };
    radix_select<false>(key_of, a.P, neg_want > 0 ? neg_want : 1, &Tn, &qn, hist, ghist, sh);
  }
  // two-way partition in index order: sampled rows first, then the rest; the first S rows are the output
  {
    int my_ties = 0;
    for (int p = p_lo; p < p_hi; ++p) my_ties += (lab[p] == 0 && keys[p] == Tn);
    int tie_tot;
    int tr = block_excl_scan(my_ties, &tie_tot, scan_ws);
    int my_sel = 0;
    unsigned selbits = 0;            // per <= 32 checked on the host
    for (int p = p_lo; p < p_hi; ++p) {
      const int l = lab[p];
      bool sel = l > 0;
      if (l == 0) {
        const bool is_tie = keys[p] == Tn;
        sel = neg_want > 0 && (keys[p] > Tn || (is_tie && tr < qn));
        tr += is_tie;
      }
      if (sel) {
        ++my_sel;
        selbits |= 1u << (p - p_lo);
      }
    }
    int packed_tot;                  // low 16 bits: sampled rows, high 16 bits: the rest
    const int packed = block_excl_scan(my_sel | ((p_hi - p_lo - my_sel) << 16), &packed_tot, scan_ws);
    const int total_sel = packed_tot & 0xffff;
    int ps = packed & 0xffff, pr = total_sel + (packed >> 16);
    for (int p = p_lo; p < p_hi; ++p) {
      const bool sel = (selbits >> (p - p_lo)) & 1u;
      const int slot = sel ? ps++ : pr++;
      if (slot < a.S) sel_of[slot] = sel ? p : -1 - p;
    }
  }
  __syncthreads();
  // emit the S rows (fewer than S proposals: the tail repeats row 0 as padding with label -1)
  for (int s = threadIdx.x; s < a.S; s += blockDim.x) {
    int p = s < a.P ? sel_of[s] : -1;
    const bool ok = p >= 0;
    if (!ok) p = -1 - p;
    const float4 b = bx[p];
    const int g = mid[p];
    const size_t o = (size_t)i * a.S + s;
    out_rois[o * 5 + 0] = (float)i;
    out_rois[o * 5 + 1] = b.x;
    out_rois[o * 5 + 2] = b.y;
    out_rois[o * 5 + 3] = b.z;
    out_rois[o * 5 + 4] = b.w;
    out_labels[o] = ok ? (int64_t)lab[p] : (int64_t)-1;
    out_gidx[o] = g;
    if (out_index) out_index[o] = p;
    // box_coder.py:22-50
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (G > 0) q = gb[g];
    const float ew = __fadd_rn(__fsub_rn(b.z, b.x), 1.f), eh = __fadd_rn(__fsub_rn(b.w, b.y), 1.f);
    const float ex = __fadd_rn(b.x, __fmul_rn(0.5f, ew)), ey = __fadd_rn(b.y, __fmul_rn(0.5f, eh));
    const float gw = __fadd_rn(__fsub_rn(q.z, q.x), 1.f), gh = __fadd_rn(__fsub_rn(q.w, q.y), 1.f);
    const float gx = __fadd_rn(q.x, __fmul_rn(0.5f, gw)), gy = __fadd_rn(q.y, __fmul_rn(0.5f, gh));
    float4 t;
    t.x = __fdiv_rn(__fmul_rn(a.wx, __fsub_rn(gx, ex)), ew);
    t.y = __fdiv_rn(__fmul_rn(a.wy, __fsub_rn(gy, ey)), eh);
    t.z = __fmul_rn(a.ww, logf(__fdiv_rn(gw, ew)));
    t.w = __fmul_rn(a.wh, logf(__fdiv_rn(gh, eh)));
    out_reg[o] = t;
  }
  // mask branch: the positives among the S rows first (mask_head.py:11-32), fixed width mask_m
  if (a.mask_m > 0) {
    __syncthreads();
    const int sper = (a.S + (int)blockDim.x - 1) / (int)blockDim.x;
    const int s_lo = min((int)threadIdx.x * sper, a.S), s_hi = min(s_lo + sper, a.S);
    int my_pos = 0;
    for (int sidx = s_lo; sidx < s_hi; ++sidx) {
      const int p = sidx < a.P ? sel_of[sidx] : -1;
      my_pos += (p >= 0 && lab[p] > 0);
    }
    int packed_tot;
    const int packed = block_excl_scan(my_pos | ((s_hi - s_lo - my_pos) << 16), &packed_tot, scan_ws);
    const int total_pos = packed_tot & 0xffff;
    int pp = packed & 0xffff, pr = total_pos + (packed >> 16);
    for (int sidx = s_lo; sidx < s_hi; ++sidx) {
      int p = sidx < a.P ? sel_of[sidx] : -1;
      const bool isp = p >= 0 && lab[p] > 0;
      if (p < 0) p = -1 - p;
      const int slot = isp ? pp++ : pr++;
      if (slot < a.mask_m) {
        const float4 b = bx[p];
        const size_t o = (size_t)i * a.mask_m + slot;
        m_rois[o * 5 + 0] = (float)i;
        m_rois[o * 5 + 1] = b.x;
        m_rois[o * 5 + 2] = b.y;
        m_rois[o * 5 + 3] = b.z;
        m_rois[o * 5 + 4] = b.w;
        m_labels[o] = isp ? (int64_t)lab[p] : (int64_t)0;
        m_w[o] = isp ? 1.f : 0.f;
        m_gidx[o] = mid[p];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ 3b. box-head post-processing
// PostProcessor.forward (modeling/roi_heads/box_head/inference.py:45-149): softmax, per-class BoxCoder.decode
// (box_coder.py:52-95), clip_to_image, score threshold -> one NMS problem per (image, class > 0) in the layout of
// mrb_nms_batched (P rows each, rows below the threshold / invalid proposals carry score -1 and an empty box) ...
struct PostArgs {
  int N, P, C, ld;                    // images, proposals per image, classes (incl. background), floats per row of `outputs`
  float thresh, wx, wy, ww, wh, clip;   // wx..wh: reciprocals of the box-coder weights
};

__global__ void __launch_bounds__(256)
box_post_decode_kernel(const float* __restrict__ outputs, const float4* __restrict__ proposals, const unsigned char* __restrict__ valid,
                       const float* __restrict__ im_w, const float* __restrict__ im_h, float4* __restrict__ boxes,
                       float* __restrict__ scores, PostArgs a) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // row = image * P + proposal
  if (r >= a.N * a.P) return;
  const int i = r / a.P, p = r - i * a.P;
  const float* row = outputs + (size_t)r * a.ld;
  float m = -INFINITY;
  for (int c = lane; c < a.C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  float ssum = 0.f;
  for (int c = lane; c < a.C; c += 32) ssum += expf(row[c] - m);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, d);
  const bool ok = valid[r] != 0;
  const float4 an = proposals[r];
  const float w = __fadd_rn(__fsub_rn(an.z, an.x), 1.f), h = __fadd_rn(__fsub_rn(an.w, an.y), 1.f);
  const float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
  const float lx = __fsub_rn(im_w[i], 1.f), ly = __fsub_rn(im_h[i], 1.f);
  for (int c = 1 + lane; c < a.C; c += 32) {
    const float prob = expf(row[c] - m) / ssum;
    const size_t o = ((size_t)(i * (a.C - 1) + (c - 1))) * a.P + p;
    if (ok && prob > a.thresh) {                                          // inference.py:108 (scores > score_thresh)
      const float* d = row + a.C + 4 * c;
      const float dx = __fmul_rn(d[0], a.wx), dy = __fmul_rn(d[1], a.wy);
      const float dw = fminf(__fmul_rn(d[2], a.ww), a.clip), dh = fminf(__fmul_rn(d[3], a.wh), a.clip);
      const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
      const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
      float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
      float x2 = __fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.f), y2 = __fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 1.f);
      boxes[o] = make_float4(fminf(fmaxf(x1, 0.f), lx), fminf(fmaxf(y1, 0.f), ly), fminf(fmaxf(x2, 0.f), lx), fminf(fmaxf(y2, 0.f), ly));
      scores[o] = prob;
    } else {
      boxes[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      scores[o] = -1.f;
    }
  }
}

// ... and, after the NMS of those problems, the detections_per_img best survivors of every image over all classes
// (inference.py:131-147), in (class, proposal) order.  One cluster per image, the classes split over its CTAs.
// Ties at the cut are taken in (class, proposal) order (the reference keeps every score >= the k-th value).
struct DetArgs {
  int N, P, C, max_det, cls_per_cta;
};

__global__ void __launch_bounds__(kGlueThreads)
box_post_select_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ keep,
                       const int32_t* __restrict__ counts, float4* __restrict__ out_b, float* __restrict__ out_s,
                       int64_t* __restrict__ out_l, int32_t* __restrict__ out_n, DetArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int hist[256], ghist[256], sh[4], scan_ws[33], cnt_sh[2];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned cs = cluster.num_blocks(), rank = cluster.block_rank();
  const int i = blockIdx.x / cs;
  unsigned* keys = (unsigned*)smem_raw;                 // [cls_per_cta * P]  0 = no candidate in the slot
  const int c_lo = min(1 + (int)rank * a.cls_per_cta, a.C), c_hi = min(c_lo + a.cls_per_cta, a.C);
  const int n = (c_hi - c_lo) * a.P;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int c = c_lo + t / a.P, q = t - (t / a.P) * a.P;
    const int prob = i * (a.C - 1) + (c - 1);
    unsigned key = 0;
    if (q < counts[prob]) {
      const int64_t rr = keep[(size_t)prob * a.P + q];
      const float sc = scores[(size_t)prob * a.P + (rr < 0 ? 0 : (rr >= a.P ? a.P - 1 : rr))];
      if (sc > 0.f) key = glue_key(sc);
    }
    keys[t] = key;
  }
  __syncthreads();
  unsigned T;
  int quota;
  auto key_of = [&](int t, unsigned* k) { *k = keys[t]; return *k != 0; };
  radix_select(key_of, n, a.max_det, &T, &quota, hist, ghist, sh);
  const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
  const int t_lo = min((int)threadIdx.x * per, n), t_hi = min(t_lo + per, n);
  int my_ties = 0;
  for (int t = t_lo; t < t_hi; ++t) my_ties += (keys[t] != 0 && keys[t] == T);
  int tie_tot;
  int tr = block_excl_scan(my_ties, &tie_tot, scan_ws);
  if (threadIdx.x == 0) cnt_sh[0] = tie_tot;
  cluster.sync();
  for (unsigned r = 0; r < rank; ++r) tr += cluster.map_shared_rank(cnt_sh, r)[0];
  int my_sel = 0;
  for (int t = t_lo; t < t_hi; ++t) {
    const unsigned key = keys[t];
    const bool is_tie = key != 0 && key == T;
    const bool sel = key != 0 && (key > T || (is_tie && tr < quota));
    tr += is_tie;
    my_sel += sel;
    if (!sel) keys[t] = 0;
  }
  int sel_tot;
  int pos = block_excl_scan(my_sel, &sel_tot, scan_ws);
  if (threadIdx.x == 0) cnt_sh[1] = sel_tot;
  cluster.sync();
  int total = 0;
  for (unsigned r = 0; r < cs; ++r) {
    const int c = cluster.map_shared_rank(cnt_sh, r)[1];
    if (r < rank) pos += c;
    total += c;
  }
  for (int t = t_lo; t < t_hi; ++t)
    if (keys[t]) {
      const int c = c_lo + t / a.P, q = t - (t / a.P) * a.P;
      const int prob = i * (a.C - 1) + (c - 1);
      const size_t row = (size_t)prob * a.P + keep[(size_t)prob * a.P + q];
      if (pos < a.max_det) {
        const size_t o = (size_t)i * a.max_det + pos;
        out_b[o] = boxes[row];
        out_s[o] = scores[row];
        out_l[o] = c;
      }
      ++pos;
    }
  if (rank == 0) {
    for (int t = total + threadIdx.x; t < a.max_det; t += blockDim.x) {
      const size_t o = (size_t)i * a.max_det + t;
      out_b[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      out_s[o] = 0.f;
      out_l[o] = 0;
    }
    if (threadIdx.x == 0) out_n[i] = total < a.max_det ? total : a.max_det;
  }
  cluster.sync();      // nobody leaves while its counters may still be read
}

// ------------------------------------------------------------------------------------------ 4. RPN anchor labelling
// pass 1: per anchor best IoU / argmax; per ground-truth box the best IoU over all anchors (atomicMax on the bit pattern of
// a non-negative float).  pass 2: labels.
__global__ void __launch_bounds__(256)
rpn_anchor_iou_kernel(const float4* __restrict__ anchors, const float4* __restrict__ gt, const int32_t* __restrict__ gt_count,
                      float* __restrict__ best_iou, int32_t* __restrict__ best_gt, unsigned* __restrict__ gt_best, int A, int gmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* sg = (float4*)smem_raw;                       // [gmax]
  float* sa = (float*)(smem_raw + (size_t)gmax * 16);   // [gmax] areas
  unsigned* sm = (unsigned*)(smem_raw + (size_t)gmax * 20);   // [gmax] CTA-local best per gt
  const int i = blockIdx.y;
  const int G = gt_count[i];
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float4 q = gt[(size_t)i * gmax + g];
    sg[g] = q;
    sa[g] = box_area1(q);
    sm[g] = 0u;
  }
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < A; p += gridDim.x * blockDim.x) {
    const float4 b = anchors[p];
    const float pa = box_area1(b);
    float best = -1.f;
    int bi = 0;
    for (int g = 0; g < G; ++g) {
      const float v = box_iou1(sg[g], sa[g], b, pa);
      if (v > best) {
        best = v;
        bi = g;
      }
      if (v > 0.f && __float_as_uint(v) > sm[g]) atomicMax(&sm[g], __float_as_uint(v));
    }
    best_iou[(size_t)i * A + p] = best;
    best_gt[(size_t)i * A + p] = bi;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x)
    if (sm[g]) atomicMax(&gt_best[(size_t)i * gmax + g], sm[g]);
}

__global__ void __launch_bounds__(256)
rpn_anchor_label_kernel(const float4* __restrict__ anchors, const float4* __restrict__ gt, const int32_t* __restrict__ gt_count,
                        const float* __restrict__ best_iou, int32_t* __restrict__ best_gt, const unsigned* __restrict__ gt_best,
                        const float* __restrict__ im_w, const float* __restrict__ im_h, float* __restrict__ labels, int A,
                        int gmax, float fg_thr, float bg_thr, float straddle) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* sg = (float4*)smem_raw;
  float* sa = (float*)(smem_raw + (size_t)gmax * 16);
  float* sb = (float*)(smem_raw + (size_t)gmax * 20);
  const int i = blockIdx.y;
  const int G = gt_count[i];
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float4 q = gt[(size_t)i * gmax + g];
    sg[g] = q;
    sa[g] = box_area1(q);
    sb[g] = __uint_as_float(gt_best[(size_t)i * gmax + g]);
  }
  __syncthreads();
  const float W = im_w[i], H = im_h[i];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < A; p += gridDim.x * blockDim.x) {
    const float4 b = anchors[p];
    const float pa = box_area1(b);
    const float best = best_iou[(size_t)i * A + p];
    // matcher.py:60-81, then set_low_quality_matches_ (:83-112): an anchor that attains some gt's best IoU keeps its argmax
    int m = best < bg_thr ? -1 : (best < fg_thr ? -2 : 0);
    if (m < 0) {
      for (int g = 0; g < G; ++g)
        if (box_iou1(sg[g], sa[g], b, pa) == sb[g]) {
          m = 0;
          break;
        }
    }
    // rpn/loss.py:66-84: label 1 for matches, 0 below the low threshold, -1 between thresholds or not visible
    float l = m == 0 ? 1.f : (m == -1 ? 0.f : -1.f);
    if (straddle >= 0.f) {   // anchor_generator.py:100-116
      const bool vis = b.x >= -straddle && b.y >= -straddle && b.z < W + straddle && b.w < H + straddle;
      if (!vis) l = -1.f;
    }
    labels[(size_t)i * A + p] = l;
  }
}

// ------------------------------------------------------------------------------------------ 5. RPN anchor sampling + encode
// BalancedPositiveNegativeSampler over the anchors of an image (balanced_positive_negative_sampler.py:19-68: up to
// batch * fraction random positives, the rest random negatives) and BoxCoder.encode of the sampled positives
// (rpn/loss.py:86-90, box_coder.py:22-50), for the whole batch in one launch.  One cluster of CTAs per image splits the
// ~268 k anchors; "n random elements" = the n smallest of caller-supplied iid keys (== randperm[:n]), found by a radix
// select across the cluster; the sampled anchors are emitted in anchor order.
struct SampleArgs {
  int N, A, slice, gmax, P, B;       // images, anchors, anchors per CTA, gt slots, positive cap, batch size per image
  float wx, wy, ww, wh;
};

__global__ void __launch_bounds__(kGlueThreads)
rpn_sample_kernel(const float* __restrict__ labels, const int32_t* __restrict__ matched, const float* __restrict__ rnd,
                  const float4* __restrict__ anchors, const float4* __restrict__ gt, int64_t* __restrict__ pos_idx,
                  unsigned char* __restrict__ pos_ok, float4* __restrict__ reg_t, int64_t* __restrict__ sel_idx,
                  float* __restrict__ sel_label, float* __restrict__ sel_w, SampleArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int hist[256], ghist[256], sh[4], scan_ws[33], cnt_sh[4];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned cs = cluster.num_blocks(), rank = cluster.block_rank();
  const int i = blockIdx.x / cs;
  unsigned* keys = (unsigned*)smem_raw;                              // [slice] inverted key (larger = earlier), bit 0 cleared
  signed char* lab = (signed char*)(smem_raw + (size_t)a.slice * 4);   // [slice] 1 positive, 0 negative, -1 ignore
  const int a_lo = min((int)rank * a.slice, a.A), a_hi = min(a_lo + a.slice, a.A), n = a_hi - a_lo;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const size_t g = (size_t)i * a.A + a_lo + t;
    unsigned k = ~glue_key(rnd[g]);
    keys[t] = k == 0 ? 1u : k;
    const float l = labels[g];
    lab[t] = l >= 1.f ? 1 : (l == 0.f ? 0 : -1);
  }
  __syncthreads();
  const int per = (n + (int)blockDim.x - 1) / (int)blockDim.x;
  const int t_lo = min((int)threadIdx.x * per, n), t_hi = min(t_lo + per, n);
  int64_t* pidx = pos_idx + (size_t)i * a.P;
  unsigned char* pok = pos_ok + (size_t)i * a.P;
  float4* preg = reg_t + (size_t)i * a.P;
  const size_t S = (size_t)a.P + a.B;
  int64_t* sidx = sel_idx + (size_t)i * S;
  float* slab = sel_label + (size_t)i * S;
  float* sw = sel_w + (size_t)i * S;
  int n_pos_total = 0;
  for (int phase = 0; phase < 2; ++phase) {
    const signed char want_lab = phase == 0 ? 1 : 0;
    const int cap = phase == 0 ? a.P : a.B;
    int want = phase == 0 ? a.P : a.B - n_pos_total;
    const bool none = want <= 0;
    if (none) want = 1;
    unsigned T;
    int quota;
    auto key_of = [&](int t, unsigned* k) { *k = keys[t]; return lab[t] == want_lab; };
    radix_select(key_of, n, want, &T, &quota, hist, ghist, sh);
    int my_ties = 0;
    for (int t = t_lo; t < t_hi; ++t) my_ties += (lab[t] == want_lab && keys[t] == T);
    int tie_tot;
    int tr = block_excl_scan(my_ties, &tie_tot, scan_ws);
    if (threadIdx.x == 0) cnt_sh[0] = tie_tot;
    cluster.sync();
    for (unsigned r = 0; r < rank; ++r) tr += cluster.map_shared_rank(cnt_sh, r)[0];
    int my_sel = 0;
    unsigned long long selbits = 0;             // per <= 64 (slice <= 64 * 1024, checked on the host)
    for (int t = t_lo; t < t_hi; ++t) {
      if (lab[t] != want_lab) continue;
      const bool is_tie = keys[t] == T;
      const bool sel = !none && (keys[t] > T || (is_tie && tr < quota));
      tr += is_tie;
      if (sel) {
        ++my_sel;
        selbits |= 1ull << (t - t_lo);
      }
    }
    int sel_tot;
    int pos = block_excl_scan(my_sel, &sel_tot, scan_ws);
    if (threadIdx.x == 0) cnt_sh[1] = sel_tot;
    cluster.sync();
    int total = 0;
    for (unsigned r = 0; r < cs; ++r) {
      const int c = cluster.map_shared_rank(cnt_sh, r)[1];
      if (r < rank) pos += c;
      total += c;
    }
    for (int t = t_lo; t < t_hi; ++t)
      if ((selbits >> (t - t_lo)) & 1ull) {
        const int g = a_lo + t;
        if (phase == 0 && pos < a.P) {
          pidx[pos] = g;
          pok[pos] = 1;
          sidx[pos] = g;
          slab[pos] = 1.f;
          sw[pos] = 1.f;
          // box_coder.py:22-50 with the RPN's weights
          const float4 b = anchors[g];
          const float4 q = gt[(size_t)i * a.gmax + matched[(size_t)i * a.A + g]];
          const float ew = __fadd_rn(__fsub_rn(b.z, b.x), 1.f), eh = __fadd_rn(__fsub_rn(b.w, b.y), 1.f);
          const float ex = __fadd_rn(b.x, __fmul_rn(0.5f, ew)), ey = __fadd_rn(b.y, __fmul_rn(0.5f, eh));
          const float gw = __fadd_rn(__fsub_rn(q.z, q.x), 1.f), gh = __fadd_rn(__fsub_rn(q.w, q.y), 1.f);
          const float gx = __fadd_rn(q.x, __fmul_rn(0.5f, gw)), gy = __fadd_rn(q.y, __fmul_rn(0.5f, gh));
          float4 tt;
          tt.x = __fdiv_rn(__fmul_rn(a.wx, __fsub_rn(gx, ex)), ew);
          tt.y = __fdiv_rn(__fmul_rn(a.wy, __fsub_rn(gy, ey)), eh);
          tt.z = __fmul_rn(a.ww, logf(__fdiv_rn(gw, ew)));
          tt.w = __fmul_rn(a.wh, logf(__fdiv_rn(gh, eh)));
          preg[pos] = tt;
        } else if (phase == 1 && pos < a.B) {
          sidx[a.P + pos] = g;
          slab[a.P + pos] = 0.f;
          sw[a.P + pos] = 1.f;
        }
        ++pos;
      }
    // padding of the list (rank 0): index 0, weight 0
    if (rank == 0) {
      const int filled = total < cap ? total : cap;
      for (int t = filled + threadIdx.x; t < cap; t += blockDim.x) {
        if (phase == 0) {
          pidx[t] = 0;
          pok[t] = 0;
          preg[t] = make_float4(0.f, 0.f, 0.f, 0.f);
          sidx[t] = 0;
          slab[t] = 1.f;
          sw[t] = 0.f;
        } else {
          sidx[a.P + t] = 0;
          slab[a.P + t] = 0.f;
          sw[a.P + t] = 0.f;
        }
      }
    }
    if (phase == 0) n_pos_total = total < a.P ? total : a.P;
    cluster.sync();        // cnt_sh is reused by the next phase; nobody leaves while its counters may still be read
  }
}

}  // namespace mrb
using namespace mrb;

MRB_API int mrb_rpn_decode(const float* logits, const float* deltas, const float* anchors, const int64_t* topk_idx,
                           const float* image_w, const float* image_h, float* boxes, float* scores, int num_images,
                           int num_anchors, int k, const float* weights_host, float xform_clip, mrb_stream_t stream) {
  if (num_images < 0 || num_anchors <= 0 || k < 0 || !weights_host) return MRB_ERR_BAD_ARG;
  if (num_images == 0 || k == 0) return MRB_OK;
  if (!logits || !deltas || !anchors || !topk_idx || !image_w || !image_h || !boxes || !scores) return MRB_ERR_BAD_ARG;
  if (((uintptr_t)deltas & 15) || ((uintptr_t)anchors & 15) || ((uintptr_t)boxes & 15)) return MRB_ERR_BAD_ARG;
  const int total = num_images * k;
  rpn_decode_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      logits, (const float4*)deltas, (const float4*)anchors, topk_idx, image_w, image_h, (float4*)boxes, scores, num_images,
      num_anchors, k, 1.f / weights_host[0], 1.f / weights_host[1], 1.f / weights_host[2], 1.f / weights_host[3], xform_clip, 0, 0);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_rpn_decode_packed(const float* head_output, int anchors_per_location, int pixel_stride, const float* anchors,
                                  const int64_t* topk_idx,
                                  const float* image_w, const float* image_h, float* boxes, float* scores, int num_images,
                                  int num_anchors, int k, const float* weights_host, float xform_clip, mrb_stream_t stream) {
  if (num_images < 0 || num_anchors <= 0 || k < 0 || !weights_host || anchors_per_location <= 0 ||
      num_anchors % anchors_per_location || pixel_stride < 5 * anchors_per_location)
    return MRB_ERR_BAD_ARG;
  if (num_images == 0 || k == 0) return MRB_OK;
  if (!head_output || !anchors || !topk_idx || !image_w || !image_h || !boxes || !scores) return MRB_ERR_BAD_ARG;
  if (((uintptr_t)anchors & 15) || ((uintptr_t)boxes & 15)) return MRB_ERR_BAD_ARG;
  const int total = num_images * k;
  rpn_decode_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      head_output, nullptr, (const float4*)anchors, topk_idx, image_w, image_h, (float4*)boxes, scores, num_images, num_anchors, k,
      1.f / weights_host[0], 1.f / weights_host[1], 1.f / weights_host[2], 1.f / weights_host[3], xform_clip, anchors_per_location, pixel_stride);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_rpn_topk_decode(const float* head_output, int anchors_per_location, int pixel_stride, const float* anchors,
                                const float* image_w, const float* image_h, float* boxes, float* scores, int num_images, int num_anchors,
                                int k, const float* weights_host, float xform_clip, mrb_stream_t stream) {
  if (num_images <= 0 || num_anchors <= 0 || k <= 0 || k > num_anchors || !weights_host || anchors_per_location <= 0 ||
      num_anchors % anchors_per_location || pixel_stride < 5 * anchors_per_location)
    return MRB_ERR_BAD_ARG;
  if (!head_output || !anchors || !image_w || !image_h || !boxes || !scores) return MRB_ERR_BAD_ARG;
  if (((uintptr_t)anchors & 15) || ((uintptr_t)boxes & 15)) return MRB_ERR_BAD_ARG;
  if (k > 8192) return MRB_ERR_UNSUPPORTED;
  // CTAs per image: the fewest (1, 2, 4, 8) that keep a slice's keys within ~100 KB of shared memory
  int cs = 1;
  while (cs < 8 && (num_anchors + cs - 1) / cs > 25600) cs <<= 1;
  TopkArgs a;
  a.N = num_images; a.A = num_anchors; a.k = k; a.apl = anchors_per_location; a.ld = pixel_stride;
  a.slice = (num_anchors + cs - 1) / cs;
  if (a.slice > 50000) return MRB_ERR_UNSUPPORTED;
  a.wx = 1.f / weights_host[0]; a.wy = 1.f / weights_host[1]; a.ww = 1.f / weights_host[2]; a.wh = 1.f / weights_host[3];
  a.clip = xform_clip;
  int p2 = 1;
  while (p2 < k) p2 <<= 1;
  const size_t smem = (((size_t)a.slice * 4 + 15) & ~(size_t)15) + (size_t)p2 * 8;
  if (smem > 220 * 1024) return MRB_ERR_UNSUPPORTED;
  MRB_CUDA_TRY(cudaFuncSetAttribute(rpn_topk_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_images * cs);
  cfg.blockDim = dim3(kGlueThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MRB_CUDA_TRY(cudaLaunchKernelEx(&cfg, rpn_topk_decode_kernel, head_output, (const float4*)anchors, image_w, image_h, (float4*)boxes,
                                  scores, a));
  return MRB_OK;
}

MRB_API int mrb_rpn_collect(const float* boxes, const float* scores, const int64_t* keep, const int32_t* num_keep,
                            const int* k_per_level_host, int num_levels, int num_images, int post_nms_top_n, int fpn_post_nms_top_n,
                            int per_batch, int sorted, const float* gt_boxes, const int32_t* gt_count, int gmax, float* out_boxes,
                            float* out_scores, uint8_t* out_valid, mrb_stream_t stream) {
  if (num_levels <= 0 || num_levels > kMaxLevels || num_images <= 0 || !k_per_level_host || gmax < 0) return MRB_ERR_BAD_ARG;
  if (!boxes || !scores || !keep || !num_keep || !out_boxes || !out_scores || !out_valid) return MRB_ERR_BAD_ARG;
  if (gmax > 0 && (!gt_boxes || !gt_count)) return MRB_ERR_BAD_ARG;
  if (((uintptr_t)boxes & 15) || ((uintptr_t)out_boxes & 15) || ((uintptr_t)gt_boxes & 15)) return MRB_ERR_BAD_ARG;
  CollectArgs a;
  a.L = num_levels;
  a.N = num_images;
  a.post_n = post_nms_top_n;
  a.gmax = gmax;
  a.per_batch = per_batch ? 1 : 0;
  a.sorted = sorted ? 1 : 0;
  int slot = 0, row = 0;
  for (int l = 0; l < num_levels; ++l) {
    a.k[l] = k_per_level_host[l];
    a.m[l] = a.k[l] < post_nms_top_n ? a.k[l] : post_nms_top_n;
    a.slot0[l] = slot;
    a.row0[l] = row;
    slot += a.m[l];
    row += num_images * a.k[l];
  }
  for (int l = num_levels; l < kMaxLevels; ++l) a.k[l] = a.m[l] = a.slot0[l] = a.row0[l] = 0;
  a.C = slot;
  const long long domain = a.per_batch ? (long long)slot * num_images : slot;
  a.topn = (int)(fpn_post_nms_top_n < domain ? fpn_post_nms_top_n : domain);
  a.W = a.topn;
  if (a.C <= 0 || a.topn <= 0) return MRB_ERR_BAD_ARG;
  if (a.per_batch && num_images > 8) return MRB_ERR_UNSUPPORTED;     // one CTA per image in one (portable) cluster
  if (a.sorted && a.per_batch) return MRB_ERR_UNSUPPORTED;
  int p2 = 1;
  while (p2 < a.W) p2 <<= 1;
  const size_t smem = (size_t)a.C * 8 + (a.sorted ? (size_t)p2 * 8 : 0);
  if (smem > 200 * 1024 || a.C >= 65536) return MRB_ERR_UNSUPPORTED;
  MRB_CUDA_TRY(cudaFuncSetAttribute(rpn_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_images);
  cfg.blockDim = dim3(kGlueThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.per_batch ? num_images : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MRB_CUDA_TRY(cudaLaunchKernelEx(&cfg, rpn_collect_kernel, (const float4*)boxes, scores, keep, num_keep, (const float4*)gt_boxes,
                                  gt_count, (float4*)out_boxes, out_scores, (unsigned char*)out_valid, a));
  return MRB_OK;
}

MRB_API int mrb_rpn_collect_width(const int* k_per_level_host, int num_levels, int num_images, int post_nms_top_n,
                                  int fpn_post_nms_top_n, int per_batch) {
  long long slot = 0;
  for (int l = 0; l < num_levels; ++l) slot += k_per_level_host[l] < post_nms_top_n ? k_per_level_host[l] : post_nms_top_n;
  const long long domain = per_batch ? slot * num_images : slot;
  return (int)(fpn_post_nms_top_n < domain ? fpn_post_nms_top_n : domain);
}

MRB_API int mrb_roi_assign_sample(const float* boxes, const uint8_t* valid, const float* rand_keys, const float* gt_boxes,
                                  const int64_t* gt_labels, const int32_t* gt_count, int num_images, int num_proposals, int gmax,
                                  int batch_size_per_image, float positive_fraction, float fg_iou, float bg_iou,
                                  const float* weights_host, int mask_rois_per_image, float* out_rois, int64_t* out_labels,
                                  float* out_reg_targets, int64_t* out_gt_index, float* mask_rois, int64_t* mask_labels,
                                  float* mask_weight, int64_t* mask_gt_index, int64_t* out_proposal_index, mrb_stream_t stream) {
  if (num_images <= 0 || num_proposals <= 0 || gmax <= 0 || batch_size_per_image <= 0 || !weights_host) return MRB_ERR_BAD_ARG;
  if (!boxes || !valid || !rand_keys || !gt_boxes || !gt_labels || !gt_count || !out_rois || !out_labels || !out_reg_targets ||
      !out_gt_index)
    return MRB_ERR_BAD_ARG;
  if (mask_rois_per_image > 0 && (!mask_rois || !mask_labels || !mask_weight || !mask_gt_index)) return MRB_ERR_BAD_ARG;
  if (mask_rois_per_image > batch_size_per_image) return MRB_ERR_BAD_ARG;
  if (((uintptr_t)boxes & 15) || ((uintptr_t)gt_boxes & 15) || ((uintptr_t)out_reg_targets & 15)) return MRB_ERR_BAD_ARG;
  AssignArgs a;
  a.N = num_images;
  a.P = num_proposals;
  a.gmax = gmax;
  a.S = batch_size_per_image;
  a.pos_cap = (int)(batch_size_per_image * positive_fraction);
  if (a.pos_cap > num_proposals) a.pos_cap = num_proposals;
  if (a.pos_cap < 1) return MRB_ERR_BAD_ARG;
  a.mask_m = mask_rois_per_image;
  a.fg_thr = fg_iou;
  a.bg_thr = bg_iou;
  a.wx = weights_host[0];
  a.wy = weights_host[1];
  a.ww = weights_host[2];
  a.wh = weights_host[3];
  const size_t smem = (size_t)a.P * 12 + (size_t)a.S * 4;
  if (smem > 200 * 1024 || num_proposals > 32 * kGlueThreads || num_proposals >= 65536) return MRB_ERR_UNSUPPORTED;
  MRB_CUDA_TRY(cudaFuncSetAttribute(roi_assign_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_images);
  cfg.blockDim = dim3(kGlueThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MRB_CUDA_TRY(cudaLaunchKernelEx(&cfg, roi_assign_sample_kernel, (const float4*)boxes, (const unsigned char*)valid, rand_keys,
                                  (const float4*)gt_boxes, gt_labels, gt_count, out_rois, out_labels, (float4*)out_reg_targets,
                                  out_gt_index, mask_rois, mask_labels, mask_weight, mask_gt_index, out_proposal_index, a));
  return MRB_OK;
}

MRB_API int mrb_box_post_decode(const float* outputs, int ld, int num_classes, const float* proposals, const uint8_t* valid,
                                const float* image_w, const float* image_h, int num_images, int proposals_per_image,
                                float score_thresh, const float* weights_host, float xform_clip, float* boxes, float* scores,
                                mrb_stream_t stream) {
  if (!outputs || !proposals || !valid || !image_w || !image_h || !weights_host || !boxes || !scores || num_images <= 0 ||
      proposals_per_image <= 0 || num_classes < 2 || ld < 5 * num_classes)
    return MRB_ERR_BAD_ARG;
  if (((uintptr_t)proposals & 15) || ((uintptr_t)boxes & 15)) return MRB_ERR_BAD_ARG;
  PostArgs a;
  a.N = num_images; a.P = proposals_per_image; a.C = num_classes; a.ld = ld; a.thresh = score_thresh;
  a.wx = 1.f / weights_host[0]; a.wy = 1.f / weights_host[1]; a.ww = 1.f / weights_host[2]; a.wh = 1.f / weights_host[3];
  a.clip = xform_clip;
  box_post_decode_kernel<<<ceil_div((int64_t)num_images * proposals_per_image, 8), 256, 0, (cudaStream_t)stream>>>(
      outputs, (const float4*)proposals, (const unsigned char*)valid, image_w, image_h, (float4*)boxes, scores, a);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}

MRB_API int mrb_box_post_select(const float* boxes, const float* scores, const int64_t* keep, const int32_t* num_keep, int num_images,
                                int proposals_per_image, int num_classes, int detections_per_img, float* out_boxes, float* out_scores,
                                int64_t* out_labels, int32_t* out_count, mrb_stream_t stream) {
  if (!boxes || !scores || !keep || !num_keep || !out_boxes || !out_scores || !out_labels || !out_count || num_images <= 0 ||
      proposals_per_image <= 0 || num_classes < 2 || detections_per_img <= 0)
    return MRB_ERR_BAD_ARG;
  if (((uintptr_t)boxes & 15) || ((uintptr_t)out_boxes & 15)) return MRB_ERR_BAD_ARG;
  DetArgs a;
  a.N = num_images; a.P = proposals_per_image; a.C = num_classes; a.max_det = detections_per_img;
  int cs = 1;
  while (cs < 8 && (size_t)((num_classes - 1 + cs - 1) / cs) * proposals_per_image * 4 > 96 * 1024) cs <<= 1;
  a.cls_per_cta = (num_classes - 1 + cs - 1) / cs;
  const size_t smem = (size_t)a.cls_per_cta * proposals_per_image * 4;
  if (smem > 200 * 1024) return MRB_ERR_UNSUPPORTED;
  MRB_CUDA_TRY(cudaFuncSetAttribute(box_post_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_images * cs);
  cfg.blockDim = dim3(kGlueThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MRB_CUDA_TRY(cudaLaunchKernelEx(&cfg, box_post_select_kernel, (const float4*)boxes, scores, keep, num_keep, (float4*)out_boxes,
                                  out_scores, out_labels, out_count, a));
  return MRB_OK;
}

MRB_API int mrb_rpn_sample(const float* labels, const int32_t* matched_gt, const float* rand_keys, const float* anchors,
                           const float* gt_boxes, int num_images, int num_anchors, int gmax, int batch_size_per_image,
                           float positive_fraction, const float* weights_host, int64_t* pos_idx, uint8_t* pos_ok, float* reg_targets,
                           int64_t* sel_idx, float* sel_label, float* sel_weight, mrb_stream_t stream) {
  if (!labels || !matched_gt || !rand_keys || !anchors || !gt_boxes || !weights_host || !pos_idx || !pos_ok || !reg_targets ||
      !sel_idx || !sel_label || !sel_weight || num_images <= 0 || num_anchors <= 0 || gmax <= 0 || batch_size_per_image <= 0)
    return MRB_ERR_BAD_ARG;
  if (((uintptr_t)anchors & 15) || ((uintptr_t)gt_boxes & 15) || ((uintptr_t)reg_targets & 15)) return MRB_ERR_BAD_ARG;
  SampleArgs a;
  a.N = num_images; a.A = num_anchors; a.gmax = gmax; a.B = batch_size_per_image;
  a.P = (int)(batch_size_per_image * positive_fraction);
  if (a.P < 1 || a.P > batch_size_per_image) return MRB_ERR_BAD_ARG;
  a.wx = weights_host[0]; a.wy = weights_host[1]; a.ww = weights_host[2]; a.wh = weights_host[3];
  int cs = 1;
  while (cs < 8 && (num_anchors + cs - 1) / cs > 36 * 1024) cs <<= 1;
  a.slice = (num_anchors + cs - 1) / cs;
  if (a.slice > 40000) return MRB_ERR_UNSUPPORTED;   // 5 B of shared memory per anchor; 16-CTA clusters are not portable
  const size_t smem = (size_t)a.slice * 5 + 16;
  if (smem > 200 * 1024) return MRB_ERR_UNSUPPORTED;
  MRB_CUDA_TRY(cudaFuncSetAttribute(rpn_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_images * cs);
  cfg.blockDim = dim3(kGlueThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MRB_CUDA_TRY(cudaLaunchKernelEx(&cfg, rpn_sample_kernel, labels, matched_gt, rand_keys, (const float4*)anchors, (const float4*)gt_boxes,
                                  pos_idx, (unsigned char*)pos_ok, (float4*)reg_targets, sel_idx, sel_label, sel_weight, a));
  return MRB_OK;
}

MRB_API size_t mrb_rpn_anchor_match_workspace_bytes(int num_images, int num_anchors, int gmax) {
  return (size_t)num_images * num_anchors * 8 + (size_t)num_images * (gmax > 0 ? gmax : 1) * 4 + 64;
}

MRB_API int mrb_rpn_anchor_match(const float* anchors, const float* gt_boxes, const int32_t* gt_count, const float* image_w,
                                 const float* image_h, int num_images, int num_anchors, int gmax, float fg_iou, float bg_iou,
                                 float straddle_thresh, float* labels, int32_t* matched_gt, void* workspace, size_t workspace_bytes,
                                 mrb_stream_t stream) {
  if (num_images <= 0 || num_anchors <= 0 || gmax <= 0) return MRB_ERR_BAD_ARG;
  if (!anchors || !gt_boxes || !gt_count || !image_w || !image_h || !labels || !matched_gt || !workspace) return MRB_ERR_BAD_ARG;
  if (workspace_bytes < mrb_rpn_anchor_match_workspace_bytes(num_images, num_anchors, gmax)) return MRB_ERR_WORKSPACE;
  if (((uintptr_t)anchors & 15) || ((uintptr_t)gt_boxes & 15) || ((uintptr_t)workspace & 15)) return MRB_ERR_BAD_ARG;
  float* best_iou = (float*)workspace;
  unsigned* gt_best = (unsigned*)((unsigned char*)workspace + (size_t)num_images * num_anchors * 4);
  const size_t smem = (size_t)gmax * 24;
  if (smem > 48 * 1024) return MRB_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  MRB_CUDA_TRY(cudaMemsetAsync(gt_best, 0, (size_t)num_images * gmax * 4, st));
  const int gx = grid_for(num_anchors, 256, 4, 1);
  rpn_anchor_iou_kernel<<<dim3(gx, num_images), 256, smem, st>>>((const float4*)anchors, (const float4*)gt_boxes, gt_count, best_iou,
                                                                  matched_gt, gt_best, num_anchors, gmax);
  MRB_LAUNCH_CHECK();
  rpn_anchor_label_kernel<<<dim3(gx, num_images), 256, smem, st>>>((const float4*)anchors, (const float4*)gt_boxes, gt_count, best_iou,
                                                                    matched_gt, gt_best, image_w, image_h, labels, num_anchors, gmax,
                                                                    fg_iou, bg_iou, straddle_thresh);
  MRB_LAUNCH_CHECK();
  return MRB_OK;
}
