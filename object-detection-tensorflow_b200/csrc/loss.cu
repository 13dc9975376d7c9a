// RetinaNet training-loss forward (BASELINE config 3 "focal-loss + decode + NMS"):
// anchor/GT matching, softmax focal loss and smooth-L1 box loss, forward only.
// Restates RetinaNet.py:357-474 (never copied):
//   * GT rows (y,x,h,w,id) padded with -1; valid count = index of first -1 in col 0
//   * IoU[G,A]; for every GT its arg-max anchor is a positive ("best", duplicates kept);
//     the remaining anchors: max IoU > 0.5 -> positive with the arg-max GT,
//     max IoU < 0.4 -> negative (label = background = last class), else ignored
//   * focal: p = clip(softmax(logits)[label], 1e-8, 1); -alpha (1-p)^gamma log p,
//     the same alpha for positives and negatives; sum / #positives
//   * box: smooth-L1 of (t_yx - (g_yx-a_yx)/a_hw) and (t_hw - log(g_hw/a_hw)), mean over positives
// Sums are reduced in a fixed order (per-block partials, then a serial final
// pass) so the result is run-to-run deterministic.
#include "tail_common.cuh"

namespace odt {

constexpr int kLossBlocks = 64;   // anchor blocks per image
constexpr int kLossThreads = 256;
constexpr int kMaxGT = 128;

__device__ __forceinline__ float iou_match(const Anchor& a, float gy1, float gx1, float gy2,
                                           float gx2, float garea) {
  // ref RetinaNet.py:383-388
  float iy1 = fmaxf(a.y1, gy1), ix1 = fmaxf(a.x1, gx1);
  float iy2 = fminf(a.y2, gy2), ix2 = fminf(a.x2, gx2);
  float inter = __fmul_rn(fmaxf(__fsub_rn(iy2, iy1), 0.f), fmaxf(__fsub_rn(ix2, ix1), 0.f));
  float aarea = __fmul_rn(a.h, a.w);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, garea), inter));
}

__device__ __forceinline__ int gt_count(const float* gt, int G) {
  // tf.argmin(gt, axis=0)[0]: first index of the column minimum (ref :359)
  int arg = 0;
  float best = gt[0];
  for (int i = 1; i < G; ++i) {
    float v = gt[i * 5];
    if (v < best) {
      best = v;
      arg = i;
    }
  }
  return arg;
}

// K1: per (gt, image) arg-max anchor (first maximum)
__global__ void __launch_bounds__(kLossThreads)
    loss_best_anchor_kernel(const __grid_constant__ TailP tp, const float* __restrict__ gt, int G,
                            int* __restrict__ best) {
  pdl_launch_dependents();
  const odt_tail_params& p = tp.p;
  const int g = blockIdx.x, b = blockIdx.y;
  const float* gb = gt + (long long)b * G * 5;
  __shared__ int s_cnt;
  __shared__ float s_v[kLossThreads];
  __shared__ int s_i[kLossThreads];
  if (threadIdx.x == 0) s_cnt = gt_count(gb, G);
  __syncthreads();
  if (g >= s_cnt) {
    if (threadIdx.x == 0) best[b * G + g] = -1;
    return;
  }
  const float gy = gb[g * 5 + 0], gx = gb[g * 5 + 1], gh = gb[g * 5 + 2], gw = gb[g * 5 + 3];
  const float hh = __fmul_rn(gh, 0.5f), hw = __fmul_rn(gw, 0.5f);
  const float gy1 = __fsub_rn(gy, hh), gx1 = __fsub_rn(gx, hw);
  const float gy2 = __fadd_rn(gy, hh), gx2 = __fadd_rn(gx, hw);
  const float garea = __fmul_rn(gh, gw);
  float bv = -1.f;
  int bi = 0x7fffffff;
  for (int n = threadIdx.x; n < p.N; n += blockDim.x) {
    Cell c = locate(p, n);
    Anchor a = anchor_ssd(p, c);
    float v = iou_match(a, gy1, gx1, gy2, gx2, garea);
    if (v > bv) {
      bv = v;
      bi = n;
    }
  }
  s_v[threadIdx.x] = bv;
  s_i[threadIdx.x] = bi;
  __syncthreads();
  for (int o = kLossThreads / 2; o; o >>= 1) {
    if (threadIdx.x < o) {
      float v2 = s_v[threadIdx.x + o];
      int i2 = s_i[threadIdx.x + o];
      if (v2 > s_v[threadIdx.x] || (v2 == s_v[threadIdx.x] && i2 < s_i[threadIdx.x])) {
        s_v[threadIdx.x] = v2;
        s_i[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) best[b * G + g] = s_i[0];
}

__device__ __forceinline__ float smooth_l1(float x) {
  float ax = fabsf(x);
  return ax < 1.f ? __fmul_rn(__fmul_rn(0.5f, x), x) : __fsub_rn(ax, 0.5f);
}

__device__ __forceinline__ void softmax21(const float* r, float* prob) {
  float m = r[0];
#pragma unroll
  for (int i = 1; i < 21; ++i) m = fmaxf(m, r[i]);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 21; ++i) {
    prob[i] = expf(__fsub_rn(r[i], m));
    s = __fadd_rn(s, prob[i]);
  }
  float inv = __fdiv_rn(1.f, s);
#pragma unroll
  for (int i = 0; i < 21; ++i) prob[i] = __fmul_rn(prob[i], inv);
}

__device__ __forceinline__ float focal(float pr, float alpha, float gamma) {
  float q = fminf(fmaxf(pr, 1e-8f), 1.f);
  return -alpha * powf(1.f - q, gamma) * logf(q);
}

__device__ __forceinline__ float coord_loss(const float* r, const Anchor& a, float gy, float gx,
                                            float gh, float gw) {
  float ty = __fdiv_rn(__fsub_rn(gy, a.cy), a.h), tx = __fdiv_rn(__fsub_rn(gx, a.cx), a.w);
  float th = logf(__fdiv_rn(gh, a.h)), tw = logf(__fdiv_rn(gw, a.w));
  float lyx = __fadd_rn(smooth_l1(__fsub_rn(r[21], ty)), smooth_l1(__fsub_rn(r[22], tx)));
  float lhw = __fadd_rn(smooth_l1(__fsub_rn(r[23], th)), smooth_l1(__fsub_rn(r[24], tw)));
  return __fadd_rn(lyx, lhw);
}

// K2: per-anchor assignment + loss partial sums. partial[b][blk] = (conf, coord, npos)
__global__ void __launch_bounds__(kLossThreads)
    loss_anchor_kernel(const float* __restrict__ head, const __grid_constant__ TailP tp,
                       const float* __restrict__ gt, int G, const int* __restrict__ best,
                       float alpha, float gamma, float* __restrict__ partial) {
  pdl_launch_dependents();
  const odt_tail_params& p = tp.p;
  const int b = blockIdx.y, blk = blockIdx.x;
  const float* gb = gt + (long long)b * G * 5;
  __shared__ float s_gt[kMaxGT][6];  // y1,x1,y2,x2,area,(unused)
  __shared__ int s_best[kMaxGT];
  __shared__ int s_cnt;
  __shared__ float s_red[3][kLossThreads];
  if (threadIdx.x == 0) s_cnt = gt_count(gb, G);
  __syncthreads();
  const int cnt = s_cnt;
  for (int g = threadIdx.x; g < cnt; g += blockDim.x) {
    float gy = gb[g * 5], gx = gb[g * 5 + 1], gh = gb[g * 5 + 2], gw = gb[g * 5 + 3];
    float hh = __fmul_rn(gh, 0.5f), hw = __fmul_rn(gw, 0.5f);
    s_gt[g][0] = __fsub_rn(gy, hh);
    s_gt[g][1] = __fsub_rn(gx, hw);
    s_gt[g][2] = __fadd_rn(gy, hh);
    s_gt[g][3] = __fadd_rn(gx, hw);
    s_gt[g][4] = __fmul_rn(gh, gw);
    s_best[g] = best[b * G + g];
  }
  __syncthreads();
  const float* hb = head + (long long)b * p.N * kRow;
  float conf = 0.f, coord = 0.f, npos = 0.f;
  const int per = (p.N + kLossBlocks - 1) / kLossBlocks;
  const int n_begin = blk * per, n_end = min(p.N, n_begin + per);
  for (int n = n_begin + threadIdx.x; n < n_end; n += blockDim.x) {
    bool is_best = false;
    for (int g = 0; g < cnt; ++g) is_best |= (s_best[g] == n);
    if (is_best) continue;  // handled by the per-GT pass below (ref :397-411)
    Cell c = locate(p, n);
    Anchor a = anchor_ssd(p, c);
    float bv = -1.f;
    int bg = 0;
    for (int g = 0; g < cnt; ++g) {
      float v = iou_match(a, s_gt[g][0], s_gt[g][1], s_gt[g][2], s_gt[g][3], s_gt[g][4]);
      if (v > bv) {
        bv = v;
        bg = g;
      }
    }
    const bool pos = bv > 0.5f, neg = bv < 0.4f;
    if (!pos && !neg) continue;
    const float* r = hb + (long long)n * kRow;
    float prob[21];
    softmax21(r, prob);
    if (pos) {
      int label = (int)gb[bg * 5 + 4];
      conf += focal(prob[label], alpha, gamma);
      coord += coord_loss(r, a, gb[bg * 5], gb[bg * 5 + 1], gb[bg * 5 + 2], gb[bg * 5 + 3]);
      npos += 1.f;
    } else {
      conf += focal(prob[20], alpha, gamma);
    }
  }
  // the G "best anchor" positives (duplicates kept), once per image (block 0)
  if (blk == 0) {
    for (int g = threadIdx.x; g < cnt; g += blockDim.x) {
      int n = s_best[g];
      Cell c = locate(p, n);
      Anchor a = anchor_ssd(p, c);
      const float* r = hb + (long long)n * kRow;
      float prob[21];
      softmax21(r, prob);
      int label = (int)gb[g * 5 + 4];
      conf += focal(prob[label], alpha, gamma);
      coord += coord_loss(r, a, gb[g * 5], gb[g * 5 + 1], gb[g * 5 + 2], gb[g * 5 + 3]);
      npos += 1.f;
    }
  }
  s_red[0][threadIdx.x] = conf;
  s_red[1][threadIdx.x] = coord;
  s_red[2][threadIdx.x] = npos;
  __syncthreads();
  for (int o = kLossThreads / 2; o; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int q = 0; q < 3; ++q) s_red[q][threadIdx.x] += s_red[q][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) partial[((long long)b * kLossBlocks + blk) * 4 + threadIdx.x] = s_red[threadIdx.x][0];
}

__global__ void loss_final_kernel(const float* __restrict__ partial, int B, float* __restrict__ out) {
  pdl_launch_dependents();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float conf = 0.f, coord = 0.f, npos = 0.f;
  for (int i = 0; i < kLossBlocks; ++i) {
    const float* q = partial + ((long long)b * kLossBlocks + i) * 4;
    conf += q[0];
    coord += q[1];
    npos += q[2];
  }
  out[b] = conf / npos + coord / npos;
}

// ============================ SSD loss forward ==============================
// Restates SSD300.py:345-453 / SSD512.py (never copied):
//   * matching as above but with ONE threshold: other anchors with best IoU > 0.5 are
//     positives, all the rest are negatives;
//   * cross-entropy (sparse_softmax_cross_entropy = logsumexp(x) - x[label]), positives
//     averaged; smooth-L1 box term averaged over the positives;
//   * hard-negative mining: tf.image.non_max_suppression over the negative ANCHOR boxes
//     (yx -+ hw/2 of the anchor's centre form) scored by their background cross-entropy,
//     max_output = min(#neg, 3 * #pos), IoU 0.7; the loss is the mean of the selected ones.
__device__ __forceinline__ float xent21(const float* r, int label) {
  float m = r[0];
#pragma unroll
  for (int i = 1; i < 21; ++i) m = fmaxf(m, r[i]);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 21; ++i) s = __fadd_rn(s, expf(__fsub_rn(r[i], m)));
  return __fsub_rn(logf(s), __fsub_rn(r[label], m));
}

// K2': per-anchor assignment; positives -> partial sums, negatives -> (loss, row) keys
__global__ void __launch_bounds__(kLossThreads)
    ssd_loss_anchor_kernel(const float* __restrict__ head, const __grid_constant__ TailP tp,
                           const float* __restrict__ gt, int G, const int* __restrict__ best,
                           float* __restrict__ partial, unsigned long long* __restrict__ neg_keys,
                           int* __restrict__ neg_count) {
  pdl_launch_dependents();
  const odt_tail_params& p = tp.p;
  const int b = blockIdx.y, blk = blockIdx.x;
  const float* gb = gt + (long long)b * G * 5;
  __shared__ float s_gt[kMaxGT][6];
  __shared__ int s_best[kMaxGT];
  __shared__ int s_cnt;
  __shared__ float s_red[3][kLossThreads];
  if (threadIdx.x == 0) s_cnt = gt_count(gb, G);
  __syncthreads();
  const int cnt = s_cnt;
  for (int g = threadIdx.x; g < cnt; g += blockDim.x) {
    float gy = gb[g * 5], gx = gb[g * 5 + 1], gh = gb[g * 5 + 2], gw = gb[g * 5 + 3];
    float hh = __fmul_rn(gh, 0.5f), hw = __fmul_rn(gw, 0.5f);
    s_gt[g][0] = __fsub_rn(gy, hh);
    s_gt[g][1] = __fsub_rn(gx, hw);
    s_gt[g][2] = __fadd_rn(gy, hh);
    s_gt[g][3] = __fadd_rn(gx, hw);
    s_gt[g][4] = __fmul_rn(gh, gw);
    s_best[g] = best[b * G + g];
  }
  __syncthreads();
  const float* hb = head + (long long)b * p.N * kRow;
  unsigned long long* nk = neg_keys + (long long)b * p.N;
  float conf = 0.f, coord = 0.f, npos = 0.f;
  const int per = (p.N + kLossBlocks - 1) / kLossBlocks;
  const int n_begin = blk * per, n_end = min(p.N, n_begin + per);
  for (int n0 = n_begin; n0 < n_end; n0 += blockDim.x) {
    const int n = n0 + threadIdx.x;
    bool is_neg = false;
    float nloss = 0.f;
    if (n < n_end) {
      bool is_best = false;
      for (int g = 0; g < cnt; ++g) is_best |= (s_best[g] == n);
      if (!is_best) {
        Cell c = locate(p, n);
        Anchor a = anchor_ssd(p, c);
        float bv = -1.f;
        int bg = 0;
        for (int g = 0; g < cnt; ++g) {
          float v = iou_match(a, s_gt[g][0], s_gt[g][1], s_gt[g][2], s_gt[g][3], s_gt[g][4]);
          if (v > bv) {
            bv = v;
            bg = g;
          }
        }
        const float* r = hb + (long long)n * kRow;
        if (bv > 0.5f) {
          conf += xent21(r, (int)gb[bg * 5 + 4]);
          coord += coord_loss(r, a, gb[bg * 5], gb[bg * 5 + 1], gb[bg * 5 + 2], gb[bg * 5 + 3]);
          npos += 1.f;
        } else {
          is_neg = true;
          nloss = xent21(r, 20);
        }
      }
    }
    // warp-aggregated append of the negatives (order is irrelevant: keys are unique)
    const unsigned mask = __ballot_sync(0xffffffffu, is_neg);
    if (mask) {
      const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&neg_count[b], __popc(mask));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (is_neg)
        nk[base + __popc(mask & ((1u << lane) - 1))] =
            ((unsigned long long)__float_as_uint(nloss) << 32) | (0xFFFFFFFFu - (unsigned)n);
    }
  }
  if (blk == 0) {  // the G "best anchor" positives (duplicates kept)
    for (int g = threadIdx.x; g < cnt; g += blockDim.x) {
      int n = s_best[g];
      Cell c = locate(p, n);
      Anchor a = anchor_ssd(p, c);
      const float* r = hb + (long long)n * kRow;
      conf += xent21(r, (int)gb[g * 5 + 4]);
      coord += coord_loss(r, a, gb[g * 5], gb[g * 5 + 1], gb[g * 5 + 2], gb[g * 5 + 3]);
      npos += 1.f;
    }
  }
  s_red[0][threadIdx.x] = conf;
  s_red[1][threadIdx.x] = coord;
  s_red[2][threadIdx.x] = npos;
  __syncthreads();
  for (int o = kLossThreads / 2; o; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int q = 0; q < 3; ++q) s_red[q][threadIdx.x] += s_red[q][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) partial[((long long)b * kLossBlocks + blk) * 4 + threadIdx.x] = s_red[threadIdx.x][0];
}

// negative anchor box in the reference's op order: yx -+ hw/2 of the centre form (:419)
__device__ __forceinline__ float4 neg_anchor_box(const odt_tail_params& p, int n) {
  Cell c = locate(p, n);
  Anchor a = anchor_ssd(p, c);
  const float hh = __fdiv_rn(a.h, 2.f), hw = __fdiv_rn(a.w, 2.f);
  return make_float4(__fsub_rn(a.cy, hh), __fsub_rn(a.cx, hw), __fadd_rn(a.cy, hh), __fadd_rn(a.cx, hw));
}

// K3': one CTA per image: hard-negative mining (exact greedy NMS by repeated arg-max, as in
// tail.cu) + the final combination.  loss_out[b] = mean(selected neg) + conf/npos + coord/npos
__global__ void __launch_bounds__(kLossThreads)
    ssd_loss_mine_kernel(const __grid_constant__ TailP tp, const float* __restrict__ partial,
                         unsigned long long* __restrict__ neg_keys, const int* __restrict__ neg_count,
                         float* __restrict__ loss_out, int* __restrict__ info_out) {
  pdl_launch_dependents();
  const odt_tail_params& p = tp.p;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ unsigned long long s_wkey[kLossThreads / 32];
  __shared__ int s_wpos[kLossThreads / 32];
  float conf = 0.f, coord = 0.f, nposf = 0.f;
  for (int i = 0; i < kLossBlocks; ++i) {
    const float* q = partial + ((long long)b * kLossBlocks + i) * 4;
    conf += q[0];
    coord += q[1];
    nposf += q[2];
  }
  const int npos = (int)nposf, cnt = neg_count[b];
  const int chosen = cnt > 3 * npos ? 3 * npos : cnt;
  unsigned long long* keys = neg_keys + (long long)b * p.N;
  double neg_sum = 0.0;  // thread 0 only
  int nsel = 0;
  while (nsel < chosen) {
    unsigned long long bk = 0ull;
    int bp = 0x7fffffff;
    for (int i = tid; i < cnt; i += blockDim.x) {
      const unsigned long long k = keys[i];
      if (k > bk) {
        bk = k;
        bp = i;
      }
    }
    {
      const unsigned hi = (unsigned)(bk >> 32), lo = (unsigned)bk;
      const unsigned whi = __reduce_max_sync(0xffffffffu, hi);
      const unsigned wlo = __reduce_max_sync(0xffffffffu, hi == whi ? lo : 0u);
      const int wpos = __reduce_min_sync(0xffffffffu, (hi == whi && lo == wlo) ? bp : 0x7fffffff);
      if (lane == 0) {
        s_wkey[warp] = ((unsigned long long)whi << 32) | wlo;
        s_wpos[warp] = wpos;
      }
    }
    __syncthreads();
    unsigned long long hk = 0ull;
    int hp = -1;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) {
      if (s_wkey[w] > hk) {
        hk = s_wkey[w];
        hp = s_wpos[w];
      }
    }
    if (hk == 0ull) break;  // every negative is selected or suppressed
    const int hn = (int)(0xFFFFFFFFu - (unsigned)(hk & 0xFFFFFFFFull));
    const float4 cur = neg_anchor_box(p, hn);
    if (tid == 0) {
      neg_sum += (double)__uint_as_float((unsigned)(hk >> 32));
      keys[hp] = 0ull;
    }
    ++nsel;
    if (nsel >= chosen) break;
    for (int j = tid; j < cnt; j += blockDim.x) {
      const unsigned long long kj = keys[j];
      if (kj == 0ull || j == hp) continue;
      const int n = (int)(0xFFFFFFFFu - (unsigned)(kj & 0xFFFFFFFFull));
      if (iou_tf(neg_anchor_box(p, n), cur) > 0.7f) keys[j] = 0ull;
    }
    __syncthreads();
  }
  if (tid == 0) {
    loss_out[b] = (float)(neg_sum / (double)nsel) + conf / nposf + coord / nposf;
    if (info_out) {
      info_out[b * 3 + 0] = npos;
      info_out[b * 3 + 1] = cnt;
      info_out[b * 3 + 2] = nsel;
    }
  }
}

// ============================ FCOS loss forward =============================
// Restates FCOS.py:153-187 (level assignment) and :266-348 (per-level loss), never copied:
//   * GT rows go to pyramid levels by sqrt(h*w) with inclusive, overlapping bounds
//     (<=64 | 64..128 | 128..256 | 256..512 | >=512); a level without GT contributes 0;
//   * per location (grid without +0.5, stride units): inside-box mask per GT; the regression
//     target is the inside GT of minimal area (ties: element-wise max of l,r,t,b);
//   * IoU loss -log(iou + 1e-12) over inside locations, centre-ness BCE-with-logits over ALL
//     locations of the level, sigmoid focal loss (alpha .25, gamma 2) on the per-class inside mask;
//   * level loss = (iou + focal + centre) / #(positive class cells); image loss = sum of levels.
constexpr int kFcosLevels = 5;

__device__ __forceinline__ float softplus_f(float x) {  // log(1 + exp(x)), stable
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

__global__ void __launch_bounds__(kLossThreads)
    fcos_loss_kernel(const float* __restrict__ head, const __grid_constant__ TailP tp,
                     const float* __restrict__ gt, int G, float* __restrict__ partial,
                     int* __restrict__ level_cnt) {
  pdl_launch_dependents();
  const odt_tail_params& p = tp.p;
  const int b = blockIdx.y, blk = blockIdx.x;
  const float* gb = gt + (long long)b * G * 5;
  __shared__ float s_box[kMaxGT][kFcosLevels][4];  // y1,x1,y2,x2 in stride units of each level
  __shared__ int s_cls[kMaxGT];
  __shared__ int s_lv[kMaxGT];                      // level membership bits
  __shared__ int s_cnt;
  __shared__ float s_red[kLossThreads];
  if (threadIdx.x == 0) s_cnt = gt_count(gb, G);
  __syncthreads();
  const int cnt = s_cnt;
  for (int g = threadIdx.x; g < cnt; g += blockDim.x) {
    const float gy = gb[g * 5], gx = gb[g * 5 + 1], gh = gb[g * 5 + 2], gw = gb[g * 5 + 3];
    const float size = __fsqrt_rn(__fmul_rn(gh, gw));
    int bits = 0;
    if (size <= 64.f) bits |= 1;
    if (size >= 64.f && size <= 128.f) bits |= 2;
    if (size >= 128.f && size <= 256.f) bits |= 4;
    if (size >= 256.f && size <= 512.f) bits |= 8;
    if (size >= 512.f) bits |= 16;
    s_lv[g] = bits;
    s_cls[g] = (int)gb[g * 5 + 4];
    for (int l = 0; l < kFcosLevels; ++l) {
      const float st = p.level[l].out_mul;
      const float y = __fdiv_rn(gy, st), x = __fdiv_rn(gx, st), h = __fdiv_rn(gh, st), w = __fdiv_rn(gw, st);
      const float hh = __fdiv_rn(h, 2.f), hw = __fdiv_rn(w, 2.f);
      s_box[g][l][0] = __fsub_rn(y, hh);
      s_box[g][l][1] = __fsub_rn(x, hw);
      s_box[g][l][2] = __fadd_rn(y, hh);
      s_box[g][l][3] = __fadd_rn(x, hw);
    }
  }
  __syncthreads();
  if (blk == 0 && threadIdx.x < kFcosLevels) {  // GT count per level (a level without GT contributes nothing)
    int c = 0;
    for (int g = 0; g < cnt; ++g) c += (s_lv[g] >> threadIdx.x) & 1;
    level_cnt[b * kFcosLevels + threadIdx.x] = c;
  }
  const float* hb = head + (long long)b * p.N * kRow;
  float acc[kFcosLevels][4];  // iou, focal, centre, #positive class cells
#pragma unroll
  for (int l = 0; l < kFcosLevels; ++l)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[l][q] = 0.f;
  const int per = (p.N + kLossBlocks - 1) / kLossBlocks;
  const int n_begin = blk * per, n_end = min(p.N, n_begin + per);
  for (int n = n_begin + threadIdx.x; n < n_end; n += blockDim.x) {
    const Cell c = locate(p, n);
    const int l = c.lvl;
    const float yy = (float)c.y, xx = (float)c.x;
    bool loc = false;
    float amin = 3.0e38f, dl = 0.f, dr = 0.f, dt = 0.f, db = 0.f;
    unsigned cmask = 0u;
    for (int g = 0; g < cnt; ++g) {
      if (!((s_lv[g] >> l) & 1)) continue;
      const float l_ = __fsub_rn(xx, s_box[g][l][1]), r_ = __fsub_rn(s_box[g][l][3], xx);
      const float t_ = __fsub_rn(yy, s_box[g][l][0]), b_ = __fsub_rn(s_box[g][l][2], yy);
      if (!(l_ > 0.f && r_ > 0.f && t_ > 0.f && b_ > 0.f)) continue;
      loc = true;
      cmask |= 1u << s_cls[g];
      const float area = __fmul_rn(__fadd_rn(l_, r_), __fadd_rn(t_, b_));
      if (area < amin) {
        amin = area;
        dl = l_; dr = r_; dt = t_; db = b_;
      } else if (area == amin) {  // tied minimal areas: element-wise maximum (:300-307)
        dl = fmaxf(dl, l_); dr = fmaxf(dr, r_); dt = fmaxf(dt, t_); db = fmaxf(db, b_);
      }
    }
    const float* r = hb + (long long)n * kRow;
    // centre-ness BCE with logits, every location (:331-334); target 0 where no GT covers the cell
    const float lrmin = fminf(dl, dr), tbmin = fminf(dt, db), lrmax = fmaxf(dl, dr), tbmax = fmaxf(dt, db);
    const float cgt = __fsqrt_rn(__fdiv_rn(__fmul_rn(lrmin, tbmin), __fadd_rn(__fmul_rn(lrmax, tbmax), 1e-12f)));
    const float cx = r[20];
    float centre = fmaxf(cx, 0.f) - cx * cgt + log1pf(expf(-fabsf(cx)));
    float iou_l = 0.f;
    if (loc) {
      const float pl = expf(r[21]), pr = expf(r[22]), pt = expf(r[23]), pb = expf(r[24]);
      const float iw = __fadd_rn(fminf(dl, pl), fminf(dr, pr)), ih = __fadd_rn(fminf(dt, pt), fminf(db, pb));
      const float inter = __fmul_rn(iw, ih);
      const float uni = __fsub_rn(__fadd_rn(__fmul_rn(__fadd_rn(dl, dr), __fadd_rn(dt, db)),
                                            __fmul_rn(__fadd_rn(pl, pr), __fadd_rn(pt, pb))), inter);
      const float iou = __fdiv_rn(inter, __fadd_rn(uni, 1e-12f));
      iou_l = -logf(__fadd_rn(iou, 1e-12f));
    }
    float focal_l = 0.f, npos = 0.f;
#pragma unroll 4
    for (int k = 0; k < 20; ++k) {
      const float x = r[k];
      const float ls = -softplus_f(-x);                 // log sigmoid(x)
      const float sg = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));
      if ((cmask >> k) & 1u) {
        const float om = 1.f - sg;
        focal_l += -0.25f * om * om * ls;
        npos += 1.f;
      } else {
        focal_l += -0.25f * sg * sg * (-x + ls);
      }
    }
#pragma unroll
    for (int q = 0; q < kFcosLevels; ++q) {
      if (q == l) {
        acc[q][0] += iou_l;
        acc[q][1] += focal_l;
        acc[q][2] += centre;
        acc[q][3] += npos;
      }
    }
  }
  // fixed-order block reduction of the 5 x 4 sums
  for (int l = 0; l < kFcosLevels; ++l) {
    for (int q = 0; q < 4; ++q) {
      s_red[threadIdx.x] = acc[l][q];
      __syncthreads();
      for (int o = kLossThreads / 2; o; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
      }
      if (threadIdx.x == 0) partial[(((long long)b * kLossBlocks + blk) * kFcosLevels + l) * 4 + q] = s_red[0];
      __syncthreads();
    }
  }
}

__global__ void fcos_loss_final_kernel(const float* __restrict__ partial, const int* __restrict__ level_cnt,
                                       int B, float* __restrict__ out) {
  pdl_launch_dependents();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float total = 0.f;
  for (int l = 0; l < kFcosLevels; ++l) {
    if (level_cnt[b * kFcosLevels + l] == 0) continue;
    double s3[4] = {0, 0, 0, 0};
    for (int i = 0; i < kLossBlocks; ++i)
      for (int q = 0; q < 4; ++q) s3[q] += (double)partial[(((long long)b * kLossBlocks + i) * kFcosLevels + l) * 4 + q];
    total += (float)((s3[0] + s3[1] + s3[2]) / s3[3]);
  }
  out[b] = total;
}

// ============================ YOLOv3 loss forward ===========================
// Restates YOLOv3.py:115-318 (+ :37-41 priors, :419-433 anchors, :435-442 GT normalisation), never copied.
// Quirks kept: level k divides the GT by 32 / 16 / 8 while its priors are priors[k] / (8, 16, 32)[k];
// the GT x anchor intersections are products of UNCLAMPED differences; the no-object anchors are built
// from (yx - hw/2, yx + hw/2) re-used as (centre, size).
__device__ __forceinline__ float sig_xent(float z, float x) {  // sigmoid_cross_entropy_with_logits(labels=z, logits=x)
  return fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float iou_unclamped(float gy1, float gx1, float gy2, float gx2, float ay1, float ax1,
                                               float ay2, float ax2, float aarea) {
  const float ih = __fsub_rn(fminf(gy2, ay2), fmaxf(gy1, ay1)), iw = __fsub_rn(fminf(gx2, ax2), fmaxf(gx1, ax1));
  const float inter = __fmul_rn(ih, iw);
  const float garea = __fmul_rn(__fsub_rn(gy2, gy1), __fsub_rn(gx2, gx1));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, garea), inter));
}
constexpr int kYoloLevels = 3;

// K1: one block per image, one thread per GT: level / prior assignment, positive terms, cell occupancy
__global__ void __launch_bounds__(kMaxGT)
    yolo_loss_pos_kernel(const float* __restrict__ head, const __grid_constant__ TailP tp,
                         const float* __restrict__ gt, int G, unsigned char* __restrict__ occ,
                         float* __restrict__ pos_out) {
  pdl_launch_dependents();
  const odt_tail_params& p = tp.p;
  const int b = blockIdx.x, g = threadIdx.x;
  const float* gb = gt + (long long)b * G * 5;
  __shared__ int s_cnt;
  __shared__ float s_red[3][kMaxGT];
  if (threadIdx.x == 0) s_cnt = gt_count(gb, G);
  __syncthreads();
  const int cnt = s_cnt;
  const float norm[kYoloLevels] = {32.f, 16.f, 8.f};
  const long long cells = p.N / 3;
  float coord = 0.f, cls = 0.f, obj = 0.f;
  if (g < cnt) {
    float mx[kYoloLevels];
    int arg[kYoloLevels], cy[kYoloLevels], cx[kYoloLevels];
    float ny[kYoloLevels], nx[kYoloLevels], nh[kYoloLevels], nw[kYoloLevels];
#pragma unroll
    for (int k = 0; k < kYoloLevels; ++k) {
      const odt_level& L = p.level[k];
      ny[k] = __fdiv_rn(gb[g * 5], norm[k]);
      nx[k] = __fdiv_rn(gb[g * 5 + 1], norm[k]);
      nh[k] = __fdiv_rn(gb[g * 5 + 2], norm[k]);
      nw[k] = __fdiv_rn(gb[g * 5 + 3], norm[k]);
      cy[k] = min(max((int)floorf(ny[k]), 0), L.H - 1);
      cx[k] = min(max((int)floorf(nx[k]), 0), L.W - 1);
      occ[(long long)b * cells + L.offset / 3 + cy[k] * L.W + cx[k]] = 1;  // every GT marks its cell on every level
      const float gy1 = __fsub_rn(ny[k], __fdiv_rn(nh[k], 2.f)), gx1 = __fsub_rn(nx[k], __fdiv_rn(nw[k], 2.f));
      const float gy2 = __fadd_rn(ny[k], __fdiv_rn(nh[k], 2.f)), gx2 = __fadd_rn(nx[k], __fdiv_rn(nw[k], 2.f));
      const float ay = (float)cy[k] + 0.5f, ax = (float)cx[k] + 0.5f;
      float bv = 0.f;
      int bi = 0;
      for (int a = 0; a < 3; ++a) {
        const float hh = __fdiv_rn(L.prior_h[a], 2.f), hw = __fdiv_rn(L.prior_w[a], 2.f);
        const float v = iou_unclamped(gy1, gx1, gy2, gx2, __fsub_rn(ay, hh), __fsub_rn(ax, hw), __fadd_rn(ay, hh),
                                      __fadd_rn(ax, hw), __fmul_rn(L.prior_h[a], L.prior_w[a]));
        if (a == 0 || v > bv) {
          bv = v;
          bi = a;
        }
      }
      mx[k] = bv;
      arg[k] = bi;
    }
    int k = 2;
    if (mx[0] > mx[1] && mx[0] > mx[2]) k = 0;
    else if (mx[1] > mx[0] && mx[1] > mx[2]) k = 1;
    const odt_level& L = p.level[k];
    const float* r = head + ((long long)b * p.N + L.offset + (long long)(cy[k] * L.W + cx[k]) * 3 + arg[k]) * kRow;
    const float ty = __fsub_rn(ny[k], floorf(ny[k])), tx = __fsub_rn(nx[k], floorf(nx[k]));
    const float th = logf(__fdiv_rn(nh[k], L.prior_h[arg[k]])), tw = logf(__fdiv_rn(nw[k], L.prior_w[arg[k]]));
    coord = sig_xent(ty, r[20]) + sig_xent(tx, r[21]);
    const float dh = __fsub_rn(r[22], th), dw = __fsub_rn(r[23], tw);
    coord += 0.5f * (dh * dh + dw * dw);
    const int label = (int)gb[g * 5 + 4];
    for (int c = 0; c < 20; ++c) cls += sig_xent(c == label ? 1.f : 0.f, r[c]);
    obj = sig_xent(1.f, r[24]);
  }
  s_red[0][threadIdx.x] = coord;
  s_red[1][threadIdx.x] = cls;
  s_red[2][threadIdx.x] = obj;
  __syncthreads();
  for (int o = kMaxGT / 2; o; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int q = 0; q < 3; ++q) s_red[q][threadIdx.x] += s_red[q][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) pos_out[b * 4 + threadIdx.x] = s_red[threadIdx.x][0];
  if (threadIdx.x == 3) pos_out[b * 4 + 3] = (float)cnt;
}

// K2: no-object term over every (cell, prior) whose cell holds no GT centre
__global__ void __launch_bounds__(kLossThreads)
    yolo_loss_noobj_kernel(const float* __restrict__ head, const __grid_constant__ TailP tp,
                           const float* __restrict__ gt, int G, const unsigned char* __restrict__ occ,
                           float* __restrict__ partial) {
  pdl_launch_dependents();
  const odt_tail_params& p = tp.p;
  const int b = blockIdx.y, blk = blockIdx.x;
  const float* gb = gt + (long long)b * G * 5;
  __shared__ float s_g[kMaxGT][kYoloLevels][4];
  __shared__ int s_cnt;
  __shared__ float s_red[kLossThreads];
  if (threadIdx.x == 0) s_cnt = gt_count(gb, G);
  __syncthreads();
  const int cnt = s_cnt;
  const float norm[kYoloLevels] = {32.f, 16.f, 8.f};
  for (int g = threadIdx.x; g < cnt; g += blockDim.x) {
#pragma unroll
    for (int k = 0; k < kYoloLevels; ++k) {
      const float ny = __fdiv_rn(gb[g * 5], norm[k]), nx = __fdiv_rn(gb[g * 5 + 1], norm[k]);
      const float nh = __fdiv_rn(gb[g * 5 + 2], norm[k]), nw = __fdiv_rn(gb[g * 5 + 3], norm[k]);
      s_g[g][k][0] = __fsub_rn(ny, __fdiv_rn(nh, 2.f));
      s_g[g][k][1] = __fsub_rn(nx, __fdiv_rn(nw, 2.f));
      s_g[g][k][2] = __fadd_rn(ny, __fdiv_rn(nh, 2.f));
      s_g[g][k][3] = __fadd_rn(nx, __fdiv_rn(nw, 2.f));
    }
  }
  __syncthreads();
  const long long cells = p.N / 3;
  const float* hb = head + (long long)b * p.N * kRow;
  float acc = 0.f;
  const int per = (p.N + kLossBlocks - 1) / kLossBlocks;
  const int n_begin = blk * per, n_end = min(p.N, n_begin + per);
  for (int n = n_begin + threadIdx.x; n < n_end; n += blockDim.x) {
    const Cell c = locate(p, n);
    const odt_level& L = p.level[c.lvl];
    if (occ[(long long)b * cells + L.offset / 3 + c.y * L.W + c.x]) continue;
    const float ay = (float)c.y + 0.5f, ax = (float)c.x + 0.5f;
    const float ph = L.prior_h[c.a], pw = L.prior_w[c.a];
    const float cyy = __fsub_rn(ay, __fdiv_rn(ph, 2.f)), cxx = __fsub_rn(ax, __fdiv_rn(pw, 2.f));  // "centre" = y1x1
    const float sh = __fadd_rn(ay, __fdiv_rn(ph, 2.f)), sw = __fadd_rn(ax, __fdiv_rn(pw, 2.f));   // "size"   = y2x2
    const float by1 = __fsub_rn(cyy, __fdiv_rn(sh, 2.f)), bx1 = __fsub_rn(cxx, __fdiv_rn(sw, 2.f));
    const float by2 = __fadd_rn(cyy, __fdiv_rn(sh, 2.f)), bx2 = __fadd_rn(cxx, __fdiv_rn(sw, 2.f));
    const float aarea = __fmul_rn(__fsub_rn(by2, by1), __fsub_rn(bx2, bx1));
    float best = -3.0e38f;
    for (int g = 0; g < cnt; ++g)
      best = fmaxf(best, iou_unclamped(s_g[g][c.lvl][0], s_g[g][c.lvl][1], s_g[g][c.lvl][2], s_g[g][c.lvl][3], by1,
                                       bx1, by2, bx2, aarea));
    if (best <= 0.5f) acc += sig_xent(0.f, hb[(long long)n * kRow + 24]);
  }
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kLossThreads / 2; o; o >>= 1) {
    if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(long long)b * kLossBlocks + blk] = s_red[0];
}

__global__ void yolo_loss_final_kernel(const float* __restrict__ pos, const float* __restrict__ partial, int B,
                                       float cs, float ns, float os, float ks, float* __restrict__ out) {
  pdl_launch_dependents();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double noobj = 0.0;
  for (int i = 0; i < kLossBlocks; ++i) noobj += (double)partial[(long long)b * kLossBlocks + i];
  const float ng = pos[b * 4 + 3];
  out[b] = (cs * pos[b * 4 + 0] + ks * pos[b * 4 + 1] + os * pos[b * 4 + 2]) / ng + ns * (float)noobj / ng;
}

}  // namespace odt

using namespace odt;

extern "C" long long odt_retina_loss_scratch_floats(int B) {
  return B > 0 ? (long long)B * kLossBlocks * 4 : -1;
}

extern "C" int odt_retina_loss_fwd(const float* head, const odt_tail_params* p, int B,
                                   const float* gt, int G, float alpha, float gamma,
                                   float* partial_scratch, int* match_scratch, float* loss_out,
                                   void* stream) {
  ODT_CHECK_ARG(head && p && gt && partial_scratch && match_scratch && loss_out, "null pointer");
  ODT_CHECK_ARG(p->kind == ODT_DECODE_SSD, "softmax-family head expected");
  ODT_CHECK_ARG(B > 0 && G > 0 && G <= kMaxGT, "B/G (G <= 128)");
  cudaStream_t st = (cudaStream_t)stream;
  TailP tp;
  tp.p = *p;
  loss_best_anchor_kernel<<<dim3(G, B), kLossThreads, 0, st>>>(tp, gt, G, match_scratch);
  ODT_LAUNCH_OK();
  loss_anchor_kernel<<<dim3(kLossBlocks, B), kLossThreads, 0, st>>>(head, tp, gt, G, match_scratch,
                                                                  alpha, gamma, partial_scratch);
  ODT_LAUNCH_OK();
  loss_final_kernel<<<(B + 63) / 64, 64, 0, st>>>(partial_scratch, B, loss_out);
  ODT_LAUNCH_OK();
  return ODT_OK;
}

// ---- SSD ----
static long long ssd_loss_layout(const odt_tail_params* p, int B, long long* o_match, long long* o_count,
                                 long long* o_info, long long* o_keys) {
  long long off = (long long)B * kLossBlocks * 4 * 4;             // partial sums (floats)
  *o_match = off;  off += (long long)B * kMaxGT * 4;              // best anchor per GT (ints)
  *o_count = off;  off += (((long long)B + 1) & ~1ll) * 4;        // negatives per image
  *o_info = off;   off += (((long long)B * 3 + 1) & ~1ll) * 4;    // (#pos, #neg, #selected) per image
  off = (off + 7) & ~7ll;
  *o_keys = off;   off += (long long)B * p->N * 8;                // negative keys (u64)
  return off;
}

extern "C" long long odt_ssd_loss_scratch_bytes(const odt_tail_params* p, int B) {
  if (!p || B <= 0 || p->N <= 0) return -1;
  long long a, b2, c, d;
  return ssd_loss_layout(p, B, &a, &b2, &c, &d);
}

extern "C" long long odt_ssd_loss_info_offset(const odt_tail_params* p, int B) {
  if (!p || B <= 0 || p->N <= 0) return -1;
  long long a, b2, c, d;
  ssd_loss_layout(p, B, &a, &b2, &c, &d);
  return c;
}

extern "C" int odt_ssd_loss_fwd(const float* head, const odt_tail_params* p, int B, const float* gt, int G,
                                void* scratch, float* loss_out, void* stream) {
  ODT_CHECK_ARG(head && p && gt && scratch && loss_out, "null pointer");
  ODT_CHECK_ARG(p->kind == ODT_DECODE_SSD, "softmax-family head expected");
  ODT_CHECK_ARG(B > 0 && G > 0 && G <= kMaxGT, "B/G (G <= 128)");
  ODT_CHECK_ARG(((uintptr_t)scratch & 7) == 0, "scratch must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  long long o_match, o_count, o_info, o_keys;
  ssd_loss_layout(p, B, &o_match, &o_count, &o_info, &o_keys);
  char* sb = static_cast<char*>(scratch);
  float* partial = reinterpret_cast<float*>(sb);
  int* match = reinterpret_cast<int*>(sb + o_match);
  int* count = reinterpret_cast<int*>(sb + o_count);
  int* info = reinterpret_cast<int*>(sb + o_info);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(sb + o_keys);
  ODT_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)B, st));
  TailP tp;
  tp.p = *p;
  loss_best_anchor_kernel<<<dim3(G, B), kLossThreads, 0, st>>>(tp, gt, G, match);
  ODT_LAUNCH_OK();
  ssd_loss_anchor_kernel<<<dim3(kLossBlocks, B), kLossThreads, 0, st>>>(head, tp, gt, G, match, partial, keys,
                                                                      count);
  ODT_LAUNCH_OK();
  ssd_loss_mine_kernel<<<B, kLossThreads, 0, st>>>(tp, partial, keys, count, loss_out, info);
  ODT_LAUNCH_OK();
  return ODT_OK;
}

// ---- FCOS ----
extern "C" long long odt_fcos_loss_scratch_bytes(int B) {
  return B > 0 ? (long long)B * kLossBlocks * kFcosLevels * 4 * 4 + (long long)B * kFcosLevels * 4 : -1;
}

extern "C" int odt_fcos_loss_fwd(const float* head, const odt_tail_params* p, int B, const float* gt, int G,
                                 void* scratch, float* loss_out, void* stream) {
  ODT_CHECK_ARG(head && p && gt && scratch && loss_out, "null pointer");
  ODT_CHECK_ARG(p->kind == ODT_DECODE_FCOS && p->num_levels == kFcosLevels && p->num_fg == 20,
                "FCOS head with 5 levels and 20 classes expected");
  ODT_CHECK_ARG(B > 0 && G > 0 && G <= kMaxGT, "B/G (G <= 128)");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = static_cast<float*>(scratch);
  int* level_cnt = reinterpret_cast<int*>(partial + (long long)B * kLossBlocks * kFcosLevels * 4);
  TailP tp;
  tp.p = *p;
  fcos_loss_kernel<<<dim3(kLossBlocks, B), kLossThreads, 0, st>>>(head, tp, gt, G, partial, level_cnt);
  ODT_LAUNCH_OK();
  fcos_loss_final_kernel<<<(B + 63) / 64, 64, 0, st>>>(partial, level_cnt, B, loss_out);
  ODT_LAUNCH_OK();
  return ODT_OK;
}

// ---- YOLOv3 ----
extern "C" long long odt_yolo_loss_scratch_bytes(const odt_tail_params* p, int B) {
  if (!p || B <= 0 || p->N <= 0) return -1;
  return (long long)B * 4 * 4 + (long long)B * kLossBlocks * 4 + (((long long)B * (p->N / 3) + 3) & ~3ll);
}

extern "C" int odt_yolo_loss_fwd(const float* head, const odt_tail_params* p, int B, const float* gt, int G,
                                 float coord_scale, float noobj_scale, float obj_scale, float class_scale,
                                 void* scratch, float* loss_out, void* stream) {
  ODT_CHECK_ARG(head && p && gt && scratch && loss_out, "null pointer");
  ODT_CHECK_ARG(p->kind == ODT_DECODE_YOLO3 && p->num_levels == kYoloLevels && p->num_fg == 20,
                "YOLOv3 head with 3 levels and 20 classes expected");
  for (int i = 0; i < kYoloLevels; ++i) ODT_CHECK_ARG(p->level[i].A == 3, "3 priors per level expected");
  ODT_CHECK_ARG(B > 0 && G > 0 && G <= kMaxGT, "B/G (G <= 128)");
  cudaStream_t st = (cudaStream_t)stream;
  float* pos = static_cast<float*>(scratch);
  float* partial = pos + (long long)B * 4;
  unsigned char* occ = reinterpret_cast<unsigned char*>(partial + (long long)B * kLossBlocks);
  ODT_CUDA_OK(cudaMemsetAsync(occ, 0, (size_t)B * (p->N / 3), st));
  TailP tp;
  tp.p = *p;
  yolo_loss_pos_kernel<<<B, kMaxGT, 0, st>>>(head, tp, gt, G, occ, pos);
  ODT_LAUNCH_OK();
  yolo_loss_noobj_kernel<<<dim3(kLossBlocks, B), kLossThreads, 0, st>>>(head, tp, gt, G, occ, partial);
  ODT_LAUNCH_OK();
  yolo_loss_final_kernel<<<(B + 63) / 64, 64, 0, st>>>(pos, partial, B, coord_scale, noobj_scale, obj_scale,
                                                       class_scale, loss_out);
  ODT_LAUNCH_OK();
  return ODT_OK;
}
