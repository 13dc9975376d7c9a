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
