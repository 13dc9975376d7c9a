// Inference tail of the five detectors on sm_100a: score activation, analytic
// anchor/prior generation, box decode, thresholding, candidate compaction and
// exact per-class TF NonMaxSuppressionV3 (iterative arg-max + IoU suppress).
//
// All arithmetic that feeds a discrete decision (class id, keep index) is done
// with explicitly rounded fp32 intrinsics (__fmul_rn/__fadd_rn/__fdiv_rn), in
// the operation order of the reference graphs, so no FMA contraction can change
// a result relative to the TF1.13 CPU kernels.
//
// Reference call sites restated here (never copied): SSD300.py:157-190,323-343;
// RetinaNet.py:224-256,328-355; YOLOv3.py:320-368,419-433; FCOS.py:130-150,197-264.
#include <stdlib.h>

#include "tail_common.cuh"

namespace odt {

constexpr int kDecodeWarps = 8;

// One warp = 32 consecutive candidate rows (3200 contiguous bytes), staged
// through shared memory with 128-bit loads, then one lane per row.
__global__ void __launch_bounds__(kDecodeWarps * 32)
    decode_candidates_kernel(const float* __restrict__ head, const __grid_constant__ TailP tp,
                             long long total_rows, unsigned long long* __restrict__ cand_keys,
                             int* __restrict__ cand_count) {
  pdl_launch_dependents();
  const odt_tail_params& p = tp.p;
  __shared__ __align__(16) float srow[kDecodeWarps][32 * kRow];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* my = srow[warp];
  const long long warps_total = (long long)gridDim.x * kDecodeWarps;
  const long long groups = (total_rows + 31) / 32;
  // Software pipeline per warp: the 128-bit loads of the NEXT 32-row group are issued into registers
  // before the current group (already staged in shared memory) is decoded, so every warp keeps
  // ~3 KB in flight while it computes.
  constexpr int kVecPerLane = (32 * kRow / 4 + 31) / 32;  // 7
  float4 pre[kVecPerLane];
  auto fetch = [&](long long gg) {
    const long long r0 = gg * 32;
    const int nr = (int)min((long long)32, total_rows - r0);
    const int nv = nr * kRow / 4;
    const float4* s4 = reinterpret_cast<const float4*>(head + r0 * kRow);
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) pre[i] = __ldcs(s4 + idx);
    }
  };
  long long g = (long long)blockIdx.x * kDecodeWarps + warp;
  if (g < groups) fetch(g);
  for (; g < groups; g += warps_total) {
    const long long row0 = g * 32;
    const int nrows = (int)min((long long)32, total_rows - row0);
    const float* src = head + row0 * kRow;
    // row0*25 floats: 32-row groups start at a multiple of 800 floats = 3200 B -> 16 B aligned
    const int nvec = nrows * kRow / 4;
    float4* dst4 = reinterpret_cast<float4*>(my);
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nvec) dst4[idx] = pre[i];
    }
    for (int i = nvec * 4 + lane; i < nrows * kRow; i += 32) my[i] = __ldg(src + i);
    __syncwarp();
    if (g + warps_total < groups) fetch(g + warps_total);

    const bool active = lane < nrows;
    const long long row = row0 + lane;
    const int b = active ? (int)(row / p.N) : 0;
    const int n = active ? (int)(row % p.N) : 0;
    const float* r = my + lane * kRow;
    float conf[20];
    bool keep_row = active;
    if (p.kind == ODT_DECODE_SSD) {
      // softmax over 21 logits, TF form exp(x-max) * (1/sum); argmax first-max; drop
      // rows whose argmax is background (last index).  ref SSD300.py:159-164
      float m = r[0];
#pragma unroll
      for (int i = 1; i < 21; ++i) m = fmaxf(m, r[i]);
      float e[21];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 21; ++i) {
        e[i] = expf(__fsub_rn(r[i], m));
        s = __fadd_rn(s, e[i]);
      }
      float inv = __fdiv_rn(1.f, s);
      float best = -1.f;
      int arg = 0;
#pragma unroll
      for (int i = 0; i < 21; ++i) {
        float pi = __fmul_rn(e[i], inv);
        if (pi > best) {
          best = pi;
          arg = i;
        }
        if (i < 20) conf[i] = pi;
      }
      keep_row = keep_row && (arg < 20);
    } else if (p.kind == ODT_DECODE_YOLO3) {
      float so = sigmoid_rn(r[24]);  // ref YOLOv3.py:338-339,349
#pragma unroll
      for (int i = 0; i < 20; ++i) conf[i] = __fmul_rn(sigmoid_rn(r[i]), so);
    } else {
      float sc = sigmoid_rn(r[20]);  // ref FCOS.py:197-201
#pragma unroll
      for (int i = 0; i < 20; ++i) conf[i] = __fmul_rn(sigmoid_rn(r[i]), sc);
    }
    // Threshold + append.  Fast path (all rows of the warp in one image, i.e. everywhere but at image
    // boundaries): the 20 per-class ballots are taken first, then lane c reserves the slots of class c
    // with ONE atomic -- 20 independent atomics in flight instead of 20 dependent round trips.
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool one_image = __all_sync(0xffffffffu, !active || b == b0);
    if (one_image) {
      unsigned my_mask = 0u;  // lane c keeps the ballot of class c
      unsigned pass_bits = 0u;
#pragma unroll
      for (int c = 0; c < 20; ++c) {
        const bool pass = c < p.num_fg && keep_row && (conf[c] >= p.score_thr);
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (lane == c) my_mask = m;
        pass_bits |= pass ? (1u << c) : 0u;
      }
      int my_base = 0;
      if (my_mask) my_base = atomicAdd(&cand_count[b0 * p.num_fg + lane], __popc(my_mask));
      const unsigned any = __ballot_sync(0xffffffffu, my_mask != 0u);  // classes with candidates
#pragma unroll
      for (int c = 0; c < 20; ++c) {
        if (!((any >> c) & 1u)) continue;  // warp-uniform
        const unsigned m = __shfl_sync(0xffffffffu, my_mask, c);
        const int base = __shfl_sync(0xffffffffu, my_base, c);
        if ((pass_bits >> c) & 1u) {
          const int slot = base + __popc(m & ((1u << lane) - 1));
          if (slot < p.cap)
            cand_keys[((long long)b0 * p.num_fg + c) * p.cap + slot] =
                ((unsigned long long)__float_as_uint(conf[c]) << 32) | (0xFFFFFFFFu - (unsigned)n);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 20; ++c) {
        if (c >= p.num_fg) break;
        const bool pass = keep_row && (conf[c] >= p.score_thr);
        const unsigned mask = __ballot_sync(0xffffffffu, pass);
        if (mask == 0) continue;
        // rows of this warp straddle two images: aggregate the counter update per image
        unsigned rem = mask;
        int slot = -1;
        while (rem) {
          const int leader = __ffs(rem) - 1;
          const int bl = __shfl_sync(0xffffffffu, b, leader);
          const unsigned grp = __ballot_sync(0xffffffffu, pass && b == bl);
          int base = 0;
          if (lane == leader) base = atomicAdd(&cand_count[bl * p.num_fg + c], __popc(grp));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (pass && b == bl) slot = base + __popc(grp & ((1u << lane) - 1));
          rem &= ~grp;
        }
        if (pass && slot < p.cap) {
          const unsigned long long key =
              ((unsigned long long)__float_as_uint(conf[c]) << 32) | (0xFFFFFFFFu - (unsigned)n);
          cand_keys[((long long)b * p.num_fg + c) * p.cap + slot] = key;
        }
      }
    }
    __syncwarp();
  }
}

// ----------------------------------------------------------------- NMS ----
constexpr int kNmsThreads = 256;
constexpr int kNmsSmemKeys = 1024;  // candidates (keys + boxes, 24 KB) a CTA works on in shared memory: 8 CTAs / SM

struct NmsSmem {
  unsigned long long keys[kNmsSmemKeys];
  float4 box[kNmsSmemKeys];
};

// Exact TF NonMaxSuppressionV3 without a sort: greedy NMS always picks the
// highest-scoring candidate that no kept box suppresses, so each round is
//   (1) block-wide arg-max over the still-alive keys   (key = score bits << 32 | ~row:
//       unique, so "highest score, lowest row on ties" is a plain u64 max),
//   (2) keep it, (3) kill every alive candidate whose IoU with it exceeds the
//   threshold (strict >).
// Rounds = number of kept boxes <= nms_max_boxes (10-20), work per round =
// alive candidates / 256 threads -- no O(n log^2 n) sort, no capacity cliff: lists
// longer than the shared-memory window are processed in place in global memory
// with boxes decoded on the fly.
//
// Long lists (> kNmsPrefilter) first try an exact prefilter: a radix select on the score
// bits finds a threshold T with kNmsSelectMin..kNmsSmemKeys candidates at or above it, and
// the rounds run on that subset S in shared memory.  Greedy NMS visits candidates in
// descending score order, so if nms_max_boxes boxes are kept inside S the result is the
// exact global one (every kept box and every candidate that could have suppressed or
// preceded it has score >= T and therefore is in S).  Only if S runs dry first is the full
// list processed (correct, slower).
constexpr int kNmsPrefilter = kNmsSmemKeys;  // longer lists: prefilter first
constexpr int kNmsSelectMin = 384;
constexpr int kNselOverflowBit = 1 << 30;  // per-(image, class) "list was truncated at cap" flag in nsel_all

__global__ void __launch_bounds__(kNmsThreads)
    nms_per_class_kernel(const float* __restrict__ head, const __grid_constant__ TailP tp, int B,
                         unsigned long long* __restrict__ cand_keys,
                         const int* __restrict__ cand_count, float* __restrict__ dets,
                         int* __restrict__ det_anchor, int* __restrict__ det_count,
                         int* __restrict__ scratch, int* __restrict__ work,
                         int* __restrict__ status, float4* __restrict__ box_pool,
                         long long box_pool_entries, long long det_img_stride) {
  pdl_launch_dependents();
  const odt_tail_params& p = tp.p;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NmsSmem& sm = *reinterpret_cast<NmsSmem*>(smem_raw);
  __shared__ long long s_pool_start;
  __shared__ unsigned long long s_wkey[kNmsThreads / 32];
  __shared__ int s_wpos[kNmsThreads / 32];
  __shared__ int s_last;
  __shared__ int s_hist[256];
  __shared__ unsigned s_ctl[4];  // radix select: [0] done flag, [1] prefix / threshold, [2] count above the bin, [3] subset fill
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int C = p.nms_classes, MB = p.max_boxes;
  // staging: per (b,c): nsel + MB * (6 floats + anchor)
  int* nsel_all = scratch;                                       // [B*C]
  float* st_det = reinterpret_cast<float*>(scratch + B * C);     // [B*C*MB*6]
  int* st_anchor = scratch + B * C + (long long)B * C * MB * 6;  // [B*C*MB]
  const long long bc = (long long)b * C + c;

  int cnt = cand_count[b * p.num_fg + c];
  const bool overflow = cnt > p.cap;
  if (overflow) {
    if (tid == 0) atomicExch(status, ODT_ERR_OVERFLOW);
    cnt = p.cap;
  }
  unsigned long long* gkeys = cand_keys + ((long long)b * p.num_fg + c) * p.cap;
  const float* hb = head + (long long)b * p.N * kRow;
  const int cnt_all = cnt;
  int nsel = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
  // ---- attempt 0 on a long list: exact prefilter (see the kernel comment) ----
  bool subset = false;
  cnt = cnt_all;
  if (attempt == 0 && cnt_all > kNmsPrefilter) {
    unsigned prefix = 0u, prefix_mask = 0u;
    bool done = false;
    for (int shift = 24; shift >= 0 && !done; shift -= 8) {
      s_hist[tid] = 0;  // kNmsThreads == 256 bins
      if (tid == 0 && shift == 24) s_ctl[2] = 0u;
      __syncthreads();
      for (int i = tid; i < cnt_all; i += blockDim.x) {
        const unsigned sc = (unsigned)(gkeys[i] >> 32);
        if ((sc & prefix_mask) == prefix) atomicAdd(&s_hist[(sc >> shift) & 255u], 1);
      }
      __syncthreads();
      if (tid == 0) {
        unsigned cum = s_ctl[2];
        s_ctl[0] = 0u;
        for (int bin = 255; bin >= 0; --bin) {
          const unsigned h = (unsigned)s_hist[bin];
          if (cum + h >= (unsigned)kNmsSelectMin) {
            s_ctl[1] = prefix | ((unsigned)bin << shift);
            if (cum + h <= (unsigned)kNmsSmemKeys) s_ctl[0] = 1u;  // threshold = low end of this bin
            else s_ctl[2] = cum;                                    // refine inside this bin
            break;
          }
          cum += h;
        }
      }
      __syncthreads();
      done = s_ctl[0] != 0u;
      prefix = s_ctl[1];
      prefix_mask |= 0xFFu << shift;
    }
    if (done) {  // (more than kNmsSmemKeys identical scores otherwise: full path)
      const unsigned thr_bits = prefix;
      if (tid == 0) s_ctl[3] = 0u;
      __syncthreads();
      for (int i = tid; i < cnt_all; i += blockDim.x) {
        const unsigned long long k = gkeys[i];
        if ((unsigned)(k >> 32) >= thr_bits) {
          const unsigned slot = atomicAdd(&s_ctl[3], 1u);
          sm.keys[slot] = k;
          const int n = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
          Cell cell = locate(p, n);
          sm.box[slot] = decode_box(p, cell, hb + (long long)n * kRow);
        }
      }
      __syncthreads();
      cnt = (int)s_ctl[3];
      subset = true;
    }
  }
  const bool in_smem = subset || cnt <= kNmsSmemKeys;
  unsigned long long* keys = in_smem ? sm.keys : gkeys;
  // boxes of a list longer than the shared-memory window are decoded ONCE into a
  // slice of the global box pool (bump-allocated per launch); only if the pool is
  // exhausted are they re-decoded on the fly every round (correct, slower)
  const float4* pbox = nullptr;
  if (subset) {
    // already staged
  } else if (in_smem) {
    for (int i = tid; i < cnt; i += blockDim.x) {
      const unsigned long long k = gkeys[i];
      sm.keys[i] = k;
      const int n = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
      Cell cell = locate(p, n);
      sm.box[i] = decode_box(p, cell, hb + (long long)n * kRow);
    }
  } else {
    if (tid == 0) {
      long long start = -1;
      if (box_pool) {
        start = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(work + ((B + 1) & ~1)),
                                     (unsigned long long)cnt);
        if (start + cnt > box_pool_entries) start = -1;
      }
      s_pool_start = start;
    }
    __syncthreads();
    if (s_pool_start >= 0) {
      float4* wb = box_pool + s_pool_start;
      for (int i = tid; i < cnt; i += blockDim.x) {
        const int n = (int)(0xFFFFFFFFu - (unsigned)(gkeys[i] & 0xFFFFFFFFull));
        Cell cell = locate(p, n);
        wb[i] = decode_box(p, cell, hb + (long long)n * kRow);
      }
      pbox = wb;
    }
  }
  __syncthreads();

  nsel = 0;
  while (nsel < MB && cnt > 0) {
    // (1) arg-max over alive keys: per-thread scan, 3 redux.sync per warp, then every
    //     thread folds the 8 warp partials itself (one barrier, no second shuffle tree)
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
    unsigned long long s_best_key = 0ull;
    int s_best_pos = -1;
#pragma unroll
    for (int w = 0; w < kNmsThreads / 32; ++w) {
      const unsigned long long k = s_wkey[w];
      if (k > s_best_key) {
        s_best_key = k;
        s_best_pos = s_wpos[w];
      }
    }
    const unsigned long long hk = s_best_key;
    const int hp = s_best_pos;
    if (hk == 0ull) break;  // nothing alive
    // (2) keep it
    const int hn = (int)(0xFFFFFFFFu - (unsigned)(hk & 0xFFFFFFFFull));
    float4 cur;
    if (in_smem) {
      cur = sm.box[hp];
    } else if (pbox) {
      cur = pbox[hp];
    } else {
      Cell cell = locate(p, hn);
      cur = decode_box(p, cell, hb + (long long)hn * kRow);
    }
    if (tid == 0) {
      float* d = st_det + (bc * MB + nsel) * 6;
      d[0] = __uint_as_float((unsigned)(hk >> 32));
      d[1] = cur.x;
      d[2] = cur.y;
      d[3] = cur.z;
      d[4] = cur.w;
      d[5] = (float)c;
      st_anchor[bc * MB + nsel] = hn;
      keys[hp] = 0ull;
    }
    ++nsel;
    if (nsel >= MB) break;
    // (3) suppress (strict >, TF IoU)
    for (int j = tid; j < cnt; j += blockDim.x) {
      const unsigned long long kj = keys[j];
      if (kj == 0ull || j == hp) continue;
      float4 bj;
      if (in_smem) {
        bj = sm.box[j];
      } else if (pbox) {
        bj = pbox[j];
      } else {
        const int n = (int)(0xFFFFFFFFu - (unsigned)(kj & 0xFFFFFFFFull));
        Cell cell = locate(p, n);
        bj = decode_box(p, cell, hb + (long long)n * kRow);
      }
      if (iou_tf(bj, cur) > p.iou_thr) keys[j] = 0ull;
    }
    __syncthreads();
  }
  __syncthreads();
  // the subset ran dry before nms_max_boxes boxes were kept: redo on the full list
  if (!(subset && nsel < MB && cnt < cnt_all)) break;
  }  // attempt
  if (tid == 0) nsel_all[bc] = nsel | (overflow ? kNselOverflowBit : 0);

  // class-major compaction by the last block of this image
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    int done = atomicAdd(&work[b], 1);
    s_last = (done == C - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ int s_off[33];
  const long long D = (long long)C * MB;
  float* img_dets = dets + (long long)b * det_img_stride;
  if (tid == 0) {
    int acc = 0, ovf = 0;
    for (int i = 0; i < C; ++i) {
      s_off[i] = acc;
      const int v = ((volatile int*)nsel_all)[(long long)b * C + i];
      acc += v & ~kNselOverflowBit;
      ovf |= v & kNselOverflowBit;
    }
    s_off[C] = acc;
    det_count[b] = acc;
    work[b] = 0;
    // packed record (the unit the multi-GPU all-gather and the host read-back ship): the two floats
    // behind an image's D rows carry its detection count and its overflow flag
    if (det_img_stride >= D * 6 + 2) {
      img_dets[D * 6] = (float)acc;
      img_dets[D * 6 + 1] = ovf ? 1.f : 0.f;
    }
  }
  __syncthreads();
  for (int i = tid; i < C * MB; i += blockDim.x) {
    int ci = i / MB, k = i % MB;
    int ns = s_off[ci + 1] - s_off[ci];
    if (k < ns) {
      long long dst = s_off[ci] + k;
      const float* s = st_det + (((long long)b * C + ci) * MB + k) * 6;
#pragma unroll
      for (int q = 0; q < 6; ++q) img_dets[dst * 6 + q] = __ldcg(s + q);
      det_anchor[(long long)b * D + dst] = __ldcg(st_anchor + ((long long)b * C + ci) * MB + k);
    }
  }
}

}  // namespace odt

using namespace odt;

static int check_tail(const odt_tail_params* p) {
  ODT_CHECK_ARG(p != nullptr, "params null");
  ODT_CHECK_ARG(p->kind >= 0 && p->kind <= 2, "kind");
  ODT_CHECK_ARG(p->num_levels >= 1 && p->num_levels <= ODT_MAX_LEVELS, "num_levels");
  ODT_CHECK_ARG(p->num_fg >= 1 && p->num_fg <= 20, "num_fg must be 1..20");
  ODT_CHECK_ARG(p->nms_classes >= 1 && p->nms_classes <= p->num_fg && p->nms_classes <= 32,
                "nms_classes");
  ODT_CHECK_ARG(p->N > 0 && p->cap > 0 && p->max_boxes > 0, "N/cap/max_boxes");
  for (int i = 0; i < p->num_levels; ++i)
    ODT_CHECK_ARG(p->level[i].A >= 1 && p->level[i].A <= ODT_MAX_PRIORS, "level.A");
  return ODT_OK;
}

extern "C" int odt_decode_candidates(const float* head, const odt_tail_params* p, int B,
                                     unsigned long long* cand_keys, int* cand_count,
                                     void* stream) {
  int rc = check_tail(p);
  if (rc) return rc;
  ODT_CHECK_ARG(head && cand_keys && cand_count && B > 0, "null pointer / B");
  ODT_CHECK_ARG(((uintptr_t)head & 15) == 0, "head must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  ODT_CUDA_OK(cudaMemsetAsync(cand_count, 0, sizeof(int) * (size_t)B * p->num_fg, st));
  TailP tp;
  tp.p = *p;
  long long rows = (long long)B * p->N;
  long long groups = (rows + 31) / 32;
  int blocks = (int)((groups + kDecodeWarps - 1) / kDecodeWarps);
  int per_sm = 8;  // measured best of 2/4/8 (80 registers: 3 blocks resident, the rest back-fills the tail)
  if (const char* e = getenv("ODT_DECODE_BLOCKS_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;
  int maxb = kNumSMs * per_sm;
  if (blocks > maxb) blocks = maxb;
  decode_candidates_kernel<<<blocks, kDecodeWarps * 32, 0, st>>>(head, tp, rows, cand_keys,
                                                                 cand_count);
  ODT_LAUNCH_OK();
  return ODT_OK;
}

extern "C" long long odt_nms_scratch_bytes(const odt_tail_params* p, int B) {
  if (!p || B <= 0) return -1;
  long long bc = (long long)B * p->nms_classes;
  return bc * 4 + bc * p->max_boxes * 6 * 4 + bc * p->max_boxes * 4;
}

extern "C" int odt_nms_per_class(const float* head, const odt_tail_params* p, int B,
                                 unsigned long long* cand_keys, const int* cand_count, float* dets,
                                 int* det_anchor, int* det_count, int* sel_scratch, int* work,
                                 int* status, float* box_pool, long long box_pool_entries,
                                 long long dets_img_stride, void* stream) {
  int rc = check_tail(p);
  if (rc) return rc;
  ODT_CHECK_ARG(head && cand_keys && cand_count && dets && det_anchor && det_count &&
                    sel_scratch && work && status && B > 0,
                "null pointer / B");
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    ODT_CUDA_OK(cudaFuncSetAttribute(nms_per_class_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(NmsSmem)));
    attr_set = true;
  }
  TailP tp;
  tp.p = *p;
  dim3 grid(p->nms_classes, B);
  int* pool_ctr = work + ((B + 1) & ~1);  // 64-bit bump pointer behind the per-image counters
  ODT_CHECK_ARG(((uintptr_t)pool_ctr & 7) == 0 || !box_pool, "work must be 8-byte aligned");
  ODT_CHECK_ARG(((uintptr_t)box_pool & 15) == 0, "box_pool must be 16-byte aligned");
  if (box_pool) ODT_CUDA_OK(cudaMemsetAsync(pool_ctr, 0, 8, st));  // pool bump pointer
  ODT_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int), st));        // the status word describes THIS launch
  const long long dense = (long long)p->nms_classes * p->max_boxes * 6;
  if (dets_img_stride == 0) dets_img_stride = dense;
  ODT_CHECK_ARG(dets_img_stride >= dense, "dets_img_stride smaller than nms_classes*max_boxes*6");
  nms_per_class_kernel<<<grid, kNmsThreads, sizeof(NmsSmem), st>>>(
      head, tp, B, cand_keys, cand_count, dets, det_anchor, det_count, sel_scratch, work, status,
      reinterpret_cast<float4*>(box_pool), box_pool ? box_pool_entries : 0, dets_img_stride);
  ODT_LAUNCH_OK();
  return ODT_OK;
}
