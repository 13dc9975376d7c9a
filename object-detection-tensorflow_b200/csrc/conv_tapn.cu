// "Taps as N": the 3x3 / stride-1 halo convolution for NARROW outputs (Cout_pad <= 64) with the three
// horizontal filter taps folded into the N dimension of the MMA.
//
// Why: a tcgen05.mma M128 x N x K16 fetches its operands from shared memory at ~64 B/clk, i.e.
// 2*(32 + N/4) clk, while the tensor pipe itself needs N/2 clk (conv_tc.cu).  At N = 64 the
// fetch (96 clk) is 3x the math (32 clk): conv1_2 of VGG and every thin RetinaNet layer sit on that floor
// (36 MMAs x 96 clk per 128-pixel tile).  Each activation row that is fetched should therefore feed more
// output columns.  Here one MMA multiplies the activation tile of filter row r by the weights of ALL THREE
// horizontal taps:
//
//   P[q, s*BN + n] = sum_{r, c} X[q + r*PW, c] * W[n, r, s, c]          (N = 3*BN, K = 3*Cin)
//   out[p, n]      = P[p-1, 0*BN + n] + P[p, 1*BN + n] + P[p+1, 2*BN + n]
//
// so a tile costs 3*cchunks*4 MMAs of N = 192 (160 clk each) instead of 9*cchunks*4 of N = 64 (96 clk):
// 1920 instead of 3456 clk at Cin = Cout = 64.  The shifted sum runs in the epilogue: a tile is 4 image rows
// x 32 columns with the COLUMN index in the lane (one warp = one row), P[p-1] / P[p+1] come from the
// neighbouring lanes (shfl), lanes 0 and 31 only serve as neighbours, tiles overlap by two columns (30 outputs
// per row and tile).  The fused 2x2/2 max-pool pairs columns inside the warp and rows through a small
// shared-memory exchange between the warps of an epilogue group.
//
// Roles / pipeline as in conv_tc.cu: w0 TMA producer (one [64 ch][32 x][4 rows] box per filter row and
// channel chunk), w1 MMA issuer, w2..9 two epilogue groups draining alternate tiles from two TMEM stages;
// the 9*cchunks weight tiles stay resident in shared memory.
//
// OPT-IN (ODT_TC_TAPN=1) until it has been A/B-timed on a B200: written at the end of round 1; its eight parity
// cases (tests/test_gpu_conv.py::test_conv_tapn_matches_reference) passed on a B200 with the last GPU seconds of
// the round, no timing yet.  ref call sites: tf.nn.conv2d SSD300.py:519 (conv1_2), tf.layers.conv2d
// RetinaNet.py:599-609, YOLOv3.py:495, FCOS.py:469-479.
#include <string.h>

#include "epilogue.cuh"
#include "tc_ptx.cuh"

namespace odt {

constexpr int TN_THREADS = 320;      // TMA, MMA, 2 x 4 epilogue warps
constexpr int TN_XV = 30;            // output columns per tile row (32 lanes - 2 neighbour lanes)
constexpr int TN_ROWS = 4;           // image rows per tile (one per TMEM lane quarter)
constexpr int TN_A_BYTES = 128 * 128;
constexpr int TN_MAX_STAGES = 8;
constexpr int TN_ACC_STRIDE = 256;   // TMEM columns between the two accumulator stages (3*BN <= 192 used)
constexpr int TN_PAR_FLOATS = 6 * 64;
constexpr int TN_XCH_PITCH = 80;     // bytes per lane in the pooling exchange (64 payload, padded against bank conflicts)
constexpr int TN_XCH_BYTES = 2 * 2 * 2 * 32 * TN_XCH_PITCH;  // [group][buffer][row pair][lane]
constexpr int TN_SMEM_LIMIT = 232448;

struct TnGeom {
  int B, H, W;
  int cchunks, klast;
  int BN;  // Cout_pad: 32 or 64
  int xblocks, yblocks, num_tiles;
  int stages;
  int out_halo;
};

template <int POOL>
__global__ void __launch_bounds__(TN_THREADS, 1)
    conv_tapn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ TnGeom g, const __grid_constant__ Epi e) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int stages = g.stages;
  const uint32_t a_base = base;
  const uint32_t b_base = base + (uint32_t)stages * TN_A_BYTES;
  const uint32_t wtile = 3u * (uint32_t)g.BN * 128u;  // weights of one (filter row, chunk): rows s*BN + n
  const uint32_t wbytes = 3u * (uint32_t)g.cchunks * wtile;
  const uint32_t bar_base = b_base + wbytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TN_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TN_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TN_MAX_STAGES + 2 + a); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * TN_MAX_STAGES + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * TN_MAX_STAGES + 5);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  float* par = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - raw));  // [scale|shift|scale2|shift2|scale3|shift3][64]
  uint8_t* xch = smem_raw + (bar_base + 256u + TN_PAR_FLOATS * 4u - raw);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);  // the four warps of the group that owns the stage
    }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    // per-channel epilogue parameters (constant data): one copy per CTA
    for (int i = threadIdx.x - 64; i < TN_PAR_FLOATS; i += 256) {
      const int which = i >> 6, c = i & 63;
      const bool ok = c < e.Cout;
      const float* src = which == 0 ? e.scale : which == 1 ? e.shift : which == 2 ? e.scale2 : which == 3 ? e.shift2
                         : which == 4 ? e.scale3 : e.shift3;
      par[i] = (ok && src) ? __ldg(src + c) : ((which & 1) ? 0.f : 1.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // the filter bank is constant data: fetch it before waiting on the previous kernel
    if (elect_one()) {
      mbar_expect_tx(wfull_bar, wbytes);
      for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < g.cchunks; ++cc)
          for (int s = 0; s < 3; ++s)
            tma_load_2d(b_base + (uint32_t)(r * g.cchunks + cc) * wtile + (uint32_t)s * (uint32_t)g.BN * 128u, &tmB,
                        wfull_bar, ((r * 3 + s) * g.cchunks + cc) * 64, 0);
    }
    __syncwarp();
  }
  pdl_launch_dependents();
  pdl_wait();

  const int per_img = g.yblocks * g.xblocks;

  if (warp == 0) {
    // ===================== TMA producer ======================================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
      const int img = tile / per_img;
      const int rem = tile - img * per_img;
      const int yq = rem / g.xblocks, xq = rem - yq * g.xblocks;
      for (int r = 0; r < 3; ++r) {
        for (int cc = 0; cc < g.cchunks; ++cc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(full_bar(stage), TN_A_BYTES);
            // padded coordinates: columns TN_XV*xq .. +31 (real x = TN_XV*xq - 1 ..), rows 4*yq + r .. +3
            // (real y = 4*yq + r - 1 ..); everything past the padded extent is TMA zero fill
            tma_load_4d(a_base + (uint32_t)stage * TN_A_BYTES, &tmA, full_bar(stage), cc * 64, TN_XV * xq,
                        TN_ROWS * yq + r, img);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =========================================
    const uint32_t idesc = make_idesc_f16(128, 3 * g.BN);
    int stage = 0;
    uint32_t phase = 0;
    int local_tile = 0;
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++local_tile) {
      const int acc = local_tile & 1;
      const uint32_t use = (uint32_t)(local_tile >> 1);
      mbar_wait(tempty_bar(acc), (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TN_ACC_STRIDE);
      for (int r = 0; r < 3; ++r) {
        for (int cc = 0; cc < g.cchunks; ++cc) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = make_desc_sw128(a_base + (uint32_t)stage * TN_A_BYTES);
          const uint64_t bdesc = make_desc_sw128(b_base + (uint32_t)(r * g.cchunks + cc) * wtile);
          const int ksteps = (cc == g.cchunks - 1) ? g.klast : 4;  // all-zero K steps of a thin layer are not issued
          const bool last = r == 2 && cc == g.cchunks - 1;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                         (uint32_t)((r | cc | k) != 0), (uint32_t)(k < ksteps));
            tc_commit(empty_bar(stage));
            if (last) tc_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================================
    const int quarter = warp & 3;          // TMEM lane quarter = image row inside the tile
    const int group = (warp - 2) >> 2;     // owns accumulator stage `group`
    const int chunks = g.BN >> 5;          // 32-channel chunks per tile
    int xbuf = 0;
    int local_tile = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++local_tile) {
      if ((local_tile & 1) != group) continue;
      const uint32_t use = (uint32_t)(local_tile >> 1);
      const int img = tile / per_img;
      const int rem = tile - img * per_img;
      const int yq = rem / g.xblocks, xq = rem - yq * g.xblocks;
      const int y = TN_ROWS * yq + quarter;
      const int x = TN_XV * xq - 1 + lane;
      const bool row_ok = lane >= 1 && lane <= TN_XV && x < g.W && y < g.H;
      long long o0_row = (long long)img * e.out0_img_stride;
      long long o1_row = 0, o2_row = 0;
      if (POOL) {
        // pooled pixel of this lane's window (used by the odd lanes of the even rows)
        const int oh = g.out_halo;
        o0_row += (long long)(((y >> 1) + oh) * ((g.W >> 1) + 2 * oh) + (x >> 1) + oh) * e.out0_pix_stride;
      } else {
        const int pix = row_ok ? y * g.W + x : 0;
        o0_row += g.out_halo ? (long long)((y + 1) * (g.W + 2) + x + 1) * e.out0_pix_stride
                             : (long long)pix * e.out0_pix_stride;
        o1_row = aux_row(e.out1_img_stride, e.out1_pix_stride, e.out1_halo, g.W, img, pix);
        o2_row = aux_row(e.out2_img_stride, e.out2_pix_stride, e.out2_halo, g.W, img, pix);
      }
      mbar_wait(tfull_bar(group), use & 1u);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (uint32_t)(group * TN_ACC_STRIDE) + ((uint32_t)(quarter * 32) << 16);
      for (int j = 0; j < chunks; ++j) {
        const int nb = j * 32;
        uint32_t r0[32], r1[32], r2[32];
        tc_ld32(taddr0 + (uint32_t)(0 * g.BN + nb), r0);
        tc_ld32(taddr0 + (uint32_t)(1 * g.BN + nb), r1);
        tc_ld32(taddr0 + (uint32_t)(2 * g.BN + nb), r2);
        tc_wait_ld();
        if (j == chunks - 1) {
          // every TMEM read of this tile has completed: hand the stage back before the math
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(group));
        }
        // out[p] = P_left[p-1] + P_centre[p] + P_right[p+1]  (lanes 0 / 31 produce unused values)
        const float4* ps = reinterpret_cast<const float4*>(par + nb);
        const float4* ph4 = reinterpret_cast<const float4*>(par + 64 + nb);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
          v[i] = __uint_as_float(r1[i]) + __uint_as_float(__shfl_up_sync(0xffffffffu, r0[i], 1)) +
                 __uint_as_float(__shfl_down_sync(0xffffffffu, r2[i], 1));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 sc = ps[i], sh = ph4[i];
          v[4 * i + 0] = apply_act(fmaf(v[4 * i + 0], sc.x, sh.x), e.act);
          v[4 * i + 1] = apply_act(fmaf(v[4 * i + 1], sc.y, sh.y), e.act);
          v[4 * i + 2] = apply_act(fmaf(v[4 * i + 2], sc.z, sh.z), e.act);
          v[4 * i + 3] = apply_act(fmaf(v[4 * i + 3], sc.w, sh.w), e.act);
        }
        if (POOL) {
          // 2x2/2 max-pool of the rounded values: columns (odd lane, next lane), rows (even warp, next warp)
          uint4 packed[4];
          __half2* ph = reinterpret_cast<__half2*>(packed);
          uint32_t* pw = reinterpret_cast<uint32_t*>(packed);
#pragma unroll
          for (int i = 0; i < 16; ++i) ph[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            uint32_t o = __shfl_down_sync(0xffffffffu, pw[i], 1);
            __half2 a = __hmax2(*reinterpret_cast<__half2*>(&pw[i]), *reinterpret_cast<__half2*>(&o));
            pw[i] = *reinterpret_cast<uint32_t*>(&a);
          }
          uint8_t* slot = xch + (size_t)(((group * 2 + xbuf) * 2 + (quarter >> 1)) * 32 + lane) * TN_XCH_PITCH;
          if (quarter & 1) {
            uint4* sp = reinterpret_cast<uint4*>(slot);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) sp[qd] = packed[qd];
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
          if (!(quarter & 1)) {
            const uint4* sp = reinterpret_cast<const uint4*>(slot);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              uint4 t = sp[qd];
              const __half2* th = reinterpret_cast<const __half2*>(&t);
              __half2* mine = reinterpret_cast<__half2*>(&packed[qd]);
#pragma unroll
              for (int i = 0; i < 4; ++i) mine[i] = __hmax2(mine[i], th[i]);
            }
            if (row_ok && (lane & 1)) {  // even x, even y: the window's top-left pixel
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out0) + o0_row + nb);
#pragma unroll
              for (int qd = 0; qd < 4; ++qd) op[qd] = packed[qd];
            }
          }
          xbuf ^= 1;  // the next chunk's writers may run ahead of this chunk's readers by one barrier only
        } else if (row_ok) {
          const long long o0 = o0_row + nb;
          if (e.residual) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(e.residual) + o0);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              uint4 t = __ldg(rp + qd);
              const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float2 f = __half22float2(h[i]);
                v[qd * 8 + 2 * i] += f.x;
                v[qd * 8 + 2 * i + 1] += f.y;
              }
            }
          }
          uint4 packed[4];
          __half2* ph = reinterpret_cast<__half2*>(packed);
#pragma unroll
          for (int i = 0; i < 16; ++i) ph[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
          if (e.out0) {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out0) + o0);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) op[qd] = packed[qd];
          }
          if (e.out1) {
            const float4* ps2 = reinterpret_cast<const float4*>(par + 128 + nb);
            const float4* ph2 = reinterpret_cast<const float4*>(par + 192 + nb);
            uint4 packed1[4];
            __half2* p1 = reinterpret_cast<__half2*>(packed1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 sc = ps2[i], sh = ph2[i];
              const float2 fa = __half22float2(ph[2 * i]);  // the consumer sees the rounded value
              const float2 fb = __half22float2(ph[2 * i + 1]);
              p1[2 * i] = __floats2half2_rn(apply_act(fmaf(fa.x, sc.x, sh.x), e.act2),
                                            apply_act(fmaf(fa.y, sc.y, sh.y), e.act2));
              p1[2 * i + 1] = __floats2half2_rn(apply_act(fmaf(fb.x, sc.z, sh.z), e.act2),
                                                apply_act(fmaf(fb.y, sc.w, sh.w), e.act2));
            }
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out1) + o1_row + nb);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) op[qd] = packed1[qd];
          }
          if (e.out2) {
            const float4* ps3 = reinterpret_cast<const float4*>(par + 256 + nb);
            const float4* ph3 = reinterpret_cast<const float4*>(par + 320 + nb);
            uint4 packed2[4];
            __half2* p2 = reinterpret_cast<__half2*>(packed2);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 sc = ps3[i], sh = ph3[i];
              const float2 fa = __half22float2(ph[2 * i]);
              const float2 fb = __half22float2(ph[2 * i + 1]);
              p2[2 * i] = __floats2half2_rn(apply_act(fmaf(fa.x, sc.x, sh.x), e.act3),
                                            apply_act(fmaf(fa.y, sc.y, sh.y), e.act3));
              p2[2 * i + 1] = __floats2half2_rn(apply_act(fmaf(fb.x, sc.z, sh.z), e.act3),
                                                apply_act(fmaf(fb.y, sc.w, sh.w), e.act3));
            }
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out2) + o2_row + nb);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) op[qd] = packed2[qd];
          }
        }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

static int g_tapn_launches = 0;  // debug: lets a test assert that the layer really took this path

// ODT_ERR_UNSUPPORTED when the layer does not qualify (the caller then takes the regular paths).
int conv_tapn_try(const void* in, const void* weights, const odt_conv_params* p, void* stream) {
  const bool shape_ok = p->in_halo == 1 && p->R == 3 && p->S == 3 && p->stride == 1 && p->dil == 1 && p->pad_t == 1 &&
                        p->pad_l == 1 && p->OH == p->H && p->OW == p->W && (p->Cout_pad == 32 || p->Cout_pad == 64) &&
                        p->in_ld % 64 == 0 && p->in_ld / 64 <= 2 && p->w_ld == p->in_ld;
  if (!shape_ok) return ODT_ERR_UNSUPPORTED;
  const bool pool = p->out0_pool == 2;
  auto aligned = [](const void* ptr, long long img, int pix, int need) {
    return ((uintptr_t)ptr & 15) == 0 && img % 8 == 0 && pix % 8 == 0 && pix >= need;
  };
  bool out_ok = p->out0_group == 0 && (!p->out0 || p->out0_dtype == ODT_F16) &&
                (!p->out0 || aligned(p->out0, p->out0_img_stride, p->out0_pix_stride, p->Cout_pad)) &&
                (!p->out1 || aligned(p->out1, p->out1_img_stride, p->out1_pix_stride, p->Cout_pad)) &&
                (!p->out2 || aligned(p->out2, p->out2_img_stride, p->out2_pix_stride, p->Cout_pad)) &&
                (!p->residual || (((uintptr_t)p->residual & 15) == 0 && p->out0_img_stride % 8 == 0 &&
                                  p->out0_pix_stride % 8 == 0 && p->out0_pix_stride >= p->Cout_pad));
  if (pool)
    out_ok = out_ok && p->out0 && !p->residual && !p->out1 && !p->out2 && p->OH % 2 == 0 && p->OW % 2 == 0;
  else
    out_ok = out_ok && p->out0_pool == 0;
  if (!out_ok) return ODT_ERR_UNSUPPORTED;

  TnGeom g;
  memset(&g, 0, sizeof(g));
  g.B = p->B;
  g.H = p->H;
  g.W = p->W;
  g.cchunks = p->in_ld / 64;
  {
    const int rem = p->Cin - (g.cchunks - 1) * 64;
    g.klast = rem >= 64 ? 4 : (rem <= 0 ? 1 : (rem + 15) / 16);
  }
  g.BN = p->Cout_pad;
  g.xblocks = (p->W + TN_XV - 1) / TN_XV;
  g.yblocks = (p->H + TN_ROWS - 1) / TN_ROWS;
  const long long tiles = (long long)p->B * g.xblocks * g.yblocks;
  if (tiles >= (1ll << 31)) return ODT_ERR_UNSUPPORTED;
  g.num_tiles = (int)tiles;
  g.out_halo = p->out0_halo;
  if (tapn_mode() == 1) {
    // same operand-fetch model as pick_tiling (conv_tc.cu): rounds of the slowest SM x MMAs per tile x clk per MMA
    // (2*(32 + N/4), A and half of B with the CTA pairs the regular flat modes use), + a fixed cost per tile.
    // Small maps lose here: a 13-wide row still occupies a 32-lane tile row.
    const long long mm = 4ll * g.cchunks;
    const long long c_tapn = (long long)ceil_div(tiles, kNumSMs) * (3 * mm * (64 + 3 * g.BN / 2) + 1200);
    long long t_reg;  // tiles of the regular path
    bool pairs;
    if (pool) {
      const long long t2 = (long long)((p->H + 1) / 2) * ((p->W + 63) / 64), t4 = (long long)((p->H + 3) / 4) * ((p->W + 31) / 32);
      t_reg = (long long)p->B * (t4 < t2 ? t4 : t2);
      pairs = true;
    } else {
      t_reg = ((long long)p->B * (p->H + 2) * (p->W + 2) + 127) / 128;
      pairs = g.cchunks >= 2;
    }
    const long long c_reg = (long long)ceil_div(t_reg, kNumSMs) * (9 * mm * (64 + (pairs ? g.BN / 4 : g.BN / 2)) + 1200);
    if (c_tapn >= c_reg) return ODT_ERR_UNSUPPORTED;
  }
  const int wbytes = 9 * g.cchunks * g.BN * 128;
  const int fixed = wbytes + 256 + TN_PAR_FLOATS * 4 + (pool ? TN_XCH_BYTES : 0) + 1024;
  int stages = (TN_SMEM_LIMIT - fixed) / TN_A_BYTES;
  if (stages > TN_MAX_STAGES) stages = TN_MAX_STAGES;
  if (stages < 3) return ODT_ERR_UNSUPPORTED;
  g.stages = stages;

  CUtensorMap tmA, tmB;
  {
    // [ld][W+2][H+2][B] view of the halo tensor; one {64, 32, 4, 1} box = the tile's rows of one filter row,
    // column index fastest (shared-memory row = 32*row + column = TMEM lane)
    const cuuint64_t PH = (cuuint64_t)p->H + 2, PW = (cuuint64_t)p->W + 2;
    cuuint64_t dims[4] = {(cuuint64_t)p->in_ld, PW, PH, (cuuint64_t)p->B};
    cuuint64_t strides[3] = {(cuuint64_t)p->in_ld * 2, PW * p->in_ld * 2, PH * PW * p->in_ld * 2};
    cuuint32_t box[4] = {64, 32, TN_ROWS, 1};
    int rc = tc_encode_tiled(&tmA, in, 4, dims, strides, box, false);
    if (rc) return rc;
  }
  {
    const cuuint64_t ktot = (cuuint64_t)9 * p->w_ld;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)p->Cout_pad};
    cuuint64_t strides[1] = {ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)g.BN};
    int rc = tc_encode_tiled(&tmB, weights, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    ODT_CUDA_OK(cudaFuncSetAttribute(conv_tapn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_LIMIT));
    ODT_CUDA_OK(cudaFuncSetAttribute(conv_tapn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_LIMIT));
    attr_set = true;
  }
  Epi e = make_epi(*p);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TN_THREADS);
  cfg.gridDim = dim3(g.num_tiles < kNumSMs ? g.num_tiles : kNumSMs);
  cfg.dynamicSmemBytes = stages * TN_A_BYTES + fixed;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (pool)
    ODT_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tapn_kernel<1>, tmA, tmB, g, e));
  else
    ODT_CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tapn_kernel<0>, tmA, tmB, g, e));
  ODT_LAUNCH_OK();
  ++g_tapn_launches;
  return ODT_OK;
}

}  // namespace odt

// Debug only: number of convolutions launched through conv_tapn_kernel by this process.
extern "C" int odt_debug_tapn_launches(void) { return odt::g_tapn_launches; }
