// Hardware probe (test-only entry point, not part of the public ABI): does a
// K-major SWIZZLE_128B UMMA operand tolerate a start address that is shifted by
// whole 128-byte rows inside the 1024-byte swizzle pattern, and which value of the
// descriptor's base_offset field does it need?  This decides whether one loaded
// [130 x 64] activation slab can feed the three horizontal taps of a 3x3 filter.
#include "tc_ptx.cuh"

namespace odt {

__global__ void __launch_bounds__(128)
    umma_rowoffset_probe_kernel(const __half* __restrict__ X /*[136][64]*/,
                                const __half* __restrict__ Wt /*[64][64]*/,
                                float* __restrict__ out /*[3][128][64]*/, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g = smem_raw + (base - raw);
  const uint32_t a_base = base;                  // 136 rows x 128 B = 17408 B
  const uint32_t b_base = base + 18432;          // 64 rows x 128 B
  const uint32_t bar = b_base + 8192;
  const uint32_t slot = bar + 16;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - raw));
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int i = t; i < 136 * 8; i += 128) {
    const int row = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(g + row * 128 + ((c ^ (row & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(X + row * 64 + c * 8);
  }
  for (int i = t; i < 64 * 8; i += 128) {
    const int row = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(g + 18432 + row * 128 + ((c ^ (row & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(Wt + row * 64 + c * 8);
  }
  if (t == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  const uint32_t idesc = make_idesc_f16(128, 64);
  uint32_t phase = 0;
  for (int s = 0; s < 3; ++s) {
    if (warp == 0) {
      if (elect_one()) {
        uint64_t adesc = make_desc_sw128(a_base + (uint32_t)s * 128u);
        if (mode == 1) adesc |= (uint64_t)(((a_base + s * 128u) >> 7) & 7u) << 49;
        const uint64_t bdesc = make_desc_sw128(b_base);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
        tc_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    uint32_t r[32];
    for (int j = 0; j < 2; ++j) {
      tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(j * 32), r);
      tc_wait_ld();
      for (int i = 0; i < 32; ++i)
        out[((size_t)s * 128 + warp * 32 + lane) * 64 + j * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

}  // namespace odt

extern "C" int odt_test_umma_rowoffset(const void* X, const void* Wt, float* out, int mode, void* stream) {
  using namespace odt;
  static bool attr = false;
  if (!attr) {
    ODT_CUDA_OK(cudaFuncSetAttribute(umma_rowoffset_probe_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    attr = true;
  }
  umma_rowoffset_probe_kernel<<<1, 128, 32768, (cudaStream_t)stream>>>((const __half*)X, (const __half*)Wt,
                                                                    out, mode);
  ODT_LAUNCH_OK();
  return ODT_OK;
}
