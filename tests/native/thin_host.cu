// TEST-ONLY: runs the per-thread body of csrc/conv_thin.cu (thin_fill / thin_pixel, __host__ __device__) on the
// CPU, thread by thread, so that tests/test_thin_host.py can check the kernel's index arithmetic, rounding chain
// and epilogue against the oracle without a GPU.  Built by that test into tests/native/_build/; never linked into
// libodt_b200.so (the product has no host path).
#include <vector>

#include "../../object-detection-tensorflow_b200/csrc/conv_thin.cu"

namespace odt {
int thin_mode() { return 2; }
void set_error(const char*, ...) {}

template <int KS, int CI8, int CO8>
static void run_thin_host(const void* in, const void* weights, const ThinGeom& g, const Epi& e) {
  std::vector<float> ws((size_t)KS * KS * CI8 * 8 * CO8 * 8 + 4), par((size_t)6 * CO8 * 8 + 4);
  // 16-byte aligned views (thin_pixel reads the bank as float4)
  float* wsp = reinterpret_cast<float*>(((uintptr_t)ws.data() + 15) & ~(uintptr_t)15);
  float* parp = reinterpret_cast<float*>(((uintptr_t)par.data() + 15) & ~(uintptr_t)15);
  for (int tid = 0; tid < THIN_THREADS; ++tid)
    thin_fill<KS, CI8, CO8>(tid, wsp, parp, reinterpret_cast<const __half*>(weights), g, e);
  for (long long m = 0; m < g.M; ++m) thin_pixel<KS, CI8, CO8>(m, wsp, parp, reinterpret_cast<const __half*>(in), g, e);
}
}  // namespace odt

// host pointers everywhere; returns ODT_OK, or ODT_ERR_UNSUPPORTED exactly where conv_thin_try would decline
extern "C" int odt_test_thin_host(const void* in, const void* weights, const odt_conv_params* p, int force) {
  using namespace odt;
  ThinGeom g;
  int ks, ci8r, co8r;
  const int rc = thin_plan(in, p, force != 0, &g, &ks, &ci8r, &co8r);
  if (rc) return rc;
  const Epi e = make_epi(*p);
#define ODT_THIN_HOST_CASE(KS_, CI_, CO_)                \
  if (ks == KS_ && ci8r == CI_ && co8r == CO_) {         \
    run_thin_host<KS_, CI_, CO_>(in, weights, g, e);     \
    return ODT_OK;                                       \
  }
  ODT_THIN_DISPATCH(ODT_THIN_HOST_CASE)
#undef ODT_THIN_HOST_CASE
  return ODT_ERR_UNSUPPORTED;
}
