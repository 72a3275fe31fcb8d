// Low-level test hooks of the C-ABI (mudg_test_*): single-op entry points used by tests/ to check each
// kernel against the oracle.  Plain pointers + sizes only.
#include "common.h"
#include "gemm.h"
#include "mudg.h"

namespace mudg {
const char* last_error_cstr();
}
using namespace mudg;

#define MUDG_API_BEGIN try {
#define MUDG_API_END                                   \
  return 0;                                            \
  }                                                    \
  catch (const std::exception& e) {                    \
    mudg::set_last_error(e.what());                    \
    return -1;                                         \
  }

extern "C" {

MUDG_EXPORT const char* mudg_last_error(void) { return last_error_cstr(); }

MUDG_EXPORT int mudg_test_tapgemm(const void* A, int B, int T, int H, int W, int Cin, int mode, const void* Wt, int N,
                                  void* D, const void* R, const float* bias, const float* bias2, int bias2_div, int nb2,
                                  float alpha, int geglu, int backend, void* stream) {
  MUDG_API_BEGIN
  TapGemm g;
  g.A = static_cast<const __half*>(A);
  g.B = B; g.T = T; g.H = H; g.W = W; g.Cin = Cin;
  if (mode == 0) { g.ntaps = 1; g.taps[0][0] = g.taps[0][1] = g.taps[0][2] = 0; }
  else if (mode == 1) { g.ntaps = 9; set_taps_3x3(g.taps); }
  else { g.ntaps = 3; set_taps_t3(g.taps); }
  g.Wt = static_cast<const __half*>(Wt);
  g.N = N;
  g.D = static_cast<__half*>(D);
  g.R = static_cast<const __half*>(R);
  g.bias = bias; g.bias2 = bias2; g.bias2_div = bias2_div; g.nb2 = nb2;
  g.alpha = alpha; g.geglu = geglu != 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (backend == 0) tapgemm_tc(g, st);
  else tapgemm_simt(g, st);
  MUDG_API_END
}

}  // extern "C"
