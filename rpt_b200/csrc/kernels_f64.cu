// kernels_f64.cu -- the parity gate: the same kernels for Real = double, keeping the
// reference's f64 formulas verbatim.  Compiled with -fmad=false so that products and
// sums round like the reference's (rustc does not contract a*b+c).
#include "launch_impl.cuh"
namespace rptb {
RPTB_DEFINE_LAUNCHERS(f64, double)
}
