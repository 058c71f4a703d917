// kernels_f32.cu -- the product path: every kernel instantiated for Real = float.
#include "launch_impl.cuh"
namespace rptb {
RPTB_DEFINE_LAUNCHERS(f32, float)
}
