// fp32 streaming Gaussian launcher (stub until the kernel lands)
#include "mp_internal.h"
#include "mp_ops_internal.h"

namespace mp {
bool gauss_stream_supported(int, int, int) { return false; }
MPStatus launch_gauss_stream(int, cudaStream_t, const Img &, const float *, float *, const mpk::GaussParams<float> &)
{
    return MP_ERROR_INVALID_ARGUMENT;
}
}  // namespace mp
