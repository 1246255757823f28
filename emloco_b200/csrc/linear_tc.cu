// placeholder until the tcgen05 path lands (next commit): refuse loudly rather than fall back silently
#include "sim.h"
cudaError_t eml_linear_tc(const float*, long long, const float*, const float*, float*, long long, long long, int, int,
                          const float*, const float*, float, int, cudaStream_t) {
    return cudaErrorNotSupported;
}
