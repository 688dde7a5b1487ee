// One resampler instance of the fused preprocess kernel (see fused_kernel.cuh).
#include "fused_kernel.cuh"
namespace mpcg {
template int fz_launch<1, 1, 1, 1>(const FzParams&, size_t, long long, cudaStream_t);
}
