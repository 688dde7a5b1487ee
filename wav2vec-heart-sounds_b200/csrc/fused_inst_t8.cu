// One resampler instance of the fused preprocess kernel (see fused_kernel.cuh).
#include "fused_kernel.cuh"
#include "fused_instances.h"
namespace mpcg {
#define MPCG_FZ_INST_(...) template int fz_launch<__VA_ARGS__>(const FzParams&, size_t, long long, cudaStream_t);
MPCG_FZ_INST_(FZ_I8)
}
