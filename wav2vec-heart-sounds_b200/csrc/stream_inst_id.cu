// One resampler instance of the row-streaming fused kernel (see stream_kernel.cuh).
#include "stream_kernel.cuh"
#include "stream_instances.h"
namespace mpcg {
#define MPCG_SK_INST_(...) template int sk_launch<__VA_ARGS__>(const SkParams&, size_t, int, cudaStream_t);
MPCG_SK_INST_(1, 1, 1, 1)
}
