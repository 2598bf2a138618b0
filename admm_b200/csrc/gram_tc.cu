// gram_tc.cu -- placeholder until the tcgen05 kernel lands (returns "shape not supported").
#include "common.cuh"
#include "kernels.h"
namespace b200 {
bool gram_tn_tensor(cudaStream_t, const float*, i64, i64, float*) { return false; }
}
