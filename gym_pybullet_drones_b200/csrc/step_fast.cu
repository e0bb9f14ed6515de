// step_fast.cu -- (stub, replaced below)
#include "qs_common.cuh"
namespace qsi {
bool step_fast_eligible(const StepArgs&) { return false; }
cudaError_t launch_step_fast(const StepArgs&, cudaStream_t) { return cudaErrorNotSupported; }
}
