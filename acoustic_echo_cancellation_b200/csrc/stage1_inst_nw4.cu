// Instantiations: 4 warps per utterance (1 mirrored-bin pair per thread), chunk = 8 frames.
#include "stage1_launch.cuh"

namespace aec {

cudaError_t launch_stage1_nw4(int P, int algo, bool echo, int minb, const Stage1Params& prm, cudaStream_t s) {
    AEC_TRY_INSTANCE(4, 8, kAlgoNlms, false, 3)
    AEC_TRY_INSTANCE(4, 8, kAlgoNlms, true, 3)
    AEC_TRY_INSTANCE(4, 8, kAlgoKalman, false, 2)
    AEC_TRY_INSTANCE(4, 8, kAlgoKalman, true, 2)
    AEC_TRY_INSTANCE(4, 16, kAlgoNlms, false, 2)
    AEC_TRY_INSTANCE(4, 16, kAlgoNlms, true, 2)
    AEC_TRY_INSTANCE(4, 16, kAlgoKalman, false, 2)
    AEC_TRY_INSTANCE(4, 16, kAlgoKalman, true, 2)
    AEC_TRY_INSTANCE(4, 4, kAlgoNlms, false, 3)
    AEC_TRY_INSTANCE(4, 4, kAlgoNlms, false, 4)
    AEC_TRY_INSTANCE(4, 4, kAlgoNlms, false, 2)
    AEC_TRY_INSTANCE(4, 4, kAlgoKalman, false, 3)
    return cudaErrorInvalidValue;
}

}  // namespace aec
