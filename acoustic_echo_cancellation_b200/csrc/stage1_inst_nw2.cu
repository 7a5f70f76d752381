// Instantiations: 2 warps per utterance (4 mirrored-bin pairs... 2 pairs per thread), chunk = 4 frames.
#include "stage1_launch.cuh"

namespace aec {

cudaError_t launch_stage1_nw2(int P, int algo, bool echo, int minb, const Stage1Params& prm, cudaStream_t s) {
    // first match for minb == 0 is the default of that (P, algo, echo)
    AEC_TRY_INSTANCE(2, 4, kAlgoNlms, false, 7)
    AEC_TRY_INSTANCE(2, 4, kAlgoNlms, false, 6)
    AEC_TRY_INSTANCE(2, 4, kAlgoNlms, false, 5)
    AEC_TRY_INSTANCE(2, 4, kAlgoNlms, false, 4)
    AEC_TRY_INSTANCE(2, 4, kAlgoNlms, true, 6)
    AEC_TRY_INSTANCE(2, 4, kAlgoKalman, false, 6)
    AEC_TRY_INSTANCE(2, 4, kAlgoKalman, true, 5)
    AEC_TRY_INSTANCE(2, 1, kAlgoNlms, false, 7)
    AEC_TRY_INSTANCE(2, 1, kAlgoNlms, true, 7)
    AEC_TRY_INSTANCE(2, 1, kAlgoKalman, false, 7)
    AEC_TRY_INSTANCE(2, 1, kAlgoKalman, true, 7)
    AEC_TRY_INSTANCE(2, 2, kAlgoNlms, false, 7)
    AEC_TRY_INSTANCE(2, 2, kAlgoNlms, true, 7)
    AEC_TRY_INSTANCE(2, 2, kAlgoKalman, false, 7)
    AEC_TRY_INSTANCE(2, 2, kAlgoKalman, true, 7)
    return cudaErrorInvalidValue;
}

}  // namespace aec
