// Instantiations: 8 warps per utterance, ONE bin per thread, chunk = 16 frames -- long filters
// (the per-bin state W_p, X[t-p], C_p of a 16-partition Kalman filter is 81 registers).
#include "stage1_launch.cuh"

namespace aec {

cudaError_t launch_stage1_nw8(int P, int algo, bool echo, int regs, const Stage1Params& prm, cudaStream_t s) {
    AEC_TRY_INSTANCE(8, 16, kAlgoKalman, false, 128)
    AEC_TRY_INSTANCE(8, 16, kAlgoKalman, true, 128)
    AEC_TRY_INSTANCE(8, 16, kAlgoNlms, false, 128)
    AEC_TRY_INSTANCE(8, 16, kAlgoNlms, true, 128)
    AEC_TRY_INSTANCE(8, 8, kAlgoKalman, false, 128)
    AEC_TRY_INSTANCE(8, 8, kAlgoKalman, true, 128)
    AEC_TRY_INSTANCE(8, 8, kAlgoNlms, false, 128)
    AEC_TRY_INSTANCE(8, 8, kAlgoNlms, true, 128)
    return kNoInstance;
}

}  // namespace aec
