// Instantiations of the frame-1024 (48 kHz) kernel: 8 warps, one mirrored pair per thread.
#include "stage1_kernel_1024.cuh"

namespace aec {

#define AEC_TRY_1024(P_, ALGO_, ECHO_, REGS_)                                                 \
    if (P == (P_) && algo == (ALGO_) && echo == (ECHO_) && (regs == (REGS_) || regs == 0))    \
        return launch_stage1_1024_instance<P_, ALGO_, ECHO_, REGS_>(prm, s);

cudaError_t launch_stage1_1024(int P, int algo, bool echo, int regs, const Stage1Params& prm, cudaStream_t s) {
    AEC_TRY_1024(8, kAlgoNlms, false, 128)
    AEC_TRY_1024(8, kAlgoNlms, true, 128)
    AEC_TRY_1024(8, kAlgoKalman, false, 128)
    AEC_TRY_1024(8, kAlgoKalman, true, 128)
    AEC_TRY_1024(4, kAlgoNlms, false, 128)
    AEC_TRY_1024(4, kAlgoNlms, true, 128)
    AEC_TRY_1024(4, kAlgoKalman, false, 128)
    AEC_TRY_1024(4, kAlgoKalman, true, 128)
    AEC_TRY_1024(2, kAlgoNlms, false, 128)
    AEC_TRY_1024(2, kAlgoNlms, true, 128)
    AEC_TRY_1024(2, kAlgoKalman, false, 128)
    AEC_TRY_1024(2, kAlgoKalman, true, 128)
    AEC_TRY_1024(1, kAlgoNlms, false, 128)
    AEC_TRY_1024(1, kAlgoNlms, true, 128)
    AEC_TRY_1024(1, kAlgoKalman, false, 128)
    AEC_TRY_1024(1, kAlgoKalman, true, 128)
    return kNoInstance;
}

}  // namespace aec
