// Instantiations of the two-warp kernel WITH the fused Stage-2 feature epilogue: the error signal
// is re-analysed on chip one frame behind its synthesis and cat[err_erb, |err_erb - far_erb|] is emitted next to it.
// 43.9 KB of shared memory per utterance -> 5 utterances per SM; the register cap that goes with five two-warp
// utterances is 65536 / (5 * 64) = 204 (the plain kernel runs seven per SM at 128).
#include "stage1_launch.cuh"

namespace aec {

cudaError_t launch_stage1_feat(int P, int algo, const Stage1Params& prm, cudaStream_t s) {
    if (P == 4 && algo == kAlgoNlms) return launch_stage1_instance<2, 4, kAlgoNlms, false, 200, true>(prm, s);
    if (P == 4 && algo == kAlgoKalman) return launch_stage1_instance<2, 4, kAlgoKalman, false, 200, true>(prm, s);
    if (P == 2 && algo == kAlgoNlms) return launch_stage1_instance<2, 2, kAlgoNlms, false, 200, true>(prm, s);
    if (P == 1 && algo == kAlgoNlms) return launch_stage1_instance<2, 1, kAlgoNlms, false, 200, true>(prm, s);
    return kNoInstance;
}

}  // namespace aec
