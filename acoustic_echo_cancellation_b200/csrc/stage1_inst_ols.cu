// Instantiations of the overlap-save PBFDAF kernel (algo = 2: NLMS step, algo = 3: Kalman step): two warps per
// utterance, 1-4 partitions.
#include "stage1_ols_kernel.cuh"

namespace aec {

template <int P, bool KAL, bool ECHO, int REGS>
static cudaError_t launch_ols(const Stage1Params& prm, cudaStream_t s) {
    auto kern = stage1_ols_kernel<P, KAL, ECHO, REGS>;
    const size_t smem = OlsSmem::total(P);
    kern<<<dim3((unsigned)prm.B), dim3(64), smem, s>>>(prm);
    return cudaGetLastError();
}

template <int P, int REGS_NLMS, int REGS_KAL>
static cudaError_t launch_ols_p(bool kalman, bool echo, const Stage1Params& prm, cudaStream_t s) {
    if (kalman) return echo ? launch_ols<P, true, true, REGS_KAL>(prm, s) : launch_ols<P, true, false, REGS_KAL>(prm, s);
    return echo ? launch_ols<P, false, true, REGS_NLMS>(prm, s) : launch_ols<P, false, false, REGS_NLMS>(prm, s);
}

cudaError_t launch_stage1_ols(int P, bool kalman, bool echo, const Stage1Params& prm, cudaStream_t s) {
    if (P == 4) return launch_ols_p<4, 128, 168>(kalman, echo, prm, s);
    if (P == 2) return launch_ols_p<2, 128, 128>(kalman, echo, prm, s);
    if (P == 1) return launch_ols_p<1, 128, 128>(kalman, echo, prm, s);
    return kNoInstance;
}

}  // namespace aec
