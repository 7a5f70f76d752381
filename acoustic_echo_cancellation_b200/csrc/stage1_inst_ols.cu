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

// regs: register cap asked for through aec_cfg.variant (0 = the default below)
cudaError_t launch_stage1_ols(int P, bool kalman, bool echo, int regs, const Stage1Params& prm, cudaStream_t s) {
    if (P == 4) {
        // four-partition Kalman step: 20 more state registers per thread.  The 128-register build (8 utterances per SM)
        // spills ~50 values per block into L1, the 168-register one (6 per SM) is spill-free: 3.42 against 4.74 ms per
        // 1024 x 10 s (the 168 build needs a second wave there), 12.4 against 11.9 ms per 4144 x 10 s -> 168 from 24
        // utterances per SM on, 128 below.
        if (kalman && regs == 0) regs = prm.B >= 24LL * prm.num_sms ? 168 : 128;
        if (kalman && regs == 168) return launch_ols_p<4, 128, 168>(kalman, echo, prm, s);
        if (regs == 0 || regs == 128) return launch_ols_p<4, 128, 128>(kalman, echo, prm, s);
        // (NLMS step at 168 registers, spill-free: 6 % shorter block latency -- 3660 against 3900 cycles alone on an SM --
        //  which does not pay for 6 instead of 8 resident utterances; variant 2168)
        if (!kalman && regs == 168) return launch_ols_p<4, 168, 168>(kalman, echo, prm, s);
        return kNoInstance;
    }
    if (regs != 0 && regs != 128) return kNoInstance;
    if (P == 2) return launch_ols_p<2, 128, 128>(kalman, echo, prm, s);
    if (P == 1) return launch_ols_p<1, 128, 128>(kalman, echo, prm, s);
    return kNoInstance;
}

}  // namespace aec
