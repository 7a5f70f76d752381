// Instantiations of the overlap-save PBFDAF kernel (algo = 2): two warps per utterance, 1-4 partitions.
#include "stage1_ols_kernel.cuh"

namespace aec {

template <int P, bool ECHO, int REGS>
static cudaError_t launch_ols(const Stage1Params& prm, cudaStream_t s) {
    auto kern = stage1_ols_kernel<P, ECHO, REGS>;
    const size_t smem = OlsSmem::total(P);
    kern<<<dim3((unsigned)prm.B), dim3(64), smem, s>>>(prm);
    return cudaGetLastError();
}

cudaError_t launch_stage1_ols(int P, bool echo, const Stage1Params& prm, cudaStream_t s) {
    if (P == 4) return echo ? launch_ols<4, true, 168>(prm, s) : launch_ols<4, false, 168>(prm, s);
    if (P == 2) return echo ? launch_ols<2, true, 128>(prm, s) : launch_ols<2, false, 128>(prm, s);
    if (P == 1) return echo ? launch_ols<1, true, 128>(prm, s) : launch_ols<1, false, 128>(prm, s);
    return kNoInstance;
}

}  // namespace aec
