// Launcher glue between the C ABI and the templated stage-1 kernels.  The kernels are
// instantiated in stage1_inst_*.cu (one translation unit per warps-per-utterance so the
// build parallelises); this header declares the per-unit entry points.
#pragma once
#include <cuda_runtime.h>

#include "stage1_kernel.cuh"

namespace aec {

// Returns cudaSuccess after launching, kNoInstance when (P, algo, echo, regs) is not instantiated in that
// unit ("not built" -> AEC_EUNSUPPORTED at the ABI) -- a code no launch path of the runtime produces, so that
// genuine cudaErrorInvalidValue failures (attributes, occupancy query, launch) surface as AEC_ECUDA
// (kNoInstance is declared in stage1_kernel.cuh).
cudaError_t launch_stage1_nw1(int P, int algo, bool echo, int regs, const Stage1Params& prm, cudaStream_t s);
cudaError_t launch_stage1_nw2(int P, int algo, bool echo, int regs, const Stage1Params& prm, cudaStream_t s);
cudaError_t launch_stage1_nw4(int P, int algo, bool echo, int regs, const Stage1Params& prm, cudaStream_t s);
cudaError_t launch_stage1_nw8(int P, int algo, bool echo, int regs, const Stage1Params& prm, cudaStream_t s);

// stage 1 with the fused Stage-2 feature epilogue (stage1_inst_feat.cu)
cudaError_t launch_stage1_feat(int P, int algo, const Stage1Params& prm, cudaStream_t s);

template <int NW, int P_, int ALGO, bool ECHO, int REGS, bool FEAT = false>
inline cudaError_t launch_stage1_instance(const Stage1Params& prm, cudaStream_t s) {
    auto kern = stage1_n512_kernel<NW, P_, ALGO, ECHO, REGS, FEAT>;
    const size_t smem = FEAT ? Stage1Smem<NW, P_>::total_feat(ECHO) : Stage1Smem<NW, P_>::total(ECHO);
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (configured_dev != dev) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // The kernel keeps its tables in shared memory / registers and does not need L1: ask for
        // the largest carve-out so that 7 two-warp utterances (31.8 KB each) fit on one SM ...
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        // ... but no larger than the resident utterances need: the long-filter kernels (2 utterances of
        // 60-90 KB per SM) still spill a few registers at their 128-register cap, and what is left of
        // the 256 KB array serves those reloads from L1 instead of L2.
        int resident = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, NW * 32, smem);
        if (e != cudaSuccess) return e;
        if (resident > 0) {
            const size_t need = size_t(resident) * (smem + 1024);
            int pct = static_cast<int>((need * 100 + 228 * 1024 - 1) / (228 * 1024));
            pct = pct > 100 ? 100 : pct;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            if (e != cudaSuccess) return e;
        }
        configured_dev = dev;
    }
    kern<<<dim3((unsigned)prm.B), dim3(NW * 32), smem, s>>>(prm);
    return cudaGetLastError();
}

#define AEC_TRY_INSTANCE(NW, P_, ALGO_, ECHO_, REGS_)                                         \
    if (P == (P_) && algo == (ALGO_) && echo == (ECHO_) && (regs == (REGS_) || regs == 0))    \
        return launch_stage1_instance<NW, P_, ALGO_, ECHO_, REGS_>(prm, s);

}  // namespace aec
