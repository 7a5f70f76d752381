// Stage-1 linear echo canceller, algo = 2 / 3: overlap-save partitioned-block FDAF with the alternated gradient
// constraint (time-domain blocks of H = 256 new samples, FFT length 512, P partitions), NLMS step (algo 2) or the
// diagonal Kalman step of algo 1 (algo 3).  One persistent CTA of two warps per utterance; filter taps, far-end
// history and the power / covariance state stay in registers for the whole utterance, like the STFT-domain kernels
// (stage1_kernel.cuh), whose half-warp FFT-256, real-FFT split and bin ownership it shares.
//
// Unlike them the recurrence crosses the transform every block -- e(t) needs W(t), W(t) needs E(t-1) = FFT(e(t-1)) --
// so there is no chunk of frames to batch and the length of the per-block dependency chain is what counts.  Five
// transforms per block (X, y, E and the two of the constraint) are arranged so that only TWO are on the chain:
//   * the constraint is applied to partition c = t mod P AS IT ENTERED the block (oracle/aec_oracle.py:pbfdaf_ols):
//     it does not wait for E, so IFFT(W_c) runs on the upper half-warp beside IFFT(Yhat) on the lower one, and
//     FFT(g) beside FFT([0, e]) -- one warp, two transform slots, both half-warps busy;
//   * X_{t+1} = FFT[x_t, x_{t+1}] does not depend on the filter at all: the OTHER warp computes it during the same phase.
// Per block t, two block barriers:
//   R  (64 threads, 4 bins each + bin 128): [t >= 1: E_{t-1}, constrained W_c back into registers, power / covariance,
//      weight update]  then  X_t into the history, Yhat = sum_p W_p X_{t-p} packed for the inverse transform, W_c packed
//   F  warp a = (slot + t) & 1:  y = IFFT(Yhat)[H:], e = d - y -> HBM, E = FFT[0, e]  ||  g = IFFT(W_c), g[H:] = 0, FFT(g)
//      warp a ^ 1:               X_{t+1}
// (a thread touches only its own bins' tile entries in R, so the update of block t-1 and the estimate of block t need
// no barrier between them).  Far-end / microphone blocks are staged HBM -> shared memory with cp.async (LDGSTS, 16 B
// per thread, no registers) two / one blocks ahead.
// Why it exists: the STFT-domain recurrence (Hann analysis window, no cross-band terms) cancels ~13 dB on the SURVEY 8d
// single-talk set; the exact linear convolution of this one reaches the 40 dB noise floor, and with the Kalman step it
// holds 14 dB through double talk.
// BUILDER-AUTHORED (the reference has no stage-1 filter): restated by oracle/aec_oracle.py:pbfdaf_ols, parity unpinned.
#pragma once
#include "stage1_kernel.cuh"

namespace aec {

struct OlsSmem {
    static constexpr size_t tile_bytes = size_t(4) * kTilePitch * sizeof(float2);   // X, Yhat / E, W_c, scratch
    static constexpr size_t blk_bytes = size_t(4 + 2) * 256 * sizeof(float);        // far-end ring [4][256], microphone [2][256]
    __host__ __device__ static constexpr size_t total(int P) {
        return tile_bytes + blk_bytes + (size_t(P) * 20 + 16 + 15) / 16 * 16 + 64;
    }
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// weight update of one bin with the error spectrum E of the block (operation order of oracle/aec_oracle.py)
template <int P, bool KAL>
__device__ __forceinline__ void ols_update(float2 (&W)[P], const float2 (&X)[P], float (&C)[KAL ? P : 1], float& sp,
                                           const float2 E, const Stage1Params& prm) {
    if constexpr (!KAL) {
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) s = fmaf(X[p].x, X[p].x, fmaf(X[p].y, X[p].y, s));
        sp = fmaf(prm.pblam, sp, prm.pboml * s);
        const float g = prm.mu * rcp_fast(sp + prm.delta);
        const float2 ge = make_float2(g * E.x, g * E.y);
#pragma unroll
        for (int p = 0; p < P; ++p) W[p] = cfmac(X[p], ge, W[p]);
    } else {
        const float e2 = fmaf(E.x, E.x, E.y * E.y);
        sp = fmaf(prm.klam, sp, prm.koml * e2);
        float d = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) d = fmaf(C[p], fmaf(X[p].x, X[p].x, X[p].y * X[p].y), d);
        d = d + sp + prm.keps;
        const float rd = __frcp_rn(d);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float x2 = fmaf(X[p].x, X[p].x, X[p].y * X[p].y);
            const float gs = C[p] * rd;
            const float2 g = make_float2(gs * X[p].x, -gs * X[p].y);      // C conj(X) / D
            float2 w = cfma(g, E, W[p]);
            w = make_float2(prm.ka * w.x, prm.ka * w.y);
            W[p] = w;
            C[p] = fmaf(prm.ka2 * (1.f - gs * x2), C[p], prm.kq * fmaf(w.x, w.x, w.y * w.y));
        }
    }
}

template <int P, bool KAL, bool ECHO, int REGS>
__global__ void __launch_bounds__(64) __maxnreg__(REGS) stage1_ols_kernel(const Stage1Params prm) {
    constexpr int PC = KAL ? P : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* tileX = reinterpret_cast<float2*>(smem_raw);
    float2* tileY = tileX + kTilePitch;
    float2* tileW = tileY + kTilePitch;
    float2* tileS = tileW + kTilePitch;                              // exchange tile of the idle half-warp of the X transform
    float* xring = reinterpret_cast<float*>(tileS + kTilePitch);     // [4][256] far-end blocks, slot = block & 3
    float* dring = xring + 4 * 256;                                  // [2][256] microphone blocks, slot = block & 1
    float2* midW = reinterpret_cast<float2*>(dring + 2 * 256);       // [P] taps of bin 128 (its own mirror)
    float2* midX = midW + P;                                         // [P] its far-end history
    float* midC = reinterpret_cast<float*>(midX + P);                // [P] covariances (Kalman)
    float* midS = midC + P;                                          // [1] smoothed power / Psi
    float* red = midS + 1;                                           // [4] ERLE energies of the two warps
    int* fft_warp_s = reinterpret_cast<int*>(red + 4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = lane >> 4, h = lane & 15;

    long long n_ll = prm.n_samples ? prm.n_samples[blockIdx.x] : prm.L;
    n_ll = n_ll < 0 ? 0 : (n_ll > prm.L ? prm.L : n_ll);
    const int nblk = static_cast<int>(n_ll / 256);
    const float* far_b = prm.far + static_cast<long long>(blockIdx.x) * prm.in_stride;
    const float* mic_b = prm.mic + static_cast<long long>(blockIdx.x) * prm.in_stride;
    float* err_b = prm.err + static_cast<long long>(blockIdx.x) * prm.out_stride;
    float* echo_b = ECHO ? prm.echo + static_cast<long long>(blockIdx.x) * prm.out_stride : nullptr;

    // Warp w of a two-warp CTA always sits on scheduler (slot + w) % 4, and the warp that carries the two-transform
    // chain is the busier one: which warp starts with it follows the hardware warp slot, so that co-resident utterances
    // on the same scheduler pair load different schedulers (same device as the bin-128 owner of the STFT-domain kernel);
    // the roles then swap every block.
    if (tid == 0) {
        unsigned hw_warp;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw_warp));
        *fft_warp_s = static_cast<int>((hw_warp >> 2) & 1u);
    }
    if (tid < P) {
        midW[tid] = make_float2(0.f, 0.f);
        midX[tid] = make_float2(0.f, 0.f);
        midC[tid] = prm.kc0;
    }
    if (tid == 0) *midS = 0.f;
    for (int i = tid; i < 256; i += 64) xring[3 * 256 + i] = 0.f;    // x_{-1} = 0 (slot of block -1)
    __syncthreads();
    const int fw = *fft_warp_s;

    TwiddleRegs twr;
    twr.w1 = __ldg(&prm.tw256[1 * 16 + h]);
    twr.w2 = __ldg(&prm.tw256[2 * 16 + h]);
    twr.w4 = __ldg(&prm.tw256[4 * 16 + h]);
    twr.w8 = __ldg(&prm.tw256[8 * 16 + h]);

    // ---- persistent per-bin state: thread owns the mirrored pairs (k, 256 - k), k = tid, tid + 64 ----
    float2 W[4][P], X[4][P];
    float C[4][PC], sp[4];
    float2 wk[2];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        sp[b] = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            W[b][p] = make_float2(0.f, 0.f);
            X[b][p] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int p = 0; p < PC; ++p) C[b][p] = prm.kc0;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) wk[i] = __ldg(&prm.tw512[tid + 64 * i]);
    const float2 w_mid = make_float2(0.f, -1.f);

    // block staging: 64 threads x 4 samples per signal
    auto stage_block = [&](const float* row, float* dst, int blk) {
        if (blk < nblk) {
            const float* p = row + static_cast<long long>(blk) * 256 + 4 * tid;
            if (prm.use_tma) {
                cp_async16(dst + 4 * tid, p);
            } else {
                *reinterpret_cast<float4*>(dst + 4 * tid) = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
            }
        }
    };
    stage_block(far_b, xring, 0);
    cp_async_commit();

    float acc_m = 0.f, acc_e = 0.f;                                   // ERLE energies (lower half-warps)
    const float k512 = 1.0f / 512.0f;

    for (int t = -1; t < nblk; ++t) {
        stage_block(far_b, xring + ((t + 2) & 3) * 256, t + 2);
        stage_block(mic_b, dring + ((t + 1) & 1) * 256, t + 1);
        cp_async_commit();
        const int a = (fw + t) & 1;                                   // the warp that carries the chain of this block
        // ---- R: update with E_{t-1}, echo estimate of block t ----
        if (t >= 0) {
            const int c = t % P, cprev = (t + P - 1) % P;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int k = tid + 64 * i, km = (256 - k) & 255;
                if (t >= 1) {
                    float2 ek, em, ca, cb;
                    unpack_pair(tileY[k], tileY[km], wk[i], ek, em);
                    unpack_pair(tileW[k], tileW[km], wk[i], ca, cb);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        if (cprev == p) {
                            W[2 * i][p] = ca;
                            W[2 * i + 1][p] = cb;
                        }
                    ols_update<P, KAL>(W[2 * i], X[2 * i], C[2 * i], sp[2 * i], ek, prm);
                    ols_update<P, KAL>(W[2 * i + 1], X[2 * i + 1], C[2 * i + 1], sp[2 * i + 1], em, prm);
                }
                float2 xk, xm, gk, gm;
                unpack_pair(tileX[k], tileX[km], wk[i], xk, xm);
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    float2* Xb = X[2 * i + b];
#pragma unroll
                    for (int p = P - 1; p > 0; --p) Xb[p] = Xb[p - 1];
                    Xb[0] = b ? xm : xk;
                }
                float2 yk = make_float2(0.f, 0.f), ym = make_float2(0.f, 0.f);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    yk = cfma(W[2 * i][p], X[2 * i][p], yk);
                    ym = cfma(W[2 * i + 1][p], X[2 * i + 1][p], ym);
                }
                pack_pair(yk, ym, wk[i], gk, gm);
                tileY[k] = gk;
                tileY[km] = gm;
                float2 wa = W[2 * i][0], wb = W[2 * i + 1][0];
#pragma unroll
                for (int p = 1; p < P; ++p)
                    if (c == p) {
                        wa = W[2 * i][p];
                        wb = W[2 * i + 1][p];
                    }
                pack_pair(wa, wb, wk[i], gk, gm);
                tileW[k] = gk;
                tileW[km] = gm;
            }
            if (tid == (a ^ 1) * 32 + 31) {                           // bin 128, on the warp with the lighter transform phase
                float2 xk, xm, gk, gm;
                float2 mw[P], mx[P];
                float mc[PC], ms = *midS;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    mw[p] = midW[p];
                    mx[p] = midX[p];
                }
#pragma unroll
                for (int p = 0; p < PC; ++p) mc[p] = midC[p];
                if (t >= 1) {
                    float2 ek, em, ca, cb;
                    unpack_pair(tileY[128], tileY[128], w_mid, ek, em);
                    unpack_pair(tileW[128], tileW[128], w_mid, ca, cb);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        if (cprev == p) mw[p] = ca;
                    ols_update<P, KAL>(mw, mx, mc, ms, ek, prm);
                    *midS = ms;
#pragma unroll
                    for (int p = 0; p < PC; ++p) midC[p] = mc[p];
                }
                unpack_pair(tileX[128], tileX[128], w_mid, xk, xm);
#pragma unroll
                for (int p = P - 1; p > 0; --p) mx[p] = mx[p - 1];
                mx[0] = xk;
                float2 y = make_float2(0.f, 0.f);
#pragma unroll
                for (int p = 0; p < P; ++p) y = cfma(mw[p], mx[p], y);
                pack_pair(y, y, w_mid, gk, gm);
                tileY[128] = gk;
                float2 wc = mw[0];
#pragma unroll
                for (int p = 1; p < P; ++p)
                    if (c == p) wc = mw[p];
                pack_pair(wc, wc, w_mid, gk, gm);
                tileW[128] = gk;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    midW[p] = mw[p];
                    midX[p] = mx[p];
                }
            }
        }
        cp_async_wait<1>();                                           // blocks staged one iteration ago have landed
        __syncthreads();
        // ---- F: the two transform slots of the chain on warp a, X_{t+1} on the other warp ----
        if (warp == a) {
            if (t >= 0) {
                float2 v[16];
                float2* tile = half == 0 ? tileY : tileW;             // lower half-warp: error path; upper: constraint
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = tile[h + 16 * j];
                __syncwarp();
                fft256_halfwarp_regs<true>(v, tile, twr, h);
                // register position p holds z[m], m = h + 16 r, r = fft16_index(p): samples 2m, 2m+1; the second half of
                // the 512 samples (r >= 8) is the linear-convolution part of y, the first half (r < 8) the part of g kept
                float2 u[16];
#pragma unroll
                for (int p = 0; p < 16; ++p) u[fft16_index(p)] = v[p];
                const float* dsrc = dring + (t & 1) * 256 + 2 * h;
                const float sg = half == 0 ? 0.f : 0.5f * k512;
                float em = 0.f, ee = 0.f;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float2 y = make_float2(u[8 + r].x * k512, u[8 + r].y * k512);
                    const float2 d = *reinterpret_cast<const float2*>(dsrc + 32 * r);
                    const float2 e = make_float2(d.x - y.x, d.y - y.y);
                    if (half == 0) {
                        float* dst = err_b + static_cast<long long>(t) * 256 + 2 * h + 32 * r;
                        if (prm.vec_out) st_stream_f2(dst, e);
                        else { st_stream_f1(dst, e.x); st_stream_f1(dst + 1, e.y); }
                        if constexpr (ECHO) {
                            float* dy = echo_b + static_cast<long long>(t) * 256 + 2 * h + 32 * r;
                            if (prm.vec_out) st_stream_f2(dy, y);
                            else { st_stream_f1(dy, y.x); st_stream_f1(dy + 1, y.y); }
                        }
                    }
                    em = fmaf(d.x, d.x, fmaf(d.y, d.y, em));
                    ee = fmaf(e.x, e.x, fmaf(e.y, e.y, ee));
                    v[r] = make_float2(u[r].x * sg, u[r].y * sg);                   // [g, 0_H] / [0_H, e]
                    v[8 + r] = half == 0 ? make_float2(0.5f * e.x, 0.5f * e.y) : make_float2(0.f, 0.f);
                }
                if (half == 0 && t >= prm.erle_skip_hops) {
                    acc_m += em;
                    acc_e += ee;
                }
                __syncwarp();
                fft256_halfwarp_regs<false>(v, tile, twr, h);
#pragma unroll
                for (int p = 0; p < 16; ++p) tile[h + 16 * fft16_index(p)] = v[p];
            }
        } else if (t + 1 < nblk) {
            // X_{t+1} = FFT[x_t, x_{t+1}]  (the 0.5 of the real-FFT split rides on the input; the upper half-warp
            // transforms along into the scratch tile)
            float2 v[16];
            const float* prev = xring + (t & 3) * 256 + 2 * h;
            const float* cur = xring + ((t + 1) & 3) * 256 + 2 * h;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 x = *reinterpret_cast<const float2*>((j < 8 ? prev : cur) + 32 * (j & 7));
                v[j] = make_float2(0.5f * x.x, 0.5f * x.y);
            }
            float2* tile = half == 0 ? tileX : tileS;
            fft256_halfwarp_regs<false>(v, tile, twr, h);
#pragma unroll
            for (int p = 0; p < 16; ++p) tile[h + 16 * fft16_index(p)] = v[p];
        }
        __syncthreads();
    }
    cp_async_wait<0>();

    // ---- epilogue: zero the output beyond the last whole block, ERLE ----
    for (long long i = static_cast<long long>(nblk) * 256 + tid; i < prm.out_stride && i < prm.L; i += 64) {
        err_b[i] = 0.f;
        if constexpr (ECHO) echo_b[i] = 0.f;
    }
    if (prm.erle_db != nullptr) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            acc_m += __shfl_xor_sync(0xffffffffu, acc_m, o);
            acc_e += __shfl_xor_sync(0xffffffffu, acc_e, o);
        }
        if (lane == 0) {
            red[2 * warp] = acc_m;
            red[2 * warp + 1] = acc_e;
        }
        __syncthreads();
        if (tid == 0)
            prm.erle_db[blockIdx.x] = 10.f * log10f(fmaxf(red[0] + red[2], 1e-20f) / fmaxf(red[1] + red[3], 1e-20f));
    }
}

// stage1_inst_ols.cu
cudaError_t launch_stage1_ols(int P, bool kalman, bool echo, const Stage1Params& prm, cudaStream_t s);

}  // namespace aec
