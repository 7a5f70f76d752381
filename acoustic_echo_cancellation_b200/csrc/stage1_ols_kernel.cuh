// Stage-1 linear echo canceller, algo = 2: overlap-save partitioned-block FDAF with the alternated gradient
// constraint (time-domain blocks of H = 256 new samples, FFT length 512, P partitions).  One persistent CTA of two
// warps per utterance; filter taps, far-end history and the smoothed input power stay in registers for the whole
// utterance, like the STFT-domain kernels (stage1_kernel.cuh), whose half-warp FFT-256, real-FFT split and bin
// ownership it shares.  Unlike them the recurrence crosses the transform every block -- e(t) needs W(t), W(t) needs
// E(t-1) = FFT(e(t-1)) -- so there is no chunk of frames to batch: per block
//   (1) stage the far-end / microphone block                                   (all threads, prefetched one block ahead)
//   (2) X_t = FFT[x_{t-1}, x_t]                                                (one half-warp)
//   (3) Yhat = sum_p W_p X_{t-p}, packed for the inverse transform             (64 threads, 4 bins each + bin 128)
//   (4) y = IFFT(Yhat)[H:],  e = d - y -> HBM,  E = FFT[0, e]                  (one half-warp, two dependent transforms)
//   (5) Pw = lam Pw + (1 - lam) sum |X_p|^2,  W_p += mu conj(X_p) E / (Pw + delta);  W_c (c = t mod P) packed
//   (6) g = IFFT(W_c), g[H:] = 0, W_c = FFT(g)                                 (one half-warp, two dependent transforms)
//   (7) W_c back into the owners' registers
// Why it exists: the STFT-domain recurrence (Hann analysis window, no cross-band terms) cancels ~13 dB on the SURVEY 8d
// single-talk set; the exact linear convolution of this one reaches the 40 dB noise floor.
// BUILDER-AUTHORED (the reference has no stage-1 filter): restated by oracle/aec_oracle.py:pbfdaf_ols, parity unpinned.
#pragma once
#include "stage1_kernel.cuh"

namespace aec {

struct OlsSmem {
    static constexpr size_t tile_bytes = size_t(3) * kTilePitch * sizeof(float2);   // X, Yhat / E, W_c
    static constexpr size_t blk_bytes = size_t(3) * 256 * sizeof(float);            // far-end ring [2][256], microphone [256]
    __host__ __device__ static constexpr size_t total(int P) {
        return tile_bytes + blk_bytes + (size_t(P) * 16 + 16 + 15) / 16 * 16 + 64;
    }
};

template <int P, bool ECHO, int REGS>
__global__ void __launch_bounds__(64) __maxnreg__(REGS) stage1_ols_kernel(const Stage1Params prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* tileX = reinterpret_cast<float2*>(smem_raw);
    float2* tileY = tileX + kTilePitch;
    float2* tileW = tileY + kTilePitch;
    float* xblk = reinterpret_cast<float*>(tileW + kTilePitch);      // [2][256] far-end blocks t-1, t (slot = t & 1)
    float* dblk = xblk + 512;                                        // [256]    microphone block t
    float2* midW = reinterpret_cast<float2*>(dblk + 256);            // [P] taps of bin 128 (its own mirror)
    float2* midX = midW + P;                                         // [P] its far-end history
    float* midPw = reinterpret_cast<float*>(midX + P);               // [1]
    int* fft_warp_s = reinterpret_cast<int*>(midPw + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = lane >> 4, h = lane & 15;

    long long n_ll = prm.n_samples ? prm.n_samples[blockIdx.x] : prm.L;
    n_ll = n_ll < 0 ? 0 : (n_ll > prm.L ? prm.L : n_ll);
    const int nblk = static_cast<int>(n_ll / 256);
    const float* far_b = prm.far + static_cast<long long>(blockIdx.x) * prm.in_stride;
    const float* mic_b = prm.mic + static_cast<long long>(blockIdx.x) * prm.in_stride;
    float* err_b = prm.err + static_cast<long long>(blockIdx.x) * prm.out_stride;
    float* echo_b = ECHO ? prm.echo + static_cast<long long>(blockIdx.x) * prm.out_stride : nullptr;

    // The transforms of an utterance run on ONE half-warp.  Warp w of a two-warp CTA always sits on scheduler
    // (slot + w) % 4, so a fixed choice would put every resident utterance's transforms on two of the four schedulers:
    // alternate with the hardware warp slot (same device as the bin-128 owner of the STFT-domain kernel).
    if (tid == 0) {
        unsigned hw_warp;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw_warp));
        *fft_warp_s = static_cast<int>((hw_warp >> 2) & 1u);
    }
    if (tid < P) {
        midW[tid] = make_float2(0.f, 0.f);
        midX[tid] = make_float2(0.f, 0.f);
    }
    if (tid == 0) *midPw = 0.f;
    for (int i = tid; i < 512; i += 64) xblk[i] = 0.f;               // x_{-1} = 0
    __syncthreads();
    const int fw = *fft_warp_s;
    const bool fft_lane = (warp == fw);
    const int mid_tid = (fw ^ 1) * 32 + 31;                          // bin 128: last lane of the other warp

    TwiddleRegs twr;
    twr.w1 = __ldg(&prm.tw256[1 * 16 + h]);
    twr.w2 = __ldg(&prm.tw256[2 * 16 + h]);
    twr.w4 = __ldg(&prm.tw256[4 * 16 + h]);
    twr.w8 = __ldg(&prm.tw256[8 * 16 + h]);

    // ---- persistent per-bin state: thread owns the mirrored pairs (k, 256 - k), k = tid, tid + 64 ----
    float2 W[4][P], X[4][P];
    float pw[4];
    float2 wk[2];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        pw[b] = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            W[b][p] = make_float2(0.f, 0.f);
            X[b][p] = make_float2(0.f, 0.f);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) wk[i] = __ldg(&prm.tw512[tid + 64 * i]);
    const float2 w_mid = make_float2(0.f, -1.f);

    // block loads: 64 threads x 4 samples per signal, prefetched one block ahead
    auto load4 = [&](const float* row, int t) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < nblk) {
            const float* p = row + static_cast<long long>(t) * 256 + 4 * tid;
            if (prm.use_tma) v = __ldg(reinterpret_cast<const float4*>(p));
            else v = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
        }
        return v;
    };
    float4 nx = load4(far_b, 0), nd = load4(mic_b, 0);
    float acc_m = 0.f, acc_e = 0.f;                                   // ERLE energies (transform half-warp only)
    const float k512 = 1.0f / 512.0f;

    for (int t = 0; t < nblk; ++t) {
        // ---- (1) stage the block ----
        *reinterpret_cast<float4*>(xblk + (t & 1) * 256 + 4 * tid) = nx;
        *reinterpret_cast<float4*>(dblk + 4 * tid) = nd;
        nx = load4(far_b, t + 1);
        nd = load4(mic_b, t + 1);
        __syncthreads();
        // ---- (2) X_t = FFT[x_{t-1}, x_t]  (the 0.5 of the real-FFT split rides on the input) ----
        if (fft_lane) {
            float2 v[16];
            const float* prev = xblk + ((t + 1) & 1) * 256 + 2 * h;
            const float* cur = xblk + (t & 1) * 256 + 2 * h;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 x = *reinterpret_cast<const float2*>((j < 8 ? prev : cur) + 32 * (j & 7));
                v[j] = make_float2(0.5f * x.x, 0.5f * x.y);
            }
            float2* tile = half == 0 ? tileX : tileW;                 // (the upper half-warp transforms along, into a free tile)
            fft256_halfwarp_regs<false>(v, tile, twr, h);
            if (half == 0) {
#pragma unroll
                for (int p = 0; p < 16; ++p) tileX[h + 16 * fft16_index(p)] = v[p];
            }
        }
        __syncthreads();
        // ---- (3) echo estimate in the frequency domain ----
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int k = tid + 64 * i, km = (256 - k) & 255;
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
        }
        if (tid == mid_tid) {
            float2 xk, xm, gk, gm;
            unpack_pair(tileX[128], tileX[128], w_mid, xk, xm);
#pragma unroll
            for (int p = P - 1; p > 0; --p) midX[p] = midX[p - 1];
            midX[0] = xk;
            float2 y = make_float2(0.f, 0.f);
#pragma unroll
            for (int p = 0; p < P; ++p) y = cfma(midW[p], midX[p], y);
            pack_pair(y, y, w_mid, gk, gm);
            tileY[128] = gk;
        }
        __syncthreads();
        // ---- (4) y = IFFT(Yhat)[H:], e = d - y, E = FFT[0, e] ----
        if (fft_lane) {
            float2 v[16];
            float2* tile = half == 0 ? tileY : tileW;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = tileY[h + 16 * j];
            __syncwarp();
            fft256_halfwarp_regs<true>(v, tile, twr, h);
            // register position p holds z[m], m = h + 16 r, r = fft16_index(p): samples 2m, 2m+1; the second half of the
            // 512 samples (r >= 8) is the linear-convolution part
            float2 u[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) u[fft16_index(p)] = v[p];
            float em = 0.f, ee = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float2 y = make_float2(u[8 + r].x * k512, u[8 + r].y * k512);
                const float2 d = *reinterpret_cast<const float2*>(dblk + 2 * h + 32 * r);
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
                v[r] = make_float2(0.f, 0.f);                          // E = FFT[0_H, e]
                v[8 + r] = make_float2(0.5f * e.x, 0.5f * e.y);
            }
            if (half == 0 && t >= prm.erle_skip_hops) {
                acc_m += em;
                acc_e += ee;
            }
            __syncwarp();
            fft256_halfwarp_regs<false>(v, tile, twr, h);
            if (half == 0) {
#pragma unroll
                for (int p = 0; p < 16; ++p) tileY[h + 16 * fft16_index(p)] = v[p];
            }
        }
        __syncthreads();
        // ---- (5) power, weight update, partition c = t mod P packed for the constraint ----
        const int c = t % P;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int k = tid + 64 * i, km = (256 - k) & 255;
            float2 ek, em2, gk, gm;
            unpack_pair(tileY[k], tileY[km], wk[i], ek, em2);
            float2 wc[2];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int bi = 2 * i + b;
                float s = 0.f;
#pragma unroll
                for (int p = 0; p < P; ++p) s = fmaf(X[bi][p].x, X[bi][p].x, fmaf(X[bi][p].y, X[bi][p].y, s));
                pw[bi] = fmaf(prm.pblam, pw[bi], prm.pboml * s);
                const float g = prm.mu * rcp_fast(pw[bi] + prm.delta);
                const float2 e = b ? em2 : ek;
                const float2 ge = make_float2(g * e.x, g * e.y);
#pragma unroll
                for (int p = 0; p < P; ++p) W[bi][p] = cfmac(X[bi][p], ge, W[bi][p]);
                wc[b] = W[bi][0];
#pragma unroll
                for (int p = 1; p < P; ++p)
                    if (c == p) wc[b] = W[bi][p];
            }
            pack_pair(wc[0], wc[1], wk[i], gk, gm);
            tileW[k] = gk;
            tileW[km] = gm;
        }
        if (tid == mid_tid) {
            float2 ek, em2, gk, gm;
            unpack_pair(tileY[128], tileY[128], w_mid, ek, em2);
            float s = 0.f;
#pragma unroll
            for (int p = 0; p < P; ++p) s = fmaf(midX[p].x, midX[p].x, fmaf(midX[p].y, midX[p].y, s));
            const float pwm = fmaf(prm.pblam, *midPw, prm.pboml * s);
            *midPw = pwm;
            const float g = prm.mu * rcp_fast(pwm + prm.delta);
            const float2 ge = make_float2(g * ek.x, g * ek.y);
#pragma unroll
            for (int p = 0; p < P; ++p) midW[p] = cfmac(midX[p], ge, midW[p]);
            const float2 wcm = midW[c];
            pack_pair(wcm, wcm, w_mid, gk, gm);
            tileW[128] = gk;
        }
        __syncthreads();
        // ---- (6) gradient constraint of partition c: g = IFFT(W_c), g[H:] = 0, W_c = FFT(g) ----
        if (fft_lane) {
            float2 v[16];
            float2* tile = half == 0 ? tileW : tileX;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = tileW[h + 16 * j];
            __syncwarp();
            fft256_halfwarp_regs<true>(v, tile, twr, h);
            float2 u[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) u[fft16_index(p)] = v[p];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                v[r] = make_float2(u[r].x * (0.5f * k512), u[r].y * (0.5f * k512));
                v[8 + r] = make_float2(0.f, 0.f);
            }
            __syncwarp();
            fft256_halfwarp_regs<false>(v, tile, twr, h);
            if (half == 0) {
#pragma unroll
                for (int p = 0; p < 16; ++p) tileW[h + 16 * fft16_index(p)] = v[p];
            }
        }
        __syncthreads();
        // ---- (7) the constrained partition back into its owners' registers ----
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int k = tid + 64 * i, km = (256 - k) & 255;
            float2 a, b2;
            unpack_pair(tileW[k], tileW[km], wk[i], a, b2);
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (c == p) {
                    W[2 * i][p] = a;
                    W[2 * i + 1][p] = b2;
                }
        }
        if (tid == mid_tid) {
            float2 a, b2;
            unpack_pair(tileW[128], tileW[128], w_mid, a, b2);
            midW[c] = a;
        }
        // (no barrier: the next block's first shared-memory writes go to xblk / dblk, its first tile write follows one)
    }

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
        if (fft_lane && lane == 0)
            prm.erle_db[blockIdx.x] = 10.f * log10f(fmaxf(acc_m, 1e-20f) / fmaxf(acc_e, 1e-20f));
    }
}

// stage1_inst_ols.cu
cudaError_t launch_stage1_ols(int P, bool echo, const Stage1Params& prm, cudaStream_t s);

}  // namespace aec
