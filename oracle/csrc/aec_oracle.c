/* CPU restatement (plain C, float32) of the stage-1 path.  TEST INFRASTRUCTURE ONLY.
 *
 * Used by tests/ (cross-check against oracle/aec_oracle.py), by bench.py's cpu_baseline leg and
 * by `bench.py --impl reference` (the reference has no CPU implementation of the FDAF to time;
 * this port is the CPU arm and is labelled "port", never "reference").  The product library
 * libaec_b200.so never links or calls this file.
 *
 * Parity status: the STFT / iSTFT steps restate the reference operators
 *   ConvSTFT.forward   Stage2_lhm/scripts/network/attention_ccrn.py:45-52  (zero pad N-H both
 *                      sides, hop-H frames, periodic Hann, rfft sign convention; kernel :8-25)
 *   ConviSTFT.forward  Stage2_lhm/scripts/network/attention_ccrn.py:82-101 (irfft * window,
 *                      overlap-add, / (sum window^2 + 1e-8), trim N-H both sides)
 * and are pinned through oracle/aec_oracle.py by the golden vectors in tests/golden/.
 * The FDAF recurrences (NLMS / Kalman) are BUILDER-AUTHORED: PARITY UNPINNED by the reference,
 * which contains no stage-1 filter.  They follow oracle/aec_oracle.py:fdaf_nlms / fdaf_kalman
 * (algos 0 / 1) and oracle/aec_oracle.py:pbfdaf_ols (algos 2 / 3, run_group_ols below) statement by statement.
 *
 * The reference computes each DFT as a dense [2K x N] convolution; this port uses an FFT
 * (same result to float32 rounding) so that the CPU baseline is not handicapped.  Since round 2 the transforms run
 * eight frames at a time in structure-of-arrays form and the recurrence is written as loops over the bins, so that the
 * compiler vectorises both (AVX2 / AVX-512 via -march=native); OpenMP parallelises over utterances.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct aec_oracle_cfg {
    int32_t frame;
    int32_t partitions;
    int32_t algo; /* 0 NLMS, 1 Kalman (STFT domain); 2 NLMS step, 3 Kalman step (overlap-save PBFDAF) */
    float mu, delta;
    float kalman_a, kalman_lambda, kalman_c0, kalman_eps;
    int32_t erle_skip_hops;
    float pb_lambda; /* algo 2: smoothing of the per-bin input power */
} aec_oracle_cfg;

typedef struct { float re, im; } cpx;

typedef struct plan {
    int n;          /* frame length */
    int m;          /* n / 2: complex FFT length */
    int log2m;
    cpx* tw;        /* m/2 twiddles exp(-2 pi i j / m) */
    cpx* tw_n;      /* m+1 twiddles exp(-2 pi i k / n) */
    int* rev;       /* bit reversal of m */
    float* win;     /* periodic Hann, n */
    float* norm;    /* 1 / (w[j]^2 + w[j+H]^2 + 1e-8), H entries */
} plan;

static plan* plan_create(int n) {
    plan* p = (plan*)calloc(1, sizeof(plan));
    p->n = n;
    p->m = n / 2;
    p->log2m = 0;
    while ((1 << p->log2m) < p->m) p->log2m++;
    p->tw = (cpx*)malloc(sizeof(cpx) * (size_t)(p->m / 2));
    p->tw_n = (cpx*)malloc(sizeof(cpx) * (size_t)(p->m + 1));
    p->rev = (int*)malloc(sizeof(int) * (size_t)p->m);
    p->win = (float*)malloc(sizeof(float) * (size_t)n);
    p->norm = (float*)malloc(sizeof(float) * (size_t)p->m);
    const double pi = 3.14159265358979323846;
    for (int j = 0; j < p->m / 2; ++j) {
        p->tw[j].re = (float)cos(-2.0 * pi * j / p->m);
        p->tw[j].im = (float)sin(-2.0 * pi * j / p->m);
    }
    for (int k = 0; k <= p->m; ++k) {
        p->tw_n[k].re = (float)cos(-2.0 * pi * k / n);
        p->tw_n[k].im = (float)sin(-2.0 * pi * k / n);
    }
    for (int i = 0; i < p->m; ++i) {
        int r = 0;
        for (int b = 0; b < p->log2m; ++b)
            if (i & (1 << b)) r |= 1 << (p->log2m - 1 - b);
        p->rev[i] = r;
    }
    for (int i = 0; i < n; ++i) p->win[i] = (float)(0.5 - 0.5 * cos(2.0 * pi * i / n)); /* attention_ccrn.py:12 */
    for (int j = 0; j < p->m; ++j) {
        const float a = p->win[j], b = p->win[j + p->m];
        p->norm[j] = 1.0f / (a * a + b * b + 1e-8f); /* attention_ccrn.py:95-97 */
    }
    return p;
}

static void plan_destroy(plan* p) {
    if (!p) return;
    free(p->tw); free(p->tw_n); free(p->rev); free(p->win); free(p->norm); free(p);
}

/* ---- transforms, SIMD across VL frames at a time ------------------------------------------------------------
 * In-place radix-2 DIT complex FFT of length m (sign = -1 forward, +1 inverse, unscaled) on the layout [point][VL]:
 * every butterfly is applied to VL independent frames at once, so the compiler can use full-width vector instructions.
 * A real frame of n samples is one complex FFT of n/2 points (even samples real lane, odd samples imaginary lane)
 * plus the usual split into K = n/2+1 bins (rfft convention, attention_ccrn.py:15-18); the inverse packs the bins
 * back (imaginary parts of DC / Nyquist ignored).  (Round-2 change, VERDICT r1 weak 7: no scalar FFT in the CPU arm.) */
#define VL 8

static void fft_c_v(const plan* p, float* restrict re, float* restrict im, int sign) {
    const int m = p->m;
    for (int i = 0; i < m; ++i) {
        const int r = p->rev[i];
        if (r > i) {
            for (int l = 0; l < VL; ++l) {
                float t = re[i * VL + l]; re[i * VL + l] = re[r * VL + l]; re[r * VL + l] = t;
                t = im[i * VL + l]; im[i * VL + l] = im[r * VL + l]; im[r * VL + l] = t;
            }
        }
    }
    for (int len = 2; len <= m; len <<= 1) {
        const int half = len >> 1, step = m / len;
        for (int s = 0; s < m; s += len) {
            for (int j = 0; j < half; ++j) {
                const cpx w = p->tw[j * step];
                const float wr = w.re, wi = sign < 0 ? w.im : -w.im;
                float* restrict ur = re + (size_t)(s + j) * VL;
                float* restrict ui = im + (size_t)(s + j) * VL;
                float* restrict vr = re + (size_t)(s + j + half) * VL;
                float* restrict vi = im + (size_t)(s + j + half) * VL;
#pragma omp simd
                for (int l = 0; l < VL; ++l) {
                    const float tr = vr[l] * wr - vi[l] * wi;
                    const float ti = vr[l] * wi + vi[l] * wr;
                    const float a = ur[l], b = ui[l];
                    ur[l] = a + tr; ui[l] = b + ti;
                    vr[l] = a - tr; vi[l] = b - ti;
                }
            }
        }
    }
}

/* one utterance */
static void run_one(const plan* p, const aec_oracle_cfg* cfg, const float* far, const float* mic, int64_t n,
                    int64_t out_len_total, float* err, float* echo, float* erle_db) {
    const int N = p->n, H = N / 2, K = H + 1, P = cfg->partitions, m = H;
    const int KP = (K + 15) & ~15;                   /* padded row length of the per-bin arrays */
    const int64_t T = n / H + 1;                     /* attention_ccrn.py:48-49 with N = 2H */
    const int64_t valid = (T - 1) * H;
    /* per-bin state, structure of arrays: [tap][bin] */
    float* Wr = (float*)calloc((size_t)P * KP, sizeof(float));
    float* Wi = (float*)calloc((size_t)P * KP, sizeof(float));
    float* Hr = (float*)calloc((size_t)P * KP, sizeof(float));   /* far-end history ring: slot (t - q) mod P */
    float* Hi = (float*)calloc((size_t)P * KP, sizeof(float));
    float* C = (float*)malloc(sizeof(float) * (size_t)P * KP);
    float* cx2 = (float*)malloc(sizeof(float) * (size_t)P * KP);
    float* psi = (float*)calloc((size_t)KP, sizeof(float));
    /* spectra of the VL frames of a chunk: [frame][bin] */
    float* Yr = (float*)malloc(sizeof(float) * (size_t)VL * KP);
    float* Yi = (float*)malloc(sizeof(float) * (size_t)VL * KP);
    float* Xr = (float*)malloc(sizeof(float) * (size_t)VL * KP);
    float* Xi = (float*)malloc(sizeof(float) * (size_t)VL * KP);
    float* Er = (float*)malloc(sizeof(float) * (size_t)VL * KP);
    float* Ei = (float*)malloc(sizeof(float) * (size_t)VL * KP);
    float* Yhr = (float*)malloc(sizeof(float) * (size_t)VL * KP);
    float* Yhi = (float*)malloc(sizeof(float) * (size_t)VL * KP);
    float* wre = (float*)malloc(sizeof(float) * (size_t)m * VL);  /* transform workspace [point][VL] */
    float* wim = (float*)malloc(sizeof(float) * (size_t)m * VL);
    float* fr = (float*)malloc(sizeof(float) * (size_t)N);
    float* prev_e = (float*)calloc((size_t)H, sizeof(float));   /* windowed second half of frame t-1 */
    float* prev_y = (float*)calloc((size_t)H, sizeof(float));
    float* yr = (float*)malloc(sizeof(float) * (size_t)KP);
    float* yi = (float*)malloc(sizeof(float) * (size_t)KP);
    float* pw = (float*)malloc(sizeof(float) * (size_t)KP);
    for (int i = 0; i < P * KP; ++i) C[i] = cfg->kalman_c0;
    const float A = cfg->kalman_a, A2 = A * A, Q = (float)(1.0 - (double)A * (double)A);
    const float lam = cfg->kalman_lambda, oml = 1.0f - lam;
    const float mu = cfg->mu, delta = cfg->delta, keps = cfg->kalman_eps;
    double pm = 0.0, pe = 0.0;

    for (int64_t t0 = 0; t0 < T; t0 += VL) {
        const int nf = (int)((T - t0) < VL ? (T - t0) : VL);
        /* ---- analysis of the chunk's frames (zero padded by H on the left, zeros beyond n on the right) ---- */
        for (int s = 0; s < 2; ++s) {
            const float* src = s == 0 ? far : mic;
            float* outr = s == 0 ? Xr : Yr;
            float* outi = s == 0 ? Xi : Yi;
            for (int f = 0; f < VL; ++f) {
                const int64_t t = t0 + f;
                for (int i = 0; i < m; ++i) {
                    const int64_t i0 = (t - 1) * H + 2 * i, i1 = i0 + 1;
                    wre[i * VL + f] = (f < nf && i0 >= 0 && i0 < n) ? src[i0] * p->win[2 * i] : 0.f;
                    wim[i * VL + f] = (f < nf && i1 >= 0 && i1 < n) ? src[i1] * p->win[2 * i + 1] : 0.f;
                }
            }
            fft_c_v(p, wre, wim, -1);
            for (int k = 0; k <= m; ++k) {                        /* rfft_frame's split, VL frames at once */
                const int ka = k & (m - 1), kb = (m - k) & (m - 1);
                const cpx w = p->tw_n[k];
                for (int f = 0; f < VL; ++f) {
                    const float ar = wre[ka * VL + f], ai = wim[ka * VL + f];
                    const float br = wre[kb * VL + f], bi = wim[kb * VL + f];
                    const float er = 0.5f * (ar + br), ei = 0.5f * (ai - bi);
                    const float orr = 0.5f * (ai + bi), oi = 0.5f * (br - ar);
                    outr[f * KP + k] = er + (orr * w.re - oi * w.im);
                    outi[f * KP + k] = (k == 0 || k == m) ? 0.f : ei + (orr * w.im + oi * w.re);
                }
            }
        }
        /* ---- recurrence: frames in order, every statement a loop over the bins ---- */
        for (int f = 0; f < nf; ++f) {
            const int64_t t = t0 + f;
            const float* restrict xr = Xr + (size_t)f * KP;
            const float* restrict xi = Xi + (size_t)f * KP;
            const float* restrict yyr = Yr + (size_t)f * KP;
            const float* restrict yyi = Yi + (size_t)f * KP;
            float* restrict er = Er + (size_t)f * KP;
            float* restrict ei = Ei + (size_t)f * KP;
            float* restrict hr0 = Hr + (size_t)(t % P) * KP;
            float* restrict hi0 = Hi + (size_t)(t % P) * KP;
            memcpy(hr0, xr, sizeof(float) * (size_t)K);
            memcpy(hi0, xi, sizeof(float) * (size_t)K);
#pragma omp simd
            for (int k = 0; k < K; ++k) { yr[k] = 0.f; yi[k] = 0.f; }
            for (int q = 0; q < P; ++q) {
                const float* restrict hr = Hr + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;   /* X[t-q], zero before start */
                const float* restrict hi = Hi + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;
                const float* restrict wr = Wr + (size_t)q * KP;
                const float* restrict wi = Wi + (size_t)q * KP;
#pragma omp simd
                for (int k = 0; k < K; ++k) {
                    yr[k] += wr[k] * hr[k] - wi[k] * hi[k];
                    yi[k] += wr[k] * hi[k] + wi[k] * hr[k];
                }
            }
#pragma omp simd
            for (int k = 0; k < K; ++k) { er[k] = yyr[k] - yr[k]; ei[k] = yyi[k] - yi[k]; }
            if (cfg->algo == 0) {
#pragma omp simd
                for (int k = 0; k < K; ++k) pw[k] = 0.f;
                for (int q = 0; q < P; ++q) {
                    const float* restrict hr = Hr + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;
                    const float* restrict hi = Hi + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;
#pragma omp simd
                    for (int k = 0; k < K; ++k) pw[k] += hr[k] * hr[k] + hi[k] * hi[k];
                }
#pragma omp simd
                for (int k = 0; k < K; ++k) pw[k] = mu / (pw[k] + delta);           /* g */
                for (int q = 0; q < P; ++q) {
                    const float* restrict hr = Hr + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;
                    const float* restrict hi = Hi + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;
                    float* restrict wr = Wr + (size_t)q * KP;
                    float* restrict wi = Wi + (size_t)q * KP;
#pragma omp simd
                    for (int k = 0; k < K; ++k) {
                        const float gr = pw[k] * er[k], gi = pw[k] * ei[k];
                        wr[k] += hr[k] * gr + hi[k] * gi;      /* conj(x) * g e */
                        wi[k] += hr[k] * gi - hi[k] * gr;
                    }
                }
            } else {
#pragma omp simd
                for (int k = 0; k < K; ++k) {
                    const float e2 = er[k] * er[k] + ei[k] * ei[k];
                    psi[k] = lam * psi[k] + oml * e2;
                    pw[k] = 0.f;                                                     /* d */
                }
                for (int q = 0; q < P; ++q) {
                    const float* restrict hr = Hr + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;
                    const float* restrict hi = Hi + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;
                    const float* restrict c = C + (size_t)q * KP;
                    float* restrict cx = cx2 + (size_t)q * KP;
#pragma omp simd
                    for (int k = 0; k < K; ++k) {
                        cx[k] = c[k] * (hr[k] * hr[k] + hi[k] * hi[k]);
                        pw[k] += cx[k];
                    }
                }
#pragma omp simd
                for (int k = 0; k < K; ++k) pw[k] = 1.0f / (pw[k] + psi[k] + keps);   /* rd */
                for (int q = 0; q < P; ++q) {
                    const float* restrict hr = Hr + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;
                    const float* restrict hi = Hi + (size_t)((t - q + 4 * (int64_t)P) % P) * KP;
                    float* restrict c = C + (size_t)q * KP;
                    const float* restrict cx = cx2 + (size_t)q * KP;
                    float* restrict wr = Wr + (size_t)q * KP;
                    float* restrict wi = Wi + (size_t)q * KP;
#pragma omp simd
                    for (int k = 0; k < K; ++k) {
                        const float gs = c[k] * pw[k];
                        const float gr = gs * hr[k], gi = -gs * hi[k];
                        const float nr = A * (wr[k] + gr * er[k] - gi * ei[k]);
                        const float ni = A * (wi[k] + gr * ei[k] + gi * er[k]);
                        wr[k] = nr; wi[k] = ni;
                        c[k] = A2 * (1.0f - cx[k] * pw[k]) * c[k] + Q * (nr * nr + ni * ni);
                    }
                }
            }
            memcpy(Yhr + (size_t)f * KP, yr, sizeof(float) * (size_t)K);
            memcpy(Yhi + (size_t)f * KP, yi, sizeof(float) * (size_t)K);
        }
        /* ---- synthesis of the chunk + overlap-add: output hop t-1 = second half of frame t-1 + first half of t ---- */
        for (int s = 0; s < (echo ? 2 : 1); ++s) {
            const float* inr = s == 0 ? Er : Yhr;
            const float* ini = s == 0 ? Ei : Yhi;
            float* prev = s == 0 ? prev_e : prev_y;
            float* dst = s == 0 ? err : echo;
            for (int k = 0; k < m; ++k) {                         /* irfft_frame's packing, VL frames at once */
                const cpx w = p->tw_n[k];
                for (int f = 0; f < VL; ++f) {
                    float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;
                    if (f < nf) {
                        ar = inr[f * KP + k]; ai = ini[f * KP + k];
                        br = inr[f * KP + m - k]; bi = ini[f * KP + m - k];
                        if (k == 0) { ai = 0.f; bi = 0.f; }
                    }
                    const float e_r = ar + br, e_i = ai - bi;
                    const float dr = ar - br, di = ai + bi;
                    const float tr = dr * w.re + di * w.im, ti = di * w.re - dr * w.im;
                    wre[k * VL + f] = e_r - ti;
                    wim[k * VL + f] = e_i + tr;
                }
            }
            fft_c_v(p, wre, wim, +1);
            const float sc = 1.0f / (float)p->n;
            for (int f = 0; f < nf; ++f) {
                const int64_t t = t0 + f;
                for (int i = 0; i < m; ++i) {
                    fr[2 * i] = wre[i * VL + f] * sc * p->win[2 * i];
                    fr[2 * i + 1] = wim[i * VL + f] * sc * p->win[2 * i + 1];
                }
                if (t >= 1) {
                    for (int j = 0; j < H; ++j) {
                        const float o = (prev[j] + fr[j]) * p->norm[j];
                        dst[(t - 1) * H + j] = o;
                        if (s == 0 && t - 1 >= cfg->erle_skip_hops) {
                            const float mv = mic[(t - 1) * H + j];
                            pe += (double)o * o;
                            pm += (double)mv * mv;
                        }
                    }
                }
                memcpy(prev, fr + H, sizeof(float) * (size_t)H);
            }
        }
    }
    for (int64_t i = valid; i < out_len_total; ++i) {
        err[i] = 0.f;
        if (echo) echo[i] = 0.f;
    }
    if (erle_db) {
        if (pm < 1e-20) pm = 1e-20;
        if (pe < 1e-20) pe = 1e-20;
        *erle_db = (float)(10.0 * log10(pm / pe));
    }
    free(Wr); free(Wi); free(Hr); free(Hi); free(C); free(cx2); free(psi); free(Yr); free(Yi); free(Xr); free(Xi);
    free(Er); free(Ei); free(Yhr); free(Yhi); free(wre); free(wim); free(fr); free(prev_e); free(prev_y);
    free(yr); free(yi); free(pw);
}

/* ---- overlap-save partitioned-block FDAF (algos 2 / 3), VL utterances at a time ----------------------------------
 * Follows oracle/aec_oracle.py:pbfdaf_ols statement by statement (BUILDER-AUTHORED, PARITY UNPINNED like the other
 * recurrences).  The recurrence crosses the transform every block, so there are no frames of one
 * utterance to batch: the SIMD lanes are VL different utterances instead (layout [bin][VL] / [point][VL]), every
 * transform and every statement of the update runs on all of them at once.  Lanes whose utterance is shorter (or absent)
 * keep computing on zeros; nothing of theirs is stored. */
static void rfft_lanes(const plan* p, float* restrict wre, float* restrict wim, float* restrict outr, float* restrict outi) {
    const int m = p->m;
    fft_c_v(p, wre, wim, -1);
    for (int k = 0; k <= m; ++k) {
        const int ka = k & (m - 1), kb = (m - k) & (m - 1);
        const cpx w = p->tw_n[k];
        const int edge = (k == 0 || k == m);
#pragma omp simd
        for (int l = 0; l < VL; ++l) {
            const float ar = wre[ka * VL + l], ai = wim[ka * VL + l];
            const float br = wre[kb * VL + l], bi = wim[kb * VL + l];
            const float er = 0.5f * (ar + br), ei = 0.5f * (ai - bi);
            const float orr = 0.5f * (ai + bi), oi = 0.5f * (br - ar);
            outr[k * VL + l] = er + (orr * w.re - oi * w.im);
            outi[k * VL + l] = edge ? 0.f : ei + (orr * w.im + oi * w.re);
        }
    }
}

/* inverse of rfft_lanes: time sample 2i in wre[i][.], 2i+1 in wim[i][.], scaled by 1/n */
static void irfft_lanes(const plan* p, const float* restrict inr, const float* restrict ini, float* restrict wre,
                        float* restrict wim) {
    const int m = p->m;
    const float sc = 1.0f / (float)p->n;
    for (int k = 0; k < m; ++k) {
        const cpx w = p->tw_n[k];
        const float z = k == 0 ? 0.f : 1.f;          /* imaginary parts of DC / Nyquist do not exist */
#pragma omp simd
        for (int l = 0; l < VL; ++l) {
            const float ar = inr[k * VL + l], ai = z * ini[k * VL + l];
            const float br = inr[(m - k) * VL + l], bi = z * ini[(m - k) * VL + l];
            const float e_r = ar + br, e_i = ai - bi;
            const float dr = ar - br, di = ai + bi;
            const float tr = dr * w.re + di * w.im, ti = di * w.re - dr * w.im;
            wre[k * VL + l] = (e_r - ti) * sc;
            wim[k * VL + l] = (e_i + tr) * sc;
        }
    }
    fft_c_v(p, wre, wim, +1);
}

static void run_group_ols(const plan* p, const aec_oracle_cfg* cfg, const float* const* far, const float* const* mic,
                          const int64_t* n, int64_t out_len_total, float* const* err, float* const* echo,
                          float* const* erle_db) {
    const int N = p->n, H = N / 2, K = H + 1, P = cfg->partitions, m = H;
    const size_t row = (size_t)K * VL;
    int64_t nblk[VL], nmax = 0;
    for (int l = 0; l < VL; ++l) {
        nblk[l] = far[l] ? n[l] / H : 0;
        if (nblk[l] > nmax) nmax = nblk[l];
    }
    float* Wr = (float*)calloc((size_t)P * row, sizeof(float));
    float* Wi = (float*)calloc((size_t)P * row, sizeof(float));
    float* Xr = (float*)calloc((size_t)P * row, sizeof(float));     /* far-end ring: slot t mod P holds Xh of block t */
    float* Xi = (float*)calloc((size_t)P * row, sizeof(float));
    float* C = (float*)malloc(sizeof(float) * (size_t)P * row);
    float* cx2 = (float*)malloc(sizeof(float) * (size_t)P * row);
    float* pw = (float*)calloc(row, sizeof(float));
    float* psi = (float*)calloc(row, sizeof(float));
    float* rd = (float*)malloc(sizeof(float) * row);
    float* Sr = (float*)malloc(sizeof(float) * row);                 /* Yhat, then E */
    float* Si = (float*)malloc(sizeof(float) * row);
    float* wre = (float*)malloc(sizeof(float) * (size_t)m * VL);
    float* wim = (float*)malloc(sizeof(float) * (size_t)m * VL);
    for (size_t i = 0; i < (size_t)P * row; ++i) C[i] = cfg->kalman_c0;
    const float A = cfg->kalman_a, A2 = A * A, Q = (float)(1.0 - (double)A * (double)A);
    const float klam = cfg->kalman_lambda, koml = 1.0f - klam, keps = cfg->kalman_eps;
    const float mu = cfg->mu, delta = cfg->delta, lam = cfg->pb_lambda, oml = 1.0f - lam;
    const int half = H / 2;                                          /* complex points per half of the 2H-sample vector */
    double pm[VL] = {0}, pe[VL] = {0};

    for (int64_t t = 0; t < nmax; ++t) {
        /* Xh_0 = rfft([x_{t-1}, x_t]) into ring slot t mod P (the shift of the partitions is the ring index) */
        for (int l = 0; l < VL; ++l) {
            const int live = t < nblk[l];
            const float* x = far[l];
            for (int i = 0; i < half; ++i) {
                wre[i * VL + l] = (live && t > 0) ? x[(t - 1) * H + 2 * i] : 0.f;
                wim[i * VL + l] = (live && t > 0) ? x[(t - 1) * H + 2 * i + 1] : 0.f;
                wre[(half + i) * VL + l] = live ? x[t * H + 2 * i] : 0.f;
                wim[(half + i) * VL + l] = live ? x[t * H + 2 * i + 1] : 0.f;
            }
        }
        rfft_lanes(p, wre, wim, Xr + (size_t)(t % P) * row, Xi + (size_t)(t % P) * row);
        /* y = irfft(sum_p W_p Xh_p)[H:] */
        memset(Sr, 0, sizeof(float) * row);
        memset(Si, 0, sizeof(float) * row);
        for (int q = 0; q < P; ++q) {
            const size_t slot = (size_t)((t - q + 4 * (int64_t)P) % P) * row;
            const float* restrict xr = Xr + slot;
            const float* restrict xi = Xi + slot;
            const float* restrict wr = Wr + (size_t)q * row;
            const float* restrict wi = Wi + (size_t)q * row;
#pragma omp simd
            for (size_t j = 0; j < row; ++j) {
                Sr[j] += wr[j] * xr[j] - wi[j] * xi[j];
                Si[j] += wr[j] * xi[j] + wi[j] * xr[j];
            }
        }
        irfft_lanes(p, Sr, Si, wre, wim);
        /* e = d_t - y -> err; E = rfft([0_H, e]) */
        for (int l = 0; l < VL; ++l) {
            const int live = t < nblk[l];
            for (int i = 0; i < half; ++i) {
                float e0 = 0.f, e1 = 0.f;
                if (live) {
                    const int64_t s = t * H + 2 * i;
                    const float y0 = wre[(half + i) * VL + l], y1 = wim[(half + i) * VL + l];
                    const float d0 = mic[l][s], d1 = mic[l][s + 1];
                    e0 = d0 - y0; e1 = d1 - y1;
                    err[l][s] = e0; err[l][s + 1] = e1;
                    if (echo && echo[l]) { echo[l][s] = y0; echo[l][s + 1] = y1; }
                    if (t >= cfg->erle_skip_hops) {
                        pe[l] += (double)e0 * e0 + (double)e1 * e1;
                        pm[l] += (double)d0 * d0 + (double)d1 * d1;
                    }
                }
                wre[i * VL + l] = 0.f; wim[i * VL + l] = 0.f;
                wre[(half + i) * VL + l] = e0; wim[(half + i) * VL + l] = e1;
            }
        }
        rfft_lanes(p, wre, wim, Sr, Si);                             /* E */
        /* gradient constraint of partition c = t mod P, as it entered the block */
        {
            float* wr = Wr + (size_t)(t % P) * row;
            float* wi = Wi + (size_t)(t % P) * row;
            irfft_lanes(p, wr, wi, wre, wim);
            memset(wre + (size_t)half * VL, 0, sizeof(float) * (size_t)half * VL);
            memset(wim + (size_t)half * VL, 0, sizeof(float) * (size_t)half * VL);
            rfft_lanes(p, wre, wim, wr, wi);
        }
        if (cfg->algo == 2) {
            memset(rd, 0, sizeof(float) * row);
            for (int q = 0; q < P; ++q) {                            /* sum_p |Xh_p|^2, newest partition first */
                const size_t slot = (size_t)((t - q + 4 * (int64_t)P) % P) * row;
                const float* restrict xr = Xr + slot;
                const float* restrict xi = Xi + slot;
#pragma omp simd
                for (size_t j = 0; j < row; ++j) rd[j] += xr[j] * xr[j] + xi[j] * xi[j];
            }
#pragma omp simd
            for (size_t j = 0; j < row; ++j) {
                pw[j] = lam * pw[j] + oml * rd[j];
                rd[j] = mu / (pw[j] + delta);                        /* g */
            }
            for (int q = 0; q < P; ++q) {
                const size_t slot = (size_t)((t - q + 4 * (int64_t)P) % P) * row;
                const float* restrict xr = Xr + slot;
                const float* restrict xi = Xi + slot;
                float* restrict wr = Wr + (size_t)q * row;
                float* restrict wi = Wi + (size_t)q * row;
#pragma omp simd
                for (size_t j = 0; j < row; ++j) {
                    const float gr = rd[j] * Sr[j], gi = rd[j] * Si[j];
                    wr[j] += xr[j] * gr + xi[j] * gi;                /* conj(Xh_p) g E */
                    wi[j] += xr[j] * gi - xi[j] * gr;
                }
            }
        } else {
#pragma omp simd
            for (size_t j = 0; j < row; ++j) {
                psi[j] = klam * psi[j] + koml * (Sr[j] * Sr[j] + Si[j] * Si[j]);
                rd[j] = 0.f;
            }
            for (int q = 0; q < P; ++q) {
                const size_t slot = (size_t)((t - q + 4 * (int64_t)P) % P) * row;
                const float* restrict xr = Xr + slot;
                const float* restrict xi = Xi + slot;
                const float* restrict c = C + (size_t)q * row;
                float* restrict cx = cx2 + (size_t)q * row;
#pragma omp simd
                for (size_t j = 0; j < row; ++j) {
                    cx[j] = c[j] * (xr[j] * xr[j] + xi[j] * xi[j]);
                    rd[j] += cx[j];
                }
            }
#pragma omp simd
            for (size_t j = 0; j < row; ++j) rd[j] = 1.0f / (rd[j] + psi[j] + keps);
            for (int q = 0; q < P; ++q) {
                const size_t slot = (size_t)((t - q + 4 * (int64_t)P) % P) * row;
                const float* restrict xr = Xr + slot;
                const float* restrict xi = Xi + slot;
                float* restrict c = C + (size_t)q * row;
                const float* restrict cx = cx2 + (size_t)q * row;
                float* restrict wr = Wr + (size_t)q * row;
                float* restrict wi = Wi + (size_t)q * row;
#pragma omp simd
                for (size_t j = 0; j < row; ++j) {
                    const float gs = c[j] * rd[j];
                    const float gr = gs * xr[j], gi = -gs * xi[j];
                    const float nr = A * (wr[j] + gr * Sr[j] - gi * Si[j]);
                    const float ni = A * (wi[j] + gr * Si[j] + gi * Sr[j]);
                    wr[j] = nr; wi[j] = ni;
                    c[j] = A2 * (1.0f - cx[j] * rd[j]) * c[j] + Q * (nr * nr + ni * ni);
                }
            }
        }
    }
    for (int l = 0; l < VL; ++l) {
        if (!far[l]) continue;
        for (int64_t i = nblk[l] * H; i < out_len_total; ++i) {
            err[l][i] = 0.f;
            if (echo && echo[l]) echo[l][i] = 0.f;
        }
        if (erle_db && erle_db[l]) {
            double a = pm[l] < 1e-20 ? 1e-20 : pm[l], b = pe[l] < 1e-20 ? 1e-20 : pe[l];
            *erle_db[l] = (float)(10.0 * log10(a / b));
        }
    }
    free(Wr); free(Wi); free(Xr); free(Xi); free(C); free(cx2); free(pw); free(psi); free(rd); free(Sr); free(Si);
    free(wre); free(wim);
}

/* Batch entry.  Buffers are host float32; n_samples nullable.  n_threads <= 0 -> all cores.
 * Returns the number of threads used (>= 1) or a negative error. */
int aec_oracle_stage1_f32(const float* far, const float* mic, float* err, float* echo, float* erle_db,
                          const int64_t* n_samples, int64_t B, int64_t L, int64_t in_stride, int64_t out_stride,
                          const aec_oracle_cfg* cfg, int n_threads) {
    if (!far || !mic || !err || !cfg || B < 0 || L < 0) return -1;
    if (cfg->frame != 512 && cfg->frame != 1024) return -2;
    if (cfg->partitions < 1) return -1;
    if (cfg->algo < 0 || cfg->algo > 3) return -2;
    plan* p = plan_create(cfg->frame);
    int used = 1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    used = n_threads;
#endif
    if (cfg->algo >= 2) {                 /* overlap-save filters: groups of VL utterances, one group per task */
        const int64_t groups = (B + VL - 1) / VL;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
        for (int64_t g = 0; g < groups; ++g) {
            const float *f[VL], *d[VL];
            float *e[VL], *y[VL], *r[VL];
            int64_t nn[VL];
            for (int l = 0; l < VL; ++l) {
                const int64_t b = g * VL + l;
                const int on = b < B;
                int64_t n = on ? (n_samples ? n_samples[b] : L) : 0;
                if (n < 0) n = 0;
                if (n > L) n = L;
                nn[l] = n;
                f[l] = on ? far + b * in_stride : NULL;
                d[l] = on ? mic + b * in_stride : NULL;
                e[l] = on ? err + b * out_stride : NULL;
                y[l] = (on && echo) ? echo + b * out_stride : NULL;
                r[l] = (on && erle_db) ? erle_db + b : NULL;
            }
            run_group_ols(p, cfg, f, d, nn, L, e, echo ? y : NULL, erle_db ? r : NULL);
        }
        plan_destroy(p);
        return used;
    }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
    for (int64_t b = 0; b < B; ++b) {
        int64_t n = n_samples ? n_samples[b] : L;
        if (n < 0) n = 0;
        if (n > L) n = L;
        run_one(p, cfg, far + b * in_stride, mic + b * in_stride, n, L, err + b * out_stride,
                echo ? echo + b * out_stride : NULL, erle_db ? erle_db + b : NULL);
    }
    plan_destroy(p);
    return used;
}

int aec_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
