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
 * statement by statement.
 *
 * The reference computes each DFT as a dense [2K x N] convolution; this port uses an FFT
 * (same result to float32 rounding) so that the CPU baseline is not handicapped.
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
    int32_t algo; /* 0 NLMS, 1 Kalman */
    float mu, delta;
    float kalman_a, kalman_lambda, kalman_c0, kalman_eps;
    int32_t erle_skip_hops;
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

/* in-place radix-2 DIT complex FFT of length m; sign = -1 forward, +1 inverse (unscaled) */
static void fft_c(const plan* p, cpx* a, int sign) {
    const int m = p->m;
    for (int i = 0; i < m; ++i) {
        const int r = p->rev[i];
        if (r > i) { cpx t = a[i]; a[i] = a[r]; a[r] = t; }
    }
    for (int len = 2; len <= m; len <<= 1) {
        const int half = len >> 1, step = m / len;
        for (int s = 0; s < m; s += len) {
            for (int j = 0; j < half; ++j) {
                const cpx w = p->tw[j * step];
                const float wi = sign < 0 ? w.im : -w.im;
                const cpx u = a[s + j], v = a[s + j + half];
                const float tr = v.re * w.re - v.im * wi;
                const float ti = v.re * wi + v.im * w.re;
                a[s + j].re = u.re + tr; a[s + j].im = u.im + ti;
                a[s + j + half].re = u.re - tr; a[s + j + half].im = u.im - ti;
            }
        }
    }
}

/* windowed real frame (n samples) -> K = n/2+1 bins, rfft convention (attention_ccrn.py:15-18) */
static void rfft_frame(const plan* p, const float* x, cpx* out, cpx* work) {
    const int m = p->m;
    for (int i = 0; i < m; ++i) { work[i].re = x[2 * i]; work[i].im = x[2 * i + 1]; }
    fft_c(p, work, -1);
    for (int k = 0; k <= m; ++k) {
        const cpx a = work[k & (m - 1)], b = work[(m - k) & (m - 1)];
        const float er = 0.5f * (a.re + b.re), ei = 0.5f * (a.im - b.im);   /* even part */
        const float orr = 0.5f * (a.im + b.im), oi = 0.5f * (b.re - a.re);  /* odd part  */
        const cpx w = p->tw_n[k];
        out[k].re = er + (orr * w.re - oi * w.im);
        out[k].im = ei + (orr * w.im + oi * w.re);
    }
    out[0].im = 0.f;
    out[m].im = 0.f;
}

/* K bins -> n real samples (irfft; imaginary parts of DC / Nyquist ignored) */
static void irfft_frame(const plan* p, const cpx* in, float* x, cpx* work) {
    const int m = p->m;
    for (int k = 0; k < m; ++k) {
        cpx a = in[k], b = in[m - k];
        if (k == 0) { a.im = 0.f; b.im = 0.f; }
        const float er = a.re + b.re, ei = a.im - b.im;           /* E[k] + conj E[m-k] */
        const float dr = a.re - b.re, di = a.im + b.im;           /* E[k] - conj E[m-k] */
        const cpx w = p->tw_n[k];                                 /* multiply by conj(w) */
        const float tr = dr * w.re + di * w.im, ti = di * w.re - dr * w.im;
        work[k].re = er - ti;
        work[k].im = ei + tr;
    }
    fft_c(p, work, +1);
    const float s = 1.0f / (float)p->n;
    for (int i = 0; i < m; ++i) { x[2 * i] = work[i].re * s; x[2 * i + 1] = work[i].im * s; }
}

/* one utterance */
static void run_one(const plan* p, const aec_oracle_cfg* cfg, const float* far, const float* mic, int64_t n,
                    int64_t out_len_total, float* err, float* echo, float* erle_db) {
    const int N = p->n, H = N / 2, K = H + 1, P = cfg->partitions;
    const int64_t T = n / H + 1;                     /* attention_ccrn.py:48-49 with N = 2H */
    const int64_t valid = (T - 1) * H;
    cpx* W = (cpx*)calloc((size_t)P * K, sizeof(cpx));
    cpx* hist = (cpx*)calloc((size_t)P * K, sizeof(cpx));   /* ring: slot (t - p) mod P */
    float* C = (float*)malloc(sizeof(float) * (size_t)P * K);
    float* psi = (float*)calloc((size_t)K, sizeof(float));
    cpx* X = (cpx*)malloc(sizeof(cpx) * (size_t)K);
    cpx* Y = (cpx*)malloc(sizeof(cpx) * (size_t)K);
    cpx* E = (cpx*)malloc(sizeof(cpx) * (size_t)K);
    cpx* Yh = (cpx*)malloc(sizeof(cpx) * (size_t)K);
    cpx* work = (cpx*)malloc(sizeof(cpx) * (size_t)H);
    float* fr = (float*)malloc(sizeof(float) * (size_t)N);
    float* prev_e = (float*)calloc((size_t)H, sizeof(float));   /* windowed second half of frame t-1 */
    float* prev_y = (float*)calloc((size_t)H, sizeof(float));
    float* cx2 = (float*)malloc(sizeof(float) * (size_t)P);
    for (int i = 0; i < P * K; ++i) C[i] = cfg->kalman_c0;
    const float A = cfg->kalman_a, A2 = A * A, Q = (float)(1.0 - (double)A * (double)A);
    const float lam = cfg->kalman_lambda, oml = 1.0f - lam;
    double pm = 0.0, pe = 0.0;

    for (int64_t t = 0; t < T; ++t) {
        /* ---- analysis (zero padded by H on the left, zeros beyond n on the right) ---- */
        for (int s = 0; s < 2; ++s) {
            const float* src = s == 0 ? far : mic;
            for (int i = 0; i < N; ++i) {
                const int64_t idx = (t - 1) * H + i;
                fr[i] = (idx >= 0 && idx < n) ? src[idx] * p->win[i] : 0.f;
            }
            rfft_frame(p, fr, s == 0 ? X : Y, work);
        }
        cpx* slot = hist + (size_t)(t % P) * K;
        memcpy(slot, X, sizeof(cpx) * (size_t)K);
        /* ---- recurrence, per bin ---- */
        for (int k = 0; k < K; ++k) {
            float yr = 0.f, yi = 0.f;
            for (int q = 0; q < P; ++q) {
                const cpx x = hist[(size_t)((t - q + 4 * (int64_t)P) % P) * K + k];   /* X[t-q], zero before start */
                const cpx w = W[(size_t)q * K + k];
                yr += w.re * x.re - w.im * x.im;
                yi += w.re * x.im + w.im * x.re;
            }
            const float er = Y[k].re - yr, ei = Y[k].im - yi;
            if (cfg->algo == 0) {
                float pw = 0.f;
                for (int q = 0; q < P; ++q) {
                    const cpx x = hist[(size_t)((t - q + 4 * (int64_t)P) % P) * K + k];
                    pw += x.re * x.re + x.im * x.im;
                }
                const float g = cfg->mu / (pw + cfg->delta);
                const float gr = g * er, gi = g * ei;
                for (int q = 0; q < P; ++q) {
                    const cpx x = hist[(size_t)((t - q + 4 * (int64_t)P) % P) * K + k];
                    cpx* w = &W[(size_t)q * K + k];
                    w->re += x.re * gr + x.im * gi;      /* conj(x) * g e */
                    w->im += x.re * gi - x.im * gr;
                }
            } else {
                const float e2 = er * er + ei * ei;
                psi[k] = lam * psi[k] + oml * e2;
                float d = 0.f;
                for (int q = 0; q < P; ++q) {
                    const cpx x = hist[(size_t)((t - q + 4 * (int64_t)P) % P) * K + k];
                    cx2[q] = C[(size_t)q * K + k] * (x.re * x.re + x.im * x.im);
                    d += cx2[q];
                }
                d = d + psi[k] + cfg->kalman_eps;
                const float rd = 1.0f / d;
                for (int q = 0; q < P; ++q) {
                    const cpx x = hist[(size_t)((t - q + 4 * (int64_t)P) % P) * K + k];
                    float* c = &C[(size_t)q * K + k];
                    cpx* w = &W[(size_t)q * K + k];
                    const float gs = *c * rd;
                    const float gr = gs * x.re, gi = -gs * x.im;
                    const float wr = A * (w->re + gr * er - gi * ei);
                    const float wi = A * (w->im + gr * ei + gi * er);
                    w->re = wr; w->im = wi;
                    *c = A2 * (1.0f - cx2[q] * rd) * *c + Q * (wr * wr + wi * wi);
                }
            }
            E[k].re = er; E[k].im = ei;
            Yh[k].re = yr; Yh[k].im = yi;
        }
        /* ---- synthesis + overlap-add: output hop t-1 = second half of frame t-1 + first half of t ---- */
        for (int s = 0; s < (echo ? 2 : 1); ++s) {
            float* prev = s == 0 ? prev_e : prev_y;
            float* dst = s == 0 ? err : echo;
            irfft_frame(p, s == 0 ? E : Yh, fr, work);
            for (int i = 0; i < N; ++i) fr[i] *= p->win[i];
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
    for (int64_t i = valid; i < out_len_total; ++i) {
        err[i] = 0.f;
        if (echo) echo[i] = 0.f;
    }
    if (erle_db) {
        if (pm < 1e-20) pm = 1e-20;
        if (pe < 1e-20) pe = 1e-20;
        *erle_db = (float)(10.0 * log10(pm / pe));
    }
    free(W); free(hist); free(C); free(psi); free(X); free(Y); free(E); free(Yh); free(work); free(fr);
    free(prev_e); free(prev_y); free(cx2);
}

/* Batch entry.  Buffers are host float32; n_samples nullable.  n_threads <= 0 -> all cores.
 * Returns the number of threads used (>= 1) or a negative error. */
int aec_oracle_stage1_f32(const float* far, const float* mic, float* err, float* echo, float* erle_db,
                          const int64_t* n_samples, int64_t B, int64_t L, int64_t in_stride, int64_t out_stride,
                          const aec_oracle_cfg* cfg, int n_threads) {
    if (!far || !mic || !err || !cfg || B < 0 || L < 0) return -1;
    if (cfg->frame != 512 && cfg->frame != 1024) return -2;
    if (cfg->partitions < 1) return -1;
    plan* p = plan_create(cfg->frame);
    int used = 1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    used = n_threads;
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
