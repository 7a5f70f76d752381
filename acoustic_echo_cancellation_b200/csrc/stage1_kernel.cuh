// Stage-1 linear echo canceller: fused STFT -> partitioned FDAF -> iSTFT, one persistent
// CTA per utterance, N = 512 (16 kHz live configuration of the reference:
// Stage2_lhm/scripts/configs.py:1-8, network/ERB.py:223-224).
//
// The reference has no stage-1 filter; the STFT/iSTFT conventions follow its operators
// (Stage2_lhm/scripts/network/attention_ccrn.py:45-52 and :82-101) and the recurrence is the
// builder-authored one frozen in DESIGN.md and restated by oracle/aec_oracle.py.
//
// Per chunk of F = 2*NW frames (NW = warps per CTA):
//   phase A  analysis : each warp transforms one frame at a time; lanes 0-15 far-end, 16-31
//                       microphone (two half-warp FFT-256 of the even/odd-packed real frame).
//   phase B  filter   : every thread owns mirrored bin pairs (k, 256-k); unpacks X,Y from the
//                       half-size spectra, runs the NLMS / Kalman recurrence with the filter
//                       taps, far-end history and covariances resident in REGISTERS for the
//                       whole utterance, and packs E (and Yhat) for the inverse transform
//                       in place.
//   phase C  synthesis: each warp inverts two consecutive frames (one per half-warp), applies
//                       the synthesis window / overlap-add normaliser, adds the in-warp
//                       overlap with one shuffle and the cross-warp overlap through 1 KB
//                       tails in shared memory, and streams the error signal to HBM.
// Far-end / microphone hops are staged HBM -> shared memory by the bulk-copy engine (TMA,
// cp.async.bulk + mbarrier) one chunk ahead of the analysis phase -- one or two copies per signal
// per chunk, issued by an elected lane; spectra never touch HBM.
// Bin 128 (its own mirror, the 257th bin on 256 bin slots) is a serial pass: on one lane inside the
// frame loop for short filters (owner warp alternating with the hardware warp slot), and one chunk
// AHEAD on the warps without synthesis work for the 16-partition ring kernels, where X[128], Y[128]
// are evaluated directly from the staged samples ((-i)^n sums, no FFT).  Measurements and the
// reasons for every one of these choices: DESIGN.md section 5, profiles/.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "fft_warp.cuh"

namespace aec {

enum { kAlgoNlms = 0, kAlgoKalman = 1, kAlgoPbfdaf = 2, kAlgoPbfkf = 3 };

// "this (P, algo, echo, register cap) is not instantiated": a code no launch path of the runtime produces
constexpr cudaError_t kNoInstance = cudaErrorStubLibrary;

struct Stage1Params {
    const float* far;
    const float* mic;
    float* err;
    float* echo;       // nullable (only read by ECHO instantiations)
    float* erle_db;    // nullable
    const long long* n_samples;  // nullable -> every utterance has L samples
    long long B, L, in_stride, out_stride;
    float mu, delta;                         // NLMS
    float ka, ka2, kq, klam, koml, kc0, keps;  // Kalman: A, A^2, 1-A^2, lambda, 1-lambda, c0, eps
    float pblam, pboml;                        // overlap-save PBFDAF (algo 2): power smoothing lambda, 1 - lambda
    int erle_skip_hops;
    int use_tma;      // inputs 16-byte aligned per hop -> bulk copies
    int vec_out;      // outputs 8-byte aligned -> float2 stores
    int stagger_ns;   // start-up skew between utterances sharing an SM (de-synchronises the phases)
    int num_sms;
    const float2* tw256;   // [16][16]  exp(-2 pi i h q / 256) at [q*16 + h]
    const float2* tw512;   // [129]     exp(-2 pi i k / 512)
    const float2* win_a;   // [256]     0.5 * hann[2m], 0.5 * hann[2m+1]
    const float2* win_s;   // [256]     hann[n] / (512 * (coff[n] + 1e-8)), n = 2m, 2m+1
    // fused Stage-2 feature epilogue (FEAT instantiations only; Stage2_lhm/scripts/network/ERB.py:262-290, in_norm off)
    float* feat;           // [B][feat_frames][64]  cat[err_erb, |err_erb - far_erb|]
    const float* erb;      // [257][32] dense cosine bank (ERB.py:10-71), device memory
    long long feat_frames; // rows per utterance (frames of an L-sample signal)
#ifdef AEC_PHASE_TIMING
    long long* dbg;        // developer build only: [B][NW][12] cycles per phase (tools/phase_timing.py)
#endif
};

// Developer instrumentation (compiled out of the product library): lane 0 of every warp accumulates
// the cycles between consecutive AEC_TICK points in shared memory.
#ifdef AEC_PHASE_TIMING
#define AEC_TICK(i)                                                   \
    do {                                                              \
        if (lane == 0) {                                              \
            const long long now_ = clock64();                         \
            dbg_sm[warp][i] += now_ - dbg_sm[warp][11];               \
            dbg_sm[warp][11] = now_;                                  \
        }                                                             \
    } while (0)
#else
#define AEC_TICK(i) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "AEC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra AEC_DONE;\n"
        "bra AEC_WAIT;\n"
        "AEC_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// one lane of a converged warp (SASS: ELECT); keeps the bulk-copy operands in uniform registers
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void st_stream_f2(float* p, float2 v) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream_f1(float* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// predicated forms: a store that only part of a warp performs, without a divergent branch around it (eight
// BSSY / BRA / BSYNC regions in the overlap-save kernels' transform chain cost ~40 cycles of branch resolution each)
__device__ __forceinline__ void st_stream_f2_if(float* p, float2 v, bool on) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %3, 0;\n@q st.global.cs.v2.f32 [%0], {%1, %2};\n}" ::"l"(p), "f"(v.x), "f"(v.y),
                 "r"(static_cast<unsigned>(on))
                 : "memory");
}
__device__ __forceinline__ void st_stream_f1_if(float* p, float v, bool on) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.global.cs.f32 [%0], %1;\n}" ::"l"(p), "f"(v),
                 "r"(static_cast<unsigned>(on))
                 : "memory");
}

// MUFU.RCP without the range fix-up code of __fdividef / the Newton step of 1.0f / x (the argument
// is a regularised power, far from the denormal / overflow ranges; 1 ulp is ample for 1e-4 parity)
// MUFU.SQRT (1 ulp) for the magnitudes of the fused feature epilogue (arguments >= 1e-9; same form as aec_features)
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------------------------------
// per-bin recurrence state, resident in registers
// ------------------------------------------------------------------------------------------
template <int P, int ALGO>
struct BinState {
    float2 W[P];   // filter taps
    float2 X[P];   // X[p] = far-end spectrum p hops ago (X[0] = current after the shift)
    float C[ALGO == kAlgoKalman ? P : 1];
    float psi;
    __device__ __forceinline__ void init(float c0) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            W[p] = make_float2(0.f, 0.f);
            X[p] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int p = 0; p < (ALGO == kAlgoKalman ? P : 1); ++p) C[p] = c0;
        psi = 0.f;
    }
    static constexpr int kFloats = 4 * P + (ALGO == kAlgoKalman ? P : 1) + 1;
    __device__ __forceinline__ void store(float* m) const {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            m[4 * p + 0] = W[p].x; m[4 * p + 1] = W[p].y; m[4 * p + 2] = X[p].x; m[4 * p + 3] = X[p].y;
        }
#pragma unroll
        for (int p = 0; p < (ALGO == kAlgoKalman ? P : 1); ++p) m[4 * P + p] = C[p];
        m[kFloats - 1] = psi;
    }
    __device__ __forceinline__ void load(const float* m) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            W[p] = make_float2(m[4 * p + 0], m[4 * p + 1]);
            X[p] = make_float2(m[4 * p + 2], m[4 * p + 3]);
        }
#pragma unroll
        for (int p = 0; p < (ALGO == kAlgoKalman ? P : 1); ++p) C[p] = m[4 * P + p];
        psi = m[kFloats - 1];
    }
};

// One frame of the recurrence for one bin.  Operation order mirrors oracle/aec_oracle.py
// (fdaf_nlms / fdaf_kalman).
// SHIFT = false: the caller has already placed X[t-1..t-P+1] in s.X[1..P-1] (history kept in a
// shared-memory ring for long filters, so that it is not live in registers across the FFT phases).
template <int P, int ALGO, bool SHIFT = true>
__device__ __forceinline__ void bin_step(BinState<P, ALGO>& s, const float2 Xn, const float2 Y,
                                         const Stage1Params& prm, float2& E, float2& Yh) {
    if constexpr (SHIFT) {
#pragma unroll
        for (int p = P - 1; p > 0; --p) s.X[p] = s.X[p - 1];
    }
    s.X[0] = Xn;
    float2 yh = make_float2(0.f, 0.f);
#pragma unroll
    for (int p = 0; p < P; ++p) yh = cfma(s.W[p], s.X[p], yh);
    const float2 e = csub(Y, yh);
    if constexpr (ALGO == kAlgoNlms) {
        float pw = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) pw = fmaf(s.X[p].x, s.X[p].x, fmaf(s.X[p].y, s.X[p].y, pw));
        const float g = prm.mu * rcp_fast(pw + prm.delta);
        const float2 ge = make_float2(g * e.x, g * e.y);
#pragma unroll
        for (int p = 0; p < P; ++p) s.W[p] = cfmac(s.X[p], ge, s.W[p]);
    } else {
        const float e2 = fmaf(e.x, e.x, e.y * e.y);
        s.psi = fmaf(prm.klam, s.psi, prm.koml * e2);
        // |X|^2 is recomputed in the update loop rather than kept in a P-entry array: two more
        // instructions per tap, P fewer live registers (what decides spilling at P = 16)
        float d = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float x2 = fmaf(s.X[p].x, s.X[p].x, s.X[p].y * s.X[p].y);
            d = fmaf(s.C[p], x2, d);
        }
        d = d + s.psi + prm.keps;
        const float rd = __frcp_rn(d);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float x2 = fmaf(s.X[p].x, s.X[p].x, s.X[p].y * s.X[p].y);
            const float gs = s.C[p] * rd;
            const float2 g = make_float2(gs * s.X[p].x, -gs * s.X[p].y);   // C conj(X) / D
            float2 w = cfma(g, e, s.W[p]);
            w = make_float2(prm.ka * w.x, prm.ka * w.y);
            s.W[p] = w;
            const float w2 = fmaf(w.x, w.x, w.y * w.y);
            s.C[p] = fmaf(prm.ka2 * (1.f - gs * x2), s.C[p], prm.kq * w2);
        }
    }
    E = e;
    Yh = yh;
}

// half-size spectra -> the two real-signal bins of the mirrored pair (k, 256-k).
// fa = Zc[k], fb = Zc[256-k], w = exp(-2 pi i k/512).  The analysis window carries the 1/2.
__device__ __forceinline__ void unpack_pair(float2 fa, float2 fb, float2 w, float2& xk, float2& xm) {
    const float2 a = make_float2(fa.x + fb.x, fa.y - fb.y);          // Zc[k] + conj Zc[256-k]
    const float2 d = make_float2(fa.y + fb.y, fb.x - fa.x);          // (Zc[k] - conj Zc[256-k]) / i
    const float2 t = cmul(w, d);
    xk = make_float2(a.x + t.x, a.y + t.y);
    xm = make_float2(a.x - t.x, t.y - a.y);                          // conj(a - t)
}
// two real-signal bins (k, 256-k) -> half-size inverse spectrum entries G[k], G[256-k]
// (scale 1/512 lives in the synthesis table).
__device__ __forceinline__ void pack_pair(float2 ek, float2 em, float2 w, float2& gk, float2& gm) {
    const float2 a = make_float2(ek.x + em.x, ek.y - em.y);          // E[k] + conj E[256-k]
    const float2 d = make_float2(ek.x - em.x, ek.y + em.y);          // E[k] - conj E[256-k]
    const float2 t = cmulc(d, w);                                    // d * conj(w)
    gk = make_float2(a.x - t.y, a.y + t.x);
    gm = make_float2(a.x + t.y, t.x - a.y);
}

// The self-mirrored bin (k = N/4 on the half-size spectrum) has the split / packing twiddle -i, for which unpack_pair /
// pack_pair reduce to 2 conj(.) -- the same values, without their multiplies by 0 and -1.
__device__ __forceinline__ float2 mid_conj2(float2 z) { return make_float2(2.f * z.x, -2.f * z.y); }

constexpr int kRingPitch = 256;   // far-end history ring: one float2 column per thread and slot (bin 128 has its own history);
                                  // a power of two, so that a slot offset is a shift instead of an integer multiply

template <int NW, int P>
struct Stage1Smem {
    // long filters on the one-bin-per-thread kernel keep X[t-p] in a shared-memory ring
    static constexpr bool kRing = (NW == 8 && P >= 16);
    static constexpr int F = kRing ? 8 : 2 * NW;   // frames per chunk
    static constexpr int R = F + 1;                // staging ring, hops per signal
    static constexpr size_t zbuf_bytes = size_t(F) * 2 * kTilePitch * sizeof(float2);
    static constexpr size_t stage_bytes = size_t(2) * R * 256 * sizeof(float);
    // window tables: analysis half-table [128] float2 + synthesis table [256] float2
    static constexpr size_t win_bytes = size_t(128 + 256) * sizeof(float2);
    // state of the self-mirrored bin 128 (one thread's worth; kept out of everybody's registers)
    // state of bin 128: W, X, C per tap (+ Psi).  Ring kernels keep X as a contiguous history H[P + F] (the P past
    // frames followed by the F frames of the chunk being looked ahead: P*8 bytes of it are the X slot above) and add
    // Y[128] of those frames (float2 each) and the resulting E / Yhat (2 float2 each)
    static constexpr size_t mid_bytes = (size_t(P) * (8 + 8 + 4) + 4 + 15) / 16 * 16 + (kRing ? size_t(P + F) * 8 + size_t(F) * 24 : 0);
    static constexpr size_t ring_bytes = kRing ? size_t(P) * kRingPitch * sizeof(float2) : 0;
    // overlap-add: in-chunk tails live in the (dead after the inverse FFT) Zbuf tile of the frame
    // that produced them; only the last warp's tail crosses a chunk boundary -> one carry slot.
    __host__ __device__ static constexpr size_t tails_bytes(bool echo) {
        return size_t(echo ? 2 : 1) * 128 * sizeof(float2);
    }
    __host__ __device__ static constexpr size_t total(bool echo) {
        return zbuf_bytes + stage_bytes + win_bytes + tails_bytes(echo) + mid_bytes + ring_bytes + 16;
    }
    // fused feature epilogue: error-signal blocks [F+1][256], |X| of the chunk's frames [F][kMagPitch] (|STFT(err)| goes
    // into dead spectrum tiles), far-end band energies [F+1][32], band-major non-zero bank coefficients [512], per-band
    // silent-frame energy / first task / task count [3][32], projection tasks [64] x (first bin, bins, coefficient offset)
    // (the tasks' partial sums go into a dead tile too): 43.9 KB per utterance with the rest -> 5 utterances per SM
    static constexpr int kMagPitch = 260;
    static constexpr size_t feat_bytes =
        (size_t(F + 1) * 256 + size_t(F) * kMagPitch + (F + 1) * 32 + 512 + 3 * 32 + 3 * 64) * sizeof(float);
    __host__ __device__ static constexpr size_t total_feat(bool echo) { return total(echo) + feat_bytes; }
};

// REGS caps the registers per thread (occupancy knob: resident utterances per SM =
// 65536 / (32 * NW * REGS), also bounded by shared memory).
template <int NW, int P, int ALGO, bool ECHO, int REGS, bool FEAT = false>
__global__ void __launch_bounds__(NW * 32) __maxnreg__(REGS) stage1_n512_kernel(const Stage1Params prm) {
    using SM = Stage1Smem<NW, P>;
    constexpr bool kRing = SM::kRing;
    constexpr int F = SM::F, R = SM::R, NT = NW * 32;
    // NW <= 4: every thread owns PPT mirrored pairs (k, 256-k), i.e. 2*PPT bins.
    // NW == 8: one bin per thread (long filters: the per-bin state is what fills the registers);
    //          lanes 2i / 2i+1 hold the two bins of pair i.
    constexpr bool kSingleBin = (NW == 8);
    constexpr int PPT = kSingleBin ? 1 : 128 / (kSingleBin ? 128 : NT);   // mirrored pairs per thread
    constexpr int NBIN = kSingleBin ? 1 : 2 * PPT;                        // bins per thread
    static_assert(NW == 1 || NW == 2 || NW == 4 || NW == 8, "1, 2, 4 or 8 warps per utterance");
    static_assert(!FEAT || (NW == 2 && !kRing), "the fused feature epilogue is built on the two-warp kernel");
    constexpr int NSIG = ECHO ? 2 : 1;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2* zbuf = reinterpret_cast<float2*>(smem_raw);                               // [F][2][kTilePitch]
    float* stage = reinterpret_cast<float*>(smem_raw + SM::zbuf_bytes);               // [2][R][256]
    float2* win_a = reinterpret_cast<float2*>(smem_raw + SM::zbuf_bytes + SM::stage_bytes);      // [128]
    float2* win_s = win_a + 128;                                                               // [256]
    float2* carry = win_s + 256;                       // [NSIG][128]   last warp's tail, crosses chunks
    float* mid_state = reinterpret_cast<float*>(carry + NSIG * 128);                             // [<= 96]
    float2* xring = reinterpret_cast<float2*>(smem_raw + SM::zbuf_bytes + SM::stage_bytes + SM::win_bytes +
                                              SM::tails_bytes(ECHO) + SM::mid_bytes);        // [P][kRingPitch]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + SM::zbuf_bytes + SM::stage_bytes + SM::win_bytes +
                                                 SM::tails_bytes(ECHO) + SM::mid_bytes + SM::ring_bytes);

    // fused Stage-2 feature epilogue (FEAT): see feat_epilogue below
    constexpr int kMagPitch = SM::kMagPitch;
    float* f_ring = reinterpret_cast<float*>(smem_raw + SM::total(ECHO));   // [F+1][256]  error-signal blocks, slot = block % (F+1)
    float* f_magX = f_ring + (F + 1) * 256;                                 // [F][kMagPitch]  |X| of the chunk's frames
    float* f_xerb = f_magX + F * kMagPitch;                                 // [F+1][32] far-end band energies, slot = frame % (F+1)
    float* f_cb = f_xerb + (F + 1) * 32;                                    // [512]     non-zero bank coefficients, band-major
    float* f_csum = f_cb + 512;                                             // [32]      band energy of a silent frame
    int* f_tb0 = reinterpret_cast<int*>(f_csum + 32);                       // [32]      first projection task of the band
    int* f_tbn = f_tb0 + 32;                                                // [32]      tasks of the band
    int* f_tk = f_tbn + 32;                                                 // [64]      task: first bin
    int* f_tn = f_tk + 64;                                                  // [64]      task: bins (0 = idle lane)
    int* f_tc = f_tn + 64;                                                  // [64]      task: offset into f_cb
    float* f_part = reinterpret_cast<float*>(zbuf + 1 * kTilePitch);         // [64][8]   partial band energies of the tasks: in the
                                                                             //           first microphone tile, free after phase D
    // |STFT(err)| of frame t0 - 1 + hw lives in the (dead) far-end tile of half-warp hw during the epilogue
    auto mag_e = [&](int hw) { return reinterpret_cast<float*>(zbuf + (hw * 2 + 0) * kTilePitch); };
    static_assert(!FEAT || kMagPitch * sizeof(float) <= kTilePitch * sizeof(float2), "magnitude row must fit a tile");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, h = lane & 15;
    const long long b = blockIdx.x;

    long long n_ll = prm.n_samples ? prm.n_samples[b] : prm.L;
    n_ll = n_ll < 0 ? 0 : (n_ll > prm.L ? prm.L : n_ll);
    const int T = static_cast<int>(n_ll) / 256 + 1;               // frames (attention_ccrn.py:48-49 with N = 2H)
    // Row pointers are recomputed where they are used (a handful of integer instructions per chunk)
    // instead of living in registers across the FFT phases, where the compiler would spill them:
    // with the 228 KB shared-memory carve-out there is no L1, so a spill reload costs an L2 round trip.
    auto row_off = [&](long long stride) {
        unsigned bx;
        asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));   // re-read where used: nothing hoisted, nothing spilled
        return static_cast<long long>(bx) * stride;
    };

    // Owner of the self-mirrored bin 128 (the 257th bin on 256 bin slots; its serial pass costs one warp
    // ~70 instructions per frame).  Warp w of a two-warp utterance always sits on scheduler (slot + w) % 4
    // of its SM, so a fixed owner loads one scheduler of each pair more than the other for every resident
    // utterance at once.  Alternate the owner with the hardware warp slot: co-resident utterances on the
    // same scheduler pair then put the extra pass on different schedulers.
    int* mid_owner = reinterpret_cast<int*>(mbar + 1);
    if (tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
        unsigned hw_warp;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw_warp));
        *mid_owner = (NW == 2) ? static_cast<int>((hw_warp >> 2) & 1u) : NW - 1;
    }
#ifdef AEC_PHASE_TIMING
    __shared__ long long dbg_sm[NW][12];
    if (lane == 0) {
        for (int i = 0; i < 11; ++i) dbg_sm[warp][i] = 0;
        dbg_sm[warp][11] = clock64();
    }
#endif
    // Utterances resident on one SM start together and would otherwise run their FADD-heavy FFT
    // phases and FFMA-heavy filter phases in lock-step; skew them by a fraction of a chunk.
    if (prm.stagger_ns > 0) {
        const unsigned slot = (unsigned)(blockIdx.x / (unsigned)prm.num_sms) & 7u;
        if (slot) __nanosleep(slot * (unsigned)prm.stagger_ns);
    }
    __syncthreads();

    // hops (padded-signal blocks) [b0, b1] -> staging ring.  Block beta holds samples
    // [(beta-1)*256, beta*256); block 0 is the reference's left zero pad.
    auto produce = [&](int b0, int b1) {
        if (b1 > T) b1 = T;
        const long long in_off = row_off(prm.in_stride);
        const float* far_b = prm.far + in_off;
        const float* mic_b = prm.mic + in_off;
        // Interior chunk (every hop of the range lies inside the signal, rows 16-byte aligned): the hops
        // are contiguous in HBM and contiguous in the ring up to its wrap-around, so the whole range is
        // one or two bulk copies per signal instead of one per hop (the per-hop issue loop used to cost
        // the producing warp ~60 instructions per frame on the critical path of the filter phase).
        if (prm.use_tma && b0 >= 1 && b1 >= b0 && b1 <= T - 1) {   // b1 * 256 <= n
            if (elect_one()) {
                const int nb = b1 - b0 + 1;
                const int s0 = b0 % R;
                const int run1 = nb < R - s0 ? nb : R - s0;
                const float* src_f = far_b + (b0 - 1) * 256;
                const float* src_m = mic_b + (b0 - 1) * 256;
                fence_proxy_async();
                mbar_arrive_expect_tx(mbar, static_cast<uint32_t>(nb) * 2048u);
                tma_load_1d(stage + (0 * R + s0) * 256, src_f, static_cast<uint32_t>(run1) * 1024u, mbar);
                tma_load_1d(stage + (1 * R + s0) * 256, src_m, static_cast<uint32_t>(run1) * 1024u, mbar);
                if (run1 < nb) {
                    const uint32_t rest = static_cast<uint32_t>(nb - run1) * 1024u;
                    tma_load_1d(stage + (0 * R) * 256, src_f + run1 * 256, rest, mbar);
                    tma_load_1d(stage + (1 * R) * 256, src_m + run1 * 256, rest, mbar);
                }
            }
            return;
        }
        // (slow path, first chunk / ragged edge / unaligned rows only: the sample count is re-read here so
        //  that it does not occupy a register across the whole chunk loop)
        long long n_re = prm.n_samples ? prm.n_samples[blockIdx.x] : prm.L;
        n_re = n_re < 0 ? 0 : (n_re > prm.L ? prm.L : n_re);
        const int n = static_cast<int>(n_re);
        if (lane == 0) {
            int nt = 0;
            for (int beta = b0; beta <= b1; ++beta) nt += (prm.use_tma && beta >= 1 && beta * 256 <= n) ? 1 : 0;
            fence_proxy_async();
            mbar_arrive_expect_tx(mbar, static_cast<uint32_t>(nt) * 2048u);
            for (int beta = b0; beta <= b1; ++beta) {
                if (prm.use_tma && beta >= 1 && beta * 256 <= n) {
                    const int slot = beta % R;
                    tma_load_1d(stage + (0 * R + slot) * 256, far_b + (beta - 1) * 256, 1024u, mbar);
                    tma_load_1d(stage + (1 * R + slot) * 256, mic_b + (beta - 1) * 256, 1024u, mbar);
                }
            }
        }
        for (int beta = b0; beta <= b1; ++beta) {
            if (!(prm.use_tma && beta >= 1 && beta * 256 <= n)) {
                const int slot = beta % R;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int o = lane + 32 * i;
                    const int idx = (beta - 1) * 256 + o;
                    const bool ok = beta >= 1 && idx < n;
                    stage[(0 * R + slot) * 256 + o] = ok ? __ldg(far_b + idx) : 0.f;
                    stage[(1 * R + slot) * 256 + o] = ok ? __ldg(mic_b + idx) : 0.f;
                }
            }
        }
    };

    if (warp == 0) produce(0, F);

    // window tables -> shared memory (so the kernel does not depend on L1 residency: with 7
    // utterances per SM the shared-memory carve-out leaves almost no L1), twiddles -> registers
    if constexpr (kRing) {
        for (int i = tid; i < P * kRingPitch; i += NT) xring[i] = make_float2(0.f, 0.f);   // X[t<0] = 0
    }
    for (int i = tid; i < 128; i += NT) win_a[i] = __ldg(&prm.win_a[i]);
    for (int i = tid; i < 256; i += NT) win_s[i] = __ldg(&prm.win_s[i]);
    if constexpr (FEAT) {
        // The cosine bank is ~2 non-zeros per bin (483 of 257 x 32 for the reference's bank).  Each band's non-zero range is
        // staged contiguously, and the projection work is cut into at most 64 TASKS of equal length (a band and a run of
        // its bins), one per thread: band-parallel, the 51-bin band would set the length of the loop for everybody.
        if (tid < 32) {
            int lo = 257, hi = 0;
            for (int k = 0; k < 257; ++k)
                if (__ldg(prm.erb + k * 32 + tid) != 0.f) {
                    lo = lo < k ? lo : k;
                    hi = k + 1;
                }
            int len = hi > lo ? hi - lo : 0;
            int off = len;                               // inclusive prefix sum over the bands
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, off, o);
                if (lane >= o) off += v;
            }
            off -= len;
            len = off + len <= 512 ? len : (off < 512 ? 512 - off : 0);   // (the host side rejects larger banks)
            off = off < 512 ? off : 0;
            float cs = 0.f;
            const float silent = sqrtf(1e-9f);           // magnitude of an all-zero frame (ERB.py:277-278)
            for (int i = 0; i < len; ++i) {
                const float c = __ldg(prm.erb + (lo + i) * 32 + tid);
                f_cb[off + i] = c;
                cs = fmaf(silent, c, cs);
            }
            f_csum[tid] = cs;
            // smallest run length for which the bands need no more than 64 tasks
            int run = 8, ntask = 0, base = 0;
            for (;; ++run) {
                ntask = (len + run - 1) / run;
                int tot = ntask;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, tot, o);
                    if (lane >= o) tot += v;
                }
                base = tot - ntask;
                if (__shfl_sync(0xffffffffu, tot, 31) <= 64) break;
            }
            f_tb0[tid] = base;
            f_tbn[tid] = ntask;
            for (int i = 0; i < ntask; ++i) {
                f_tk[base + i] = lo + i * run;
                f_tn[base + i] = (len - i * run) < run ? (len - i * run) : run;
                f_tc[base + i] = off + i * run;
            }
            const int total = __shfl_sync(0xffffffffu, base + ntask, 31);
            for (int i = total + tid; i < 64; i += 32) {
                f_tk[i] = 0;
                f_tn[i] = 0;
                f_tc[i] = 0;
            }
        }
    }
    TwiddleRegs twr;
    twr.w1 = __ldg(&prm.tw256[1 * 16 + h]);
    twr.w2 = __ldg(&prm.tw256[2 * 16 + h]);
    twr.w4 = __ldg(&prm.tw256[4 * 16 + h]);
    twr.w8 = __ldg(&prm.tw256[8 * 16 + h]);

    // ---- persistent recurrence state -------------------------------------------------------
    BinState<P, ALGO> st[NBIN];
    float2 wk[PPT];
#pragma unroll
    for (int i = 0; i < NBIN; ++i) st[i].init(prm.kc0);
#pragma unroll
    for (int i = 0; i < PPT; ++i) wk[i] = __ldg(&prm.tw512[kSingleBin ? (tid >> 1) : tid + i * NT]);
    // single-bin mode: the same formulas serve both bins of a pair with per-thread twiddles
    //   unpack: X = A + wu * D,  A = own + conj(other), D = (own - conj(other)) / i
    //           wu = w for bin k, -conj(w) for bin 256-k
    //   pack  : G = a + i * (d * ws),  a = E_own + conj(E_other), d = E_own - conj(E_other)
    //           ws = conj(w) for bin k, -w for bin 256-k
    const bool side = kSingleBin && (tid & 1);
    const float2 wu = side ? make_float2(-wk[0].x, wk[0].y) : wk[0];
    const float2 ws = side ? make_float2(-wk[0].x, -wk[0].y) : make_float2(wk[0].x, -wk[0].y);
    const int k_own = side ? ((256 - (tid >> 1)) & 255) : (tid >> 1);
    const int k_oth = side ? (tid >> 1) : ((256 - (tid >> 1)) & 255);
    const int k_bin = side ? 256 - (tid >> 1) : (tid >> 1);      // true bin index 0..256 (ring column)
    // Bin 128 is its own mirror, the 257th bin on 256 bin slots.
    //  * short filters (two-warp kernel): serial on one lane inside the frame loop, state in shared memory.
    //  * 8 partitions and more: a serial pass on one lane would double the filter phase of its warp (every
    //    other warp then waits at the barrier: measured 25 % of the frame time of the 16-partition Kalman
    //    kernel).  The last warp runs it TAP-PARALLEL instead (lane p owns tap p, the sums over taps are
    //    shuffle reductions) in a loop over the chunk's frames AFTER its regular bins -- no pair slot touches
    //    entry 128 of the tiles -- with the tap state in registers for the duration of that loop only
    //    (loaded from / stored to shared memory once per chunk).
    static_assert(P <= 32, "one lane per tap");
    // (measured, ms per 2048 x 10 s: 16-partition Kalman 9.57 -> 9.02 against the former in-loop shared-memory
    //  form; 16-partition NLMS on 8 warps 7.19 -> 5.96; 8 partitions are faster serial -- 2.12 vs 2.34 ms per
    //  1024 -- as is the 4-warp 16-partition NLMS kernel, whose chunk is only 8 frames of a cheap serial pass)
    constexpr bool kMidTapParallel = (P >= 16) && (ALGO == kAlgoKalman || NW == 8);
    // Ring kernels (8 warps, 8 frames per chunk): only warps 0-3 have synthesis work, warps 4-7 used to wait at
    // the barrier.  They now run bin 128 of the NEXT chunk in that window.  X[128] and Y[128] need no FFT:
    // e^{-2 pi i 128 n / 512} = (-i)^n, so they are four interleaved sums of the windowed samples, which the
    // bulk copies for the next chunk have already delivered.  The serial recurrence of that bin -- 25 % of the
    // frame time when it ran inside the filter phase -- is off the critical path altogether.
    constexpr bool kMidAhead = kRing;
    static_assert(!kMidAhead || (kMidTapParallel && F == 8 && NW == 8), "look-ahead needs the idle synthesis warps");
    static_assert(BinState<P, ALGO>::kFloats * sizeof(float) <= SM::mid_bytes, "mid-bin state does not fit");
    float2* midW = reinterpret_cast<float2*>(mid_state);           // [P]
    float2* midX = midW + P;                                       // [P]
    float* midC = reinterpret_cast<float*>(midX + P);              // [P]
    float* midPsi = midC + P;                                      // [1]
    // look-ahead buffers (ring kernels); midH[i] = X[128] at frame tc0 - P + i while chunk tc0 is being looked ahead
    float2* midH = reinterpret_cast<float2*>(mid_state + (size_t(P) * 20 + 4 + 15) / 16 * 4);   // [P + F]
    float2* midY = midH + P + F;                                   // [F]    Y[128] of the chunk's frames
    float2* midE = midY + F;                                       // [F][2] E[128], Yhat[128] (packed when injected)
    if constexpr (kMidTapParallel) {
        if (warp == NW - 1 && lane < P) {
            midW[lane] = make_float2(0.f, 0.f);
            midX[lane] = make_float2(0.f, 0.f);
            if constexpr (kRing) midH[lane] = make_float2(0.f, 0.f);       // X[t < 0] = 0
            midC[lane] = prm.kc0;
            if (lane == 0) *midPsi = 0.f;
        }
    } else if (tid == NT - 1) {
        BinState<P, ALGO> st_mid;
        st_mid.init(prm.kc0);
        st_mid.store(mid_state);
    }

    // ---- bin 128 of the chunk starting at frame tc0, run by warps F/2 .. NW-1 (ring kernels) --------
    // step 1 (all four warps, two frames each): X[128], Y[128] straight from the staged samples;
    // step 2 (last warp, one lane per tap): the recurrence over the chunk's frames, tap state in registers.
    auto mid_ahead = [&](int tc0, uint32_t parity) {
      if constexpr (kMidAhead) {
        mbar_wait(mbar, parity);                                   // the chunk's hops have landed
        {
            // step 1: this warp's two frames x two signals as four independent sums (one joint reduction)
            const int tl0 = (warp - F / 2) * 2;
            float r[4], q[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int t = tc0 + tl0 + j;
                t = t < T ? t : T - 1;                             // (past the end: recomputes a valid frame, not stored)
#pragma unroll
                for (int sig = 0; sig < 2; ++sig) {
                    const float* s0 = stage + (sig * R + (t % R)) * 256 + 2 * lane;
                    const float* s1 = stage + (sig * R + ((t + 1) % R)) * 256 + 2 * lane;
                    float a = 0.f, b2 = 0.f;                       // sums over n = 0 (2) mod 4 and n = 1 (3) mod 4
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 w = win_a[lane + 32 * i];     // 0.5 hann[2m], 0.5 hann[2m+1], m = lane + 32 i
                        const float2 x0 = *reinterpret_cast<const float2*>(s0 + 64 * i);
                        const float2 x1 = *reinterpret_cast<const float2*>(s1 + 64 * i);
                        a = fmaf(x0.x, w.x, a);
                        b2 = fmaf(x0.y, w.y, b2);
                        a = fmaf(x1.x, 0.5f - w.x, a);             // hann[n + 256] = 1 - hann[n]
                        b2 = fmaf(x1.y, 0.5f - w.y, b2);
                    }
                    // (-i)^n: even lanes hold n = 0, 1 (mod 4): +a, -i b ; odd lanes n = 2, 3: -a, +i b
                    r[2 * j + sig] = (lane & 1) ? -a : a;
                    q[2 * j + sig] = (lane & 1) ? b2 : -b2;
                }
            }
            // eight sums over the 32 lanes in 9 shuffles instead of 40 (the shuffle unit is shared by the SM and was the
            // longest part of this step): at every level a lane keeps half of its values and hands the other half to its
            // partner; after three levels it owns one value, two more levels finish it.  The value a lane ends up with:
            // bit 4 of the lane selects q over r, bit 3 the frame (j), bit 2 the signal.
            float v8[8] = {r[0], r[1], r[2], r[3], q[0], q[1], q[2], q[3]};    // index = 4 * (q ? 1 : 0) + 2 * j + sig
            float v4[4], v2[2], v1;
            {
                const bool up = lane & 16;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float mine = up ? v8[4 + k] : v8[k], other = up ? v8[k] : v8[4 + k];
                    v4[k] = mine + __shfl_xor_sync(0xffffffffu, other, 16);
                }
            }
            {
                const bool up = lane & 8;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float mine = up ? v4[2 + k] : v4[k], other = up ? v4[k] : v4[2 + k];
                    v2[k] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
                }
            }
            {
                const bool up = lane & 4;
                const float mine = up ? v2[1] : v2[0], other = up ? v2[0] : v2[1];
                v1 = mine + __shfl_xor_sync(0xffffffffu, other, 4);
            }
            v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
            // lane 4 * i holds value i: [r: (j0 far, j0 mic, j1 far, j1 mic)] in lanes 0, 4, 8, 12, [q: ...] in 16 .. 28
            {
                const int j = (lane >> 3) & 1, sig = (lane >> 2) & 1, isq = lane >> 4;
                if ((lane & 3) == 0 && tc0 + tl0 + j < T) {            // (the table carries the 1/2)
                    float* dstf = reinterpret_cast<float*>(sig == 0 ? midH + P + tl0 + j : midY + tl0 + j);
                    dstf[isq] = 2.f * v1;
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"((NW - F / 2) * 32) : "memory");     // warps F/2 .. NW-1 only
        if (warp == NW - 1) {
            // step 2: the recurrence, one lane per tap, tap state in registers for the chunk.  Lane p reads its far-end
            // value of frame tl straight from the history (H[P + tl - p]): nothing but the filter taps, the covariances
            // and Psi is carried from frame to frame, and only the sums over taps cross lanes.
            constexpr int kLanes = P;
            static_assert((kLanes & (kLanes - 1)) == 0 && kLanes <= 32, "one lane per tap");
            auto wsum = [](float v) {
#pragma unroll
                for (int o = 1; o < kLanes; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                return v;
            };
            const bool tap = lane < kLanes;   // the other lanes carry zeros
            const int pl = tap ? lane : 0;
            float2 w = tap ? midW[pl] : make_float2(0.f, 0.f);
            float cv = (ALGO == kAlgoKalman && tap) ? midC[pl] : 0.f;
            float psi = (ALGO == kAlgoKalman) ? *midPsi : 0.f;
            const float2* hp = midH + (P - pl);
            // fully unrolled (8 frames).  A chunk that lies completely inside the utterance (all but the last one) runs
            // without per-frame conditions, so that every load is hoisted to the top and only arithmetic and shuffles
            // remain on the frame-to-frame chain; the reciprocal is MUFU + one Newton step written out (what __frcp_rn
            // does on its fast path, without its range-check branch: D >= eps is far from the denormals).
            auto run = [&](auto guard_tag) {
                constexpr bool kGuard = decltype(guard_tag)::value;
#pragma unroll
                for (int tl = 0; tl < F; ++tl) {
                    if (!kGuard || tc0 + tl < T) {
                        float2 x = hp[tl];
                        if (!tap) x = make_float2(0.f, 0.f);
                        const float2 yn = midY[tl];
                        const float2 prod = cmul(w, x);
                        const float x2 = fmaf(x.x, x.x, x.y * x.y);
                        const float2 yh = make_float2(wsum(prod.x), wsum(prod.y));
                        const float2 e = csub(yn, yh);
                        if constexpr (ALGO == kAlgoNlms) {
                            const float g = prm.mu * rcp_fast(wsum(x2) + prm.delta);
                            w = cfmac(x, make_float2(g * e.x, g * e.y), w);
                        } else {
                            const float sx = wsum(cv * x2);
                            psi = fmaf(prm.klam, psi, prm.koml * fmaf(e.x, e.x, e.y * e.y));
                            const float d = sx + psi + prm.keps;
                            const float r0 = rcp_fast(d);
                            const float rd = fmaf(r0, -fmaf(d, r0, -1.f), r0);
                            const float gs = cv * rd;
                            w = cfma(make_float2(gs * x.x, -gs * x.y), e, w);
                            w = make_float2(prm.ka * w.x, prm.ka * w.y);
                            cv = fmaf(prm.ka2 * (1.f - gs * x2), cv, prm.kq * fmaf(w.x, w.x, w.y * w.y));
                        }
                        if (lane == 0) {
                            midE[2 * tl] = e;
                            midE[2 * tl + 1] = yh;
                        }
                    }
                }
            };
            if (tc0 + F <= T) run(std::false_type{});
            else run(std::true_type{});
            if (tap) {
                midW[pl] = w;
                if constexpr (ALGO == kAlgoKalman) midC[pl] = cv;
            }
            if constexpr (ALGO == kAlgoKalman) if (lane == 0) *midPsi = psi;
            // slide the history: the last P frames become the past of the next chunk
            const float2 keep = tap ? midH[F + pl] : make_float2(0.f, 0.f);
            __syncwarp();
            if (tap) midH[pl] = keep;
        }
      }
    };
    if constexpr (kMidAhead) {
        __syncthreads();                      // window table, tap state and the producer's manual-path stores
        if (warp >= F / 2) mid_ahead(0, 0u);
    }

    // ---- fused Stage-2 feature epilogue, parity-preserving form -----------------------------
    // The reference front end takes STFTs of TIME-DOMAIN signals (ERB.py:262-264).  The error spectrum E the filter holds
    // is not the STFT of the error signal it synthesises (overlap-add + re-analysis projects it; measured 48 % median
    // difference of the band energies), so the epilogue re-analyses the synthesised error hops --
    // which are on chip in phase C -- one frame behind the synthesis: frame tau of STFT(err) needs output blocks tau and
    // tau + 1.  The far-end band energies come from |X| of the analysis the filter already did.  Per chunk:
    //   phase D   4 half-warps = 4 frames tau = t0-1 .. t0+2: window the error blocks, FFT-256, unpack, magnitudes
    //   project   lane pair (2j, 2j+1) owns band j: two frames each of |X| (this chunk) and |STFT(err)| (one behind),
    //             a loop over the band's non-zero bins; feat[tau] = cat[err_erb, |err_erb - far_erb|]
    // `in_norm` (the batch-global mean / std shift of ERB.py:254-256) is off in this path: the shift of the error signal
    // is not known before the whole batch has been processed (aec_features[_dev] on the stored error signal covers it).
    auto feat_epilogue = [&](int t0) {
      if constexpr (FEAT) {
        {
            const int hw = 2 * warp + half;
            const int tau = t0 - 1 + hw;
            const bool live = tau >= 0 && tau <= T - 1;
            float2 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int beta = tau + (j >> 3);                       // block beta = output hop beta - 1; 0 and T are zero pads
                const bool ok = live && beta >= 1 && beta <= T - 1;
                float2 x = make_float2(0.f, 0.f);
                if (ok) x = *reinterpret_cast<const float2*>(f_ring + (beta % (F + 1)) * 256 + 2 * h + 32 * (j & 7));
                float2 w = win_a[h + 16 * (j & 7)];
                if (j >= 8) w = make_float2(0.5f - w.x, 0.5f - w.y);
                v[j] = make_float2(x.x * w.x, x.y * w.y);
            }
            float2* tile = zbuf + (hw * 2 + 1) * kTilePitch;           // (every spectrum tile is dead by now)
            fft256_halfwarp_regs<false>(v, tile, twr, h);
#pragma unroll
            for (int p = 0; p < 16; ++p) tile[h + 16 * fft16_index(p)] = v[p];
            __syncwarp();
            // (the far-end tile of this half-warp is dead as well: it takes the magnitudes)
            float* me = mag_e(hw);
            if (live) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int k = h + 16 * q;
                    float2 xk, xm;
                    unpack_pair(tile[k], tile[(256 - k) & 255], __ldg(&prm.tw512[k]), xk, xm);
                    if (k == 0) {
                        xk.y = 0.f;
                        xm.y = 0.f;
                    }
                    me[k] = sqrt_approx(fmaf(xk.x, xk.x, fmaf(xk.y, xk.y, 1e-9f)));        // ERB.py:277-278
                    me[256 - k] = sqrt_approx(fmaf(xm.x, xm.x, fmaf(xm.y, xm.y, 1e-9f)));
                }
                if (h == 0) {
                    float2 xk, xm;
                    unpack_pair(tile[128], tile[128], make_float2(0.f, -1.f), xk, xm);
                    me[128] = sqrt_approx(fmaf(xk.x, xk.x, fmaf(xk.y, xk.y, 1e-9f)));
                }
            }
        }
        __syncthreads();
        {
            // one task per thread: a run of one band's bins for the four far-end and the four error frames of the chunk
            const int k0 = f_tk[tid], n = f_tn[tid];
            const float* c = f_cb + f_tc[tid];
            float ax[F], ae[F];
#pragma unroll
            for (int r = 0; r < F; ++r) {
                ax[r] = 0.f;
                ae[r] = 0.f;
            }
            for (int i = 0; i < n; ++i) {                              // ERB.py:282-283
                const float cc = c[i];
#pragma unroll
                for (int r = 0; r < F; ++r) {
                    ax[r] = fmaf(f_magX[r * kMagPitch + k0 + i], cc, ax[r]);
                    ae[r] = fmaf(mag_e(r)[k0 + i], cc, ae[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < F; ++r) {
                f_part[tid * 8 + r] = ax[r];
                f_part[tid * 8 + F + r] = ae[r];
            }
        }
        __syncthreads();
        {
            // lane pair (2j, 2j+1) finishes band j: the band's tasks in ascending order, two frames per lane
            const int j = tid >> 1, fp = tid & 1;
            const int b0 = f_tb0[j], bn = f_tbn[j];
            float ax0 = 0.f, ax1 = 0.f, ae0 = 0.f, ae1 = 0.f;
            for (int i = 0; i < bn; ++i) {
                const float* pp = f_part + (b0 + i) * 8 + 2 * fp;
                ax0 += pp[0];
                ax1 += pp[1];
                ae0 += pp[F];
                ae1 += pp[F + 1];
            }
            const int tx = t0 + 2 * fp;                                // far-end frames of this chunk
            if (tx < T) f_xerb[(tx % (F + 1)) * 32 + j] = ax0;
            if (tx + 1 < T) f_xerb[((tx + 1) % (F + 1)) * 32 + j] = ax1;
            __syncwarp();                                              // the pair partner's far-end energies
            float* fb = prm.feat + (static_cast<long long>(blockIdx.x) * prm.feat_frames) * 64;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int tau = t0 - 1 + 2 * fp + r;
                if (tau >= 0 && tau <= T - 1 && tau < prm.feat_frames) {
                    const float e = r ? ae1 : ae0;
                    const float xe = f_xerb[(tau % (F + 1)) * 32 + j];
                    fb[static_cast<long long>(tau) * 64 + j] = e;
                    fb[static_cast<long long>(tau) * 64 + 32 + j] = fabsf(e - xe);   // ERB.py:287-290
                }
            }
        }
      }
    };

    float acc_e = 0.f;     // ERLE energies in ONE register: lanes 16-31 accumulate the microphone, lanes 0-15 the error

    // (chunk count not kept in a register: every use is a comparison of the frame index with T)
    for (int c = 0; c * F < T; ++c) {
        const int t0 = c * F;
        AEC_TICK(0);                         // loop tail / prologue
        __syncthreads();                     // manual-path staging stores of the producer visible
        mbar_wait(mbar, static_cast<uint32_t>(c & 1));
        AEC_TICK(1);                         // barrier + staged-data wait

        // ================= phase A : analysis =================
#pragma unroll 1
        for (int tl = warp; tl < F; tl += NW) {
            const int t = t0 + tl;
            if (t < T) {
                const float* s0 = stage + (half * R + (t % R)) * 256 + 2 * h;
                const float* s1 = stage + (half * R + ((t + 1) % R)) * 256 + 2 * h;
                float2 v[16];
                float e_acc = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 x = *reinterpret_cast<const float2*>((j < 8 ? s0 : s1) + 32 * (j & 7));
                    // periodic Hann: w[n + 256] = 1 - w[n]; the table holds 0.5 w[n], n < 256
                    float2 w = win_a[h + 16 * (j & 7)];
                    if (j >= 8) w = make_float2(0.5f - w.x, 0.5f - w.y);
                    v[j] = make_float2(x.x * w.x, x.y * w.y);
                    if (j >= 8) e_acc = fmaf(x.x, x.x, fmaf(x.y, x.y, e_acc));
                }
                // second half of frame t is output hop t (block t+1): inside the ERLE span?
                if (prm.erle_db != nullptr && half == 1 && t + 1 <= T - 1 && t >= prm.erle_skip_hops) acc_e += e_acc;
                float2* tile = zbuf + (tl * 2 + half) * kTilePitch;
                fft256_halfwarp_regs<false>(v, tile, twr, h);
#pragma unroll
                for (int p = 0; p < 16; ++p) tile[h + 16 * fft16_index(p)] = v[p];
            }
        }
        AEC_TICK(2);                         // phase A
        __syncthreads();
        AEC_TICK(3);                         // barrier after A

        // stage the next chunk's hops while the filter and synthesis phases run
        if (warp == 0 && t0 + F < T) produce(t0 + F + 1, t0 + 2 * F);
        AEC_TICK(4);                         // produce

        // ================= phase B : per-bin recurrence =================
        if constexpr (kMidAhead) {
            // entries 128 of the tiles (no pair slot touches them): E / Yhat of bin 128, computed one chunk ahead
            if (tid < F && t0 + tid < T) {
                const float2 e = midE[2 * tid], yh = midE[2 * tid + 1];
                zbuf[(tid * 2 + 0) * kTilePitch + 128] = mid_conj2(e);
                if constexpr (ECHO) zbuf[(tid * 2 + 1) * kTilePitch + 128] = mid_conj2(yh);
            }
        }
        // (frame loop deliberately NOT unrolled: the chunk loop is ~30 KB of executed SASS against a 32 KB
        //  instruction cache; unrolling by 2 / 4 removes the register moves of the history shift and costs
        //  +3.7 % / +12 % on the whole kernel -- re-measured after the synthesis loop had shrunk)
#pragma unroll 1
        for (int tl = 0; tl < F; ++tl) {
            if (t0 + tl < T) {
                float2* zf = zbuf + (tl * 2 + 0) * kTilePitch;
                float2* zm = zbuf + (tl * 2 + 1) * kTilePitch;
                if constexpr (kSingleBin) {
                    const float2 fo = zf[k_own], fx = zf[k_oth], mo = zm[k_own], mx = zm[k_oth];
                    auto unpack1 = [&](float2 own, float2 oth) {
                        const float2 a = make_float2(own.x + oth.x, own.y - oth.y);
                        const float2 d = make_float2(own.y + oth.y, oth.x - own.x);
                        return cadd(a, cmul(wu, d));
                    };
                    auto pack1 = [&](float2 eo) {
                        const float2 ex = make_float2(__shfl_xor_sync(0xffffffffu, eo.x, 1),
                                                      __shfl_xor_sync(0xffffffffu, eo.y, 1));
                        const float2 a = make_float2(eo.x + ex.x, eo.y - ex.y);
                        const float2 d = make_float2(eo.x - ex.x, eo.y + ex.y);
                        const float2 t = cmul(d, ws);
                        return make_float2(a.x - t.y, a.y + t.x);
                    };
                    float2 e, yh;
                    const float2 xn = unpack1(fo, fx);
                    if constexpr (kRing) {
                        const int tt = t0 + tl;
                        float2* col = xring + tid;       // (any one-to-one column assignment: a thread reads back its own)
#pragma unroll
                        for (int p = 1; p < P; ++p) st[0].X[p] = col[((tt - p) & (P - 1)) * kRingPitch];
                        col[(tt & (P - 1)) * kRingPitch] = xn;
                    }
                    bin_step<P, ALGO, !kRing>(st[0], xn, unpack1(mo, mx), prm, e, yh);
                    zf[k_own] = pack1(e);
                    if constexpr (ECHO) zm[k_own] = pack1(yh);
                } else
#pragma unroll
                for (int i = 0; i < PPT; ++i) {
                    const int k = tid + i * NT;
                    const int km = (256 - k) & 255;
                    float2 xk, xm, yk, ym, ek, em, hk, hm, gk, gm;
                    unpack_pair(zf[k], zf[km], wk[i], xk, xm);
                    unpack_pair(zm[k], zm[km], wk[i], yk, ym);
                    if constexpr (FEAT) {                 // |X| for the far-end band energies (bins k and 256 - k)
                        float* mx = f_magX + tl * kMagPitch;
                        mx[k] = sqrt_approx(fmaf(xk.x, xk.x, fmaf(xk.y, xk.y, 1e-9f)));
                        mx[256 - k] = sqrt_approx(fmaf(xm.x, xm.x, fmaf(xm.y, xm.y, 1e-9f)));
                    }
                    bin_step<P, ALGO>(st[2 * i], xk, yk, prm, ek, hk);
                    bin_step<P, ALGO>(st[2 * i + 1], xm, ym, prm, em, hm);
                    // (k == 0 is the DC / Nyquist pair: X, Y are exactly real there, so W, E and Yhat stay
                    //  exactly real and need no masking of the imaginary parts before packing)
                    pack_pair(ek, em, wk[i], gk, gm);
                    zf[k] = gk;
                    zf[km] = gm;
                    if constexpr (ECHO) {
                        pack_pair(hk, hm, wk[i], gk, gm);
                        zm[k] = gk;
                        zm[km] = gm;
                    }
                }
                if constexpr (!kMidTapParallel) {
                    // self-mirrored bin 128, serial on the last lane of the owning warp (owner re-read from shared
                    // memory: a register live across the chunk loop tips the 128-register build into spilling)
                    if (tid == *static_cast<volatile int*>(mid_owner) * 32 + 31) {
                        // (split / packing twiddle of this bin is -i: the real-signal bin is 2 conj(Z[128]) and back -- the
                        //  same values as unpack_pair / pack_pair produce, without their multiplies by 0 and -1)
                        float2 ek, hk;
                        const float2 xk = mid_conj2(zf[128]), yk = mid_conj2(zm[128]);
                        if constexpr (FEAT) f_magX[tl * kMagPitch + 128] = sqrt_approx(fmaf(xk.x, xk.x, fmaf(xk.y, xk.y, 1e-9f)));
                        BinState<P, ALGO> st_mid;
                        st_mid.load(mid_state);
                        bin_step<P, ALGO>(st_mid, xk, yk, prm, ek, hk);
                        st_mid.store(mid_state);
                        zf[128] = mid_conj2(ek);
                        if constexpr (ECHO) zm[128] = mid_conj2(hk);
                    }
                }
            }
        }
        if constexpr (kMidTapParallel && !kMidAhead) {
            if (warp == NW - 1) {             // self-mirrored bin 128, one lane per tap, whole chunk
                // butterfly over the (power-of-two padded) tap lanes only: log2(P) steps
                constexpr int kTapLanes = P <= 8 ? 8 : P <= 16 ? 16 : 32;
                auto wsum = [](float v) {
#pragma unroll
                    for (int o = kTapLanes / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    return v;
                };
                const bool tap = lane < P;    // lanes beyond the filter length carry zeros
                float2 w = tap ? midW[lane] : make_float2(0.f, 0.f);
                float2 xp = tap ? midX[lane] : make_float2(0.f, 0.f);     // this tap's spectrum one frame ago
                float cv = (ALGO == kAlgoKalman && tap) ? midC[lane] : 0.f;
                float psi = (ALGO == kAlgoKalman) ? *midPsi : 0.f;
#pragma unroll 1
                for (int tl = 0; tl < F; ++tl) {
                    if (t0 + tl < T) {
                        float2* zf = zbuf + (tl * 2 + 0) * kTilePitch;
                        float2* zm = zbuf + (tl * 2 + 1) * kTilePitch;
                        const float2 xn = mid_conj2(zf[128]), yn = mid_conj2(zm[128]);
                        // tap p sees the spectrum tap p-1 saw one frame ago
                        float2 x = make_float2(__shfl_up_sync(0xffffffffu, xp.x, 1), __shfl_up_sync(0xffffffffu, xp.y, 1));
                        if (lane == 0) x = xn;
                        if (!tap) x = make_float2(0.f, 0.f);
                        const float2 prod = cmul(w, x);
                        const float2 yh = make_float2(wsum(prod.x), wsum(prod.y));
                        const float2 e = csub(yn, yh);
                        const float x2 = fmaf(x.x, x.x, x.y * x.y);
                        if constexpr (ALGO == kAlgoNlms) {
                            const float g = prm.mu * rcp_fast(wsum(x2) + prm.delta);
                            w = cfmac(x, make_float2(g * e.x, g * e.y), w);
                        } else {
                            psi = fmaf(prm.klam, psi, prm.koml * fmaf(e.x, e.x, e.y * e.y));
                            const float rd = __frcp_rn(wsum(cv * x2) + psi + prm.keps);
                            const float gs = cv * rd;
                            w = cfma(make_float2(gs * x.x, -gs * x.y), e, w);
                            w = make_float2(prm.ka * w.x, prm.ka * w.y);
                            cv = fmaf(prm.ka2 * (1.f - gs * x2), cv, prm.kq * fmaf(w.x, w.x, w.y * w.y));
                        }
                        xp = x;
                        if (lane == 0) {
                            zf[128] = mid_conj2(e);
                            if constexpr (ECHO) zm[128] = mid_conj2(yh);
                        }
                    }
                }
                if (tap) {
                    midW[lane] = w;
                    midX[lane] = xp;
                    if constexpr (ALGO == kAlgoKalman) midC[lane] = cv;
                }
                if constexpr (ALGO == kAlgoKalman) if (lane == 0) *midPsi = psi;
            }
        }
        AEC_TICK(5);                         // phase B
        __syncthreads();
        AEC_TICK(6);                         // barrier after B

        // ================= phase C : synthesis + overlap-add =================
        const int tl = 2 * warp + half;      // this half-warp's frame inside the chunk
        const int t = t0 + tl;
        const long long out_off = row_off(prm.out_stride);
        float* out_b[2] = {prm.err + out_off, ECHO ? prm.echo + out_off : nullptr};
        float2 head[NSIG][8];
        const bool synth_warp = (2 * warp < F);      // F/2 frame pairs per chunk; surplus warps sit out
        if (synth_warp)
#pragma unroll
        for (int sgn = 0; sgn < NSIG; ++sgn) {
            float2 v[16];
            float2* tile = zbuf + (tl * 2 + sgn) * kTilePitch;
            if (t < T) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = tile[h + 16 * j];
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = make_float2(0.f, 0.f);
            }
            __syncwarp();
            fft256_halfwarp_regs<true>(v, tile, twr, h);
            // register position p holds z[m], m = h + 16 r, r = fft16_index(p): samples 2m, 2m+1
            float2 u[16];                    // u[r], windowed + normalised
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int r = fft16_index(p);
                const float2 w = win_s[h + 16 * r];
                u[r] = make_float2(v[p].x * w.x, v[p].y * w.y);
            }
            // in-warp overlap: block t_lo + 1 = second half of the lower frame + first half of the upper
            // tail of the upper frame -> the (now dead) tile that held its spectrum
            float2* tail_dst = zbuf + ((2 * warp + 1) * 2 + sgn) * kTilePitch;
            // every condition of the store loop is decided once per chunk, so that the loop body is
            // predicated straight-line code (it used to carry a branch nest per value)
            const bool lower = (half == 0);
            const bool st_ok = lower && (t + 1 <= T - 1);            // output block t + 1, hop index t
            const bool en_ok = st_ok && sgn == 0 && t >= prm.erle_skip_hops;
            const bool vec = prm.vec_out != 0;
            float* dst = out_b[sgn] + (long long)t * 256 + 2 * h;
            float en = 0.f;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float ux = __shfl_down_sync(0xffffffffu, u[r].x, 16);
                const float uy = __shfl_down_sync(0xffffffffu, u[r].y, 16);
                const float2 o = make_float2(u[8 + r].x + ux, u[8 + r].y + uy);
                if (st_ok && vec) st_stream_f2(dst + 32 * r, o);
                if (st_ok && !vec) { st_stream_f1(dst + 32 * r, o.x); st_stream_f1(dst + 32 * r + 1, o.y); }
                en = fmaf(o.x, o.x, fmaf(o.y, o.y, en));
                if constexpr (FEAT) {                  // output block t + 1 of the error signal stays on chip for the re-analysis
                    if (lower && sgn == 0)
                        *reinterpret_cast<float2*>(f_ring + ((t + 1) % (F + 1)) * 256 + 2 * h + 32 * r) = o;
                }
                if (!lower) tail_dst[h + 16 * r] = u[8 + r];
                // unconditional: a predicated assignment would make the old value loop-carried
                // (16 registers live across the whole chunk loop -> spills)
                head[sgn][r] = u[r];
            }
            if (en_ok) acc_e += en;
        }
        if constexpr (kMidAhead) {
            // the warps without synthesis work run bin 128 of the next chunk (see mid_ahead)
            if (!synth_warp && t0 + F < T) mid_ahead(t0 + F, static_cast<uint32_t>((c + 1) & 1));
        }
        AEC_TICK(7);                         // phase C (inverse FFT + in-warp overlap)
        __syncthreads();
        AEC_TICK(8);                         // barrier after C
        // cross-warp overlap: block t (lower frame) = predecessor's tail + this frame's first half
        if (synth_warp && half == 0 && t >= 1 && t <= T - 1) {
#pragma unroll
            for (int sgn = 0; sgn < NSIG; ++sgn) {
                const float2* tail_src =
                    (warp == 0) ? carry + sgn * 128 : zbuf + ((2 * warp - 1) * 2 + sgn) * kTilePitch;
                const bool vec = prm.vec_out != 0;
                float* dst = out_b[sgn] + (long long)(t - 1) * 256 + 2 * h;
                float en = 0.f;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float2 tl2 = tail_src[h + 16 * r];
                    const float2 o = make_float2(head[sgn][r].x + tl2.x, head[sgn][r].y + tl2.y);
                    if (vec) st_stream_f2(dst + 32 * r, o);
                    else { st_stream_f1(dst + 32 * r, o.x); st_stream_f1(dst + 32 * r + 1, o.y); }
                    en = fmaf(o.x, o.x, fmaf(o.y, o.y, en));
                    if constexpr (FEAT) {              // output block t
                        if (sgn == 0) *reinterpret_cast<float2*>(f_ring + (t % (F + 1)) * 256 + 2 * h + 32 * r) = o;
                    }
                }
                if (sgn == 0 && t - 1 >= prm.erle_skip_hops) acc_e += en;
            }
        }
        // the last warp's tail crosses into the next chunk: warp 0 (which has just consumed the old
        // carry) moves it out of Zbuf before the next analysis phase overwrites the tile.
        if (warp == 0 && t0 + F < T) {
            __syncwarp();
#pragma unroll
            for (int sgn = 0; sgn < NSIG; ++sgn) {
                const float2* src = zbuf + ((F - 1) * 2 + sgn) * kTilePitch;
#pragma unroll
                for (int i = 0; i < 4; ++i) carry[sgn * 128 + lane + 32 * i] = src[lane + 32 * i];
            }
        }
        if constexpr (FEAT) {
            __syncthreads();                 // the chunk's error blocks are complete, every spectrum tile is dead
            feat_epilogue(t0);
        }
    }
    if constexpr (FEAT) {
        // frames the one-frame lag left behind (at most F of them: the last chunk emitted up to its third frame) ...
        const int t_last = ((T - 1) / F) * F;
        __syncthreads();
        feat_epilogue(t_last + F);
        // ... and rows beyond this utterance's frames: what the front end gives for silence (ragged batches)
        float* fb = prm.feat + (static_cast<long long>(blockIdx.x) * prm.feat_frames) * 64;
        for (long long i = static_cast<long long>(T) * 64 + tid; i < prm.feat_frames * 64; i += NT)
            fb[i] = (i & 32) ? 0.f : f_csum[i & 31];
    }

#ifdef AEC_PHASE_TIMING
    AEC_TICK(9);
    if (lane == 0 && prm.dbg) {
        unsigned hw_warp, hw_sm;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw_warp));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(hw_sm));
        dbg_sm[warp][10] = (long long)hw_sm * 1000 + hw_warp;      // placement of this warp
        for (int i = 0; i < 12; ++i) prm.dbg[((long long)blockIdx.x * NW + warp) * 12 + i] = dbg_sm[warp][i];
    }
#endif
    // ---- epilogue: zero the output beyond (T-1)*256, ERLE ------------------------------------
    {
        const long long valid = (long long)(T - 1) * 256;
        float* out_b[2] = {prm.err + row_off(prm.out_stride), ECHO ? prm.echo + row_off(prm.out_stride) : nullptr};
        for (long long i = valid + tid; i < prm.out_stride && i < prm.L; i += NT) {
            out_b[0][i] = 0.f;
            if constexpr (ECHO) out_b[1][i] = 0.f;
        }
    }
    if (prm.erle_db != nullptr) {
        __syncthreads();
        float* red = reinterpret_cast<float*>(zbuf);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) acc_e += __shfl_xor_sync(0xffffffffu, acc_e, o);   // within each half-warp
        if (h == 0) red[2 * warp + (half ^ 1)] = acc_e;      // [2w] microphone (lane 16), [2w+1] error (lane 0)
        __syncthreads();
        if (tid == 0) {
            float pm = 0.f, pe = 0.f;
            for (int w = 0; w < NW; ++w) {
                pm += red[2 * w];
                pe += red[2 * w + 1];
            }
            prm.erle_db[blockIdx.x] = 10.f * log10f(fmaxf(pm, 1e-20f) / fmaxf(pe, 1e-20f));
        }
    }
}

}  // namespace aec
