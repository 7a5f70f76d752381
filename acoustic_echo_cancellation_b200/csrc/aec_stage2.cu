// Stage-2 residual-echo suppressor inference (the consumer of the stage-1 output):
//   aec_stage2_mask   <-> Little_net.forward lines Stage2_lhm/scripts/network/ERB.py:293-304
//                         (GRU 64->32, Linear 64->32 + ReLU, Linear 32->32 + sigmoid, mask * mic_erb)
//   aec_stage2_synth  <-> ERB.py:306-316 (est_erb @ erb^T applied to the microphone spectrum, iSTFT, + 1e-9)
// The feature tensor fed to the GRU comes from aec_features (ERB.py:262-290).  Pinned by
// tests/golden/reference_stage2.npz (the reference module itself, seeded weights).
#include <cstdint>
#include <type_traits>

#include "aec_common.cuh"
#include "fft_warp.cuh"

namespace aec {
namespace {

constexpr int kH = 32;          // hidden units == ERB bands (Little_net: GRU(2*bands -> bands))
constexpr int kIn = 64;         // 2 * bands
// utterances (= warps) per CTA: 4 for large batches (the 45 KB of weights are staged once per CTA), 1 for batches that would
// otherwise leave SMs empty -- the recurrence streams ~400 shared-memory wavefronts per utterance per frame, and four
// utterances on one SM pay them one after the other while other SMs sit idle

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

// One warp per utterance, lane j = hidden unit / band j.  Weights are staged transposed in shared
// memory (groups of four inputs per unit, 16-byte loads); the frame's feature
// vector and the hidden state are broadcast through a per-warp scratch.
template <int kMaskWarps>
__global__ void __launch_bounds__(kMaskWarps * 32) stage2_mask_kernel(const float* __restrict__ feat,
                                                                      aec_stage2_weights w, float* __restrict__ est,
                                                                      long long B, long long T) {
    extern __shared__ __align__(16) float sm[];
    float* wih = sm;                       // [64][96]
    float* whh = wih + kIn * 96;           // [32][96]
    float* w1 = whh + kH * 96;             // [64][32]
    float* w2 = w1 + kIn * kH;             // [32][32]
    float* bias = w2 + kH * kH;            // b_ih[96] b_hh[96] b1[32] b2[32]
    float* scratch = bias + 256;           // per warp: x[64] h[32] o2[32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // weights as float4 groups of four consecutive inputs per (gate, unit): one 16-byte load feeds four FFMAs
    //   wih : [gate 3][input/4 16][unit 32][4]     whh : [gate 3][input/4 8][unit 32][4]
    //   w1  : [input/4 16][unit 32][4]             w2  : [input/4 8][unit 32][4]
    for (int i = tid; i < 96 * kIn; i += blockDim.x) {
        const int row = i / kIn, col = i % kIn, g = row >> 5, j = row & 31;
        wih[(((g * (kIn / 4) + (col >> 2)) * 32 + j) << 2) + (col & 3)] = __ldg(w.gru_w_ih + i);
    }
    for (int i = tid; i < 96 * kH; i += blockDim.x) {
        const int row = i / kH, col = i % kH, g = row >> 5, j = row & 31;
        whh[(((g * (kH / 4) + (col >> 2)) * 32 + j) << 2) + (col & 3)] = __ldg(w.gru_w_hh + i);
    }
    for (int i = tid; i < kH * kIn; i += blockDim.x) {
        const int j = i / kIn, col = i % kIn;
        w1[((((col >> 2)) * 32 + j) << 2) + (col & 3)] = __ldg(w.lin1_w + i);
    }
    for (int i = tid; i < kH * kH; i += blockDim.x) {
        const int j = i / kH, col = i % kH;
        w2[((((col >> 2)) * 32 + j) << 2) + (col & 3)] = __ldg(w.lin2_w + i);
    }
    for (int i = tid; i < 96; i += blockDim.x) {
        bias[i] = __ldg(w.gru_b_ih + i);
        bias[96 + i] = __ldg(w.gru_b_hh + i);
    }
    for (int i = tid; i < kH; i += blockDim.x) {
        bias[192 + i] = __ldg(w.lin1_b + i);
        bias[224 + i] = __ldg(w.lin2_b + i);
    }
    __syncthreads();
    const long long b = (long long)blockIdx.x * kMaskWarps + warp;
    if (b >= B) return;
    float* xs = scratch + warp * 128;
    float* hs = xs + 64;
    float* os = xs + 96;
    const float* fb = feat + b * T * kIn;
    float* eb = est + b * T * kH;
    const float b_r = bias[lane] + bias[96 + lane], b_z = bias[32 + lane] + bias[128 + lane];
    const float b_in = bias[64 + lane], b_hn = bias[160 + lane];
    const float b_1 = bias[192 + lane], b_2 = bias[224 + lane];
    float h = 0.f;
    hs[lane] = 0.f;
    float x0 = T > 0 ? __ldg(fb + lane) : 0.f, x1 = T > 0 ? __ldg(fb + 32 + lane) : 0.f;
    for (long long t = 0; t < T; ++t) {
        __syncwarp();
        xs[lane] = x0;
        xs[32 + lane] = x1;
        const float merb = x0;                       // lane j: mic_erb[j] (first half of the feature vector)
        __syncwarp();
        if (t + 1 < T) {                             // prefetch the next frame's features
            x0 = __ldg(fb + (t + 1) * kIn + lane);
            x1 = __ldg(fb + (t + 1) * kIn + 32 + lane);
        }
        // four partial sums per gate (one per position inside a group of four inputs): the recurrence is latency-bound
        // on one warp, and a single accumulator per gate made every step a chain of 96 dependent FFMAs
        float ar[4] = {b_r, 0.f, 0.f, 0.f}, az[4] = {b_z, 0.f, 0.f, 0.f}, ain[4] = {b_in, 0.f, 0.f, 0.f},
              ahn[4] = {b_hn, 0.f, 0.f, 0.f};
        const float4* wih4 = reinterpret_cast<const float4*>(wih) + lane;
        const float4* whh4 = reinterpret_cast<const float4*>(whh) + lane;
#pragma unroll 4
        for (int i4 = 0; i4 < kIn / 4; ++i4) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + 4 * i4);
            const float4 wr = wih4[(0 * (kIn / 4) + i4) * 32], wz = wih4[(1 * (kIn / 4) + i4) * 32], wn = wih4[(2 * (kIn / 4) + i4) * 32];
            ar[0] = fmaf(wr.x, xv.x, ar[0]); ar[1] = fmaf(wr.y, xv.y, ar[1]); ar[2] = fmaf(wr.z, xv.z, ar[2]); ar[3] = fmaf(wr.w, xv.w, ar[3]);
            az[0] = fmaf(wz.x, xv.x, az[0]); az[1] = fmaf(wz.y, xv.y, az[1]); az[2] = fmaf(wz.z, xv.z, az[2]); az[3] = fmaf(wz.w, xv.w, az[3]);
            ain[0] = fmaf(wn.x, xv.x, ain[0]); ain[1] = fmaf(wn.y, xv.y, ain[1]); ain[2] = fmaf(wn.z, xv.z, ain[2]); ain[3] = fmaf(wn.w, xv.w, ain[3]);
        }
#pragma unroll 4
        for (int i4 = 0; i4 < kH / 4; ++i4) {
            const float4 hv = *reinterpret_cast<const float4*>(hs + 4 * i4);
            const float4 wr = whh4[(0 * (kH / 4) + i4) * 32], wz = whh4[(1 * (kH / 4) + i4) * 32], wn = whh4[(2 * (kH / 4) + i4) * 32];
            ar[0] = fmaf(wr.x, hv.x, ar[0]); ar[1] = fmaf(wr.y, hv.y, ar[1]); ar[2] = fmaf(wr.z, hv.z, ar[2]); ar[3] = fmaf(wr.w, hv.w, ar[3]);
            az[0] = fmaf(wz.x, hv.x, az[0]); az[1] = fmaf(wz.y, hv.y, az[1]); az[2] = fmaf(wz.z, hv.z, az[2]); az[3] = fmaf(wz.w, hv.w, az[3]);
            ahn[0] = fmaf(wn.x, hv.x, ahn[0]); ahn[1] = fmaf(wn.y, hv.y, ahn[1]); ahn[2] = fmaf(wn.z, hv.z, ahn[2]); ahn[3] = fmaf(wn.w, hv.w, ahn[3]);
        }
        const float r = sigmoidf_((ar[0] + ar[1]) + (ar[2] + ar[3])), z = sigmoidf_((az[0] + az[1]) + (az[2] + az[3]));
        const float n = tanhf(fmaf(r, (ahn[0] + ahn[1]) + (ahn[2] + ahn[3]), (ain[0] + ain[1]) + (ain[2] + ain[3])));
        h = fmaf(z, h - n, n);                       // (1 - z) n + z h
        __syncwarp();
        hs[lane] = h;
        __syncwarp();
        float a1[4] = {b_1, 0.f, 0.f, 0.f}, a1x[4] = {0.f, 0.f, 0.f, 0.f};   // linear1 on cat[h, mic_erb]  (ERB.py:295-298)
        const float4* w14 = reinterpret_cast<const float4*>(w1) + lane;
#pragma unroll 4
        for (int i4 = 0; i4 < kH / 4; ++i4) {
            const float4 hv = *reinterpret_cast<const float4*>(hs + 4 * i4);
            const float4 xv = *reinterpret_cast<const float4*>(xs + 4 * i4);
            const float4 wa = w14[i4 * 32], wb = w14[(kH / 4 + i4) * 32];      // inputs 0..31 = h, 32..63 = mic_erb
            a1[0] = fmaf(wa.x, hv.x, a1[0]); a1[1] = fmaf(wa.y, hv.y, a1[1]); a1[2] = fmaf(wa.z, hv.z, a1[2]); a1[3] = fmaf(wa.w, hv.w, a1[3]);
            a1x[0] = fmaf(wb.x, xv.x, a1x[0]); a1x[1] = fmaf(wb.y, xv.y, a1x[1]); a1x[2] = fmaf(wb.z, xv.z, a1x[2]); a1x[3] = fmaf(wb.w, xv.w, a1x[3]);
        }
        os[lane] = fmaxf(((a1[0] + a1[1]) + (a1[2] + a1[3])) + ((a1x[0] + a1x[1]) + (a1x[2] + a1x[3])), 0.f);
        __syncwarp();
        float a2[4] = {b_2, 0.f, 0.f, 0.f};          // linear2 + sigmoid  (ERB.py:301)
        const float4* w24 = reinterpret_cast<const float4*>(w2) + lane;
#pragma unroll 4
        for (int i4 = 0; i4 < kH / 4; ++i4) {
            const float4 ov = *reinterpret_cast<const float4*>(os + 4 * i4);
            const float4 wv = w24[i4 * 32];
            a2[0] = fmaf(wv.x, ov.x, a2[0]); a2[1] = fmaf(wv.y, ov.y, a2[1]); a2[2] = fmaf(wv.z, ov.z, a2[2]); a2[3] = fmaf(wv.w, ov.w, a2[3]);
        }
        eb[t * kH + lane] = sigmoidf_((a2[0] + a2[1]) + (a2[2] + a2[3])) * merb;    // est_erb = mask * mic_erb  (ERB.py:304)
    }
}

// ---- synthesis: out = iSTFT((est_erb @ erb^T) * STFT(mic - shift)) + 1e-9 -------------------
constexpr int kTT = 16, kThreads = 128, kK = 257, kPitch = kTT + 1;

__device__ __forceinline__ void unpack_pair_2(float2 fa, float2 fb, float2 w, float2& xk, float2& xm) {
    const float2 a = make_float2(fa.x + fb.x, fa.y - fb.y);
    const float2 d = make_float2(fa.y + fb.y, fb.x - fa.x);
    const float2 t = cmul(w, d);
    xk = make_float2(a.x + t.x, a.y + t.y);
    xm = make_float2(a.x - t.x, t.y - a.y);
}
__device__ __forceinline__ void pack_pair_2(float2 ek, float2 em, float2 w, float2& gk, float2& gm) {
    const float2 a = make_float2(ek.x + em.x, ek.y - em.y);
    const float2 d = make_float2(ek.x - em.x, ek.y + em.y);
    const float2 t = cmulc(d, w);
    gk = make_float2(a.x - t.y, a.y + t.x);
    gm = make_float2(a.x + t.y, t.x - a.y);
}

__global__ void __launch_bounds__(kThreads) stage2_synth_kernel(const float* __restrict__ mic,
                                                                const float* __restrict__ est,
                                                                const float* __restrict__ erb, float* __restrict__ y,
                                                                long long L, long long in_stride, long long out_stride,
                                                                long long T, float shift_val,
                                                                const float* __restrict__ shift_dev, Tables tab) {
    extern __shared__ __align__(16) unsigned char smem[];
    float2* tiles = reinterpret_cast<float2*>(smem);                                   // [8][kTilePitch]
    float2* fr = tiles + 8 * kTilePitch;                                               // [kTT][256]
    float* erbs = reinterpret_cast<float*>(fr + kTT * 256);                            // [257][32]
    float* es = erbs + kK * kH;                                                        // [kTT][32]
    int* blo = reinterpret_cast<int*>(es + kTT * kH);                                  // [257]
    int* bhi = blo + kK;                                                               // [257]
    const int tid = threadIdx.x, lane = tid & 31, hw = tid >> 4, h = lane & 15;
    const long long b = blockIdx.y, g0 = (long long)blockIdx.x * (kTT - 1);
    const float shift = shift_dev ? __ldg(shift_dev) : shift_val;      // ERB.py:254, by value or from aec_batch_shift
    for (int i = tid; i < kK * kH; i += kThreads) erbs[i] = __ldg(erb + i);
    for (int i = tid; i < kTT * kH; i += kThreads) {
        const long long t = g0 + i / kH;
        es[i] = t < T ? __ldg(est + (b * T + t) * kH + (i % kH)) : 0.f;
    }
    __syncthreads();
    for (int k = tid; k < kK; k += kThreads) {     // non-zero band range of bin k (the bank is ~2 bands per bin)
        int lo = kH, hi = 0;
        for (int j = 0; j < kH; ++j)
            if (erbs[k * kH + j] != 0.f) {
                lo = min(lo, j);
                hi = j + 1;
            }
        blo[k] = lo;
        bhi[k] = hi;
    }
    __syncthreads();
    const float* xb = mic + b * in_stride;
    float2* tile = tiles + hw * kTilePitch;
    auto gain = [&](int tt, int k) {               // ERB.py:306-307: (mask * mic_erb) @ erb^T
        float g = 0.f;
        for (int j = blo[k]; j < bhi[k]; ++j) g = fmaf(es[tt * kH + j], erbs[k * kH + j], g);
        return g;
    };
    for (int i = 0; i < 2; ++i) {
        const int tt = hw + 8 * i;
        const long long t = g0 + tt;
        // ---- analysis of the (shifted) microphone frame ----
        {
            float2 v[16];
            const long long base = (t - 1) * 256;
            const long long Lv = t < T ? L : 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const long long s = base + 2 * h + 32 * j;
                const float2 wv = __ldg(&tab.win_a[h + 16 * j]);
                const float x0 = (s >= 0 && s < Lv) ? __ldg(xb + s) - shift : 0.f;
                const float x1 = (s + 1 >= 0 && s + 1 < Lv) ? __ldg(xb + s + 1) - shift : 0.f;
                v[j] = make_float2(x0 * wv.x, x1 * wv.y);
            }
            __syncwarp();
            fft256_halfwarp<false>(v, tile, tab.tw256, h);
#pragma unroll
            for (int p = 0; p < 16; ++p) tile[h + 16 * fft16_index(p)] = v[p];
            __syncwarp();
        }
        // ---- per-bin gain, repack for the inverse transform (in place) ----
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = h + 16 * q, km = (256 - k) & 255;
            const float2 wk = __ldg(&tab.tw512[k]);
            float2 xk, xm, gk, gm;
            unpack_pair_2(tile[k], tile[km], wk, xk, xm);
            const float ga = gain(tt, k), gb = gain(tt, 256 - k);
            xk = make_float2(ga * xk.x, k == 0 ? 0.f : ga * xk.y);       // ERB.py:309-310
            xm = make_float2(gb * xm.x, k == 0 ? 0.f : gb * xm.y);
            pack_pair_2(xk, xm, wk, gk, gm);
            tile[k] = gk;
            tile[km] = gm;
        }
        if (h == 0) {
            float2 xk, xm, gk, gm;
            unpack_pair_2(tile[128], tile[128], make_float2(0.f, -1.f), xk, xm);
            const float ga = gain(tt, 128);
            xk = make_float2(ga * xk.x, ga * xk.y);
            pack_pair_2(xk, xk, make_float2(0.f, -1.f), gk, gm);
            tile[128] = gk;
        }
        __syncwarp();
        float2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = tile[h + 16 * j];
        __syncwarp();
        fft256_halfwarp<true>(v, tile, tab.tw256, h);
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int r = fft16_index(p);
            const float2 wv = __ldg(&tab.win_s[h + 16 * r]);
            fr[tt * 256 + h + 16 * r] = make_float2(v[p].x * wv.x, v[p].y * wv.y);
        }
        __syncwarp();
    }
    __syncthreads();
    float* yb = y + b * out_stride;
    for (int idx = tid; idx < (kTT - 1) * 128; idx += kThreads) {
        const int tt = idx / 128, m = idx % 128;
        const long long g = g0 + tt;
        if (g + 1 <= T - 1) {
            const float2 a = fr[tt * 256 + 128 + m];
            const float2 c = fr[(tt + 1) * 256 + m];
            yb[g * 256 + 2 * m] = a.x + c.x + 1e-9f;                     // ERB.py:316
            yb[g * 256 + 2 * m + 1] = a.y + c.y + 1e-9f;
        }
    }
}

}  // namespace
}  // namespace aec

using namespace aec;

extern "C" int aec_stage2_mask(const float* feat, const aec_stage2_weights* w, float* est_erb, int64_t B, int64_t T,
                               int32_t bands, void* cuda_stream) {
    if (B < 0 || T < 0 || !w) return AEC_EINVAL;
    if (bands != kH) return AEC_EUNSUPPORTED;        // Little_net is built for 32 ERB bands (configs.py:21-27)
    if (B == 0 || T == 0) return AEC_OK;
    if (!feat || !est_erb || !w->gru_w_ih || !w->gru_w_hh || !w->gru_b_ih || !w->gru_b_hh || !w->lin1_w ||
        !w->lin1_b || !w->lin2_w || !w->lin2_b)
        return AEC_EINVAL;
    Tables tab;
    int rc = get_tables(&tab);                        // device check (sm_100 only)
    if (rc != AEC_OK) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto launch = [&](auto warps_tag) -> int {
        constexpr int kW = decltype(warps_tag)::value;
        const size_t smem = (size_t)(kIn * 96 + kH * 96 + kIn * kH + kH * kH + 256 + kW * 128) * sizeof(float);
        AEC_CUDA_CHECK(cudaFuncSetAttribute(stage2_mask_kernel<kW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)((B + kW - 1) / kW);
        stage2_mask_kernel<kW><<<grid, kW * 32, smem, static_cast<cudaStream_t>(cuda_stream)>>>(feat, *w, est_erb, B, T);
        return AEC_OK;
    };
    rc = (B <= 4LL * sms) ? launch(std::integral_constant<int, 1>{}) : launch(std::integral_constant<int, 4>{});
    if (rc != AEC_OK) return rc;
    AEC_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return AEC_OK;
}

static int synth_impl(const float* mic, const float* est_erb, const float* erb, float* out, int64_t B, int64_t L,
                      int64_t in_stride, int64_t out_stride, int32_t frame, int32_t bands, float shift_mic,
                      const float* shift_mic_dev, void* cuda_stream) {
    if (B < 0 || L < 0 || in_stride < L) return AEC_EINVAL;
    if (frame != 512 || bands != kH) return AEC_EUNSUPPORTED;
    const long long T = aec_num_frames(L, frame);
    if (T >= 1 && out_stride < (T - 1) * 256) return AEC_EINVAL;
    if (B == 0 || T <= 1) return AEC_OK;
    if (!mic || !est_erb || !erb || !out) return AEC_EINVAL;
    if (B > 65535) {            // grid.y carries the utterance index: larger batches go in slices
        for (int64_t off = 0; off < B; off += 65535) {
            const int rc2 = synth_impl(mic + off * in_stride, est_erb + off * T * kH, erb, out + off * out_stride,
                                       (B - off < 65535) ? B - off : 65535, L, in_stride, out_stride, frame, bands,
                                       shift_mic, shift_mic_dev, cuda_stream);
            if (rc2 != AEC_OK) return rc2;
        }
        return AEC_OK;
    }
    Tables tab;
    int rc = get_tables(&tab);
    if (rc != AEC_OK) return rc;
    const size_t smem = (size_t)(8 * kTilePitch + kTT * 256) * sizeof(float2) + (size_t)(kK * kH + kTT * kH) * sizeof(float) +
                        (size_t)2 * kK * sizeof(int);
    AEC_CUDA_CHECK(cudaFuncSetAttribute(stage2_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((T - 1 + kTT - 2) / (kTT - 1)), (unsigned)B);
    stage2_synth_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(cuda_stream)>>>(
        mic, est_erb, erb, out, L, in_stride, out_stride, T, shift_mic, shift_mic_dev, tab);
    AEC_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return AEC_OK;
}

extern "C" int aec_stage2_synth(const float* mic, const float* est_erb, const float* erb, float* out, int64_t B,
                                int64_t L, int64_t in_stride, int64_t out_stride, int32_t frame, int32_t bands,
                                float shift_mic, void* cuda_stream) {
    return synth_impl(mic, est_erb, erb, out, B, L, in_stride, out_stride, frame, bands, shift_mic, nullptr, cuda_stream);
}

extern "C" int aec_stage2_synth_dev(const float* mic, const float* est_erb, const float* erb, float* out, int64_t B,
                                    int64_t L, int64_t in_stride, int64_t out_stride, int32_t frame, int32_t bands,
                                    const float* shift_mic_dev, void* cuda_stream) {
    return synth_impl(mic, est_erb, erb, out, B, L, in_stride, out_stride, frame, bands, 0.f, shift_mic_dev, cuda_stream);
}
