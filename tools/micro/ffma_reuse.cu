// Developer microbenchmark (round 2): does the operand-reuse cache lift three-register FFMAs above the 0.64
// warp-instructions / cycle / scheduler that tools/micro/fp32_pipes.cu measured for FFMAs whose three sources are
// all fresh registers?  The kernels replay the complex multiply-accumulate sequences of the stage-1 filter phase
// (stage1_kernel.cuh bin_step: Yhat += W_p X_p and W_p += conj(X_p) gE), which share W.x / W.y / X.x / X.y between
// neighbouring FFMAs -- exactly where ptxas sets `.reuse` -- and an SGEMM-style outer product as the best case.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ab/ffma_reuse tools/micro/ffma_reuse.cu && ab/ffma_reuse
//   cuobjdump -sass ab/ffma_reuse | grep FFMA     (the .reuse flags of each variant)
#include <cstdio>
#include <cuda_runtime.h>

struct c2 { float x, y; };

// MODE 0: echo-estimate order of the kernel:   acc += W_p * X_p, p = 0..3, as cfma() writes it
//         (fma(-a.y, b.y, fma(a.x, b.x, acc.x)), fma(a.y, b.x, fma(a.x, b.y, acc.y)))
// MODE 1: same products ordered so that consecutive FFMAs share one multiplicand in the same operand slot:
//         (W.x X.x -> ax), (W.x X.y -> ay), (W.y X.y -> ax), (W.y X.x -> ay)
// MODE 2: 4 x 4 real outer product  acc[i][j] += a[i] * b[j]   (SGEMM inner step: every a[i] reused four times)
// MODE 3: weight update order of the kernel:   W_p += conj(X_p) * ge   (W changes, X and ge fixed per frame)
// MODE 4: three fresh registers per FFMA, nothing to reuse (the old worst case, for reference in the same harness)
template <int MODE>
__global__ void __launch_bounds__(128) probe(float* out, const float* in, int iters) {
    // per-thread state as in the two-warp kernel: four bins, four taps each, W and X in registers
    c2 W[4][4], X[4][4], acc[4], ge[4];
    float a[4], b[4], m[16];
    const float t = in[threadIdx.x];               // thread-dependent: nothing lands in uniform registers
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            W[k][p] = {t * (4 * k + p + 1), t * (4 * k + p + 2) + 1.f};
            X[k][p] = {t * (4 * k + p + 3) - 1.f, t * (4 * k + p + 5)};
        }
        acc[k] = {0.f, 0.f};
        ge[k] = {t * 0.5f + k, t * 0.25f - k};
        a[k] = t * (k + 7);
        b[k] = t * (k + 11) + 2.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = t * i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    acc[k].x = fmaf(-W[k][p].y, X[k][p].y, fmaf(W[k][p].x, X[k][p].x, acc[k].x));
                    acc[k].y = fmaf(W[k][p].y, X[k][p].x, fmaf(W[k][p].x, X[k][p].y, acc[k].y));
                }
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {      // (bins interleaved: four independent chains between dependent FFMAs)
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[k].x) : "f"(W[k][p].x), "f"(X[k][p].x));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[k].y) : "f"(W[k][p].x), "f"(X[k][p].y));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[k].x) : "f"(-W[k][p].y), "f"(X[k][p].y));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[k].y) : "f"(W[k][p].y), "f"(X[k][p].x));
                }
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) m[4 * i + j] = fmaf(a[i], b[j], m[4 * i + j]);
        } else if (MODE == 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    W[k][p].x = fmaf(X[k][p].y, ge[k].y, fmaf(X[k][p].x, ge[k].x, W[k][p].x));
                    W[k][p].y = fmaf(-X[k][p].y, ge[k].x, fmaf(X[k][p].x, ge[k].y, W[k][p].y));
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(m[(i + 5) & 15]), "f"(m[(i + 10) & 15]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s += acc[k].x + acc[k].y;
#pragma unroll
        for (int p = 0; p < 4; ++p) s += W[k][p].x + W[k][p].y;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += m[i];
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, int sms, double clock_ghz, int ffma_per_iter) {
    float *d, *in;
    cudaMalloc(&d, 256);
    cudaMalloc(&in, 128 * sizeof(float));
    float hin[128];
    for (int i = 0; i < 128; ++i) hin[i] = 1e-3f * (i + 1);
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int occ : {2, 4}) {                      // 128-thread blocks per SM: 2 / 4 warps per scheduler
        const int blocks = sms * occ;
        probe<MODE><<<blocks, 128>>>(d, in, iters);
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            probe<MODE><<<blocks, 128>>>(d, in, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double warp_inst = (double)ffma_per_iter * iters * 4.0 * blocks;       // 4 warps per block
        const double per_sched = warp_inst / (best * 1e-3) / (clock_ghz * 1e9) / sms / 4.0;
        printf("%-46s warps/scheduler %d  %.3f ms  %.3f FFMA warp-instr / cycle / scheduler\n", name, occ, best, per_sched);
    }
    cudaFree(d);
    cudaFree(in);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    run<0>("echo estimate, kernel order (cfma)", p.multiProcessorCount, ghz, 64);
    run<1>("echo estimate, reuse-paired order", p.multiProcessorCount, ghz, 64);
    run<2>("4x4 outer product (SGEMM step)", p.multiProcessorCount, ghz, 64);
    run<3>("weight update, kernel order (cfmac)", p.multiProcessorCount, ghz, 64);
    run<4>("three fresh registers (no reuse possible)", p.multiProcessorCount, ghz, 64);
    return 0;
}
