// Developer microbenchmark: issue rate of FFMA / FADD / FMUL / mixes on one B200 (per SM per cycle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ab/fp32_pipes tools/micro/fp32_pipes.cu && ab/fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            if (MODE == 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            if (MODE == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
            if (MODE == 3) {   // alternate add / fma
                if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
                else asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            }
            if (MODE == 4) {   // add written as fma(x, 1, b)
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(1.0f), "f"(b));
            }
            if (MODE == 5) {   // 2 adds : 1 fma : 1 mul (FFT-like mix)
                if ((i & 3) == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
                else if ((i & 3) == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
                else asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            }
            if (MODE == 6) {   // add with two register operands (x[i] += x[(i+1)&15]) -- different banks pattern
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(x[(i + 5) & 15]));
            }
            if (MODE == 7) {   // fma with three register operands
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(x[(i + 5) & 15]), "f"(x[(i + 9) & 15]));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s;
}

// packed f32x2: 8 register pairs per thread; MODE 0: fma2 pair,const,const  1: fma2 pair,pair,pair  2: add2 pair,pair
// 3: mul2 pair,pair   4: complex MAC the scalar way (4 FFMA, 3 regs each) for comparison, same data volume
template <int MODE>
__global__ void __launch_bounds__(256) probe2(float* out, int iters, float a, float b) {
    unsigned long long x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo = threadIdx.x * 1e-3f + i, hi = lo + 0.5f;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
    }
    unsigned long long ca, cb;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(ca) : "f"(a));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cb) : "f"(b));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(ca), "l"(cb));
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(x[(i + 3) & 7]), "l"(x[(i + 5) & 7]));
            if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(x[(i + 3) & 7]));
            if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(x[(i + 3) & 7]));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 12345ull) out[0] = 1.f;
}

template <int MODE>
void run2(const char* name, int sms, float clock_ghz, int occ) {
    float* d;
    cudaMalloc(&d, 256);
    const int iters = 4096, blocks = sms * occ;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe2<MODE><<<blocks, 256>>>(d, iters, 0.999f, 1e-3f);
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        probe2<MODE><<<blocks, 256>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warp_inst = 8.0 * iters * 8.0 * blocks;          // 8 warps per block
    const double per_sched = warp_inst / (best * 1e-3) / (clock_ghz * 1e9) / sms / 4.0;
    printf("%-28s blocks/SM %d  %.3f ms  %.2f packed warp-instr / cycle / scheduler (= %.2f scalar-equivalent)\n", name, occ, best,
           per_sched, 2 * per_sched);
    cudaFree(d);
}

// FFMA d=x[i], x[i+O1], x[i+O2], x[i]: does the register-bank pattern of the three sources matter?
template <int O1, int O2>
__global__ void __launch_bounds__(256) probe3(float* out, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(x[(i + O1) & 15]), "f"(x[(i + O2) & 15]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 12345.678f) out[0] = s;
}
template <int O1, int O2>
void run3(int sms, float clock_ghz) {
    float* d;
    cudaMalloc(&d, 256);
    const int iters = 4096, blocks = sms * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe3<O1, O2><<<blocks, 256>>>(d, iters);
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        probe3<O1, O2><<<blocks, 256>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double per = 16.0 * iters * 8.0 * blocks / (best * 1e-3) / (clock_ghz * 1e9) / sms / 4.0;
    printf("FFMA x[i], x[i+%d], x[i+%d]        %.2f warp-instr / cycle / scheduler\n", O1, O2, per);
    cudaFree(d);
}

template <int MODE>
void run(const char* name, int sms, float clock_ghz, int warps_per_sm_sched) {
    float* d;
    cudaMalloc(&d, 256);
    const int iters = 4096;
    const int blocks = sms * warps_per_sm_sched;       // one 8-warp block = 2 warps per scheduler
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(d, iters, 0.999f, 1e-3f);
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        probe<MODE><<<blocks, 256>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double thread_inst = 16.0 * iters * 256.0 * blocks;
    const double per_sm_per_cycle = thread_inst / (best * 1e-3) / (clock_ghz * 1e9) / sms;
    printf("%-28s blocks/SM %d  %.3f ms  %.1f lane-instr / cycle / SM  (%.2f warp-instr / cycle / scheduler)\n", name,
           warps_per_sm_sched, best, per_sm_per_cycle, per_sm_per_cycle / 128.0);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const float ghz = clk * 1e-6f;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    for (int occ : {1, 2, 4}) {
        run<0>("FFMA", p.multiProcessorCount, ghz, occ);
        run<1>("FADD", p.multiProcessorCount, ghz, occ);
        run<2>("FMUL", p.multiProcessorCount, ghz, occ);
        run<3>("FADD/FFMA alternating", p.multiProcessorCount, ghz, occ);
        run<4>("FFMA(x,1,b) as add", p.multiProcessorCount, ghz, occ);
        run<5>("2 FADD : 1 FFMA : 1 FMUL", p.multiProcessorCount, ghz, occ);
        run<6>("FADD reg,reg", p.multiProcessorCount, ghz, occ);
        run<7>("FFMA reg,reg,reg", p.multiProcessorCount, ghz, occ);
        run2<0>("FFMA2 pair,const,const", p.multiProcessorCount, ghz, occ);
        run2<1>("FFMA2 pair,pair,pair", p.multiProcessorCount, ghz, occ);
        run2<2>("FADD2 pair,pair", p.multiProcessorCount, ghz, occ);
        run2<3>("FMUL2 pair,pair", p.multiProcessorCount, ghz, occ);
    }
    run3<1, 2>(p.multiProcessorCount, ghz);
    run3<1, 3>(p.multiProcessorCount, ghz);
    run3<2, 4>(p.multiProcessorCount, ghz);
    run3<4, 8>(p.multiProcessorCount, ghz);
    run3<1, 1>(p.multiProcessorCount, ghz);
    run3<2, 2>(p.multiProcessorCount, ghz);
    run3<3, 6>(p.multiProcessorCount, ghz);
    run3<5, 10>(p.multiProcessorCount, ghz);
    run3<0, 1>(p.multiProcessorCount, ghz);
    run3<0, 0>(p.multiProcessorCount, ghz);
    return 0;
}
