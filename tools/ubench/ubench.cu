// ubench.cu — pipe-throughput probes for the instruction mix of the walk (sm_100a).  Development tool:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench/ubench.cu && /tmp/ubench
// Each probe runs 8 independent dependency chains per thread, 148*8 blocks of 256 threads, and prints
// warp-instructions per cycle per SM (4 = every scheduler issues one of them every cycle).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
template <int OP>
__global__ void __launch_bounds__(256) probe(float* out, float a, float b, double da) {
    float v[8];
    double d[8];
    unsigned long long q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x + i; d[i] = threadIdx.x + i; q[i] = (unsigned long long)(threadIdx.x + i) * 0x9E3779B97F4A7C15ull; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) v[i] = fmaf(v[i], a, b);                                    // FFMA
            if (OP == 1) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i]));  // MUFU.RSQ
            if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(q[(i + 1) & 7]));   // FADD2
            if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q[i]) : "l"(q[(i + 1) & 7]));   // FFMA2
            if (OP == 4) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(v[i])); d[i] = t; v[i] += 1.0f; }   // F2F.F64.F32 (+FADD)
            if (OP == 5) d[i] = d[i] + da;                                           // DADD
            if (OP == 6) d[i] = fma(d[i], da, da);                                   // DFMA
            if (OP == 7) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(v[i])); d[i] += t; v[i] = v[i] * a; }   // F2F + DADD + FMUL
            if (OP == 8) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i])); double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(v[i])); d[i] += t; }   // MUFU + F2F + DADD
            if (OP == 9) v[i] = v[i] + a;                                            // FADD
        }
    }
    float s = 0.f;
    double ds = 0.0;
    unsigned long long qs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += v[i]; ds += d[i]; qs ^= q[i]; }
    if (s == 12345.678f && ds == 1.0 && qs == 3) out[0] = s;
}

template <int OP>
void run(const char* name, int instr_per_iter, int sms) {
    float* d;
    cudaMalloc(&d, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = sms * 8;
    probe<OP><<<blocks, 256>>>(d, 1.0001f, 0.5f, 1.000001);
    cudaEventRecord(a);
    probe<OP><<<blocks, 256>>>(d, 1.0001f, 0.5f, 1.000001);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_instr = (double)blocks * 8 /*warps*/ * ITERS * 8 * instr_per_iter;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, warp_instr / cycles / sms, clk / 1000);
    cudaFree(d);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("FFMA", 1, sms);
    run<9>("FADD", 1, sms);
    run<1>("MUFU.RSQ", 1, sms);
    run<2>("FADD2", 1, sms);
    run<3>("FFMA2", 1, sms);
    run<4>("F2F.F64.F32 + FADD", 2, sms);
    run<5>("DADD", 1, sms);
    run<6>("DFMA", 1, sms);
    run<7>("F2F + DADD + FMUL", 3, sms);
    run<8>("MUFU + F2F + DADD", 3, sms);
    return cudaGetLastError() != cudaSuccess;
}
