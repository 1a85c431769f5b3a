// ubench.cu -- issue-rate microbenchmarks that size the kernels:
//   FADD / FMUL / packed add.rn.f32x2 / mul.rn.f32x2, IMAD.WIDE (+funnel shift) chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

__global__ void k_fadd(float *out, float a) {
    float v[ILP];
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = v[i] + a;
    }
    float s = 0;
    for (int i = 0; i < ILP; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fmul_fadd(float *out, float a, float b) {
    float v[ILP];
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = v[i] * a + b;  // -fmad=false: FMUL + FADD
    }
    float s = 0;
    for (int i = 0; i < ILP; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd2(float *out, float a) {
    unsigned long long v[ILP], aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    for (int i = 0; i < ILP; i++) {
        float x = threadIdx.x * 0.001f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(x));
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(aa));
    }
    float s = 0;
    for (int i = 0; i < ILP; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fmul2_fadd2(float *out, float a, float b) {
    unsigned long long v[ILP], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    for (int i = 0; i < ILP; i++) {
        float x = threadIdx.x * 0.001f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(v[i]) : "f"(x));
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(aa));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(bb));
        }
    }
    float s = 0;
    for (int i = 0; i < ILP; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// biquad-like: 5 IMAD.WIDE + funnel shift per sample, ILP independent lanes per thread
__global__ void k_imad_wide(int *out, int b0, int b1, int b2, int a1, int a2) {
    int x1[ILP], x2[ILP], y1[ILP], y2[ILP];
    for (int i = 0; i < ILP; i++) { x1[i] = threadIdx.x + i; x2[i] = i; y1[i] = 3 * i; y2[i] = 7; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            int x0 = it + i;
            unsigned long long acc = (unsigned long long)((long long)b0 * x0) + (unsigned long long)((long long)b1 * x1[i]) +
                                     (unsigned long long)((long long)b2 * x2[i]) + (unsigned long long)((long long)a1 * y1[i]) +
                                     (unsigned long long)((long long)a2 * y2[i]);
            int y0 = (int)__funnelshift_r((unsigned)acc, (unsigned)(acc >> 32), 30);
            x2[i] = x1[i]; x1[i] = x0; y2[i] = y1[i]; y1[i] = y0;
        }
    }
    int s = 0;
    for (int i = 0; i < ILP; i++) s += y1[i] + y2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; i++) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    float *o; cudaMalloc(&o, sizeof(float) * sms * 8 * 1024);
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
        double n = (double)blocks * threads * ITERS * ILP;
        float t;
        t = timeit([&] { k_fadd<<<blocks, threads>>>(o, 1.0001f); });
        printf("warps/SM %2d  FADD        %8.1f Gop/s (thread-ops)  %.2f op/clk/SM @1.9GHz\n", warps_per_sm, n / t / 1e6, n / t / 1e6 / sms / 1.9);
        t = timeit([&] { k_fmul_fadd<<<blocks, threads>>>(o, 1.0001f, 0.5f); });
        printf("             FMUL+FADD   %8.1f Gop/s\n", 2 * n / t / 1e6);
        t = timeit([&] { k_fadd2<<<blocks, threads>>>(o, 1.0001f); });
        printf("             FADD2       %8.1f Gop/s (scalar flops, 2 per instr)\n", 2 * n / t / 1e6);
        t = timeit([&] { k_fmul2_fadd2<<<blocks, threads>>>(o, 1.0001f, 0.5f); });
        printf("             FMUL2+FADD2 %8.1f Gop/s (scalar flops)\n", 4 * n / t / 1e6);
        t = timeit([&] { k_imad_wide<<<blocks, threads>>>((int *)o, 1 << 28, 1 << 29, 1 << 28, 1227265970, -443242341); });
        printf("             biquad-i32  %8.1f GSa/s (5 IMAD.WIDE + SHF per sample)\n", n / t / 1e6);
    }
    return 0;
}
