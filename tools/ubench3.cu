// ubench3.cu -- integer multiply issue rates on sm_100a (what bounds Cascade<N>, the i64 biquad and the lock-in):
//   mad.wide.s32 with a 64-bit accumulator (IMAD.WIDE), mad.lo.s32 (IMAD), mul.hi.s32 (IMAD.HI), mul.wide + add.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3 tools/ubench3.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

__global__ void k_madwide(long long *out, int b) {
    long long acc[ILP];
    int a[ILP];
    for (int i = 0; i < ILP; i++) { acc[i] = threadIdx.x + i; a[i] = threadIdx.x * 3 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {  // mul.wide + add.s64 from the compiler: ptxas fuses them into IMAD.WIDE Rd, Ra, Ub, Rc
            acc[i] += (long long)a[i] * b;
            a[i] ^= (int)(acc[i] >> 32);
        }
    }
    long long s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mulwide(long long *out, int b) {  // product only (no accumulate operand)
    long long acc[ILP];
    int a[ILP];
    for (int i = 0; i < ILP; i++) { acc[i] = 0; a[i] = threadIdx.x * 3 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            asm volatile("mul.wide.s32 %0, %1, %2;" : "=l"(acc[i]) : "r"(a[i]), "r"(b));
            a[i] = (int)acc[i];
        }
    }
    long long s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_madlo(long long *out, int b) {
    int acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(acc[i]) : "r"(b));
    }
    long long s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mulhi(long long *out, int b) {
    int acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 77777 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(acc[i]) : "r"(b));
    }
    long long s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(long long *out, int b) {
    int acc[ILP];
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("add.s32 %0, %0, %1;" : "+r"(acc[i]) : "r"(b));
    }
    long long s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// IMAD.WIDE and IADD3 streams together (two pipes)
__global__ void k_mix(long long *out, int b) {
    long long acc[ILP];
    int a[ILP], c[ILP];
    for (int i = 0; i < ILP; i++) { acc[i] = threadIdx.x + i; a[i] = threadIdx.x * 3 + i; c[i] = i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            acc[i] += (long long)a[i] * b;
            a[i] ^= (int)(acc[i] >> 32);
            asm volatile("add.s32 %0, %0, %1;" : "+r"(c[i]) : "r"(b));
        }
    }
    long long s = 0;
    for (int i = 0; i < ILP; i++) s += acc[i] + c[i];
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
    printf("%s, %d SMs\n", p.name, sms);
    long long *o; cudaMalloc(&o, sizeof(long long) * sms * 8 * 1024);
    for (int warps_per_sm : {8, 16, 32}) {
        int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
        double n = (double)blocks * threads * ITERS * ILP;
        auto rep = [&](const char *name, float t, double per) {
            printf("warps/SM %2d  %-28s %9.1f Gop/s  %6.2f thread-op/clk/SM @1.9GHz\n", warps_per_sm, name, per * n / t / 1e6, per * n / t / 1e6 / sms / 1.9);
        };
        rep("IMAD.WIDE acc + LOP3", timeit([&] { k_madwide<<<blocks, threads>>>(o, 12345); }), 1);
        rep("mul.wide.s32", timeit([&] { k_mulwide<<<blocks, threads>>>(o, 12345); }), 1);
        rep("mad.lo.s32", timeit([&] { k_madlo<<<blocks, threads>>>(o, 12345); }), 1);
        rep("mul.hi.s32", timeit([&] { k_mulhi<<<blocks, threads>>>(o, 12345); }), 1);
        rep("add.s32", timeit([&] { k_iadd3<<<blocks, threads>>>(o, 12345); }), 1);
        rep("IMAD.WIDE acc + LOP3 + IADD", timeit([&] { k_mix<<<blocks, threads>>>(o, 12345); }), 1);
    }
    return 0;
}
