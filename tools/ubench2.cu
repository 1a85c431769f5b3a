// ubench2.cu -- issue-rate of the half-band inner step  acc = acc + (wa + wb) * c  (no FMA):
// scalar FADD/FMUL/FADD vs packed add/mul/add.rn.f32x2, R independent accumulators per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench2 tools/ubench2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define R 8
#define W 16

__device__ __forceinline__ float fa(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fm(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d; }

__global__ void k_scalar(float *out, const float *in, float c0, float c1) {
    float w[W], acc[R];
    for (int i = 0; i < W; i++) w[i] = in[threadIdx.x + 32 * i];
    for (int i = 0; i < R; i++) acc[i] = 0.f;
    for (int it0 = 0; it0 < ITERS; it0 += W) {
#pragma unroll
      for (int it = 0; it < W; it++) {
#pragma unroll
        for (int i = 0; i < R; i++) {
            acc[i] = fa(acc[i], fm(fa(w[(i + it) % W], w[(W - 1 - i + it) % W]), c0));
            acc[i] = fa(acc[i], fm(fa(w[(i + 1 + it) % W], w[(W - 2 - i + it) % W]), c1));
        }
      }
      w[0] = fa(w[0], acc[0]);
    }
    float s = 0;
    for (int i = 0; i < R; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed(float *out, const float *in, float c0, float c1) {
    u64 w[W], acc[R];
    for (int i = 0; i < W; i++) w[i] = pk(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 7]);
    for (int i = 0; i < R; i++) acc[i] = pk(0.f, 0.f);
    const u64 cc0 = pk(c0, c0), cc1 = pk(c1, c1), one = pk(1.f, 1.f);
    for (int it0 = 0; it0 < ITERS; it0 += W) {
#pragma unroll
      for (int it = 0; it < W; it++) {
#pragma unroll
        for (int i = 0; i < R; i++) {
            acc[i] = add2(acc[i], mul2(add2(w[(i + it) % W], w[(W - 1 - i + it) % W]), cc0));
            acc[i] = add2(acc[i], mul2(add2(w[(i + 1 + it) % W], w[(W - 2 - i + it) % W]), cc1));
        }
      }
      w[0] = add2(w[0], acc[0]);
    }
    float s = 0;
    for (int i = 0; i < R; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// packed adds, scalar muls (FADD2, 2x FMUL, FADD2)
__global__ void k_mixed(float *out, const float *in, float c0, float c1) {
    u64 w[W], acc[R];
    for (int i = 0; i < W; i++) w[i] = pk(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 7]);
    for (int i = 0; i < R; i++) acc[i] = pk(0.f, 0.f);
    const u64 one = pk(1.f, 1.f);
    for (int it0 = 0; it0 < ITERS; it0 += W) {
#pragma unroll
      for (int it = 0; it < W; it++) {
#pragma unroll
        for (int i = 0; i < R; i++) {
            float lo, hi;
            u64 s = add2(w[(i + it) % W], w[(W - 1 - i + it) % W]);
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s));
            acc[i] = add2(acc[i], pk(fm(lo, c0), fm(hi, c0)));
            s = add2(w[(i + 1 + it) % W], w[(W - 2 - i + it) % W]);
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s));
            acc[i] = add2(acc[i], pk(fm(lo, c1), fm(hi, c1)));
        }
      }
      w[0] = add2(w[0], acc[0]);
    }
    float s = 0;
    for (int i = 0; i < R; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
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
    float *o, *in; cudaMalloc(&o, sizeof(float) * sms * 8 * 1024); cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20);
    for (int warps_per_sm : {8, 16, 32}) {
        int threads = 128, blocks = sms * warps_per_sm * 32 / threads;
        double taps = (double)blocks * threads * ITERS * R * 2;  // (add, mul, add) triples per thread-lane
        float t;
        t = timeit([&] { k_scalar<<<blocks, threads>>>(o, in, 1.0001f, 0.5f); });
        printf("warps/SM %2d  scalar  %8.1f G tap-steps/s (x3 = %.1f Gflop/s)\n", warps_per_sm, taps / t / 1e6, 3 * taps / t / 1e6);
        t = timeit([&] { k_packed<<<blocks, threads>>>(o, in, 1.0001f, 0.5f); });
        printf("             packed  %8.1f G tap-steps/s (2 per packed triple)\n", 2 * taps / t / 1e6);
        t = timeit([&] { k_mixed<<<blocks, threads>>>(o, in, 1.0001f, 0.5f); });
        printf("             mixed   %8.1f G tap-steps/s (FADD2 + 2 FMUL + FADD2)\n", 2 * taps / t / 1e6);
    }
    return 0;
}
