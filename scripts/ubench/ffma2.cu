// Microbenchmark: does FFMA2 (fma.rn.f32x2, sm_100) free issue slots?  Compares, at equal FLOPs,
//   A: scalar FFMA only            B: FFMA2 only
//   C: 2 FFMA + 2 FMNMX per step   D: 1 FFMA2 + 2 FMNMX per step
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r;}
__device__ __forceinline__ void unpack(u64 v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;}
__device__ __forceinline__ float fma1(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;}
__device__ __forceinline__ float mn(float a, float b){ float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s)
{
    float a[8], m[8];
    u64 p[4];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 0.001f + i; m[i] = a[i] + 1.f; }
#pragma unroll
    for (int i = 0; i < 4; i++) p[i] = pack(a[2 * i], a[2 * i + 1]);
    const u64 s2 = pack(s, s);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (MODE == 0 || MODE == 2) { a[2 * i] = fma1(a[2 * i], s, s); a[2 * i + 1] = fma1(a[2 * i + 1], s, s); }
            if (MODE == 1 || MODE == 3) p[i] = fma2(p[i], s2, s2);
            if (MODE >= 2) { m[2 * i] = mn(m[2 * i], s); m[2 * i + 1] = mn(m[2 * i + 1], s); s += 1.f; }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) r += a[i] + m[i];
#pragma unroll
    for (int i = 0; i < 4; i++) { float x, y; unpack(p[i], x, y); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> float run(float* out, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 16, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    const double fmas = 148.0 * 8 * 256 * iters * 8;   // scalar-FMA equivalents per kernel
    const char* names[4] = {"A scalar FFMA", "B FFMA2", "C 2 FFMA + 2 FMNMX (+FADD)", "D 1 FFMA2 + 2 FMNMX (+FADD)"};
    float ms[4] = {run<0>(out, iters), run<1>(out, iters), run<2>(out, iters), run<3>(out, iters)};
    for (int i = 0; i < 4; i++) printf("%-30s %8.3f ms  %7.2f TFMA/s\n", names[i], ms[i], fmas / (ms[i] * 1e-3) / 1e12);
    return 0;
}
