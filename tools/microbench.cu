// microbench.cu -- per-SM issue throughput of the instructions the step kernel is made of, measured
// on the box (B200): FFMA, packed FFMA2 / FMUL2, MUFU.{RCP,LG2,EX2}, I2FP, IADD, LDS.64 / LDS.128.
// Every kernel runs ILP independent chains per thread, 1024 threads per SM (8 warps per scheduler),
// and reports warp-instructions per cycle per SM (4.0 = every scheduler issues every cycle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

constexpr int ITER = 4096;
constexpr int ILP = 8;

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo(u64 v) {
    float a, b;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a + b;
}

#define KERNEL(NAME, DECL, BODY, FIN)                                                  \
    __global__ void NAME(float* out, long long* cycles, float seed) {                  \
        DECL;                                                                          \
        long long t0 = clock64();                                                      \
        _Pragma("unroll 1") for (int it = 0; it < ITER; ++it) {                        \
            _Pragma("unroll") for (int k = 0; k < ILP; ++k) { BODY; }                  \
        }                                                                              \
        long long t1 = clock64();                                                      \
        float r = 0;                                                                   \
        _Pragma("unroll") for (int k = 0; k < ILP; ++k) { FIN; }                       \
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;                                \
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;                            \
    }

KERNEL(k_ffma, float v[ILP]; for (int k = 0; k < ILP; ++k) v[k] = seed + k + threadIdx.x,
       v[k] = fmaf(v[k], 1.0000001f, 0.5f), r += v[k])
KERNEL(k_ffma2, u64 v[ILP]; u64 c1 = pack(1.0000001f, 0.9999999f); u64 c2 = pack(0.5f, 0.25f);
       for (int k = 0; k < ILP; ++k) v[k] = pack(seed + k, seed + threadIdx.x),
       asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[k]) : "l"(c1), "l"(c2)), r += lo(v[k]))
KERNEL(k_fmul2, u64 v[ILP]; u64 c1 = pack(1.0000001f, 0.9999999f);
       for (int k = 0; k < ILP; ++k) v[k] = pack(seed + k, seed + threadIdx.x),
       asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[k]) : "l"(c1)), r += lo(v[k]))
KERNEL(k_rcp, float v[ILP]; for (int k = 0; k < ILP; ++k) v[k] = seed + k + threadIdx.x,
       asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[k])), r += v[k])
KERNEL(k_lg2, float v[ILP]; for (int k = 0; k < ILP; ++k) v[k] = seed + k + threadIdx.x,
       asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[k])), r += v[k])
KERNEL(k_ex2, float v[ILP]; for (int k = 0; k < ILP; ++k) v[k] = (seed + k + threadIdx.x) * 1e-3f,
       asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[k])), r += v[k])
KERNEL(k_rsqrt, float v[ILP]; for (int k = 0; k < ILP; ++k) v[k] = seed + k + threadIdx.x,
       asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[k])), r += v[k])
KERNEL(k_i2f, int v[ILP]; for (int k = 0; k < ILP; ++k) v[k] = (int)seed + k + threadIdx.x,
       v[k] = __float_as_int(__int2float_rn(v[k])) + 3, r += v[k])
KERNEL(k_iadd, int v[ILP]; int s = (int)seed; for (int k = 0; k < ILP; ++k) v[k] = (int)seed + k + threadIdx.x,
       asm volatile("sub.s32 %0, %0, %1;" : "+r"(v[k]) : "r"(s)), r += v[k])
// mixed: the instruction mix of one packed pair-of-pairs body (2 pairs): see DESIGN.md
KERNEL(k_mix, u64 v[ILP]; float w[ILP]; u64 c1 = pack(1.0000001f, 0.9999999f); u64 c2 = pack(0.5f, 0.25f);
       for (int k = 0; k < ILP; ++k) { v[k] = pack(seed + k, seed + threadIdx.x); w[k] = seed + k; },
       { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[k]) : "l"(c1), "l"(c2));
         asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[k]) : "l"(c1), "l"(c2));
         asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[k]) : "l"(c1), "l"(c2));
         asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(w[k])); },
       r += lo(v[k]) + w[k])

__global__ void k_lds64(float* out, long long* cycles, float seed) {
    __shared__ uint2 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_uint2(i, i + 1);
    __syncthreads();
    unsigned acc = 0;
    int base = (threadIdx.x >> 2) & 1023;  // 8 distinct addresses per warp, like neighbouring cells
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            uint2 v = buf[(base + k + it) & 2047];
            acc += v.x ^ v.y;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
__global__ void k_lds128(float* out, long long* cycles, float seed) {
    __shared__ uint4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    unsigned acc = 0;
    int base = (threadIdx.x >> 2) & 511;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            uint4 v = buf[(base + k + it) & 1023];
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename K>
void run(const char* name, K kernel, int sms, double inst_per_body) {
    const int threads = 1024;
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    kernel<<<sms, threads>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kernel<<<sms, threads>>>(out, cyc, 1.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto c : h) avg += (double)c;
    avg /= sms;
    double warp_inst = (double)ITER * ILP * inst_per_body * (threads / 32);
    printf("%-8s %8.3f warp-inst/clk/SM   (%.0f cycles, %.3f ms, err=%s)\n", name, warp_inst / avg, avg, ms,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %.0f MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1e3);
    int sms = p.multiProcessorCount;
    run("FFMA", k_ffma, sms, 1);
    run("FFMA2", k_ffma2, sms, 1);
    run("FMUL2", k_fmul2, sms, 1);
    run("RCP", k_rcp, sms, 1);
    run("LG2", k_lg2, sms, 1);
    run("EX2", k_ex2, sms, 1);
    run("RSQRT", k_rsqrt, sms, 1);
    run("I2FP+IADD", k_i2f, sms, 2);
    run("IADD", k_iadd, sms, 1);
    run("3FFMA2+RCP", k_mix, sms, 4);
    run("LDS.64", k_lds64, sms, 1);
    run("LDS.128", k_lds128, sms, 1);
    return 0;
}
