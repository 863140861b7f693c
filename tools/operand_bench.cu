// operand_bench.cu -- what a packed fp32x2 operation costs by operand form (sm_100a): issue cycles per warp instruction and
// scheduler for FFMA2 / FMUL2 / FADD2 with 64-bit register operands, with one 32-bit register broadcast to both lanes
// (.F32), with the same register twice, and against scalar FFMA. 8 independent chains per thread, 8 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o operand_bench operand_bench.cu && ./operand_bench
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

constexpr int ITER = 2048;
constexpr int ILP = 8;
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

#define KERNEL(NAME, BODY)                                                                          \
    __global__ void __launch_bounds__(1024, 1) NAME(float* out, long long* cycles, float seed) {    \
        float2 v[ILP], a[ILP];                                                                      \
        for (int k = 0; k < ILP; ++k) {                                                             \
            v[k] = make_float2(seed + k, seed + threadIdx.x);                                       \
            a[k] = make_float2(1.f + seed * 1e-7f * (k + threadIdx.x), 1.f - seed * 1e-7f * (k + 2 * threadIdx.x));                   \
        }                                                                                           \
        const float2 c = make_float2(seed * 1e-9f, seed * 2e-9f);                                   \
        const float s1 = 1.f + seed * 1e-8f, s2 = seed * 1e-9f;                                     \
        long long t0 = clock64();                                                                   \
        _Pragma("unroll 1") for (int it = 0; it < ITER; ++it) {                                     \
            _Pragma("unroll") for (int k = 0; k < ILP; ++k) { BODY; }                               \
        }                                                                                           \
        long long t1 = clock64();                                                                   \
        float r = 0;                                                                                \
        for (int k = 0; k < ILP; ++k) r += v[k].x + v[k].y + a[k].x;                                \
        out[blockIdx.x * blockDim.x + threadIdx.x] = r + c.x + s1 + s2;                             \
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;                                         \
    }

KERNEL(k_ffma2_rrr, v[k] = __ffma2_rn(v[k], a[k], c))                       // three 64-bit registers, all distinct
KERNEL(k_ffma2_rrr_acc, v[k] = __ffma2_rn(a[k], c, v[k]))                   // accumulate form
KERNEL(k_ffma2_rr_same, v[k] = __ffma2_rn(v[k], v[k], c))                   // a * a + c
KERNEL(k_ffma2_r_s_s, v[k] = __ffma2_rn(v[k], splat(s1), splat(s2)))        // 64-bit * .F32 + .F32
KERNEL(k_ffma2_r_r_s, v[k] = __ffma2_rn(v[k], a[k], splat(s2)))             // 64-bit * 64-bit + .F32
KERNEL(k_ffma2_r_s_r, v[k] = __ffma2_rn(v[k], splat(s1), a[k]))             // 64-bit * .F32 + 64-bit
KERNEL(k_ffma2_imm, v[k] = __ffma2_rn(v[k], a[k], splat(1.0f)))             // immediate addend
KERNEL(k_fmul2_rr, v[k] = __fmul2_rn(v[k], a[k]))
KERNEL(k_fmul2_same, v[k] = __fmul2_rn(v[k], v[k]))
KERNEL(k_fmul2_r_s, v[k] = __fmul2_rn(v[k], splat(s1)))
KERNEL(k_fadd2_rr, v[k] = __fadd2_rn(v[k], a[k]))
KERNEL(k_fadd2_r_s, v[k] = __fadd2_rn(v[k], splat(s2)))
KERNEL(k_ffma_scalar, v[k].x = fmaf(v[k].x, a[k].x, c.x))
KERNEL(k_ffma_scalar2, { v[k].x = fmaf(v[k].x, a[k].x, c.x); v[k].y = fmaf(v[k].y, a[k].y, c.y); })
KERNEL(k_mix_ffma2_ffma, { v[k] = __ffma2_rn(v[k], a[k], c); a[k].x = fmaf(a[k].x, s1, s2); })  // does a scalar FFMA hide behind a packed one?
KERNEL(k_mix_ffma2_iadd, { v[k] = __ffma2_rn(v[k], a[k], c); a[k].x = __int_as_float(__float_as_int(a[k].x) + 1); })

template <typename K>
void run(const char* name, K kernel, int sms, double inst_per_body) {
    const int threads = 1024;
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    kernel<<<sms, threads>>>(out, cyc, 1.5f);
    kernel<<<sms, threads>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto c : h) avg += (double)c;
    avg /= sms;
    // per scheduler: 8 warps, each ITER * ILP bodies
    const double bodies = (double)ITER * ILP * 8;
    printf("%-44s %6.2f cycles per body and scheduler (%g counted instructions per body)  %s\n", name, avg / bodies, inst_per_body,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs; loop overhead: 3 instructions per 8 bodies\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run("FFMA2 r64 * r64 + r64 (all distinct)", k_ffma2_rrr, sms, 1);
    run("FFMA2 acc += r64 * r64", k_ffma2_rrr_acc, sms, 1);
    run("FFMA2 a * a + r64", k_ffma2_rr_same, sms, 1);
    run("FFMA2 r64 * .F32 + .F32", k_ffma2_r_s_s, sms, 1);
    run("FFMA2 r64 * r64 + .F32", k_ffma2_r_r_s, sms, 1);
    run("FFMA2 r64 * .F32 + r64", k_ffma2_r_s_r, sms, 1);
    run("FFMA2 r64 * r64 + 1.0", k_ffma2_imm, sms, 1);
    run("FMUL2 r64 * r64", k_fmul2_rr, sms, 1);
    run("FMUL2 a * a", k_fmul2_same, sms, 1);
    run("FMUL2 r64 * .F32", k_fmul2_r_s, sms, 1);
    run("FADD2 r64 + r64", k_fadd2_rr, sms, 1);
    run("FADD2 r64 + .F32", k_fadd2_r_s, sms, 1);
    run("FFMA scalar", k_ffma_scalar, sms, 1);
    run("2 x FFMA scalar", k_ffma_scalar2, sms, 2);
    run("FFMA2 + FFMA scalar", k_mix_ffma2_ffma, sms, 2);
    run("FFMA2 + IADD", k_mix_ffma2_iadd, sms, 2);
    return 0;
}
