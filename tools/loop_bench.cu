// loop_bench.cu -- what bounds the pair loop of step_kernel_c?  Variants of the loop body on the same synthetic tile
// (128 couples, three windows of 13 staged records, 9 CTAs per SM like the product), timed alone: SM cycles per trip and
// warp (one trip = one neighbour against a couple = two pair forces).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o loop_bench loop_bench.cu && ./loop_bench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

constexpr int kCouples = 128;
constexpr int kRowCap = 300;  // records per staged row (the windows reach index 285)

__device__ __forceinline__ float mufu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Poly { float d0, d1, d2, d3; };

enum Variant {
    kAsBuilt = 0,    // 13 packed + 2 RCP + 2 LG2
    kNoLg2,          // integer exponents only: 10 packed + 2 RCP
    kNoMufu,         // every MUFU replaced by one packed multiply: the FMA-pipe / issue bound
    kMufuOnly,       // geometry + 4 MUFU + accumulate: 6 packed + 4 MUFU
    kSharedRcp,      // one RCP for the two lanes: 13 packed + 3 scalar + 1 RCP + 2 LG2
    kEx2ForRcp,      // q = 2^(-lg2 r2): 2 LG2 + 2 EX2
    kLg2First,       // as built, the source asks for LG2 after RCP (the compiler schedules them anyway)
    kTwoLg2Only,     // 2 LG2, no RCP (q from a multiply)
    kTwoRcpOnly,     // 2 RCP, LG2 replaced by a multiply
    kRsqForRcp,      // 2 RSQ instead of 2 RCP (same count, other function)
    kScalar,         // the same arithmetic in 26 scalar operations (no packed fp32x2 at all)
    kHybrid,         // three-operand operations scalar (14 FFMA), two-operand ones packed (4 FMUL2, 2 FADD2)
    kConstUniform,   // as built, the cubic's constants straight from the kernel parameters (uniform registers / constant bank)
    kVariants
};
const char* kNames[kVariants] = {"as built (13 FP2, 2 RCP, 2 LG2)", "no LG2 (10 FP2, 2 RCP)", "no MUFU (15 FP2)",
                                 "MUFU only (6 FP2, 2 RCP, 2 LG2)", "shared RCP (13 FP2 + 3 FP, 1 RCP, 2 LG2)",
                                 "EX2 for RCP (13 FP2, 2 LG2, 2 EX2)", "as built, other source order",
                                 "2 LG2, no RCP (14 FP2)", "2 RCP, no LG2 (14 FP2)", "RSQ for RCP (13 FP2, 2 RSQ, 2 LG2)",
                                 "scalar (26 FP, 2 RCP, 2 LG2)", "hybrid (14 FFMA + 6 FP2, 2 RCP, 2 LG2)",
                                 "as built, constants uniform"};

__device__ __forceinline__ float sfma(float a, float b, float c) { float r; asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float smul(float a, float b) { float r; asm("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float sadd(float a, float b) { float r; asm("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

// one lane of the couple, scalar
__device__ __forceinline__ void lane_scalar(float x, float y, const Poly& pc, float& gx, float& gy) {
    const float r2 = sfma(y, y, smul(x, x));
    const float q = mufu_rcp(r2), l = mufu_lg2(r2);
    const float q2 = smul(q, q), q4 = smul(q2, q2), pn = smul(q4, q4);
    float e = sfma(l, pc.d3, pc.d2);
    e = sfma(l, e, pc.d1);
    e = sfma(l, e, pc.d0);
    const float g = sfma(pn, e, q4);
    gx = sfma(g, x, gx);
    gy = sfma(g, y, gy);
}

template <int V>
__device__ __forceinline__ void body(float2 x, float2 y, const Poly& pc, float2& gx, float2& gy) {
    if (V == kScalar) {
        lane_scalar(x.x, y.x, pc, gx.x, gy.x);
        lane_scalar(x.y, y.y, pc, gx.y, gy.y);
        return;
    }
    if (V == kHybrid) {
        const float2 xx = __fmul2_rn(x, x);
        const float2 r2 = make_float2(sfma(y.x, y.x, xx.x), sfma(y.y, y.y, xx.y));
        const float2 q = make_float2(mufu_rcp(r2.x), mufu_rcp(r2.y)), l = make_float2(mufu_lg2(r2.x), mufu_lg2(r2.y));
        const float2 q2 = __fmul2_rn(q, q), q4 = __fmul2_rn(q2, q2), pn = __fmul2_rn(q4, q4);
        float2 e = make_float2(sfma(l.x, pc.d3, pc.d2), sfma(l.y, pc.d3, pc.d2));
        e = make_float2(sfma(l.x, e.x, pc.d1), sfma(l.y, e.y, pc.d1));
        e = make_float2(sfma(l.x, e.x, pc.d0), sfma(l.y, e.y, pc.d0));
        const float2 g = make_float2(sfma(pn.x, e.x, q4.x), sfma(pn.y, e.y, q4.y));
        gx = make_float2(sfma(g.x, x.x, gx.x), sfma(g.y, x.y, gx.y));
        gy = make_float2(sfma(g.x, y.x, gy.x), sfma(g.y, y.y, gy.y));
        return;
    }
    float2 r2 = __ffma2_rn(y, y, __fmul2_rn(x, x));
    float2 q, l;
    if (V == kSharedRcp) {
        const float rr = mufu_rcp(r2.x * r2.y);
        q = make_float2(r2.y * rr, r2.x * rr);
    } else if (V == kNoMufu || V == kTwoLg2Only) {
        q = __fmul2_rn(r2, splat(pc.d1));
    } else if (V == kEx2ForRcp) {
        l = make_float2(mufu_lg2(r2.x), mufu_lg2(r2.y));
        q = make_float2(mufu_ex2(-l.x), mufu_ex2(-l.y));
    } else if (V == kRsqForRcp) {
        q = make_float2(mufu_rsq(r2.x), mufu_rsq(r2.y));
    } else {
        q = make_float2(mufu_rcp(r2.x), mufu_rcp(r2.y));
    }
    if (V == kMufuOnly) {
        l = make_float2(mufu_lg2(r2.x), mufu_lg2(r2.y));
        const float2 g = __fmul2_rn(q, l);
        gx = __ffma2_rn(g, x, gx);
        gy = __ffma2_rn(g, y, gy);
        return;
    }
    float2 q2 = __fmul2_rn(q, q);
    float2 q4 = __fmul2_rn(q2, q2);
    float2 pn = __fmul2_rn(q4, q4);
    float2 e;
    if (V == kNoLg2) {
        e = splat(pc.d0);
    } else {
        if (V == kNoMufu || V == kTwoRcpOnly) l = __fmul2_rn(r2, splat(pc.d2));
        else if (V != kEx2ForRcp) l = make_float2(mufu_lg2(r2.x), mufu_lg2(r2.y));
        e = __ffma2_rn(l, splat(pc.d3), splat(pc.d2));
        e = __ffma2_rn(l, e, splat(pc.d1));
        e = __ffma2_rn(l, e, splat(pc.d0));
    }
    float2 g = __ffma2_rn(pn, e, q4);
    gx = __ffma2_rn(g, x, gx);
    gy = __ffma2_rn(g, y, gy);
}

template <int V, int UNROLL>
__global__ void __launch_bounds__(kCouples, 9) loop_kernel(float* __restrict__ out, int reps, float seed, Poly pf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* rec = reinterpret_cast<float4*>(smem_raw);
    for (int i = threadIdx.x; i < 3 * kRowCap; i += blockDim.x) {
        const int row = i / kRowCap, k = i % kRowCap;
        const float x = 0.2222f * k + 0.01f * ((k * 7 + row * 3) % 11), y = (float)row + 0.013f * ((k * 5) % 7);
        rec[i] = make_float4(x, y, x - 1.f, y);
    }
    __syncthreads();
    const int i0 = 16 + 2 * threadIdx.x, i1 = i0 + 1;
    const int cell = (i0 * 2) / 9;
    const int ws = ((cell - 1) * 9 + 1) / 2;
    const float2 nx = make_float2(-rec[kRowCap + i0].x, -rec[kRowCap + i1].x);
    const float2 ny0 = make_float2(-rec[kRowCap + i0].y + 0.37f, -rec[kRowCap + i1].y + 0.41f);
    const float zero = seed * 0.f;
    Poly pc;
    if (V == kConstUniform) pc = pf;
    else pc.d0 = pf.d0 + zero, pc.d1 = pf.d1 + zero, pc.d2 = pf.d2 + zero, pc.d3 = pf.d3 + zero;
    float2 gx = splat(0.f), gy = splat(0.f);
#pragma unroll 1
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float2 ny = __fadd2_rn(ny0, splat((float)(d - 1)));
            uint32_t pa = smem_u32(rec + d * kRowCap + ws);
            const uint32_t pa_end = smem_u32(rec + d * kRowCap + ws + (UNROLL == 2 ? 14 : 13));
#pragma unroll 1
            for (; pa < pa_end; pa += 16u * UNROLL) {
                float2 j, k;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(j.x), "=f"(j.y) : "r"(pa));
                if (UNROLL == 2) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+16];" : "=f"(k.x), "=f"(k.y) : "r"(pa));
                body<V>(__fadd2_rn(nx, splat(j.x)), __fadd2_rn(ny, splat(j.y)), pc, gx, gy);
                if (UNROLL == 2) body<V>(__fadd2_rn(nx, splat(k.x)), __fadd2_rn(ny, splat(k.y)), pc, gx, gy);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = gx.x + gx.y + gy.x + gy.y;
}

template <int V, int UNROLL>
void run(int ctas_per_sm, int sms, float clock_ghz, float* out) {
    const int reps = 300;
    size_t bytes = (size_t)(227 * 1024) / ctas_per_sm - 1024;
    bytes &= ~(size_t)127;
    cudaFuncSetAttribute(loop_kernel<V, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, loop_kernel<V, UNROLL>, kCouples, bytes);
    const Poly pf{0.9f, -0.05f, 0.002f, -0.0001f};
    const int grid = sms * resident;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    loop_kernel<V, UNROLL><<<grid, kCouples, bytes>>>(out, reps, 1.5f, pf);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int k = 0; k < 5; ++k) {
        cudaEventRecord(e0);
        loop_kernel<V, UNROLL><<<grid, kCouples, bytes>>>(out, reps, 1.5f, pf);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // trips per scheduler: every CTA has one warp on each of the 4 schedulers of its SM
    const double trips = (double)resident * reps * 3 * (UNROLL == 2 ? 14 : 13);
    const double cycles = best * 1e-3 * clock_ghz * 1e9;
    printf("%-44s unroll %d, %d CTAs/SM: %7.3f ms  %6.2f cycles per trip and scheduler  (%s)\n", kNames[V], UNROLL, resident, best,
           cycles / trips, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float ghz = khz * 1e-6f;
    printf("%s, %d SMs, %.3f GHz (nominal: cycles below assume the SMs run at it)\n", p.name, sms, ghz);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 16 * kCouples);
    run<kAsBuilt, 1>(9, sms, ghz, out);
    run<kAsBuilt, 2>(9, sms, ghz, out);
    run<kAsBuilt, 1>(12, sms, ghz, out);
    run<kAsBuilt, 1>(6, sms, ghz, out);
    run<kNoLg2, 1>(9, sms, ghz, out);
    run<kNoMufu, 1>(9, sms, ghz, out);
    run<kMufuOnly, 1>(9, sms, ghz, out);
    run<kSharedRcp, 1>(9, sms, ghz, out);
    run<kEx2ForRcp, 1>(9, sms, ghz, out);
    run<kTwoLg2Only, 1>(9, sms, ghz, out);
    run<kTwoRcpOnly, 1>(9, sms, ghz, out);
    run<kRsqForRcp, 1>(9, sms, ghz, out);
    run<kNoMufu, 2>(9, sms, ghz, out);
    run<kSharedRcp, 2>(9, sms, ghz, out);
    run<kScalar, 1>(9, sms, ghz, out);
    run<kHybrid, 1>(9, sms, ghz, out);
    run<kConstUniform, 1>(9, sms, ghz, out);
    run<kScalar, 2>(9, sms, ghz, out);
    run<kHybrid, 2>(9, sms, ghz, out);
    cudaFree(out);
    return 0;
}
