// halfstencil_bench.cu -- what would Newton's third law buy the pair loops of step_kernel_c?  
//
// A best-case model of the half-stencil variant, measured instead of estimated: the pair loop of step_float.cuh (same
// arithmetic: 13 packed operations, 2 MUFU.RCP, 2 MUFU.LG2, 1 LDS.64 per neighbour and couple) against the loop a half
// stencil would run -- half the neighbours, each with the reaction on top (2 FMUL + 2 FFMA for the sum over the couple's
// two particles, one STS.64 into a deterministic slot[k mod NS][j]), the slots cleared before and summed after the loops
// (NS = 12 sources per target). Everything the real thing would add on top is LEFT OUT in the half stencil's favour: the
// halo ring that still needs directed passes, multi-row tiles, boundary columns, divergent windows, the global loads and
// the epilogue. What is kept is the one structural cost that cannot be avoided: the slots live in shared memory (96 B per
// particle = 24 KB per 256-particle tile on top of the 22 KB of staged records), so fewer CTAs fit an SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o halfstencil_bench halfstencil_bench.cu && ./halfstencil_bench
//
// Output: time per tile-equivalent (128 couples x the stencil of a crystal: 39 neighbours) for the directed loop at 9
// CTAs per SM (the product's occupancy) and for the half-stencil loop at 9 (if the slots were free), 4 (what its
// shared-memory footprint allows) and 3 CTAs per SM.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int kCouples = 128;
constexpr int kRowCap = 448;
constexpr int kSlots = 12;  // sources per target: the couples of the 3 x 3 - 1 neighbouring cells that come later in the order

__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Poly {
    float d0, d1, d2, d3;
};

// the force law of step_float.cuh's pairc_xy<8, kFracPoly, false>: g = q^4 + q^8 * cubic(lg2 r^2)
__device__ __forceinline__ float2 force(float2 x, float2 y, const Poly& pc) {
    float2 r2 = __ffma2_rn(y, y, __fmul2_rn(x, x));
    float2 q = make_float2(fast_rcp(r2.x), fast_rcp(r2.y));
    float2 q2 = __fmul2_rn(q, q);
    float2 q4 = __fmul2_rn(q2, q2);
    float2 pn = __fmul2_rn(q4, q4);
    float2 l = make_float2(fast_lg2(r2.x), fast_lg2(r2.y));
    float2 e = __ffma2_rn(l, splat(pc.d3), splat(pc.d2));
    e = __ffma2_rn(l, e, splat(pc.d1));
    e = __ffma2_rn(l, e, splat(pc.d0));
    return __ffma2_rn(pn, e, q4);
}

// Shared memory of a CTA: float4 rec[3][kRowCap], then (HALF) `slot_rows` rows of float2 slot[kSlotRow] -- the model keeps
// as many slot rows as the CTA's share of the SM holds and lets the 12 logical rows wrap onto them (same instructions, same
// bank behaviour, any occupancy can be asked for) -- then padding that sets the number of CTAs per SM.
constexpr int kSlotRow = 2 * 303;  // own row + row above, padded so that consecutive lanes fall into different banks

template <bool HALF>
__global__ void __launch_bounds__(kCouples, 9) loop_kernel(float* __restrict__ out, int reps, float seed, Poly pf, int slot_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* rec = reinterpret_cast<float4*>(smem_raw);
    float2* slot = reinterpret_cast<float2*>(smem_raw + sizeof(float4) * 3 * kRowCap);

    // a lattice: 4.5 particles per cell, cells one unit wide, rows one unit apart
    for (int i = threadIdx.x; i < 3 * kRowCap; i += blockDim.x) {
        const int row = i / kRowCap, k = i % kRowCap;
        const float x = 0.2222f * k + 0.01f * ((k * 7 + row * 3) % 11), y = (float)row + 0.013f * ((k * 5) % 7);
        rec[i] = make_float4(x, y, x - 1.f, y);
    }
    __syncthreads();

    // the thread's couple: two neighbours of the own row, in cells of 4 and 5 particles in turn; its windows hold the 13
    // records from the start of the cell to the left
    const int i0 = 16 + 2 * threadIdx.x, i1 = i0 + 1;
    const int cell = (i0 * 2) / 9;
    const int ws = ((cell - 1) * 9 + 1) / 2;
    const float2 nx = make_float2(-rec[kRowCap + i0].x, -rec[kRowCap + i1].x);
    const float2 ny0 = make_float2(-rec[kRowCap + i0].y + 0.37f, -rec[kRowCap + i1].y + 0.41f);
    const float zero = seed * 0.f;
    Poly pc;
    pc.d0 = pf.d0 + zero, pc.d1 = pf.d1 + zero, pc.d2 = pf.d2 + zero, pc.d3 = pf.d3 + zero;

    float2 gx = splat(0.f), gy = splat(0.f);
#pragma unroll 1
    for (int rep = 0; rep < reps; ++rep) {
        if (HALF) {
            // clear the 12 slots of this tile's targets (2 x 128 own particles + as many of the row above)
            float4* s4 = reinterpret_cast<float4*>(slot);
            const int n4 = slot_rows * kSlotRow / 2;
            for (int i = threadIdx.x; i < kSlots * 2 * 2 * kCouples / 2; i += blockDim.x)
                s4[i < n4 ? i : i - n4 * (i / n4)] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncthreads();
        }
#pragma unroll
        for (int d = HALF ? 1 : 0; d < 3; ++d) {
            const float2 ny = __fadd2_rn(ny0, splat((float)(d - 1)));
            // HALF: the own row from the couple onwards (6 of 13 records), the row above in full; the row below is the
            // business of the tile below
            const int first = (HALF && d == 1) ? ws + 7 : ws;
            uint32_t pa = smem_u32(rec + d * kRowCap + first);
            const uint32_t pa_end = smem_u32(rec + d * kRowCap + ws + 13);
            // the reaction of record j goes to slot[k mod NS][j]: a second induction variable (8-byte entries against
            // 16-byte records)
            uint32_t ps = smem_u32(slot + ((int)threadIdx.x % slot_rows) * kSlotRow + (d - 1) * 303 + first);
#pragma unroll 1
            for (; pa < pa_end; pa += 16u) {
                float2 j;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(j.x), "=f"(j.y) : "r"(pa));
                const float2 x = __fadd2_rn(nx, splat(j.x)), y = __fadd2_rn(ny, splat(j.y));
                const float2 g = force(x, y, pc);
                gx = __ffma2_rn(g, x, gx);
                gy = __ffma2_rn(g, y, gy);
                if (HALF) {  // what the couple does to j (the sign is applied when the slots are summed)
                    const float rx = fmaf(g.y, x.y, g.x * x.x), ry = fmaf(g.y, y.y, g.x * y.x);
                    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ps), "f"(rx), "f"(ry) : "memory");
                    ps += 8u;
                }
            }
        }
        if (HALF) {
            // sum the slots of the thread's own two particles
            __syncthreads();
            for (int s = 0; s < kSlots; ++s) {
                const int r = s < slot_rows ? s : s % slot_rows;
                const float2 a = slot[r * kSlotRow + i0], b = slot[r * kSlotRow + i1];
                gx = __fadd2_rn(gx, make_float2(-a.x, -b.x));
                gy = __fadd2_rn(gy, make_float2(-a.y, -b.y));
            }
            __syncthreads();
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = gx.x + gx.y + gy.x + gy.y;
}

template <bool HALF>
double run(const char* name, int ctas_per_sm, int sms, int reps, float* out) {
    // shared memory per CTA that lets exactly ctas_per_sm CTAs fit (227 KB per SM, 1 KB reserved per CTA)
    size_t bytes = (size_t)(227 * 1024) / ctas_per_sm - 1024;
    bytes &= ~(size_t)127;
    const size_t recs = sizeof(float4) * 3 * kRowCap;
    int slot_rows = HALF ? (int)((bytes - recs) / (sizeof(float2) * kSlotRow)) : 0;
    if (slot_rows > kSlots) slot_rows = kSlots;
    if (bytes < recs || (HALF && slot_rows < 1)) {
        printf("%-28s %d CTAs/SM: the records alone do not fit\n", name, ctas_per_sm);
        return 0;
    }
    cudaFuncSetAttribute(loop_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, loop_kernel<HALF>, kCouples, bytes);
    const Poly pf{0.9f, -0.05f, 0.002f, -0.0001f};
    const int grid = sms * resident;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    loop_kernel<HALF><<<grid, kCouples, bytes>>>(out, reps, 1.5f, pf, slot_rows);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int k = 0; k < 5; ++k) {
        cudaEventRecord(e0);
        loop_kernel<HALF><<<grid, kCouples, bytes>>>(out, reps, 1.5f, pf, slot_rows);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double tiles = (double)grid * reps;
    const double ns_per_tile_sm = best * 1e6 / (tiles / sms);  // ns of one SM's time per tile-equivalent
    printf("%-28s %d CTAs/SM resident (asked %d, %2d slot rows held): %8.3f ms, %7.1f ns of SM time per tile  (%s)\n", name,
           resident, ctas_per_sm, slot_rows, best, ns_per_tile_sm, cudaGetErrorString(cudaGetLastError()));
    return ns_per_tile_sm;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 16 * kCouples);
    const int reps = 400;
    const double full = run<false>("directed (as built), 39 nb", 9, sms, reps, out);
    run<false>("directed, 39 nb", 6, sms, reps, out);
    run<false>("directed, 39 nb", 4, sms, reps, out);
    const double h9 = run<true>("half stencil, 19 nb + slots", 9, sms, reps, out);
    const double h6 = run<true>("half stencil, 19 nb + slots", 6, sms, reps, out);
    const double h4 = run<true>("half stencil, 19 nb + slots", 4, sms, reps, out);
    const double h3 = run<true>("half stencil, 19 nb + slots", 3, sms, reps, out);
    printf("half stencil / directed at 9 CTAs per SM -- loops, slots and their two passes only, everything else left out:\n");
    if (full > 0 && h9 > 0) printf("  9 CTAs/SM (slots for free):                        %.3f\n", h9 / full);
    if (full > 0 && h6 > 0) printf("  6 CTAs/SM (24 warps: multi-row tiles of 256):      %.3f\n", h6 / full);
    if (full > 0 && h4 > 0) printf("  4 CTAs/SM (records + 96 B of slots per particle):  %.3f\n", h4 / full);
    if (full > 0 && h3 > 0) printf("  3 CTAs/SM:                                         %.3f\n", h3 / full);
    cudaFree(out);
    return 0;
}
