// stepper.cu -- the B200 (sm_100a) particle stepper behind include/psim_b200.h.
//
// What it replaces in the reference (paths under /root/reference/cuda_simulator/src):
//   kernel_prepare_frame  kernel.cuh:200-250      -> psim_upload_frame  (GPU stable counting sort)
//   bucket_move_kernel    kernel_bucket.cuh:5-39  -> rebin()            (same sort, on the live state)
//   bucket_step_kernel    kernel_bucket.cuh:40-94 -> step_kernel        (fused force + kick + drift)
//   bucket_kernel_run_async kernel_bucket.cuh:181-206 -> run_frame()    (same step / re-bin schedule)
//   Kernel::{write,read,sync,write_metadata} kernel.cuh:88-129 -> upload / download / sync / set_metadata
//
// Data layout in HBM (all cell-sorted, structure of arrays, no per-cell capacity, no null slots):
//   pos[2][n]  uint2  fixed-point (x, y), ping-pong: a step reads pos[cur] and writes pos[cur^1]
//   vel[n]     float2 half-step velocities, updated in place by a step
//   ty[n]      int32  species label, untouched by a step
//   cell_start[cells+1] uint32 exclusive prefix sum of per-cell counts (CSR); cell = cx + cy*BX
//   cell_id[n] uint32 cell of every particle as of the last binning (membership, not position)
//   tiles[ceil(n/128)] TileDesc: what the CTA of each tile of 128 consecutive particles stages
// Membership is by the LAST binning, exactly like the reference's slot array: between re-bins a
// particle keeps its index and its cell even if it has drifted out of it (kernel_bucket.cuh:71-91).
//
// There is no CPU path in this file and nothing here includes or links oracle/.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "psim_b200.h"

namespace {

// ------------------------------------------------------------------------------------------------
// Kernel parameter blocks
// ------------------------------------------------------------------------------------------------

constexpr int kTile = 128;        // particles per CTA of the step kernel (= threads per CTA)
constexpr int kCsCap = 288;       // cell_start entries staged per stencil row (multiple of 4)
constexpr int kPosCap = 640;      // neighbour positions staged per stencil row (multiple of 2)
constexpr int kScanItems = 8;     // cells per thread in the scan kernels
constexpr int kScanThreads = 256;
constexpr int kScanBlock = kScanItems * kScanThreads;
constexpr uint32_t kNoKey = 0xFFFFFFFFu;
constexpr int kPadCells = 8;      // readable slack after cell_start[cells] for 16-byte bulk copies
constexpr int kPadParticles = 4;  // readable slack after pos[n] for 16-byte bulk copies

struct Grid {
    uint32_t lx, ly;    // log2 cells in x / y
    uint32_t bx, by;    // cells in x / y
    uint32_t cells;     // bx * by
    uint32_t sx, sy;    // 32 - lx, 32 - ly (shift that maps a fixed-point coordinate to its cell)
};

// How the non-integer part of the repulsive exponent is evaluated (see pair2 below).
enum FracMode { kFracNone = 0, kFracPoly = 1, kFracEx2 = 2 };

// Everything a step needs from FrameMetadata, pre-digested on the host once per metadata change
// (the reference rebuilds ParticleParams, including a powf, in every thread of every step:
// kernel_bucket.cuh:52, particle.cuh:53-55).
//
// Pair force in the units the kernel works in.  With q = sigma^2 / r^2:
//   F_vec = C eps (m (s/r)^m - n (s/r)^n) / r^2 * r_vec                      (particle.cuh:63-66,97-103)
//         = (C eps m / sigma^2) * (q^(m/2+1) - (n/m) q^(n/2+1)) * r_vec
// r_vec is kept in raw fixed-point x units (dx, dy * yscale), so
//   F_vec = pair_scale * sum_j g_j * (dx, dy')      with pair_scale = C eps m kx / sigma^2.
struct Phys {
    float inv_c2;        // kx^2 / sigma^2: (raw x units)^2 -> r^2 / sigma^2
    float yscale;        // ky / kx (1 for square cells): raw y units -> raw x units
    float nm;            // n / m
    float fn, fm;        // exponents n/2+1 = kn + fn, m/2+1 = km + fm  (|fn|, |fm| <= 0.5)
    int kn, km;
    float c1, c2, c3;    // 2^z ~ 1 + z (c1 + z (c2 + z c3)) on the z range of kFracPoly
    float pair_scale;    // scaled pair sum -> newtons (x), see above
    float pair_scale_y;  // same for y: pair_scale (dy' is already in x units)
    float wall_scale;    // C * eps * m
    float sigma;
    float inv_mass;
    float dt;
    float kx, ky;        // box / 2^32: fixed-point units -> metres
    float ux, uy;        // dt * 2^32 / box: velocity -> fixed-point displacement per step
    float cursor_x, cursor_y, cursor_r2;  // cursor_r2 = cursor_size^2 / 4
    int wall_m6;         // m == 6: wall term by multiplication
    float m;
};

// One tile of kTile consecutive particles: what its CTA stages in shared memory. Written at re-bin
// time (tile_desc_kernel), read by every step until the next re-bin.  64 bytes.
struct __align__(16) TileDesc {
    uint32_t fits;       // 1: the three stencil rows fit the staging buffers
    uint32_t first;      // first / last cell touched by the tile's own particles
    uint32_t last;
    uint32_t _pad;
    uint32_t cs_lo[3];   // first cell_start entry staged per row (multiple of 4)
    uint32_t cs_cnt[3];  // entries staged per row (multiple of 4; 0: row outside the grid)
    uint32_t p_lo[3];    // first particle staged per row (even)
    uint32_t p_cnt[3];   // particles staged per row (even)
};
static_assert(sizeof(TileDesc) == 64, "TileDesc is read as four 16-byte words");

// ------------------------------------------------------------------------------------------------
// Device helpers
// ------------------------------------------------------------------------------------------------

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
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

__device__ __forceinline__ float2 powi2(float2 b, int k) {  // k is uniform across the grid
    float2 r = splat(1.f);
    while (k) {
        if (k & 1) r = __fmul2_rn(r, b);
        b = __fmul2_rn(b, b);
        k >>= 1;
    }
    return r;
}

__device__ __forceinline__ uint32_t cell_of(uint2 p, const Grid& g) {
    // kernel.cuh:224-226
    return (p.x >> g.sx) + ((p.y >> g.sy) << g.lx);
}

// mbarrier + 1-D bulk copy (TMA) wrappers: global -> shared::cta, completion counted in bytes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// Two pairs at a time, on the packed fp32x2 pipe (FMUL2 / FFMA2, new on sm_100): the separation
// i -> j (f_dist, particle.cuh:41-47: exact u32 difference, one int -> float conversion) and the Mie
// force written on r^2 so that no square root is needed (see Phys).  With r2 = r^2 / sigma^2 and
// q = 1 / r2, per pair   g = q^km - (n/m) q^kn * q^fn   (km = 4 for m = 6).
// q^fn, the non-integer sliver of the repulsive exponent (n = 14.08 -> kn = 8, fn = 0.04), is
// 2^(-fn log2 r2): MUFU.LG2, then either a host-fitted cubic in z = -fn log2 r2 (kFracPoly: |z| is
// small, the cubic is exact to ~1e-8 where the term matters) or MUFU.EX2 (kFracEx2).  Integer powers
// are products; each term keeps a relative error of a few 1e-7, no worse than the reference's own
// fp32 `powf(sigma / len, n)`.  MUFU runs at 16 lanes/clk/SM (measured, tools/microbench.cu), so the
// two MUFU ops per pair (RCP, LG2) are what bounds this loop, with issue slots a close second.
// MASK1: the second pair of the packed couple does not exist (odd tail) and contributes exactly 0.
// CLAMP: the window may contain i itself (own row): r2 is clamped away from 0 so that g stays
// finite and the zero separation gives an exact 0 (kernel_bucket.cuh:85 skips j == i).
// ------------------------------------------------------------------------------------------------
template <int KN, int FRAC, bool ANISO, bool MASK1, bool CLAMP>
__device__ __forceinline__ void pair2(uint2 pi, uint2 pj0, uint2 pj1, const Phys& ph, float2& gx, float2& gy) {
    float2 x = make_float2(__int2float_rn((int)(pj0.x - pi.x)), __int2float_rn((int)(pj1.x - pi.x)));
    float2 y = make_float2(__int2float_rn((int)(pj0.y - pi.y)), __int2float_rn((int)(pj1.y - pi.y)));
    if (ANISO) y = __fmul2_rn(y, splat(ph.yscale));
    float2 r2 = __ffma2_rn(y, y, __fmul2_rn(x, x));
    r2 = __fmul2_rn(r2, splat(ph.inv_c2));
    if (MASK1) r2.y = 1e12f;
    if (CLAMP) {
        r2.x = fmaxf(r2.x, 1e-2f);
        r2.y = fmaxf(r2.y, 1e-2f);
    }
    float2 q = make_float2(fast_rcp(r2.x), fast_rcp(r2.y));
    float2 q2 = __fmul2_rn(q, q);
    float2 q4 = __fmul2_rn(q2, q2);
    float2 pm, pn;
    if (KN > 0) {  // m = 6 (q^4) and a compile-time integer part of n/2 + 1
        pm = q4;
        if (KN == 5) pn = __fmul2_rn(q4, q);
        else if (KN == 6) pn = __fmul2_rn(q4, q2);
        else if (KN == 7) pn = __fmul2_rn(__fmul2_rn(q4, q2), q);
        else if (KN == 8) pn = __fmul2_rn(q4, q4);
        else if (KN == 9) pn = __fmul2_rn(__fmul2_rn(q4, q4), q);
        else pn = __fmul2_rn(__fmul2_rn(q4, q4), q2);
    } else {  // any exponents: run-time integer parts
        pm = powi2(q, ph.km);
        pn = powi2(q, ph.kn);
    }
    if (FRAC != kFracNone) {
        float2 l = make_float2(fast_lg2(r2.x), fast_lg2(r2.y));
        float2 z = __fmul2_rn(l, splat(-ph.fn));
        float2 e;
        if (FRAC == kFracPoly) {
            e = __ffma2_rn(z, splat(ph.c3), splat(ph.c2));
            e = __ffma2_rn(z, e, splat(ph.c1));
            e = __ffma2_rn(z, e, splat(1.f));
        } else {
            e = make_float2(fast_ex2(z.x), fast_ex2(z.y));
        }
        pn = __fmul2_rn(pn, e);
        if (KN == 0 && ph.fm != 0.f) {
            float2 zm = __fmul2_rn(l, splat(-ph.fm));
            pm = __fmul2_rn(pm, make_float2(fast_ex2(zm.x), fast_ex2(zm.y)));
        }
    }
    float2 g = __ffma2_rn(pn, splat(-ph.nm), pm);
    gx = __ffma2_rn(g, x, gx);
    gy = __ffma2_rn(g, y, gy);
}

// All of one window [0, count) of staged neighbours, two at a time, in ascending index order.
template <int KN, int FRAC, bool ANISO, bool CLAMP>
__device__ __forceinline__ void window_accumulate(const uint2* __restrict__ pj, int count, uint2 pi, const Phys& ph,
                                                  float2& gx, float2& gy) {
    int k = 0;
#pragma unroll 2
    for (; k + 1 < count; k += 2) pair2<KN, FRAC, ANISO, false, CLAMP>(pi, pj[k], pj[k + 1], ph, gx, gy);
    if (k < count) pair2<KN, FRAC, ANISO, true, CLAMP>(pi, pj[k], pi, ph, gx, gy);
}

// Repulsive wall term C eps m (sigma/d)^m / d (particle.cuh:68-71).
__device__ __forceinline__ float wall_term(float d, const Phys& ph) {
    float inv_d = fast_rcp(d);
    float q = ph.sigma * inv_d;
    float pw;
    if (ph.wall_m6) {
        float q2 = q * q;
        pw = q2 * q2 * q2;
    } else {
        pw = fast_ex2(ph.m * fast_lg2(q));
    }
    return ph.wall_scale * pw * inv_d;
}

// Cursor + wall forces on one particle (kernel_bucket.cuh:54-69, particle.cuh:125-144).
__device__ __forceinline__ float2 field_force(uint2 p, const Phys& ph) {
    const float inv32 = 1.f / 4294967296.f;
    float2 f = make_float2(0.f, 0.f);
    float dx = ph.cursor_x - __uint2float_rn(p.x) * inv32;
    float dy = ph.cursor_y - __uint2float_rn(p.y) * inv32;
    float sq = dx * dx + dy * dy;
    if (sq < ph.cursor_r2) {
        float c = 8e-12f * fast_rcp(sq + 1.f);
        f.x = dx > 0 ? -c : c;
        f.y = dy > 0 ? -c : c;
    }
    if (p.x < 0xFFFFFFFFu / 2) f.x += wall_term(__uint2float_rn(p.x) * ph.kx, ph);
    else f.x -= wall_term(__uint2float_rn(0xFFFFFFFFu - p.x) * ph.kx, ph);
    if (p.y < 0xFFFFFFFFu / 2) f.y += wall_term(__uint2float_rn(p.y) * ph.ky, ph);
    else f.y -= wall_term(__uint2float_rn(0xFFFFFFFFu - p.y) * ph.ky, ph);
    return f;
}

// Leapfrog kick + drift on half-step velocities with wrapping fixed-point positions
// (f_apply_force, particle.cuh:105-123): v += F/m dt; x += round(v dt / box * 2^32) (wrapping).
// The reference's divisions by the constants mass and box are multiplications by their
// reciprocals here (a 1-ulp difference, far inside the 1e-5 tolerance).
__device__ __forceinline__ void integrate(uint2 p, float2 v, float2 f, const Phys& ph, uint2& p_out, float2& v_out) {
    v_out.x = fmaf(f.x * ph.inv_mass, ph.dt, v.x);
    v_out.y = fmaf(f.y * ph.inv_mass, ph.dt, v.y);
    p_out.x = p.x + (uint32_t)(long long)roundf(v_out.x * ph.ux);
    p_out.y = p.y + (uint32_t)(long long)roundf(v_out.y * ph.uy);
}

// largest c in [0, count) with a[c] <= i, given a[0] <= i  (a is non-decreasing)
__device__ __forceinline__ int last_le(const uint32_t* a, int count, uint32_t i) {
    int lo = 0, hi = count;  // invariant: a[lo] <= i, (hi == count or a[hi] > i)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// The step kernel: force over the 3x3 cell stencil + kick + drift, one HBM round trip of the state.
//
// CTA b owns particles [b*kTile, (b+1)*kTile). Because the arrays are cell-sorted and cells are
// row-major, everything those particles interact with lies in three contiguous index ranges, one
// per stencil row: cells [first-1, last+1] shifted by -BX, 0, +BX. One thread reads the tile's
// descriptor and issues up to six 1-D bulk copies (TMA, cp.async.bulk -> mbarrier) that stage the
// cell_start entries and the positions of those ranges in shared memory; meanwhile every thread
// loads its own particle.  Then every thread walks its own three windows (cells cx-1..cx+1 of rows
// cy-1..cy+1, clipped at the grid edge exactly like kernel_bucket.cuh:74-77) in ascending index
// order, the (row, column, slot) order of the reference's loop.
// Tiles whose stencil does not fit the staging buffers (very sparse or very clustered spots) run
// the same code with the pointers aimed at global memory instead.
// ------------------------------------------------------------------------------------------------

struct StepArgs {
    const uint2* __restrict__ pos_in;
    uint2* __restrict__ pos_out;
    float2* __restrict__ vel;
    const uint32_t* __restrict__ cell_id;
    const uint32_t* __restrict__ cell_start;
    const TileDesc* __restrict__ tiles;
    uint32_t n;
    Grid g;
    Phys ph;
};

template <int KN, int FRAC, bool ANISO>
__device__ __forceinline__ void step_particle(uint32_t i, uint2 pi, float2 vi, uint32_t cell,
                                              const uint32_t* const cs[3], const uint32_t cs_lo[3],
                                              const uint2* const pp[3], const uint32_t pp_lo[3], const StepArgs& a) {
    const Grid& g = a.g;
    uint32_t cx = cell & (g.bx - 1), cy = cell >> g.lx;
    uint32_t x0 = cx == 0 ? 0 : cx - 1, x1 = cx == g.bx - 1 ? cx : cx + 1;
    float2 gx = splat(0.f), gy = splat(0.f);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int row = (int)cy + d - 1;
        if (row < 0 || row >= (int)g.by) continue;
        uint32_t c0 = ((uint32_t)row << g.lx) + x0, c1 = ((uint32_t)row << g.lx) + x1;
        uint32_t s = cs[d][c0 - cs_lo[d]], e = cs[d][c1 + 1 - cs_lo[d]];
        const uint2* win = pp[d] + (s - pp_lo[d]);  // window [s, e) of this row
        if (d == 1) window_accumulate<KN, FRAC, ANISO, true>(win, (int)(e - s), pi, a.ph, gx, gy);  // contains i
        else window_accumulate<KN, FRAC, ANISO, false>(win, (int)(e - s), pi, a.ph, gx, gy);
    }
    float2 f = field_force(pi, a.ph);
    f.x = fmaf(a.ph.pair_scale, gx.x + gx.y, f.x);
    f.y = fmaf(a.ph.pair_scale_y, gy.x + gy.y, f.y);
    uint2 po;
    float2 vo;
    integrate(pi, vi, f, a.ph, po, vo);
    a.pos_out[i] = po;
    a.vel[i] = vo;
}

template <int KN, int FRAC, bool ANISO>
__global__ void __launch_bounds__(kTile, 8) step_kernel(const StepArgs a) {
    __shared__ __align__(16) uint32_t s_cs[3][kCsCap];
    __shared__ __align__(16) uint2 s_pos[3][kPosCap];
    __shared__ __align__(8) uint64_t s_bar;

    const uint32_t b = blockIdx.x;
    const uint32_t i = b * kTile + threadIdx.x;
    const TileDesc t = a.tiles[b];

    if (t.fits) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            uint32_t bytes = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) bytes += t.cs_cnt[d] * 4u + t.p_cnt[d] * 8u;
            mbar_arrive_expect_tx(&s_bar, bytes);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (t.cs_cnt[d]) bulk_copy_g2s(s_cs[d], a.cell_start + t.cs_lo[d], t.cs_cnt[d] * 4u, &s_bar);
                if (t.p_cnt[d]) bulk_copy_g2s(s_pos[d], a.pos_in + t.p_lo[d], t.p_cnt[d] * 8u, &s_bar);
            }
        }
        __syncthreads();  // the barrier is initialised before anyone polls it
    }
    const bool live = i < a.n;
    uint2 pi = make_uint2(0, 0);
    float2 vi = make_float2(0.f, 0.f);
    uint32_t cell = 0;
    if (live) {
        pi = a.pos_in[i];
        vi = a.vel[i];
        cell = a.cell_id[i];
    }
    if (t.fits) {
        mbar_wait(&s_bar, 0);
        if (!live) return;
        const uint32_t* cs[3] = {s_cs[0], s_cs[1], s_cs[2]};
        const uint2* pp[3] = {s_pos[0], s_pos[1], s_pos[2]};
        step_particle<KN, FRAC, ANISO>(i, pi, vi, cell, cs, t.cs_lo, pp, t.p_lo, a);
    } else {
        if (!live) return;
        const uint32_t* cs[3] = {a.cell_start, a.cell_start, a.cell_start};
        const uint2* pp[3] = {a.pos_in, a.pos_in, a.pos_in};
        const uint32_t zero[3] = {0, 0, 0};
        step_particle<KN, FRAC, ANISO>(i, pi, vi, cell, cs, zero, pp, zero, a);
    }
}

// ------------------------------------------------------------------------------------------------
// Binning: stable counting sort by cell (count -> scan -> scatter -> order fix-up + gather).
//
//   key_count : key = cell(pos); rank = atomicAdd(count[key], 1)          (arbitrary rank in cell)
//   scan      : cell_start = exclusive prefix sum of count                 (3 small kernels)
//   scatter   : perm[cell_start[key] + rank] = source index
//   gather    : slot p holds source index i = perm[p]; its final place inside its cell is the number
//               of cell-mates with a smaller source index, which makes the result the STABLE sort
//               whatever order the atomics were served in -- the order the reference's serial
//               append (kernel.cuh:219-229) and its (row, column, slot) pull (kernel_bucket.cuh:17-33)
//               both produce.
// ------------------------------------------------------------------------------------------------

struct Source {  // where the particles to be binned come from
    const Particle* aos;  // ingest: records as they arrived (may contain nulls, ty < 0)
    const uint2* pos;     // re-bin: the live state
    const float2* vel;
    const int32_t* ty;
};

template <bool AOS>
__device__ __forceinline__ uint32_t source_key(const Source& s, uint32_t i, const Grid& g) {
    if (AOS) {
        const Particle& p = s.aos[i];
        if (p.ty < 0) return kNoKey;  // kernel.cuh:222
        return cell_of(make_uint2(p.x, p.y), g);
    }
    return cell_of(s.pos[i], g);
}

template <bool AOS>
__global__ void key_count_kernel(Source src, uint32_t count, Grid g, uint32_t* __restrict__ cell_count,
                                 uint32_t* __restrict__ rank) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t key = source_key<AOS>(src, i, g);
    if (key != kNoKey) rank[i] = atomicAdd(&cell_count[key], 1u);
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t* __restrict__ in, uint32_t count,
                                                                   uint32_t* __restrict__ block_sum) {
    __shared__ uint32_t warp_sum[kScanThreads / 32];
    uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < count) v += in[base + k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += warp_sum[w];
        block_sum[blockIdx.x] = t;
    }
}

// single block: exclusive scan of block_sum in place, total -> *total_out
__global__ void __launch_bounds__(1024) scan_top_kernel(uint32_t* __restrict__ block_sum, uint32_t blocks,
                                                        uint32_t* __restrict__ total_out) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < blocks; base += 1024) {
        uint32_t idx = base + threadIdx.x;
        uint32_t v = idx < blocks ? block_sum[idx] : 0;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sum[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if (threadIdx.x >= o) wi += t;
            }
            warp_sum[threadIdx.x] = wi - w;  // exclusive
        }
        __syncthreads();
        uint32_t excl = carry + warp_sum[threadIdx.x >> 5] + incl - v;
        if (idx < blocks) block_sum[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t count,
                                                                  const uint32_t* __restrict__ block_offset,
                                                                  uint32_t* __restrict__ out) {
    __shared__ uint32_t warp_sum[kScanThreads / 32];
    uint32_t base = blockIdx.x * kScanBlock + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < count ? in[base + k] : 0;
        t += v[k];
    }
    uint32_t incl = t;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += u;
    }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += warp_sum[w];
    uint32_t run = block_offset[blockIdx.x] + woff + incl - t;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < count) out[base + k] = run;
        run += v[k];
    }
}

template <bool AOS>
__global__ void scatter_kernel(Source src, uint32_t count, Grid g, const uint32_t* __restrict__ cell_start,
                               const uint32_t* __restrict__ rank, uint32_t* __restrict__ perm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t key = source_key<AOS>(src, i, g);
    if (key != kNoKey) perm[cell_start[key] + rank[i]] = i;
}

template <bool AOS>
__global__ void gather_kernel(Source src, uint32_t live, Grid g, const uint32_t* __restrict__ cell_start,
                              const uint32_t* __restrict__ perm, uint2* __restrict__ pos_out,
                              float2* __restrict__ vel_out, int32_t* __restrict__ ty_out,
                              uint32_t* __restrict__ cell_id_out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= live) return;
    uint32_t i = perm[p];
    uint2 pos;
    float2 vel;
    int32_t ty;
    if (AOS) {
        Particle q = src.aos[i];
        pos = make_uint2(q.x, q.y);
        vel = make_float2(q.vx, q.vy);
        ty = q.ty;
    } else {
        pos = src.pos[i];
        vel = src.vel[i];
        ty = src.ty[i];
    }
    uint32_t key = cell_of(pos, g);
    uint32_t s = cell_start[key], e = cell_start[key + 1];
    uint32_t r = 0;
    for (uint32_t k = s; k < e; ++k) r += perm[k] < i ? 1u : 0u;
    uint32_t dst = s + r;
    pos_out[dst] = pos;
    vel_out[dst] = vel;
    ty_out[dst] = ty;
    cell_id_out[dst] = key;
}

// One descriptor per tile of kTile consecutive particles: the cell range of the tile, and for each of
// the three stencil rows the (16-byte aligned) slices of cell_start and of the position array that its
// CTA stages in shared memory (see step_kernel).
__global__ void tile_desc_kernel(const uint32_t* __restrict__ cell_start, Grid g, uint32_t n,
                                 TileDesc* __restrict__ tiles) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ntiles = (n + kTile - 1) / kTile;
    if (b >= ntiles) return;
    uint32_t i0 = b * kTile, i1 = min(n, i0 + kTile) - 1;
    TileDesc t;
    t.first = (uint32_t)last_le(cell_start, (int)g.cells, i0);
    t.last = (uint32_t)last_le(cell_start, (int)g.cells, i1);
    t._pad = 0;
    bool fits = true;
    for (int d = 0; d < 3; ++d) {
        long long shift = (long long)(d - 1) * (long long)g.bx;
        long long lo = max((long long)t.first - 1 + shift, 0ll);
        long long hi = min((long long)t.last + 1 + shift, (long long)g.cells - 1);
        if (hi < lo) {  // the whole row lies outside the grid
            t.cs_lo[d] = t.cs_cnt[d] = t.p_lo[d] = t.p_cnt[d] = 0;
            continue;
        }
        uint32_t cs_lo = (uint32_t)lo & ~3u;
        uint32_t cs_cnt = (((uint32_t)hi + 2 - cs_lo) + 3u) & ~3u;  // entries lo .. hi+1
        uint32_t p_lo = cell_start[lo] & ~1u;
        uint32_t p_cnt = ((cell_start[hi + 1] - p_lo) + 1u) & ~1u;
        t.cs_lo[d] = cs_lo;
        t.cs_cnt[d] = cs_cnt;
        t.p_lo[d] = p_lo;
        t.p_cnt[d] = p_cnt;
        fits = fits && cs_cnt <= (uint32_t)kCsCap && p_cnt <= (uint32_t)kPosCap;
    }
    t.fits = fits ? 1u : 0u;
    tiles[b] = t;
}

// snapshot: pack the structure of arrays back into wire-format records (particle.rs:10-18)
__global__ void pack_kernel(const uint2* __restrict__ pos, const float2* __restrict__ vel,
                            const int32_t* __restrict__ ty, uint32_t n, Particle* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 p = pos[i];
    float2 v = vel[i];
    Particle q;
    q.x = p.x;
    q.y = p.y;
    q.vx = v.x;
    q.vy = v.y;
    q.ty = ty[i];
    out[i] = q;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------

struct PsimStepper {
    PsimConfig cfg{};
    Grid grid{};
    Phys phys{};
    int kernel_kn = 0;         // step-kernel variant: integer part of n/2+1 (0: run-time exponents)
    int kernel_frac = kFracEx2;
    bool kernel_aniso = false;
    FrameMetadata meta{};
    int device = 0;

    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;       // where the step loop runs (own_stream or the caller's)
    cudaStream_t copy_stream = nullptr;  // snapshot download
    cudaEvent_t snapshot_ready = nullptr;
    cudaEvent_t snapshot_consumed = nullptr;

    // live state
    uint2* pos[2] = {nullptr, nullptr};
    float2* vel[2] = {nullptr, nullptr};
    int32_t* ty[2] = {nullptr, nullptr};
    int cur_pos = 0, cur_vel = 0, cur_ty = 0;
    uint32_t* cell_start = nullptr;  // cells + 1
    uint32_t* cell_count = nullptr;  // cells
    uint32_t* block_sum = nullptr;
    uint32_t* rank = nullptr;
    uint32_t* perm = nullptr;
    uint32_t* cell_id = nullptr;  // n: cell of every particle as of the last binning
    TileDesc* tiles = nullptr;    // ceil(n / kTile)
    Particle* staging = nullptr;  // ingest (AoS) and snapshot (AoS) buffer
    uint32_t* h_total = nullptr;  // pinned

    uint32_t n = 0;            // live particles
    uint32_t snapshot_n = 0;   // particles in the packed snapshot
    FrameMetadata snapshot_meta{};
    bool has_scene = false;
    bool has_snapshot = false;
    int native_countdown = 0;  // native schedule: steps until the next re-bin

    uint64_t steps_executed = 0, rebins_executed = 0, launches = 0;

    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    size_t timing_used = 0;
    double timing_total_ms = 0;
    uint64_t timing_launches = 0;

    std::string error;
};

namespace {

std::string g_create_error;

int fail(PsimStepper* s, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (s) s->error = buf;
    else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t err__ = (call);                                                                       \
        if (err__ != cudaSuccess)                                                                         \
            return fail(s, PSIM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, \
                        __LINE__);                                                                        \
    } while (0)

// Cubic for 2^z on z in [z_lo, z_hi] (weighted least squares on Chebyshev nodes; the constant term is
// pinned to 1 so that z = 0 is exact). Returns the largest error of (n/m) q^(kn+fn) it causes, i.e. the
// absolute error of g, over the q range that maps onto [z_lo, z_hi].
double fit_exp2_cubic(double fn, double kn, double nm, double q_lo, double q_hi, float c[3]) {
    const int N = 256;
    const double ln2 = 0.69314718055994530942;
    // normal equations for e(z) - 1 = z (c1 + c2 z + c3 z^2), weight w = how much an error costs in g
    double A[3][3] = {{0}}, B[3] = {0};
    auto weight = [&](double q) { return std::min(nm * std::pow(q, kn + fn), 0.15) + 1e-6; };
    for (int k = 0; k < N; ++k) {
        double t = std::cos(3.14159265358979323846 * (k + 0.5) / N);
        double lq = 0.5 * (std::log2(q_hi) + std::log2(q_lo)) + 0.5 * (std::log2(q_hi) - std::log2(q_lo)) * t;
        double q = std::exp2(lq), z = fn * lq, w = weight(q);
        double basis[3] = {z, z * z, z * z * z}, rhs = std::exp(z * ln2) - 1.0;
        for (int r = 0; r < 3; ++r) {
            for (int cc = 0; cc < 3; ++cc) A[r][cc] += w * w * basis[r] * basis[cc];
            B[r] += w * w * basis[r] * rhs;
        }
    }
    // 3x3 solve (Gaussian elimination with partial pivoting)
    int idx[3] = {0, 1, 2};
    double M[3][4];
    for (int r = 0; r < 3; ++r) {
        for (int cc = 0; cc < 3; ++cc) M[r][cc] = A[r][cc];
        M[r][3] = B[r];
    }
    for (int col = 0; col < 3; ++col) {
        int piv = col;
        for (int r = col + 1; r < 3; ++r)
            if (std::fabs(M[r][col]) > std::fabs(M[piv][col])) piv = r;
        std::swap(idx[col], idx[piv]);
        for (int cc = 0; cc < 4; ++cc) std::swap(M[col][cc], M[piv][cc]);
        if (M[col][col] == 0.0) {  // fn == 0 or a degenerate range: Taylor coefficients
            c[0] = (float)ln2;
            c[1] = (float)(ln2 * ln2 / 2);
            c[2] = (float)(ln2 * ln2 * ln2 / 6);
            return fn == 0.0 ? 0.0 : 1.0;
        }
        for (int r = 0; r < 3; ++r) {
            if (r == col) continue;
            double f = M[r][col] / M[col][col];
            for (int cc = col; cc < 4; ++cc) M[r][cc] -= f * M[col][cc];
        }
    }
    for (int r = 0; r < 3; ++r) c[r] = (float)(M[r][3] / M[r][r]);
    double worst = 0;
    for (int k = 0; k <= 4 * N; ++k) {
        double lq = std::log2(q_lo) + (std::log2(q_hi) - std::log2(q_lo)) * k / (4.0 * N);
        double q = std::exp2(lq), z = fn * lq;
        double approx = 1.0 + z * ((double)c[0] + z * ((double)c[1] + z * (double)c[2]));
        double err = std::fabs(approx - std::exp(z * ln2));
        // absolute error of g, and -- where the repulsive term dominates -- relative to the term itself
        double cost = std::min(nm * std::pow(q, kn + fn) * err, err / 2e-7 * 3e-8);
        worst = std::max(worst, cost);
    }
    return worst;
}

Phys make_phys(const FrameMetadata& m, int* kernel_kn, int* kernel_frac, bool* aniso) {
    const MiePotentialParams& p = m.particles[0];  // kernel_bucket.cuh:52
    Phys ph{};
    const float two32 = 4294967296.f;
    ph.kx = m.box_width / two32;
    ph.ky = m.box_height / two32;
    ph.yscale = ph.ky / ph.kx;
    *aniso = ph.yscale != 1.f;
    ph.inv_c2 = (ph.kx / p.sigma) * (ph.kx / p.sigma);
    ph.nm = p.n / p.m;
    // exponents of q = sigma^2 / r^2: m/2 + 1 and n/2 + 1, split into the nearest integer and a rest
    float em = p.m / 2.f + 1.f, en = p.n / 2.f + 1.f;
    ph.km = (int)lrintf(em);
    ph.kn = (int)lrintf(en);
    ph.fm = em - (float)ph.km;
    ph.fn = en - (float)ph.kn;
    ph.m = p.m;
    float C = (p.n / (p.n - p.m)) * powf(p.n / p.m, p.m / (p.n - p.m));  // particle.cuh:53-55
    ph.pair_scale = C * p.epsilon * p.m * ph.kx / (p.sigma * p.sigma);
    ph.pair_scale_y = ph.pair_scale;
    ph.wall_scale = C * p.epsilon * p.m;
    ph.wall_m6 = p.m == 6.f;
    ph.sigma = p.sigma;
    ph.inv_mass = 1.f / (float)6.63352599e-26;  // particle.cuh:51
    ph.dt = m.step_dt;
    ph.ux = m.step_dt / m.box_width * two32;
    ph.uy = m.step_dt / m.box_height * two32;
    ph.cursor_x = m.cursor_pos[0];
    ph.cursor_y = m.cursor_pos[1];
    ph.cursor_r2 = m.cursor_size * m.cursor_size / 4;

    // Which step-kernel variant evaluates these exponents.
    //   m = 6 and 5 <= kn <= 10: the integer powers are compile-time products (KN = kn);
    //   the rest fn of the repulsive exponent: none / cubic in z / MUFU.EX2, see pair2().
    // The cubic is accepted only if the error it adds to g stays below 3e-8 (g is O(0.1..1); fp32
    // rounding of the terms themselves is ~1e-7) over r from 0.5 sigma to beyond the stencil reach.
    ph.c1 = 0.69314718f;
    ph.c2 = 0.24022651f;
    ph.c3 = 0.05550411f;
    bool fixed_powers = ph.km == 4 && ph.fm == 0.f && ph.kn >= 5 && ph.kn <= 10;
    *kernel_kn = fixed_powers ? ph.kn : 0;
    if (ph.fn == 0.f && (fixed_powers || ph.fm == 0.f)) {
        *kernel_frac = kFracNone;
    } else if (fixed_powers) {
        float c[3];
        double cost = fit_exp2_cubic(ph.fn, ph.kn, ph.nm, 1e-3, 4.0, c);
        if (cost <= 3e-8) {
            ph.c1 = c[0];
            ph.c2 = c[1];
            ph.c3 = c[2];
            *kernel_frac = kFracPoly;
        } else {
            *kernel_frac = kFracEx2;
        }
    } else {
        *kernel_frac = kFracEx2;
    }
    return ph;
}

void apply_metadata(PsimStepper* s, const FrameMetadata& m) {
    s->meta = m;
    s->phys = make_phys(m, &s->kernel_kn, &s->kernel_frac, &s->kernel_aniso);
}

template <int KN, int FRAC>
void launch_step_aniso(PsimStepper* s, const StepArgs& a, uint32_t tiles) {
    if (s->kernel_aniso) step_kernel<KN, FRAC, true><<<tiles, kTile, 0, s->stream>>>(a);
    else step_kernel<KN, FRAC, false><<<tiles, kTile, 0, s->stream>>>(a);
}

template <int KN>
void launch_step_frac(PsimStepper* s, const StepArgs& a, uint32_t tiles) {
    switch (s->kernel_frac) {
        case kFracNone: launch_step_aniso<KN, kFracNone>(s, a, tiles); break;
        case kFracPoly: launch_step_aniso<KN, kFracPoly>(s, a, tiles); break;
        default: launch_step_aniso<KN, kFracEx2>(s, a, tiles); break;
    }
}

void launch_step(PsimStepper* s, const StepArgs& a, uint32_t tiles) {
    switch (s->kernel_kn) {
        case 5: launch_step_frac<5>(s, a, tiles); break;
        case 6: launch_step_frac<6>(s, a, tiles); break;
        case 7: launch_step_frac<7>(s, a, tiles); break;
        case 8: launch_step_frac<8>(s, a, tiles); break;
        case 9: launch_step_frac<9>(s, a, tiles); break;
        case 10: launch_step_frac<10>(s, a, tiles); break;
        default:  // run-time exponents; kFracPoly is never selected for them
            if (s->kernel_frac == kFracNone) launch_step_aniso<0, kFracNone>(s, a, tiles);
            else launch_step_aniso<0, kFracEx2>(s, a, tiles);
            break;
    }
}

inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// cell_start = exclusive scan of cell_count; total (live particles) -> cell_start[cells]
int enqueue_scan(PsimStepper* s) {
    uint32_t cells = s->grid.cells;
    uint32_t blocks = div_up(cells, kScanBlock);
    scan_reduce_kernel<<<blocks, kScanThreads, 0, s->stream>>>(s->cell_count, cells, s->block_sum);
    scan_top_kernel<<<1, 1024, 0, s->stream>>>(s->block_sum, blocks, s->cell_start + cells);
    scan_apply_kernel<<<blocks, kScanThreads, 0, s->stream>>>(s->cell_count, cells, s->block_sum, s->cell_start);
    s->launches += 3;
    CK(cudaGetLastError());
    return PSIM_OK;
}

int enqueue_tiles(PsimStepper* s) {
    uint32_t tiles = div_up(s->n, kTile);
    if (tiles == 0) return PSIM_OK;
    tile_desc_kernel<<<div_up(tiles, 128), 128, 0, s->stream>>>(s->cell_start, s->grid, s->n, s->tiles);
    s->launches += 1;
    CK(cudaGetLastError());
    return PSIM_OK;
}

// Re-bin the live state (bucket_move, kernel_bucket.cuh:5-39).
int enqueue_rebin(PsimStepper* s) {
    if (s->n == 0) return PSIM_OK;
    const uint32_t n = s->n, tb = 256;
    Source src{nullptr, s->pos[s->cur_pos], s->vel[s->cur_vel], s->ty[s->cur_ty]};
    CK(cudaMemsetAsync(s->cell_count, 0, sizeof(uint32_t) * s->grid.cells, s->stream));
    key_count_kernel<false><<<div_up(n, tb), tb, 0, s->stream>>>(src, n, s->grid, s->cell_count, s->rank);
    s->launches += 1;
    int rc = enqueue_scan(s);
    if (rc) return rc;
    scatter_kernel<false><<<div_up(n, tb), tb, 0, s->stream>>>(src, n, s->grid, s->cell_start, s->rank, s->perm);
    gather_kernel<false><<<div_up(n, tb), tb, 0, s->stream>>>(src, n, s->grid, s->cell_start, s->perm,
                                                              s->pos[s->cur_pos ^ 1], s->vel[s->cur_vel ^ 1],
                                                              s->ty[s->cur_ty ^ 1], s->cell_id);
    s->launches += 2;
    CK(cudaGetLastError());
    s->cur_pos ^= 1;
    s->cur_vel ^= 1;
    s->cur_ty ^= 1;
    rc = enqueue_tiles(s);
    if (rc) return rc;
    s->rebins_executed += 1;
    return PSIM_OK;
}

int enqueue_step(PsimStepper* s) {
    if (s->n == 0) {
        s->steps_executed += 1;
        return PSIM_OK;
    }
    StepArgs a;
    a.pos_in = s->pos[s->cur_pos];
    a.pos_out = s->pos[s->cur_pos ^ 1];
    a.vel = s->vel[s->cur_vel];
    a.cell_id = s->cell_id;
    a.cell_start = s->cell_start;
    a.tiles = s->tiles;
    a.n = s->n;
    a.g = s->grid;
    a.ph = s->phys;
    uint32_t tiles = div_up(s->n, kTile);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (s->timing) {
        if (s->timing_used == s->timing_events.size()) {
            cudaEvent_t x, y;
            CK(cudaEventCreate(&x));
            CK(cudaEventCreate(&y));
            s->timing_events.push_back({x, y});
        }
        e0 = s->timing_events[s->timing_used].first;
        e1 = s->timing_events[s->timing_used].second;
        s->timing_used += 1;
        CK(cudaEventRecord(e0, s->stream));
    }
    launch_step(s, a, tiles);
    if (s->timing) CK(cudaEventRecord(e1, s->stream));
    CK(cudaGetLastError());
    s->launches += 1;
    s->cur_pos ^= 1;
    s->steps_executed += 1;
    return PSIM_OK;
}

int enqueue_snapshot(PsimStepper* s) {
    // the previous snapshot must have left the staging buffer before it is overwritten
    CK(cudaStreamWaitEvent(s->stream, s->snapshot_consumed, 0));
    if (s->n) {
        pack_kernel<<<div_up(s->n, 256), 256, 0, s->stream>>>(s->pos[s->cur_pos], s->vel[s->cur_vel], s->ty[s->cur_ty],
                                                              s->n, s->staging);
        s->launches += 1;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(s->snapshot_ready, s->stream));
    s->snapshot_n = s->n;
    s->snapshot_meta = s->meta;
    s->has_snapshot = true;
    return PSIM_OK;
}

int collect_timing(PsimStepper* s) {
    for (size_t k = 0; k < s->timing_used; ++k) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, s->timing_events[k].first, s->timing_events[k].second));
        s->timing_total_ms += ms;
        s->timing_launches += 1;
    }
    s->timing_used = 0;
    return PSIM_OK;
}

// Bin `count` wire-format records that sit in s->staging (device).
int ingest_staged(PsimStepper* s, uint32_t count) {
    const uint32_t tb = 256;
    Source src{s->staging, nullptr, nullptr, nullptr};
    CK(cudaMemsetAsync(s->cell_count, 0, sizeof(uint32_t) * s->grid.cells, s->stream));
    if (count) {
        key_count_kernel<true><<<div_up(count, tb), tb, 0, s->stream>>>(src, count, s->grid, s->cell_count, s->rank);
        s->launches += 1;
    }
    int rc = enqueue_scan(s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s->h_total, s->cell_start + s->grid.cells, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                       s->stream));
    CK(cudaStreamSynchronize(s->stream));
    uint32_t live = *s->h_total;
    if (live > s->cfg.max_particles)
        return fail(s, PSIM_ECAPACITY, "%u live particles exceed max_particles = %u", live, s->cfg.max_particles);
    s->n = live;
    s->cur_pos = s->cur_vel = s->cur_ty = 0;
    if (live) {
        scatter_kernel<true><<<div_up(count, tb), tb, 0, s->stream>>>(src, count, s->grid, s->cell_start, s->rank,
                                                                      s->perm);
        gather_kernel<true><<<div_up(live, tb), tb, 0, s->stream>>>(src, live, s->grid, s->cell_start, s->perm,
                                                                    s->pos[0], s->vel[0], s->ty[0], s->cell_id);
        s->launches += 2;
        CK(cudaGetLastError());
    }
    rc = enqueue_tiles(s);
    if (rc) return rc;
    s->has_scene = true;
    s->native_countdown = 0;
    rc = enqueue_snapshot(s);  // the ingested scene itself can be downloaded (cuda_simulator.cu:28-31)
    if (rc) return rc;
    CK(cudaStreamSynchronize(s->stream));
    return PSIM_OK;
}

}  // namespace

extern "C" {

PsimConfig psim_default_config(void) {
    PsimConfig c;
    std::memset(&c, 0, sizeof c);
    c.grid_x_log2 = 6;  // kernel.cuh:15-16
    c.grid_y_log2 = 6;
    c.max_particles = 65536;  // kernel.cuh:20
    c.schedule = PSIM_SCHEDULE_REFERENCE;
    c.rebin_every = 0;
    c.device = -1;
    c.use_graph = 0;
    return c;
}

const char* psim_last_error(const PsimStepper* s) { return s ? s->error.c_str() : g_create_error.c_str(); }

void psim_destroy(PsimStepper* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    for (auto& ev : s->timing_events) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    for (int k = 0; k < 2; ++k) {
        cudaFree(s->pos[k]);
        cudaFree(s->vel[k]);
        cudaFree(s->ty[k]);
    }
    cudaFree(s->cell_start);
    cudaFree(s->cell_count);
    cudaFree(s->block_sum);
    cudaFree(s->rank);
    cudaFree(s->perm);
    cudaFree(s->cell_id);
    cudaFree(s->tiles);
    cudaFree(s->staging);
    if (s->h_total) cudaFreeHost(s->h_total);
    if (s->snapshot_ready) cudaEventDestroy(s->snapshot_ready);
    if (s->snapshot_consumed) cudaEventDestroy(s->snapshot_consumed);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    delete s;
}

int psim_create(const PsimConfig* config, PsimStepper** out) {
    PsimStepper* s = nullptr;  // for CK / fail before the object exists
    if (!config || !out) return fail(s, PSIM_EINVAL, "psim_create: null argument");
    *out = nullptr;
    if (config->grid_x_log2 > 15 || config->grid_y_log2 > 15 || config->grid_x_log2 + config->grid_y_log2 > 28)
        return fail(s, PSIM_EINVAL, "psim_create: grid 2^%u x 2^%u is out of range", config->grid_x_log2,
                    config->grid_y_log2);
    // separations inside the 3x3 stencil must fit a signed 32-bit fixed-point difference
    if (config->grid_x_log2 < 3 || config->grid_y_log2 < 3)
        return fail(s, PSIM_EINVAL, "psim_create: the grid needs at least 8 cells per axis");
    if (config->max_particles == 0 || config->max_particles > 0x7FFFFF00u)
        return fail(s, PSIM_EINVAL, "psim_create: max_particles out of range");
    if (config->schedule > PSIM_SCHEDULE_NATIVE) return fail(s, PSIM_EINVAL, "psim_create: unknown schedule");

    int device = config->device;
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0)
        return fail(s, PSIM_ECUDA, "psim_create: no CUDA device (%s); this library has no CPU path",
                    err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= ndev) return fail(s, PSIM_EINVAL, "psim_create: device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(s, PSIM_ECUDA, "psim_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);

    PsimStepper* st = new PsimStepper;
    st->cfg = *config;
    if (st->cfg.rebin_every == 0) st->cfg.rebin_every = 17;
    st->device = device;
    Grid& g = st->grid;
    g.lx = config->grid_x_log2;
    g.ly = config->grid_y_log2;
    g.bx = 1u << g.lx;
    g.by = 1u << g.ly;
    g.cells = g.bx * g.by;
    g.sx = 32 - g.lx;
    g.sy = 32 - g.ly;
    s = st;  // from here on failures are recorded on the object (and it is destroyed before returning)
#define CKC(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t err__ = (call);                                                                      \
        if (err__ != cudaSuccess) {                                                                      \
            fail(nullptr, PSIM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, \
                 __LINE__);                                                                              \
            psim_destroy(st);                                                                            \
            return PSIM_ECUDA;                                                                           \
        }                                                                                                \
    } while (0)
    const size_t cap = config->max_particles;
    CKC(cudaStreamCreateWithFlags(&st->own_stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking));
    st->stream = st->own_stream;
    CKC(cudaEventCreateWithFlags(&st->snapshot_ready, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&st->snapshot_consumed, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) {
        CKC(cudaMalloc(&st->pos[k], sizeof(uint2) * (cap + kPadParticles)));
        CKC(cudaMemset(st->pos[k], 0, sizeof(uint2) * (cap + kPadParticles)));
        CKC(cudaMalloc(&st->vel[k], sizeof(float2) * cap));
        CKC(cudaMalloc(&st->ty[k], sizeof(int32_t) * cap));
    }
    CKC(cudaMalloc(&st->cell_start, sizeof(uint32_t) * ((size_t)g.cells + 1 + kPadCells)));
    CKC(cudaMalloc(&st->cell_count, sizeof(uint32_t) * (size_t)g.cells));
    CKC(cudaMalloc(&st->block_sum, sizeof(uint32_t) * (size_t)div_up(g.cells, kScanBlock)));
    CKC(cudaMalloc(&st->rank, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&st->perm, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&st->cell_id, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&st->tiles, sizeof(TileDesc) * (size_t)div_up((uint32_t)cap, kTile)));
    CKC(cudaMalloc(&st->staging, sizeof(Particle) * cap));
    CKC(cudaMallocHost(&st->h_total, sizeof(uint32_t)));
    CKC(cudaMemset(st->cell_start, 0, sizeof(uint32_t) * ((size_t)g.cells + 1 + kPadCells)));
#undef CKC
    *out = st;
    return PSIM_OK;
}

int psim_set_stream(PsimStepper* s, void* cuda_stream) {
    if (!s) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    s->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->own_stream;
    return PSIM_OK;
}

int psim_upload_frame(PsimStepper* s, const FrameHeader* frame) {
    if (!s || !frame) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    if (frame->particle_count > s->cfg.max_particles)
        return fail(s, PSIM_ECAPACITY, "frame holds %u particles, max_particles = %u", frame->particle_count,
                    s->cfg.max_particles);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaStreamSynchronize(s->copy_stream));
    apply_metadata(s, frame->metadata);
    if (frame->particle_count)
        CK(cudaMemcpyAsync(s->staging, frame->particles, sizeof(Particle) * (size_t)frame->particle_count,
                           cudaMemcpyHostToDevice, s->stream));
    return ingest_staged(s, frame->particle_count);
}

int psim_upload_device(PsimStepper* s, const FrameMetadata* meta, const void* d_particles, uint32_t count) {
    if (!s || !meta || (count && !d_particles)) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    if (count > s->cfg.max_particles)
        return fail(s, PSIM_ECAPACITY, "%u particles, max_particles = %u", count, s->cfg.max_particles);
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaStreamSynchronize(s->copy_stream));
    apply_metadata(s, *meta);
    if (count)
        CK(cudaMemcpyAsync(s->staging, d_particles, sizeof(Particle) * (size_t)count, cudaMemcpyDeviceToDevice,
                           s->stream));
    return ingest_staged(s, count);
}

int psim_set_metadata(PsimStepper* s, const FrameMetadata* meta) {
    if (!s || !meta) return PSIM_EINVAL;
    apply_metadata(s, *meta);  // captured by value at the next enqueue, like kernel_bucket.cuh:121
    return PSIM_OK;
}

int psim_get_metadata(const PsimStepper* s, FrameMetadata* out) {
    if (!s || !out) return PSIM_EINVAL;
    *out = s->meta;
    return PSIM_OK;
}

int psim_step_async(PsimStepper* s, uint32_t steps) {
    if (!s) return PSIM_EINVAL;
    if (!s->has_scene) return fail(s, PSIM_ESTATE, "psim_step_async: no scene uploaded");
    CK(cudaSetDevice(s->device));
    for (uint32_t k = 0; k < steps; ++k) {
        int rc = enqueue_step(s);
        if (rc) return rc;
    }
    return PSIM_OK;
}

int psim_rebin_async(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    if (!s->has_scene) return fail(s, PSIM_ESTATE, "psim_rebin_async: no scene uploaded");
    CK(cudaSetDevice(s->device));
    return enqueue_rebin(s);
}

int psim_snapshot_async(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    if (!s->has_scene) return fail(s, PSIM_ESTATE, "psim_snapshot_async: no scene uploaded");
    CK(cudaSetDevice(s->device));
    return enqueue_snapshot(s);
}

int psim_run_frame_async(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    if (!s->has_scene) return fail(s, PSIM_ESTATE, "psim_run_frame_async: no scene uploaded");
    CK(cudaSetDevice(s->device));
    const uint32_t target = s->meta.steps_per_frame;
    int rc;
    if (s->cfg.schedule == PSIM_SCHEDULE_REFERENCE) {
        // bucket_kernel_run_async, kernel_bucket.cuh:181-206: the reference always runs one step,
        // then alternates "re-bin + 1 step" with pairs of steps, 16 steps between re-bins counted
        // from the first re-bin, the countdown restarting with every frame. Pairs make it overshoot
        // an odd remainder by one step.
        const int move_every_n = 16;
        int countdown = 0;
        uint32_t steps = 0;
        if ((rc = enqueue_step(s))) return rc;
        steps += 1;
        while (steps < target) {
            if (countdown <= 0) {
                if ((rc = enqueue_rebin(s))) return rc;
                countdown = move_every_n;
                if ((rc = enqueue_step(s))) return rc;
                countdown -= 1;
                steps += 1;
            } else {
                if ((rc = enqueue_step(s))) return rc;
                if ((rc = enqueue_step(s))) return rc;
                countdown -= 2;
                steps += 2;
            }
        }
    } else {
        for (uint32_t k = 0; k < target; ++k) {
            if (s->native_countdown <= 0) {
                // a freshly ingested scene is already binned
                if (s->steps_executed != 0 || s->rebins_executed != 0 || k != 0) {
                    if ((rc = enqueue_rebin(s))) return rc;
                }
                s->native_countdown = (int)s->cfg.rebin_every;
            }
            if ((rc = enqueue_step(s))) return rc;
            s->native_countdown -= 1;
        }
    }
    return enqueue_snapshot(s);
}

int psim_sync(PsimStepper* s) {
    if (!s) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    if (s->timing) return collect_timing(s);
    return PSIM_OK;
}

int psim_download_frame(PsimStepper* s, FrameHeader* dst) {
    if (!s || !dst) return PSIM_EINVAL;
    if (!s->has_snapshot) return fail(s, PSIM_ESTATE, "psim_download_frame: no snapshot has been packed");
    CK(cudaSetDevice(s->device));
    if (dst->particle_count < s->snapshot_n)
        return fail(s, PSIM_ECAPACITY, "psim_download_frame: destination holds %u particles, snapshot has %u",
                    dst->particle_count, s->snapshot_n);
    CK(cudaStreamWaitEvent(s->copy_stream, s->snapshot_ready, 0));
    if (s->snapshot_n)
        CK(cudaMemcpyAsync(dst->particles, s->staging, sizeof(Particle) * (size_t)s->snapshot_n,
                           cudaMemcpyDeviceToHost, s->copy_stream));
    CK(cudaEventRecord(s->snapshot_consumed, s->copy_stream));
    CK(cudaStreamSynchronize(s->copy_stream));
    // FrameHeader::new (particle.rs:214-223)
    static const uint8_t sig0[4] = {0x36, 0xbc, 0xe9, 0xbd}, sig1[4] = {0xac, 0xc4, 0x12, 0xec};
    std::memcpy(dst->signature_start, sig0, 4);
    std::memcpy(dst->signature_end, sig1, 4);
    dst->_padding = 0;
    dst->metadata = s->snapshot_meta;
    dst->particle_count = s->snapshot_n;
    return PSIM_OK;
}

uint32_t psim_particle_count(const PsimStepper* s) { return s ? s->n : 0; }
uint64_t psim_steps_executed(const PsimStepper* s) { return s ? s->steps_executed : 0; }
uint64_t psim_rebins_executed(const PsimStepper* s) { return s ? s->rebins_executed : 0; }
uint64_t psim_kernel_launches(const PsimStepper* s) { return s ? s->launches : 0; }
uint32_t psim_cell_count(const PsimStepper* s) { return s ? s->grid.cells : 0; }

int psim_get_cell_start(PsimStepper* s, uint32_t* out) {
    if (!s || !out) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy(out, s->cell_start, sizeof(uint32_t) * ((size_t)s->grid.cells + 1), cudaMemcpyDeviceToHost));
    return PSIM_OK;
}

int psim_enable_step_timing(PsimStepper* s, int enable) {
    if (!s) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    int rc = collect_timing(s);
    if (rc) return rc;
    s->timing = enable != 0;
    s->timing_total_ms = 0;
    s->timing_launches = 0;
    return PSIM_OK;
}

int psim_get_step_timing(PsimStepper* s, double* total_ms, uint64_t* launches) {
    if (!s) return PSIM_EINVAL;
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    int rc = collect_timing(s);
    if (rc) return rc;
    if (total_ms) *total_ms = s->timing_total_ms;
    if (launches) *launches = s->timing_launches;
    return PSIM_OK;
}

int psim_device_state(PsimStepper* s, const void** pos, const void** vel, const void** ty, const void** cell_start) {
    if (!s) return PSIM_EINVAL;
    if (pos) *pos = s->pos[s->cur_pos];
    if (vel) *vel = s->vel[s->cur_vel];
    if (ty) *ty = s->ty[s->cur_ty];
    if (cell_start) *cell_start = s->cell_start;
    return PSIM_OK;
}

}  // extern "C"
