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
//   tile_first/tile_last[ceil(n/P)] first / last cell touched by each tile of P consecutive particles
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
constexpr int kCsCap = 288;       // cell_start entries staged per stencil row (cells spanned + 3)
constexpr int kPosCap = 640;      // neighbour positions staged per stencil row
constexpr int kScanItems = 8;     // cells per thread in the scan kernels
constexpr int kScanThreads = 256;
constexpr int kScanBlock = kScanItems * kScanThreads;
constexpr uint32_t kNoKey = 0xFFFFFFFFu;

struct Grid {
    uint32_t lx, ly;    // log2 cells in x / y
    uint32_t bx, by;    // cells in x / y
    uint32_t cells;     // bx * by
    uint32_t sx, sy;    // 32 - lx, 32 - ly (shift that maps a fixed-point coordinate to its cell)
};

// Everything a step needs from FrameMetadata, pre-digested on the host once per metadata change
// (the reference rebuilds ParticleParams, including a powf, in every thread of every step:
// kernel_bucket.cuh:52, particle.cuh:53-55).
struct Phys {
    float ax, ay;        // (box / 2^32) / sigma : fixed-point units -> separation in units of sigma
    float kx, ky;        // box / 2^32           : fixed-point units -> metres
    float n, m;          // Mie exponents of species 0 (the only ones the reference uses)
    float fn, fm;        // fractional parts of n/2 and m/2
    int kn, km;          // integer parts of n/2 and m/2
    float pair_scale;    // C * eps / sigma   : scaled pair sum -> newtons
    float wall_scale;    // C * eps * m
    float sigma;
    float mass;
    float dt;
    float box_w, box_h;
    float cursor_x, cursor_y, cursor_r2;  // cursor_r2 = cursor_size^2 / 4
};

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

__device__ __forceinline__ float powi(float b, int k) {  // k is uniform across the grid
    float r = 1.f;
    while (k) {
        if (k & 1) r *= b;
        b *= b;
        k >>= 1;
    }
    return r;
}

__device__ __forceinline__ uint32_t cell_of(uint2 p, const Grid& g) {
    // kernel.cuh:224-226; a shift by 32 is undefined, so a 1-cell axis is handled explicitly
    uint32_t cx = g.lx ? p.x >> g.sx : 0u;
    uint32_t cy = g.ly ? p.y >> g.sy : 0u;
    return cx + (cy << g.lx);
}

// One pair: separation i -> j in units of sigma (f_dist, particle.cuh:41-47), Mie force
// (particle.cuh:63-66,97-103) written on r^2 so that no square root is needed:
//   F_vec = C eps (m (s/r)^m - n (s/r)^n) / r^2 * r_vec = (C eps / sigma) * (m q^(m/2) - n q^(n/2)) q * r'_vec
// with r' = r / sigma and q = 1 / r'^2.  q^(n/2) = q^kn * 2^(fn * log2 q): the integer part by
// multiplication, only the small fractional part through the approximate MUFU units, which keeps
// the relative error of each term at a few 1e-7 (the fp32 evaluation in the reference, sigma/len
// rounded and raised to the 14th power, is no better).
template <bool FAST>
__device__ __forceinline__ void pair_accumulate(uint2 pi, uint2 pj, const Phys& ph, float& gx, float& gy) {
    float rx = __int2float_rn((int)(pj.x - pi.x)) * ph.ax;
    float ry = __int2float_rn((int)(pj.y - pi.y)) * ph.ay;
    float r2 = fmaf(ry, ry, rx * rx);
    float q = fast_rcp(r2);
    float l = fast_lg2(q);
    float pm, pn;
    if (FAST) {  // m = 6, floor(n/2) = 7 (the reference's default nitrogen parameters)
        float q2 = q * q;
        pm = q2 * q;
        pn = (pm * pm) * q * fast_ex2(ph.fn * l);
    } else {
        pm = powi(q, ph.km);
        if (ph.fm != 0.f) pm *= fast_ex2(ph.fm * l);
        pn = powi(q, ph.kn);
        if (ph.fn != 0.f) pn *= fast_ex2(ph.fn * l);
    }
    float g = fmaf(ph.m, pm, -ph.n * pn) * q;
    gx = fmaf(g, rx, gx);
    gy = fmaf(g, ry, gy);
}

template <bool FAST>
__device__ __forceinline__ void range_accumulate(const uint2* __restrict__ pj, int count, uint2 pi, const Phys& ph,
                                                 float& gx, float& gy) {
#pragma unroll 4
    for (int k = 0; k < count; ++k) pair_accumulate<FAST>(pi, pj[k], ph, gx, gy);
}

// Repulsive wall term C eps m (sigma/d)^m / d (particle.cuh:68-71).
__device__ __forceinline__ float wall_term(float d, const Phys& ph) {
    float q = ph.sigma / d;
    float pw;
    if (ph.km == 3 && ph.fm == 0.f) {
        float q2 = q * q;
        pw = q2 * q2 * q2;
    } else {
        pw = fast_ex2(ph.m * fast_lg2(q));
    }
    return ph.wall_scale * pw / d;
}

// Cursor + wall forces on one particle (kernel_bucket.cuh:54-69, particle.cuh:125-144).
__device__ __forceinline__ float2 field_force(uint2 p, const Phys& ph) {
    const float inv32 = 1.f / 4294967296.f;
    float2 f = make_float2(0.f, 0.f);
    float dx = ph.cursor_x - __uint2float_rn(p.x) * inv32;
    float dy = ph.cursor_y - __uint2float_rn(p.y) * inv32;
    float sq = dx * dx + dy * dy;
    if (sq < ph.cursor_r2) {
        float c = 8e-12f / (sq + 1.f);
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
// (f_apply_force, particle.cuh:105-123). Per particle, so the exact divisions are kept.
__device__ __forceinline__ void integrate(uint2 p, float2 v, float2 f, const Phys& ph, uint2& p_out, float2& v_out) {
    const float two32 = 4294967296.f;
    float axl = f.x / ph.mass;
    float ayl = f.y / ph.mass;
    v_out.x = v.x + axl * ph.dt;
    v_out.y = v.y + ayl * ph.dt;
    float dx = v_out.x * ph.dt;
    float dy = v_out.y * ph.dt;
    p_out.x = p.x + (uint32_t)(long long)roundf((dx / ph.box_w) * two32);
    p_out.y = p.y + (uint32_t)(long long)roundf((dy / ph.box_h) * two32);
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
// per stencil row: cells [first-1, last+1] shifted by -BX, 0, +BX. The CTA stages the cell_start
// entries and the positions of those three ranges in shared memory once, then every thread walks
// its own three windows (cells cx-1..cx+1 of rows cy-1..cy+1, clipped at the grid edge exactly like
// kernel_bucket.cuh:74-77) in ascending index order -- the same (row, column, slot) order in which
// the reference accumulates, so the fp32 sum is formed in the same sequence.
// Tiles whose stencil does not fit the staging buffers (very sparse or very clustered spots) take
// the same code path with the pointers aimed at global memory instead.
// ------------------------------------------------------------------------------------------------

struct StepArgs {
    const uint2* __restrict__ pos_in;
    uint2* __restrict__ pos_out;
    float2* __restrict__ vel;
    const uint32_t* __restrict__ cell_start;
    const uint32_t* __restrict__ tile_first;
    const uint32_t* __restrict__ tile_last;
    uint32_t n;
    Grid g;
    Phys ph;
};

template <bool FAST>
__device__ __forceinline__ void step_particle(uint32_t i, uint32_t cell, const uint32_t* const cs[3],
                                              const int cs_lo[3], const uint2* const pp[3], const uint32_t pp_lo[3],
                                              const StepArgs& a) {
    const Grid& g = a.g;
    uint2 pi = a.pos_in[i];
    float2 vi = a.vel[i];
    uint32_t cx = cell & (g.bx - 1), cy = cell >> g.lx;
    uint32_t x0 = cx == 0 ? 0 : cx - 1, x1 = cx == g.bx - 1 ? cx : cx + 1;
    float gx = 0.f, gy = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int row = (int)cy + d - 1;
        if (row < 0 || row >= (int)g.by) continue;
        int c0 = (row << g.lx) + (int)x0, c1 = (row << g.lx) + (int)x1;
        uint32_t s = cs[d][c0 - cs_lo[d]], e = cs[d][c1 + 1 - cs_lo[d]];
        const uint2* win = pp[d] + (s - pp_lo[d]);  // window [s, e) of this row
        if (d == 1) {  // own row: skip j == i (kernel_bucket.cuh:85)
            range_accumulate<FAST>(win, (int)(i - s), pi, a.ph, gx, gy);
            range_accumulate<FAST>(win + (i + 1 - s), (int)(e - i - 1), pi, a.ph, gx, gy);
        } else {
            range_accumulate<FAST>(win, (int)(e - s), pi, a.ph, gx, gy);
        }
    }
    float2 f = field_force(pi, a.ph);
    f.x = fmaf(a.ph.pair_scale, gx, f.x);
    f.y = fmaf(a.ph.pair_scale, gy, f.y);
    uint2 po;
    float2 vo;
    integrate(pi, vi, f, a.ph, po, vo);
    a.pos_out[i] = po;
    a.vel[i] = vo;
}

template <bool FAST>
__global__ void __launch_bounds__(kTile) step_kernel(const StepArgs a) {
    __shared__ uint32_t s_cs[3][kCsCap];
    __shared__ uint2 s_pos[3][kPosCap];
    __shared__ int s_fits;

    const Grid& g = a.g;
    const uint32_t b = blockIdx.x;
    const uint32_t i = b * kTile + threadIdx.x;
    const int first = (int)a.tile_first[b], last = (int)a.tile_last[b];

    // linear cell range of each stencil row, clipped to the grid
    int lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int shift = (d - 1) * (int)g.bx;
        lo[d] = max(first - 1 + shift, 0);
        hi[d] = min(last + 1 + shift, (int)g.cells - 1);
    }
    const bool cs_fits = last - first + 4 <= kCsCap;
    if (threadIdx.x == 0) s_fits = cs_fits ? 1 : 0;
    if (cs_fits) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            int cnt = hi[d] - lo[d] + 2;  // entries lo..hi+1 (may be <= 0 for a clipped-away row)
            for (int k = threadIdx.x; k < cnt; k += kTile) s_cs[d][k] = a.cell_start[lo[d] + k];
        }
    }
    __syncthreads();
    uint32_t plo[3] = {0, 0, 0};
    if (cs_fits) {
        bool fits = true;
        uint32_t pcnt[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (hi[d] >= lo[d]) {
                plo[d] = s_cs[d][0];
                pcnt[d] = s_cs[d][hi[d] - lo[d] + 1] - plo[d];
            } else {
                pcnt[d] = 0;
            }
            fits = fits && pcnt[d] <= (uint32_t)kPosCap;
        }
        if (fits) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
                for (uint32_t k = threadIdx.x; k < pcnt[d]; k += kTile) s_pos[d][k] = a.pos_in[plo[d] + k];
        } else if (threadIdx.x == 0) {
            s_fits = 0;  // every thread computes the same `fits`; one writer is enough
        }
    }
    __syncthreads();
    if (i >= a.n) return;

    if (s_fits) {
        const uint32_t* cs[3] = {s_cs[0], s_cs[1], s_cs[2]};
        const uint2* pp[3] = {s_pos[0], s_pos[1], s_pos[2]};
        // own cell: the cell c in [first, last] with cell_start[c] <= i < cell_start[c+1]
        int off = first - lo[1];
        uint32_t cell = (uint32_t)(first + last_le(s_cs[1] + off, last - first + 1, i));
        step_particle<FAST>(i, cell, cs, lo, pp, plo, a);
    } else {
        const uint32_t* cs[3] = {a.cell_start, a.cell_start, a.cell_start};
        const uint2* pp[3] = {a.pos_in, a.pos_in, a.pos_in};
        const int zero[3] = {0, 0, 0};
        const uint32_t uzero[3] = {0, 0, 0};
        uint32_t cell = (uint32_t)(first + last_le(a.cell_start + first, last - first + 1, i));
        step_particle<FAST>(i, cell, cs, zero, pp, uzero, a);
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
                              float2* __restrict__ vel_out, int32_t* __restrict__ ty_out) {
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
}

// first / last cell of every tile of kTile consecutive particles
__global__ void tile_cells_kernel(const uint32_t* __restrict__ cell_start, uint32_t cells, uint32_t n,
                                  uint32_t* __restrict__ tile_first, uint32_t* __restrict__ tile_last) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t tiles = (n + kTile - 1) / kTile;
    if (b >= tiles) return;
    uint32_t i0 = b * kTile, i1 = min(n, i0 + kTile) - 1;
    tile_first[b] = (uint32_t)last_le(cell_start, (int)cells, i0);
    tile_last[b] = (uint32_t)last_le(cell_start, (int)cells, i1);
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
    bool fast_path = false;
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
    uint32_t* tile_first = nullptr;
    uint32_t* tile_last = nullptr;
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

Phys make_phys(const FrameMetadata& m) {
    const MiePotentialParams& p = m.particles[0];  // kernel_bucket.cuh:52
    Phys ph{};
    const float two32 = 4294967296.f;
    ph.kx = m.box_width / two32;
    ph.ky = m.box_height / two32;
    ph.ax = ph.kx / p.sigma;
    ph.ay = ph.ky / p.sigma;
    ph.n = p.n;
    ph.m = p.m;
    float hn = p.n / 2.f, hm = p.m / 2.f;
    ph.kn = (int)floorf(hn);
    ph.km = (int)floorf(hm);
    ph.fn = hn - (float)ph.kn;
    ph.fm = hm - (float)ph.km;
    float C = (p.n / (p.n - p.m)) * powf(p.n / p.m, p.m / (p.n - p.m));  // particle.cuh:53-55
    ph.pair_scale = C * p.epsilon / p.sigma;
    ph.wall_scale = C * p.epsilon * p.m;
    ph.sigma = p.sigma;
    ph.mass = (float)6.63352599e-26;  // particle.cuh:51
    ph.dt = m.step_dt;
    ph.box_w = m.box_width;
    ph.box_h = m.box_height;
    ph.cursor_x = m.cursor_pos[0];
    ph.cursor_y = m.cursor_pos[1];
    ph.cursor_r2 = m.cursor_size * m.cursor_size / 4;
    return ph;
}

void apply_metadata(PsimStepper* s, const FrameMetadata& m) {
    s->meta = m;
    s->phys = make_phys(m);
    s->fast_path = s->phys.km == 3 && s->phys.fm == 0.f && s->phys.kn == 7 && s->phys.kn >= 0;
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
    tile_cells_kernel<<<div_up(tiles, 128), 128, 0, s->stream>>>(s->cell_start, s->grid.cells, s->n, s->tile_first,
                                                                 s->tile_last);
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
                                                              s->ty[s->cur_ty ^ 1]);
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
    a.cell_start = s->cell_start;
    a.tile_first = s->tile_first;
    a.tile_last = s->tile_last;
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
    if (s->fast_path) step_kernel<true><<<tiles, kTile, 0, s->stream>>>(a);
    else step_kernel<false><<<tiles, kTile, 0, s->stream>>>(a);
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
                                                                    s->pos[0], s->vel[0], s->ty[0]);
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
    cudaFree(s->tile_first);
    cudaFree(s->tile_last);
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
        CKC(cudaMalloc(&st->pos[k], sizeof(uint2) * cap));
        CKC(cudaMalloc(&st->vel[k], sizeof(float2) * cap));
        CKC(cudaMalloc(&st->ty[k], sizeof(int32_t) * cap));
    }
    CKC(cudaMalloc(&st->cell_start, sizeof(uint32_t) * ((size_t)g.cells + 1)));
    CKC(cudaMalloc(&st->cell_count, sizeof(uint32_t) * (size_t)g.cells));
    CKC(cudaMalloc(&st->block_sum, sizeof(uint32_t) * (size_t)div_up(g.cells, kScanBlock)));
    CKC(cudaMalloc(&st->rank, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&st->perm, sizeof(uint32_t) * cap));
    CKC(cudaMalloc(&st->tile_first, sizeof(uint32_t) * (size_t)div_up((uint32_t)cap, kTile)));
    CKC(cudaMalloc(&st->tile_last, sizeof(uint32_t) * (size_t)div_up((uint32_t)cap, kTile)));
    CKC(cudaMalloc(&st->staging, sizeof(Particle) * cap));
    CKC(cudaMallocHost(&st->h_total, sizeof(uint32_t)));
    CKC(cudaMemset(st->cell_start, 0, sizeof(uint32_t) * ((size_t)g.cells + 1)));
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
