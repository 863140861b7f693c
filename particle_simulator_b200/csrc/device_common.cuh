// device_common.cuh -- kernel parameter blocks, fixed-point / packed fp32x2 helpers and the pair arithmetic all step kernels share.
// Included by stepper.cu inside its anonymous namespace (one translation unit: the kernels, their parameter blocks and
// the host code that launches them are compiled together). Not a stand-alone header.

// ------------------------------------------------------------------------------------------------
// Kernel parameter blocks
// ------------------------------------------------------------------------------------------------

constexpr int kTile = 128;        // particles per CTA of the step kernel (= threads per CTA)
constexpr int kCsCap = 288;       // cell_start entries staged per stencil row (multiple of 4)
constexpr int kPosCap = 640;      // neighbour positions staged per stencil row (multiple of 2)
constexpr int kScanItems = 8;     // cells per thread in the scan kernels
constexpr int kScanThreads = 256;
constexpr int kScanBlock = kScanItems * kScanThreads;
constexpr int kPadCells = 8;      // readable slack after cell_start[cells] for 16-byte bulk copies
constexpr int kPadParticles = 4;  // readable slack after pos[n] for 16-byte bulk copies

struct Grid {
    uint32_t lx;         // log2 cells in x
    uint32_t bx;         // cells in x
    uint32_t by;         // LOCAL cell rows held by this stepper (owned rows + ghost rows)
    uint32_t cells;      // bx * by (local)
    uint32_t sx, sy;     // 32 - log2(cells in x / GLOBAL cells in y): fixed-point coordinate -> global cell
    int32_t row_offset;  // local row = global row - row_offset
    uint32_t own_row0;   // first owned local row (1 when there is a lower ghost row, else 0)
    uint32_t own_rows;   // owned rows
    uint32_t rows_below, rows_above;  // rows the adjacent slabs own (0: no such slab): how far a migrant can be delivered
};

// Where a position falls relative to the rows this stepper owns.
constexpr uint32_t kKeyDown = 0xFFFFFFFEu;  // below the slab: migrates to the lower neighbour
constexpr uint32_t kKeyUp = 0xFFFFFFFDu;    // above the slab: migrates to the upper neighbour

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
    int cursor_on;       // the cursor can reach a particle of the box at all
    float m;
    // A particle further than this from a wall (fixed-point units) skips that wall's term: beyond
    // d = sigma 10^(8/(m+1)) the term is below 1e-8 C eps m / sigma, i.e. < 1e-7 of the largest attraction between two
    // particles and less than one fp32 rounding of any force sum that matters (0xFFFFFFFF: never skip; make_phys).
    uint32_t wall_skip_x, wall_skip_y;
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

// Local cell of a position (kernel.cuh:224-226 with the slab's row offset), or kKeyDown / kKeyUp when
// the position lies outside the owned rows. With a single slab every position is inside.
__device__ __forceinline__ uint32_t cell_of(uint2 p, const Grid& g) {
    int32_t row = (int32_t)(p.y >> g.sy) - g.row_offset;
    if (row < (int32_t)g.own_row0) return kKeyDown;
    if (row >= (int32_t)(g.own_row0 + g.own_rows)) return kKeyUp;
    return (p.x >> g.sx) + ((uint32_t)row << g.lx);
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
// f_dist for any two particles of the box (particle.cuh:41-47): the unsigned difference can exceed 2^31.
__device__ __forceinline__ float wide_diff(uint32_t b, uint32_t a) {
    return a < b ? __uint2float_rn(b - a) : -__uint2float_rn(a - b);
}

// WIDE: the two particles may be further apart than half the box (all-pairs mode); inside a 3x3 stencil of a
// grid of >= 8 cells per axis the wrapping signed difference is the separation.
template <int KN, int FRAC, bool ANISO, bool MASK1, bool CLAMP, bool WIDE = false>
__device__ __forceinline__ void pair2(uint2 pi, uint2 pj0, uint2 pj1, const Phys& ph, float2& gx, float2& gy) {
    float2 x, y;
    if (WIDE) {
        x = make_float2(wide_diff(pj0.x, pi.x), wide_diff(pj1.x, pi.x));
        y = make_float2(wide_diff(pj0.y, pi.y), wide_diff(pj1.y, pi.y));
    } else {
        x = make_float2(__int2float_rn((int)(pj0.x - pi.x)), __int2float_rn((int)(pj1.x - pi.x)));
        y = make_float2(__int2float_rn((int)(pj0.y - pi.y)), __int2float_rn((int)(pj1.y - pi.y)));
    }
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
// CG: the window lies in global memory; loads go to L2 (a ghost row is written by the neighbour slab while
// this kernel runs, and another tile on this SM may have pulled a stale copy of its sector into L1).
template <bool CG>
__device__ __forceinline__ uint2 load_pos(const uint2* p) {
    return CG ? __ldcg(p) : *p;
}

template <int KN, int FRAC, bool ANISO, bool CLAMP, bool CG>
__device__ __forceinline__ void window_accumulate(const uint2* __restrict__ pj, int count, uint2 pi, const Phys& ph,
                                                  float2& gx, float2& gy) {
    int k = 0;
#pragma unroll 2
    for (; k + 1 < count; k += 2)
        pair2<KN, FRAC, ANISO, false, CLAMP>(pi, load_pos<CG>(pj + k), load_pos<CG>(pj + k + 1), ph, gx, gy);
    if (k < count) pair2<KN, FRAC, ANISO, true, CLAMP>(pi, load_pos<CG>(pj + k), pi, ph, gx, gy);
}

// Repulsive wall term C eps m (sigma/d)^m / d (particle.cuh:68-71).
// M6: m == 6 is known at compile time (every kernel variant with compile-time powers): no predicated-off MUFU pair.
template <bool M6>
__device__ __forceinline__ float wall_term(float d, const Phys& ph) {
    float inv_d = fast_rcp(d);
    float q = ph.sigma * inv_d;
    float pw;
    if (M6 || ph.wall_m6) {
        float q2 = q * q;
        pw = q2 * q2 * q2;
    } else {
        pw = fast_ex2(ph.m * fast_lg2(q));
    }
    return ph.wall_scale * pw * inv_d;
}

// Cursor + wall forces on one particle (kernel_bucket.cuh:54-69, particle.cuh:125-144).
template <bool M6>
__device__ __forceinline__ float2 field_force(uint2 p, const Phys& ph) {
    const float inv32 = 1.f / 4294967296.f;
    float2 f = make_float2(0.f, 0.f);
    if (ph.cursor_on) {  // uniform: the editor parks the cursor at (-1, -1), out of reach of every particle
        float dx = ph.cursor_x - __uint2float_rn(p.x) * inv32;
        float dy = ph.cursor_y - __uint2float_rn(p.y) * inv32;
        float sq = dx * dx + dy * dy;
        if (sq < ph.cursor_r2) {
            float c = 8e-12f * fast_rcp(sq + 1.f);
            f.x = dx > 0 ? -c : c;
            f.y = dy > 0 ? -c : c;
        }
    }
    // nearest wall per axis (particle.cuh:125-144: side chosen by p.x < UINT32_MAX / 2), one term per axis
    const bool left = p.x < 0xFFFFFFFFu / 2, low = p.y < 0xFFFFFFFFu / 2;
    const uint32_t dxw = left ? p.x : 0xFFFFFFFFu - p.x, dyw = low ? p.y : 0xFFFFFFFFu - p.y;
    if (dxw <= ph.wall_skip_x) {
        const float w = wall_term<M6>(__uint2float_rn(dxw) * ph.kx, ph);
        f.x += left ? w : -w;
    }
    if (dyw <= ph.wall_skip_y) {
        const float w = wall_term<M6>(__uint2float_rn(dyw) * ph.ky, ph);
        f.y += low ? w : -w;
    }
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

