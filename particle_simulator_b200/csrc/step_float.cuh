// step_float.cuh -- the step kernel fine grids (>= 1024 cells per axis) run.  Included by stepper.cu inside its
// anonymous namespace.
//
// Three ideas on top of step_kernel (stepper.cu):
//
// 1. EXACT fp32 OFFSETS.  f_dist (particle.cuh:41-47) subtracts the u32 coordinates exactly and rounds the
//    difference to fp32 once.  An integer of magnitude < 2^24 is exact in fp32 and so is the difference of two
//    such numbers, so  float(xj - O) - float(xi - O) == float(xj - xi)  whenever both offsets from a common
//    origin O stay below 2^24 fixed-point units: neighbours are staged in shared memory as fp32 offsets and the
//    pair loop needs no integer subtract and no int -> float conversion.  A cell is 2^(32-LX) units wide, so an
//    origin serves only a few cells:
//      y: a tile lies inside one cell row; its three stencil rows share the centre of that row (TileC::yc);
//      x: cell columns are grouped into ZONES of 2^zl columns (zl = 2 for LX >= 11, 1 for LX = 10).  A thread
//         whose cell is column cx works in zone z = max(cx-1, 0) >> zl; its three stencil columns lie within the
//         first 2^zl + 2 columns of that zone.  Every column belongs to exactly one even and one odd zone (z and
//         z-1 of its own zone z), so each staged neighbour carries two x offsets, from the centre of the used
//         span of its even zone and of its odd zone, and a thread reads the one with the parity of its own zone.
//    Offsets are pre-multiplied by a power of two (exact) that brings (r/sigma)^2 to within [1/4, 1) of its true
//    value; the remaining factor f in [1, 2) of kx/sigma is folded into the constants of the force law (PhysF).
//
// 2. TWO CELL-MATES PER THREAD.  A thread steps a COUPLE: two consecutive particles of one cell (a cell with an
//    odd count ends in a half-empty couple).  Both see the same stencil, so one neighbour is loaded once and the
//    packed fp32x2 pipe (FADD2 / FMUL2 / FFMA2) evaluates it against both: per neighbour 1 LDS.64, 16 packed
//    instructions and 4 MUFU for two pair forces, and all per-thread set-up is shared by two particles.
//
// 3. ONE CONTIGUOUS STENCIL.  The tile's band (3 rows x its columns +- 1) is laid out in shared memory
//    COLUMN-major: for each column the cells of the row below, the own row and the row above.  The 3x3 stencil of
//    a thread is then one contiguous range, walked by one loop (split in three only to confine the r = 0 guard
//    to the own cell).  The accumulation order differs from the reference's (dy, dx, slot) order; results stay
//    within the 1e-5 tolerance and are independent of the slab decomposition.
//
// Tiles are row-aligned: a tile is up to 128 consecutive couples of ONE cell row (TileC, built at re-bin time).

constexpr int kColCap = 128;   // cell columns a tile's band may span (its own columns +- 1)
constexpr int kCsRow = 136;    // cell_start entries staged per row: kColCap + 1, + 3 alignment slack, multiple of 4
constexpr int kColStride = 132; // words per tile in the column-offset table (kColCap + 1, multiple of 4)
constexpr int kBandCap = 1280;  // neighbours of the whole band (a crystal at r0 alternates rows of 4 and 6 per cell)
constexpr int kCouples = 128;  // couples (= threads) per tile

struct PhysF {
    float sx, sy;          // fixed-point units -> scaled units (sx a power of two; sy = sx * ky/kx, also one)
    float zone_shift;      // zone stride * cell width * sx: distance between the even-zone and odd-zone origins
    float d0, d1, d2, d3;  // -(n/m) f^(-2(kn-km)) q^fn as a cubic in l = log2(scaled r^2)  (d0 alone if fn == 0)
    float pair_scale;      // scaled pair sum -> newtons
    uint32_t zl;           // log2 of the zone stride in cell columns
    uint32_t half_span;    // (2^zl + 2) cells / 2 in fixed-point units: centre of a zone's used span
    uint32_t sxbits;       // 32 - LX
};

// One tile: couples [k0, k0 + nk) of local cell row `row`, whose cells are columns [c_first, c_last] of it.
// The band it stages: columns [col_lo, col_lo + ncol) of rows row-1, row, row+1 (rows outside the grid: cnt 0).
struct __align__(16) TileC {
    uint32_t k0, nk;
    uint32_t row;
    uint32_t col_lo, ncol;
    uint32_t yc;          // fixed-point y of the centre of the row (global coordinates)
    uint32_t fits;        // 0: the band exceeds the staging buffers -> global-memory path
    uint32_t _pad0;
    uint32_t cs_lo[3];    // first cell_start entry staged per row (multiple of 4)
    uint32_t cs_cnt[3];   // entries staged per row (multiple of 4; 0: row outside the grid)
    uint32_t _pad1[2];
};
static_assert(sizeof(TileC) == 64, "TileC is read as four 16-byte words");

struct StepArgsC {
    const uint32_t* __restrict__ couple_i0;  // per couple: index of its first particle | (has a second one) << 31
    const TileC* __restrict__ tiles;
    const uint32_t* __restrict__ col_start;  // per tile, kColStride words: first band slot of each of its columns
    PhysF pf;
};

// One staged neighbour (x_even, y, x_odd, y) against the thread's two particles.  nx, ny: minus their offsets.
// g f^(2 km) = qs^4 - (n/m) f^(-2(kn-4)) qs^KN q^fn  with qs = 1 / (scaled r^2); see make_phys_f().
template <int KN, int FRAC, bool CLAMP>
__device__ __forceinline__ void pairc(float xj, float yj, float2 nx, float2 ny, const PhysF& pf, float2& gx, float2& gy) {
    float2 x = __fadd2_rn(nx, splat(xj));
    float2 y = __fadd2_rn(ny, splat(yj));
    float2 r2 = __ffma2_rn(y, y, __fmul2_rn(x, x));
    if (CLAMP) {  // the own cell contains the particle itself: an exact 0 instead of 0 * inf
        r2.x = fmaxf(r2.x, 1e-3f);
        r2.y = fmaxf(r2.y, 1e-3f);
    }
    float2 q = make_float2(fast_rcp(r2.x), fast_rcp(r2.y));
    float2 q2 = __fmul2_rn(q, q);
    float2 q4 = __fmul2_rn(q2, q2);
    float2 pn;
    if (KN == 5) pn = __fmul2_rn(q4, q);
    else if (KN == 6) pn = __fmul2_rn(q4, q2);
    else if (KN == 7) pn = __fmul2_rn(__fmul2_rn(q4, q2), q);
    else if (KN == 8) pn = __fmul2_rn(q4, q4);
    else if (KN == 9) pn = __fmul2_rn(__fmul2_rn(q4, q4), q);
    else pn = __fmul2_rn(__fmul2_rn(q4, q4), q2);
    float2 e;
    if (FRAC == kFracPoly) {
        float2 l = make_float2(fast_lg2(r2.x), fast_lg2(r2.y));
        e = __ffma2_rn(l, splat(pf.d3), splat(pf.d2));
        e = __ffma2_rn(l, e, splat(pf.d1));
        e = __ffma2_rn(l, e, splat(pf.d0));
    } else {
        e = splat(pf.d0);
    }
    float2 g = __ffma2_rn(pn, e, q4);
    gx = __ffma2_rn(g, x, gx);
    gy = __ffma2_rn(g, y, gy);
}

template <int KN, int FRAC>
__global__ void __launch_bounds__(kCouples, 9) step_kernel_c(const StepArgs a, const StepArgsC ac) {
    __shared__ __align__(16) float4 s_nb[kBandCap];          // the band, column-major: (x_even, y, x_odd, y)
    __shared__ __align__(16) uint32_t s_cs[3][kCsRow];       // cell_start slices of the three rows
    __shared__ __align__(16) uint32_t s_col[kColStride];     // first band slot of every staged column
    __shared__ __align__(8) uint64_t s_bar;

    const Grid& g = a.g;
    const PhysF& pf = ac.pf;
    const uint32_t tile = halo_tile_order(a, blockIdx.x, gridDim.x);
    const TileC t = ac.tiles[tile];
    const bool live = threadIdx.x < t.nk;

    if (a.push) {  // uniform over the grid
        if (threadIdx.x == 0) {
            if (blockIdx.x == 0) halo_publish_empty(a);
            halo_wait(a, tile);
        }
        if (!t.fits) __syncthreads();  // the global-memory path reads the ghost rows directly
    }
    if (!t.fits) {  // very sparse or very crowded spot: same physics straight from global memory, one particle at a time
        if (!live) return;
        const uint32_t w = ac.couple_i0[t.k0 + threadIdx.x];
        const uint32_t i0 = w & 0x7FFFFFFFu;
        const uint32_t* cs[3] = {a.cell_start, a.cell_start, a.cell_start};
        const uint2* pp[3] = {a.pos_in, a.pos_in, a.pos_in};
        const uint32_t zero[3] = {0, 0, 0};
        for (uint32_t i = i0; i <= i0 + (w >> 31); ++i)
            step_particle<KN, FRAC, true, true>(i, a.pos_in[i], a.vel[i], a.cell_id[i], cs, zero, pp, zero, a);
        return;
    }

    // TMA: the cell_start slices of the three rows and the tile's column offsets (built at re-bin time)
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        const uint32_t col_bytes = ((t.ncol + 1u + 3u) & ~3u) * 4u;
        mbar_arrive_expect_tx(&s_bar, (t.cs_cnt[0] + t.cs_cnt[1] + t.cs_cnt[2]) * 4u + col_bytes);
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (t.cs_cnt[d]) bulk_copy_g2s(s_cs[d], a.cell_start + t.cs_lo[d], t.cs_cnt[d] * 4u, &s_bar);
        bulk_copy_g2s(s_col, ac.col_start + (size_t)tile * kColStride, col_bytes, &s_bar);
    }

    // this thread's couple (the loads fly while the slices arrive)
    uint32_t i0 = a.own_lo, has1 = 0;
    if (live) {
        const uint32_t w = ac.couple_i0[t.k0 + threadIdx.x];
        i0 = w & 0x7FFFFFFFu;
        has1 = w >> 31;
    }
    const uint32_t i1 = i0 + has1;  // a half-empty couple computes its only particle twice
    const uint2 p0 = a.pos_in[i0], p1 = a.pos_in[i1];
    const float2 v0 = a.vel[i0], v1 = a.vel[i1];
    const uint32_t cx = a.cell_id[i0] & (g.bx - 1);

    // entry of column c (band-relative) of row d in s_cs[d]
    const uint32_t row0 = t.row - 1;  // wraps for row 0: that row has cs_cnt == 0 and is never read
    uint32_t off[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) off[d] = ((row0 + d) << g.lx) + t.col_lo - t.cs_lo[d];

    __syncthreads();  // the barrier is initialised before anyone polls it
    mbar_wait(&s_bar, 0);

    // fixed-point positions -> exact scaled fp32 offsets, one thread per cell of the band. Threads run along a row
    // (consecutive cells are consecutive in HBM: coalesced loads); the band itself is column-major.
    for (uint32_t idx = threadIdx.x; idx < 3u * t.ncol; idx += kCouples) {
        const uint32_t d = idx < t.ncol ? 0u : (idx < 2u * t.ncol ? 1u : 2u);
        const uint32_t c = idx - d * t.ncol;
        uint32_t dst = s_col[c], e;
        if (d == 0) {
            if (t.cs_cnt[0] == 0) continue;
            e = off[0] + c;
        } else {
            if (t.cs_cnt[0]) dst += s_cs[0][off[0] + c + 1] - s_cs[0][off[0] + c];
            if (d == 1) {
                e = off[1] + c;
            } else {
                if (t.cs_cnt[2] == 0) continue;
                dst += s_cs[1][off[1] + c + 1] - s_cs[1][off[1] + c];
                e = off[2] + c;
            }
        }
        const uint32_t r0 = s_cs[d][e], r1 = s_cs[d][e + 1];
        const uint32_t zq = (t.col_lo + c) >> pf.zl;
        const uint32_t xo0 = (((zq & ~1u) << pf.zl) << pf.sxbits) + pf.half_span;  // centre of the even zone's span
        const float to_odd = (zq & 1u) ? -pf.zone_shift : pf.zone_shift;
        const uint2* __restrict__ src = a.pos_in + r0;
        float4* out = s_nb + dst;
        const uint32_t n = r1 - r0;
        for (uint32_t base = 0; base < n; base += 4) {  // four loads in flight: a cell holds 4-6 particles
            uint2 p[4];
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u)
                if (base + u < n) p[u] = __ldcg(src + base + u);  // L2: a ghost row is written by the neighbour meanwhile
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                if (base + u < n) {
                    const float xe = __int2float_rn((int)(p[u].x - xo0)) * pf.sx;
                    const float y = __int2float_rn((int)(p[u].y - t.yc)) * pf.sy;
                    out[base + u] = make_float4(xe, y, xe + to_odd, y);
                }
            }
        }
    }
    __syncthreads();
    if (!live) return;

    // the stencil of this thread: columns [xa, xb] of the band = slots [w0, w1); its own cell = [o0, o1)
    const uint32_t cb = cx - t.col_lo;
    const uint32_t xa = cx == 0 ? cb : cb - 1, xb = cx == g.bx - 1 ? cb : cb + 1;
    const uint32_t w0 = s_col[xa], w1 = s_col[xb + 1];
    uint32_t o0 = s_col[cb];
    if (t.cs_cnt[0]) o0 += s_cs[0][off[0] + cb + 1] - s_cs[0][off[0] + cb];
    const uint32_t o1 = o0 + s_cs[1][off[1] + cb + 1] - s_cs[1][off[1] + cb];

    const uint32_t zt = (cx == 0 ? 0u : cx - 1) >> pf.zl;
    const uint32_t xo = ((zt << pf.zl) << pf.sxbits) + pf.half_span;
    const float2 nx = make_float2(-(__int2float_rn((int)(p0.x - xo)) * pf.sx), -(__int2float_rn((int)(p1.x - xo)) * pf.sx));
    const float2 ny = make_float2(-(__int2float_rn((int)(p0.y - t.yc)) * pf.sy), -(__int2float_rn((int)(p1.y - t.yc)) * pf.sy));
    // (x_even, y) or (x_odd, y): the 8 bytes at offset 0 or 8 of a slot
    const float2* nb = reinterpret_cast<const float2*>(s_nb) + (zt & 1u);
    float2 gx = splat(0.f), gy = splat(0.f);
    for (uint32_t k = w0; k < o0; ++k) {
        const float2 j = nb[2 * k];
        pairc<KN, FRAC, false>(j.x, j.y, nx, ny, pf, gx, gy);
    }
    for (uint32_t k = o0; k < o1; ++k) {
        const float2 j = nb[2 * k];
        pairc<KN, FRAC, true>(j.x, j.y, nx, ny, pf, gx, gy);
    }
    for (uint32_t k = o1; k < w1; ++k) {
        const float2 j = nb[2 * k];
        pairc<KN, FRAC, false>(j.x, j.y, nx, ny, pf, gx, gy);
    }
    finish_particle(i0, p0, v0, gx.x, gy.x, pf.pair_scale, pf.pair_scale, a);
    if (has1) finish_particle(i1, p1, v1, gx.y, gy.y, pf.pair_scale, pf.pair_scale, a);
}

// ---- re-bin side of the couples ------------------------------------------------------------------------

// couple_i0[k] for the couples of every cell; couples are numbered by pad_start / 2.
__global__ void couple_build_kernel(const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ pad_start,
                                    uint32_t cells, uint32_t* __restrict__ couple_i0) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const uint32_t s = cell_start[c], n = cell_start[c + 1] - s;
    uint32_t k = pad_start[c] >> 1;
    for (uint32_t m = 0; m < n; m += 2, ++k) couple_i0[k] = (s + m) | (m + 1 < n ? 0x80000000u : 0u);
}

// Tiles of every owned row: ceil(couples of the row / kCouples); single block: exclusive scan into
// tile_base[0..own_rows], total -> *total_out.
__global__ void __launch_bounds__(1024) row_tiles_kernel(const uint32_t* __restrict__ pad_start, Grid g,
                                                         uint32_t* __restrict__ tile_base, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < g.own_rows; base += 1024) {
        const uint32_t r = base + threadIdx.x;
        uint32_t v = 0;
        if (r < g.own_rows) {
            const uint32_t row = g.own_row0 + r;
            const uint32_t m = (pad_start[(row + 1) << g.lx] - pad_start[row << g.lx]) >> 1;
            v = (m + kCouples - 1) / kCouples;
        }
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += u;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sum[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if (threadIdx.x >= o) wi += u;
            }
            warp_sum[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_sum[threadIdx.x >> 5] + incl - v;
        if (r < g.own_rows) tile_base[r] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_base[g.own_rows] = carry;
        *total_out = carry;
    }
}

// One TileC per tile index b in [0, tile_base[own_rows]), and its row of the column-offset table. One warp per tile:
// the lanes share the scan of the band's column counts.
__global__ void tile_build_kernel(const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ pad_start,
                                  const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ couple_i0,
                                  const uint32_t* __restrict__ cell_id, Grid g, TileC* __restrict__ tiles,
                                  uint32_t* __restrict__ col_start) {
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (b >= tile_base[g.own_rows]) return;  // whole warps leave together
    // the last row whose base is <= b (rows without tiles share their successor's base and are skipped over)
    const uint32_t r = (uint32_t)last_le(tile_base, (int)g.own_rows, b);
    const uint32_t row = g.own_row0 + r;
    const uint32_t row_k0 = pad_start[row << g.lx] >> 1, row_k1 = pad_start[(row + 1) << g.lx] >> 1;
    TileC t;
    t.k0 = row_k0 + (b - tile_base[r]) * kCouples;
    t.nk = min(row_k1 - t.k0, (uint32_t)kCouples);
    t.row = row;
    const uint32_t c_first = cell_id[couple_i0[t.k0] & 0x7FFFFFFFu] & (g.bx - 1);
    const uint32_t c_last = cell_id[couple_i0[t.k0 + t.nk - 1] & 0x7FFFFFFFu] & (g.bx - 1);
    t.col_lo = c_first == 0 ? 0 : c_first - 1;
    const uint32_t col_hi = c_last == g.bx - 1 ? c_last : c_last + 1;
    t.ncol = col_hi - t.col_lo + 1;
    t.yc = (uint32_t)((2ll * ((long long)row + g.row_offset) + 1) << (g.sy - 1));
    t._pad0 = t._pad1[0] = t._pad1[1] = 0;
    bool fits = t.ncol <= (uint32_t)kColCap;
    uint32_t lo[3];
    bool has[3];
    for (int d = 0; d < 3; ++d) {
        const long long rd = (long long)row + d - 1;
        has[d] = rd >= 0 && rd < (long long)g.by;
        lo[d] = has[d] ? ((uint32_t)rd << g.lx) + t.col_lo : 0u;
        t.cs_lo[d] = lo[d] & ~3u;
        t.cs_cnt[d] = has[d] ? ((lo[d] + t.ncol + 1 - t.cs_lo[d]) + 3u) & ~3u : 0u;  // entries lo .. lo + ncol
        fits = fits && t.cs_cnt[d] <= (uint32_t)kCsRow;
    }
    uint32_t band = 0;
    if (fits) {  // exclusive scan of the three rows' counts over the band's columns, 32 columns at a time
        uint32_t* col = col_start + (size_t)b * kColStride;
        const uint32_t padded = (t.ncol + 1u + 3u) & ~3u;
        for (uint32_t c0 = 0; c0 < padded; c0 += 32) {
            const uint32_t c = c0 + lane;
            uint32_t n = 0;
            if (c < t.ncol)
                for (int d = 0; d < 3; ++d)
                    if (has[d]) n += cell_start[lo[d] + c + 1] - cell_start[lo[d] + c];
            uint32_t incl = n;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= (uint32_t)o) incl += u;
            }
            if (c < padded) col[c] = band + incl - n;  // entries past ncol repeat the total
            band += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
    }
    t.fits = fits && band <= (uint32_t)kBandCap ? 1u : 0u;
    if (lane == 0) tiles[b] = t;
}
