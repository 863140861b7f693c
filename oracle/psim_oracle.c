/*
 * psim_oracle.c -- TEST INFRASTRUCTURE ONLY (see psim_oracle.h).
 *
 * CPU restatement of the reference's hot path.  Every function names the reference lines it
 * follows (paths relative to /root/reference/).  Arithmetic is kept operation for operation in
 * fp32 (compile with -ffp-contract=off) so that it agrees bit-for-bit with the reference's own
 * `__host__` code built by oracle/Makefile.
 */
#include "psim_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define U32_MAX_F ((float)UINT32_MAX) /* rounds to 2^32, as in the reference */

uint64_t oracle_slot_count(OracleGrid g) { return ((uint64_t)g.capacity) << (g.lx + g.ly); }

/* ---- physics primitives: cuda_simulator/src/particle.cuh ---------------------------------- */

typedef struct MieF {
    float sigma, epsilon, n, m, C, mass;
} MieF;

/* particle.cuh:49-55 */
static MieF mie_of(MiePotentialParams p) {
    MieF q;
    q.sigma = p.sigma;
    q.epsilon = p.epsilon;
    q.n = p.n;
    q.m = p.m;
    q.C = (p.n / (p.n - p.m)) * powf(p.n / p.m, p.m / (p.n - p.m));
    q.mass = (float)6.63352599e-26; /* a double literal narrowed to float, particle.cuh:51 */
    return q;
}

float oracle_params_C(MiePotentialParams p) { return mie_of(p).C; }

/* particle.cuh:63-66 */
static float mie_force(const MieF* q, float r) {
    float sr = q->sigma / r;
    return q->C * q->epsilon * (q->m * powf(sr, q->m) - q->n * powf(sr, q->n)) / r;
}

float oracle_f_force(MiePotentialParams p, float r) {
    MieF q = mie_of(p);
    return mie_force(&q, r);
}

/* particle.cuh:68-71 */
static float mie_force_repulsive(const MieF* q, float r) {
    float sr = q->sigma / r;
    return q->C * q->epsilon * q->m * powf(sr, q->m) / r;
}

/* particle.cuh:41-47: separation from a to b; the u32 difference is exact, then one conversion */
static void separation(const Particle* a, const Particle* b, const FrameMetadata* f, float* rx, float* ry) {
    float dx = a->x < b->x ? (float)(b->x - a->x) : -(float)(a->x - b->x);
    float dy = a->y < b->y ? (float)(b->y - a->y) : -(float)(a->y - b->y);
    *rx = (dx / U32_MAX_F) * f->box_width;
    *ry = (dy / U32_MAX_F) * f->box_height;
}

/* particle.cuh:97-103 */
static void pair_force(const MieF* q, float rx, float ry, float* fx, float* fy) {
    float len = hypotf(rx, ry);
    float f = mie_force(q, len);
    f /= len;
    *fx = f * rx;
    *fy = f * ry;
}

/* particle.cuh:125-144 */
static void wall_force(const MieF* q, const Particle* p, const FrameMetadata* f, float* fx, float* fy) {
    if (p->x < UINT32_MAX / 2) {
        float d = ((float)p->x / U32_MAX_F) * f->box_width;
        *fx = mie_force_repulsive(q, d);
    } else {
        float d = ((float)(UINT32_MAX - p->x) / U32_MAX_F) * f->box_width;
        *fx = -mie_force_repulsive(q, d);
    }
    if (p->y < UINT32_MAX / 2) {
        float d = ((float)p->y / U32_MAX_F) * f->box_height;
        *fy = mie_force_repulsive(q, d);
    } else {
        float d = ((float)(UINT32_MAX - p->y) / U32_MAX_F) * f->box_height;
        *fy = -mie_force_repulsive(q, d);
    }
}

/* kernel_bucket.cuh:54-67 (same code in kernel_compact.cuh:10-23) */
static void cursor_force(const Particle* p, const FrameMetadata* f, float* fx, float* fy) {
    float dx = f->cursor_pos[0] - (float)p->x / U32_MAX_F;
    float dy = f->cursor_pos[1] - (float)p->y / U32_MAX_F;
    float sq = dx * dx + dy * dy;
    *fx = 0.f;
    *fy = 0.f;
    if (sq < f->cursor_size * f->cursor_size / 4) {
        *fx = 8e-12f / (sq + 1.f);
        *fy = 8e-12f / (sq + 1.f);
        if (dx > 0) *fx = -*fx;
        if (dy > 0) *fy = -*fy;
    }
}

/* particle.cuh:105-123: leapfrog kick + drift on half-step velocities, wrapping u32 positions */
static void integrate(const MieF* q, Particle* dst, const Particle* src, float fx, float fy, const FrameMetadata* f) {
    float ax = fx / q->mass;
    float ay = fy / q->mass;
    dst->vx = src->vx + ax * f->step_dt;
    dst->vy = src->vy + ay * f->step_dt;
    float dx = dst->vx * f->step_dt;
    float dy = dst->vy * f->step_dt;
    dst->x = src->x + (uint32_t)(int64_t)roundf((dx / f->box_width) * U32_MAX_F);
    dst->y = src->y + (uint32_t)(int64_t)roundf((dy / f->box_height) * U32_MAX_F);
    dst->ty = src->ty;
}

/* ---- binning ------------------------------------------------------------------------------ */

/* kernel.cuh:210-239 */
uint32_t oracle_prepare(const FrameHeader* src, Particle* slots, OracleGrid g) {
    uint64_t cells = 1ull << (g.lx + g.ly);
    uint32_t* len = (uint32_t*)calloc(cells, sizeof(uint32_t));
    uint32_t dropped = 0;
    for (uint32_t i = 0; i < src->particle_count; ++i) {
        Particle p = src->particles[i];
        if (p.ty < 0) continue;
        uint64_t cx = g.lx ? p.x >> (32 - g.lx) : 0;
        uint64_t cy = g.ly ? p.y >> (32 - g.ly) : 0;
        uint64_t cell = cx + (cy << g.lx);
        if (len[cell] >= g.capacity) {
            ++dropped;
            continue;
        }
        slots[cell * g.capacity + len[cell]++] = p;
    }
    for (uint64_t cell = 0; cell < cells; ++cell)
        for (uint32_t k = len[cell]; k < g.capacity; ++k) slots[cell * g.capacity + k].ty = -1;
    free(len);
    return dropped;
}

/* kernel_bucket.cuh:5-39, one cell. The reference strides rows by BUCKETS_Y (:21); the grids it is
 * ever built with are square, and the restatement uses the x extent, which is what a row stride is. */
static uint32_t move_cell(const Particle* src, Particle* dst, OracleGrid g, uint32_t cell) {
    uint32_t bx = 1u << g.lx, by = 1u << g.ly;
    uint32_t cx = cell % bx, cy = cell / bx;
    int x0 = cx == 0 ? 0 : -1, x1 = cx == bx - 1 ? 0 : 1;
    int y0 = cy == 0 ? 0 : -1, y1 = cy == by - 1 ? 0 : 1;
    uint32_t filled = 0;
    Particle* out = dst + (uint64_t)cell * g.capacity;
    for (int dy = y0; dy <= y1; ++dy) {
        for (int dx = x0; dx <= x1; ++dx) {
            const Particle* in = src + ((uint64_t)(cx + dx) + (uint64_t)(cy + dy) * bx) * g.capacity;
            for (uint32_t k = 0; k < g.capacity; ++k) {
                if (in[k].ty < 0) continue;
                uint32_t px = g.lx ? in[k].x >> (32 - g.lx) : 0;
                uint32_t py = g.ly ? in[k].y >> (32 - g.ly) : 0;
                if (px != cx || py != cy) continue;
                out[filled++] = in[k];
                if (filled == g.capacity) return filled;
            }
        }
    }
    for (uint32_t k = filled; k < g.capacity; ++k) out[k].ty = -1;
    return filled;
}

uint64_t oracle_move(const Particle* src, Particle* dst, OracleGrid g) {
    uint64_t cells = 1ull << (g.lx + g.ly), live = 0;
    for (uint64_t c = 0; c < cells; ++c) live += move_cell(src, dst, g, (uint32_t)c);
    return live;
}

/* ---- force + integrate -------------------------------------------------------------------- */

/* kernel_bucket.cuh:40-94 for one slot; if `out_f` is given the force is reported instead of integrated */
/* `species`: 0 = the reference (q[0] for every particle, kernel_bucket.cuh:52); 1 = the per-species EXTENSION
 * (oracle_step_species): q[0], q[1], q[2] are the species pairs 00, 01 (mixed), 11; the wall term uses the particle's own. */
static int species_of(int32_t ty) { return ty > 0 ? 1 : 0; }

static void step_slot(const Particle* src, Particle* dst, const FrameMetadata* f, OracleGrid g, const MieF* q,
                      uint64_t i, float* out_f, int species) {
    if (dst) dst[i].ty = src[i].ty;
    if (src[i].ty < 0) {
        if (out_f) out_f[0] = out_f[1] = out_f[2] = 0.f;
        return;
    }
    uint32_t bx = 1u << g.lx, by = 1u << g.ly;
    float fx, fy, wx, wy;
    const int si = species ? species_of(src[i].ty) : 0;
    cursor_force(&src[i], f, &fx, &fy);
    wall_force(species ? &q[2 * si] : q, &src[i], f, &wx, &wy);
    fx += wx;
    fy += wy;

    uint64_t cell = i / g.capacity;
    uint32_t cx = (uint32_t)(cell % bx), cy = (uint32_t)(cell / bx);
    int x0 = cx == 0 ? 0 : -1, x1 = cx == bx - 1 ? 0 : 1;
    int y0 = cy == 0 ? 0 : -1, y1 = cy == by - 1 ? 0 : 1;
    float max_pair = 0.f;
    for (int dy = y0; dy <= y1; ++dy) {
        for (int dx = x0; dx <= x1; ++dx) {
            uint64_t j0 = ((uint64_t)(cx + dx) + (uint64_t)(cy + dy) * bx) * g.capacity;
            for (uint32_t k = 0; k < g.capacity; ++k) {
                uint64_t j = j0 + k;
                if (j == i || src[j].ty < 0) continue;
                float rx, ry, px, py;
                separation(&src[i], &src[j], f, &rx, &ry);
                pair_force(species ? &q[si + species_of(src[j].ty)] : q, rx, ry, &px, &py);
                fx += px;
                fy += py;
                if (out_f) {
                    float mag = hypotf(px, py);
                    if (mag > max_pair) max_pair = mag;
                }
            }
        }
    }
    if (out_f) {
        out_f[0] = fx;
        out_f[1] = fy;
        out_f[2] = max_pair;
    }
    if (dst) integrate(q, &dst[i], &src[i], fx, fy, f); /* one mass for every species (particle.cuh:51) */
}

typedef struct StepJob {
    const Particle* src;
    Particle* dst;
    const FrameMetadata* f;
    OracleGrid g;
    uint64_t i0, i1;
    int species;
} StepJob;

/* Lorentz-Berthelot mix of the two species (the extension's unlike pair): mean sigma, geometric-mean epsilon, mean exponents */
static MiePotentialParams mie_mix(MiePotentialParams a, MiePotentialParams b) {
    MiePotentialParams m;
    m.sigma = (a.sigma + b.sigma) * 0.5f;
    m.epsilon = sqrtf(a.epsilon * b.epsilon);
    m.n = (a.n + b.n) * 0.5f;
    m.m = (a.m + b.m) * 0.5f;
    return m;
}

static void* step_job(void* arg) {
    StepJob* j = (StepJob*)arg;
    MieF q[3];
    q[0] = mie_of(j->f->particles[0]); /* the reference only ever uses species 0: kernel_bucket.cuh:52 */
    q[1] = mie_of(mie_mix(j->f->particles[0], j->f->particles[1]));
    q[2] = mie_of(j->f->particles[1]);
    for (uint64_t i = j->i0; i < j->i1; ++i) step_slot(j->src, j->dst, j->f, j->g, q, i, NULL, j->species);
    return NULL;
}

static void step_all(const Particle* src, Particle* dst, const FrameMetadata* meta, OracleGrid g, uint32_t threads,
                     int species) {
    uint64_t n = oracle_slot_count(g);
    if (threads <= 1) {
        StepJob j = {src, dst, meta, g, 0, n, species};
        step_job(&j);
        return;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    StepJob* jobs = (StepJob*)malloc(sizeof(StepJob) * threads);
    for (uint32_t t = 0; t < threads; ++t) {
        StepJob j = {src, dst, meta, g, n * t / threads, n * (t + 1) / threads, species};
        jobs[t] = j;
        pthread_create(&th[t], NULL, step_job, &jobs[t]);
    }
    for (uint32_t t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

void oracle_step(const Particle* src, Particle* dst, const FrameMetadata* meta, OracleGrid g, uint32_t threads) {
    step_all(src, dst, meta, g, threads, 0);
}

void oracle_step_species(const Particle* src, Particle* dst, const FrameMetadata* meta, OracleGrid g, uint32_t threads) {
    step_all(src, dst, meta, g, threads, 1);
}

void oracle_forces(const Particle* src, const FrameMetadata* meta, OracleGrid g, float* fx, float* fy,
                   float* max_pair) {
    uint64_t n = oracle_slot_count(g);
    MieF q = mie_of(meta->particles[0]);
    for (uint64_t i = 0; i < n; ++i) {
        float out[3];
        step_slot(src, NULL, meta, g, &q, i, out, 0);
        fx[i] = out[0];
        fy[i] = out[1];
        max_pair[i] = out[2];
    }
}

/* kernel_bucket.cuh:181-206 with the buffer roles of kernel.cuh:10-12: buf0 = src, buf1 = dst,
 * buf2 = D_BUFFER_INTERNAL. */
uint32_t oracle_run_frame(Particle* buf0, Particle* buf1, Particle* buf2, const FrameMetadata* meta, OracleGrid g,
                          uint32_t threads, uint32_t* moves) {
    const int move_every_n = 16;
    int countdown = 0;
    uint32_t steps = 0, nmoves = 0;
    oracle_step(buf0, buf1, meta, g, threads);
    steps += 1;
    while (steps < meta->steps_per_frame) {
        if (countdown <= 0) {
            oracle_move(buf1, buf2, g);
            ++nmoves;
            countdown = move_every_n;
            oracle_step(buf2, buf1, meta, g, threads);
            countdown -= 1;
            steps += 1;
        } else {
            oracle_step(buf1, buf2, meta, g, threads);
            oracle_step(buf2, buf1, meta, g, threads);
            countdown -= 2;
            steps += 2;
        }
    }
    if (moves) *moves = nmoves;
    return steps;
}

/* particle.rs:371-379 */
void oracle_compact(const Particle* slots, const FrameMetadata* meta, OracleGrid g, FrameHeader* dst) {
    uint64_t n = oracle_slot_count(g);
    uint32_t count = 0;
    dst->metadata = *meta;
    for (uint64_t i = 0; i < n; ++i)
        if (slots[i].ty >= 0) dst->particles[count++] = slots[i];
    dst->particle_count = count;
}

/* Energies follow the potential the force derives from, V(r) = C eps ((s/r)^n - (s/r)^m)
 * (particle.cuh:12) and, for the walls, the antiderivative of particle.cuh:68-71:
 * V_wall(d) = C eps (s/d)^m. Double precision; pairs are those of the 3x3 stencil, counted once. */
void oracle_diagnostics(const Particle* slots, const FrameMetadata* meta, OracleGrid g, double out[6]) {
    uint64_t n = oracle_slot_count(g);
    uint32_t bx = 1u << g.lx, by = 1u << g.ly;
    MieF q = mie_of(meta->particles[0]);
    double C = q.C, eps = q.epsilon, sig = q.sigma, en = q.n, em = q.m, mass = q.mass;
    double two32 = 4294967296.0, bw = meta->box_width, bh = meta->box_height;
    double ke = 0, pe = 0, pw = 0, px = 0, py = 0, live = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const Particle* a = &slots[i];
        if (a->ty < 0) continue;
        live += 1;
        ke += 0.5 * mass * ((double)a->vx * a->vx + (double)a->vy * a->vy);
        px += mass * a->vx;
        py += mass * a->vy;
        double dxw = a->x < UINT32_MAX / 2 ? (double)a->x : (double)(UINT32_MAX - a->x);
        double dyw = a->y < UINT32_MAX / 2 ? (double)a->y : (double)(UINT32_MAX - a->y);
        pw += C * eps * (pow(sig / (dxw / two32 * bw), em) + pow(sig / (dyw / two32 * bh), em));
        uint64_t cell = i / g.capacity;
        uint32_t cx = (uint32_t)(cell % bx), cy = (uint32_t)(cell / bx);
        int x0 = cx == 0 ? 0 : -1, x1 = cx == bx - 1 ? 0 : 1;
        int y0 = cy == 0 ? 0 : -1, y1 = cy == by - 1 ? 0 : 1;
        for (int dy = y0; dy <= y1; ++dy)
            for (int dx = x0; dx <= x1; ++dx) {
                uint64_t j0 = ((uint64_t)(cx + dx) + (uint64_t)(cy + dy) * bx) * g.capacity;
                for (uint32_t k = 0; k < g.capacity; ++k) {
                    uint64_t j = j0 + k;
                    if (j <= i || slots[j].ty < 0) continue;
                    const Particle* b = &slots[j];
                    double rx = ((double)b->x - (double)a->x) / two32 * bw;
                    double ry = ((double)b->y - (double)a->y) / two32 * bh;
                    double r = sqrt(rx * rx + ry * ry);
                    pe += C * eps * (pow(sig / r, en) - pow(sig / r, em));
                }
            }
    }
    out[0] = ke;
    out[1] = pe;
    out[2] = pw;
    out[3] = px;
    out[4] = py;
    out[5] = live;
}
