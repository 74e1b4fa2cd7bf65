// gravity.cu -- Barnes-Hut walk of the linear octree with the reference's opening rule and
// spline-softened leaf interactions.
//
// Replaces GJL.compute_g / gravity_acc / Kernels / min_distance2_point_to_cell
// (F/gravOctree_Single.jl:5-29, :231-304).  The reference walks the tree once per particle with a FIFO
// queue; here one warp walks it for 32 key-adjacent targets at once with a shared depth-first stack of
// cells TO BE OPENED, each with the mask of lanes that have not accepted one of its ancestors.  Opening a
// cell visits its (BFS-contiguous) children: every lane of the mask evaluates the reference's two-clause
// acceptance test for ITS OWN particle
//        s^2/d^2 < theta^2  (d to the node COM)   and   h_i^2 / mindist^2(p_i, cell) < 0.25        (:265)
// lanes that accept add the monopole, the others form the mask pushed with the child; leaf children are
// evaluated on the spot.  Decisions are exactly per-particle as in the reference - d^2 and mindist^2 are
// formed with the reference's roundings, the two quotient comparisons are decided by products and fall
// back to the IEEE division inside a 1e-15 band around the threshold - only the floating-point
// summation order differs (the reference's own order already drifts through its leaf list surgery,
// :293-300), and 1/d, 1/d^3 come from one rsqrt (<= 4 ulp; the parity bar for gravity is 1e-6).
// Node data is read with warp-uniform addresses (32 B per double4, broadcast to the warp).
// walk_pairs_kernel (default) adds a pair queue for the cells only a few lanes still have to open (see its header);
// walk_kernel is the shared walk alone (also used when node ids do not fit the pair encoding).
#include "sph_internal.cuh"

namespace {

constexpr int GW_WARPS = 4;
constexpr int GW_STACK = 192;          // shared stack of the walks: 7 pending siblings per level + 8 covers 26 levels
constexpr int GW_STACK_DEEP = 304;     // ... and all 42 levels (7 * 42 + 8): the DEEP variants, see sph_launch_walk
constexpr int GW_REC = SPH_WALK_REC;   // double4 per walk record: {rCOM, Mass | h_j}, {(2L)^2, radius, child info, range}

// coefficients of the softened kernels below, read as constant-bank operands (as literals each one costs two
// uniform-register moves in front of the FP64 instruction that uses it)
__constant__ double c_gk[13] = {4.0 / 3, 6.0 / 5, 1.0 / 2, 2.0 / 3, 3.0 / 10, 1.0 / 10, 7.0 / 5,
                                8.0 / 3, 1.0 / 6, 1.0 / 15, 1.0 / 30, 8.0 / 5, 3.0};

// Kernels (F/gravOctree_Single.jl:5-29): grad(PHI)/r and PHI of the spline-softened potential, written in
// q = r/h and 1/h (same polynomials; one reciprocal and one rsqrt instead of seven divisions)
#ifdef WALK_NOINLINE_NEAR
#define NEAR_INLINE __noinline__
#else
#define NEAR_INLINE __forceinline__
#endif
__device__ NEAR_INLINE void grav_pair(double d_sq, double h, double &gPHI, double &PHI) {
    const double rinv = d_sq > 0.0 ? fast_rsqrt(d_sq) : 0.0;
    const double r = d_sq * rinv;
    const double hinv = fast_rcp(h);
    const double q = r * hinv;
    if (q > 2.0) {
        gPHI = rinv * rinv * rinv;   // 1/r^3  (:19)
        PHI = -rinv;                 // -1/r   (:20)
        return;
    }
    const double q2 = q * q, q3 = q2 * q;
    const double hinv3 = hinv * hinv * hinv;
    if (q <= 1.0) {
        gPHI = hinv3 * ((c_gk[0] - c_gk[1] * q2) + c_gk[2] * q3);                                    // (:10)
        PHI = hinv * (((c_gk[3] * q2 - c_gk[4] * (q2 * q2)) + c_gk[5] * (q2 * q3)) - c_gk[6]);       // (:11)
    } else {
        const double qi = rinv * h;  // 1/q
        gPHI = hinv3 * ((((c_gk[7] - c_gk[12] * q) + c_gk[1] * q2) - c_gk[8] * q3) - c_gk[9] * (qi * qi * qi));              // (:14)
        PHI = hinv * (((((c_gk[0] * q2 - q3) + c_gk[4] * (q2 * q2)) - c_gk[10] * (q2 * q3)) - c_gk[11]) + c_gk[9] * qi);   // (:15)
    }
}

// fl(a / b) < c  decided without the division except within a relative band of 1e-15 around equality
// (a, b, c >= 0; b == 0 gives +inf or NaN, i.e. false, exactly like the quotient)
__device__ __forceinline__ bool quotient_less(double a, double b, double c) {
    const double t = c * b;
    if (a < t * (1.0 - 1e-15)) return true;
    if (a > t * (1.0 + 1e-15)) return false;
    asm volatile("");   // keeps the IEEE division inside the band: the compiler otherwise if-converts it and every
                        // lane that fails the first comparison pays for a division
    return a / b < c;
}

// max(max(lo - p, 0), p - hi)  (min_distance2_point_to_cell, :231-236) with compare/select instead of the
// NaN-aware fmax sequence; lo <= hi, so at most one of the two differences is positive
__device__ __forceinline__ double axis_dist(double lo, double hi, double p) {
    const double a = lo - p, b = p - hi;
    const double mx = a > b ? a : b;
    return mx > 0.0 ? mx : 0.0;
}
// the same value from the sign bits (integer pipe instead of two FP64 compares): a > 0 excludes b > 0, and a zero
// of either sign yields 0 like the maximum does
__device__ __forceinline__ double axis_dist_bits(double lo, double hi, double p) {
    const double a = lo - p, b = p - hi;
    const double bz = __double2hiint(b) >= 0 ? b : 0.0;
    return __double2hiint(a) >= 0 ? a : bz;
}

// build-time experiments (tools/build_variants.sh): resident blocks of the pair walk, rare paths out of line
#ifndef WALK_MINB
#define WALK_MINB 8
#endif
#ifdef WALK_NOINLINE_C2
#define C2_INLINE __noinline__
#else
#define C2_INLINE __forceinline__
#endif

// clause 2 of the acceptance rule evaluated with the reference's expression: h_i^2 / mindist^2(p_i, cell) < 0.25
__device__ C2_INLINE bool clause2_exact(const double4 *__restrict__ nodeBC, int n, double px, double py, double pz, double hi2) {
    const double4 B = nodeBC[2 * (int64_t)n];
    const double4 C = nodeBC[2 * (int64_t)n + 1];
    const double ex = axis_dist_bits(B.x, B.w, px), ey = axis_dist_bits(B.y, C.x, py), ez = axis_dist_bits(B.z, C.y, pz);
    return quotient_less(hi2, sph_d2_exact(ex, ey, ez), 0.25);
}

// the reference's acceptance rule (:265) for one particle and one internal cell (both walks use it); bc[2 n], bc[2 n + 1]
// is the cell's {box, ..} record
__device__ __forceinline__ bool cell_accepted(const double4 *__restrict__ bc, int n, double s_sq, double radius, double d_sq,
                                              double px, double py, double pz, double hi2, double h2x,
                                              double theta_sq, double th_lo, double th_hi) {
    // clause 1: s*s/d_sq < theta_sq
    bool accept;
    if (s_sq < d_sq * th_lo) accept = true;
    else if (s_sq > d_sq * th_hi) accept = false;
    else {
        asm volatile("");   // see quotient_less
        accept = s_sq / d_sq < theta_sq;
    }
    // clause 2: h_i*h_i / mind2 < 0.25, proven from d > radius + 2 h_i, else the reference's expression
    if (accept) {
        const double w = radius + h2x;
        if (!(d_sq > w * w)) accept = clause2_exact(bc, n, px, py, pz, hi2);
    }
    return accept;
}

// leaf = one particle j with smoothing length hj (:259-264): grad(PHI)/r and PHI per unit mass
__device__ __forceinline__ void leaf_pair(double d_sq, double hi, double hj, double &gP, double &pot) {
    const double hs = hi + hj;                           // 2 h_ij (:259)
    if (d_sq > hs * hs) {                                // q > 2: Newtonian (:19-20); (2 h_ij)^2 is 4 h_ij^2 bit for bit
        const double rinv = fast_rsqrt(d_sq);
        gP = rinv * rinv * rinv;
        pot = -rinv;
    } else {
        grav_pair(d_sq, hs / 2, gP, pot);
    }
}

// Root children handled by block row y of a tile.  The children are ordered o = 0, 1, ..: o = 0 is the child that
// CONTAINS the tile (its walk is by far the longest: the whole near field), o >= 1 the others in cyclic order; row y of
// `rows` takes o = y, y + rows, ...  The hardware starts blocks in grid order, so every long work item begins before
// any short one and the short ones fill the tail.
__device__ __forceinline__ int walk_root_near(const double4 *__restrict__ BC, int2 R, int64_t first_slot) {
    const int nch = R.y & 0xff;
    int near = 0;
    for (int c = 0; c < nch; ++c) {
        const int2 rg = unpack_i2(BC[2 * (int64_t)(R.x + c) + 1].w);        // {nstart, ncount}
        if (first_slot >= rg.x && first_slot < (int64_t)rg.x + rg.y) near = c;
    }
    return near;
}

// true: this launch has nothing to do.  The regular variants run unless an error is flagged; the DEEP variants run only
// when the regular walk of this evaluation overflowed its stack (and nothing else went wrong) and redo every tile.
template <bool DEEP>
__device__ __forceinline__ bool walk_skipped(const unsigned long long *__restrict__ scal) {
    const unsigned long long f = scal[SC_ERR];
    return DEEP ? f != (unsigned long long)ERRF_STACK : f != 0ull;
}

// one work item (tile bx of this rank, block row `row` of `rows`) of the shared walk
template <bool COUNT, bool DEEP, int STACK>
__device__ __forceinline__ void walk_tile(int bx, int row, int rows, int4 *__restrict__ stack, int64_t N, int nranks, int rank,
                                          int64_t chunk, const double4 *__restrict__ pos4, const double *__restrict__ hs,
                                          const SphTree &t, double theta_sq, double m, unsigned long long *__restrict__ scal,
                                          double *__restrict__ part /* [rows][4][chunk] */) {
    const int lane = threadIdx.x & 31;
    const double4 *__restrict__ W = t.nodeW;
    // tiles of 128 key-adjacent targets are dealt round-robin in groups of SPH_WALK_DEAL consecutive tiles
    const int64_t local = (int64_t)bx * (GW_WARPS * 32) + threadIdx.x;
    const int64_t gtile = ((int64_t)(bx / SPH_WALK_DEAL) * nranks + rank) * SPH_WALK_DEAL + bx % SPH_WALK_DEAL;
    const int64_t s = gtile * (GW_WARPS * 32) + threadIdx.x;
    const bool active = s < N;
    double px = 0, py = 0, pz = 0, hi = 1.0;
    if (active) {
        const double4 p = pos4[s];
        px = p.x; py = p.y; pz = p.z; hi = hs[s];
    }
    const double hi2 = hi * hi;
    const double h2x = 2.0 * hi * (1.0 + 1e-9);      // clause 2 is certainly true when mindist > h2x
    const double th_lo = theta_sq * (1.0 - 1e-15), th_hi = theta_sq * (1.0 + 1e-15);
    double gx = 0.0, gy = 0.0, gz = 0.0, ph = 0.0;
    unsigned long long visits = 0;
    const unsigned amask = __ballot_sync(0xffffffffu, active);
    int sp = 0;
    // the walk starts by opening the root: the root itself is never tested (:246-249).  The root's children are
    // dealt to the gridDim.y block rows: more, shorter work items (wave quantisation matters when a rank owns ~1 wave
    // of tiles); the partial sums are added in row order by walk_reduce_kernel, so results stay deterministic.
    const int2 R = unpack_i2(W[1].z);
    const int rnch = R.y & 0xff;
    if (row >= rnch) return;
    if (amask) {
        const int near = walk_root_near(t.nodeBC, R, gtile * (GW_WARPS * 32));
        if (lane == 0) {
            int o = row;
            while (o + rows < rnch) o += rows;
            for (; o >= 0; o -= rows) {               // pushed far to near: the near child is walked first
                const int rc = near + o >= rnch ? near + o - rnch : near + o;
                stack[sp++] = make_int4(R.x + rc, 1 | ((((R.y >> 8) >> rc) & 1) << 8), (int)amask, 0);
            }
        }
        sp = __shfl_sync(0xffffffffu, sp, 0);
    }
    __syncwarp();
    while (sp > 0) {
        const int4 top = stack[--sp];
        __syncwarp();
        const bool mine = (((unsigned)top.z) >> lane) & 1u;
        const int first = top.x, nch = top.y & 0xff, leafmask = top.y >> 8;
        if (sp + nch > STACK) {     // never write out of bounds: the DEEP variant redoes the walk (it cannot overflow)
            if (lane == 0) atomicOr(scal + SC_ERR, (unsigned long long)(DEEP ? ERRF_STACK2 : ERRF_STACK));
            break;
        }
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
            const int n = first + c;
            const double4 A = W[GW_REC * (int64_t)n], V = W[GW_REC * (int64_t)n + 1];   // A = {rCOM, Mass | h_j}, V = {(2L)^2, radius, child info, range}
            const double dx = px - A.x, dy = py - A.y, dz = pz - A.z;   // p_i - rCOM (:255)
            const double d_sq = sph_d2_exact(dx, dy, dz);                // (:256)
            if (COUNT && mine) ++visits;
            if ((leafmask >> c) & 1) {
                // leaf = one particle j; A.w carries h_j, its mass is m.  The target's own leaf is skipped
                // (the reference removes it from its parent's child list, :293-294).
                if (mine && (int64_t)unpack_i2(V.z).x != s) {
                    double gP, pot;
                    leaf_pair(d_sq, hi, A.w, gP, pot);
                    const double mg = m * gP;
                    gx += mg * dx; gy += mg * dy; gz += mg * dz;         // (:263)
                    ph += m * pot;                                       // (:264)
                }
            } else {
                bool open = false;
                if (mine) {
                    if (cell_accepted(t.nodeBC, n, V.x, V.y, d_sq, px, py, pz, hi2, h2x, theta_sq, th_lo, th_hi)) {
                        const double rinv = fast_rsqrt(d_sq);
                        const double f = A.w * (rinv * rinv * rinv);         // Mass / d^3  (:266-268)
                        gx += f * dx; gy += f * dy; gz += f * dz;
                        ph -= A.w * rinv;                                    // -Mass / d   (:269)
                    } else {
                        open = true;
                    }
                }
                const unsigned om = __ballot_sync(0xffffffffu, open);
                if (om) {
                    if (lane == 0) {
                        const int2 ci = unpack_i2(V.z);
                        stack[sp] = make_int4(ci.x, ci.y, (int)om, 0);
                    }
                    ++sp;
                }
            }
        }
        __syncwarp();
    }
    if (active) {
        double *out = part + (size_t)row * 4 * chunk;
        out[local] = gx; out[local + chunk] = gy; out[local + 2 * chunk] = gz;
        out[local + 3 * chunk] = row == 0 ? ph - (m * (7.0 / 5) / hi) : ph;       // (:303), once
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) visits += __shfl_xor_sync(0xffffffffu, visits, o);
        if (lane == 0) atomicAdd(scal + SC_VISITS, visits);
    }
}

// The regular variants launch one block per work item.  The DEEP variants are launched after every regular walk but
// have work only when it overflowed: a small persistent grid loops over the work items, so that the usual case costs a
// few hundred blocks that leave at once instead of one empty block per work item (36 us at N = 1e6).
template <bool COUNT, bool DEEP>
__global__ void __launch_bounds__(GW_WARPS * 32, DEEP ? 4 : 8) walk_kernel(int64_t N, int nranks, int rank, int64_t chunk,
                                                                 const double4 *__restrict__ pos4,
                                                                 const double *__restrict__ hs, SphTree t,
                                                                 double theta_sq, double m,
                                                                 unsigned long long *__restrict__ scal,
                                                                 double *__restrict__ part /* [rows][4][chunk] */,
                                                                 int nbx, int rows) {
    constexpr int STACK = DEEP ? GW_STACK_DEEP : GW_STACK;
    __shared__ int4 s_stack[GW_WARPS][STACK];   // {first child, nch | leafmask << 8, lane mask, -}
    if (walk_skipped<DEEP>(scal)) return;
    int4 *stack = s_stack[threadIdx.x >> 5];
    if (!DEEP) {
        walk_tile<COUNT, DEEP, STACK>(blockIdx.x, blockIdx.y, rows, stack, N, nranks, rank, chunk, pos4, hs, t, theta_sq, m, scal, part);
    } else {
        for (int w = blockIdx.x; w < nbx * rows; w += gridDim.x) {
            walk_tile<COUNT, DEEP, STACK>(w % nbx, w / nbx, rows, stack, N, nranks, rank, chunk, pos4, hs, t, theta_sq, m, scal, part);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Walk with a pair queue for the sparse part (default; SPH_B200_WALK_DFS=1 selects walk_kernel above).
//
// In the shared walk a cell that only a few lanes still have to open costs a full warp iteration per child:
// at N = 1e6 cells with <= 12 interested lanes are half of all warp iterations but carry a seventh of the
// particle-cell visits.  Here such a cell is not opened by the warp; instead every (child, target) PAIR goes
// to a per-warp LIFO queue in shared memory, and whenever 32 pairs are queued one round evaluates them with
// one pair per lane: the lane fetches its target {x, y, z, h} from the owning lane by shuffle, loads the
// pair's node, decides the reference's rule for that target exactly as the shared walk does
// (cell_accepted / leaf_pair are the same code), and either produces the interaction or queues the children
// of the cell for the same target.  Interactions of a round are added to per-target accumulators in shared
// memory; pairs of one round that belong to the same target are serialised by their lane rank, so the sums
// are deterministic.  The set of (particle, node) visits - hence every decision - is the one of the
// reference; only the summation order differs.
// The queue cannot overflow in a tree of up to 21 levels: a round pops B <= (SOFT - qn) / 7 pairs (each pushes at most 8
// children); B = 1 is a depth-first descent whose excursion is bounded by 7 per level + 8 entries (the slack above
// SOFT), and cells are only expanded into the queue while it holds < 32 pairs.  Deeper trees (two-word keys, up to 42
// levels) almost never need more - single-child chains do not grow a stack - but the bound no longer holds, so stack
// and queue are guarded: an overflow raises ERRF_STACK and sph_launch_walk's DEEP variant (sized for 42 levels, half
// the occupancy) redoes the walk of this evaluation.
// ---------------------------------------------------------------------------------------------------
constexpr int GP_TMAX = 16;              // cells with <= sparse_t <= GP_TMAX interested lanes go to the pair queue
constexpr int GP_STACK = 160;            // shared stack of dense cells: 7 per level + 8 covers 21 levels
constexpr int GP_SOFT = 352;
constexpr int GP_SLACK = 160;            // depth-first excursion of the queue above GP_SOFT: 7 per level + 8
constexpr int GP_DEEP = 304;             // both, for all 42 levels (DEEP variant)

template <int STACK, int SLACK>
struct GpWarp {
    int2 stack[STACK];                   // {first child | nch << 27, lane mask}
    double4 acc[32];                     // sparse-path sums {gx, gy, gz, phi} of the warp's 32 targets
    int q[GP_SOFT + SLACK];              // pairs: node | target lane << 27
    double2 rec[32];                     // the walk records of the popped cell's <= 8 children (64 B each)
};

template <bool COUNT, bool DEEP, int STACK, int QCAP>
__device__ __forceinline__ void walk_pairs_tile(int bx, int row, int rows, GpWarp<STACK, QCAP - GP_SOFT> &sm, int64_t N, int nranks,
                                                int rank, int64_t chunk, const double4 *__restrict__ pos4,
                                                const double *__restrict__ hs, const SphTree &t, double theta_sq, double th_lo,
                                                double th_hi, double m, int sparse_t, unsigned long long *__restrict__ scal,
                                                double *__restrict__ part /* [rows][4][chunk] */) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const double4 *__restrict__ W = t.nodeW;
    const int64_t local = (int64_t)bx * (GW_WARPS * 32) + threadIdx.x;
    const int64_t gtile = ((int64_t)(bx / SPH_WALK_DEAL) * nranks + rank) * SPH_WALK_DEAL + bx % SPH_WALK_DEAL;
    const int64_t s = gtile * (GW_WARPS * 32) + threadIdx.x;
    const bool active = s < N;
    double px = 0, py = 0, pz = 0, hi = 1.0;
    if (active) {
        const double4 p = pos4[s];
        px = p.x; py = p.y; pz = p.z; hi = hs[s];
    }
    const int sbase = (int)(s - lane);               // sorted slot of lane 0's target
    const double hi2 = hi * hi;
    const double h2x = 2.0 * hi * (1.0 + 1e-9);      // clause 2 is certainly true when mindist > h2x
    double gx = 0.0, gy = 0.0, gz = 0.0, ph = 0.0;
    unsigned long long visits = 0;
    const unsigned amask = __ballot_sync(0xffffffffu, active);
    // the root is opened unconditionally (:246-249); its children are dealt to the block rows (see walk_kernel)
    const int2 R = unpack_i2(W[1].z);
    const int rnch = R.y & 0xff;
    if (row >= rnch) return;
    sm.acc[lane] = make_double4(0.0, 0.0, 0.0, 0.0);
    int sp = 0;
    if (amask) {
        const int near = walk_root_near(t.nodeBC, R, gtile * (GW_WARPS * 32));
        if (lane == 0) {
            int o = row;
            while (o + rows < rnch) o += rows;
            for (; o >= 0; o -= rows) {               // pushed far to near: the near child is walked first
                const int rc = near + o >= rnch ? near + o - rnch : near + o;
                sm.stack[sp++] = make_int2((R.x + rc) | (1 << 27), (int)amask);
            }
        }
        sp = __shfl_sync(0xffffffffu, sp, 0);
    }
    __syncwarp();
    int qn = 0;
    for (;;) {
        if (qn >= 32 || (sp == 0 && qn > 0)) {
            // ---- one round of the pair queue: lane k evaluates pair k of the top B
            int B = (GP_SOFT - qn) / 7;
            B = B < 1 ? 1 : (B > 32 ? 32 : B);
            B = B < qn ? B : qn;
            qn -= B;
            const bool v = lane < B;
            unsigned pr = 0;
            if (v) pr = (unsigned)sm.q[qn + lane];
            __syncwarp();
            const int n = (int)(pr & 0x7ffffffu);
            const int tl = v ? (int)(pr >> 27) : lane;
            const double tx = __shfl_sync(0xffffffffu, px, tl), ty = __shfl_sync(0xffffffffu, py, tl),
                         tz = __shfl_sync(0xffffffffu, pz, tl), th = __shfl_sync(0xffffffffu, hi, tl);
            double fx = 0.0, fy = 0.0, fz = 0.0, fp = 0.0;
            bool open = false, has = false;
            int cfirst = 0, cn = 0;
            if (v) {
                const double4 A = W[GW_REC * (int64_t)n], V = W[GW_REC * (int64_t)n + 1];
                const double dx = tx - A.x, dy = ty - A.y, dz = tz - A.z;   // p_i - rCOM (:255)
                const double d_sq = sph_d2_exact(dx, dy, dz);                // (:256)
                if (COUNT) ++visits;
                const int2 ci = unpack_i2(V.z);
                if ((ci.y & 0xff) == 0) {
                    // leaf = particle in sorted slot ci.x; the target's own leaf is skipped (:293-294)
                    if (ci.x != sbase + tl) {
                        double gP, pot;
                        leaf_pair(d_sq, th, A.w, gP, pot);
                        const double mg = m * gP;
                        fx = mg * dx; fy = mg * dy; fz = mg * dz;            // (:263)
                        fp = m * pot;                                        // (:264)
                        has = true;
                    }
                } else if (cell_accepted(t.nodeBC, n, V.x, V.y, d_sq, tx, ty, tz, th * th, 2.0 * th * (1.0 + 1e-9), theta_sq, th_lo, th_hi)) {
                    const double rinv = fast_rsqrt(d_sq);
                    const double f = A.w * (rinv * rinv * rinv);             // Mass / d^3  (:266-268)
                    fx = f * dx; fy = f * dy; fz = f * dz;
                    fp = -(A.w * rinv);                                      // -Mass / d   (:269)
                    has = true;
                } else {
                    open = true;
                    cfirst = ci.x; cn = ci.y & 0xff;
                }
            }
            // children of the opened cells: lane k's cell takes cn slots after those of the lanes below it
            if (__any_sync(0xffffffffu, open)) {
                int off = cn;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int a = __shfl_up_sync(0xffffffffu, off, o);
                    if (lane >= o) off += a;
                }
                const int total = __shfl_sync(0xffffffffu, off, 31);
                if (qn + total > QCAP) {       // only in trees deeper than 21 levels: the DEEP variant takes over
                    if (lane == 0) atomicOr(scal + SC_ERR, (unsigned long long)(DEEP ? ERRF_STACK2 : ERRF_STACK));
                    sp = 0; qn = 0;
                    continue;
                }
                int *dst = sm.q + (qn + off - cn);
                const unsigned tag = (unsigned)tl << 27;
#pragma unroll 1
                for (int c = 0; c < cn; ++c) dst[c] = (int)((unsigned)(cfirst + c) | tag);
                qn += total;
            }
            // interactions -> per-target sums; pairs of the same target take turns in lane order
            const unsigned peers = __match_any_sync(0xffffffffu, has ? tl : 32 + lane);
            const int prank = __popc(peers & lt);
            const int mult = __reduce_max_sync(0xffffffffu, has ? __popc(peers) : 0);
            for (int r = 0; r < mult; ++r) {
                if (has && prank == r) {
                    double4 a = sm.acc[tl];
                    a.x += fx; a.y += fy; a.z += fz; a.w += fp;
                    sm.acc[tl] = a;
                }
                __syncwarp();
            }
            __syncwarp();
            continue;
        }
        if (sp == 0) break;
        const int2 top = sm.stack[--sp];
        __syncwarp();
        const unsigned mask = (unsigned)top.y;
        const bool mine = (mask >> lane) & 1u;
        const int first = top.x & 0x7ffffff, nch = top.x >> 27;
        const int pc = __popc(mask);
        if (pc <= sparse_t) {
            // few lanes are interested: their (child, target) pairs go to the queue, child-major (qn < 32 here)
            if (mine) {
                const int r = __popc(mask & lt);
                for (int c = 0; c < nch; ++c) sm.q[qn + c * pc + r] = (int)((unsigned)(first + c) | ((unsigned)lane << 27));
            }
            qn += pc * nch;
            __syncwarp();
            continue;
        }
        if (sp + nch > STACK) {     // only in trees deeper than 21 levels: the DEEP variant takes over
            if (lane == 0) atomicOr(scal + SC_ERR, (unsigned long long)(DEEP ? ERRF_STACK2 : ERRF_STACK));
            break;
        }
        // the records of all children (contiguous in BFS order, 64 B each) come into shared memory with ONE coalesced
        // load per popped cell; the child loop then reads them with the short, fixed latency of shared memory instead
        // of exposing a global-load latency per child (walk alone at N = 1e6: 5.76 -> 5.56 ms)
        if (lane < 4 * nch) sm.rec[lane] = reinterpret_cast<const double2 *>(W + GW_REC * (int64_t)first)[lane];
        __syncwarp();
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
            const int n = first + c;
            const double4 A = reinterpret_cast<const double4 *>(sm.rec)[2 * c], V = reinterpret_cast<const double4 *>(sm.rec)[2 * c + 1];   // A = {rCOM, Mass | h_j}, V = {(2L)^2, radius, child info, range}
            const double dx = px - A.x, dy = py - A.y, dz = pz - A.z;   // p_i - rCOM (:255)
            const double d_sq = sph_d2_exact(dx, dy, dz);                // (:256)
            if (COUNT && mine) ++visits;
            const int2 ci = unpack_i2(V.z);
            if ((ci.y & 0xff) == 0) {
                // leaf = one particle j; A.w carries h_j, its mass is m; the target's own leaf is skipped (:293-294)
                if (mine && ci.x != sbase + lane) {
                    double gP, pot;
                    leaf_pair(d_sq, hi, A.w, gP, pot);
                    const double mg = m * gP;
                    gx += mg * dx; gy += mg * dy; gz += mg * dz;         // (:263)
                    ph += m * pot;                                       // (:264)
                }
            } else {
                bool open = false;
                if (mine) {
                    if (cell_accepted(t.nodeBC, n, V.x, V.y, d_sq, px, py, pz, hi2, h2x, theta_sq, th_lo, th_hi)) {
                        const double rinv = fast_rsqrt(d_sq);
                        const double f = A.w * (rinv * rinv * rinv);         // Mass / d^3  (:266-268)
                        gx += f * dx; gy += f * dy; gz += f * dz;
                        ph -= A.w * rinv;                                    // -Mass / d   (:269)
                    } else {
                        open = true;
                    }
                }
                const unsigned om = __ballot_sync(0xffffffffu, open);
                if (om) {
                    sm.stack[sp] = make_int2(ci.x | ((ci.y & 0xff) << 27), (int)om);   // every lane stores the same entry (no lane test)
                    ++sp;
                }
            }
        }
        __syncwarp();
    }
    if (active) {
        const double4 a = sm.acc[lane];
        gx += a.x; gy += a.y; gz += a.z; ph += a.w;
        double *out = part + (size_t)row * 4 * chunk;
        out[local] = gx; out[local + chunk] = gy; out[local + 2 * chunk] = gz;
        out[local + 3 * chunk] = row == 0 ? ph - (m * (7.0 / 5) / hi) : ph;       // (:303), once
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) visits += __shfl_xor_sync(0xffffffffu, visits, o);
        if (lane == 0) atomicAdd(scal + SC_VISITS, visits);
    }
}

template <bool COUNT, bool DEEP>
__global__ void __launch_bounds__(GW_WARPS * 32, DEEP ? 4 : WALK_MINB) walk_pairs_kernel(int64_t N, int nranks, int rank, int64_t chunk,
                                                                       const double4 *__restrict__ pos4,
                                                                       const double *__restrict__ hs, SphTree t,
                                                                       double theta_sq, double th_lo, double th_hi, double m,
                                                                       int sparse_t, unsigned long long *__restrict__ scal,
                                                                       double *__restrict__ part /* [rows][4][chunk] */,
                                                                       int nbx, int rows) {
    constexpr int STACK = DEEP ? GP_DEEP : GP_STACK, QCAP = GP_SOFT + (DEEP ? GP_DEEP : GP_SLACK);
    __shared__ GpWarp<STACK, QCAP - GP_SOFT> s_w[GW_WARPS];
    if (walk_skipped<DEEP>(scal)) return;
    GpWarp<STACK, QCAP - GP_SOFT> &sm = s_w[threadIdx.x >> 5];
    if (!DEEP) {
        walk_pairs_tile<COUNT, DEEP, STACK, QCAP>(blockIdx.x, blockIdx.y, rows, sm, N, nranks, rank, chunk, pos4, hs, t, theta_sq, th_lo,
                                                  th_hi, m, sparse_t, scal, part);
    } else {   // see walk_kernel
        for (int w = blockIdx.x; w < nbx * rows; w += gridDim.x) {
            walk_pairs_tile<COUNT, DEEP, STACK, QCAP>(w % nbx, w / nbx, rows, sm, N, nranks, rank, chunk, pos4, hs, t, theta_sq, th_lo,
                                                      th_hi, m, sparse_t, scal, part);
            __syncwarp();
        }
    }
}

// SPH_B200_WALK_FORCE_DEEP=1 (tests): pretend the regular walk overflowed, so that the DEEP variant produces the result
__global__ void walk_force_overflow_kernel(unsigned long long *__restrict__ scal) {
    if (scal[SC_ERR] == 0ull) scal[SC_ERR] = (unsigned long long)ERRF_STACK;
}

// after the DEEP variant: the regular walk's overflow is dealt with
__global__ void walk_clear_overflow_kernel(unsigned long long *__restrict__ scal) {
    if (scal[SC_ERR] & (unsigned long long)ERRF_STACK) scal[SC_ERR] &= ~(unsigned long long)ERRF_STACK;
}

// sum of the per-row partial results in row order -> this rank's section of walk_buf
__global__ void walk_reduce_kernel(int64_t n4 /* 4 * chunk */, int rows, const double *__restrict__ part,
                                   const double4 *__restrict__ W, const unsigned long long *__restrict__ scal,
                                   double *__restrict__ out) {
    if (scal[SC_ERR] != 0ull) return;
    const int nch = unpack_i2(W[1].z).y & 0xff;
    const int nr = rows < nch ? rows : nch;          // rows beyond the root's child count wrote nothing
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < nr; ++c) s += part[(size_t)c * n4 + i];
        out[i] = s;
    }
}

// after the search: one 64-byte walk record per node (leaves carry h_j instead of their constant mass m)
__global__ void pack_nodes_kernel(SphTree t, const double *__restrict__ hs, const unsigned long long *__restrict__ scal) {
    if (scal[SC_ERR] != 0ull) return;
    const int64_t M = (int64_t)scal[SC_NNODES];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < M; k += (int64_t)gridDim.x * blockDim.x) {
        const int2 I = t.nodeI[k];
        double4 A = t.nodeA[k];
        double2 D = make_double2(0.0, 0.0);
        if (I.y == 0) A.w = hs[I.x];
        else D = t.nodeD[k];
        double4 *rec = t.nodeW + GW_REC * k;
        rec[0] = A;
        rec[1] = make_double4(D.x, D.y, pack_i2(I.x, I.y), pack_i2(t.nstart[k], t.ncount[k]));
    }
}

}  // namespace

cudaError_t sph_launch_walk(sph_handle *h) {
    static_assert(GW_WARPS * 32 == 128, "walk tiles are 128 targets (finish_kernel and sph_comm_init assume it)");
    sph_note(1);
    pack_nodes_kernel<<<148 * 8, 256, 0, h->stream>>>(h->tree, h->hs, h->scal);
    const int64_t tiles = (h->N + 127) / 128;
    const int64_t groups = (tiles + SPH_WALK_DEAL - 1) / SPH_WALK_DEAL;
    const int64_t blocks = (groups - h->rank + h->nranks - 1) / h->nranks * SPH_WALK_DEAL;   // groups rank, rank + P, ...
    if (blocks <= 0) {
        cudaEventRecord(h->wev[0], h->stream);
        cudaEventRecord(h->wev[1], h->stream);
        return cudaGetLastError();
    }
    const double th2 = h->p.theta * h->p.theta;
    double *out = h->walk_buf + (size_t)h->rank * 4 * h->walk_chunk;
    // block rows = work items per tile (the root's children are dealt to them): 8 - with more, shorter items the tail of
    // the launch stays short when the grid is only a wave or two of blocks (N = 1e5 on one GPU: 1.08 vs 1.50 ms with 2
    // rows).  A rank that owns many waves of tiles (>= 6144 tiles = 5 waves of 148 x 8 blocks; 4 rows from 4096 tiles, as
    // measured earlier in the round) takes ONE: the tail no
    // longer matters, the kernel writes its sums straight into walk_buf, and the per-row partial sums (32 B per target
    // and row, written here, read by walk_reduce - the largest DRAM item of the launch) and the reduce kernel go away.
    // Evaluation at N = 1e6 beside the SPH kernels with 8 / 4 / 3 / 2 / 1 rows: 9.39 / 9.30 / 9.29 / 9.23 / 9.16 ms
    // (profiles/r02_walk_variants.txt, r4n / r4o).  SPH_B200_WALK_ROWS overrides (1..8).
    static const int rows_env = getenv("SPH_B200_WALK_ROWS") ? atoi(getenv("SPH_B200_WALK_ROWS")) : 0;
    int rows = rows_env > 0 ? rows_env : (blocks >= 6144 ? 1 : (blocks >= 4096 ? 4 : 8));
    rows = rows < 1 ? 1 : (rows > 8 ? 8 : rows);
    double *part = rows == 1 ? out : h->walk_part;
    const dim3 grid((unsigned)blocks, (unsigned)rows);
    sph_note(1);
    cudaEventRecord(h->wev[0], h->stream);
    static const bool count_env = getenv("SPH_B200_COUNT_VISITS") != nullptr;
    const bool count = count_env || (h->p.flags & SPH_FLAG_COUNT_VISITS) != 0;
    static const bool shared_only = getenv("SPH_B200_WALK_DFS") != nullptr;
    // cells with <= sparse_t interested lanes are evaluated pair-wise (walk_pairs_kernel); 0 = never
    static const int sparse_t = [] {
        const char *e = getenv("SPH_B200_WALK_T");
        const int v = e ? atoi(e) : 10;
        return v < 0 ? 0 : (v > GP_TMAX ? GP_TMAX : v);
    }();
    const double lo = th2 * (1.0 - 1e-15), hi = th2 * (1.0 + 1e-15);
    const int64_t N = h->N;
    const bool shared = shared_only || h->tree.cap >= (1ll << 27) || h->N >= (1ll << 31);   // node ids that do not fit the pair encoding
    const int nbx = (int)blocks;
    const dim3 grid_deep(148 * 4);       // persistent: loops over the nbx x rows work items when it has to run
#define WALK_LAUNCH(DEEP, GRID)                                                                                                 \
    do {                                                                                                                        \
        if (shared) {                                                                                                           \
            if (count) walk_kernel<true, DEEP><<<GRID, GW_WARPS * 32, 0, h->stream>>>(N, h->nranks, h->rank, h->walk_chunk, h->pos4, h->hs, h->tree, th2, h->p.m, h->scal, part, nbx, rows);   \
            else walk_kernel<false, DEEP><<<GRID, GW_WARPS * 32, 0, h->stream>>>(N, h->nranks, h->rank, h->walk_chunk, h->pos4, h->hs, h->tree, th2, h->p.m, h->scal, part, nbx, rows);        \
        } else {                                                                                                                \
            if (count) walk_pairs_kernel<true, DEEP><<<GRID, GW_WARPS * 32, 0, h->stream>>>(N, h->nranks, h->rank, h->walk_chunk, h->pos4, h->hs, h->tree, th2, lo, hi, h->p.m, sparse_t, h->scal, part, nbx, rows);   \
            else walk_pairs_kernel<false, DEEP><<<GRID, GW_WARPS * 32, 0, h->stream>>>(N, h->nranks, h->rank, h->walk_chunk, h->pos4, h->hs, h->tree, th2, lo, hi, h->p.m, sparse_t, h->scal, part, nbx, rows);        \
        }                                                                                                                       \
    } while (0)
    WALK_LAUNCH(false, grid);
    cudaEventRecord(h->wev[1], h->stream);
    // trees deeper than 21 levels may overflow the regular variant's stack: the DEEP variant (a no-op otherwise: its
    // blocks leave at once) then redoes every tile with stacks sized for 42 levels
    sph_note(2);
    static const bool force_deep = getenv("SPH_B200_WALK_FORCE_DEEP") != nullptr;
    if (force_deep) walk_force_overflow_kernel<<<1, 1, 0, h->stream>>>(h->scal);
    WALK_LAUNCH(true, grid_deep);
#undef WALK_LAUNCH
    walk_clear_overflow_kernel<<<1, 1, 0, h->stream>>>(h->scal);
    if (rows > 1) {
        sph_note(1);
        walk_reduce_kernel<<<148 * 8, 256, 0, h->stream>>>(4 * h->walk_chunk, rows, h->walk_part, h->tree.nodeW, h->scal, out);
    }
    return cudaGetLastError();
}
