// knn.cu -- exact K-nearest-neighbour search on the key-sorted particles and their octree.
//
// Replaces NearestNeighbors.KDTree + knn(tree, ri', K, true) as called by HJL.getNeighbors
// (F/isothermal_hydroKDTree.jl:128-142): for every particle the K nearest particles INCLUDING itself,
// ascending distance.  Distances are evaluated exactly like the oracle's restatement
// (d2 = (dx*dx + dy*dy) + dz*dz, no FMA), ties are ordered by the caller's particle index, so the
// emitted lists are bit-identical to the CPU path.
//
// Two kernels:
//   knn_kernel       one warp per target.  Guaranteed radius R0: the K particles around the target in key order
//                    all lie within max_j d2(target, j), so that ball holds at least K particles; the trial radius
//                    is (1.1 * 2 h_prev)^2 from the previous evaluation when smaller.  Depth-first walk: up to 8
//                    children of a cell are tested by 8 lanes (point-to-box distance with an absolute slack for
//                    the <= 1 ulp mismatch between the reference's classification centre and its stored bounds);
//                    cells holding <= 32 particles are scanned as contiguous ranges (coalesced), hits go to a
//                    shared-memory buffer by ballot prefix; a full buffer is compacted by a bitonic sort that keeps
//                    the K best and tightens the radius.  If the trial ball held < K particles the search restarts
//                    from R0.  Used for the first evaluation (no hint), for arbitrary query points (density_plot)
//                    and for the targets the second kernel hands over.
//   knn_quad_kernel  four targets per warp once hints exist (see its header); exactness never depends on the hint.
#include "sph_internal.cuh"

#include <climits>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int KNN_WARPS = 8;
constexpr int KNN_BUCKET = 32;
constexpr int KNN_STACK = 304;   // cells to visit: 7 pending siblings per level + 8, 42 levels

__device__ __forceinline__ bool cand_less(double da, int ia, double db, int ib, const int *__restrict__ perm) {
    if (da < db) return true;
    if (da > db) return false;
    if (ia == ib) return false;
    const int oa = ia < 0 ? INT_MAX : (perm ? perm[ia] : ia);
    const int ob = ib < 0 ? INT_MAX : (perm ? perm[ib] : ib);
    return oa < ob;
}

template <int S>
__device__ __forceinline__ void warp_bitonic(double *d2, int *id, const int *__restrict__ perm, int lane) {
#pragma unroll 1
    for (int k = 2; k <= S; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int t = lane; t < S / 2; t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = (i & k) == 0;
                const double a = d2[i], b = d2[p];
                const int ia = id[i], ib = id[p];
                const bool sw = up ? cand_less(b, ib, a, ia, perm) : cand_less(a, ia, b, ib, perm);
                if (sw) { d2[i] = b; d2[p] = a; id[i] = ib; id[p] = ia; }
            }
            __syncwarp();
        }
    }
}

// buffer full: keep the K best candidates and tighten the search radius to the K-th of them
template <int CAP>
__device__ __forceinline__ void knn_compact(double *bd2, int *bid, const int *__restrict__ perm, int lane, int K,
                                            int &cnt, double &R2) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    for (int i = cnt + lane; i < CAP; i += 32) { bd2[i] = INF; bid[i] = -1; }
    __syncwarp();
    warp_bitonic<CAP>(bd2, bid, perm, lane);
    if (cnt >= K) { cnt = K; R2 = bd2[K - 1]; }
    __syncwarp();
}

// SELF = true : queries are the sorted particles themselves (targets t0..t1), output = neighbour lists
// SELF = false: queries are arbitrary points (density_plot), output = the K sorted squared distances
template <int CAP, bool SELF>
__global__ void __launch_bounds__(KNN_WARPS * 32, 5) knn_kernel(int64_t N, int64_t NL, int K, int64_t t0, int64_t t1,
                                                              const double4 *__restrict__ pos4,
                                                              const double *__restrict__ qpts, int64_t qstride,
                                                              const int *__restrict__ perm, SphTree t,
                                                              const double *__restrict__ hint_h, double hint_fac2,
                                                              const int *__restrict__ list,
                                                              unsigned long long *__restrict__ scal,
                                                              int *__restrict__ nbr, double *__restrict__ d2k,
                                                              int *__restrict__ kid, double *__restrict__ d2_out) {
    __shared__ double s_d2[KNN_WARPS][CAP];
    __shared__ int s_id[KNN_WARPS][CAP];
    __shared__ int s_stack[KNN_WARPS][KNN_STACK];
    if (scal[SC_ERR] != 0ull) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    double *bd2 = s_d2[warp];
    int *bid = s_id[warp];
    int *stack = s_stack[warp];
    const double ldom = __longlong_as_double((long long)scal[SC_LDOM]);
    const double eps = ldom * 1e-14;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    unsigned long long retries = 0;

    // targets: the range [t0, t1), or - for the second pass of the tiled search - the listed slots
    const int64_t nwarps = (int64_t)gridDim.x * KNN_WARPS;
    const int64_t ntargets = list ? (int64_t)scal[SC_KNN_RETRY] : t1 - t0;
    for (int64_t it = (int64_t)blockIdx.x * KNN_WARPS + warp; it < ntargets; it += nwarps) {
        const int64_t s = list ? (int64_t)list[it] : t0 + it;
        double qx, qy, qz;
        if (SELF) {
            const double4 q = pos4[s];
            qx = q.x; qy = q.y; qz = q.z;
        } else {
            qx = qpts[s]; qy = qpts[s + qstride]; qz = qpts[s + 2 * qstride];
        }
        // ---- radii
        double R0sq = INF;
        if (SELF) {
            int64_t w0 = s - K / 2;
            if (w0 > N - K) w0 = N - K;
            if (w0 < 0) w0 = 0;
            double mx = 0.0;
            for (int j = lane; j < K; j += 32) {
                const double4 p = pos4[w0 + j];
                mx = fmax(mx, sph_d2_exact(qx - p.x, qy - p.y, qz - p.z));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            R0sq = mx;
        }
        double Rsq = R0sq;
        bool guaranteed = true;
        if (SELF && hint_h) {
            const double hh = hint_h[perm[s]];
            const double tr = 4.0 * hh * hh * hint_fac2;
            if (hh > 0.0 && tr < R0sq) { Rsq = tr; guaranteed = false; }
        }

        int cnt = 0;
        double R2 = Rsq;
        for (;;) {
            cnt = 0;
            R2 = Rsq;
            int sp = 1;
            if (lane == 0) stack[0] = 0;
            __syncwarp();
            while (sp > 0) {
                const int n = stack[--sp];
                __syncwarp();
                const int2 I = t.nodeI[n];
                const int nch = I.y & 0xff, first = I.x;
                bool pass = false;
                int cstart = 0, ccount = 0;
                if (lane < nch) {
                    const int c = first + lane;
                    const double4 B = t.nodeB[c];
                    const double4 C = t.nodeC[c];
                    // point-to-box distance per axis, shrunk by eps (compare/select instead of NaN-aware fmax)
                    double ax = B.x - qx, bx = qx - B.w, ay = B.y - qy, by = qy - C.x, az = B.z - qz, bz = qz - C.y;
                    ax = (ax > bx ? ax : bx) - eps; ay = (ay > by ? ay : by) - eps; az = (az > bz ? az : bz) - eps;
                    ax = ax > 0.0 ? ax : 0.0; ay = ay > 0.0 ? ay : 0.0; az = az > 0.0 ? az : 0.0;
                    const double md2 = ax * ax + ay * ay + az * az;
                    pass = md2 * (1.0 - 1e-12) <= R2;
                    cstart = t.nstart[c];
                    ccount = t.ncount[c];
                }
                const bool is_bucket = ccount <= KNN_BUCKET;
                unsigned bm = __ballot_sync(0xffffffffu, pass && is_bucket);
                const unsigned im = __ballot_sync(0xffffffffu, pass && !is_bucket);
                if (pass && !is_bucket) stack[sp + __popc(im & lt)] = first + lane;
                sp += __popc(im);
                __syncwarp();
                while (bm) {
                    const int cl = __ffs(bm) - 1;
                    bm &= bm - 1;
                    const int bs = __shfl_sync(0xffffffffu, cstart, cl);
                    const int bc = __shfl_sync(0xffffffffu, ccount, cl);
                    const int j = bs + lane;
                    const bool v = lane < bc;
                    double d2 = INF;
                    if (v) {
                        const double4 p = pos4[j];
                        d2 = sph_d2_exact(qx - p.x, qy - p.y, qz - p.z);
                    }
                    const bool ok = v && d2 <= R2;
                    const unsigned om = __ballot_sync(0xffffffffu, ok);
                    if (ok) {
                        const int slot = cnt + __popc(om & lt);
                        bd2[slot] = d2;
                        bid[slot] = j;
                    }
                    cnt += __popc(om);
                    __syncwarp();
                    if (cnt > CAP - 32) knn_compact<CAP>(bd2, bid, perm, lane, K, cnt, R2);
                }
            }
            if (guaranteed || cnt >= K) break;
            Rsq = R0sq;      // the hinted ball held fewer than K particles: fall back to the guaranteed radius
            guaranteed = true;
            ++retries;
        }
        // ---- final ordering by (d2, particle id)
        if (cnt <= 64 && CAP >= 64) {
            for (int i = cnt + lane; i < 64; i += 32) { bd2[i] = INF; bid[i] = -1; }
            __syncwarp();
            warp_bitonic<64>(bd2, bid, perm, lane);
        } else {
            for (int i = cnt + lane; i < CAP; i += 32) { bd2[i] = INF; bid[i] = -1; }
            __syncwarp();
            warp_bitonic<CAP>(bd2, bid, perm, lane);
        }
        if (SELF) {
            for (int j = lane; j < K; j += 32) nbr[s + (int64_t)j * NL] = bid[j];
            // kid: caller's particle id of the K-th entry.  A particle at exactly the K-th distance belongs to this
            // list iff its id is <= kid (the tie order of the sort above): hydro.cu decides list membership from
            // (d2k, kid) alone, without reading other targets' lists
            if (lane == 0) { d2k[s] = bd2[K - 1]; kid[s] = perm[bid[K - 1]]; }
        } else {
            for (int j = lane; j < K; j += 32) d2_out[s + (int64_t)j * qstride] = bd2[j];
        }
        __syncwarp();
    }
    if (!list && retries && lane == 0) atomicAdd(scal + SC_KNN_RETRY, retries);
}

// sph_get_neighbors: one warp per row orders the (possibly unordered) list by (distance, particle id) and writes
// the reference's layout: N x K column-major, 1-based caller ids, ascending distance (F/isothermal_hydroKDTree.jl:131-142)
template <int CAP>
__global__ void __launch_bounds__(KNN_WARPS * 32) export_sorted_kernel(int64_t N, int64_t NL, int K, const double4 *__restrict__ pos4,
                                                                        const int *__restrict__ perm,
                                                                        const int *__restrict__ nbr,
                                                                        int *__restrict__ idx_out, double *__restrict__ r_out) {
    __shared__ double s_d2[KNN_WARPS][CAP];
    __shared__ int s_id[KNN_WARPS][CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *bd2 = s_d2[warp];
    int *bid = s_id[warp];
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    for (int64_t s = (int64_t)blockIdx.x * KNN_WARPS + warp; s < N; s += (int64_t)gridDim.x * KNN_WARPS) {
        const double4 q = pos4[s];
        for (int j = lane; j < CAP; j += 32) {
            if (j < K) {
                const int nj = nbr[s + (int64_t)j * NL];
                const double4 p = pos4[nj];
                bd2[j] = sph_d2_exact(q.x - p.x, q.y - p.y, q.z - p.z);
                bid[j] = nj;
            } else { bd2[j] = INF; bid[j] = -1; }
        }
        __syncwarp();
        warp_bitonic<CAP>(bd2, bid, perm, lane);
        const int64_t i = perm[s];
        for (int j = lane; j < K; j += 32) {
            if (idx_out) idx_out[i + (int64_t)j * N] = perm[bid[j]] + 1;
            if (r_out) r_out[i + (int64_t)j * N] = sqrt(bd2[j]);
        }
        __syncwarp();
    }
}

// density_plot (F/isothermal_hydroKDTree.jl:291-297): h = r_K/2, rho = m * sum_j W(r_j, h), columns in order
__global__ void point_density_kernel(int64_t M, int K, const double *__restrict__ d2s, double m, int poly,
                                     double *__restrict__ rho) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const double PI = 3.141592653589793;
    const double h = sqrt(d2s[i + (int64_t)(K - 1) * M]) / 2;
    const double ct = 1 / (PI * (h * h * h));
    double s = 0.0;
    for (int j = 0; j < K; ++j) {
        const double q = sqrt(d2s[i + (int64_t)j * M]) / h;
        double w = 0.0;
        if (q <= 1.0) w = ct * ((1 - 3.0 / 2 * (q * q)) + 3.0 / 4 * (q * q * q));
        else if (poly || q <= 2.0) { const double u = 2 - q; w = (ct * 1 / 4) * (u * u * u); }
        s += w;
    }
    rho[i] = m * s;
}

// ---------------------------------------------------------------------------------------------------
// Hinted search, four targets per warp (the default from the second force evaluation on).
//
// A warp owns 4 key-adjacent targets, each with its trial ball R_t = f * 2 h_prev (f = 1.06).  Lanes are
// CANDIDATES, as in the warp-per-target kernel, but one tree walk (pruned against the bounding box of the 4
// balls) serves all 4 targets, and the particles of the overlapping buckets are first expanded into a dense
// shared-memory index list so that every test round has 32 busy lanes (buckets hold ~12 particles on average).
// Each lane tests its candidate against the 4 balls and appends it to the matching buffers by ballot prefix.
// A target whose ball held >= K particles (and fit its 96-entry buffer) owns its exact K nearest.  The lists need
// no order (density / force skip the self entry by index, export_sorted_kernel orders on demand), so the K-th
// smallest squared distance is SELECTED: starting from the previous evaluation's (2 h_prev)^2 the warp steps over
// distinct key values and recounts (2-3 steps between two evaluations), then emits the keys <= it in buffer order;
// the register bitonic network / rank sort remains as the fallback after KQ_SEL_STEPS steps (SPH_B200_KNN_SORT=1:
// always).  Targets without a usable hint, with < K particles in the ball, an overflowing buffer or a tie AT the
// K-th distance are queued for the warp-per-target kernel, which restarts from the guaranteed radius and breaks
// ties by particle id: exactness never depends on the hint.
// The walk pops up to 4 cells per iteration (lane group g tests the <= 8 children of cell g against the box); the
// children's {box, child info, particle range} come from the compact nodeBC records the tree build writes, and the
// stack entries carry {first child, count}, so a pop issues no dependent load.
// ---------------------------------------------------------------------------------------------------
constexpr int KQ_T = 4;
constexpr int KQ_CAP = 96;
constexpr int KQ_WARPS = 5;
constexpr int KQ_BLOCKS = 4;     // resident blocks per SM (the register budget of 5 spills: 3.0 vs 2.5 ms)
constexpr int KQ_BUCKET = 64;   // largest cell scanned as a range (list sizing); the default threshold is 32 (SPH_B200_KNN_BUCKET)
constexpr int KQ_BKS = 64;      // buckets gathered before a flush
constexpr int KQ_CAND = 512;    // candidates gathered before a flush (>= 8 * KQ_BUCKET)
constexpr int KQ_STACK = 96;    // cells to visit; a walk that needs more hands its targets to the per-target search
constexpr int KQ_SEL_STEPS = 8;  // selection steps around the previous K-th distance before the sort takes over
constexpr int KQ_MAXCAND = 3072; // quads whose box spans more candidates (key-order jumps) go to the per-target search

struct KqWarp {
    double d2[KQ_T][KQ_CAP];
    int id[KQ_T][KQ_CAP];
    int2 bks[KQ_BKS];
    int cand[KQ_CAND];
    int2 stack[KQ_STACK];   // {first child, number of children} of the cells still to visit
};

// compare-exchange step of the register bitonic sort: keys are the bit patterns of non-negative doubles
__device__ __forceinline__ void kq_cex(unsigned long long &k, int &v, unsigned long long ok, int ov, bool keep_small) {
    const bool other_smaller = ok < k;
    if (other_smaller == keep_small) { k = ok; v = ov; }
}

template <int BLOCKS>
__global__ void __launch_bounds__(KQ_WARPS * 32, BLOCKS) knn_quad_kernel(int64_t N, int64_t NL, int K, int64_t t0, int64_t t1,
                                                                    const double4 *__restrict__ pos4,
                                                                    const int *__restrict__ perm, SphTree t,
                                                                    const double *__restrict__ hint_h, double hint_fac2,
                                                                    int sel_steps, int bucket, unsigned long long *__restrict__ scal,
                                                                    int *__restrict__ retry_list,
                                                                    int *__restrict__ nbr, double *__restrict__ d2k,
                                                                    int *__restrict__ kid) {
    __shared__ KqWarp s_w[KQ_WARPS];
    if (scal[SC_ERR] != 0ull) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    KqWarp &sm = s_w[warp];
    const double ldom = __longlong_as_double((long long)scal[SC_LDOM]);
    const double eps = ldom * 1e-14;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const double inv_fac2 = 1.0 / hint_fac2;

    const int64_t nquads = (t1 - t0 + KQ_T - 1) / KQ_T;
    for (int64_t quad = (int64_t)blockIdx.x * KQ_WARPS + warp; quad < nquads; quad += (int64_t)gridDim.x * KQ_WARPS) {
        const int64_t s0 = t0 + quad * KQ_T;
        double qx[KQ_T], qy[KQ_T], qz[KQ_T], R2[KQ_T];
        int cnt[KQ_T];
        bool ok[KQ_T];       // target is being searched with a finite trial radius
        double blo[3] = {INF, INF, INF}, bhi[3] = {-INF, -INF, -INF};
        bool any = false;
#pragma unroll
        for (int k = 0; k < KQ_T; ++k) {
            const int64_t s = s0 + k;
            qx[k] = qy[k] = qz[k] = 0.0; R2[k] = -1.0; cnt[k] = 0; ok[k] = false;
            if (s < t1) {
                const double4 q = pos4[s];
                qx[k] = q.x; qy[k] = q.y; qz[k] = q.z;
                const double hh = hint_h[perm[s]];
                if (hh > 0.0 && hh < INF) {
                    R2[k] = 4.0 * hh * hh * hint_fac2;
                    ok[k] = true; any = true;
                    const double R = sqrt(R2[k]) * (1.0 + 1e-12) + eps;
                    blo[0] = fmin(blo[0], q.x - R); blo[1] = fmin(blo[1], q.y - R); blo[2] = fmin(blo[2], q.z - R);
                    bhi[0] = fmax(bhi[0], q.x + R); bhi[1] = fmax(bhi[1], q.y + R); bhi[2] = fmax(bhi[2], q.z + R);
                }
            }
        }
        // ---- one walk for the 4 balls (all values above are warp-uniform)
        if (any) {
            int sp = 1, nb = 0, ncand = 0, tested = 0, width = 4;
            bool flush_now = false;
            if (lane == 0) { const int2 I0 = t.nodeI[0]; sm.stack[0] = make_int2(I0.x, I0.y & 0xff); }
            __syncwarp();
            for (;;) {
                const bool last = sp == 0;
                // ---- flush: expand the gathered buckets into a dense candidate list and test it in rounds of 32
                if (nb > 0 && (last || flush_now)) {
                    flush_now = false;
                    tested += ncand;
                    if (tested > KQ_MAXCAND) {
                        // the box of these 4 targets straddles a jump of the key order: four small balls are cheaper
#pragma unroll
                        for (int k = 0; k < KQ_T; ++k) ok[k] = false;
                        break;
                    }
                    // exclusive prefix of the bucket sizes (<= 64 buckets: two per lane)
                    const int c0 = lane < nb ? sm.bks[lane].y : 0, c1 = lane + 32 < nb ? sm.bks[lane + 32].y : 0;
                    int i0 = c0, i1 = c1;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int a0 = __shfl_up_sync(0xffffffffu, i0, o), a1 = __shfl_up_sync(0xffffffffu, i1, o);
                        if (lane >= o) { i0 += a0; i1 += a1; }
                    }
                    const int tot0 = __shfl_sync(0xffffffffu, i0, 31);
                    {
                        const int st0 = lane < nb ? sm.bks[lane].x : 0, st1 = lane + 32 < nb ? sm.bks[lane + 32].x : 0;
                        int *o0 = sm.cand + (i0 - c0), *o1 = sm.cand + (tot0 + i1 - c1);
                        for (int i = 0; i < c0; ++i) o0[i] = st0 + i;
                        for (int i = 0; i < c1; ++i) o1[i] = st1 + i;
                    }
                    __syncwarp();
                    // the positions of the next 32 candidates are requested before the current ones are tested
                    int jn = 0;
                    double4 pn = make_double4(0.0, 0.0, 0.0, 0.0);
                    if (lane < ncand) { jn = sm.cand[lane]; pn = pos4[jn]; }
                    for (int base = 0; base < ncand; base += 32) {
                        const bool v = base + lane < ncand;
                        const int j = jn;
                        const double px = pn.x, py = pn.y, pz = pn.z;
                        if (base + 32 + lane < ncand) { jn = sm.cand[base + 32 + lane]; pn = pos4[jn]; }
#pragma unroll
                        for (int k = 0; k < KQ_T; ++k) {
                            const double d2 = sph_d2_exact(qx[k] - px, qy[k] - py, qz[k] - pz);
                            const bool hit = v && d2 <= R2[k];
                            const unsigned om = __ballot_sync(0xffffffffu, hit);
                            if (om) {
                                const int slot = cnt[k] + __popc(om & lt);
                                if (hit && slot < KQ_CAP) { sm.d2[k][slot] = d2; sm.id[k][slot] = j; }
                                cnt[k] += __popc(om);     // may exceed KQ_CAP: detected below
                            }
                        }
                    }
                    __syncwarp();
                    nb = 0; ncand = 0;
                }
                if (last) break;
                // ---- pop up to 4 cells: lane group g = lane / 8 tests the (up to 8) children of cell g
                int m = width < sp ? width : sp;
                if (sp - m + 8 * m > KQ_STACK) m = 1;
                if (sp + 7 > KQ_STACK) {   // not seen for these pruned walks: the per-target search takes the four targets
#pragma unroll
                    for (int k = 0; k < KQ_T; ++k) ok[k] = false;
                    break;
                }
                const int g = lane >> 3, c8 = lane & 7;
                int2 I = make_int2(0, 0);
                if (g < m) I = sm.stack[sp - 1 - g];
                const int nch = I.y, first = I.x;
                bool pass = false;
                int cstart = 0, ccount = 0;
                int2 Ic = make_int2(0, 0);     // the child's own children: fetched with its box, so a pop needs no dependent load
                if (c8 < nch) {
                    const int c = first + c8;
                    const double4 B = t.nodeBC[2 * (int64_t)c];         // {lo.xyz, hi.x}
                    const double4 C = t.nodeBC[2 * (int64_t)c + 1];     // {hi.y, hi.z, child bits, range bits}
                    Ic = unpack_i2(C.z);
                    pass = B.x <= bhi[0] && B.w >= blo[0] && B.y <= bhi[1] && C.x >= blo[1] && B.z <= bhi[2] && C.y >= blo[2];
                    const int2 rg = unpack_i2(C.w);
                    cstart = rg.x;
                    ccount = rg.y;
                }
                const bool is_bucket = ccount <= bucket;
                const unsigned bm = __ballot_sync(0xffffffffu, pass && is_bucket);
                const unsigned im = __ballot_sync(0xffffffffu, pass && !is_bucket);
                const int add = __reduce_add_sync(0xffffffffu, (pass && is_bucket) ? ccount : 0);
                if (nb + __popc(bm) > KQ_BKS || ncand + add > KQ_CAND) {
                    // does not fit: test what is gathered and come back (nothing was consumed); the children of a
                    // single cell always fit the empty lists (8 buckets, 8 * KQ_BUCKET candidates)
                    if (nb > 0) flush_now = true;
                    else width = 1;
                    continue;
                }
                __syncwarp();   // every group has read its stack entry before the pushes reuse the slots
                sp -= m;
                if (pass && !is_bucket) sm.stack[sp + __popc(im & lt)] = make_int2(Ic.x, Ic.y & 0xff);
                sp += __popc(im);
                if (pass && is_bucket) sm.bks[nb + __popc(bm & lt)] = make_int2(cstart, ccount);
                nb += __popc(bm);
                ncand += add;
                width = 4;
                __syncwarp();
            }
        }
        __syncwarp();
        // ---- per target: order the buffer, emit the first K, or queue for the warp-per-target search
        unsigned retry_mask = 0;
#pragma unroll
        for (int k = 0; k < KQ_T; ++k) {
            const int64_t s = s0 + k;
            if (s >= t1) continue;
            const int n = cnt[k];
            if (!ok[k] || n < K || n > KQ_CAP) { retry_mask |= 1u << k; continue; }
            const double *bd2 = sm.d2[k];
            const int *bid = sm.id[k];
            bool tie = false;
            // ---- selection instead of a sort: the lists need no order (density / force skip the self entry by index,
            // the export sorts on demand).  Keys are the bit patterns of the squared distances (order-preserving).
            // The K-th smallest is reached from the previous evaluation's K-th distance (2 h_prev)^2 by stepping over
            // distinct key values and recounting; between two evaluations the count inside that radius moves by a few.
            bool found = false;
            {
                unsigned long long key[KQ_CAP / 32];
                int vid[KQ_CAP / 32];
#pragma unroll
                for (int r = 0; r < KQ_CAP / 32; ++r) {
                    const int e = lane + 32 * r;
                    key[r] = e < n ? (unsigned long long)__double_as_longlong(bd2[e]) : ~0ull;
                    vid[r] = e < n ? bid[e] : -1;
                }
                unsigned long long cur = (unsigned long long)__double_as_longlong(R2[k] * inv_fac2), kth = 0;
                int cl = 0;
#pragma unroll
                for (int r = 0; r < KQ_CAP / 32; ++r) cl += key[r] <= cur;
                int c = __reduce_add_sync(0xffffffffu, cl);
                if (c < K) {
                    // the (K - c)-th distinct value above cur, at most KQ_SEL_STEPS steps
                    for (int it = 0; it < sel_steps && !found; ++it) {
                        unsigned long long mn = ~0ull;
#pragma unroll
                        for (int r = 0; r < KQ_CAP / 32; ++r)
                            if (key[r] > cur && key[r] < mn) mn = key[r];
                        const unsigned mh = __reduce_min_sync(0xffffffffu, (unsigned)(mn >> 32));
                        const unsigned ml = __reduce_min_sync(0xffffffffu, (unsigned)(mn >> 32) == mh ? (unsigned)mn : 0xffffffffu);
                        cur = ((unsigned long long)mh << 32) | ml;
                        cl = 0;
#pragma unroll
                        for (int r = 0; r < KQ_CAP / 32; ++r) cl += key[r] <= cur;
                        c = __reduce_add_sync(0xffffffffu, cl);
                        if (c >= K) { found = true; kth = cur; tie = c > K; }
                    }
                } else {
                    // walk down over the distinct values <= cur until fewer than K keys lie below
                    for (int it = 0; it < sel_steps && !found; ++it) {
                        unsigned long long mx = 0ull;
#pragma unroll
                        for (int r = 0; r < KQ_CAP / 32; ++r)
                            if (key[r] <= cur && key[r] > mx) mx = key[r];
                        const unsigned mh = __reduce_max_sync(0xffffffffu, (unsigned)(mx >> 32));
                        const unsigned ml = __reduce_max_sync(0xffffffffu, (unsigned)(mx >> 32) == mh ? (unsigned)mx : 0u);
                        mx = ((unsigned long long)mh << 32) | ml;
                        cl = 0;
#pragma unroll
                        for (int r = 0; r < KQ_CAP / 32; ++r) cl += key[r] == mx;
                        const int ceq = __reduce_add_sync(0xffffffffu, cl);
                        if (c - ceq < K) { found = true; kth = mx; tie = c > K; }
                        else { c -= ceq; cur = mx - 1ull; }   // mx > 0 here: the keys equal to 0 alone never reach K
                    }
                }
                if (found && !tie) {
                    // emit the keys <= kth in buffer order (exactly K of them)
                    int base = 0;
#pragma unroll
                    for (int r = 0; r < KQ_CAP / 32; ++r) {
                        const bool sel = key[r] <= kth;
                        const unsigned sb = __ballot_sync(0xffffffffu, sel);
                        if (sel) nbr[s + (int64_t)(base + __popc(sb & lt)) * NL] = vid[r];
                        base += __popc(sb);
                    }
                    if (lane == 0) { d2k[s] = __longlong_as_double((long long)kth); kid[s] = INT_MAX; }   // no tie at the K-th distance
                }
            }
            if (found) {
                // tie at the K-th distance: the exact tie-breaking path decides (below)
            } else if (n <= 64) {
                // register bitonic sort of 64 keys (element e = lane + 32 r), shuffles for the 5 low strides
                unsigned long long k0 = lane < n ? (unsigned long long)__double_as_longlong(bd2[lane]) : ~0ull;
                unsigned long long k1 = lane + 32 < n ? (unsigned long long)__double_as_longlong(bd2[lane + 32]) : ~0ull;
                int v0 = lane < n ? bid[lane] : -1, v1 = lane + 32 < n ? bid[lane + 32] : -1;
#pragma unroll
                for (int kk = 2; kk <= 64; kk <<= 1) {
#pragma unroll
                    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                        if (jj == 32) {
                            // partner of element e is e ^ 32: the lane's own other register; ascending overall (kk == 64)
                            if (k1 < k0) { const unsigned long long tk = k0; k0 = k1; k1 = tk; const int tv = v0; v0 = v1; v1 = tv; }
                        } else {
                            const unsigned long long p0 = __shfl_xor_sync(0xffffffffu, k0, jj), p1 = __shfl_xor_sync(0xffffffffu, k1, jj);
                            const int w0 = __shfl_xor_sync(0xffffffffu, v0, jj), w1 = __shfl_xor_sync(0xffffffffu, v1, jj);
                            const bool lower = (lane & jj) == 0;                 // this lane holds the lower index of the pair
                            const bool up0 = kk == 64 || (lane & kk) == 0;       // element lane      : ascending block?
                            const bool up1 = kk == 64 || ((lane + 32) & kk) == 0;  // element lane + 32 : ascending block?
                            kq_cex(k0, v0, p0, w0, lower == up0);
                            kq_cex(k1, v1, p1, w1, lower == up1);
                        }
                    }
                }
                // equal neighbouring keys = tied distances: let the exact tie-breaking path handle the target
                const unsigned long long nx0 = __shfl_down_sync(0xffffffffu, k0, 1), nx1 = __shfl_down_sync(0xffffffffu, k1, 1);
                const unsigned long long f1 = __shfl_sync(0xffffffffu, k1, 0);
                const bool t0e = (lane < 31 ? nx0 : f1) == k0 && k0 != ~0ull;
                const bool t1e = lane < 31 && nx1 == k1 && k1 != ~0ull;
                tie = __any_sync(0xffffffffu, t0e || t1e);
                if (!tie) {
                    if (lane < K) nbr[s + (int64_t)lane * NL] = v0;
                    if (lane + 32 < K) nbr[s + (int64_t)(lane + 32) * NL] = v1;
                    if (lane == K - 1) d2k[s] = __longlong_as_double((long long)k0);
                    if (lane + 32 == K - 1) d2k[s] = __longlong_as_double((long long)k1);
                    if (lane == 0) kid[s] = INT_MAX;
                }
            } else {
                // rank sort: count the entries that precede each of the lane's three
                unsigned long long e_k[KQ_CAP / 32];
                int e_id[KQ_CAP / 32], e_rank[KQ_CAP / 32];
#pragma unroll
                for (int r = 0; r < KQ_CAP / 32; ++r) {
                    const int e = lane + 32 * r;
                    e_k[r] = e < n ? (unsigned long long)__double_as_longlong(bd2[e]) : ~0ull;
                    e_id[r] = e < n ? bid[e] : -1;
                    e_rank[r] = 0;
                }
                for (int j = 0; j < n; ++j) {
                    const unsigned long long kj = (unsigned long long)__double_as_longlong(bd2[j]);
#pragma unroll
                    for (int r = 0; r < KQ_CAP / 32; ++r) e_rank[r] += kj < e_k[r];
                }
                // ranks are a permutation of 0..n-1 unless two keys are equal
                int rs = 0;
#pragma unroll
                for (int r = 0; r < KQ_CAP / 32; ++r) rs += e_id[r] >= 0 ? e_rank[r] : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
                tie = rs != n * (n - 1) / 2;
                if (!tie) {
#pragma unroll
                    for (int r = 0; r < KQ_CAP / 32; ++r) {
                        if (e_id[r] >= 0 && e_rank[r] < K) {
                            nbr[s + (int64_t)e_rank[r] * NL] = e_id[r];
                            if (e_rank[r] == K - 1) { d2k[s] = __longlong_as_double((long long)e_k[r]); kid[s] = INT_MAX; }
                        }
                    }
                }
            }
            if (tie) retry_mask |= 1u << k;
        }
        if (retry_mask) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(scal + SC_KNN_RETRY, (unsigned long long)__popc(retry_mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lane < KQ_T && ((retry_mask >> lane) & 1u))
                retry_list[base + __popc(retry_mask & lt)] = (int)(s0 + lane);
        }
        __syncwarp();
    }
}

// Search-radius hints for the FIRST evaluation of a handle (no previous smoothing lengths yet): climb from the particle's
// leaf to the smallest cell that holds >= 2 K particles, take the mean density of that cell and size the ball for
// ~1.3 K particles.  Only a hint: the 4-target search hands targets whose ball held < K or > 96 particles to the
// warp-per-target search, which falls back to the guaranteed radius - exactness never depends on it.  Replaces a first
// evaluation at the guaranteed radii (36.7 ms at N = 1e6) by the steady-state search path.
__global__ void __launch_bounds__(256) tree_hint_kernel(int64_t N, int K, const int *__restrict__ perm, SphTree t,
                                                         const unsigned long long *__restrict__ scal, double fac, double hint_k,
                                                         double *__restrict__ hint_h) {
    if (scal[SC_ERR] != 0ull) return;
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    int k = t.leaf_of[s];
    while (t.ncount[k] < 2 * K && t.parent[k] >= 0) k = t.parent[k];
    const double L = t.nodeC[k].w;                              // half-width of the cell
    const double dens = (double)t.ncount[k] / (8.0 * L * L * L);
    const double r = cbrt(3.0 * (hint_k * K) / (4.0 * 3.141592653589793 * dens));
    hint_h[perm[s]] = r / (2.0 * fac);                          // the search multiplies 2 h by fac
}

inline int knn_blocks(int64_t n) {
    int64_t blocks = (n + KNN_WARPS - 1) / KNN_WARPS;
    const int64_t cap = 148 * 8 * 4;
    return (int)(blocks > cap ? cap : blocks);
}

}  // namespace

cudaError_t sph_launch_knn(sph_handle *h, int64_t t0, int64_t t1) {
    if (t1 <= t0) return cudaSuccess;
    // radius hint: h of the previous evaluation (caller's particle order), valid once one evaluation completed
    const double *hint = (h->hint_valid && !h->no_hint) ? h->o_h : nullptr;
    double fac2 = 1.1 * 1.1;
    const double quad_fac = 1.06;      // trial radius = 1.06 x 2 h_prev (measured optimum of 1.03 .. 1.15 at N = 1e6)
    if (!hint && !h->no_hint && h->K <= 64 && h->N >= 4 * h->K) {
        // first evaluation: hints from the tree instead of the previous smoothing lengths
        sph_note(1);
        const double hint_k = 1.3;         // ball sized for 1.3 K particles: 8.7 % of the targets are handed over at N = 1e6 (1.5: 10 %, 1.7: 23 %)
        tree_hint_kernel<<<(int)((h->N + 255) / 256), 256, 0, h->stream>>>(h->N, h->K, h->perm, h->tree, h->scal, quad_fac, hint_k, h->o_h);
        hint = h->o_h;
    }
    if (hint && h->K <= 64) {
        // 4 targets per warp for the hinted targets, then the warp-per-target search for whatever it queued
        sph_note(2);
        fac2 = quad_fac * quad_fac;
        const int64_t quads = (t1 - t0 + KQ_T - 1) / KQ_T;
        int64_t blocks = (quads + KQ_WARPS - 1) / KQ_WARPS;
        if (blocks > 148 * 5 * 8) blocks = 148 * 5 * 8;
        // SPH_B200_KNN_SORT=1: always order the hits with the sort network instead of selecting the K-th distance
        static const int sel_steps = getenv("SPH_B200_KNN_SORT") ? 0 : KQ_SEL_STEPS;
        const int bucket = 32;             // cells up to this many particles are scanned as ranges (<= KQ_BUCKET)
        knn_quad_kernel<KQ_BLOCKS><<<(int)blocks, KQ_WARPS * 32, 0, h->stream>>>(h->N, h->NL, h->K, t0, t1, h->pos4, h->perm, h->tree, hint, fac2,
                                                                                sel_steps, bucket, h->scal, h->cnt, h->nbr, h->d2k, h->kid);
        // queued targets keep their own hinted ball (only the shared box was too wide); a failing hint falls back to
        // the guaranteed radius inside the kernel
        knn_kernel<128, true><<<148 * 5, KNN_WARPS * 32, 0, h->stream>>>(
            h->N, h->NL, h->K, t0, t1, h->pos4, nullptr, 0, h->perm, h->tree, hint, 1.1 * 1.1, h->cnt, h->scal, h->nbr, h->d2k, h->kid, nullptr);
        return cudaGetLastError();
    }
    sph_note(1);
    const int blocks = knn_blocks(t1 - t0);
    if (h->K <= 96)
        knn_kernel<128, true><<<blocks, KNN_WARPS * 32, 0, h->stream>>>(
            h->N, h->NL, h->K, t0, t1, h->pos4, nullptr, 0, h->perm, h->tree, hint, fac2, nullptr, h->scal, h->nbr, h->d2k, h->kid, nullptr);
    else
        knn_kernel<256, true><<<blocks, KNN_WARPS * 32, 0, h->stream>>>(
            h->N, h->NL, h->K, t0, t1, h->pos4, nullptr, 0, h->perm, h->tree, hint, fac2, nullptr, h->scal, h->nbr, h->d2k, h->kid, nullptr);
    return cudaGetLastError();
}

cudaError_t sph_launch_export_neighbors(sph_handle *h, int *idx_out_dev, double *r_out_dev) {
    sph_note(1);
    if (h->K <= 64)
        export_sorted_kernel<64><<<148 * 8, KNN_WARPS * 32, 0, h->stream>>>(h->N, h->NL, h->K, h->pos4, h->perm, h->nbr, idx_out_dev, r_out_dev);
    else if (h->K <= 128)
        export_sorted_kernel<128><<<148 * 8, KNN_WARPS * 32, 0, h->stream>>>(h->N, h->NL, h->K, h->pos4, h->perm, h->nbr, idx_out_dev, r_out_dev);
    else
        export_sorted_kernel<256><<<148 * 8, KNN_WARPS * 32, 0, h->stream>>>(h->N, h->NL, h->K, h->pos4, h->perm, h->nbr, idx_out_dev, r_out_dev);
    return cudaGetLastError();
}

// pts_dev: M x 3 column-major device points; d2s: M x K scratch for the sorted squared distances
cudaError_t sph_launch_knn_points(sph_handle *h, const double *pts_dev, int64_t M, double *d2s, double *rho_out_dev) {
    if (M <= 0) return cudaSuccess;
    sph_note(2);
    const int blocks = knn_blocks(M);
    if (h->K <= 96)
        knn_kernel<128, false><<<blocks, KNN_WARPS * 32, 0, h->stream>>>(
            h->N, h->NL, h->K, 0, M, h->pos4, pts_dev, M, h->perm, h->tree, nullptr, 1.0, nullptr, h->scal, nullptr, nullptr, nullptr, d2s);
    else
        knn_kernel<256, false><<<blocks, KNN_WARPS * 32, 0, h->stream>>>(
            h->N, h->NL, h->K, 0, M, h->pos4, pts_dev, M, h->perm, h->tree, nullptr, 1.0, nullptr, h->scal, nullptr, nullptr, nullptr, d2s);
    point_density_kernel<<<(int)((M + 127) / 128), 128, 0, h->stream>>>(M, h->K, d2s, h->p.m,
                                                                         h->p.eos == SPH_EOS_POLYTROPIC, rho_out_dev);
    return cudaGetLastError();
}
