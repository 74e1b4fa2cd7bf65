// knn.cu -- exact K-nearest-neighbour search on the key-sorted particles and their octree.
//
// Replaces NearestNeighbors.KDTree + knn(tree, ri', K, true) as called by HJL.getNeighbors
// (F/isothermal_hydroKDTree.jl:128-142): for every particle the K nearest particles INCLUDING itself,
// ascending distance.  Distances are evaluated exactly like the oracle's restatement
// (d2 = (dx*dx + dy*dy) + dz*dz, no FMA), ties are ordered by the caller's particle index, so the
// emitted lists are bit-identical to the CPU path.
//
// One warp per target, ball collection + one sort:
//   1. guaranteed radius R0: the K particles around the target in key order all lie within
//      max_j d2(target, j), so the ball of that radius holds at least K particles;
//   2. trial radius: (2 h_prev * 1.1)^2 from the previous force evaluation of the same particle when that
//      is smaller than R0 (positions move by <= v dt/2 between evaluations, F/isothermal_sim.jl:197);
//   3. depth-first walk of the octree: up to 8 children of a cell are tested by 8 lanes (point-to-box
//      distance with an absolute slack covering the <= 1 ulp mismatch between the reference's
//      classification centre and its stored bounds); cells holding <= 32 particles are scanned as
//      contiguous ranges of the sorted array (coalesced), hits are appended to a shared-memory buffer
//      with a ballot prefix.  A full buffer is compacted by a warp bitonic sort that keeps the K best and
//      tightens the radius (only taken on a cold start, when the radius is the loose R0);
//   4. if the trial ball held fewer than K particles the search is repeated with R0 (exactness never
//      depends on the hint); finally one bitonic sort by (d2, particle id) and the first K are emitted.
#include "sph_internal.cuh"

#include <climits>

namespace {

constexpr int KNN_WARPS = 8;
constexpr int KNN_BUCKET = 32;
constexpr int KNN_STACK = 192;

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
__global__ void __launch_bounds__(KNN_WARPS * 32, 5) knn_kernel(int64_t N, int K, int64_t t0, int64_t t1,
                                                              const double4 *__restrict__ pos4,
                                                              const double *__restrict__ qpts, int64_t qstride,
                                                              const int *__restrict__ perm, SphTree t,
                                                              const double *__restrict__ hint_h, double hint_fac2,
                                                              unsigned long long *__restrict__ scal,
                                                              int *__restrict__ nbr, double *__restrict__ d2k,
                                                              double *__restrict__ d2_out) {
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

    const int64_t nwarps = (int64_t)gridDim.x * KNN_WARPS;
    for (int64_t s = t0 + (int64_t)blockIdx.x * KNN_WARPS + warp; s < t1; s += nwarps) {
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
                    double ax = fmax(fmax(B.x - qx, qx - B.w), 0.0);
                    double ay = fmax(fmax(B.y - qy, qy - C.x), 0.0);
                    double az = fmax(fmax(B.z - qz, qz - C.y), 0.0);
                    ax = fmax(ax - eps, 0.0); ay = fmax(ay - eps, 0.0); az = fmax(az - eps, 0.0);
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
            for (int j = lane; j < K; j += 32) nbr[s + (int64_t)j * N] = bid[j];
            if (lane == 0) d2k[s] = bd2[K - 1];
        } else {
            for (int j = lane; j < K; j += 32) d2_out[s + (int64_t)j * qstride] = bd2[j];
        }
        __syncwarp();
    }
    if (retries && lane == 0) atomicAdd(scal + SC_KNN_RETRY, retries);
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

inline int knn_blocks(int64_t n) {
    int64_t blocks = (n + KNN_WARPS - 1) / KNN_WARPS;
    const int64_t cap = 148 * 8 * 4;
    return (int)(blocks > cap ? cap : blocks);
}

}  // namespace

cudaError_t sph_launch_knn(sph_handle *h, int64_t t0, int64_t t1) {
    if (t1 <= t0) return cudaSuccess;
    sph_note(1);
    const int blocks = knn_blocks(t1 - t0);
    // radius hint: h of the previous evaluation (caller's particle order), valid once one evaluation completed
    const double *hint = (h->hint_valid && !h->no_hint) ? h->o_h : nullptr;
    const double fac2 = 1.1 * 1.1;
    if (h->K <= 96)
        knn_kernel<128, true><<<blocks, KNN_WARPS * 32, 0, h->stream>>>(
            h->N, h->K, t0, t1, h->pos4, nullptr, 0, h->perm, h->tree, hint, fac2, h->scal, h->nbr, h->d2k, nullptr);
    else
        knn_kernel<256, true><<<blocks, KNN_WARPS * 32, 0, h->stream>>>(
            h->N, h->K, t0, t1, h->pos4, nullptr, 0, h->perm, h->tree, hint, fac2, h->scal, h->nbr, h->d2k, nullptr);
    return cudaGetLastError();
}

// pts_dev: M x 3 column-major device points
cudaError_t sph_launch_knn_points(sph_handle *h, const double *pts_dev, int64_t M, double *rho_out_dev) {
    if (M <= 0) return cudaSuccess;
    sph_note(2);
    double *d2s = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d2s, (size_t)M * h->K * sizeof(double), h->stream);
    if (e != cudaSuccess) return e;
    const int blocks = knn_blocks(M);
    if (h->K <= 96)
        knn_kernel<128, false><<<blocks, KNN_WARPS * 32, 0, h->stream>>>(
            h->N, h->K, 0, M, h->pos4, pts_dev, M, h->perm, h->tree, nullptr, 1.0, h->scal, nullptr, nullptr, d2s);
    else
        knn_kernel<256, false><<<blocks, KNN_WARPS * 32, 0, h->stream>>>(
            h->N, h->K, 0, M, h->pos4, pts_dev, M, h->perm, h->tree, nullptr, 1.0, h->scal, nullptr, nullptr, d2s);
    point_density_kernel<<<(int)((M + 127) / 128), 128, 0, h->stream>>>(M, h->K, d2s, h->p.m,
                                                                         h->p.eos == SPH_EOS_POLYTROPIC, rho_out_dev);
    e = cudaGetLastError();
    cudaFreeAsync(d2s, h->stream);
    return e;
}
