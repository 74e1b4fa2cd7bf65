// hydro.cu -- density + EOS and pressure / artificial-viscosity force over the neighbour lists, in GATHER form:
// no atomics on floating-point data, every sum in a fixed order, so results are bit-reproducible from run to run.
//
// Replaces HJL.W / getDensity / getPressure / gradW / getAV / hydroCalculation and the scatter part of
// evolve_K! (F/isothermal_hydroKDTree.jl:5-245, F/polytrope_hydroKDTree.jl:5-341).  The reference
// materialises ~35 N x K Float64 matrices per call; here every pair quantity lives in registers and only
// the per-particle reductions its caller consumes are written:
//   rho, h, P/rho^2, c_i                                       (density + EOS; pos4.w carries d2k = (2h)^2 throughout)
//   a_hyd, sum_j v_ij.gradW_ij, max_j mu_ij, dK/dt sum         (force)
//
// The reference's pair loop is a scatter (F/isothermal_hydroKDTree.jl:226-242): for j in N(i), j != i:
//       a_i -= ct_ij gradW_ij ;  a_j += ct_ij gradW_ij          (gradW_ij uses h_i)
// Here particle i GATHERS both kinds of terms that land on it:
//   forward   sum over its own list j in N(i)                    -ct_ij g(r, h_i) d_ij
//   reverse   sum over the particles k whose lists contain i     -ct_ki g(r, h_k) d_ik          (d_ik = x_i - x_k)
// A reverse partner k is either in i's own list (mutual pair, 93 % of the pairs at N = 1e6: both terms share
// d, r, mu, rho_bar) or it is not; those few (3.7 per particle on average) are collected per particle in the "extras"
// table E(i) by the density pass.  List membership "a in N(b)" is decided from b's K-th squared distance d2k_b and the
// tie id kid_b (knn.cu) alone:  d2_ab < d2k_b, or d2_ab == d2k_b and id_a <= kid_b -- exactly the set the search
// emitted, ties included, without reading b's list (which another rank may own).
//
// Lists are K x NL int32 (NL = N rounded up to 128; column-major like the Julia matrix), so the lanes of a warp -
// consecutive targets in key order - read consecutive list entries (coalesced) and gather spatially close particles
// (L1/L2 hits).  The *_tile kernels stage the list tile of a block and a key-order window of the particle
// records in shared memory with bulk async copies (TMA, cp.async.bulk + mbarrier).
#include "sph_internal.cuh"

#include <cstdlib>

namespace {

constexpr int HB = 128;
// build-time experiments (tools/build_variants.sh): L1 prefetch distance of the gathers (0 = off), resident blocks of the force pass
#ifndef SPH_PREFETCH
#define SPH_PREFETCH 0
#endif
#ifndef FORCE_MINB
#define FORCE_MINB 4
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
constexpr double PI_D = 3.141592653589793;
constexpr double INV_PI_D = 0.3183098861837907;

// W  (F/isothermal_hydroKDTree.jl:22-31; polytropic second mask = !mask1, F/polytrope_hydroKDTree.jl:158)
__device__ __forceinline__ double kernel_W(double ct, double q, bool poly) {
    if (q <= 1.0) return ct * ((1 - 3.0 / 2 * (q * q)) + 3.0 / 4 * (q * q * q));
    if (poly || q <= 2.0) {
        const double u = 2 - q;
        return (ct * 1 / 4) * (u * u * u);
    }
    return 0.0;
}
// (dW/dr)/r  (F/isothermal_hydroKDTree.jl:57-70)
// hinv = 1/h, rinv = 1/r (approximations to 1e-13: the parity bar for the hydro force is 1e-9)
__device__ __forceinline__ double kernel_dWdr(double ct4, double hinv, double q, double rinv, bool poly) {
    if (q <= 1.0) return ct4 * hinv * (9.0 / 4 * q - 3);          // 9/4 r/h^2 - 3/h
    if (poly || q <= 2.0) {
        const double u = 2 - q;
        return ct4 * (-3.0 / 4 * (u * u)) * rinv;
    }
    return 0.0;
}

// "a in N(b)": a at squared distance d2 from b, b's K-th squared distance d2k_b (see the file header)
__device__ __forceinline__ bool in_list_of(double d2, double d2k_b, int a, int b, const int *__restrict__ perm,
                                           const int *__restrict__ kid) {
    if (d2 < d2k_b) return true;
    if (d2 > d2k_b) return false;
    return perm[a] <= kid[b];          // at the K-th distance: the search kept the smaller particle ids
}

// where the density pass records "k is a reverse partner of j that j will not find in its own list"
struct ExtrasOut {
    int *ecnt;              // [NL] entries per particle (may exceed ecap: the surplus is in the overflow list)
    int *ext;               // [SPH_ECAP][NL] column-major
    int2 *ovf;              // overflow pairs {j, k}
    int2 *outbox;           // pairs for particles of other ranks; entry 0 = {count, 0}
    int ecap, ovcap, obcap;
    int64_t own0, own1;     // sorted slots this rank owns
    int64_t NL;
};

__device__ __forceinline__ void push_extra(const ExtrasOut &x, int j, int k, unsigned long long *__restrict__ scal) {
    if (j >= x.own0 && j < x.own1) {
        const int slot = atomicAdd(&x.ecnt[j], 1);
        if (slot < x.ecap) {
            x.ext[(int64_t)slot * x.NL + j] = k;
        } else {
            const unsigned long long o = atomicAdd(scal + SC_OVF, 1ull);
            if (o < (unsigned long long)x.ovcap) x.ovf[o] = make_int2(j, k);
            else atomicOr(scal + SC_ERR, (unsigned long long)ERRF_EXTRAS);
        }
    } else {
        const unsigned long long o = atomicAdd(scal + SC_OUTBOX, 1ull);
        if (o < (unsigned long long)x.obcap) x.outbox[1 + o] = make_int2(j, k);
        else atomicOr(scal + SC_ERR, (unsigned long long)ERRF_HALO);
    }
}

// ---------------------------------------------------------------------------------------------------
// bulk async copies global -> shared (TMA, non-tensor form) completing on an mbarrier
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// bytes: multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------------
// density + extras pass.  pos4.w holds d2k (the K-th squared distance of every particle).
// TILE: the block's K x 128 list tile and the records of the key-order window [tile - TW, tile + 128 + TW) are staged
// in shared memory by bulk async copies; neighbours outside the window are gathered from global memory.
// ---------------------------------------------------------------------------------------------------
constexpr int TW = 128;                 // window margin on each side of the 128-target tile
constexpr int TWIN = HB + 2 * TW;       // records in the window

// EOS closure of one particle: isothermal P = cs^2 rho (F/isothermal_hydroKDTree.jl:190), c = cs; polytropic
// P = K rho^gamma (F/polytrope_hydroKDTree.jl:216), c_i = sqrt(gamma K rho^(gamma-1)) (:186).
// Writes hr = {h, rho} and the records the force pass gathers per neighbour: fa = {x, y, z, h}, fb = {vx, vy, vz, rho},
// fc = {P/rho^2, c}.  The isothermal force needs only fa and fb (P/rho^2 = cs^2 / rho, c = cs): two 32-byte sectors per
// neighbour instead of three.
// (Measured and dropped: ONE 128-byte record per particle, i.e. one cache line per gathered neighbour instead of
// sectors in three lines - force 0.73 -> 1.08 ms at N = 1e6: a line of a 32-byte array holds four key-adjacent
// particles, which are mostly neighbours too, so separate arrays use L1 far better.)
struct ForceRecs {
    double4 *fa, *fb;
    double2 *fc;
};
__device__ __forceinline__ void eos_store(int64_t s, const double4 &pi, const double4 &vi, double h, double rho, int poly,
                                          double cs, double gamma, double2 *__restrict__ hr, const ForceRecs &f) {
    double P, c;
    if (!poly) {
        P = cs * cs * rho;
        c = cs;
    } else {
        const double Kent = vi.w;
        c = sqrt(gamma * Kent * pow(rho, gamma - 1));
        P = Kent * pow(rho, gamma);
    }
    hr[s] = make_double2(h, rho);
    f.fa[s] = make_double4(pi.x, pi.y, pi.z, h);
    f.fb[s] = make_double4(vi.x, vi.y, vi.z, rho);
    f.fc[s] = make_double2(P / (rho * rho), c);
}

// EOS: with one rank the density pass closes the EOS of its target on the spot (every particle is a target);
// with several ranks rho is all-gathered first and eos_kernel runs over all particles.
template <bool TILE, bool EOS>
__global__ void __launch_bounds__(HB) density_kernel(int64_t N, int64_t NL, int K, int64_t t0, int64_t t1,
                                                      const double4 *__restrict__ pos4, const int *__restrict__ nbr,
                                                      const int *__restrict__ perm, const int *__restrict__ kid,
                                                      double m, int poly, unsigned long long *__restrict__ scal,
                                                      ExtrasOut x, double *__restrict__ rho_out,
                                                      const double4 *__restrict__ vel4, double cs, double gamma,
                                                      double2 *__restrict__ hr, ForceRecs frec) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    if (SPH_ERR_BLOCKING(scal) != 0ull) return;
    const int64_t s0 = t0 + (int64_t)blockIdx.x * HB;
    const int64_t s = s0 + threadIdx.x;
    // ---- staging
    const int *lst = nbr + s;           // entry j of this target at lst[j * NL]
    int64_t lstride = NL;
    const double4 *wpos = nullptr;
    int64_t w0 = 0;
    int wn = 0;
    if (TILE) {
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(dyn_smem);
        double4 *s_pos = reinterpret_cast<double4 *>(dyn_smem + 128);
        int *s_idx = reinterpret_cast<int *>(dyn_smem + 128 + sizeof(double4) * TWIN);
        w0 = s0 - TW < 0 ? 0 : s0 - TW;
        const int64_t w1 = s0 + HB + TW > N ? N : s0 + HB + TW;
        wn = (int)(w1 - w0);
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (unsigned)(wn * sizeof(double4) + (size_t)K * HB * sizeof(int)));
            bulk_g2s(s_pos, pos4 + w0, (unsigned)(wn * sizeof(double4)), bar);
            for (int j = 0; j < K; ++j) bulk_g2s(s_idx + j * HB, nbr + (int64_t)j * NL + s0, HB * sizeof(int), bar);
        }
        mbar_wait(bar, 0);
        lst = s_idx + threadIdx.x;
        lstride = HB;
        wpos = s_pos;
    }
    if (s >= t1) return;
    const double4 pi = TILE ? wpos[s - w0] : pos4[s];
    const double h = sqrt(pi.w) / 2;                 // h = r[:, end] ./ 2   (:151); pi.w = d2k
    const double ct = 1 / (PI_D * (h * h * h));
    const double hinv = 1 / h;
    double sum = 0.0;
    unsigned long long defer = 0ull;     // bit j: the j-th neighbour's list does not hold s (told after the loop)
    for (int j = 0; j < K; ++j) {
        const int nj = lst[j * lstride];
        if (SPH_PREFETCH > 0 && !TILE && j + SPH_PREFETCH < K) prefetch_l1(pos4 + lst[(j + SPH_PREFETCH) * lstride]);
        double4 pj;
        if (TILE) {
            const unsigned loc = (unsigned)(nj - (int)w0);
            const double4 *pp = loc < (unsigned)wn ? wpos + loc : pos4 + nj;     // generic load: window or global
            pj = *pp;
        } else {
            pj = pos4[nj];
        }
        // q = r / h from one rsqrt seed + Newton step (1e-13 relative; the parity bar for rho is 1e-9)
        const double d2 = sph_d2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        const double q = d2 > 0.0 ? d2 * fast_rsqrt(d2) * hinv : 0.0;
        sum += kernel_W(ct, q, poly != 0);           // rho_i = m * sum_j w_ij  (:175)
        // s is in nj's list <=> nj finds s itself when it gathers; otherwise nj must be told (its extras table).
        // The atomics that hand out the table slots are issued after the loop: inside it nearly every warp iteration
        // would wait for the round trip of the one or two lanes that have something to tell.
        if (!(d2 < pj.w) && nj != (int)s) {
            if (!in_list_of(d2, pj.w, (int)s, nj, perm, kid)) {
                if (j < 64) defer |= 1ull << j;
                else push_extra(x, nj, (int)s, scal);
            }
        }
    }
    // first all the slot requests of up to 8 entries (independent atomics: one round trip instead of one per entry),
    // then the stores; whatever is left takes the plain path
    {
        int tj[8], slot[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            tj[u] = -1; slot[u] = 0;
            if (defer) {
                const int j = __ffsll((long long)defer) - 1;
                defer &= defer - 1ull;
                const int nj = lst[j * lstride];
                if (nj >= x.own0 && nj < x.own1) { tj[u] = nj; slot[u] = atomicAdd(&x.ecnt[nj], 1); }
                else push_extra(x, nj, (int)s, scal);             // another rank's particle: exchange buffer
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (tj[u] >= 0) {
                if (slot[u] < x.ecap) {
                    x.ext[(int64_t)slot[u] * x.NL + tj[u]] = (int)s;
                } else {
                    const unsigned long long o = atomicAdd(scal + SC_OVF, 1ull);
                    if (o < (unsigned long long)x.ovcap) x.ovf[o] = make_int2(tj[u], (int)s);
                    else atomicOr(scal + SC_ERR, (unsigned long long)ERRF_EXTRAS);
                }
            }
        }
    }
    while (defer) {
        const int j = __ffsll((long long)defer) - 1;
        defer &= defer - 1ull;
        push_extra(x, lst[j * lstride], (int)s, scal);
    }
    if (EOS) eos_store(s, pi, vel4[s], h, m * sum, poly, cs, gamma, hr, frec);
    else rho_out[s] = m * sum;
}

// extras of other ranks' targets that point into this rank's range: append them (order is fixed afterwards)
__global__ void __launch_bounds__(256) extras_merge_kernel(const int2 *__restrict__ inbox, int nranks, int rank, int64_t stride,
                                                            ExtrasOut x, unsigned long long *__restrict__ scal) {
    if (SPH_ERR_BLOCKING(scal) != 0ull) return;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) continue;
        const int2 *box = inbox + (int64_t)r * stride;
        const int n = box[0].x;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
            const int2 p = box[1 + e];
            if (p.x >= x.own0 && p.x < x.own1) push_extra(x, p.x, p.y, scal);
        }
    }
}

__global__ void outbox_header_kernel(int2 *__restrict__ outbox, int obcap, const unsigned long long *__restrict__ scal) {
    const unsigned long long n = scal[SC_OUTBOX];
    outbox[0] = make_int2((int)(n < (unsigned long long)obcap ? n : (unsigned long long)obcap), 0);
}

// fixed order of every particle's extras (the slots were handed out by atomics): ascending sorted-space index
__global__ void __launch_bounds__(HB) extras_sort_kernel(int64_t NL, int64_t t0, int64_t t1, int ecap,
                                                          const int *__restrict__ ecnt, int *__restrict__ ext,
                                                          const unsigned long long *__restrict__ scal) {
    if (SPH_ERR_BLOCKING(scal) != 0ull) return;
    const int64_t s = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t1) return;
    int n = ecnt[s];
    n = n < ecap ? n : ecap;
    if (n < 2) return;
    if (n <= 8) {
        // up to 8 entries (nearly every particle): all loads at once, then each value goes to the slot of its rank
        int v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = e < n ? ext[(int64_t)e * NL + s] : 0x7fffffff;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int r = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) r += v[b] < v[e];
            if (e < n) ext[(int64_t)r * NL + s] = v[e];
        }
        return;
    }
    for (int a = 0; a + 1 < n; ++a) {              // selection sort on the column of this particle
        int best = ext[(int64_t)a * NL + s], bi = a;
        for (int b = a + 1; b < n; ++b) {
            const int v = ext[(int64_t)b * NL + s];
            if (v < best) { best = v; bi = b; }
        }
        if (bi != a) {
            ext[(int64_t)bi * NL + s] = ext[(int64_t)a * NL + s];
            ext[(int64_t)a * NL + s] = best;
        }
    }
}

// EOS closure for ALL particles after the density all-gather of multi-GPU runs (pos4.w = d2k)
__global__ void __launch_bounds__(HB) eos_kernel(int64_t N, const double *__restrict__ rho_s, const double4 *__restrict__ pos4,
                                                  const double4 *__restrict__ vel4, int poly, double cs, double gamma,
                                                  const unsigned long long *__restrict__ scal, double2 *__restrict__ hr,
                                                  ForceRecs frec) {
    if (SPH_ERR_BLOCKING(scal) != 0ull) return;
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const double4 pi = pos4[s];
    const double h = sqrt(pi.w) / 2;                 // h = r[:, end] ./ 2   (:151)
    eos_store(s, pi, vel4[s], h, rho_s[s], poly, cs, gamma, hr, frec);
}

// pos4.w = d2k for ALL particles (after the d2k all-gather in multi-GPU runs): the density and force passes read the
// K-th distance of every neighbour next to its position; also clears the extras counters
__global__ void __launch_bounds__(HB) smoothing_kernel(int64_t N, const double *__restrict__ d2k,
                                                        const unsigned long long *__restrict__ scal,
                                                        double4 *__restrict__ pos4, double *__restrict__ hs, int *__restrict__ ecnt) {
    if (SPH_ERR_BLOCKING(scal) != 0ull) return;
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const double v = d2k[s];
    pos4[s].w = v;
    hs[s] = sqrt(v) / 2;                             // h = r[:, end] ./ 2   (:151): the tree walk starts from it
    ecnt[s] = 0;
}

// ---------------------------------------------------------------------------------------------------
// One pair (i, k) of the force pass.  Quantities shared by both directions: d = x_i - x_k, r, v.d, h_bar, rho_bar, mu
// (getAV, F/isothermal_hydroKDTree.jl:200-213).  FWD: the term of i's own list (kernel of i, sound speed of i);
// rev: the reaction of k's list on i (kernel of k, sound speed of k).  Both are -ct g d on a_i.
// ---------------------------------------------------------------------------------------------------
struct Target {
    double x, y, z, h, vx, vy, vz, rho, prr, cs, hinv, ct4;
};

// A = {x, y, z, h} and B = {vx, vy, vz, rho} of the partner, prr_j = P_j / rho_j^2, cs_j its sound speed
template <bool POLY, bool FWD>
__device__ __forceinline__ void pair_terms(const Target &t, const double4 &A, const double4 &B, double prr_j, double cs_j, bool rev,
                                           double m, double alpha, double beta, double &ax, double &ay, double &az,
                                           double &dk, double &svdw, double &mmax) {
    const double dx = t.x - A.x, dy = t.y - A.y, dz = t.z - A.z;              // getTreeDiffs: f_i - f_j (:93)
    const double d2 = sph_d2_exact(dx, dy, dz);
    const double rinv = d2 > 0.0 ? fast_rsqrt(d2) : 0.0;
    const double r = d2 * rinv;
    const double h_avg = (t.h + A.w) / 2;                                    // getVectorTreeAvgs (:111)
    const double rho_avg = (t.rho + B.w) / 2;
    const double vx = t.vx - B.x, vy = t.vy - B.y, vz = t.vz - B.z;
    const double vdr = (vx * dx + vy * dy) + vz * dz;                        // (:210)
    const double mu = fmin(h_avg * vdr * fast_rcp(d2 + 0.01 * (h_avg * h_avg)), 0.0);   // (:211)
    const double rinv_rho = fast_rcp(rho_avg);
    const double bmu2 = beta * (mu * mu);
    double coef = 0.0;
    if (FWD) {
        const double q = r * t.hinv;
        const double dW = kernel_dWdr(t.ct4, t.hinv, q, rinv, POLY);
        const double Pi = ((-alpha) * t.cs * mu + bmu2) * rinv_rho;          // (:213), c = c_i
        const double vdw = dW * vdr;                                         // v_ij . gradW_ij
        svdw += vdw;
        mmax = fmax(mmax, mu);
        double ct;
        if (!POLY) ct = m * (t.prr + Pi / 2);                                // iso :232
        else ct = m * ((t.prr + prr_j) + Pi) / 2;                            // poly :235
        coef = ct * dW;
        if (POLY) dk += m * Pi * vdw / 2;                                    // evolve_K! poly :305-311
    }
    if (rev) {
        const double hinv_j = fast_rcp(A.w);
        const double hi2 = hinv_j * hinv_j;
        const double ct4_j = INV_PI_D * (hi2 * hi2);                         // 1 / (pi h_j^4)
        const double q = r * hinv_j;
        const double dW = kernel_dWdr(ct4_j, hinv_j, q, rinv, POLY);
        const double Pi = ((-alpha) * (POLY ? cs_j : t.cs) * mu + bmu2) * rinv_rho;     // c = c_j (the list owner's)
        double ct;
        if (!POLY) ct = m * (prr_j + Pi / 2);
        else ct = m * ((prr_j + t.prr) + Pi) / 2;
        coef += ct * dW;
        if (POLY) dk += m * Pi * (dW * vdr) / 2;
    }
    ax -= coef * dx; ay -= coef * dy; az -= coef * dz;
}

// P/rho^2 and c of a partner: polytropic - the fc record; isothermal - cs^2 / rho_j and cs (no third gather)
template <bool POLY>
__device__ __forceinline__ void partner_eos(const double2 *__restrict__ fc, int64_t j, double rho_j, double cs2, double cs,
                                            double &prr_j, double &cs_j) {
    if (POLY) {
        const double2 c = fc[j];
        prr_j = c.x; cs_j = c.y;
    } else {
        prr_j = cs2 * fast_rcp(rho_j);
        cs_j = cs;
    }
}

// "s in N(j)" from j's smoothing length when the distance is clearly inside / outside (2 h_j)^2 = d2k_j, else exactly
__device__ __forceinline__ bool in_list_fast(double d2, double h_j, int a, int b, const double *__restrict__ d2k,
                                             const int *__restrict__ perm, const int *__restrict__ kid) {
    const double t4 = 4.0 * (h_j * h_j);
    if (d2 < t4 * (1.0 - 1e-13)) return true;
    if (d2 > t4 * (1.0 + 1e-13)) return false;
    return in_list_of(d2, d2k[b], a, b, perm, kid);
}

template <bool POLY, bool TILE>
__global__ void __launch_bounds__(HB, TILE ? 3 : FORCE_MINB) force_kernel(int64_t N, int64_t NL, int64_t NS, int K, int64_t t0, int64_t t1,
                                                    const double4 *__restrict__ fa, const double4 *__restrict__ fb,
                                                    const double2 *__restrict__ fc, const double *__restrict__ d2k,
                                                    const int *__restrict__ nbr,
                                                    const int *__restrict__ perm, const int *__restrict__ kid,
                                                    const int *__restrict__ ecnt, const int *__restrict__ ext, int ecap,
                                                    double m, double alpha, double beta, double cs2,
                                                    const unsigned long long *__restrict__ scal,
                                                    double *__restrict__ ahyd, double *__restrict__ dkdt,
                                                    double *__restrict__ sumvdw, double *__restrict__ mumax) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    if (SPH_ERR_BLOCKING(scal) != 0ull) return;
    const int64_t s0 = t0 + (int64_t)blockIdx.x * HB;
    const int64_t s = s0 + threadIdx.x;
    const int *lst = nbr + s;
    int64_t lstride = NL;
    const double4 *wa = nullptr, *wb = nullptr;
    const double2 *wc = nullptr;
    int64_t w0 = 0;
    int wn = 0;
    if (TILE) {
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(dyn_smem);
        double4 *s_a = reinterpret_cast<double4 *>(dyn_smem + 128);
        double4 *s_b = s_a + TWIN;
        double2 *s_c = reinterpret_cast<double2 *>(s_b + TWIN);
        int *s_idx = reinterpret_cast<int *>(s_c + TWIN);
        w0 = s0 - TW < 0 ? 0 : s0 - TW;
        const int64_t w1 = s0 + HB + TW > N ? N : s0 + HB + TW;
        wn = (int)(w1 - w0);
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned wb4 = (unsigned)(wn * sizeof(double4)), wb2 = POLY ? (unsigned)(wn * sizeof(double2)) : 0u;
            mbar_expect_tx(bar, 2 * wb4 + wb2 + (unsigned)((size_t)K * HB * sizeof(int)));
            bulk_g2s(s_a, fa + w0, wb4, bar);
            bulk_g2s(s_b, fb + w0, wb4, bar);
            if (POLY) bulk_g2s(s_c, fc + w0, wb2, bar);
            for (int j = 0; j < K; ++j) bulk_g2s(s_idx + j * HB, nbr + (int64_t)j * NL + s0, HB * sizeof(int), bar);
        }
        mbar_wait(bar, 0);
        lst = s_idx + threadIdx.x;
        lstride = HB;
        wa = s_a; wb = s_b; wc = s_c;
    }
    if (s >= t1) return;
    Target t;
    {
        const double4 pi = fa[s], vi = fb[s];
        const double2 ci = fc[s];
        t.x = pi.x; t.y = pi.y; t.z = pi.z; t.h = pi.w;
        t.vx = vi.x; t.vy = vi.y; t.vz = vi.z; t.rho = vi.w;
        t.prr = ci.x; t.cs = ci.y;
        const double h2 = t.h * t.h;
        t.ct4 = 1 / (PI_D * (h2 * h2));
        t.hinv = 1 / t.h;
    }
    double ax = 0.0, ay = 0.0, az = 0.0, svdw = 0.0, dk = 0.0;
    // the self pair (column 1 of the reference's matrices) has v_ii = 0: v.gradW = 0 and mu = min(0, 0) = 0 enter the
    // row sum and the row maximum; hydroCalculation and evolve_K! start at column 2 (:226, poly :301)
    double mmax = 0.0;
    for (int j = 0; j < K; ++j) {
        const int nj = lst[j * lstride];
        if (SPH_PREFETCH > 0 && !TILE && j + SPH_PREFETCH < K) {
            const int np = lst[(j + SPH_PREFETCH) * lstride];
            prefetch_l1(fa + np); prefetch_l1(fb + np);
        }
        if (nj == (int)s) continue;      // lists arrive unordered from the grouped search: self is recognised by index
        const double4 *pa = fa + nj, *pb = fb + nj;
        const double2 *pcj = fc + nj;
        if (TILE) {
            const unsigned loc = (unsigned)(nj - (int)w0);
            if (loc < (unsigned)wn) { pa = wa + loc; pb = wb + loc; pcj = wc + loc; }      // generic loads: window or global
        }
        const double4 A = *pa, B = *pb;
        double prr_j, cs_j;
        partner_eos<POLY>(pcj, 0, B.w, cs2, t.cs, prr_j, cs_j);
        // does nj's list contain s?  then its reaction on s is gathered here (mutual pair)
        const double d2 = sph_d2_exact(t.x - A.x, t.y - A.y, t.z - A.z);
        const bool rev = in_list_fast(d2, A.w, (int)s, nj, d2k, perm, kid);
        pair_terms<POLY, true>(t, A, B, prr_j, cs_j, rev, m, alpha, beta, ax, ay, az, dk, svdw, mmax);
    }
    // reverse partners outside the own list (3.7 on average), ascending index (extras_sort_kernel)
    int ne = ecnt[s];
    ne = ne < ecap ? ne : ecap;
    double dummy_s = 0.0, dummy_m = 0.0;
    for (int e = 0; e < ne; ++e) {
        const int k = ext[(int64_t)e * NL + s];
        const double4 A = fa[k], B = fb[k];
        double prr_j, cs_j;
        partner_eos<POLY>(fc, k, B.w, cs2, t.cs, prr_j, cs_j);
        pair_terms<POLY, false>(t, A, B, prr_j, cs_j, true, m, alpha, beta, ax, ay, az, dk, dummy_s, dummy_m);
    }
    ahyd[s] = ax; ahyd[s + NS] = ay; ahyd[s + 2 * NS] = az;
    if (POLY) dkdt[s] = dk;
    sumvdw[s] = svdw;
    mumax[s] = mmax;
}

// Particles with more reverse partners than the table holds: the surplus sits in the overflow list (unordered).
// Each such particle scans the list for its entries and adds them in ascending index order (rare, deterministic).
template <bool POLY>
__global__ void __launch_bounds__(HB) force_overflow_kernel(int64_t NS, int64_t t0, int64_t t1,
                                                             const double4 *__restrict__ fa, const double4 *__restrict__ fb,
                                                             const double2 *__restrict__ fc, const int *__restrict__ ecnt,
                                                             int ecap, const int2 *__restrict__ ovf, int ovcap, double m,
                                                             double alpha, double beta, double cs2,
                                                             const unsigned long long *__restrict__ scal,
                                                             double *__restrict__ ahyd, double *__restrict__ dkdt) {
    if (SPH_ERR_BLOCKING(scal) != 0ull) return;
    const unsigned long long no = scal[SC_OVF];
    if (no == 0ull) return;
    const int n = (int)(no < (unsigned long long)ovcap ? no : (unsigned long long)ovcap);
    const int64_t s = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t1 || ecnt[s] <= ecap) return;
    Target t;
    {
        const double4 pi = fa[s], vi = fb[s];
        const double2 ci = fc[s];
        t.x = pi.x; t.y = pi.y; t.z = pi.z; t.h = pi.w;
        t.vx = vi.x; t.vy = vi.y; t.vz = vi.z; t.rho = vi.w;
        t.prr = ci.x; t.cs = ci.y;
        t.ct4 = 0.0; t.hinv = 0.0;
    }
    double ax = 0.0, ay = 0.0, az = 0.0, dk = 0.0, d0 = 0.0, d1 = 0.0;
    int last = -1;
    for (;;) {
        int best = 0x7fffffff;
        for (int e = 0; e < n; ++e) {
            const int2 p = ovf[e];
            if (p.x == (int)s && p.y > last && p.y < best) best = p.y;
        }
        if (best == 0x7fffffff) break;
        const double4 A = fa[best], B = fb[best];
        double prr_j, cs_j;
        partner_eos<POLY>(fc, best, B.w, cs2, t.cs, prr_j, cs_j);
        pair_terms<POLY, false>(t, A, B, prr_j, cs_j, true, m, alpha, beta, ax, ay, az, dk, d0, d1);
        last = best;
    }
    ahyd[s] += ax; ahyd[s + NS] += ay; ahyd[s + 2 * NS] += az;
    if (POLY) dkdt[s] += dk;
}

inline ExtrasOut extras_of(sph_handle *h, int64_t t0, int64_t t1) {
    ExtrasOut x;
    x.ecnt = h->ecnt; x.ext = h->ext; x.ovf = h->ovf; x.outbox = h->outbox;
    x.ecap = h->ecap; x.ovcap = (int)h->ovcap; x.obcap = (int)h->obcap;
    x.own0 = t0; x.own1 = t1; x.NL = h->NL;
    return x;
}

constexpr size_t DENS_SMEM = 128 + sizeof(double4) * TWIN;          // + K * HB * 4
constexpr size_t FORCE_SMEM = 128 + (2 * sizeof(double4) + sizeof(double2)) * TWIN;     // + K * HB * 4

}  // namespace

// SPH_B200_SPH_TILE=0 selects the direct-gather kernels, =1 the shared-memory tile kernels (see sph_launch_density)
static bool use_tile_kernels() {
    static const int v = [] {
        const char *e = getenv("SPH_B200_SPH_TILE");
        return e ? atoi(e) : SPH_TILE_DEFAULT;
    }();
    return v != 0;
}

cudaError_t sph_launch_smoothing(sph_handle *h) {
    sph_note(1);
    smoothing_kernel<<<(int)((h->N + HB - 1) / HB), HB, 0, h->stream>>>(h->N, h->d2k, h->scal, h->pos4, h->hs, h->ecnt);
    return cudaGetLastError();
}

template <bool EOS>
static cudaError_t launch_density(sph_handle *h, int64_t t0, int64_t t1) {
    const int64_t nt = t1 - t0;
    const int blocks = (int)((nt + HB - 1) / HB);
    const ExtrasOut x = extras_of(h, t0, t1);
    const int poly = h->p.eos == SPH_EOS_POLYTROPIC;
    // the tile kernels need 16-byte aligned list tiles: t0 is a multiple of 128 (set_partition) and so is NL
    if (use_tile_kernels() && (t0 % HB) == 0) {
        const size_t smem = DENS_SMEM + (size_t)h->K * HB * sizeof(int);
        static bool attr = false;
        if (!attr) {
            cudaError_t e = cudaFuncSetAttribute(density_kernel<true, EOS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
            if (e != cudaSuccess) return e;
            attr = true;
        }
        density_kernel<true, EOS><<<blocks, HB, smem, h->stream>>>(h->N, h->NL, h->K, t0, t1, h->pos4, h->nbr, h->perm, h->kid,
                                                                   h->p.m, poly, h->scal, x, h->rho_s, h->vel4, h->p.cs, h->p.gamma,
                                                                   h->hr, ForceRecs{h->fa, h->fb, h->fc});
    } else {
        density_kernel<false, EOS><<<blocks, HB, 0, h->stream>>>(h->N, h->NL, h->K, t0, t1, h->pos4, h->nbr, h->perm, h->kid,
                                                                 h->p.m, poly, h->scal, x, h->rho_s, h->vel4, h->p.cs, h->p.gamma,
                                                                 h->hr, ForceRecs{h->fa, h->fb, h->fc});
    }
    return cudaGetLastError();
}

// with_eos: one rank - the EOS of every target is closed in the same kernel (hr and the force records written); several ranks - only
// rho_s of the owned targets, sph_launch_eos follows the all-gather
cudaError_t sph_launch_density(sph_handle *h, int64_t t0, int64_t t1, bool with_eos) {
    if (t1 <= t0) return cudaSuccess;
    sph_note(1);
    return with_eos ? launch_density<true>(h, t0, t1) : launch_density<false>(h, t0, t1);
}

cudaError_t sph_launch_outbox_header(sph_handle *h) {
    sph_note(1);
    outbox_header_kernel<<<1, 1, 0, h->stream>>>(h->outbox, (int)h->obcap, h->scal);
    return cudaGetLastError();
}

cudaError_t sph_launch_extras_merge(sph_handle *h, int64_t t0, int64_t t1) {
    sph_note(1);
    extras_merge_kernel<<<148 * 2, 256, 0, h->stream>>>(h->inbox, h->nranks, h->rank, h->obcap + 1, extras_of(h, t0, t1), h->scal);
    return cudaGetLastError();
}

cudaError_t sph_launch_extras_sort(sph_handle *h, int64_t t0, int64_t t1) {
    if (t1 <= t0) return cudaSuccess;
    sph_note(1);
    extras_sort_kernel<<<(int)((t1 - t0 + HB - 1) / HB), HB, 0, h->stream>>>(h->NL, t0, t1, h->ecap, h->ecnt, h->ext, h->scal);
    return cudaGetLastError();
}

cudaError_t sph_launch_eos(sph_handle *h) {
    sph_note(1);
    eos_kernel<<<(int)((h->N + HB - 1) / HB), HB, 0, h->stream>>>(h->N, h->rho_s, h->pos4, h->vel4, h->p.eos == SPH_EOS_POLYTROPIC,
                                                                  h->p.cs, h->p.gamma, h->scal, h->hr, ForceRecs{h->fa, h->fb, h->fc});
    return cudaGetLastError();
}

template <bool POLY>
static cudaError_t launch_force(sph_handle *h, int64_t t0, int64_t t1) {
    const int64_t nt = t1 - t0;
    const int blocks = (int)((nt + HB - 1) / HB);
    const double cs2 = h->p.cs * h->p.cs;
    if (use_tile_kernels() && (t0 % HB) == 0) {
        const size_t smem = FORCE_SMEM + (size_t)h->K * HB * sizeof(int);
        static bool attr = false;
        if (!attr) {
            cudaError_t e = cudaFuncSetAttribute(force_kernel<POLY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            if (e != cudaSuccess) return e;
            attr = true;
        }
        force_kernel<POLY, true><<<blocks, HB, smem, h->stream>>>(h->N, h->NL, h->NS, h->K, t0, t1, h->fa, h->fb, h->fc, h->d2k, h->nbr,
                                                                  h->perm, h->kid, h->ecnt, h->ext, h->ecap, h->p.m, h->p.alpha,
                                                                  h->p.beta, cs2, h->scal, h->s_ahyd, h->s_dkdt, h->s_sumvdw, h->s_mumax);
    } else {
        force_kernel<POLY, false><<<blocks, HB, 0, h->stream>>>(h->N, h->NL, h->NS, h->K, t0, t1, h->fa, h->fb, h->fc, h->d2k, h->nbr,
                                                                h->perm, h->kid, h->ecnt, h->ext, h->ecap, h->p.m, h->p.alpha,
                                                                h->p.beta, cs2, h->scal, h->s_ahyd, h->s_dkdt, h->s_sumvdw, h->s_mumax);
    }
    force_overflow_kernel<POLY><<<blocks, HB, 0, h->stream>>>(h->NS, t0, t1, h->fa, h->fb, h->fc, h->ecnt, h->ecap, h->ovf,
                                                              (int)h->ovcap, h->p.m, h->p.alpha, h->p.beta, cs2, h->scal, h->s_ahyd,
                                                              h->s_dkdt);
    return cudaGetLastError();
}

cudaError_t sph_launch_force(sph_handle *h, int64_t t0, int64_t t1) {
    if (t1 <= t0) return cudaSuccess;
    sph_note(2);
    return h->p.eos == SPH_EOS_POLYTROPIC ? launch_force<true>(h, t0, t1) : launch_force<false>(h, t0, t1);
}
