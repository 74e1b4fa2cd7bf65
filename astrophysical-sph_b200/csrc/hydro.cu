// hydro.cu -- density + EOS and pressure / artificial-viscosity force over the neighbour lists.
//
// Replaces HJL.W / getDensity / getPressure / gradW / getAV / hydroCalculation and the scatter part of
// evolve_K! (F/isothermal_hydroKDTree.jl:5-245, F/polytrope_hydroKDTree.jl:5-341).  The reference
// materialises ~35 N x K Float64 matrices per call; here every pair quantity lives in registers and only
// the per-particle reductions its caller consumes are written:
//   rho, h, P/rho^2, c_i                                       (density + EOS)
//   a_hyd, sum_j v_ij.gradW_ij, max_j mu_ij, dK/dt scatter sum (force)
// Lists are N x K int32 column-major (the Julia layout), so the lanes of a warp - consecutive targets in
// key order - read consecutive list entries (coalesced) and gather spatially close particles (L1/L2 hits).
#include "sph_internal.cuh"

namespace {

constexpr int HB = 128;
constexpr double PI_D = 3.141592653589793;

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

__global__ void __launch_bounds__(HB) density_kernel(int64_t N, int K, int64_t t0, int64_t t1,
                                                      const double4 *__restrict__ pos4, const int *__restrict__ nbr,
                                                      const double *__restrict__ d2k, double m, int poly,
                                                      const unsigned long long *__restrict__ scal,
                                                      double2 *__restrict__ hr) {
    if (scal[SC_ERR] != 0ull) return;
    const int64_t s = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t1) return;
    const double4 pi = pos4[s];
    const double h = sqrt(d2k[s]) / 2;               // h = r[:, end] ./ 2   (:151)
    const double ct = 1 / (PI_D * (h * h * h));
    const double hinv = 1 / h;
    double sum = 0.0;
    for (int j = 0; j < K; ++j) {
        const int nj = nbr[s + (int64_t)j * N];
        const double4 pj = pos4[nj];
        // q = r / h from one rsqrt seed + Newton step (1e-13 relative; the parity bar for rho is 1e-9)
        const double d2 = sph_d2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        const double q = d2 > 0.0 ? d2 * fast_rsqrt(d2) * hinv : 0.0;
        sum += kernel_W(ct, q, poly != 0);           // rho_i = m * sum_j w_ij  (:175)
    }
    hr[s] = make_double2(h, m * sum);
}

// EOS closure for ALL particles (after the density all-gather in multi-GPU runs):
//   isothermal  P = cs^2 rho (F/isothermal_hydroKDTree.jl:190), c = cs
//   polytropic  P = K rho^gamma (F/polytrope_hydroKDTree.jl:216), c_i = sqrt(gamma K rho^(gamma-1)) (:186)
// also refreshes pos4.w = h (used by the leaf interactions of the tree walk).
__global__ void __launch_bounds__(HB) eos_kernel(int64_t N, const double2 *__restrict__ hr,
                                                  const double4 *__restrict__ vel4, int poly, double cs, double gamma,
                                                  const unsigned long long *__restrict__ scal,
                                                  double *__restrict__ prr, double *__restrict__ cs_s,
                                                  double4 *__restrict__ pos4 /* null: pos4.w already holds h */) {
    if (scal[SC_ERR] != 0ull) return;
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const double2 a = hr[s];
    const double rho = a.y;
    double P, c;
    if (!poly) {
        P = cs * cs * rho;
        c = cs;
    } else {
        const double Kent = vel4[s].w;
        c = sqrt(gamma * Kent * pow(rho, gamma - 1));
        P = Kent * pow(rho, gamma);
    }
    prr[s] = P / (rho * rho);
    cs_s[s] = c;
    if (pos4) pos4[s].w = a.x;
}

// h = r[:, end] ./ 2 (:151) straight from the search result into pos4.w: lets the tree walk start while the density
// runs on the second stream (single GPU; same expression as density_kernel, so the value is the one it stores in hr)
__global__ void __launch_bounds__(HB) smoothing_kernel(int64_t N, const double *__restrict__ d2k,
                                                        const unsigned long long *__restrict__ scal,
                                                        double4 *__restrict__ pos4) {
    if (scal[SC_ERR] != 0ull) return;
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    pos4[s].w = sqrt(d2k[s]) / 2;
}

// Pair loop of hydroCalculation / getAV / evolve_K!.  The target's own update is kept in registers and
// stored once; the reaction on the neighbour (a_j += ct*gradW, dK_j += c2) is a double-precision RED to L2.
template <bool POLY>
__global__ void __launch_bounds__(HB, 6) force_kernel(int64_t N, int64_t NS, int K, int64_t t0, int64_t t1,
                                                    const double4 *__restrict__ pos4, const double4 *__restrict__ vel4,
                                                    const double2 *__restrict__ hr, const double *__restrict__ prr,
                                                    const double *__restrict__ cs_s, const int *__restrict__ nbr,
                                                    double m, double alpha, double beta,
                                                    const unsigned long long *__restrict__ scal,
                                                    double *__restrict__ ahyd, double *__restrict__ dkdt,
                                                    double *__restrict__ sumvdw, double *__restrict__ mumax) {
    if (scal[SC_ERR] != 0ull) return;
    const int64_t s = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= t1) return;
    const double4 pi = pos4[s];
    const double4 vi = vel4[s];
    const double2 hri = hr[s];
    const double hi = hri.x, rhoi = hri.y;
    const double prri = prr[s], ci = cs_s[s];
    const double h2 = hi * hi;
    const double ct4 = 1 / (PI_D * (h2 * h2));
    const double hinv = 1 / hi;
    double ax = 0.0, ay = 0.0, az = 0.0, svdw = 0.0, dk = 0.0;
    double mmax = -__longlong_as_double(0x7ff0000000000000LL);
    for (int j = 0; j < K; ++j) {
        const int nj = nbr[s + (int64_t)j * N];
        const double4 pj = pos4[nj];
        const double4 vj = vel4[nj];
        const double2 hrj = hr[nj];
        const double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;   // getTreeDiffs: f_i - f_j (:93)
        const double d2 = sph_d2_exact(dx, dy, dz);
        const double rinv = d2 > 0.0 ? fast_rsqrt(d2) : 0.0;
        const double q = d2 * rinv * hinv;
        const double dW = kernel_dWdr(ct4, hinv, q, rinv, POLY);
        const double gx = dW * dx, gy = dW * dy, gz = dW * dz;
        const double h_avg = (hi + hrj.x) / 2;                               // getVectorTreeAvgs (:111)
        const double rho_avg = (rhoi + hrj.y) / 2;
        const double vx = vi.x - vj.x, vy = vi.y - vj.y, vz = vi.z - vj.z;
        const double vdr = (vx * dx + vy * dy) + vz * dz;                    // (:210)
        const double mu = fmin(h_avg * vdr * fast_rcp(d2 + 0.01 * (h_avg * h_avg)), 0.0);   // (:211)
        const double Pi = ((-alpha) * ci * mu + beta * (mu * mu)) * fast_rcp(rho_avg);       // (:213)
        const double vdw = (vx * gx + vy * gy) + vz * gz;
        svdw += vdw;
        mmax = fmax(mmax, mu);
        // hydroCalculation and evolve_K! start at column 2 (:226, poly :301): column 1 is the particle itself.
        // (lists arrive unordered from the grouped search, so the self entry is recognised by its index)
        if (nj == s) continue;
        double ct;
        if (!POLY) ct = m * (prri + Pi / 2);                                 // iso :232
        else ct = m * ((prri + prr[nj]) + Pi) / 2;                           // poly :235
        const double fx = ct * gx, fy = ct * gy, fz = ct * gz;
        ax -= fx; ay -= fy; az -= fz;
        atomicAdd(&ahyd[nj], fx);
        atomicAdd(&ahyd[nj + NS], fy);
        atomicAdd(&ahyd[nj + 2 * NS], fz);
        if (POLY) {
            const double c2 = m * Pi * vdw / 2;                              // evolve_K! poly :305-311
            dk += c2;
            atomicAdd(&dkdt[nj], c2);
        }
    }
    atomicAdd(&ahyd[s], ax);
    atomicAdd(&ahyd[s + NS], ay);
    atomicAdd(&ahyd[s + 2 * NS], az);
    if (POLY) atomicAdd(&dkdt[s], dk);
    sumvdw[s] = svdw;
    mumax[s] = mmax;
}

}  // namespace

cudaError_t sph_launch_density(sph_handle *h, int64_t t0, int64_t t1) {
    if (t1 <= t0) return cudaSuccess;
    sph_note(1);
    const int64_t nt = t1 - t0;
    density_kernel<<<(int)((nt + HB - 1) / HB), HB, 0, h->stream>>>(h->N, h->K, t0, t1, h->pos4, h->nbr, h->d2k,
                                                                    h->p.m, h->p.eos == SPH_EOS_POLYTROPIC, h->scal,
                                                                    h->hr);
    return cudaGetLastError();
}

cudaError_t sph_launch_eos(sph_handle *h, bool write_h) {
    sph_note(1);
    eos_kernel<<<(int)((h->N + HB - 1) / HB), HB, 0, h->stream>>>(h->N, h->hr, h->vel4,
                                                                  h->p.eos == SPH_EOS_POLYTROPIC, h->p.cs, h->p.gamma,
                                                                  h->scal, h->prr, h->cs_s, write_h ? h->pos4 : nullptr);
    return cudaGetLastError();
}

cudaError_t sph_launch_smoothing(sph_handle *h) {
    sph_note(1);
    smoothing_kernel<<<(int)((h->N + HB - 1) / HB), HB, 0, h->stream>>>(h->N, h->d2k, h->scal, h->pos4);
    return cudaGetLastError();
}

cudaError_t sph_launch_force(sph_handle *h, int64_t t0, int64_t t1) {
    const int64_t N = h->N;
    cudaMemsetAsync(h->s_red, 0, sizeof(double) * 6 * h->NS, h->stream);
    if (t1 <= t0) return cudaGetLastError();
    sph_note(1);
    const int64_t nt = t1 - t0;
    if (h->p.eos == SPH_EOS_POLYTROPIC)
        force_kernel<true><<<(int)((nt + HB - 1) / HB), HB, 0, h->stream>>>(
            N, h->NS, h->K, t0, t1, h->pos4, h->vel4, h->hr, h->prr, h->cs_s, h->nbr, h->p.m, h->p.alpha, h->p.beta,
            h->scal, h->s_ahyd, h->s_dkdt, h->s_sumvdw, h->s_mumax);
    else
        force_kernel<false><<<(int)((nt + HB - 1) / HB), HB, 0, h->stream>>>(
            N, h->NS, h->K, t0, t1, h->pos4, h->vel4, h->hr, h->prr, h->cs_s, h->nbr, h->p.m, h->p.alpha, h->p.beta,
            h->scal, h->s_ahyd, h->s_dkdt, h->s_sumvdw, h->s_mumax);
    return cudaGetLastError();
}
