// integrate.cu -- acceleration assembly, adaptive time step, statistics row, predictor / corrector,
// entropy-function update.  All kernels are single streaming passes in the caller's particle order.
//
// Replaces F/isothermal_sim.jl:41-46 (acc = a_hyd - G g), :158-166 (dt), :168-192 (statistics),
// :197-200 (predictor), :206-212 (corrector) and F/polytrope_hydroKDTree.jl:313-315 (K update);
// polytropic twins at F/polytrope_sim.jl:42-48, :165-174, :177-205, :211-231.
// Products and sums of the integrator are rounded separately (__dmul_rn/__dadd_rn) so that the state
// advances bit-identically to a scalar evaluation of the reference's expressions given equal inputs.
#include "sph_internal.cuh"

namespace {

constexpr int IB = 256;
constexpr int RED_BLOCKS = 592;  // 4 x 148

inline int grid_for(int64_t n) {
    int64_t g = (n + IB - 1) / IB;
    const int64_t cap = 148 * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// stat_dev layout
enum { DV_T = 0, DV_DT = 1, DV_SUM = 2 /* 12 sums */, DV_L = 16 /* 3 sums */, DV_ROW = 20 /* 10 */ };

// walk results of sorted slot s: tile = s / 128 belongs to group tile / SPH_WALK_DEAL, which went to rank
// group % nranks as its (group / nranks)-th group
__device__ __forceinline__ const double *walk_slot(const double *__restrict__ walk_buf, int nranks, int64_t wchunk, int64_t s) {
    const int64_t tile = s >> 7;
    const int64_t group = tile / SPH_WALK_DEAL;
    return walk_buf + (size_t)(group % nranks) * 4 * wchunk + ((group / nranks) * SPH_WALK_DEAL + tile % SPH_WALK_DEAL) * 128 + (s & 127);
}

// per evaluation: total acceleration and h in the caller's particle order (integrator, next search radius)
__global__ void __launch_bounds__(IB) finish_kernel(int64_t N, int64_t NS, const int *__restrict__ perm,
                                                     const double *__restrict__ s_ahyd, const double *__restrict__ walk_buf,
                                                     int nranks, int64_t wchunk, const double2 *__restrict__ hr, double G,
                                                     double *__restrict__ acc, double *__restrict__ o_h) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = perm[s];
        const double *w = walk_slot(walk_buf, nranks, wchunk, s);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            acc[i + k * N] = __dsub_rn(s_ahyd[s + k * NS], __dmul_rn(G, w[k * wchunk]));   // ax -= G * g[:, 1]  (F/isothermal_sim.jl:41-43)
        o_h[i] = hr[s].x;
    }
}

// on demand (getters): the remaining per-particle results in the caller's particle order
__global__ void __launch_bounds__(IB) unpermute_kernel(int64_t N, int64_t NS, const int *__restrict__ perm,
                                                        const double *__restrict__ s_red, const double *__restrict__ walk_buf,
                                                        int nranks, int64_t wchunk, const double2 *__restrict__ hr,
                                                        const double2 *__restrict__ fc, double *__restrict__ o_ahyd,
                                                        double *__restrict__ o_g, double *__restrict__ o_rho,
                                                        double *__restrict__ o_phi, double *__restrict__ o_sumvdw,
                                                        double *__restrict__ o_mumax, double *__restrict__ o_cs,
                                                        double *__restrict__ o_dkdt) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = perm[s];
        const double *w = walk_slot(walk_buf, nranks, wchunk, s);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            o_ahyd[i + k * N] = s_red[s + k * NS];
            o_g[i + k * N] = w[k * wchunk];
        }
        o_rho[i] = hr[s].y;
        o_phi[i] = w[3 * wchunk];
        o_dkdt[i] = s_red[s + 3 * NS];
        o_sumvdw[i] = s_red[s + 4 * NS];
        o_mumax[i] = s_red[s + 5 * NS];
        o_cs[i] = fc[s].y;
    }
}

__device__ __forceinline__ double block_min(double v) {
    __shared__ double sm[IB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < IB / 32; ++w) v = fmin(v, sm[w]);
    return v;
}

__global__ void dt_init_kernel(unsigned long long *scal) { scal[SC_DT] = 0x7ff0000000000000ull; }

// dt = 0.3 * min(min 1/|div v|, min h/|v|, min sqrt(h/|a|), min h/(c + 1.2(alpha c + beta max_j mu)))
// evaluated in the sorted order of the evaluation it follows (vel4 = the velocities that evaluation was given)
__global__ void __launch_bounds__(IB) dt_kernel(int64_t N, int64_t NS, const double4 *__restrict__ vel4,
                                                 const double *__restrict__ s_red, const double *__restrict__ walk_buf,
                                                 int nranks, int64_t wchunk, const double2 *__restrict__ hr,
                                                 const double2 *__restrict__ fc, double G, double m, double alpha, double beta,
                                                 unsigned long long *__restrict__ scal) {
    double best = __longlong_as_double(0x7ff0000000000000LL);
    bool bad = false;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        const double4 v = vel4[s];
        const double *w = walk_slot(walk_buf, nranks, wchunk, s);
        const double ax = __dsub_rn(s_red[s], __dmul_rn(G, w[0]));
        const double ay = __dsub_rn(s_red[s + NS], __dmul_rn(G, w[wchunk]));
        const double az = __dsub_rn(s_red[s + 2 * NS], __dmul_rn(G, w[2 * wchunk]));
        const double vel_r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)), __dmul_rn(v.z, v.z)));
        const double a_r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
        const double2 hrs = hr[s];
        const double h = hrs.x, c = fc[s].y;
        const double abs_div_v = fabs(-(__dmul_rn(m, s_red[s + 4 * NS])) / hrs.y);
        const double c1 = 1 / abs_div_v;
        const double c2 = h / vel_r;
        const double c3 = sqrt(h / a_r);
        const double c4 = h / __dadd_rn(c, __dmul_rn(1.2, __dadd_rn(__dmul_rn(alpha, c), __dmul_rn(beta, s_red[s + 5 * NS]))));
        best = fmin(best, fmin(fmin(c1, c2), fmin(c3, c4)));
        // fmin drops NaNs, the reference's minimum() propagates them (a blown-up state ends its loop): remember them
        bad |= (c1 != c1) | (c2 != c2) | (c3 != c3) | (c4 != c4);
    }
    best = block_min(best);
    if (threadIdx.x == 0) atomicMin(&scal[SC_DT], (unsigned long long)__double_as_longlong(best));
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(&scal[SC_ERR], (unsigned long long)ERRF_NAN);
}

__global__ void dt_final_kernel(unsigned long long *__restrict__ scal, double *__restrict__ dv) {
    const double dt = 0.3 * __longlong_as_double((long long)scal[SC_DT]);
    const bool nan = (scal[SC_ERR] & (unsigned long long)ERRF_NAN) != 0ull;
    dv[DV_DT] = nan ? __longlong_as_double(0x7ff8000000000000LL) : dt;
    if (nan) scal[SC_STICKY] |= (unsigned long long)ERRF_NAN;    // survives the reset at the next evaluation's start
}

// deterministic two-stage sum of NV values per particle
template <int NV>
__device__ __forceinline__ void block_sum_store(double (&v)[NV], double *__restrict__ partial) {
    __shared__ double sm[NV][IB / 32];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < IB / 32; ++w) s += sm[threadIdx.x][w];
        partial[(int64_t)blockIdx.x * NV + threadIdx.x] = s;
    }
}

// sums: 0 sum |v|^2, 1 sum PHI, 2-4 sum pos, 5-7 sum vel, 8 sum K/(gamma-1) rho^(gamma-1)
__global__ void __launch_bounds__(IB) stats1_kernel(int64_t N, const double *__restrict__ pos, const double *__restrict__ vel,
                                                     const double *__restrict__ walk_buf, int nranks, int64_t wchunk,
                                                     const double2 *__restrict__ hr, const double4 *__restrict__ vel4, int poly,
                                                     double gamma, double *__restrict__ partial) {
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double vx = vel[i], vy = vel[i + N], vz = vel[i + 2 * N];
        const double vr = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
        v[0] += __dmul_rn(vr, vr);
        v[1] += walk_slot(walk_buf, nranks, wchunk, i)[3 * wchunk];   // sums do not care about the order: slot i of the sorted arrays
        v[2] += pos[i]; v[3] += pos[i + N]; v[4] += pos[i + 2 * N];
        v[5] += vx; v[6] += vy; v[7] += vz;
        if (poly) v[8] += vel4[i].w / (gamma - 1) * pow(hr[i].y, gamma - 1);   // K_i and rho_i of sorted slot i
    }
    block_sum_store<9>(v, partial);
}

// sum of the per-block partials, one warp per value, in a fixed order (lane l adds blocks l, l + 32, ..; then a
// shuffle tree): deterministic
template <int NV>
__global__ void __launch_bounds__(32 * NV) final_sum_kernel(const double *__restrict__ partial, int nblocks, double *__restrict__ out) {
    const int v = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += partial[(int64_t)b * NV + v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[v] = s;
}

// angular momentum about the centre of mass: l = m * sum (pos - r_com) x vel  (F/isothermal_sim.jl:186)
__global__ void __launch_bounds__(IB) stats2_kernel(int64_t N, const double *__restrict__ pos, const double *__restrict__ vel,
                                                     const double *__restrict__ dv, double *__restrict__ partial) {
    const double rcx = dv[DV_SUM + 2] / (double)N, rcy = dv[DV_SUM + 3] / (double)N, rcz = dv[DV_SUM + 4] / (double)N;
    double v[3] = {0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double ax = pos[i] - rcx, ay = pos[i + N] - rcy, az = pos[i + 2 * N] - rcz;
        const double bx = vel[i], by = vel[i + N], bz = vel[i + 2 * N];
        v[0] += __dsub_rn(__dmul_rn(ay, bz), __dmul_rn(az, by));
        v[1] += __dsub_rn(__dmul_rn(az, bx), __dmul_rn(ax, bz));
        v[2] += __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
    }
    block_sum_store<3>(v, partial);
}

// stats row [t, T, V, U, Etot, rcx, rcy, rcz, |p|, |L|] and the step log entry
__global__ void stats_row_kernel(int64_t N, double m, double G, int poly, double U_iso, double *__restrict__ dv,
                                 double *__restrict__ log_row /* 11 doubles or null */) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double *S = dv + DV_SUM;
    const double T = 0.5 * m * S[0];
    const double V = G / 2 * m * S[1];
    double U, Etot;
    if (poly) { U = m * S[8]; Etot = T + V + U; }          // F/polytrope_sim.jl:186,189
    else { U = U_iso; Etot = T + V + 2 * U; }              // F/isothermal_sim.jl:177
    const double px = S[5] * m, py = S[6] * m, pz = S[7] * m;
    const double lx = dv[DV_L] * m, ly = dv[DV_L + 1] * m, lz = dv[DV_L + 2] * m;
    double *row = dv + DV_ROW;
    row[0] = dv[DV_T]; row[1] = T; row[2] = V; row[3] = U; row[4] = Etot;
    row[5] = S[2] / (double)N; row[6] = S[3] / (double)N; row[7] = S[4] / (double)N;
    row[8] = sqrt((px * px + py * py) + pz * pz);
    row[9] = sqrt((lx * lx + ly * ly) + lz * lz);
    if (log_row) {
        log_row[0] = dv[DV_DT];
        for (int k = 0; k < 10; ++k) log_row[1 + k] = row[k];
    }
}

// pos_half = pos + vel*dt/2 ; vel_half = vel + acc*dt/2   (F/isothermal_sim.jl:197,200)
__global__ void __launch_bounds__(IB) predict_kernel(int64_t n3, const double *__restrict__ pos, const double *__restrict__ vel,
                                                      const double *__restrict__ acc, const double *__restrict__ dv,
                                                      double *__restrict__ pos_half, double *__restrict__ vel_half) {
    const double dt = dv[DV_DT];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
        pos_half[i] = __dadd_rn(pos[i], __dmul_rn(vel[i], dt) / 2);
        vel_half[i] = __dadd_rn(vel[i], __dmul_rn(acc[i], dt) / 2);
    }
}

// vel += acc*dt ; pos += vel*dt - (1/2)*acc*dt^2  with the UPDATED vel  (F/isothermal_sim.jl:206,209)
__global__ void __launch_bounds__(IB) correct_kernel(int64_t n3, double *__restrict__ pos, double *__restrict__ vel,
                                                      const double *__restrict__ acc, const double *__restrict__ dv) {
    const double dt = dv[DV_DT];
    const double dt2 = __dmul_rn(dt, dt);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = acc[i];
        const double v = __dadd_rn(vel[i], __dmul_rn(a, dt));
        vel[i] = v;
        pos[i] = __dadd_rn(pos[i], __dsub_rn(__dmul_rn(v, dt), __dmul_rn(__dmul_rn(0.5, a), dt2)));
    }
}

__global__ void advance_time_kernel(double *dv) { dv[DV_T] = __dadd_rn(dv[DV_T], dv[DV_DT]); }   // t += dt (:212)

// K .+= 1/2*(gamma-1) ./ rho.^(gamma-1) .* dK * dt   called with dt/2  (F/polytrope_hydroKDTree.jl:313-315)
__global__ void __launch_bounds__(IB) evolve_k_kernel(int64_t N, double *__restrict__ kent, const int *__restrict__ perm,
                                                       const double2 *__restrict__ hr, const double *__restrict__ s_dkdt,
                                                       double gamma, const double *__restrict__ dv) {
    const double hdt = dv[DV_DT] / 2;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = perm[s];
        kent[i] = __dadd_rn(kent[i], __dmul_rn(__dmul_rn(1.0 / 2 * (gamma - 1) / pow(hr[s].y, gamma - 1), s_dkdt[s]), hdt));
    }
}

__global__ void set_time_kernel(double *dv, double t) { dv[DV_T] = t; }

}  // namespace

cudaError_t sph_launch_finish(sph_handle *h, double *acc_out) {
    sph_note(1);
    finish_kernel<<<grid_for(h->N), IB, 0, h->stream>>>(h->N, h->NS, h->perm, h->s_ahyd, h->walk_buf, h->nranks, h->walk_chunk,
                                                         h->hr, h->p.G, acc_out, h->o_h);
    h->outputs_fresh = false;
    return cudaGetLastError();
}

// getters: un-permute the remaining results of the last evaluation once
cudaError_t sph_launch_unpermute(sph_handle *h) {
    if (h->outputs_fresh) return cudaSuccess;
    sph_note(1);
    unpermute_kernel<<<grid_for(h->N), IB, 0, h->stream>>>(h->N, h->NS, h->perm, h->s_red, h->walk_buf, h->nranks, h->walk_chunk,
                                                            h->hr, h->fc, h->o_ahyd, h->o_g, h->o_rho, h->o_phi, h->o_sumvdw,
                                                            h->o_mumax, h->o_cs, h->o_dkdt);
    h->outputs_fresh = true;
    return cudaGetLastError();
}

cudaError_t sph_launch_dt(sph_handle *h) {
    sph_note(3);
    dt_init_kernel<<<1, 1, 0, h->stream>>>(h->scal);
    dt_kernel<<<RED_BLOCKS, IB, 0, h->stream>>>(h->N, h->NS, h->vel4, h->s_red, h->walk_buf, h->nranks, h->walk_chunk, h->hr,
                                                 h->fc, h->p.G, h->p.m, h->p.alpha, h->p.beta, h->scal);
    dt_final_kernel<<<1, 1, 0, h->stream>>>(h->scal, h->stat_dev);
    return cudaGetLastError();
}

cudaError_t sph_launch_stats(sph_handle *h, double *log_row) {
    const bool poly = h->p.eos == SPH_EOS_POLYTROPIC;
    sph_note(5);
    stats1_kernel<<<RED_BLOCKS, IB, 0, h->stream>>>(h->N, h->pos, h->vel, h->walk_buf, h->nranks, h->walk_chunk, h->hr, h->vel4,
                                                     poly, h->p.gamma, h->red_partial);
    final_sum_kernel<9><<<1, 32 * 9, 0, h->stream>>>(h->red_partial, RED_BLOCKS, h->stat_dev + DV_SUM);
    stats2_kernel<<<RED_BLOCKS, IB, 0, h->stream>>>(h->N, h->pos, h->vel, h->stat_dev, h->red_partial);
    final_sum_kernel<3><<<1, 32 * 3, 0, h->stream>>>(h->red_partial, RED_BLOCKS, h->stat_dev + DV_L);
    stats_row_kernel<<<1, 32, 0, h->stream>>>(h->N, h->p.m, h->p.G, poly, h->p.U_iso, h->stat_dev, log_row);
    return cudaGetLastError();
}

cudaError_t sph_launch_predict(sph_handle *h) {
    sph_note(1);
    predict_kernel<<<grid_for(3 * h->N), IB, 0, h->stream>>>(3 * h->N, h->pos, h->vel, h->acc, h->stat_dev,
                                                              h->pos_half, h->vel_half);
    return cudaGetLastError();
}

cudaError_t sph_launch_correct(sph_handle *h) {
    sph_note(2);
    correct_kernel<<<grid_for(3 * h->N), IB, 0, h->stream>>>(3 * h->N, h->pos, h->vel, h->acc, h->stat_dev);
    advance_time_kernel<<<1, 1, 0, h->stream>>>(h->stat_dev);
    return cudaGetLastError();
}

cudaError_t sph_launch_evolve_k(sph_handle *h) {
    sph_note(1);
    evolve_k_kernel<<<grid_for(h->N), IB, 0, h->stream>>>(h->N, h->kent, h->perm, h->hr, h->s_dkdt, h->p.gamma, h->stat_dev);
    return cudaGetLastError();
}

cudaError_t sph_launch_set_time(sph_handle *h, double t) {
    sph_note(1);
    set_time_kernel<<<1, 1, 0, h->stream>>>(h->stat_dev, t);
    return cudaGetLastError();
}
