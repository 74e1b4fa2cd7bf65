// gravity.cu -- Barnes-Hut walk of the linear octree with the reference's opening rule and
// spline-softened leaf interactions.
//
// Replaces GJL.compute_g / gravity_acc / Kernels / min_distance2_point_to_cell
// (F/gravOctree_Single.jl:5-29, :231-304).  The reference walks the tree once per particle with a FIFO
// queue; here one warp walks it for 32 key-adjacent targets at once with a shared depth-first stack whose
// entries carry a lane mask: every lane evaluates the reference's two-clause acceptance test for ITS OWN
// particle (s^2/d^2 < theta^2 with d to the node COM, and h_i^2/mindist^2(p_i, cell) < 0.25, :265); lanes
// that accept add the monopole, the others stay in the mask pushed with the children.  Decisions are
// therefore exactly per-particle as in the reference; only the floating-point summation order differs
// (the reference's own order already drifts through its leaf list surgery, :293-300).
// Node data is read with warp-uniform addresses (one 32 B sector per double4, broadcast to the warp).
#include "sph_internal.cuh"

namespace {

constexpr int GW_WARPS = 4;
constexpr int GW_STACK = 192;

// Kernels (F/gravOctree_Single.jl:5-29): returns grad(PHI)/r and PHI of the spline-softened potential
__device__ __forceinline__ void grav_kernels(double r, double h, double &gPHI, double &PHI) {
    const double q = r / h;
    if (q > 2.0) {
        const double r3 = r * r * r;
        gPHI = 1 / r3;
        PHI = -1 / r;
        return;
    }
    const double h2 = h * h;
    const double q2 = q * q, q3 = q2 * q, q4 = q2 * q2, q5 = q4 * q;
    if (q <= 1.0) {
        const double h3 = h2 * h, h4 = h2 * h2;
        const double r2 = r * r, r3 = r2 * r;
        gPHI = (1 / h2) * ((4.0 / 3 / h - 6.0 / 5 * (r2 / h3)) + 1.0 / 2 * (r3 / h4));
        PHI = (1 / h) * (((2.0 / 3 * q2 - 3.0 / 10 * q4) + 1.0 / 10 * q5) - 7.0 / 5);
    } else {
        gPHI = ((1 / h2) * ((((8.0 / 3 * q - 3 * q2) + 6.0 / 5 * q3) - 1.0 / 6 * q4) - 1.0 / 15 * (1 / q2))) / r;
        PHI = (1 / h) * (((((4.0 / 3 * q2 - q3) + 3.0 / 10 * q4) - 1.0 / 30 * q5) - 8.0 / 5) + 1.0 / 15 / q);
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(GW_WARPS * 32) walk_kernel(int64_t NS, int64_t t0, int64_t t1,
                                                              const double4 *__restrict__ pos4, SphTree t,
                                                              double theta_sq, double m,
                                                              unsigned long long *__restrict__ scal,
                                                              double *__restrict__ g, double *__restrict__ phi) {
    __shared__ int2 s_stack[GW_WARPS][GW_STACK];
    if (scal[SC_ERR] != 0ull) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int2 *stack = s_stack[warp];
    const int64_t s = t0 + ((int64_t)blockIdx.x * GW_WARPS + warp) * 32 + lane;
    const bool active = s < t1;
    double px = 0, py = 0, pz = 0, hi = 1.0;
    if (active) {
        const double4 p = pos4[s];  // .w = h_i
        px = p.x; py = p.y; pz = p.z; hi = p.w;
    }
    const double hi2 = hi * hi;
    double gx = 0.0, gy = 0.0, gz = 0.0, ph = 0.0;
    unsigned long long visits = 0;
    const unsigned amask = __ballot_sync(0xffffffffu, active);
    int sp = 0;
    if (amask) {
        // the walk starts at the root's children; the root itself is never tested (:246-249)
        const int2 R = t.nodeI[0];
        if (lane < R.y) stack[lane] = make_int2(R.x + lane, (int)amask);
        sp = R.y;
    }
    __syncwarp();
    while (sp > 0) {
        const int2 top = stack[--sp];
        __syncwarp();
        const int n = top.x;
        const bool mine = (((unsigned)top.y) >> lane) & 1u;
        const int2 I = t.nodeI[n];
        const double4 A = t.nodeA[n];
        const double dx = px - A.x, dy = py - A.y, dz = pz - A.z;   // p_i - rCOM (:255)
        const double d_sq = (dx * dx + dy * dy) + dz * dz;
        if (COUNT && mine) ++visits;
        if (I.y == 0) {
            // leaf = one particle j (sorted slot I.x); A.w carries h_j.  The target's own leaf is skipped
            // (the reference removes it from its parent's child list, :293-294).
            if (mine && (int64_t)I.x != s) {
                const double h_ij = (hi + A.w) / 2;                  // (:259)
                double gP, pot;
                grav_kernels(sqrt(d_sq), h_ij, gP, pot);
                gx += m * (gP * dx); gy += m * (gP * dy); gz += m * (gP * dz);   // (:263)
                ph += m * pot;                                                   // (:264)
            }
        } else {
            bool open = false;
            if (mine) {
                const double4 B = t.nodeB[n];
                const double4 C = t.nodeC[n];
                const double ex = fmax(fmax(B.x - px, 0.0), px - B.w);           // (:231-236)
                const double ey = fmax(fmax(B.y - py, 0.0), py - C.x);
                const double ez = fmax(fmax(B.z - pz, 0.0), pz - C.y);
                const double md2 = (ex * ex + ey * ey) + ez * ez;
                const bool accept = (C.z / d_sq < theta_sq) && (hi2 / md2 < 0.25);   // (:265)
                if (accept) {
                    const double d = sqrt(d_sq);
                    const double f = A.w / (d * d * d);                          // (:266-268)
                    gx += f * dx; gy += f * dy; gz += f * dz;
                    ph += -A.w / d;                                              // (:269)
                } else {
                    open = true;
                }
            }
            const unsigned om = __ballot_sync(0xffffffffu, open);
            if (om) {
                if (sp + I.y > GW_STACK) {  // cannot happen for depth <= 21; never write out of bounds
                    if (lane == 0) atomicOr(scal + SC_ERR, (unsigned long long)ERRF_STACK);
                    break;
                }
                if (lane < I.y) stack[sp + lane] = make_int2(I.x + lane, (int)om);
                sp += I.y;
            }
        }
        __syncwarp();
    }
    if (active) {
        g[s] = gx; g[s + NS] = gy; g[s + 2 * NS] = gz;
        phi[s] = ph - (m * (7.0 / 5) / hi);                                      // (:303)
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) visits += __shfl_xor_sync(0xffffffffu, visits, o);
        if (lane == 0) atomicAdd(scal + SC_VISITS, visits);
    }
    (void)lt;
}

// after the search: leaves carry h_j in nodeA.w (their mass is the constant m)
__global__ void leaf_h_kernel(SphTree t, const double4 *__restrict__ pos4, const unsigned long long *__restrict__ scal) {
    if (scal[SC_ERR] != 0ull) return;
    const int64_t M = (int64_t)scal[SC_NNODES];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < M; k += (int64_t)gridDim.x * blockDim.x) {
        const int2 I = t.nodeI[k];
        if (I.y == 0) t.nodeA[k].w = pos4[I.x].w;
    }
}

}  // namespace

cudaError_t sph_launch_walk(sph_handle *h, int64_t t0, int64_t t1) {
    sph_note(1);
    leaf_h_kernel<<<148 * 8, 256, 0, h->stream>>>(h->tree, h->pos4, h->scal);
    if (t1 <= t0) return cudaGetLastError();
    sph_note(1);
    const int64_t nt = t1 - t0;
    const int64_t blocks = (nt + GW_WARPS * 32 - 1) / (GW_WARPS * 32);
    const double th2 = h->p.theta * h->p.theta;
    static const bool count = getenv("SPH_B200_COUNT_VISITS") != nullptr;
    if (count)
        walk_kernel<true><<<(int)blocks, GW_WARPS * 32, 0, h->stream>>>(h->NS, t0, t1, h->pos4, h->tree, th2, h->p.m,
                                                                        h->scal, h->s_g, h->s_phi);
    else
        walk_kernel<false><<<(int)blocks, GW_WARPS * 32, 0, h->stream>>>(h->NS, t0, t1, h->pos4, h->tree, th2, h->p.m,
                                                                         h->scal, h->s_g, h->s_phi);
    return cudaGetLastError();
}
