// tree.cu -- domain size, octant keys, permutation into key order, and the linear octree.
//
// Replaces GJL.Octree / addNodes! / build_octree! / setCOMs! (F/gravOctree_Single.jl:78-227).
// The reference builds the tree breadth-first by re-bucketing particle lists; here the tree is derived
// from the sorted keys with one thread per particle / per node and no recursion:
//   * particles are ordered by their octant path: sorted by key word 0 (levels 0..20), and the rare runs of particles
//     that share all of word 0 are ordered by word 1 (levels 21..41) in place (key_runs_kernel);
//   * particle i shares cpl(i-1), cpl(i) leading octant levels with its sorted neighbours;
//   * the cells that START at particle i have depths cpl(i-1)+1 .. max(cpl(i-1),cpl(i))+1 (the last one
//     is the leaf that holds i alone; the root is the depth-0 cell starting at particle 0);
//   * a stable counting sort of that (start, depth) list by depth gives exactly the node order of
//     build_octree!'s breadth-first loop (:217-223), so children of a node are contiguous;
//   * cell geometry replays the reference's centre/bounds recurrence bit for bit (sph_cell_of);
//   * masses / centres of mass are accumulated bottom-up in one launch (arrival counters), children in
//     octant order, as setCOMs! does in its reverse sweep (:183-211).
#include <cstdlib>

#include "sph_internal.cuh"

namespace {

constexpr int TB = 256;
inline int grid_for(int64_t n, int tb = TB) {
    int64_t g = (n + tb - 1) / tb;
    const int64_t cap = 148 * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---- l_domain = maximum(abs.(pos))  (F/isothermal_sim.jl:33) -----------------------------------
__global__ void __launch_bounds__(TB) absmax_kernel(const double *__restrict__ pos, int64_t n3,
                                                     unsigned long long *__restrict__ scal) {
    double mx = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x)
        mx = fmax(mx, fabs(pos[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __shared__ double sm[TB / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TB / 32; ++w) mx = fmax(mx, sm[w]);
        // non-negative doubles order like their bit patterns
        atomicMax(&scal[SC_LDOM], (unsigned long long)__double_as_longlong(mx));
        // an error of an earlier evaluation that the host has not collected yet (sph_step enqueues several evaluations
        // before it reads the flags) keeps the following evaluations from running on the broken state
        if (blockIdx.x == 0 && scal[SC_STICKY] != 0ull) atomicOr(&scal[SC_ERR], scal[SC_STICKY]);
    }
}

__global__ void __launch_bounds__(TB) keys_kernel(const double *__restrict__ pos, int64_t N,
                                                   const unsigned long long *__restrict__ scal,
                                                   uint64_t *__restrict__ keys, int *__restrict__ vals) {
    const double l = __longlong_as_double((long long)scal[SC_LDOM]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = sph_octant_key(pos[i], pos[i + N], pos[i + 2 * N], l);
        vals[i] = (int)i;
    }
}

__global__ void __launch_bounds__(TB) permute_kernel(const double *__restrict__ pos, const double *__restrict__ vel,
                                                      const double *__restrict__ kent, const int *__restrict__ perm,
                                                      int64_t N, double4 *__restrict__ pos4, double4 *__restrict__ vel4) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < N; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = perm[s];
        pos4[s] = make_double4(pos[i], pos[i + N], pos[i + 2 * N], 0.0);
        vel4[s] = make_double4(vel[i], vel[i + N], vel[i + 2 * N], kent ? kent[i] : 0.0);
    }
}

// leading octant levels shared by sorted slots a and b (0..42); word 1 is only read (and only valid) when word 0 agrees
__device__ __forceinline__ int common_levels2(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ klo, int64_t a, int64_t b) {
    const int c = sph_common_levels(keys[a], keys[b]);
    if (c < SPH_KEY_LEVELS) return c;
    return SPH_KEY_LEVELS + sph_common_levels(klo[a], klo[b]);
}

// Runs of particles that share key word 0 (closer than l / 2^21 per axis): order each run by word 1, then by particle
// id, in place (insertion sort by the thread of the run's first slot; the runs are a handful of particles).
__global__ void __launch_bounds__(TB) key_runs_kernel(const double *__restrict__ pos, int64_t N, const uint64_t *__restrict__ keys,
                                                       const unsigned long long *__restrict__ scal, int *__restrict__ perm,
                                                       uint64_t *__restrict__ klo) {
    const double l = __longlong_as_double((long long)scal[SC_LDOM]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        if (i > 0 && keys[i - 1] == k) continue;            // not the first slot of a run
        if (i + 1 >= N || keys[i + 1] != k) continue;       // a run of one: word 1 is never consulted
        int64_t e = i + 2;
        while (e < N && keys[e] == k) ++e;
        for (int64_t j = i; j < e; ++j) {
            const int p = perm[j];
            const uint64_t lo = sph_octant_key_word(pos[p], pos[p + N], pos[p + 2 * N], l, 1);
            int64_t t = j;
            while (t > i && (klo[t - 1] > lo || (klo[t - 1] == lo && perm[t - 1] > p))) {
                klo[t] = klo[t - 1];
                perm[t] = perm[t - 1];
                --t;
            }
            klo[t] = lo;
            perm[t] = p;
        }
    }
}

// ---- tree: per-particle node counts ---------------------------------------------------------------
__global__ void __launch_bounds__(TB) node_count_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ klo, int64_t N,
                                                         int *__restrict__ cnt, unsigned long long *__restrict__ scal) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = i > 0 ? common_levels2(keys, klo, i - 1, i) : -1;
        const int b = i + 1 < N ? common_levels2(keys, klo, i, i + 1) : -1;
        if (a >= SPH_LEVELS || b >= SPH_LEVELS) atomicOr(&scal[SC_ERR], (unsigned long long)ERRF_DEPTH);
        const int mx = a > b ? a : b;
        cnt[i] = mx - a + 1;  // depths a+1 .. mx+1
    }
}

// emits the (start, depth) node list in particle order; depth doubles as the 8-bit sort key
__global__ void __launch_bounds__(TB) node_emit_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ klo, int64_t N,
                                                        const int *__restrict__ cnt, const int *__restrict__ base,
                                                        int64_t cap, int *__restrict__ old_start,
                                                        int *__restrict__ old_depth, uint64_t *__restrict__ dkey,
                                                        int *__restrict__ dval, unsigned long long *__restrict__ scal) {
    const int64_t total = base[N];
    // identical keys (flagged by node_count_kernel, a previous launch) make the node list meaningless
    const bool bad = total > cap || (scal[SC_ERR] & (unsigned long long)ERRF_DEPTH) != 0ull;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[SC_NNODES] = (unsigned long long)(bad ? 0 : total);
        if (total > cap) atomicOr(&scal[SC_ERR], (unsigned long long)ERRF_NODES);
    }
    if (bad) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = i > 0 ? common_levels2(keys, klo, i - 1, i) : -1;
        const int n = cnt[i];
        const int o = base[i];
        for (int k = 0; k < n; ++k) {
            old_start[o + k] = (int)i;
            old_depth[o + k] = a + 1 + k;
            dkey[o + k] = (uint64_t)(a + 1 + k);
            dval[o + k] = o + k;
        }
    }
}

__global__ void __launch_bounds__(TB) level_init_kernel(int *__restrict__ level_start,
                                                         const unsigned long long *__restrict__ scal) {
    if (threadIdx.x < SPH_LEVELS + 3) level_start[threadIdx.x] = (int)scal[SC_NNODES];
}

__global__ void __launch_bounds__(TB) node_inverse_kernel(const int *__restrict__ bfs_old, const int *__restrict__ old_depth,
                                                           const unsigned long long *__restrict__ scal,
                                                           int *__restrict__ bfs_of_old, int *__restrict__ level_start) {
    const int64_t M = (int64_t)scal[SC_NNODES];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < M; k += (int64_t)gridDim.x * blockDim.x) {
        const int o = bfs_old[k];
        bfs_of_old[o] = (int)k;
        const int d = old_depth[o];
        if (k == 0 || old_depth[bfs_old[k - 1]] != d) level_start[d] = (int)k;
    }
}

// first index in [lo, hi) whose key is >= v
__device__ __forceinline__ int lower_bound_key(const uint64_t *__restrict__ keys, int lo, int hi, uint64_t v) {
    while (lo < hi) {
        const int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
        if (keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one thread per node (BFS id k): particle range, children, geometry, leaf payload
__global__ void __launch_bounds__(TB) node_build_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ klo, int64_t N,
                                                         const int *__restrict__ cnt, const int *__restrict__ base,
                                                         const int *__restrict__ bfs_old, const int *__restrict__ old_start,
                                                         const int *__restrict__ old_depth, const int *__restrict__ bfs_of_old,
                                                         const double4 *__restrict__ pos4, double mass,
                                                         const unsigned long long *__restrict__ scal, SphTree t) {
    const int64_t M = (int64_t)scal[SC_NNODES];
    const double l = __longlong_as_double((long long)scal[SC_LDOM]);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < M; k += (int64_t)gridDim.x * blockDim.x) {
        const int o = bfs_old[k];
        const int s = old_start[o];
        const int d = old_depth[o];
        const uint64_t key = keys[s];
        const int dleaf = (base[s + 1] - base[s]) + (d - (o - base[s])) - 1;  // deepest depth starting at s
        const bool leaf = d == dleaf;
        (void)cnt;
        const uint64_t lo = d > SPH_KEY_LEVELS ? klo[s] : 0ull;   // word 1: cells below level 21 lie inside a run of equal word 0
        int e;  // end of the particle range
        if (d == 0) e = (int)N;
        else if (leaf) e = s + 1;
        else if (d <= SPH_KEY_LEVELS) {
            const int sh = 3 * (SPH_KEY_LEVELS - d);
            const uint64_t next_prefix = ((key >> sh) + 1ull) << sh;  // cannot overflow 64 bits: key < 2^63
            e = lower_bound_key(keys, s + 1, (int)N, next_prefix);
        } else {
            const int re = lower_bound_key(keys, s + 1, (int)N, key + 1ull);     // end of the run that shares word 0
            const int sh = 3 * (2 * SPH_KEY_LEVELS - d);
            e = lower_bound_key(klo, s + 1, re, ((lo >> sh) + 1ull) << sh);
        }
        t.nstart[k] = s;
        t.ncount[k] = e - s;
        t.ndepth[k] = d;
        const SphCell g = sph_cell_of(key, lo, d, l);
        t.nodeB[k] = make_double4(g.lo[0], g.lo[1], g.lo[2], g.hi[0]);
        const double s2 = (g.L * 2) * (g.L * 2);  // s = node.Length*2 ; s^2  (F/gravOctree_Single.jl:257,265)
        t.nodeC[k] = make_double4(g.hi[1], g.hi[2], s2, g.L);
        int2 I;
        if (leaf) {
            I = make_int2(s, 0);
            t.nodeI[k] = I;
            t.leaf_of[s] = (int)k;
            const double4 p = pos4[s];
            t.nodeA[k] = make_double4(p.x, p.y, p.z, mass);  // leaf: rCOM = particle, Mass = m (:186-194, :69)
        } else {
            // non-empty octants of this cell = its children, contiguous in BFS order starting at the
            // depth-(d+1) cell that starts at the same particle; their digits are level d of the path
            const bool w1 = d >= SPH_KEY_LEVELS;
            const uint64_t *kw = w1 ? klo : keys;
            const uint64_t kv = w1 ? klo[s] : key;
            const int dl = w1 ? d - SPH_KEY_LEVELS : d;               // level inside the word
            const int sh = 3 * (SPH_KEY_LEVELS - 1 - dl);
            const uint64_t pre = (dl == 0) ? 0ull : ((kv >> (sh + 3)) << 3);
            int nch = 0, lo_i = s, leafmask = 0;
            for (int c = 1; c <= 8 && lo_i < e; ++c) {
                const int nb = (c == 8) ? e : lower_bound_key(kw, lo_i, e, (pre + (uint64_t)c) << sh);
                if (nb > lo_i) {
                    if (nb - lo_i == 1) leafmask |= 1 << nch;   // child holds one particle = leaf
                    ++nch;
                }
                lo_i = nb;
            }
            const int fc = bfs_of_old[o + 1];
            I = make_int2(fc, nch | (leafmask << 8));
            t.nodeI[k] = I;
            for (int c = 0; c < nch; ++c) t.parent[fc + c] = (int)k;
            if (k == 0) t.parent[0] = -1;
        }
        // compact copy for the 4-target search (box + children + range in two adjacent sectors) and the walk's exact clause 2
        t.nodeBC[2 * k] = make_double4(g.lo[0], g.lo[1], g.lo[2], g.hi[0]);
        t.nodeBC[2 * k + 1] = make_double4(g.hi[1], g.hi[2], pack_i2(I.x, I.y), pack_i2(s, e - s));
    }
}

// Mass / rCOM / cell radius of one internal node from its (finished) children, children in octant order
// (setCOMs!, F/gravOctree_Single.jl:197-208).  No FMA: the reference rounds each product.
__device__ __forceinline__ void com_of_children(const SphTree &t, int k, int2 I) {
    double tm = 0.0, wx = 0.0, wy = 0.0, wz = 0.0;
    for (int c = 0; c < (I.y & 0xff); ++c) {
        const double2 *ap = reinterpret_cast<const double2 *>(&t.nodeA[I.x + c]);   // L2 reads: written by other SMs in this launch
        const double2 a0 = __ldcg(ap), a1 = __ldcg(ap + 1);
        const double4 A = make_double4(a0.x, a0.y, a1.x, a1.y);
        tm = __dadd_rn(tm, A.w);
        wx = __dadd_rn(wx, __dmul_rn(A.w, A.x));
        wy = __dadd_rn(wy, __dmul_rn(A.w, A.y));
        wz = __dadd_rn(wz, __dmul_rn(A.w, A.z));
    }
    const double cx = wx / tm, cy = wy / tm, cz = wz / tm;
    t.nodeA[k] = make_double4(cx, cy, cz, tm);
    // radius of the cell about its COM (upper bound): lets the walk prove clause 2 of the acceptance test
    // (h_i^2 / mindist^2 < 0.25) from d alone, since mindist >= d - radius
    const double4 B = t.nodeB[k];
    const double4 C = t.nodeC[k];
    const double rx = fmax(cx - B.x, B.w - cx), ry = fmax(cy - B.y, C.x - cy), rz = fmax(cz - B.z, C.y - cz);
    t.nodeD[k] = make_double2(C.z, sqrt(rx * rx + ry * ry + rz * rz) * (1.0 + 1e-12));
}

// Bottom-up sweep in ONE launch (replaces a launch per level): every leaf climbs towards the root; the thread that
// completes a cell's last child (atomic arrival counter) computes the cell from its children and keeps climbing.
// Results do not depend on the arrival order: a cell is always summed over all of its children in octant order.
__global__ void __launch_bounds__(TB) com_bottomup_kernel(SphTree t, const unsigned long long *__restrict__ scal) {
    if (scal[SC_ERR] != 0ull) return;
    const int64_t M = (int64_t)scal[SC_NNODES];
    for (int64_t k0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k0 < M; k0 += (int64_t)gridDim.x * blockDim.x) {
        if (t.nodeI[k0].y != 0) continue;          // start at the leaves only
        int k = (int)k0;
        for (;;) {
            const int p = t.parent[k];
            if (p < 0) break;
            const int2 I = t.nodeI[p];
            __threadfence();                                   // my cell's data before my arrival
            const int old = atomicAdd(&t.arrive[p], 1);
            if (old != (I.y & 0xff) - 1) break;                // siblings still pending: their last one continues
            __threadfence();
            com_of_children(t, p, I);
            k = p;
        }
    }
}

}  // namespace

cudaError_t sph_launch_domain_keys(sph_handle *h, const double *pos) {
    cudaStream_t st = h->stream;
    cudaMemsetAsync(h->scal, 0, sizeof(unsigned long long) * SC_RESET, st);
    sph_note(2);
    absmax_kernel<<<grid_for(3 * h->N), TB, 0, st>>>(pos, 3 * h->N, h->scal);
    keys_kernel<<<grid_for(h->N), TB, 0, st>>>(pos, h->N, h->scal, h->keys_alt, h->perm_alt);
    cudaError_t e = sph_sort_pairs(h->keys_alt, h->perm_alt, h->keys, h->perm, h->N, nullptr, 0, 64, h->sort_tmp,
                                   h->sort_tmp_bytes, st);
    if (e != cudaSuccess) return e;
    sph_note(1);
    key_runs_kernel<<<grid_for(h->N), TB, 0, st>>>(pos, h->N, h->keys, h->scal, h->perm, h->klo);
    return cudaGetLastError();
}

cudaError_t sph_launch_permute(sph_handle *h, const double *pos, const double *vel, const double *kent) {
    sph_note(1);
    permute_kernel<<<grid_for(h->N), TB, 0, h->stream>>>(pos, vel, kent, h->perm, h->N, h->pos4, h->vel4);
    return cudaGetLastError();
}

cudaError_t sph_launch_tree(sph_handle *h) {
    cudaStream_t st = h->stream;
    SphTree &t = h->tree;
    const int64_t N = h->N;
    sph_note(5);
    node_count_kernel<<<grid_for(N), TB, 0, st>>>(h->keys, h->klo, N, h->cnt, h->scal);
    cudaError_t e = sph_exclusive_scan(h->cnt, h->base, N, h->sort_tmp, h->sort_tmp_bytes, st);
    if (e != cudaSuccess) return e;
    node_emit_kernel<<<grid_for(N), TB, 0, st>>>(h->keys, h->klo, N, h->cnt, h->base, t.cap, t.old_start, t.old_depth,
                                                  t.dkey_in, t.dval_in, h->scal);
    level_init_kernel<<<1, TB, 0, st>>>(t.level_start, h->scal);
    // stable counting sort by depth (one 8-bit pass) = breadth-first node order of build_octree!
    e = sph_sort_pairs(t.dkey_in, t.dval_in, t.dkey_out, t.dval_out, t.cap, h->scal + SC_NNODES, 0, 8,
                       h->sort_tmp, h->sort_tmp_bytes, st);
    if (e != cudaSuccess) return e;
    node_inverse_kernel<<<grid_for(t.cap), TB, 0, st>>>(t.dval_out, t.old_depth, h->scal, t.bfs_of_old, t.level_start);
    node_build_kernel<<<grid_for(t.cap), TB, 0, st>>>(h->keys, h->klo, N, h->cnt, h->base, t.dval_out, t.old_start,
                                                       t.old_depth, t.bfs_of_old, h->pos4, h->p.m, h->scal, t);
    return cudaGetLastError();
}

// setCOMs! (F/gravOctree_Single.jl:183-211).  A launch of its own because only the walk reads Mass / rCOM: the
// evaluation runs it on the second stream beside the neighbour search (sph_api.cu:eval_internal).
cudaError_t sph_launch_com(sph_handle *h) {
    cudaStream_t st = h->stream;
    SphTree &t = h->tree;
    sph_note(1);
    cudaMemsetAsync(t.arrive, 0, sizeof(int) * (size_t)t.cap, st);
    com_bottomup_kernel<<<grid_for(t.cap), TB, 0, st>>>(t, h->scal);
    return cudaGetLastError();
}
