// sph_internal.cuh -- shared declarations of libsph_b200.so (not part of the public ABI).
//
// Device data layout (per handle, all FP64 unless noted; "sorted" = octant-key order of the
// CURRENT evaluation, "orig" = the caller's particle order):
//   state   pos[3N] vel[3N] kent[N] acc[3N]                      orig, column-major like the Julia matrices
//   sort    keys[N] u64 (21 levels x 3 bits), perm[N] i32        perm[s] = orig id of sorted slot s
//   sorted  pos4[N] double4 {x,y,z,d2k = (2h)^2}, vel4[N] double4 {vx,vy,vz,K_i}, hr[N] double2 {h,rho},
//           force records fa[N] double4 {x,y,z,h}, fb[N] double4 {vx,vy,vz,rho}, fc[N] double2 {P/rho^2,c}
//   tree    BFS-ordered linear octree: nodeI int2 {first child | particle, nchild | leafmask << 8 (0 = leaf)},
//           nodeA double4 {com, mass}, nodeB double4 {lo.xyz, hi.x}, nodeC double4 {hi.y, hi.z, (2L)^2, L}, nodeD double2 {(2L)^2, max |corner - com|},
//           nstart/ncount i32 particle range of the node in the sorted arrays
//   lists   nbr[K x NL] i32 column-major (NL = N rounded up to 128), sorted-space rows and entries (0-based), d2k[N],
//           kid[N] (tie id of the K-th entry); extras ext[SPH_ECAP x NL] + ecnt[N]: reverse partners outside the own list
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/sph_b200.h"

#define SPH_KEY_LEVELS 21       // octant levels held by one 63-bit key word
#define SPH_LEVELS 42           // deepest cell: two key words (the second one only for particles that share the first)
#define SPH_MAX_RANKS 16
#define SPH_WALK_REC 2           // double4 per walk record of the octree (two nodes per 128-byte line)
#define SPH_ECAP 64              // reverse partners per particle held in the extras table (more: overflow list)
#define SPH_TILE_DEFAULT 0       // density / force: 1 = shared-memory tile kernels (TMA-staged), 0 = direct gathers
#define SPH_WALK_DEAL 16         // walk tiles (128 targets) are dealt to the ranks in groups of 16 consecutive tiles:
                                 // round-robin balances the load, consecutive tiles keep the tree nodes hot in L2

#define HD __host__ __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// Octant path key.  Replays the child-centre recurrence of addNodes! (F/gravOctree_Single.jl:110-126,
// :143-148): child_l = parent_l/2, child centre = parent centre -/+ child_l, octant bit = (x - c) > 0.
// Only additions, subtractions and halvings: bit-identical to the Julia evaluation, so a particle
// on a cell boundary lands in the same child as in the reference.
// Word 0 holds levels 0..20, word 1 levels 21..41.  The sort uses word 0; word 1 is computed only for the (rare)
// particles that share all of word 0 with a sorted neighbour and orders them inside that run (tree.cu), so trees
// deeper than 21 levels - which build_octree! produces whenever two particles are closer than l / 2^21 per axis
// (:213-227 subdivides until every leaf holds one particle) - are built like any other.
// ---------------------------------------------------------------------------------------------
HD uint64_t sph_octant_key_word(double x, double y, double z, double l, int word) {
    double cx = 0.0, cy = 0.0, cz = 0.0, L = l;
    uint64_t key = 0;
    const int last = SPH_KEY_LEVELS * (word + 1);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int lev = 0; lev < last; ++lev) {
        const double cl = L / 2;
        const unsigned ox = (x - cx) > 0, oy = (y - cy) > 0, oz = (z - cz) > 0;
        key = (key << 3) | (uint64_t)((oz << 2) | (oy << 1) | ox);      // the bits of earlier words shift out
        cx = ox ? cx + cl : cx - cl;
        cy = oy ? cy + cl : cy - cl;
        cz = oz ? cz + cl : cz - cl;
        L = cl;
    }
    return key & 0x7fffffffffffffffull;
}
HD uint64_t sph_octant_key(double x, double y, double z, double l) { return sph_octant_key_word(x, y, z, l, 0); }

// Geometry of the depth-d cell on the path (khi, klo): centre, bounds and half-width exactly as
// addNodes! produces them (F/gravOctree_Single.jl:110-140): for a parent (pc, pl): cl = pl/2,
// lc = pc-cl, rc = pc+cl, mn = lc-cl, ctr = lc+cl, mx = rc+cl; low child = {lc,[mn,ctr]}, high = {rc,[ctr,mx]}.
struct SphCell {
    double c[3], lo[3], hi[3], L;
};
HD SphCell sph_cell_of(uint64_t khi, uint64_t klo, int depth, double l) {
    SphCell g;
    g.L = l;
    for (int a = 0; a < 3; ++a) { g.c[a] = 0.0; g.lo[a] = -l; g.hi[a] = l; }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int lev = 0; lev < depth; ++lev) {
        const double cl = g.L / 2;
        const unsigned oct = lev < SPH_KEY_LEVELS ? (unsigned)(khi >> (3 * (SPH_KEY_LEVELS - 1 - lev))) & 7u
                                                  : (unsigned)(klo >> (3 * (2 * SPH_KEY_LEVELS - 1 - lev))) & 7u;
        for (int a = 0; a < 3; ++a) {
            const double lc = g.c[a] - cl, rc = g.c[a] + cl;
            const double mn = lc - cl, ctr = lc + cl, mx = rc + cl;
            if ((oct >> a) & 1u) { g.c[a] = rc; g.lo[a] = ctr; g.hi[a] = mx; }
            else                 { g.c[a] = lc; g.lo[a] = mn;  g.hi[a] = ctr; }
        }
        g.L = cl;
    }
    return g;
}

// number of leading octant levels two key words share (21 = identical words)
HD int sph_common_levels(uint64_t a, uint64_t b) {
    const uint64_t x = a ^ b;
    if (x == 0) return SPH_KEY_LEVELS;
#ifdef __CUDA_ARCH__
    const int lz = __clzll((long long)x);
#else
    const int lz = __builtin_clzll(x);
#endif
    return (lz - 1) / 3;
}

#ifdef __CUDACC__
// 1/sqrt(x) and 1/x from the hardware seed (MUFU.RSQ64H / RCP64H read the high word of x: ~2^-21 relative) plus one
// Newton step each: relative error <= 4e-13 - the parity bars are 1e-9 (hydro sums) and 1e-6 (gravity) - in five
// and three instructions instead of the IEEE sequences with their slow-path calls.
// x must be a normal positive number (squared distances / smoothing lengths of distinct particles are).
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);        // 1 - x y^2
    return fma(y * e, 0.5, y);                     // y (1 + e/2)
}
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, e, y);                           // y (1 + e)
}

// two ints in the bit pattern of a double (node records keep their integer fields next to the FP64 ones)
__device__ __forceinline__ int2 unpack_i2(double v) {
    const long long b = __double_as_longlong(v);
    return make_int2((int)(b & 0xffffffffLL), (int)(b >> 32));
}
__device__ __forceinline__ double pack_i2(int x, int y) {
    return __longlong_as_double((long long)(unsigned)x | ((long long)y << 32));
}

// Squared distance exactly as NearestNeighbors' Euclidean metric evaluates it in the oracle's
// restatement: (dx*dx + dy*dy) + dz*dz with every product and sum rounded (never contracted to FMA).
__device__ __forceinline__ double sph_d2_exact(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
#endif

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
enum { PH_SORT = 0, PH_TREE, PH_KNN, PH_DENSITY, PH_FORCE, PH_GRAV, PH_FINISH, PH_COUNT };

struct SphTree {
    int64_t cap = 0;  // node capacity
    int2 *nodeI = nullptr;
    double4 *nodeA = nullptr, *nodeB = nullptr, *nodeC = nullptr;
    double4 *nodeW = nullptr;  // walk records, SPH_WALK_REC x double4 per node: {com.xyz, mass | h_j}, {(2L)^2, radius, bits{first|slot, nch|leafmask<<8}, bits{nstart, ncount}}, nodeB, nodeC
    double4 *nodeBC = nullptr; // compact search / clause-2 record of node k: [2k] {lo.xyz, hi.x}, [2k+1] {hi.y, hi.z, bits{first child | slot, nchild | leafmask << 8}, bits{nstart, ncount}}
    double2 *nodeD = nullptr;  // internal nodes: {(2 Length)^2, upper bound of the distance from rCOM to any point of the cell}
    int *nstart = nullptr, *ncount = nullptr, *ndepth = nullptr;
    int *parent = nullptr, *arrive = nullptr;   // bottom-up COM sweep: parent id, number of finished children
    int *leaf_of = nullptr;                      // leaf node of every sorted slot
    // build scratch
    int *old_start = nullptr, *old_depth = nullptr;  // node list in (start, depth) order
    uint64_t *dkey_in = nullptr, *dkey_out = nullptr;
    int *dval_in = nullptr, *dval_out = nullptr;      // bfs -> old
    int *bfs_of_old = nullptr;
    int *level_start = nullptr;                       // [SPH_LEVELS + 3]
};

struct sph_handle {
    sph_params p{};
    int64_t N = 0;
    int64_t NS = 0;  // stride of the sorted-space component arrays = nranks * chunk (chunk: targets per rank, a multiple of 128)
    int64_t NS_alloc = 0;  // allocation of those arrays: covers every rank count up to SPH_MAX_RANKS
    int64_t NL = 0;  // row stride of the neighbour lists and the extras table: N rounded up to 128
    int K = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    bool have_state = false, have_eval = false;
    double t = 0.0;

    // state (orig order)
    double *pos = nullptr, *vel = nullptr, *kent = nullptr, *acc = nullptr;
    double *pos_half = nullptr, *vel_half = nullptr;
    // staging for sph_eval_acc
    double *in_pos = nullptr, *in_vel = nullptr, *in_kent = nullptr, *in_acc = nullptr;
    const double *last_acc = nullptr;  // acceleration array written by the last evaluation
    bool lists_valid = false;          // neighbour lists match the current sort
    bool outputs_fresh = false;        // o_* arrays (other than o_h) hold the last evaluation's results
    bool hint_valid = false;           // o_h holds smoothing lengths of a completed evaluation (search radius hint)
    bool no_hint = false;              // SPH_B200_NO_HINT: always search from the guaranteed radius
    // per-evaluation outputs, orig order
    double *o_rho = nullptr, *o_h = nullptr, *o_phi = nullptr, *o_sumvdw = nullptr, *o_mumax = nullptr,
           *o_cs = nullptr, *o_dkdt = nullptr, *o_ahyd = nullptr, *o_g = nullptr;
    // sort
    uint64_t *keys = nullptr, *keys_alt = nullptr;
    uint64_t *klo = nullptr;    // second key word, valid for particles that share word 0 with a sorted neighbour
    int *perm = nullptr, *perm_alt = nullptr;
    void *sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    // sorted working set
    double4 *pos4 = nullptr, *vel4 = nullptr;
    double2 *hr = nullptr;      // {h, rho}
    double4 *fa = nullptr, *fb = nullptr;   // {x, y, z, h}, {vx, vy, vz, rho}: what the force pass gathers per neighbour
    double2 *fc = nullptr;                  // {P/rho^2, c} (gathered by the polytropic force only)
    double *rho_s = nullptr;    // density of the owned targets (all-gathered in multi-GPU runs)
    double *hs = nullptr;       // smoothing length of every particle, set right after the search (the walk starts from it)
    double *d2k = nullptr;      // K-th squared distance
    int *kid = nullptr;         // caller's id of the K-th list entry when the K-th distance is tied, else INT_MAX
    int *nbr = nullptr;         // K x NL
    // reverse partners outside the own list (hydro.cu): table, overflow list, pairs for / from other ranks
    int *ecnt = nullptr, *ext = nullptr;
    int2 *ovf = nullptr, *outbox = nullptr, *inbox = nullptr;
    int ecap = SPH_ECAP;
    int64_t ovcap = 0, obcap = 0;
    // one buffer [6][NS]: a_hyd x,y,z, dK/dt sum, sum_vdw, mumax; every rank writes its own targets and the six
    // component arrays are all-gathered in multi-GPU runs
    double *s_red = nullptr;
    double *s_ahyd = nullptr, *s_dkdt = nullptr, *s_sumvdw = nullptr, *s_mumax = nullptr;   // views into s_red
    // walk results [rank][4][walk_chunk]: tiles of 128 targets are dealt round-robin to the ranks (load balance), each
    // rank writes its tiles compactly -> a single in-place all-gather
    double *walk_buf = nullptr;
    double *walk_part = nullptr;   // [8 root children][4][walk_chunk] partial walk results of this rank
    int64_t walk_chunk = 0;
    int *cnt = nullptr, *base = nullptr;  // per-particle node counts / offsets (N+1)
    SphTree tree;
    // device scalars: [0] l_domain bits, [1] n_nodes, [2] error flags, [3] dt bits, [4] visits
    unsigned long long *scal = nullptr;
    unsigned long long *h_scal = nullptr;  // pinned mirror
    double *stat_dev = nullptr;            // 32 doubles: t, dt, reduction results, stats row (integrate.cu)
    double *h_stat = nullptr;              // pinned mirror
    double *red_partial = nullptr;         // per-block partial sums of the statistics kernels
    double *log_dev = nullptr, *h_log = nullptr;   // step log {dt, stats row} of sph_step: device + pinned mirror
    size_t log_cap = 0;                            // steps
    cudaGraphExec_t step_graph = nullptr;          // one captured step (small N: launch-latency bound), see sph_step
    long long graph_launches = 0;                  // kernels in it
    void *scratch = nullptr;                       // grow-only scratch of the getters / density_at
    size_t scratch_bytes = 0;
    // timing
    cudaEvent_t ev[PH_COUNT + 1]{};
    cudaEvent_t cev[8]{};   // begin/end of the four collectives of an evaluation
    cudaEvent_t wev[2]{};   // the walk kernel alone
    // the force kernel (+ its all-reduce) runs on a second stream, concurrently with the tree walk: both only need
    // the density/EOS results, and the latency-bound force kernel fills the issue slots the walk's tail leaves idle
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_tree = nullptr, ev_com = nullptr, fev[2]{}, dev[2]{};   // fev / dev: force and density phases on the stream they ran on
    void *nccl2 = nullptr;  // communicator of stream2 (NCCL calls of one communicator must not run concurrently)
    bool overlap = true;
    bool ev_valid = false;
    // multi-GPU
    int nranks = 1, rank = 0;
    void *nccl = nullptr;
    int64_t chunk = 0;  // targets per rank (padded)
};

// scal[0..SC_RESET) is cleared at the start of every force evaluation; SC_STICKY accumulates error flags
enum { SC_LDOM = 0, SC_NNODES, SC_ERR, SC_DT, SC_VISITS, SC_KNN_RETRY, SC_OVF, SC_OUTBOX, SC_RESET = 15, SC_STICKY = 15, SC_COUNT = 16 };
// ERRF_STACK is transient: the regular walk raises it, the DEEP walk variant answers and clears it (gravity.cu).  Kernels
// that run beside the walk (density, force on the second stream) must not mistake it for a failed evaluation.
#define SPH_ERR_BLOCKING(scal) ((scal)[SC_ERR] & ~(unsigned long long)ERRF_STACK)
enum { ERRF_DEPTH = 1, ERRF_NODES = 2, ERRF_STACK = 4, ERRF_NAN = 8, ERRF_EXTRAS = 16, ERRF_HALO = 32, ERRF_STACK2 = 64 };

int sph_fail(sph_handle *h, int code, const std::string &msg);
// cumulative count of kernel launches issued by the library in this process (bench.py's gpu_launches)
extern long long g_sph_launches;
inline void sph_note(int n) { g_sph_launches += n; }
#define SPH_CUDA(h, call)                                                                            \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return sph_fail((h), SPH_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// ---- radix_sort.cu ------------------------------------------------------------------------------
size_t sph_sort_temp_bytes(int64_t n);
// Stable LSD radix sort of (key, value) pairs on bits [begin_bit, end_bit).  Result lands in
// keys_out/vals_out (ping-pongs through the *_in buffers, which are clobbered).
cudaError_t sph_sort_pairs(uint64_t *keys_in, int *vals_in, uint64_t *keys_out, int *vals_out, int64_t n,
                           const unsigned long long *n_dev /* optional device-side n (<= n) */,
                           int begin_bit, int end_bit, void *temp, size_t temp_bytes, cudaStream_t st);
// exclusive scan of n ints; out[n] = total (out has n+1 entries)
cudaError_t sph_exclusive_scan(const int *in, int *out, int64_t n, void *temp, size_t temp_bytes,
                               cudaStream_t st);

// ---- tree.cu -------------------------------------------------------------------------------------
cudaError_t sph_launch_domain_keys(sph_handle *h, const double *pos);
cudaError_t sph_launch_permute(sph_handle *h, const double *pos, const double *vel, const double *kent);
cudaError_t sph_launch_tree(sph_handle *h);      // node table (ranges, children, geometry): all the search needs
cudaError_t sph_launch_com(sph_handle *h);       // Mass / rCOM bottom-up (setCOMs!): needed by the walk only

// ---- knn.cu --------------------------------------------------------------------------------------
cudaError_t sph_launch_knn(sph_handle *h, int64_t t0, int64_t t1);
cudaError_t sph_launch_export_neighbors(sph_handle *h, int *idx_out_dev, double *r_out_dev);
cudaError_t sph_launch_knn_points(sph_handle *h, const double *pts_dev, int64_t M, double *d2s_dev /* M x K scratch */, double *rho_out_dev);

// ---- hydro.cu ------------------------------------------------------------------------------------
cudaError_t sph_launch_smoothing(sph_handle *h);                          // pos4.w = d2k of ALL particles, extras counters cleared
cudaError_t sph_launch_density(sph_handle *h, int64_t t0, int64_t t1, bool with_eos);   // rho (+ EOS) and extras of [t0, t1)
cudaError_t sph_launch_outbox_header(sph_handle *h);
cudaError_t sph_launch_extras_merge(sph_handle *h, int64_t t0, int64_t t1);
cudaError_t sph_launch_extras_sort(sph_handle *h, int64_t t0, int64_t t1);     // fixed (ascending) order of every particle's extras
cudaError_t sph_launch_eos(sph_handle *h);                                 // several ranks: hr and the force records of ALL particles
cudaError_t sph_launch_force(sph_handle *h, int64_t t0, int64_t t1);

// ---- gravity.cu ----------------------------------------------------------------------------------
cudaError_t sph_launch_walk(sph_handle *h);

// ---- integrate.cu --------------------------------------------------------------------------------
cudaError_t sph_launch_finish(sph_handle *h, double *acc_out);
cudaError_t sph_launch_dt(sph_handle *h);
cudaError_t sph_launch_unpermute(sph_handle *h);
cudaError_t sph_launch_stats(sph_handle *h, double *log_row_dev);
cudaError_t sph_launch_set_time(sph_handle *h, double t);
cudaError_t sph_launch_predict(sph_handle *h);
cudaError_t sph_launch_correct(sph_handle *h);
cudaError_t sph_launch_evolve_k(sph_handle *h);
