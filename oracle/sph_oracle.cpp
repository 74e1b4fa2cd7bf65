// =============================================================================
// sph_oracle.cpp -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A C++17, FP64, no-FMA-contraction restatement of the per-step SPH core of the
// reference engine  /root/reference/julia_version/fastv1_kd&single_oc/  (below "F/").
// It exists to CHECK the CUDA path (tests/, __graft_entry__.smoke(), and the
// cpu_baseline / --impl reference legs of bench.py).  Nothing in the product path
// (astrophysical-sph_b200/) may import, link or call it.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
//  and Julia is not available in the build container, so
// this restatement cannot be pinned against outputs of the reference itself.  It is
// instead cross-checked against an independent numpy/scipy twin (oracle/sph_numpy.py),
// brute-force kNN, and direct-sum gravity (tests/test_oracle.py).
//
// Conventions (identical to the Julia side):
//   * matrices are column-major: pos = [x_0..x_{N-1}, y_0.., z_0..]  (N x 3)
//   * neighbour indices are 1-based Int32, N x K column-major, column 1 = self
//   * every expression is evaluated in the reference's association order;
//     compile with -O2 -ffp-contract=off (see oracle/Makefile).
//   * third-party arithmetic: NearestNeighbors.jl (unpinned, not vendored in the
//     reference) supplies exact Euclidean kNN; restated here as an exact KD-tree
//     search with d2 = (dx*dx + dy*dy) + dz*dz, r = sqrt(d2), ascending (d2, index).
//     Tie order among equal distances is a documented choice of this oracle
//     (smaller index first); nothing in the reference pins it.
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <limits>
#include <numeric>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double PI = 3.141592653589793;  // Float64(pi)

thread_local std::string g_err;

// -----------------------------------------------------------------------------
// Exact kNN  (replaces NearestNeighbors.KDTree / knn,
//             call sites F/isothermal_hydroKDTree.jl:128,131)
// -----------------------------------------------------------------------------
struct KDTree {
    int N = 0;
    const double *x = nullptr, *y = nullptr, *z = nullptr;
    std::vector<int32_t> perm;  // point ids in tree order
    struct Node {
        int lo, hi;      // range in perm
        int left, right; // children (-1 = leaf)
        int dim;
        double split;
    };
    std::vector<Node> nodes;
    static constexpr int LEAF = 12;

    double coord(int id, int d) const { return d == 0 ? x[id] : (d == 1 ? y[id] : z[id]); }

    int build(int lo, int hi) {
        int me = (int)nodes.size();
        nodes.push_back({lo, hi, -1, -1, 0, 0.0});
        if (hi - lo <= LEAF) return me;
        double mn[3], mx[3];
        for (int d = 0; d < 3; d++) { mn[d] = 1e300; mx[d] = -1e300; }
        for (int i = lo; i < hi; i++)
            for (int d = 0; d < 3; d++) {
                double c = coord(perm[i], d);
                mn[d] = std::min(mn[d], c);
                mx[d] = std::max(mx[d], c);
            }
        int dim = 0;
        for (int d = 1; d < 3; d++)
            if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
        if (!(mx[dim] > mn[dim])) return me;  // all coincident: keep as (big) leaf
        int mid = (lo + hi) / 2;
        std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                         [&](int32_t a, int32_t b) { return coord(a, dim) < coord(b, dim); });
        double split = coord(perm[mid], dim);
        int l = build(lo, mid);
        int r = build(mid, hi);
        nodes[me].left = l;
        nodes[me].right = r;
        nodes[me].dim = dim;
        nodes[me].split = split;
        return me;
    }

    void init(int n, const double *pos) {
        N = n; x = pos; y = pos + n; z = pos + 2 * (size_t)n;
        perm.resize(n);
        std::iota(perm.begin(), perm.end(), 0);
        nodes.clear();
        nodes.reserve(2 * (size_t)n / LEAF + 16);
        if (n > 0) build(0, n);
    }
};

struct Cand {
    double d2;
    int32_t id;
};
inline bool cand_less(const Cand &a, const Cand &b) {
    return a.d2 < b.d2 || (a.d2 == b.d2 && a.id < b.id);
}

struct KnnQuery {
    const KDTree &t;
    int K;
    double qx, qy, qz;
    std::vector<Cand> heap;  // max-heap on (d2,id)
    KnnQuery(const KDTree &tt, int k) : t(tt), K(k) { heap.reserve(k + 1); }

    void offer(int id) {
        double dx = qx - t.x[id], dy = qy - t.y[id], dz = qz - t.z[id];
        double d2 = (dx * dx + dy * dy) + dz * dz;
        Cand c{d2, id};
        if ((int)heap.size() < K) {
            heap.push_back(c);
            std::push_heap(heap.begin(), heap.end(), cand_less);
        } else if (cand_less(c, heap.front())) {
            std::pop_heap(heap.begin(), heap.end(), cand_less);
            heap.back() = c;
            std::push_heap(heap.begin(), heap.end(), cand_less);
        }
    }
    void search(int n) {
        const KDTree::Node &nd = t.nodes[n];
        if (nd.left < 0) {
            for (int i = nd.lo; i < nd.hi; i++) offer(t.perm[i]);
            return;
        }
        double q = nd.dim == 0 ? qx : (nd.dim == 1 ? qy : qz);
        double diff = q - nd.split;
        int near = diff < 0 ? nd.left : nd.right;
        int far = diff < 0 ? nd.right : nd.left;
        search(near);
        // fl(diff*diff) is a lower bound of fl(d2) of every point beyond the plane
        // (rounding is monotone), so "<=" keeps tie candidates reachable.
        if ((int)heap.size() < K || diff * diff <= heap.front().d2) search(far);
    }
    void run(double x, double y, double z) {
        qx = x; qy = y; qz = z;
        heap.clear();
        if (t.N > 0) search(0);
        std::sort_heap(heap.begin(), heap.end(), cand_less);  // ascending
    }
};

// idx: M x K col-major 1-based; r: M x K col-major
int knn_all(int M, const double *q, int N, const double *pos, int K, int32_t *idx, double *r,
            int nthreads) {
    if (K > N) { g_err = "knn: K > N"; return -1; }
    KDTree tree;
    tree.init(N, pos);
#pragma omp parallel num_threads(nthreads)
    {
        KnnQuery Q(tree, K);
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < M; i++) {
            Q.run(q[i], q[i + (size_t)M], q[i + 2 * (size_t)M]);
            for (int j = 0; j < K; j++) {
                idx[i + (size_t)j * M] = Q.heap[j].id + 1;
                r[i + (size_t)j * M] = std::sqrt(Q.heap[j].d2);
            }
        }
    }
    return 0;
}

// -----------------------------------------------------------------------------
// Hydro (F/isothermal_hydroKDTree.jl, F/polytrope_hydroKDTree.jl)
// -----------------------------------------------------------------------------
struct HydroOut {
    // per particle
    double *ax, *ay, *az, *rho, *h;
    double *sum_vdw;  // sum_j (v_ij . gradW_ij), j=1..K   (F/isothermal_sim.jl:160)
    double *mumax;    // max_j mu_ij                       (F/isothermal_sim.jl:165)
    double *cs_i;     // poly only (may be null)
    double *dkdt;     // poly only: evolve_K! scatter sum (may be null)
};

// W  F/isothermal_hydroKDTree.jl:5-35 (poly: mask2 = !mask1, F/polytrope_hydroKDTree.jl:158)
inline double kernel_W(double h, double q, bool poly) {
    double ct = 1 / (PI * (h * h * h));
    if (q <= 1.0) return ct * ((1 - 3.0 / 2 * (q * q)) + 3.0 / 4 * (q * q * q));
    if (poly || q <= 2.0) {
        double t = 2 - q;
        return (ct * 1 / 4) * (t * t * t);
    }
    return 0.0;
}
// dWdr / r   F/isothermal_hydroKDTree.jl:38-73
inline double kernel_dWdr(double h, double q, double r, bool poly) {
    double h2 = h * h;
    double ct = 1 / (PI * (h2 * h2));
    if (q <= 1.0) return ct * (9.0 / 4 * r / h2 - 3 / h);
    if (poly || q <= 2.0) {
        double t = 2 - q;
        return ct * (-3.0 / 4 * (t * t)) / r;
    }
    return 0.0;
}

// One full hydrodynamics() call.  eos: 0 = isothermal, 1 = polytropic.
int hydro(int N, const double *pos, const double *vel, double m, int eos, double cs,
          const double *Kent, double gamma, double alpha, double beta, int K, int32_t *idx,
          double *r, HydroOut o, int nthreads) {
    const bool poly = eos == 1;
    const double *x = pos, *y = pos + N, *z = pos + 2 * (size_t)N;
    const double *vx = vel, *vy = vel + N, *vz = vel + 2 * (size_t)N;
    if (knn_all(N, pos, N, pos, K, idx, r, nthreads)) return -1;
    auto IDX = [&](int i, int j) { return idx[i + (size_t)j * N] - 1; };
    auto R = [&](int i, int j) { return r[i + (size_t)j * N]; };

    // h = r[:, end] ./ 2      (:151)
    for (int i = 0; i < N; i++) o.h[i] = R(i, K - 1) / 2;
    // rho = m * sum_j W        (:258-262), sequential over columns
    for (int i = 0; i < N; i++) {
        double s = 0.0;
        for (int j = 0; j < K; j++) {
            double q = R(i, j) / o.h[i];
            s += kernel_W(o.h[i], q, poly);
        }
        o.rho[i] = m * s;
    }
    // EOS
    std::vector<double> P(N), csv(N);
    for (int i = 0; i < N; i++) {
        if (!poly) {
            P[i] = cs * cs * o.rho[i];  // :190
            csv[i] = cs;
        } else {
            csv[i] = std::sqrt(gamma * Kent[i] * std::pow(o.rho[i], gamma - 1));  // poly :186
            P[i] = Kent[i] * std::pow(o.rho[i], gamma);                             // poly :216
            if (o.cs_i) o.cs_i[i] = csv[i];
        }
    }
    for (int i = 0; i < N; i++) {
        o.ax[i] = o.ay[i] = o.az[i] = 0.0;
        o.sum_vdw[i] = 0.0;
        o.mumax[i] = -std::numeric_limits<double>::infinity();
        if (o.dkdt) o.dkdt[i] = 0.0;
    }
    // pair terms; loop order "for j in 2:K, for i in 1:N" of hydroCalculation (:226-242)
    // is kept so the accumulation order into a[i] is the reference's.
    auto pair = [&](int i, int j, double &gx, double &gy, double &gz, double &Pi, double &mu,
                    double &vdw, int &nj) {
        nj = IDX(i, j);
        double dx = x[i] - x[nj], dy = y[i] - y[nj], dz = z[i] - z[nj];  // getTreeDiffs :93
        double rr = R(i, j), hi = o.h[i];
        double q = rr / hi;
        double dWdr = kernel_dWdr(hi, q, rr, poly);
        gx = dWdr * dx; gy = dWdr * dy; gz = dWdr * dz;
        double h_avg = (hi + o.h[nj]) / 2;              // getVectorTreeAvgs :111
        double rho_avg = (o.rho[i] + o.rho[nj]) / 2;
        double vijx = vx[i] - vx[nj], vijy = vy[i] - vy[nj], vijz = vz[i] - vz[nj];
        double v_dot_r = (vijx * dx + vijy * dy) + vijz * dz;                        // :210
        mu = std::min(h_avg * v_dot_r / (rr * rr + 0.01 * (h_avg * h_avg)), 0.0);    // :211
        Pi = ((-alpha) * csv[i] * mu + beta * (mu * mu)) / rho_avg;                   // :213
        vdw = (vijx * gx + vijy * gy) + vijz * gz;
    };
    // column 1 (self) contributes to the row reductions only
    for (int i = 0; i < N; i++) {
        double gx, gy, gz, Pi, mu, vdw; int nj;
        pair(i, 0, gx, gy, gz, Pi, mu, vdw, nj);
        o.sum_vdw[i] += vdw;
        o.mumax[i] = std::max(o.mumax[i], mu);
    }
    for (int j = 1; j < K; j++) {
        for (int i = 0; i < N; i++) {
            double gx, gy, gz, Pi, mu, vdw; int nj;
            pair(i, j, gx, gy, gz, Pi, mu, vdw, nj);
            double ct;
            if (!poly)
                ct = m * ((P[i] / (o.rho[i] * o.rho[i])) + Pi / 2);  // iso :232
            else
                ct = m * (((P[i] / (o.rho[i] * o.rho[i])) + (P[nj] / (o.rho[nj] * o.rho[nj]))) + Pi) / 2;  // poly :235
            o.ax[i] -= ct * gx; o.ay[i] -= ct * gy; o.az[i] -= ct * gz;
            o.ax[nj] += ct * gx; o.ay[nj] += ct * gy; o.az[nj] += ct * gz;
            o.sum_vdw[i] += vdw;
            o.mumax[i] = std::max(o.mumax[i], mu);
            if (o.dkdt) {  // evolve_K! poly :301-312
                double c2 = m * Pi * vdw / 2;
                o.dkdt[i] += c2;
                o.dkdt[nj] += c2;
            }
        }
    }
    return 0;
}

// -----------------------------------------------------------------------------
// Gravity (F/gravOctree_Single.jl)
// -----------------------------------------------------------------------------
struct CellNode {            // :33-64
    double Length;
    double Center[3];
    double lo[3], hi[3];     // axis_bounds
    double Mass = 0.0;
    double rCOM[3] = {0, 0, 0};
    int32_t parentID = -1;
    int32_t particle_count = 0;
    std::vector<int32_t> particle_list;
    std::vector<int32_t> child_nodes;
    bool is_leaf = false;
    int depth = 0;
};

struct Octree {
    double l, m, theta_sq;
    int N;
    const double *x, *y, *z, *h;
    std::vector<CellNode> nodes;
    std::vector<int32_t> leaf_list;
    int max_depth = 0;

    Octree(int n, double l_, double m_, const double *pos, double theta, const double *h_)
        : l(l_), m(m_), theta_sq(theta * theta), N(n), x(pos), y(pos + n), z(pos + 2 * (size_t)n), h(h_) {
        nodes.reserve((size_t)(1.6 * n) + 16);
        CellNode root;  // :94-104
        root.Length = l;
        root.Center[0] = root.Center[1] = root.Center[2] = 0.0;
        for (int d = 0; d < 3; d++) { root.lo[d] = -l; root.hi[d] = l; }
        root.particle_list.resize(n);
        std::iota(root.particle_list.begin(), root.particle_list.end(), 0);
        nodes.push_back(std::move(root));
    }

    // addNodes!  :107-181
    void addNodes(int pid) {
        double parent_l = nodes[pid].Length;
        double pc[3] = {nodes[pid].Center[0], nodes[pid].Center[1], nodes[pid].Center[2]};
        double child_l = parent_l / 2;
        double lc[3], rc[3], mn[3], ctr[3], mx[3];
        for (int d = 0; d < 3; d++) {
            lc[d] = pc[d] - child_l;  // left_x / down_y / outw_z
            rc[d] = pc[d] + child_l;  // right_x / up_y / inw_z
            mn[d] = lc[d] - child_l;  // left_minx
            ctr[d] = lc[d] + child_l; // centerx  (NOT recomputed from the parent centre)
            mx[d] = rc[d] + child_l;  // right_maxx
        }
        std::vector<int32_t> buckets[8];
        {
            const std::vector<int32_t> &pl = nodes[pid].particle_list;
            for (int32_t p : pl) {
                int ox = (x[p] - pc[0]) > 0, oy = (y[p] - pc[1]) > 0, oz = (z[p] - pc[2]) > 0;  // :143-148
                buckets[4 * oz + 2 * oy + ox].push_back(p);
            }
        }
        int pdepth = nodes[pid].depth;
        for (int c = 0; c < 8; c++) {
            if (buckets[c].empty()) continue;  // :161
            CellNode nn;
            nn.Length = child_l;
            int b[3] = {c & 1, (c >> 1) & 1, (c >> 2) & 1};
            for (int d = 0; d < 3; d++) {
                nn.Center[d] = b[d] ? rc[d] : lc[d];
                nn.lo[d] = b[d] ? ctr[d] : mn[d];
                nn.hi[d] = b[d] ? mx[d] : ctr[d];
            }
            // addParticles! :67-75
            nn.Mass += m * (double)buckets[c].size();
            nn.particle_count += (int32_t)buckets[c].size();
            nn.particle_list = std::move(buckets[c]);
            nn.parentID = pid;
            nn.depth = pdepth + 1;
            max_depth = std::max(max_depth, nn.depth);
            int32_t id = (int32_t)nodes.size();
            nodes.push_back(std::move(nn));
            nodes[pid].child_nodes.push_back(id);
        }
        std::vector<int32_t>().swap(nodes[pid].particle_list);  // :179
    }

    // build_octree!  :213-227
    int build(int depth_cap) {
        size_t i = 0;
        while (i < nodes.size()) {
            if (nodes[i].particle_count != 1) {
                if (nodes[i].depth >= depth_cap) {
                    g_err = "octree: depth cap reached (coincident particles?)";
                    return -1;
                }
                addNodes((int)i);
            }
            i++;
        }
        setCOMs();
        return 0;
    }

    // setCOMs!  :183-211
    void setCOMs() {
        leaf_list.clear();
        for (int i = (int)nodes.size() - 1; i >= 0; i--) {
            CellNode &nd = nodes[i];
            if (nd.particle_count == 1) {
                nd.is_leaf = true;
                leaf_list.push_back(i);
                int j = nd.particle_list[0];
                nd.rCOM[0] = x[j]; nd.rCOM[1] = y[j]; nd.rCOM[2] = z[j];
            } else {
                double total_mass = 0;
                double w[3] = {0, 0, 0};
                for (int32_t cid : nd.child_nodes) {
                    const CellNode &c = nodes[cid];
                    total_mass += c.Mass;
                    for (int d = 0; d < 3; d++) w[d] += c.Mass * c.rCOM[d];
                }
                nd.Mass = total_mass;
                for (int d = 0; d < 3; d++) nd.rCOM[d] = w[d] / total_mass;
            }
        }
    }
};

// Kernels :5-29
inline void grav_kernels(double r, double h, double &gPHI, double &PHI) {
    double q = r / h;
    double h2 = h * h, h3 = h * h * h, h4 = h2 * h2;
    double q2 = q * q, q3 = q * q * q, q4 = q2 * q2, q5 = q4 * q;
    double r2 = r * r, r3 = r * r * r;
    if (q <= 1) {
        gPHI = (1 / h2) * ((4.0 / 3 / h - 6.0 / 5 * (r2 / h3)) + 1.0 / 2 * (r3 / h4));
        PHI = (1 / h) * (((2.0 / 3 * q2 - 3.0 / 10 * q4) + 1.0 / 10 * q5) - 7.0 / 5);
    } else if (q <= 2) {
        gPHI = ((1 / h2) * ((((8.0 / 3 * q - 3 * q2) + 6.0 / 5 * q3) - 1.0 / 6 * q4) - 1.0 / 15 * (1 / q2))) / r;
        PHI = (1 / h) * (((((4.0 / 3 * q2 - q3) + 3.0 / 10 * q4) - 1.0 / 30 * q5) - 8.0 / 5) + 1.0 / 15 / q);
    } else {
        gPHI = 1 / r3;
        PHI = -1 / r;
    }
}

// min_distance2_point_to_cell :231-236
inline double mind2(const double p[3], const CellNode &n) {
    double s = 0;
    double d[3];
    for (int k = 0; k < 3; k++) d[k] = std::max(std::max(n.lo[k] - p[k], 0.0), p[k] - n.hi[k]);
    s = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    return s;
}

struct WalkStats {
    long long leaf = 0, mono = 0, opened = 0;
};

// compute_g :239-278.  skip_leaf >= 0: treat that leaf as absent (used by the
// threaded variant instead of the list surgery of gravity_acc).
inline void compute_g(const Octree &t, int i, int skip_leaf, double g[3], double &PHI, WalkStats *ws) {
    g[0] = g[1] = g[2] = 0.0;
    PHI = 0;
    double p[3] = {t.x[i], t.y[i], t.z[i]};
    double h_i = t.h[i];
    std::deque<int32_t> q;
    for (int32_t c : t.nodes[0].child_nodes) q.push_back(c);
    while (!q.empty()) {
        int32_t id = q.front();
        q.pop_front();
        if (id == skip_leaf) continue;
        const CellNode &n = t.nodes[id];
        double dx = p[0] - n.rCOM[0], dy = p[1] - n.rCOM[1], dz = p[2] - n.rCOM[2];
        double d_sq = (dx * dx + dy * dy) + dz * dz;
        double s = n.Length * 2;
        if (n.is_leaf) {
            int j = n.particle_list[0];
            double h_ij = (h_i + t.h[j]) / 2;
            double gP, pot;
            grav_kernels(std::sqrt(d_sq), h_ij, gP, pot);
            g[0] += n.Mass * (gP * dx); g[1] += n.Mass * (gP * dy); g[2] += n.Mass * (gP * dz);
            PHI += n.Mass * pot;
            if (ws) ws->leaf++;
        } else if ((s * s / d_sq < t.theta_sq) && (h_i * h_i / mind2(p, n) < 0.25)) {
            double d = std::sqrt(d_sq);
            double factor = n.Mass / (d * d * d);
            g[0] += factor * dx; g[1] += factor * dy; g[2] += factor * dz;
            PHI += -n.Mass / d;
            if (ws) ws->mono++;
        } else {
            for (int32_t c : n.child_nodes) q.push_back(c);
            if (ws) ws->opened++;
        }
    }
}

// gravity_acc :280-304 + gravity :307-319.  g: N x 3 col-major.
int gravity(int N, double l_domain, double m, const double *pos, double theta, const double *h,
            double *g, double *PHI, double *tree_stats, int nthreads) {
    Octree t(N, l_domain, m, pos, theta, h);
    if (t.build(200)) return -1;
    WalkStats total;
    int nleaf = (int)t.leaf_list.size();
    if (nthreads <= 1) {
        for (int k = 0; k < nleaf; k++) {
            int leafID = t.leaf_list[k];
            CellNode &leaf = t.nodes[leafID];
            int p_i = leaf.particle_list[0];
            // list surgery :293-300 (child order of the parent drifts; FP order only)
            std::vector<int32_t> &ch = t.nodes[leaf.parentID].child_nodes;
            ch.erase(std::find(ch.begin(), ch.end(), (int32_t)leafID));
            double gi[3], ph;
            compute_g(t, p_i, -1, gi, ph, &total);
            g[p_i] = gi[0]; g[p_i + (size_t)N] = gi[1]; g[p_i + 2 * (size_t)N] = gi[2];
            PHI[p_i] = ph;
            ch.push_back(leafID);
        }
    } else {
#pragma omp parallel num_threads(nthreads)
        {
            WalkStats loc;
#pragma omp for schedule(dynamic, 64)
            for (int k = 0; k < nleaf; k++) {
                int leafID = t.leaf_list[k];
                int p_i = t.nodes[leafID].particle_list[0];
                double gi[3], ph;
                compute_g(t, p_i, leafID, gi, ph, &loc);
                g[p_i] = gi[0]; g[p_i + (size_t)N] = gi[1]; g[p_i + 2 * (size_t)N] = gi[2];
                PHI[p_i] = ph;
            }
#pragma omp critical
            { total.leaf += loc.leaf; total.mono += loc.mono; total.opened += loc.opened; }
        }
    }
    for (int i = 0; i < N; i++) PHI[i] = PHI[i] - (m * (7.0 / 5) / h[i]);  // :303
    if (tree_stats) {
        tree_stats[0] = (double)t.nodes.size();
        tree_stats[1] = (double)t.max_depth;
        tree_stats[2] = (double)total.leaf;
        tree_stats[3] = (double)total.mono;
        tree_stats[4] = (double)total.opened;
    }
    return 0;
}

// -----------------------------------------------------------------------------
// getAcc (F/isothermal_sim.jl:16-49, F/polytrope_sim.jl:17-51)
// -----------------------------------------------------------------------------
struct AccWork {
    int N, K;
    std::vector<int32_t> idx;
    std::vector<double> r, ax, ay, az, rho, h, sum_vdw, mumax, cs_i, dkdt, g, phi;
    AccWork(int n, int k) : N(n), K(k), idx((size_t)n * k), r((size_t)n * k), ax(n), ay(n), az(n), rho(n),
                            h(n), sum_vdw(n), mumax(n), cs_i(n), dkdt(n), g(3 * (size_t)n), phi(n) {}
};

int get_acc(AccWork &w, const double *pos, const double *vel, double m, int eos, double cs,
            const double *Kent, double gamma, double G, double theta, double alpha, double beta,
            double *acc, double *tree_stats, int nthreads) {
    int N = w.N;
    double l_domain = 0.0;  // maximum(abs.(pos)) :33
    for (size_t i = 0; i < 3 * (size_t)N; i++) l_domain = std::max(l_domain, std::fabs(pos[i]));
    HydroOut o{w.ax.data(), w.ay.data(), w.az.data(), w.rho.data(), w.h.data(), w.sum_vdw.data(),
               w.mumax.data(), eos == 1 ? w.cs_i.data() : nullptr, eos == 1 ? w.dkdt.data() : nullptr};
    if (hydro(N, pos, vel, m, eos, cs, Kent, gamma, alpha, beta, w.K, w.idx.data(), w.r.data(), o, nthreads))
        return -1;
    if (gravity(N, l_domain, m, pos, theta, w.h.data(), w.g.data(), w.phi.data(), tree_stats, nthreads))
        return -1;
    for (int i = 0; i < N; i++) {  // :41-46
        acc[i] = w.ax[i] - G * w.g[i];
        acc[i + (size_t)N] = w.ay[i] - G * w.g[i + (size_t)N];
        acc[i + 2 * (size_t)N] = w.az[i] - G * w.g[i + 2 * (size_t)N];
    }
    return 0;
}

// adaptive dt  F/isothermal_sim.jl:158-166 / F/polytrope_sim.jl:165-174
double adaptive_dt(const AccWork &w, const double *vel, const double *acc, double m, int eos, double cs,
                   double alpha, double beta) {
    int N = w.N;
    double inf = std::numeric_limits<double>::infinity();
    double c1 = inf, c2 = inf, c3 = inf, c4 = inf;
    for (int i = 0; i < N; i++) {
        double vx = vel[i], vy = vel[i + (size_t)N], vz = vel[i + 2 * (size_t)N];
        double axx = acc[i], ayy = acc[i + (size_t)N], azz = acc[i + 2 * (size_t)N];
        double vel_r = std::sqrt((vx * vx + vy * vy) + vz * vz);
        double a_r = std::sqrt((axx * axx + ayy * ayy) + azz * azz);
        double abs_div_v = std::fabs(-(m * w.sum_vdw[i]) / w.rho[i]);
        double c = eos == 1 ? w.cs_i[i] : cs;
        c1 = std::min(c1, 1 / abs_div_v);
        c2 = std::min(c2, w.h[i] / vel_r);
        c3 = std::min(c3, std::sqrt(w.h[i] / a_r));
        c4 = std::min(c4, w.h[i] / (c + 1.2 * (alpha * c + beta * w.mumax[i])));
    }
    return 0.3 * std::min(std::min(c1, c2), std::min(c3, c4));
}

}  // namespace

// =============================================================================
// C ABI (ctypes from tests / bench cpu_baseline only)
// =============================================================================
extern "C" {

const char *oracle_last_error() { return g_err.c_str(); }

int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// getNeighbors core: queries q (M x 3) against pos (N x 3)
int oracle_knn(int M, const double *q, int N, const double *pos, int K, int32_t *idx, double *r, int nthreads) {
    return knn_all(M, q, N, pos, K, idx, r, std::max(1, nthreads));
}

// HJL.hydrodynamics.  Outputs (any may NOT be null except cs_i/dkdt for eos=0):
//   idx N x K int32 (1-based, col-major), r N x K, ahyd N x 3, rho, h, sum_vdw, mumax, cs_i, dkdt
int oracle_hydro(int N, const double *pos, const double *vel, double m, int eos, double cs, const double *Kent,
                 double gamma, double alpha, double beta, int K, int32_t *idx, double *r, double *ahyd,
                 double *rho, double *h, double *sum_vdw, double *mumax, double *cs_i, double *dkdt,
                 int nthreads) {
    HydroOut o{ahyd, ahyd + N, ahyd + 2 * (size_t)N, rho, h, sum_vdw, mumax, eos == 1 ? cs_i : nullptr,
               eos == 1 ? dkdt : nullptr};
    return hydro(N, pos, vel, m, eos, cs, Kent, gamma, alpha, beta, K, idx, r, o, std::max(1, nthreads));
}

// GJL.gravity.  tree_stats (5 doubles, may be null): nodes, max depth, leaf / monopole / opened visits
int oracle_gravity(int N, double l_domain, double m, const double *pos, double theta, const double *h,
                   double *g, double *phi, double *tree_stats, int nthreads) {
    return gravity(N, l_domain, m, pos, theta, h, g, phi, tree_stats, std::max(1, nthreads));
}

// getAcc.  Outputs: acc N x 3, rho, h, phi, sum_vdw, mumax, (cs_i, dkdt: poly)
int oracle_getacc(int N, const double *pos, const double *vel, double m, int eos, double cs, const double *Kent,
                  double gamma, double G, double theta, double alpha, double beta, int K, double *acc,
                  double *rho, double *h, double *phi, double *sum_vdw, double *mumax, double *cs_i,
                  double *dkdt, int nthreads) {
    AccWork w(N, K);
    if (get_acc(w, pos, vel, m, eos, cs, Kent, gamma, G, theta, alpha, beta, acc, nullptr, std::max(1, nthreads)))
        return -1;
    std::copy(w.rho.begin(), w.rho.end(), rho);
    std::copy(w.h.begin(), w.h.end(), h);
    std::copy(w.phi.begin(), w.phi.end(), phi);
    if (sum_vdw) std::copy(w.sum_vdw.begin(), w.sum_vdw.end(), sum_vdw);
    if (mumax) std::copy(w.mumax.begin(), w.mumax.end(), mumax);
    if (eos == 1 && cs_i) std::copy(w.cs_i.begin(), w.cs_i.end(), cs_i);
    if (eos == 1 && dkdt) std::copy(w.dkdt.begin(), w.dkdt.end(), dkdt);
    return 0;
}

// nsteps iterations of the `while t < tEnd` body (F/isothermal_sim.jl:152-213,
// F/polytrope_sim.jl:158-232) without I/O.  pos, vel, Kent updated in place.
// stats: nsteps x 10 row-major  [t, T, V, U, Etot, rcx, rcy, rcz, |p|, |L|]  (t = time at step start)
// dts: nsteps.  U_iso: the constant "U" of the isothermal snapshot.
int oracle_step(int N, double *pos, double *vel, double *Kent, double m, int eos, double cs, double gamma,
                double G, double theta, double alpha, double beta, int K, double U_iso, double *t_inout,
                int nsteps, double *dts, double *stats, int nthreads) {
    nthreads = std::max(1, nthreads);
    AccWork w(N, K);
    std::vector<double> acc(3 * (size_t)N), pos_half(3 * (size_t)N), vel_half(3 * (size_t)N), vdw0(N), rho0(N);
    double t = *t_inout;
    size_t n3 = 3 * (size_t)N;
    for (int s = 0; s < nsteps; s++) {
        if (get_acc(w, pos, vel, m, eos, cs, Kent, gamma, G, theta, alpha, beta, acc.data(), nullptr, nthreads))
            return -1;
        double dt = adaptive_dt(w, vel, acc.data(), m, eos, cs, alpha, beta);
        // ---- statistics (:168-192 / poly :177-205)
        double T = 0, sumPhi = 0, sx = 0, sy = 0, sz = 0, px = 0, py = 0, pz = 0;
        for (int i = 0; i < N; i++) {
            double vx = vel[i], vy = vel[i + (size_t)N], vz = vel[i + 2 * (size_t)N];
            double vr = std::sqrt((vx * vx + vy * vy) + vz * vz);
            T += vr * vr;
            sumPhi += w.phi[i];
            sx += pos[i]; sy += pos[i + (size_t)N]; sz += pos[i + 2 * (size_t)N];
            px += vx; py += vy; pz += vz;
        }
        T = 0.5 * m * T;
        double V = G / 2 * m * sumPhi;
        double U, Etot;
        if (eos == 1) {
            double su = 0;
            for (int i = 0; i < N; i++) su += Kent[i] / (gamma - 1) * std::pow(w.rho[i], gamma - 1);
            U = m * su;
            Etot = T + V + U;
        } else {
            U = U_iso;
            Etot = T + V + 2 * U;
        }
        double rcx = sx / N, rcy = sy / N, rcz = sz / N;
        px *= m; py *= m; pz *= m;
        double lx = 0, ly = 0, lz = 0;
        for (int i = 0; i < N; i++) {
            double ax_ = pos[i] - rcx, ay_ = pos[i + (size_t)N] - rcy, az_ = pos[i + 2 * (size_t)N] - rcz;
            double bx = vel[i], by = vel[i + (size_t)N], bz = vel[i + 2 * (size_t)N];
            lx += ay_ * bz - az_ * by;
            ly += az_ * bx - ax_ * bz;
            lz += ax_ * by - ay_ * bx;
        }
        lx *= m; ly *= m; lz *= m;
        if (stats) {
            double *row = stats + 10 * (size_t)s;
            row[0] = t; row[1] = T; row[2] = V; row[3] = U; row[4] = Etot;
            row[5] = rcx; row[6] = rcy; row[7] = rcz;
            row[8] = std::sqrt((px * px + py * py) + pz * pz);
            row[9] = std::sqrt((lx * lx + ly * ly) + lz * lz);
        }
        // ---- predictor (:197-200)
        for (size_t i = 0; i < n3; i++) {
            pos_half[i] = pos[i] + vel[i] * dt / 2;
            vel_half[i] = vel[i] + acc[i] * dt / 2;
        }
        if (eos == 1) {  // evolve_K! with full-step data (poly :217)
            for (int i = 0; i < N; i++)
                Kent[i] = Kent[i] + (1.0 / 2 * (gamma - 1) / std::pow(w.rho[i], gamma - 1) * w.dkdt[i]) * (dt / 2);
        }
        if (get_acc(w, pos_half.data(), vel_half.data(), m, eos, cs, Kent, gamma, G, theta, alpha, beta,
                    acc.data(), nullptr, nthreads))
            return -1;
        if (eos == 1) {  // evolve_K! with half-step data (poly :221)
            for (int i = 0; i < N; i++)
                Kent[i] = Kent[i] + (1.0 / 2 * (gamma - 1) / std::pow(w.rho[i], gamma - 1) * w.dkdt[i]) * (dt / 2);
        }
        // ---- corrector (:206-209)
        for (size_t i = 0; i < n3; i++) {
            vel[i] += acc[i] * dt;
            pos[i] += vel[i] * dt - (1.0 / 2) * acc[i] * (dt * dt);
        }
        t += dt;
        if (dts) dts[s] = dt;
    }
    *t_inout = t;
    return 0;
}

// adaptive dt alone, from getAcc outputs (for unit tests)
double oracle_dt(int N, const double *vel, const double *acc, const double *rho, const double *h,
                 const double *sum_vdw, const double *mumax, const double *cs_i, double m, int eos, double cs,
                 double alpha, double beta) {
    AccWork w(N, 1);
    std::copy(rho, rho + N, w.rho.begin());
    std::copy(h, h + N, w.h.begin());
    std::copy(sum_vdw, sum_vdw + N, w.sum_vdw.begin());
    std::copy(mumax, mumax + N, w.mumax.begin());
    if (eos == 1) std::copy(cs_i, cs_i + N, w.cs_i.begin());
    return adaptive_dt(w, vel, acc, m, eos, cs, alpha, beta);
}

// HJL.density_plot (F/isothermal_hydroKDTree.jl:291-297): density at M sample points
int oracle_density_at(int M, const double *pts, int N, const double *pos, double m, int K, int eos,
                      double *rho_out, int nthreads) {
    std::vector<int32_t> idx((size_t)M * K);
    std::vector<double> r((size_t)M * K);
    if (knn_all(M, pts, N, pos, K, idx.data(), r.data(), std::max(1, nthreads))) return -1;
    for (int i = 0; i < M; i++) {
        double h = r[i + (size_t)(K - 1) * M] / 2;
        double s = 0;
        for (int j = 0; j < K; j++) s += kernel_W(h, r[i + (size_t)j * M] / h, eos == 1);
        rho_out[i] = m * s;
    }
    return 0;
}

// Octree structure dump for tests: per node [Length, cx,cy,cz, lox,loy,loz, hix,hiy,hiz, Mass, comx,comy,comz, count, depth]
// Returns number of nodes (or -1); if out==null only counts.
long long oracle_octree(int N, double l_domain, double m, const double *pos, double *out, long long cap) {
    std::vector<double> h(N, 1.0);
    Octree t(N, l_domain, m, pos, 0.5, h.data());
    if (t.build(200)) return -1;
    long long n = (long long)t.nodes.size();
    if (out) {
        for (long long i = 0; i < std::min(n, cap); i++) {
            const CellNode &c = t.nodes[i];
            double *o = out + 16 * i;
            o[0] = c.Length;
            for (int d = 0; d < 3; d++) { o[1 + d] = c.Center[d]; o[4 + d] = c.lo[d]; o[7 + d] = c.hi[d]; o[11 + d] = c.rCOM[d]; }
            o[10] = c.Mass;
            o[14] = c.particle_count;
            o[15] = c.depth;
        }
    }
    return n;
}

}  // extern "C"
