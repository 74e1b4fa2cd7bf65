/*
 * sph_b200.h -- C ABI of libsph_b200.so: the per-step SPH core of the reference engine
 * julia_version/fastv1_kd&single_oc (below "F/") on one NVIDIA B200 (sm_100a) per handle.
 *
 * The reference has no FFI; its seam is the Julia call boundary
 *     getAcc(pos, vel, m, cs | (K, gamma), G, theta, alpha, beta, Kh)   F/isothermal_sim.jl:16-49, F/polytrope_sim.jl:17-51
 * and the body of `while t < tEnd`                                     F/isothermal_sim.jl:152-213, F/polytrope_sim.jl:158-232
 * Each entry point below names the reference code it replaces.  INTEGRATION.md shows the Julia `ccall` side.
 *
 * Conventions (identical to the Julia side, so Julia arrays are passed as Ptr{Float64} without copies):
 *   - every N x 3 matrix is COLUMN-MAJOR: x = p, y = p + N, z = p + 2N      (Matrix{Float64}, F/isothermal_sim.jl:16)
 *   - neighbour indices are N x Kh Int32, column-major, 1-based, column 1 = the particle itself,
 *     ascending distance                                                    (F/isothermal_hydroKDTree.jl:128-142)
 *   - particle order at the boundary is always the caller's (snapshot) order.
 *   - all pointers are HOST pointers unless the name says `_dev`; the caller owns them; the library owns
 *     all device memory.  Every call returns 0 on success or a negative SPH_ERR_* code; the message is
 *     available from sph_last_error().  No exceptions, no callbacks.  A handle is not re-entrant.
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_B200_ABI_VERSION 2

enum sph_eos {
    SPH_EOS_ISOTHERMAL = 0, /* P = cs^2 rho            F/isothermal_hydroKDTree.jl:181-193 */
    SPH_EOS_POLYTROPIC = 1  /* P = K_i rho^gamma       F/polytrope_hydroKDTree.jl:207-219  */
};

enum sph_status {
    SPH_OK = 0,
    SPH_ERR_INVALID = -1,     /* bad argument (null pointer, K > N, N < 64, ...)                              */
    SPH_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed; message holds cudaGetErrorString        */
    SPH_ERR_NO_DEVICE = -3,   /* no sm_100 device visible: the library never falls back to the CPU             */
    SPH_ERR_TREE_DEPTH = -4,  /* two particles share all 42 octant levels (coincident, or closer than l/2^42 per
                                 axis); the reference loops forever on coincident particles
                                 (F/gravOctree_Single.jl:217-223)                                               */
    SPH_ERR_TREE_NODES = -5,  /* node pool (3 N + 1024 nodes; SPH_B200_NODE_FACTOR enlarges it) exhausted         */
    SPH_ERR_NCCL = -6,        /* NCCL call failed                                                              */
    SPH_ERR_STATE = -7,       /* call sequence error (e.g. sph_step before sph_upload)                         */
    SPH_ERR_NAN = -8          /* the adaptive time step came out NaN (non-finite state).  The reference's
                                 minimum() propagates NaN, t becomes NaN and `while t < tEnd` ends
                                 (F/isothermal_sim.jl:152,158-166); here the step is refused instead            */
};

/* Run-time physics parameters = the `constants` row of a snapshot (F/SnapshotRW.jl:86-97) that
 * run_simulation unpacks at F/isothermal_sim.jl:87-105 / F/polytrope_sim.jl:92-112. */
typedef struct sph_params {
    int64_t N;      /* number of particles                                                   */
    int32_t Kh;     /* neighbours per particle incl. self (default 50, F/iniconds.jl:539)    */
    int32_t eos;    /* enum sph_eos                                                          */
    double m;       /* particle mass                                                         */
    double cs;      /* isothermal sound speed (ignored for polytropic)                       */
    double gamma;   /* adiabatic index (ignored for isothermal)                              */
    double G;       /* gravitational constant                                                */
    double theta;   /* Barnes-Hut opening angle                                              */
    double alpha;   /* artificial viscosity alpha                                            */
    double beta;    /* artificial viscosity beta                                             */
    double U_iso;   /* constant thermal energy "U" of an isothermal snapshot (stats only)    */
    int32_t device; /* CUDA device ordinal of this handle (one process per GPU)              */
    int32_t flags;  /* SPH_FLAG_* bits, 0 by default                                         */
} sph_params;

/* sph_params.flags */
#define SPH_FLAG_COUNT_VISITS 1 /* count the node visits of the tree walk (sph_timings.walk_visits); slows the walk */
#define SPH_FLAG_SERIAL_PHASES 2 /* run density / force BEFORE the walk on one stream instead of beside it (same results):
                                    sph_timings then holds the time of every phase alone (measurement; slower overall) */

/* One row of the reference's stats matrix (F/isothermal_sim.jl:189-192) plus the step's dt. */
typedef struct sph_step_info {
    double dt;        /* adaptive time step used (F/isothermal_sim.jl:158-166)                          */
    double stats[10]; /* [t, T, V, U, Etot, rcom_x, rcom_y, rcom_z, |p|, |L|], t = time at step start   */
} sph_step_info;

/* Device time of the phases of the LAST force evaluation, milliseconds (CUDA events on the handle's
 * stream).  Replaces the @debug phase timers of F/isothermal_hydroKDTree.jl:250-285 and
 * F/gravOctree_Single.jl:308-316. */
typedef struct sph_timings {
    double sort_ms;      /* domain max, keys, radix sort, permute          */
    double tree_ms;      /* linear octree build + centre-of-mass sweep     */
    double knn_ms;       /* exact K-nearest search                         */
    double density_ms;   /* density + EOS                                  */
    double force_ms;     /* pressure + artificial-viscosity force          */
    double gravity_ms;   /* tree walk                                      */
    double finish_ms;    /* assemble acc, un-permute, collectives          */
    double total_ms;
    double walk_visits;  /* sum over targets of node visits in the last walk (0 unless SPH_B200_COUNT_VISITS) */
    double knn_retries;  /* targets whose hinted search radius held < Kh particles and was repeated           */
    double comm_ms;      /* of the above: time inside the NCCL collectives (0 on a single GPU)               */
    double walk_kernel_ms; /* the tree-walk kernel alone (gravity_ms also holds the record packing, the partial-sum
                              reduction and, on several GPUs, the result all-gather)                            */
} sph_timings;

typedef struct sph_handle sph_handle;

/* ---- lifecycle ---------------------------------------------------------------------------------- */
/* Creates the device context for N particles.  Replaces the unpacking at F/isothermal_sim.jl:87-105. */
int sph_create(const sph_params *params, sph_handle **out);
int sph_destroy(sph_handle *h);
/* Message of the last failing call on this handle (h may be NULL: last error of sph_create). */
const char *sph_last_error(const sph_handle *h);
int sph_abi_version(void);
/* Cumulative number of CUDA kernels this library has launched in the calling process (diagnostics; bench.py
 * reports the per-step difference as gpu_launches). */
int64_t sph_launch_count(void);
/* Number of CUDA devices visible (0 = none; never an error).  */
int sph_device_count(void);
/* Measured FP64 FMA throughput of `device` in TFLOP/s (own DFMA microbenchmark, a few milliseconds): the denominator of
 * the FP64 roofline bench.py reports for the tree walk and the SPH sums (not a reference interface). */
int sph_measure_fp64_peak(int device, double *tflops);
/* Run the handle's work on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL = own stream). */
int sph_set_stream(sph_handle *h, void *cuda_stream);
int sph_synchronize(sph_handle *h);

/* ---- state --------------------------------------------------------------------------------------- */
/* Host -> device: pos, vel (N x 3 column-major), K (N, polytropic only, else NULL), time t.
 * Replaces read_snapshot's hand-over at F/isothermal_sim.jl:78-85 / F/polytrope_sim.jl:116-117.
 * On a handle that joined a communicator this is a COLLECTIVE call: every rank passes the same arrays, moves only its
 * 1/nranks slice of them over PCIe and the ranks exchange the slices over NVLink. */
int sph_upload(sph_handle *h, const double *pos, const double *vel, const double *K_or_null, double t);
/* Device -> host in the caller's particle order (any pointer may be NULL). */
int sph_download(sph_handle *h, double *pos, double *vel, double *K_or_null, double *t);

/* ---- hot path ------------------------------------------------------------------------------------ */
/* One getAcc (F/isothermal_sim.jl:16-49 / F/polytrope_sim.jl:17-51) on caller-supplied pos/vel(/K):
 * acc N x 3, rho, hsml, phi (any output may be NULL).  Does not touch the uploaded state. */
int sph_eval_acc(sph_handle *h, const double *pos, const double *vel, const double *K_or_null,
                 double *acc, double *rho, double *hsml, double *phi);
/* Same, on the uploaded state (no host<->device traffic); results stay on the device and are read
 * with the sph_get_* calls. */
int sph_eval_state(sph_handle *h);
/* nsteps iterations of the loop body F/isothermal_sim.jl:155-212 / F/polytrope_sim.jl:162-231 on the
 * uploaded state: getAcc, adaptive dt, statistics, predictor, getAcc, corrector (and evolve_K! twice
 * for polytropic, F/polytrope_hydroKDTree.jl:296-341).  info: nsteps entries or NULL. */
int sph_step(sph_handle *h, int nsteps, sph_step_info *info);

/* ---- inspection of the last force evaluation (parity tests) --------------------------------------- */
/* HJL.getNeighbors output (F/isothermal_hydroKDTree.jl:118-163): idx N x Kh (1-based, column-major),
 * r N x Kh distances (either may be NULL). */
int sph_get_neighbors(sph_handle *h, int32_t *idx, double *r);
/* HJL.hydrodynamics outputs reduced to what its caller consumes (F/isothermal_hydroKDTree.jl:287,
 * F/isothermal_sim.jl:160,165): ahyd N x 3, rho, hsml, sum_vdw = sum_j v_ij.gradW_ij, mumax = max_j mu_ij,
 * cs_i and dkdt (polytropic: F/polytrope_hydroKDTree.jl:186 and :301-312).  Any may be NULL. */
int sph_get_hydro(sph_handle *h, double *ahyd, double *rho, double *hsml, double *sum_vdw, double *mumax,
                  double *cs_i, double *dkdt);
/* GJL.gravity outputs (F/gravOctree_Single.jl:307-319): g N x 3 (NOT multiplied by G), PHI. */
int sph_get_grav(sph_handle *h, double *g, double *phi);
/* Total acceleration of the last evaluation, N x 3. */
int sph_get_acc(sph_handle *h, double *acc);
/* Octree of the last evaluation in the reference's BFS node order (F/gravOctree_Single.jl:213-227).
 * nodes: n_nodes x 16 row-major [Length, cx,cy,cz, lox,loy,loz, hix,hiy,hiz, Mass, comx,comy,comz,
 * particle_count, depth]; pass NULL to query only the count. */
int sph_get_octree(sph_handle *h, double *nodes, int64_t cap, int64_t *n_nodes);
int sph_get_timings(sph_handle *h, sph_timings *out);
/* Adaptive dt of the last evaluation against the uploaded velocities (F/isothermal_sim.jl:158-166). */
int sph_get_dt(sph_handle *h, double *dt);

/* ---- snapshot helper ------------------------------------------------------------------------------ */
/* HJL.density_plot (F/isothermal_hydroKDTree.jl:291-297): SPH density at M arbitrary points (M x 3
 * column-major) from the uploaded positions. */
int sph_density_at(sph_handle *h, const double *pts, int64_t M, double *rho_out);

/* ---- multi-GPU (one process per GPU; the reference is single-process) ---------------- */
/* 128-byte NCCL unique id, created on rank 0 and distributed by the host (torch.distributed / MPI). */
int sph_comm_unique_id(void *id128);
/* Join a communicator of nranks (<= 16) handles.  Afterwards every rank holds the full state; the targets of
 * search / density / force / walk are split by Morton-key range and the results exchanged with NCCL all-gathers
 * (nothing is reduced across ranks: the force is evaluated in gather form).  sph_upload, sph_eval_*, sph_step and the
 * sph_get_* calls are then collective: every rank must make the same calls in the same order. */
int sph_comm_init(sph_handle *h, int nranks, int rank, const void *id128);

#ifdef __cplusplus
}
#endif
#endif /* SPH_B200_H */
