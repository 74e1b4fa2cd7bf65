// radix_sort.cu -- CUB-free stable LSD radix sort of (u64 key, i32 value) pairs and a device-wide
// exclusive scan, both hand-written for sm_100a.
//
// Replaces the implicit grouping of particles by octant in addNodes! (F/gravOctree_Single.jl:151-156;
// the reference re-buckets the particle list of every node it subdivides) and the KD-tree build of
// NearestNeighbors.jl (F/isothermal_hydroKDTree.jl:128): one sort by 63-bit octant path key orders the
// particles so that EVERY octree cell of every level is a contiguous range.
//
// One-sweep passes (8 bits per pass, tile = 256 threads x 8 keys): one histogram kernel for all passes, then
// ONE kernel per pass that ranks its tile with __match_any_sync + per-warp digit counters, learns the digits of the
// preceding tiles by decoupled look-back and scatters.  HBM traffic per pass: 1 read + 1 write of 12 B per element.
#include <cstdlib>

#include "sph_internal.cuh"

namespace {

constexpr int RS_BITS = 8;
constexpr int RS_BINS = 1 << RS_BITS;
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;

constexpr int SC_THREADS = 512;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ int64_t eff_n(int64_t n, const unsigned long long *n_dev) {
    return n_dev ? min((int64_t)*n_dev, n) : n;
}

// ---------------------------------------------------------------- one-sweep passes
// One kernel per pass instead of five: the global digit counts of ALL passes come from one read of the keys
// (os_hist_kernel), and inside a pass every tile obtains the number of equal digits in the tiles before it by a
// decoupled look-back over per-tile status words {flag:2 | count:30} (tile ids are handed out by an atomic counter,
// so a tile only ever waits for tiles whose blocks are already running).
constexpr int OS_ITEMS = 8;
constexpr int OS_TILE = RS_THREADS * OS_ITEMS;
#ifndef OS_WIN
#define OS_WIN 8                 // predecessors read per look-back round trip (build-time experiment: tools/build_variants.sh)
#endif
constexpr unsigned OS_AGG = 1u << 30, OS_INC = 2u << 30, OS_MASK = (1u << 30) - 1u;

__global__ void __launch_bounds__(RS_THREADS) os_hist_kernel(const uint64_t *__restrict__ keys, int64_t n_cap,
                                                              const unsigned long long *n_dev, int begin_bit, int npass,
                                                              unsigned *__restrict__ ghist /* [npass][256] */) {
    __shared__ unsigned hist[8][RS_BINS];
    const int64_t n = eff_n(n_cap, n_dev);
    for (int i = threadIdx.x; i < 8 * RS_BINS; i += RS_THREADS) (&hist[0][0])[i] = 0;
    __syncthreads();
    for (int64_t e = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; e < n; e += (int64_t)gridDim.x * RS_THREADS) {
        const uint64_t k = keys[e];
        for (int p = 0; p < npass; ++p) atomicAdd(&hist[p][(unsigned)((k >> (begin_bit + p * RS_BITS)) & (RS_BINS - 1))], 1u);
    }
    __syncthreads();
    for (int p = 0; p < npass; ++p) {
        const unsigned c = hist[p][threadIdx.x];
        if (c) atomicAdd(&ghist[p * RS_BINS + threadIdx.x], c);
    }
}

// exclusive prefix over the 256 digits of every pass, in place
__global__ void __launch_bounds__(RS_THREADS) os_prefix_kernel(unsigned *__restrict__ ghist, int npass) {
    __shared__ unsigned sm[RS_BINS];
    for (int p = 0; p < npass; ++p) {
        const unsigned v = ghist[p * RS_BINS + threadIdx.x];
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < RS_BINS; o <<= 1) {
            const unsigned t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0u;
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        ghist[p * RS_BINS + threadIdx.x] = sm[threadIdx.x] - v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS) os_pass_kernel(const uint64_t *__restrict__ keys_in,
                                                              const int *__restrict__ vals_in,
                                                              uint64_t *__restrict__ keys_out, int *__restrict__ vals_out,
                                                              int64_t n_cap, const unsigned long long *n_dev, int shift,
                                                              const unsigned *__restrict__ gpref /* [256] */,
                                                              volatile unsigned *status /* [ntiles][256] */,
                                                              unsigned *tile_counter) {
    __shared__ int cnt[RS_WARPS][RS_BINS];
    __shared__ int s_tile;
    const int64_t n = eff_n(n_cap, n_dev);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(tile_counter, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int tile = s_tile;
    const int64_t wbase = (int64_t)tile * OS_TILE + (int64_t)warp * (32 * OS_ITEMS);
    uint64_t key[OS_ITEMS];
    int rank[OS_ITEMS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < OS_ITEMS; ++it) {
        const int64_t e = wbase + it * 32 + lane;
        const bool ok = e < n;
        key[it] = ok ? keys_in[e] : 0ull;
        const unsigned dig = ok ? (unsigned)((key[it] >> shift) & (RS_BINS - 1)) : (0x10000u + lane);
        const unsigned grp = __match_any_sync(0xffffffffu, dig);
        const int leader = __ffs(grp) - 1;
        int pre = 0;
        if (ok && lane == leader) {
            pre = cnt[warp][dig];
            cnt[warp][dig] = pre + __popc(grp);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[it] = pre + __popc(grp & lt);
        __syncwarp();
    }
    __syncthreads();
    {   // thread d: tile total of digit d, look-back for the digits of the preceding tiles, per-warp start offsets
        const int d = threadIdx.x;
        unsigned total = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) total += (unsigned)cnt[w][d];
        unsigned excl = 0;
        if (tile == 0) {
            status[d] = OS_INC | total;
        } else {
            status[(size_t)tile * RS_BINS + d] = OS_AGG | total;
            // windows of OS_WIN predecessors: the loads of a window are independent, so one L2 round trip covers them
            // all (most predecessors only hold their local count yet, so a late tile walks back a long way)
            bool done = false;
            for (int t = tile - 1; t >= 0 && !done; t -= OS_WIN) {
                unsigned v[OS_WIN];
#pragma unroll
                for (int w = 0; w < OS_WIN; ++w) v[w] = t - w >= 0 ? status[(size_t)(t - w) * RS_BINS + d] : (2u << 30);
#pragma unroll
                for (int w = 0; w < OS_WIN; ++w) {
                    if (done) break;
                    while ((v[w] >> 30) == 0u) { __nanosleep(20); v[w] = status[(size_t)(t - w) * RS_BINS + d]; }
                    excl += v[w] & OS_MASK;
                    done = (v[w] >> 30) == 2u;
                }
            }
            status[(size_t)tile * RS_BINS + d] = OS_INC | (excl + total);
        }
        int run = (int)(gpref[d] + excl);
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const int c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < OS_ITEMS; ++it) {
        const int64_t e = wbase + it * 32 + lane;
        if (e < n) {
            const unsigned dig = (unsigned)((key[it] >> shift) & (RS_BINS - 1));
            const int dst = cnt[warp][dig] + rank[it];
            keys_out[dst] = key[it];
            vals_out[dst] = vals_in[e];
        }
    }
}

// ---------------------------------------------------------------- scan (three phases)
__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *smem /* >= 32 ints */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? smem[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;  // exclusive offset of each warp
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    const int res = smem[warp] + inc - v;
    if (total) *total = smem[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SC_THREADS) scan_tile_sums(const int *__restrict__ in, int64_t n,
                                                              int *__restrict__ tile_sum) {
    __shared__ int sm[40];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    int tot;
    block_exclusive_scan(s, &tot, sm);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

// single block: exclusive scan of the tile sums in place; also writes the grand total to tile_sum[ntiles]
__global__ void __launch_bounds__(1024) scan_of_sums(int *tile_sum, int ntiles) {
    __shared__ int sm[40];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < ntiles; b += 1024) {
        const int i = b + threadIdx.x;
        const int v = i < ntiles ? tile_sum[i] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, &tot, sm);
        if (i < ntiles) tile_sum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sum[ntiles] = carry;
}

__global__ void __launch_bounds__(SC_THREADS) scan_apply(const int *__restrict__ in, int *__restrict__ out, int64_t n,
                                                          const int *__restrict__ tile_sum, int ntiles) {
    __shared__ int sm[40];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    int v[SC_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int run = block_exclusive_scan(s, nullptr, sm) + tile_sum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tile_sum[ntiles];
}

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

static size_t onesweep_temp_bytes(int64_t n) {   // [8][256] counts, [8] tile counters, [8][ntiles][256] status words
    return align256((size_t)(8 * RS_BINS + 8) * 4) + (size_t)8 * cdiv(n, OS_TILE) * RS_BINS * 4 + 1024;
}
size_t sph_sort_temp_bytes(int64_t n) {
    const size_t a = align256((size_t)(cdiv(n + 1, SC_TILE) + 2) * 4) + 1024, b = onesweep_temp_bytes(n);   // scan | sort
    return a > b ? a : b;
}

cudaError_t sph_exclusive_scan(const int *in, int *out, int64_t n, void *temp, size_t temp_bytes, cudaStream_t st) {
    const int ntiles = (int)cdiv(n, SC_TILE);
    if ((size_t)(ntiles + 2) * 4 > temp_bytes) return cudaErrorInvalidValue;
    int *tile_sum = (int *)temp;
    sph_note(3);
    scan_tile_sums<<<ntiles, SC_THREADS, 0, st>>>(in, n, tile_sum);
    scan_of_sums<<<1, 1024, 0, st>>>(tile_sum, ntiles);
    scan_apply<<<ntiles, SC_THREADS, 0, st>>>(in, out, n, tile_sum, ntiles);
    return cudaGetLastError();
}

cudaError_t sph_sort_pairs(uint64_t *keys_in, int *vals_in, uint64_t *keys_out, int *vals_out, int64_t n,
                           const unsigned long long *n_dev, int begin_bit, int end_bit, void *temp,
                           size_t temp_bytes, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    if (temp_bytes < sph_sort_temp_bytes(n)) return cudaErrorInvalidValue;
    const int npass = (end_bit - begin_bit + RS_BITS - 1) / RS_BITS;
    if (npass > 8) return cudaErrorInvalidValue;
    const int ntiles = (int)cdiv(n, OS_TILE);
    unsigned *ghist = (unsigned *)temp;
    unsigned *counters = ghist + 8 * RS_BINS;
    unsigned *status = (unsigned *)((char *)temp + align256((size_t)(8 * RS_BINS + 8) * 4));
    cudaMemsetAsync(temp, 0, align256((size_t)(8 * RS_BINS + 8) * 4) + (size_t)npass * ntiles * RS_BINS * 4, st);
    sph_note(2 + npass);
    int hb = (int)cdiv(n, RS_THREADS * 16);
    hb = hb < 1 ? 1 : (hb > 148 * 8 ? 148 * 8 : hb);
    os_hist_kernel<<<hb, RS_THREADS, 0, st>>>(keys_in, n, n_dev, begin_bit, npass, ghist);
    os_prefix_kernel<<<1, RS_THREADS, 0, st>>>(ghist, npass);
    uint64_t *ka = keys_in, *kb = keys_out;
    int *va = vals_in, *vb = vals_out;
    for (int p = 0; p < npass; ++p) {
        os_pass_kernel<<<ntiles, RS_THREADS, 0, st>>>(ka, va, kb, vb, n, n_dev, begin_bit + p * RS_BITS, ghist + p * RS_BINS,
                                                      status + (size_t)p * ntiles * RS_BINS, counters + p);
        uint64_t *tk = ka; ka = kb; kb = tk;
        int *tv = va; va = vb; vb = tv;
    }
    if (ka != keys_out) {  // result currently in keys_in/vals_in
        cudaMemcpyAsync(keys_out, ka, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(vals_out, va, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
    }
    return cudaGetLastError();
}
