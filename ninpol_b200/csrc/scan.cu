// scan.cu — prefix sums, min / max and the stable class partition: the "scan" steps of the counting-sort /
// planned-emit formulations of K1, K2 (GLS work lists) and K3, hand-written (tile scans + recursive scan of the tile
// totals; warp-ballot ranking).  CUB remains only in the two helpers of the optional edge structures
// (build_edges=True: radix sort of the edge hashes, segmented head propagation) at the end of this file.
#include <limits.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// exclusive prefix sum of int32 (out may alias in): tile-local scans + a recursive scan of the tile totals
// ------------------------------------------------------------------------------------------------
#define SCAN_T 256
#define SCAN_PER 16
#define SCAN_TILE (SCAN_T * SCAN_PER)

__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &block_total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < SCAN_T / 32 ? s_warp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < SCAN_T / 32) s_warp[lane] = wi - w;   // exclusive warp offsets
        if (lane == SCAN_T / 32 - 1) s_warp[SCAN_T / 32] = wi;
    }
    __syncthreads();
    block_total = s_warp[SCAN_T / 32];
    return s_warp[wid] + incl - v;
}

// each CTA: exclusive scan of its tile (thread t owns SCAN_PER consecutive items), tile total to tile_sum[blockIdx]
__global__ void __launch_bounds__(SCAN_T) k_scan_tiles(const int32_t *__restrict__ in, int32_t *__restrict__ out, i64 n,
                                                       int32_t *__restrict__ tile_sum)
{
    __shared__ int s_warp[SCAN_T / 32 + 1];
    const i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_PER;
    int v[SCAN_PER];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    int total;
    int run = block_exclusive_scan(sum, s_warp, total);
#pragma unroll
    for (int k = 0; k < SCAN_PER; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_add(int32_t *__restrict__ out, i64 n, const int32_t *__restrict__ tile_off)
{
    const i64 tile = (i64)blockIdx.x + 1;        // tile 0 needs no offset
    const int off = tile_off[tile];
    const i64 base = tile * SCAN_TILE;
    for (int k = threadIdx.x; k < SCAN_TILE; k += SCAN_T)
        if (base + k < n) out[base + k] += off;
}

static int scan_rec(npb_ctx *c, const int32_t *in, int32_t *out, i64 n, int32_t *tmp)
{
    const i64 nt = (n + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<<<(unsigned)nt, SCAN_T, 0, c->stream>>>(in, out, n, tmp);
    NPB_LAUNCH(c);
    if (nt > 1) {
        NPB_TRY(scan_rec(c, tmp, tmp, nt, tmp + ((nt + 63) & ~(i64)63)));
        k_scan_add<<<(unsigned)(nt - 1), SCAN_T, 0, c->stream>>>(out, n, tmp);
        NPB_LAUNCH(c);
    }
    return NPB_OK;
}

int npb_exclusive_scan_i32(npb_ctx *c, const int32_t *in, int32_t *out, i64 n)
{
    if (n <= 0) return NPB_OK;
    if (n >= (1ll << 31)) {
        npb_set_error("scan length %lld exceeds the 32-bit range", n);
        return NPB_ERR_RANGE;
    }
    // tile totals of every level: n / 4096 + n / 4096^2 + ... (+ alignment slack)
    size_t need = 0;
    for (i64 m = n; m > 1;) {
        m = (m + SCAN_TILE - 1) / SCAN_TILE;
        need += (size_t)((m + 63) & ~(i64)63);
        if (m == 1) break;
    }
    NPB_TRY(npb_ensure(&c->scan_tmp, &c->scan_tmp_cap, sizeof(int32_t) * (need + 64)));
    return scan_rec(c, in, out, n, (int32_t *)c->scan_tmp);
}

// ------------------------------------------------------------------------------------------------
// min / max of int32
// ------------------------------------------------------------------------------------------------
__global__ void k_minmax_init(int32_t *mm)
{
    mm[0] = INT32_MIN;   // max
    mm[1] = INT32_MAX;   // min
}
__global__ void __launch_bounds__(256) k_minmax(const int32_t *__restrict__ in, i64 n, int32_t *__restrict__ mm)
{
    int mx = INT32_MIN, mn = INT32_MAX;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        int v = in[i];
        mx = max(mx, v);
        mn = min(mn, v);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    mn = __reduce_min_sync(0xffffffffu, mn);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&mm[0], mx);
        atomicMin(&mm[1], mn);
    }
}

int npb_minmax_i32(npb_ctx *c, const int32_t *in, i64 n, int32_t *h_min, int32_t *h_max)
{
    *h_min = 0;
    *h_max = -1;
    if (n <= 0) return NPB_OK;
    int32_t *d_out = (int32_t *)(c->counters + 16);
    k_minmax_init<<<1, 1, 0, c->stream>>>(d_out);
    i64 blocks = (n + 255) / 256;
    if (blocks > (i64)c->sm_count * 16) blocks = (i64)c->sm_count * 16;
    k_minmax<<<(unsigned)blocks, 256, 0, c->stream>>>(in, n, d_out);
    c->launches += 2;
    int32_t h[2];
    NPB_CUDA(cudaMemcpyAsync(h, d_out, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    NPB_CUDA(cudaStreamSynchronize(c->stream));
    *h_max = h[0];
    *h_min = h[1];
    return NPB_OK;
}

int npb_max_i32(npb_ctx *c, const int32_t *in, i64 n, int32_t *host_out)
{
    int32_t mn = 0, mx = 0;
    *host_out = 0;
    if (n <= 0) return NPB_OK;
    NPB_TRY(npb_minmax_i32(c, in, n, &mn, &mx));
    *host_out = mx;
    return NPB_OK;
}

// ------------------------------------------------------------------------------------------------
// Stable partition of the nodes [lo, hi) by size class (GLS work lists): one histogram pass, one scan of the
// (class-major) per-block histograms, one scatter pass.  out[starts[k] .. starts[k+1]) = ascending node ids of
// class k; the NCLS + 1 starts come back to the host through the mapped block with ONE synchronisation
// (round 1 ran a CUB select and a host round trip per class).
// ------------------------------------------------------------------------------------------------
#define PART_T 256
#define PART_ROUNDS 8
#define PART_TILE (PART_T * PART_ROUNDS)
#define PART_NCLS 9

__global__ void __launch_bounds__(PART_T) k_part_hist(const uint8_t *__restrict__ cls, i64 lo, i64 hi, int nblk,
                                                      int32_t *__restrict__ hist /*[PART_NCLS * nblk + 1]*/)
{
    __shared__ int s_h[PART_NCLS];
    if (threadIdx.x < PART_NCLS) s_h[threadIdx.x] = 0;
    __syncthreads();
    const i64 base = lo + (i64)blockIdx.x * PART_TILE;
    for (int r = 0; r < PART_ROUNDS; r++) {
        const i64 p = base + r * PART_T + threadIdx.x;
        if (p < hi) atomicAdd(&s_h[cls[p] < PART_NCLS ? cls[p] : PART_NCLS - 1], 1);
    }
    __syncthreads();
    if (threadIdx.x < PART_NCLS) hist[threadIdx.x * nblk + blockIdx.x] = s_h[threadIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) hist[PART_NCLS * nblk] = 0;
}

__global__ void __launch_bounds__(PART_T) k_part_scatter(const uint8_t *__restrict__ cls, i64 lo, i64 hi, int nblk,
                                                         const int32_t *__restrict__ off /*scanned hist*/, int32_t *__restrict__ out)
{
    __shared__ int s_run[PART_NCLS];
    __shared__ int s_wc[PART_T / 32][PART_NCLS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < PART_NCLS) s_run[threadIdx.x] = off[threadIdx.x * nblk + blockIdx.x];
    const i64 base = lo + (i64)blockIdx.x * PART_TILE;
    for (int r = 0; r < PART_ROUNDS; r++) {
        const i64 p = base + r * PART_T + threadIdx.x;
        const int k = p < hi ? (cls[p] < PART_NCLS ? cls[p] : PART_NCLS - 1) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, k);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (threadIdx.x < (PART_T / 32) * PART_NCLS) (&s_wc[0][0])[threadIdx.x] = 0;
        __syncthreads();
        if (k >= 0 && rank == 0) s_wc[wid][k] = __popc(peers);
        __syncthreads();
        if (k >= 0) {
            int pos = s_run[k] + rank;
            for (int w = 0; w < wid; w++) pos += s_wc[w][k];
            out[pos] = (int32_t)p;
        }
        __syncthreads();
        if (threadIdx.x < PART_NCLS) {
            int add = 0;
            for (int w = 0; w < PART_T / 32; w++) add += s_wc[w][threadIdx.x];
            s_run[threadIdx.x] += add;
        }
        __syncthreads();
    }
}

__global__ void k_part_starts(const int32_t *__restrict__ off, int nblk, int *__restrict__ dst)
{
    if (threadIdx.x <= PART_NCLS) dst[threadIdx.x] = off[threadIdx.x * nblk];
}

int npb_partition_classes(npb_ctx *c, const uint8_t *cls, i64 lo, i64 hi, int32_t *out, int *host_starts /*[PART_NCLS + 1]*/)
{
    for (int k = 0; k <= PART_NCLS; k++) host_starts[k] = 0;
    if (hi <= lo) return NPB_OK;
    const i64 n = hi - lo;
    const int nblk = (int)((n + PART_TILE - 1) / PART_TILE);
    const size_t hist_len = (size_t)PART_NCLS * nblk + 1;
    NPB_TRY(npb_ensure(&c->part_tmp, &c->part_tmp_cap, sizeof(int32_t) * (hist_len + 64)));
    int32_t *hist = (int32_t *)c->part_tmp;
    k_part_hist<<<nblk, PART_T, 0, c->stream>>>(cls, lo, hi, nblk, hist);
    NPB_LAUNCH(c);
    NPB_TRY(npb_exclusive_scan_i32(c, hist, hist, (i64)hist_len));
    k_part_scatter<<<nblk, PART_T, 0, c->stream>>>(cls, lo, hi, nblk, hist, out);
    NPB_LAUNCH(c);
    k_part_starts<<<1, 32, 0, c->stream>>>(hist, nblk, c->d_small + 32);
    NPB_LAUNCH(c);
    NPB_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k <= PART_NCLS; k++) host_starts[k] = c->h_small[32 + k];
    return NPB_OK;
}

// stable sort of (key, value) pairs by 32-bit key (CUB radix sort); outputs in keys_out / vals_out
int npb_sort_pairs_u32(npb_ctx *c, const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *vals_in, uint32_t *vals_out, i64 n)
{
    if (n <= 0) return NPB_OK;
    size_t need = 0;
    NPB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 32, c->stream));
    NPB_TRY(npb_ensure(&c->scratch, &c->scratch_cap, need));
    NPB_CUDA(cub::DeviceRadixSort::SortPairs(c->scratch, need, keys_in, keys_out, vals_in, vals_out, (int)n, 0, 32, c->stream));
    c->launches += 8;
    return NPB_OK;
}

struct HeadCopy {   // segmented "copy the run head's value forward": (flag, value) pairs
    __host__ __device__ uint2 operator()(const uint2 &a, const uint2 &b) const { return b.x ? b : make_uint2(a.x, a.y); }
};

// in/out: pairs (is_head, value); after the call every element carries the value of its run's head
int npb_propagate_heads(npb_ctx *c, uint2 *pairs, i64 n)
{
    if (n <= 0) return NPB_OK;
    size_t need = 0;
    NPB_CUDA(cub::DeviceScan::InclusiveScan(nullptr, need, pairs, pairs, HeadCopy(), (int)n, c->stream));
    NPB_TRY(npb_ensure(&c->scratch, &c->scratch_cap, need));
    NPB_CUDA(cub::DeviceScan::InclusiveScan(c->scratch, need, pairs, pairs, HeadCopy(), (int)n, c->stream));
    c->launches += 2;
    return NPB_OK;
}
