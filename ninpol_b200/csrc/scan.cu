// scan.cu — prefix sums, maxima and stream compaction (CUB device primitives compiled into this
// library; the only translation unit that instantiates CUB, to keep build times down).
// These are the "scan" steps of the counting-sort / two-pass-emit formulations of K1 and K3.
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include "common.cuh"

int npb_exclusive_scan_i32(npb_ctx *c, const int32_t *in, int32_t *out, i64 n)
{
    if (n <= 0) return NPB_OK;
    if (n >= (1ll << 31)) {
        npb_set_error("scan length %lld exceeds the 32-bit range", n);
        return NPB_ERR_RANGE;
    }
    size_t need = 0;
    NPB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n, c->stream));
    NPB_TRY(npb_ensure(&c->scratch, &c->scratch_cap, need));
    NPB_CUDA(cub::DeviceScan::ExclusiveSum(c->scratch, need, in, out, (int)n, c->stream));
    c->launches += 2;  // CUB's decoupled look-back scan: init + scan kernels
    return NPB_OK;
}

int npb_max_i32(npb_ctx *c, const int32_t *in, i64 n, int32_t *host_out)
{
    *host_out = 0;
    if (n <= 0) return NPB_OK;
    size_t need = 0;
    int32_t *d_out = (int32_t *)(c->counters + 16);
    NPB_CUDA(cub::DeviceReduce::Max(nullptr, need, in, d_out, (int)n, c->stream));
    NPB_TRY(npb_ensure(&c->scratch, &c->scratch_cap, need));
    NPB_CUDA(cub::DeviceReduce::Max(c->scratch, need, in, d_out, (int)n, c->stream));
    c->launches += 2;
    NPB_CUDA(cudaMemcpyAsync(host_out, d_out, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    NPB_CUDA(cudaStreamSynchronize(c->stream));
    return NPB_OK;
}

int npb_minmax_i32(npb_ctx *c, const int32_t *in, i64 n, int32_t *h_min, int32_t *h_max)
{
    *h_min = 0;
    *h_max = -1;
    if (n <= 0) return NPB_OK;
    size_t need = 0, need2 = 0;
    int32_t *d_out = (int32_t *)(c->counters + 16);
    NPB_CUDA(cub::DeviceReduce::Max(nullptr, need, in, d_out, (int)n, c->stream));
    NPB_CUDA(cub::DeviceReduce::Min(nullptr, need2, in, d_out + 1, (int)n, c->stream));
    if (need2 > need) need = need2;
    NPB_TRY(npb_ensure(&c->scratch, &c->scratch_cap, need));
    NPB_CUDA(cub::DeviceReduce::Max(c->scratch, need, in, d_out, (int)n, c->stream));
    NPB_CUDA(cub::DeviceReduce::Min(c->scratch, need, in, d_out + 1, (int)n, c->stream));
    c->launches += 4;
    int32_t h[2];
    NPB_CUDA(cudaMemcpyAsync(h, d_out, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    NPB_CUDA(cudaStreamSynchronize(c->stream));
    *h_max = h[0];
    *h_min = h[1];
    return NPB_OK;
}

struct ClassIs {
    const uint8_t *cls;
    int which;
    __host__ __device__ bool operator()(const int &i) const { return cls[i] == which; }
};

// out <- ascending node ids p in [lo, hi) with cls[p] == which
int npb_select_class(npb_ctx *c, const uint8_t *cls, i64 lo, i64 hi, int which, int32_t *out, int *host_count)
{
    *host_count = 0;
    if (hi <= lo) return NPB_OK;
    thrust::counting_iterator<int> it((int)lo);
    ClassIs pred{cls, which};
    int *d_num = c->d_small + 1;   // mapped host memory: the count needs no D2H copy
    size_t need = 0;
    NPB_CUDA(cub::DeviceSelect::If(nullptr, need, it, out, d_num, (int)(hi - lo), pred, c->stream));
    NPB_TRY(npb_ensure(&c->scratch, &c->scratch_cap, need));
    NPB_CUDA(cub::DeviceSelect::If(c->scratch, need, it, out, d_num, (int)(hi - lo), pred, c->stream));
    c->launches += 2;
    NPB_CUDA(cudaStreamSynchronize(c->stream));
    *host_count = c->h_small[1];
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
