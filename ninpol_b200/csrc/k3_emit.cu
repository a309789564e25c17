// k3_emit.cu — two-pass CSR emit (kernel group K3).
//
// Replaces the COO fill loop of Interpolator.interpolate (ninpol/_interpolator/interpolator.pyx:598-618)
// and scipy's coo->csr + eliminate_zeros (:622-624).  Pass 1 (the K2 kernels) leaves the final values
// `w + neumann_ws` esup-indexed in wbuf and the number of surviving entries per row in rowcnt; an
// exclusive scan turns the counts into indptr; pass 2 (here) writes the surviving (column, value)
// pairs of every row at indptr[row] in esup order — esup rows are ascending, so the indices come out
// sorted exactly as scipy's canonical CSR has them.  Entries equal to +-0.0 are dropped, NaN is kept
// (SURVEY.md Q5).  Outputs are int32 indices / float64 data, the dtypes scipy picks.
#include "common.cuh"

// W lanes per row (W = 32, or 8 for meshes whose stars have <= 16 elements: four rows per warp): the lanes
// stride over the row and ballot-compact the non-zeros
template <int W>
__global__ void __launch_bounds__(256)
k_emit_rows(const int32_t *__restrict__ esup_ptr, const int32_t *__restrict__ esup, const double *__restrict__ wbuf,
            i64 wbase, const int32_t *__restrict__ indptr, i64 lo, i64 hi, int32_t *__restrict__ indices,
            double *__restrict__ data, const int32_t *__restrict__ rowcnt, int *__restrict__ mismatch)
{
    const i64 gt = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & (W - 1);
    const int shift = (threadIdx.x & 31) & ~(W - 1);          // first warp lane of this row's group
    const unsigned gmask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << shift);
    i64 p = lo + gt / W;
    const bool live = p < hi;                                  // whole groups are live or not; ballots need every lane
    int out0 = 0, b = 0, e = 0;
    if (live) {
        out0 = indptr[p];
        const int planned = indptr[p + 1] - out0;
        if (planned != 0) {
            b = esup_ptr[p];
            e = esup_ptr[p + 1];
        }
        // planned mode (pipeline.cu): indptr was fixed before the weights were known; a row that kept fewer or
        // more entries than planned voids the plan (and is not written)
        if (rowcnt && rowcnt[p] != planned) {
            if (lane == 0) atomicAdd(mismatch, 1);
            e = b;
        }
    }
    // all groups of a warp iterate together (the ballot is warp wide): up to the longest row of the warp
    int len = e - b;
#pragma unroll
    for (int o = 16; o >= W; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    int written = 0;
    for (int k0 = 0; k0 < len; k0 += W) {
        int q = b + k0 + lane;
        double v = 0.0;
        int col = 0;
        if (q < e) {
            v = wbuf[(i64)q - wbase];
            col = esup[q];
        }
        bool keep = (q < e) && (v != 0.0);
        unsigned m = __ballot_sync(0xffffffffu, keep) & gmask;
        if (keep) {
            int pos = out0 + written + __popc(m & ((1u << (shift + lane)) - 1u));
            indices[pos] = col;
            data[pos] = v;
        }
        written += __popc(m);
    }
}

static int emit(npb_ctx *c, i64 lo, i64 hi, const int32_t *rowcnt, int *mismatch)
{
    if (hi <= lo) return NPB_OK;
    const int T = 256;
    if (c->mx_epp <= 16) {
        i64 threads = (hi - lo) * 8;
        k_emit_rows<8><<<npb_blocks(threads, T), T, 0, c->stream>>>(c->esup_ptr, c->esup, c->wbuf, c->wbase, c->indptr, lo, hi,
                                                                   c->indices, c->data, rowcnt, mismatch);
    } else {
        i64 threads = (hi - lo) * 32;
        k_emit_rows<32><<<npb_blocks(threads, T), T, 0, c->stream>>>(c->esup_ptr, c->esup, c->wbuf, c->wbase, c->indptr, lo, hi,
                                                                    c->indices, c->data, rowcnt, mismatch);
    }
    NPB_LAUNCH(c);
    NPB_CUDA(cudaGetLastError());
    return NPB_OK;
}

int npb_k3_fill(npb_ctx *c, i64 lo, i64 hi) { return emit(c, lo, hi, nullptr, nullptr); }

// rows [lo, hi) at the offsets of the optimistic plan; *mismatch_counter counts rows whose surviving entries
// differ from the planned length (an exact-zero weight): the caller then discards the result
int npb_k3_fill_planned(npb_ctx *c, i64 lo, i64 hi, int *mismatch_counter)
{
    return emit(c, lo, hi, c->rowcnt, mismatch_counter);
}
