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

// one warp per row: lanes stride over the row, ballot-compact the non-zeros
__global__ void __launch_bounds__(256)
k_emit_rows(const int32_t *__restrict__ esup_ptr, const int32_t *__restrict__ esup, const double *__restrict__ wbuf,
            i64 wbase, const int32_t *__restrict__ indptr, i64 lo, i64 hi, int32_t *__restrict__ indices,
            double *__restrict__ data)
{
    i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    i64 p = lo + warp;
    if (p >= hi) return;
    int out0 = indptr[p];
    int cnt = indptr[p + 1] - out0;
    if (cnt == 0) return;
    int b = esup_ptr[p], e = esup_ptr[p + 1];
    int written = 0;
    for (int q0 = b; q0 < e; q0 += 32) {
        int q = q0 + lane;
        double v = 0.0;
        int col = 0;
        if (q < e) {
            v = wbuf[(i64)q - wbase];
            col = esup[q];
        }
        bool keep = (q < e) && (v != 0.0);
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            int pos = out0 + written + __popc(m & ((1u << lane) - 1u));
            indices[pos] = col;
            data[pos] = v;
        }
        written += __popc(m);
    }
}

int npb_k3_fill(npb_ctx *c, i64 lo, i64 hi)
{
    if (hi <= lo) return NPB_OK;
    const int T = 256;
    i64 threads = (hi - lo) * 32;
    k_emit_rows<<<npb_blocks(threads, T), T, 0, c->stream>>>(c->esup_ptr, c->esup, c->wbuf, c->wbase, c->indptr, lo, hi,
                                                            c->indices, c->data);
    NPB_LAUNCH(c);
    NPB_CUDA(cudaGetLastError());
    return NPB_OK;
}
