// stream.cu — interpolate() as a three-stage pipeline over node chunks (single GPU).
//
// No reference counterpart: the reference computes on the host.  npb_interpolate_count / _fetch run
// "upload everything, compute everything, download everything"; at 50M cells the two PCIe legs (4.1 GB of
// permeability / diff_mag up, 2.5 GB of CSR down) are 120 ms next to 620 ms of GLS.  Here the nodes are
// cut into contiguous chunks and three streams overlap:
//   upload stream   : the slice of the cell fields chunk k+1 reads (the element range of its esup rows,
//                     minus what is already resident) while chunk k computes;
//   compute stream  : K2 (weights of the chunk) -> scan of its row counts on top of the running nnz ->
//                     K3 fill of its CSR block;
//   download stream : the chunk's indptr / indices / data / neumann block into the caller's page-locked
//                     arrays while chunk k+1 computes.
// Every node's weights are computed by the same kernels as in the unchunked path, so the result is
// bit-identical to npb_interpolate_count + npb_interpolate_fetch.  The caller sizes indices / data for the
// upper bound (all esup entries) because the exact nnz is only known at the end.
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

__global__ void k_add_base(int32_t *__restrict__ p, i64 n, int32_t base)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += base;
}

static int d2h_pieces(npb_ctx *c, void *dst, const void *src, size_t bytes)
{
    const size_t piece = (size_t)16 << 20;
    for (size_t off = 0; off < bytes; off += piece) {
        size_t n = bytes - off < piece ? bytes - off : piece;
        NPB_CUDA(cudaMemcpyAsync((char *)dst + off, (const char *)src + off, n, cudaMemcpyDeviceToHost, c->down_stream));
    }
    return NPB_OK;
}

static int ensure_streams(npb_ctx *c, int n_events)
{
    if (!c->up_stream) NPB_CUDA(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
    if (!c->down_stream) NPB_CUDA(cudaStreamCreateWithFlags(&c->down_stream, cudaStreamNonBlocking));
    while ((int)c->pipe_ev.size() < n_events) {
        cudaEvent_t e;
        NPB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->pipe_ev.push_back(e);
    }
    return NPB_OK;
}

extern "C" int npb_interpolate_streamed(npb_ctx *c, int method, int n_chunks, const double *perm_host,
                                        const double *diff_mag_host, int32_t *indptr, int32_t *indices, double *data,
                                        double *neumann, int64_t capacity, int64_t *nnz)
{
    if (!c || !nnz || !indptr || !indices || !data || !neumann) return NPB_ERR_ARG;
    if (!c->mesh_loaded) {
        npb_set_error("Grid not initialized. Please load a mesh first.");
        return NPB_ERR_STATE;
    }
    if (method != NPB_METHOD_IDW && method != NPB_METHOD_LS && method != NPB_METHOD_GLS) {
        npb_set_error("unknown method id %d", method);
        return NPB_ERR_ARG;
    }
    if (c->world != 1) {
        npb_set_error("npb_interpolate_streamed is the single-GPU pipeline; ranks of a multi-GPU run already hold a slice");
        return NPB_ERR_STATE;
    }
    if (!c->have_flags) {
        npb_set_error("neumann flags have not been set");
        return NPB_ERR_STATE;
    }
    const bool gls = method == NPB_METHOD_GLS;
    const bool upload = gls && perm_host && diff_mag_host;
    if (gls && !upload && (!c->have_perm || !c->have_dm)) {
        npb_set_error("GLS needs the 'permeability' and 'diff_mag' cell fields");
        return NPB_ERR_STATE;
    }
    if (!npb_is_pinned(indptr) || !npb_is_pinned(indices) || !npb_is_pinned(data) || !npb_is_pinned(neumann) ||
        (upload && (!npb_is_pinned(perm_host) || !npb_is_pinned(diff_mag_host)))) {
        npb_set_error("npb_interpolate_streamed needs page-locked host arrays (npb_host_alloc / npb_host_register)");
        return NPB_ERR_ARG;
    }
    NPB_CUDA(cudaSetDevice(c->device));
    const i64 np = c->n_points;
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > 64) n_chunks = 64;
    if ((i64)n_chunks > np) n_chunks = (int)np;
    cudaStream_t s = c->stream;
    NPB_TRY(ensure_streams(c, 2 * n_chunks));
    NPB_TRY(npb_ensure((void **)&c->wbuf, &c->wbuf_cap, sizeof(double) * (size_t)(c->wlen > 0 ? c->wlen : 1)));
    NPB_TRY(npb_ensure_out(c, (size_t)c->wlen));   // upper bound: every esup entry kept
    c->counted = false;
    c->filled = false;

    // chunk bounds (equal node counts) and the element range each chunk reads
    std::vector<i64> lo(n_chunks + 1);
    for (int k = 0; k <= n_chunks; k++) lo[k] = (np * k) / n_chunks;
    std::vector<i64> e_first(n_chunks, 0), e_last(n_chunks, -1);
    if (upload) {
        std::vector<int32_t> ptr(n_chunks + 1);
        for (int k = 0; k <= n_chunks; k++)
            NPB_CUDA(cudaMemcpyAsync(&ptr[k], c->esup_ptr + lo[k], sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        NPB_CUDA(cudaStreamSynchronize(s));
        for (int k = 0; k < n_chunks; k++) {
            int32_t mn = 0, mx = -1;
            NPB_TRY(npb_minmax_i32(c, c->esup + ptr[k], (i64)ptr[k + 1] - ptr[k], &mn, &mx));
            e_first[k] = mn;
            e_last[k] = mx;
        }
        if (!c->perm) NPB_TRY(npb_alloc(c, (void **)&c->perm, sizeof(double) * 9 * (size_t)c->n_elems));
        if (!c->diff_mag) NPB_TRY(npb_alloc(c, (void **)&c->diff_mag, sizeof(double) * (size_t)c->n_elems));
    }
    i64 ulo = 0, uhi = 0;   // resident element interval [ulo, uhi)
    auto enqueue_upload = [&](int k) -> int {
        if (upload && e_last[k] >= e_first[k]) {
            i64 a = e_first[k], b = e_last[k] + 1;
            i64 seg[2][2] = {{0, 0}, {0, 0}};
            if (uhi <= ulo) {
                seg[0][0] = a; seg[0][1] = b;
                ulo = a; uhi = b;
            } else {
                if (a < ulo) { seg[0][0] = a; seg[0][1] = ulo; ulo = a; }
                if (b > uhi) { seg[1][0] = uhi; seg[1][1] = b; uhi = b; }
            }
            for (int q = 0; q < 2; q++) {
                i64 n = seg[q][1] - seg[q][0];
                if (n <= 0) continue;
                NPB_CUDA(cudaMemcpyAsync(c->perm + 9 * seg[q][0], perm_host + 9 * seg[q][0], sizeof(double) * 9 * (size_t)n,
                                         cudaMemcpyHostToDevice, c->up_stream));
                NPB_CUDA(cudaMemcpyAsync(c->diff_mag + seg[q][0], diff_mag_host + seg[q][0], sizeof(double) * (size_t)n,
                                         cudaMemcpyHostToDevice, c->up_stream));
            }
        }
        NPB_CUDA(cudaEventRecord(c->pipe_ev[2 * k], c->up_stream));
        return NPB_OK;
    };

    const bool dbg = getenv("NPB_STREAM_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    NpbTimer tall(c, "streamed");
    NPB_TRY(enqueue_upload(0));
    int32_t base = 0;
    for (int k = 0; k < n_chunks; k++) {
        const i64 a = lo[k], b = lo[k + 1];
        double t0 = dbg ? now() : 0.0;
        if (k + 1 < n_chunks) NPB_TRY(enqueue_upload(k + 1));
        NPB_CUDA(cudaStreamWaitEvent(s, c->pipe_ev[2 * k], 0));
        double t1 = dbg ? now() : 0.0;
        if (gls)
            NPB_TRY(npb_k2_gls(c, a, b));
        else {
            int used = 0;
            NPB_TRY(npb_k2_idw_ls_tiles(c, method, a, b, &used));
            if (!used) NPB_TRY(npb_k2_idw_ls(c, method, a, b));
        }
        if (dbg) cudaStreamSynchronize(s);
        double t2 = dbg ? now() : 0.0;
        // indptr[a..b] = base + exclusive scan of the chunk's row counts
        NPB_CUDA(cudaMemsetAsync(c->rowcnt + b, 0, sizeof(int32_t), s));
        NPB_TRY(npb_exclusive_scan_i32(c, c->rowcnt + a, c->indptr + a, b - a + 1));
        if (base != 0) {
            k_add_base<<<npb_blocks(b - a + 1, 256), 256, 0, s>>>(c->indptr + a, b - a + 1, base);
            NPB_LAUNCH(c);
        }
        int32_t next = 0;
        NPB_TRY(npb_read_int(c, c->indptr + b, &next));
        if ((i64)next > capacity) {
            npb_set_error("npb_interpolate_streamed: capacity %lld too small (nnz so far %d)", (long long)capacity, next);
            cudaStreamSynchronize(c->up_stream);
            cudaStreamSynchronize(c->down_stream);
            return NPB_ERR_ARG;
        }
        double t3 = dbg ? now() : 0.0;
        NPB_TRY(npb_k3_fill(c, a, b));
        if (dbg) cudaStreamSynchronize(s);
        double t4 = dbg ? now() : 0.0;
        NPB_CUDA(cudaEventRecord(c->pipe_ev[2 * k + 1], s));
        NPB_CUDA(cudaStreamWaitEvent(c->down_stream, c->pipe_ev[2 * k + 1], 0));
        const i64 nk = (i64)next - base;
        // bulk downloads in 16 MB pieces: a small copy of another stream waits for one piece, not for the block
        NPB_TRY(d2h_pieces(c, indptr + a, c->indptr + a, sizeof(int32_t) * (size_t)(b - a)));
        NPB_TRY(d2h_pieces(c, neumann + a, c->neumann + a, sizeof(double) * (size_t)(b - a)));
        if (nk > 0) {
            NPB_TRY(d2h_pieces(c, indices + base, c->indices + base, sizeof(int32_t) * (size_t)nk));
            NPB_TRY(d2h_pieces(c, data + base, c->data + base, sizeof(double) * (size_t)nk));
        }
        base = next;
        if (dbg) fprintf(stderr, "[stream] chunk %d: enqueue-up %.2f k2 %.2f scan %.2f fill %.2f enqueue-down %.2f ms\n", k, t1 - t0, t2 - t1, t3 - t2, t4 - t3, now() - t4);
    }
    if (upload) {   // elements no node refers to (none on a valid mesh): keep the resident copy complete
        const i64 rest[2][2] = {{0, uhi > ulo ? ulo : c->n_elems}, {uhi > ulo ? uhi : c->n_elems, c->n_elems}};
        for (int q = 0; q < 2; q++) {
            i64 n = rest[q][1] - rest[q][0];
            if (n <= 0) continue;
            NPB_CUDA(cudaMemcpyAsync(c->perm + 9 * rest[q][0], perm_host + 9 * rest[q][0], sizeof(double) * 9 * (size_t)n,
                                     cudaMemcpyHostToDevice, c->up_stream));
            NPB_CUDA(cudaMemcpyAsync(c->diff_mag + rest[q][0], diff_mag_host + rest[q][0], sizeof(double) * (size_t)n,
                                     cudaMemcpyHostToDevice, c->up_stream));
        }
        c->have_perm = true;
        c->have_dm = true;
    }
    NPB_CUDA(cudaStreamSynchronize(c->up_stream));
    NPB_CUDA(cudaStreamSynchronize(c->down_stream));
    tall.stop();
    indptr[np] = base;
    c->nnz = base;
    c->nnz_ret = base;
    c->blk_off = 0;
    c->method = method;
    *nnz = base;
    return NPB_OK;
}
