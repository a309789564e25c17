// k2_gls.cu — per-node GLS weight construction (kernel group K2, tolerance method).
//
// Replaces GLSInterpolation.GLS and its helpers (ninpol/_methods/gls.pyx:75-474): build_ks_sv_arrays
// (:234-249), build_ls_matrices (:252-356), set_neumann_rows (:374-416), solve_ls (:420-474, LAPACK
// DGELS on an (E+3F+B) x (3E+1) system with E right-hand sides, keeping only the last row of X).
//
// Algebra: with M = [A | c] (A = first 3E columns, c = last column = 1 on the E element
// rows, 0 elsewhere) and r = c - A argmin_g |A g - c| the residual of ONE least-squares problem,
//     weights_i = X[3E, i] = r_i / |r|^2 = r_i / sum_j r_j          (i < E),
// so one Householder QR of A with the single right-hand side c replaces the reference's E solves.
// The all-zero rows the reference leaves for boundary faces (gls.pyx:340-344) do not change a
// least-squares solution and are not materialised.  Reproduced reference behaviour: Q3 (neumann[p] is
// the weight of the node's LAST element, not the Neumann term; the Neumann right-hand side is dead),
// Q4 (that value is added to every weight of the row), Q8 (all faces boundary -> zero row).
//
// Mapping: one warp per node, the dense system lives in shared memory (row-major, lanes over columns,
// loops over rows), nodes are bucketed by workspace size so small stars (hex, boundary) get many
// resident warps and large stars fall back to a global-memory workspace.  FP64-FMA bound, not HBM
// bound.
#include "common.cuh"

#define GLS_NCLASS 6
// workspace caps (bytes) of classes 1..4; class 5 = global-memory workspace; class 0 = skipped node
__constant__ int c_gls_cap[GLS_NCLASS] = {0, 12 * 1024, 40 * 1024, 80 * 1024, 112 * 1024, 0};
static const int h_gls_cap[GLS_NCLASS] = {0, 12 * 1024, 40 * 1024, 80 * 1024, 112 * 1024, 0};

struct GlsArgs {
    const int32_t *esup_ptr, *esup, *fsup_ptr, *fsup;
    const int2 *esuf2;
    const uint8_t *bpoint, *nflag;
    const double *coords, *cent, *fcent, *fnormal, *perm, *diff_mag;
    double *wbuf;
    int32_t *rowcnt;
    double *neumann;
    i64 wbase;
};

__host__ __device__ __forceinline__ size_t gls_ws_bytes(int E, int m)
{
    // M [m, 3E+1] + vv [m] + rinv/g [3E+1] doubles, then es [E] ints (padded to 8 bytes)
    size_t d = (size_t)m * (3 * E + 1) + m + (3 * E + 1);
    return d * 8 + (((size_t)E * 4 + 7) & ~(size_t)7);
}

// per node: Dirichlet / Q8 nodes are finished here (zero row); the others get a size class
__global__ void k_gls_classify(GlsArgs a, i64 lo, i64 hi, uint8_t *__restrict__ cls)
{
    i64 p = lo + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hi) return;
    int eb = a.esup_ptr[p], ee = a.esup_ptr[p + 1];
    int fb = a.fsup_ptr[p], fe = a.fsup_ptr[p + 1];
    int E = ee - eb, F = fe - fb;
    bool neu = a.nflag[p] != 0;
    int nb = 0;
    bool skip = (a.bpoint[p] && !neu);  // gls.pyx:165-166
    if (!skip) {
        for (int q = fb; q < fe; q++) nb += (a.esuf2[a.fsup[q]].y < 0) ? 1 : 0;
        if (nb >= F) skip = true;  // gls.pyx:266-267 + DGELS on a zero matrix -> zero weights (Q8)
    }
    if (skip) {
        double *w = a.wbuf + ((i64)eb - a.wbase);
        for (int k = 0; k < E; k++) w[k] = 0.0;
        a.rowcnt[p] = 0;
        a.neumann[p] = 0.0;
        cls[p] = 0;
        return;
    }
    int m = E + 3 * (F - nb) + (neu ? nb : 0);
    size_t need = gls_ws_bytes(E, m);
    int k = GLS_NCLASS - 1;
    for (int q = 1; q < GLS_NCLASS - 1; q++)
        if (need <= (size_t)c_gls_cap[q]) {
            k = q;
            break;
        }
    cls[p] = (uint8_t)k;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per node.  ws = workspace of this warp (shared or global), generic address space.
__device__ void gls_node(const GlsArgs &a, int p, double *ws)
{
    const int lane = threadIdx.x & 31;
    const int eb = a.esup_ptr[p], E = a.esup_ptr[p + 1] - eb;
    const int fb = a.fsup_ptr[p], F = a.fsup_ptr[p + 1] - fb;
    const bool neu = a.nflag[p] != 0;
    const int n = 3 * E, ld = n + 1;
    const double xv0 = a.coords[(i64)p * 3 + 0], xv1 = a.coords[(i64)p * 3 + 1], xv2 = a.coords[(i64)p * 3 + 2];

    // ---- pass 0: count interior / boundary faces ----
    int n_if = 0;
    for (int f0 = 0; f0 < F; f0 += 32) {
        int fi = f0 + lane;
        bool interior = false;
        if (fi < F) interior = a.esuf2[a.fsup[fb + fi]].y >= 0;
        n_if += __popc(__ballot_sync(0xffffffffu, interior));
    }
    const int n_bf = F - n_if;
    const int m = E + 3 * n_if + (neu ? n_bf : 0);

    double *M = ws;
    double *vv = M + (size_t)m * ld;
    double *gg = vv + m;             // [ld]: reciprocal diagonal, then the solution g
    int *es = (int *)(gg + ld);      // [E]: the node's esup row

    for (int i = lane; i < m * ld; i += 32) M[i] = 0.0;
    for (int i = lane; i < E; i += 32) es[i] = a.esup[eb + i];
    __syncwarp();

    // ---- element rows (gls.pyx:268-281): [ (x_K - x_v)^T at block i | 1 ] ----
    for (int i = lane; i < E; i += 32) {
        const double *cc = a.cent + (i64)es[i] * 3;
        double *row = M + (size_t)i * ld;
        row[3 * i + 0] = cc[0] - xv0;
        row[3 * i + 1] = cc[1] - xv1;
        row[3 * i + 2] = cc[2] - xv2;
        row[n] = 1.0;
    }
    // ---- face rows (gls.pyx:291-356) and Neumann rows (:394-416) ----
    int if_seen = 0, bf_seen = 0;
    for (int f0 = 0; f0 < F; f0 += 32) {
        int fi = f0 + lane;
        int face = -1;
        int2 e2 = make_int2(-1, -1);
        if (fi < F) {
            face = a.fsup[fb + fi];
            e2 = a.esuf2[face];
        }
        bool interior = (fi < F) && e2.y >= 0;
        bool boundary = (fi < F) && e2.y < 0;
        unsigned mi = __ballot_sync(0xffffffffu, interior);
        unsigned mb = __ballot_sync(0xffffffffu, boundary);
        unsigned below = (1u << lane) - 1u;
        if (interior) {
            int j = if_seen + __popc(mi & below);
            int I1 = 0, I2 = 0;
            for (int k = 0; k < E; k++) {
                int ek = es[k];
                if (ek == e2.x) I1 = k;
                if (ek == e2.y) I2 = k;
            }
            const double *Nn = a.fnormal + (i64)face * 3;
            const double *xs = a.fcent + (i64)face * 3;
            double N0 = Nn[0], N1 = Nn[1], N2 = Nn[2];
            double t0 = xv0 - xs[0], t1 = xv1 - xs[1], t2 = xv2 - xs[2];   // T1 = x_v - x_S
            double c0 = N1 * t2 - N2 * t1, c1 = N2 * t0 - N0 * t2, c2 = N0 * t1 - N1 * t0;  // T2 = N x T1
            double eta = fmax(fmax(0.0, a.diff_mag[e2.x]), a.diff_mag[e2.y]);
            double tau = pow(sqrt(c0 * c0 + c1 * c1 + c2 * c2), -eta);
            const double *K1 = a.perm + (i64)e2.x * 9;
            const double *K2 = a.perm + (i64)e2.y * 9;
            double *r1 = M + (size_t)(E + 3 * j) * ld;
            double *r2 = r1 + ld;
            double *r3 = r2 + ld;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                double k1n = K1[3 * q] * N0 + K1[3 * q + 1] * N1 + K1[3 * q + 2] * N2;
                double k2n = K2[3 * q] * N0 + K2[3 * q + 1] * N1 + K2[3 * q + 2] * N2;
                r1[3 * I1 + q] = -k1n;
                r1[3 * I2 + q] = k2n;
            }
            r2[3 * I1 + 0] = -t0; r2[3 * I1 + 1] = -t1; r2[3 * I1 + 2] = -t2;
            r2[3 * I2 + 0] = t0;  r2[3 * I2 + 1] = t1;  r2[3 * I2 + 2] = t2;
            r3[3 * I1 + 0] = -(tau * c0); r3[3 * I1 + 1] = -(tau * c1); r3[3 * I1 + 2] = -(tau * c2);
            r3[3 * I2 + 0] = tau * c0;    r3[3 * I2 + 1] = tau * c1;    r3[3 * I2 + 2] = tau * c2;
        }
        if (boundary && neu) {
            int j = bf_seen + __popc(mb & below);
            int Ik = 0;
            for (int k = 0; k < E; k++)
                if (es[k] == e2.x) Ik = k;
            const double *Nn = a.fnormal + (i64)face * 3;
            const double *K1 = a.perm + (i64)e2.x * 9;
            double N0 = Nn[0], N1 = Nn[1], N2 = Nn[2];
            double *rr = M + (size_t)(E + 3 * n_if + j) * ld;
#pragma unroll
            for (int q = 0; q < 3; q++) rr[3 * Ik + q] = -(K1[3 * q] * N0 + K1[3 * q + 1] * N1 + K1[3 * q + 2] * N2);
        }
        if_seen += __popc(mi);
        bf_seen += __popc(mb);
    }
    __syncwarp();

    // ---- Householder QR of [A | c], natural column order ----
    const int kmax = n < m ? n : m;
    for (int k = 0; k < kmax; k++) {
        double part = 0.0;
        for (int r = k + lane; r < m; r += 32) {
            double x = M[(size_t)r * ld + k];
            vv[r] = x;
            part += x * x;
        }
        double sigma = warp_sum(part);
        __syncwarp();
        if (sigma == 0.0) {
            if (lane == 0) gg[k] = 0.0;
            continue;
        }
        double x0 = vv[k];
        double alpha = (x0 >= 0.0) ? -sqrt(sigma) : sqrt(sigma);
        double beta = 1.0 / (sigma - x0 * alpha);
        __syncwarp();
        if (lane == 0) {
            vv[k] = x0 - alpha;
            M[(size_t)k * ld + k] = alpha;
            gg[k] = 1.0 / alpha;
        }
        __syncwarp();
        for (int j0 = k + 1; j0 < ld; j0 += 32) {
            int j = j0 + lane;
            if (j < ld) {
                double s0 = 0.0, s1 = 0.0;
                int r = k;
                for (; r + 1 < m; r += 2) {
                    s0 += vv[r] * M[(size_t)r * ld + j];
                    s1 += vv[r + 1] * M[(size_t)(r + 1) * ld + j];
                }
                if (r < m) s0 += vv[r] * M[(size_t)r * ld + j];
                double s = (s0 + s1) * beta;
                for (r = k; r < m; r++) M[(size_t)r * ld + j] -= s * vv[r];
            }
        }
        __syncwarp();
    }
    // ---- back substitution R g = z (column oriented; z lives in column n) ----
    for (int k = kmax - 1; k >= 0; k--) {
        double gk = M[(size_t)k * ld + n] * gg[k];
        __syncwarp();
        if (lane == 0) gg[k] = gk;
        for (int r = lane; r < k; r += 32) M[(size_t)r * ld + n] -= M[(size_t)r * ld + k] * gk;
        __syncwarp();
    }
    for (int k = kmax + lane; k < n; k += 32) gg[k] = 0.0;
    __syncwarp();
    // ---- residual on the element rows, weights, CSR values ----
    double *w = a.wbuf + ((i64)eb - a.wbase);
    double part = 0.0;
    for (int i = lane; i < E; i += 32) {
        const double *cc = a.cent + (i64)es[i] * 3;
        double ri = 1.0 - ((cc[0] - xv0) * gg[3 * i] + (cc[1] - xv1) * gg[3 * i + 1] + (cc[2] - xv2) * gg[3 * i + 2]);
        vv[i] = ri;
        part += ri;
    }
    double tot = warp_sum(part);
    __syncwarp();
    double nv = neu ? vv[E - 1] / tot : 0.0;   // gls.pyx:470-472 (Q3)
    int cnt = 0;
    for (int i = lane; i < E; i += 32) {
        double v = vv[i] / tot + nv;           // interpolator.pyx:618 (Q4)
        w[i] = v;
        cnt += (v != 0.0) ? 1 : 0;
    }
    cnt = (int)warp_sum((double)cnt);
    if (lane == 0) {
        a.rowcnt[p] = cnt;
        a.neumann[p] = nv;
    }
}

// persistent: one warp per CTA; nodes handed out through an atomic counter
__global__ void __launch_bounds__(32)
k_gls_nodes(GlsArgs a, const int32_t *__restrict__ list, int count, int *__restrict__ counter, double *gws,
            size_t gws_stride)
{
    extern __shared__ double smem_ws[];
    double *ws = gws ? (double *)((char *)gws + (size_t)blockIdx.x * gws_stride) : smem_ws;
    while (true) {
        int i = 0;
        if (threadIdx.x == 0) i = atomicAdd(counter, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= count) break;
        gls_node(a, list[i], ws);
        __syncwarp();
    }
}

int npb_k2_gls(npb_ctx *c, i64 lo, i64 hi)
{
    if (hi <= lo) return NPB_OK;
    cudaStream_t s = c->stream;
    GlsArgs a;
    a.esup_ptr = c->esup_ptr; a.esup = c->esup; a.fsup_ptr = c->fsup_ptr; a.fsup = c->fsup; a.esuf2 = c->esuf2;
    a.bpoint = c->bpoint; a.nflag = c->nflag; a.coords = c->coords; a.cent = c->centroids; a.fcent = c->fcent;
    a.fnormal = c->fnormal; a.perm = c->perm; a.diff_mag = c->diff_mag; a.wbuf = c->wbuf; a.rowcnt = c->rowcnt;
    a.neumann = c->neumann; a.wbase = c->wbase;
    i64 nloc = hi - lo;
    uint8_t *cls = nullptr;
    NPB_CUDA(cudaMalloc(&cls, (size_t)c->n_points));
    k_gls_classify<<<npb_blocks(nloc, 256), 256, 0, s>>>(a, lo, hi, cls);
    NPB_LAUNCH(c);
    if (!c->node_list) NPB_TRY(npb_alloc(c, (void **)&c->node_list, sizeof(int32_t) * (size_t)c->n_points));
    for (int k = 1; k < GLS_NCLASS; k++) {
        int count = 0;
        NPB_TRY(npb_select_class(c, cls, lo, hi, k, c->node_list, &count));
        if (count == 0) continue;
        int *counter = c->counters + 20 + k;
        NPB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), s));
        if (k < GLS_NCLASS - 1) {
            int smem = h_gls_cap[k];
            NPB_CUDA(cudaFuncSetAttribute(k_gls_nodes, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            int per_sm = (int)((227 * 1024) / (smem + 1024));
            if (per_sm > 32) per_sm = 32;
            int grid = c->sm_count * per_sm;
            if (grid > count) grid = count;
            k_gls_nodes<<<grid, 32, smem, s>>>(a, c->node_list, count, counter, nullptr, 0);
        } else {
            // oversized stars: global-memory workspace, sized for the largest possible system
            int E = c->mx_epp, F = c->mx_fpp;
            size_t stride = (gls_ws_bytes(E, E + 4 * F) + 255) & ~(size_t)255;
            int grid = c->sm_count * 8;
            if (grid > count) grid = count;
            NPB_TRY(npb_ensure(&c->gls_ws, &c->gls_ws_cap, stride * (size_t)grid));
            k_gls_nodes<<<grid, 32, 0, s>>>(a, c->node_list, count, counter, (double *)c->gls_ws, stride);
        }
        NPB_LAUNCH(c);
        NPB_CUDA(cudaGetLastError());
        // node_list is reused by the next class: the select below is stream-ordered after this kernel
    }
    NPB_CUDA(cudaStreamSynchronize(s));
    NPB_CUDA(cudaFree(cls));
    return NPB_OK;
}
