// k2_gls_dense.cu — dense fallback of the per-node GLS solve: Householder QR of the whole
// (E + 3 F_int + B) x (3E + 1) system [A | c] in natural column order, one warp per node, workspace in
// global memory (L2-resident).  Used only for node stars that do not fit the multifrontal kernel of
// k2_gls.cu (more than 64 elements, or an arena overflow) — same algebra, same outputs.
// Reference: ninpol/_methods/gls.pyx:252-474 (see k2_gls.cu for the row definitions).
#include "gls_common.cuh"

__host__ __device__ __forceinline__ size_t gls_dense_ws_bytes(int E, int m)
{
    // M [m, 3E+1] + vv [m] + vdiag [3E+1] + beta [3E+1] doubles, then es [E] ints (padded to 8 bytes)
    size_t d = (size_t)m * (3 * E + 1) + m + 2 * (size_t)(3 * E + 1);
    return d * 8 + (((size_t)E * 4 + 7) & ~(size_t)7);
}

__device__ void gls_node_dense(const GlsArgs &a, int p, double *ws)
{
    const int lane = threadIdx.x & 31;
    const int eb = a.esup_ptr[p], E = a.esup_ptr[p + 1] - eb;
    const int fb = a.fsup_ptr[p], F = a.fsup_ptr[p + 1] - fb;
    const bool neu = a.nflag[p] != 0;
    const int n = 3 * E, ld = n + 1;
    const double xv0 = a.coords[(i64)p * 3 + 0], xv1 = a.coords[(i64)p * 3 + 1], xv2 = a.coords[(i64)p * 3 + 2];

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
    double *vd = vv + m;             // [ld]: v_k[k] = x_k - alpha_k (the sub-diagonal part of v_k stays in column k of M)
    double *bt = vd + ld;            // [ld]: beta_k = 2 / |v_k|^2 (0 for a zero column)
    int *es = (int *)(bt + ld);      // [E]: the node's esup row

    for (int i = lane; i < m * ld; i += 32) M[i] = 0.0;
    for (int i = lane; i < E; i += 32) es[i] = a.esup[eb + i];
    __syncwarp();

    // element rows (gls.pyx:268-281): [ (x_K - x_v)^T at block i | 1 ]
    for (int i = lane; i < E; i += 32) {
        const double *cc = a.cent + (i64)es[i] * NPB_CSTRIDE;
        double *row = M + (size_t)i * ld;
        row[3 * i + 0] = cc[0] - xv0;
        row[3 * i + 1] = cc[1] - xv1;
        row[3 * i + 2] = cc[2] - xv2;
        row[n] = 1.0;
    }
    // face rows (gls.pyx:291-356) and Neumann rows (:394-416)
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
            double t0 = xv0 - xs[0], t1 = xv1 - xs[1], t2 = xv2 - xs[2];
            double c0 = gls_cross(N1, t2, N2, t1), c1 = gls_cross(N2, t0, N0, t2), c2 = gls_cross(N0, t1, N1, t0);
            double eta = fmax(fmax(0.0, a.diff_mag[e2.x]), a.diff_mag[e2.y]);
            double tau = pow(gls_norm3(c0, c1, c2), -eta);
            const double *K1 = a.perm + (i64)e2.x * 9;
            const double *K2 = a.perm + (i64)e2.y * 9;
            double *r1 = M + (size_t)(E + 3 * j) * ld;
            double *r2 = r1 + ld;
            double *r3 = r2 + ld;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                double k1n = gls_kn(K1 + 3 * q, N0, N1, N2);
                double k2n = gls_kn(K2 + 3 * q, N0, N1, N2);
                r1[3 * I1 + q] = -k1n;
                r1[3 * I2 + q] = k2n;
            }
            r2[3 * I1 + 0] = -t0; r2[3 * I1 + 1] = -t1; r2[3 * I1 + 2] = -t2;
            r2[3 * I2 + 0] = t0;  r2[3 * I2 + 1] = t1;  r2[3 * I2 + 2] = t2;
            const double tc0 = __dmul_rn(tau, c0), tc1 = __dmul_rn(tau, c1), tc2 = __dmul_rn(tau, c2);
            r3[3 * I1 + 0] = -tc0; r3[3 * I1 + 1] = -tc1; r3[3 * I1 + 2] = -tc2;
            r3[3 * I2 + 0] = tc0;  r3[3 * I2 + 1] = tc1;  r3[3 * I2 + 2] = tc2;
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
            for (int q = 0; q < 3; q++) rr[3 * Ik + q] = -gls_kn(K1 + 3 * q, N0, N1, N2);
        }
        if_seen += __popc(mi);
        bf_seen += __popc(mb);
    }
    __syncwarp();

    // Householder QR of [A | c], natural column order (the order DGELS uses); the reflectors are kept: v_k below the
    // diagonal of column k, its leading entry in vd[k], beta_k in bt[k]
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
            if (lane == 0) {
                vd[k] = 0.0;
                bt[k] = 0.0;
            }
            continue;
        }
        double x0 = vv[k];
        double alpha = (x0 >= 0.0) ? -sqrt(sigma) : sqrt(sigma);
        double beta = 1.0 / (sigma - x0 * alpha);
        __syncwarp();
        if (lane == 0) {
            vv[k] = x0 - alpha;
            vd[k] = x0 - alpha;
            bt[k] = beta;
            M[(size_t)k * ld + k] = alpha;
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
    // The weights are the element-row entries of r / |r|^2, r = c - A g the least-squares residual.  r is formed by
    // applying the reflectors BACKWARDS to what the factorisation left of c below the triangle, r = Q [0; z] - the
    // route DGELS itself takes for its E right-hand sides (X[3E, i] = (Q^T e_i)[3E] / rho) - and |r|^2 = |z|^2 is a
    // sum of squares.  Forming r as 1 - d_i . g_i from the back-substituted g, and |r|^2 as sum_i r_i, loses digits to
    // cancellation wherever the system is nearly consistent (one-sided stars at Neumann boundary nodes: weights of
    // +-75 on a 2-D corner, error 2e-9 instead of 1e-14).
    double zz = 0.0;
    for (int r = lane; r < m; r += 32) {
        double y = (r >= kmax) ? M[(size_t)r * ld + n] : 0.0;
        vv[r] = y;
        zz += y * y;
    }
    const double rho2 = warp_sum(zz);
    __syncwarp();
    for (int k = kmax - 1; k >= 0; k--) {
        const double beta = bt[k];
        if (beta == 0.0) continue;     // warp-uniform
        double part = 0.0;
        for (int r = k + 1 + lane; r < m; r += 32) part += M[(size_t)r * ld + k] * vv[r];
        double dot = warp_sum(part) + vd[k] * vv[k];
        const double sc = beta * dot;
        __syncwarp();
        for (int r = k + 1 + lane; r < m; r += 32) vv[r] -= sc * M[(size_t)r * ld + k];
        if (lane == 0) vv[k] -= sc * vd[k];
        __syncwarp();
    }
    // weights, CSR values
    double *w = a.wbuf + ((i64)eb - a.wbase);
    const double tot = rho2;
    double nv = neu ? vv[E - 1] / tot : 0.0;   // gls.pyx:470-472 (Q3)
    int cnt = 0;
    for (int i = lane; i < E; i += 32) {
        double v = vv[i] / tot + nv;           // interpolator.pyx:618 (Q4)
        w[i] = v;
        cnt += (v != 0.0) ? 1 : 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) {
        a.rowcnt[p] = cnt;
        a.neumann[p] = nv;
    }
}

__global__ void __launch_bounds__(32)
k_gls_dense(GlsArgs a, const int32_t *__restrict__ list, int count, int *__restrict__ counter, double *gws,
            size_t gws_stride)
{
    double *ws = (double *)((char *)gws + (size_t)blockIdx.x * gws_stride);
    while (true) {
        int i = 0;
        if (threadIdx.x == 0) i = atomicAdd(counter, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= count) break;
        gls_node_dense(a, list[i], ws);
        __syncwarp();
    }
}

int npb_gls_dense(npb_ctx *c, const GlsArgs &a, const int32_t *list, int count)
{
    if (count <= 0) return NPB_OK;
    int E = c->mx_epp, F = c->mx_fpp;
    size_t stride = (gls_dense_ws_bytes(E, E + 4 * F) + 255) & ~(size_t)255;
    int grid = c->sm_count * 8;
    if (grid > count) grid = count;
    NPB_TRY(npb_ensure(&c->gls_ws, &c->gls_ws_cap, stride * (size_t)grid));
    int *counter = c->counters + 30;
    NPB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), c->stream));
    k_gls_dense<<<grid, 32, 0, c->stream>>>(a, list, count, counter, (double *)c->gls_ws, stride);
    NPB_LAUNCH(c);
    NPB_CUDA(cudaGetLastError());
    return NPB_OK;
}
