// k2_idw_ls.cu — per-node IDW and LS weight construction (kernel group K2, bit-exact methods).
//
// Replaces IDWInterpolation.inverse_distance (ninpol/_methods/idw.pyx:35-84) and LSInterpolation.LS
// (ninpol/_methods/ls.pyx:33-135).  Both use only IEEE +,-,*,/,sqrt in a fixed left-to-right order
// (SURVEY.md App. C), so this file is compiled with -fmad=false and spells every operation with the
// _rn intrinsics: results are bit-identical to the reference, including the NaN rows LS produces at
// Neumann nodes with coplanar centroids.
//
// Output: the FINAL CSR values `w + neumann_ws` (interpolator.pyx:618; neumann_ws is never written by
// these two methods, so it is +0.0) stored esup-indexed in wbuf (local node range), the per-row count of
// entries that survive scipy's eliminate_zeros (value != 0, NaN kept) and neumann[p] = 0.
// Mapping: memory-bound gather; one thread per node walks its esup row serially, which is what keeps
// the sequential sums in reference order.
#include "common.cuh"

#define IDW_EPS ((double)1.0000000036274937e-15f) /* float32(1e-15), idw.pyx:53 */

__global__ void __launch_bounds__(128)
k_idw(const int32_t *__restrict__ esup_ptr, const int32_t *__restrict__ esup, const uint8_t *__restrict__ bpoint,
      const uint8_t *__restrict__ nflag, const double *__restrict__ coords, const double *__restrict__ cent, int dim,
      i64 lo, i64 hi, i64 wbase, double *__restrict__ wbuf, int32_t *__restrict__ rowcnt, double *__restrict__ neumann)
{
    i64 p = lo + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hi) return;
    int b = esup_ptr[p], e = esup_ptr[p + 1];
    double *w = wbuf + ((i64)b - wbase);
    neumann[p] = 0.0;
    if (bpoint[p] && !nflag[p]) {  // Dirichlet node: row stays zero (idw.pyx:62-63)
        for (int q = b; q < e; q++) w[q - b] = 0.0;
        rowcnt[p] = 0;
        return;
    }
    double x0 = coords[p * 3 + 0], x1 = coords[p * 3 + 1], x2 = coords[p * 3 + 2];
    double total = 0.0;
    int n_source = 0;
    int zero_at = -1;
    for (int q = b; q < e; q++) {
        const double *cc = cent + (i64)esup[q] * NPB_CSTRIDE;
        double d0 = __dsub_rn(x0, cc[0]);
        double dist = __dadd_rn(0.0, __dmul_rn(d0, d0));
        if (dim > 1) {
            double d1 = __dsub_rn(x1, cc[1]);
            dist = __dadd_rn(dist, __dmul_rn(d1, d1));
        }
        if (dim > 2) {
            double d2 = __dsub_rn(x2, cc[2]);
            dist = __dadd_rn(dist, __dmul_rn(d2, d2));
        }
        if (dist <= IDW_EPS) {  // coincident centroid: one-hot row (idw.pyx:69-74)
            zero_at = q - b;
            break;
        }
        double r = __ddiv_rn(1.0, __dsqrt_rn(dist));
        w[q - b] = __dadd_rn(0.0, r);
        total = __dadd_rn(total, r);
        n_source++;
    }
    int cnt = 0;
    if (zero_at >= 0) {
        for (int q = b; q < e; q++) w[q - b] = (q - b == zero_at) ? 1.0 : 0.0;
        cnt = 1;
    } else {
        for (int k = 0; k < n_source; k++) {
            double v = __dadd_rn(__ddiv_rn(w[k], total), 0.0);  // + neumann_ws (interpolator.pyx:618)
            w[k] = v;
            cnt += (v != 0.0) ? 1 : 0;
        }
    }
    rowcnt[p] = cnt;
}

__global__ void __launch_bounds__(128)
k_ls(const int32_t *__restrict__ esup_ptr, const int32_t *__restrict__ esup, const uint8_t *__restrict__ bpoint,
     const uint8_t *__restrict__ nflag, const double *__restrict__ coords, const double *__restrict__ cent, i64 lo, i64 hi,
     i64 wbase, double *__restrict__ wbuf, int32_t *__restrict__ rowcnt, double *__restrict__ neumann)
{
    i64 p = lo + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hi) return;
    int b = esup_ptr[p], e = esup_ptr[p + 1];
    double *w = wbuf + ((i64)b - wbase);
    neumann[p] = 0.0;
    if (bpoint[p] && !nflag[p]) {
        for (int q = b; q < e; q++) w[q - b] = 0.0;
        rowcnt[p] = 0;
        return;
    }
    double x0 = coords[p * 3 + 0], x1 = coords[p * 3 + 1], x2 = coords[p * 3 + 2];
    double Ix = 0.0, Iy = 0.0, Iz = 0.0, Ixx = 0.0, Ixy = 0.0, Ixz = 0.0, Iyy = 0.0, Iyz = 0.0, Izz = 0.0;
    for (int q = b; q < e; q++) {  // ls.pyx:64-77
        const double *cc = cent + (i64)esup[q] * NPB_CSTRIDE;
        double vx = __dsub_rn(cc[0], x0), vy = __dsub_rn(cc[1], x1), vz = __dsub_rn(cc[2], x2);
        Ix = __dadd_rn(Ix, vx);
        Iy = __dadd_rn(Iy, vy);
        Iz = __dadd_rn(Iz, vz);
        Ixx = __dadd_rn(Ixx, __dmul_rn(vx, vx));
        Ixy = __dadd_rn(Ixy, __dmul_rn(vx, vy));
        Ixz = __dadd_rn(Ixz, __dmul_rn(vx, vz));
        Iyy = __dadd_rn(Iyy, __dmul_rn(vy, vy));
        Iyz = __dadd_rn(Iyz, __dmul_rn(vy, vz));
        Izz = __dadd_rn(Izz, __dmul_rn(vz, vz));
    }
    bool flat = (Iz == 0.0 && Izz == 0.0 && Ixz == 0.0 && Iyz == 0.0);
    if (flat) Izz = 1.0;  // ls.pyx:79-80
#define M2(a, b) __dmul_rn(a, b)
#define S2(a, b) __dsub_rn(a, b)
#define A2(a, b) __dadd_rn(a, b)
    double D = A2(A2(M2(Ixx, S2(M2(Iyy, Izz), M2(Iyz, Iyz))), M2(Ixy, S2(M2(Iyz, Ixz), M2(Ixy, Izz)))),
                  M2(Ixz, S2(M2(Ixy, Iyz), M2(Iyy, Ixz))));
    int cnt = 0;
    if (D == 0.0) {  // inverse-distance fallback, ls.pyx:88-102
        double total = 0.0;
        for (int q = b; q < e; q++) {
            const double *cc = cent + (i64)esup[q] * NPB_CSTRIDE;
            double vx = S2(cc[0], x0), vy = S2(cc[1], x1), vz = S2(cc[2], x2);
            double r = __ddiv_rn(1.0, __dsqrt_rn(A2(A2(M2(vx, vx), M2(vy, vy)), M2(vz, vz))));
            w[q - b] = r;
            total = A2(total, r);
        }
        for (int q = b; q < e; q++) {
            double v = A2(__ddiv_rn(w[q - b], total), 0.0);
            w[q - b] = v;
            cnt += (v != 0.0) ? 1 : 0;
        }
        rowcnt[p] = cnt;
        return;
    }
    // ls.pyx:105-106 tests the same condition again, but Izz is 1.0 by now whenever it held, so the
    // reference never sets Izz = -1.0; restated literally
    if (Iz == 0.0 && Izz == 0.0 && Ixz == 0.0 && Iyz == 0.0) Izz = -1.0;
    double lx = __ddiv_rn(A2(A2(M2(Ix, S2(M2(Iyz, Iyz), M2(Iyy, Izz))), M2(Iy, S2(M2(Ixy, Izz), M2(Iyz, Ixz)))),
                             M2(Iz, S2(M2(Iyy, Ixz), M2(Ixy, Iyz)))), D);
    double ly = __ddiv_rn(A2(A2(M2(Ix, S2(M2(Ixy, Izz), M2(Iyz, Ixz))), M2(Iy, S2(M2(Ixz, Ixz), M2(Ixx, Izz)))),
                             M2(Iz, S2(M2(Ixx, Iyz), M2(Ixy, Ixz)))), D);
    double lz = __ddiv_rn(A2(A2(M2(Ix, S2(M2(Iyy, Ixz), M2(Ixy, Iyz))), M2(Iy, S2(M2(Ixx, Iyz), M2(Ixy, Ixz)))),
                             M2(Iz, S2(M2(Ixy, Ixy), M2(Ixx, Iyy)))), D);
    double denom = A2(A2(A2((double)(e - b), M2(lx, Ix)), M2(ly, Iy)), M2(lz, Iz));  // ls.pyx:126
    for (int q = b; q < e; q++) {
        const double *cc = cent + (i64)esup[q] * NPB_CSTRIDE;
        double vx = S2(cc[0], x0), vy = S2(cc[1], x1), vz = S2(cc[2], x2);
        double v = A2(A2(A2(1.0, M2(lx, vx)), M2(ly, vy)), M2(lz, vz));
        v = A2(__ddiv_rn(v, denom), 0.0);
        w[q - b] = v;
        cnt += (v != 0.0) ? 1 : 0;
    }
    rowcnt[p] = cnt;
#undef M2
#undef S2
#undef A2
}

int npb_k2_idw_ls(npb_ctx *c, int method, i64 lo, i64 hi)
{
    if (hi <= lo) return NPB_OK;
    const int T = 128;
    i64 wbase = c->wbase;
    if (method == NPB_METHOD_IDW)
        k_idw<<<npb_blocks(hi - lo, T), T, 0, c->stream>>>(c->esup_ptr, c->esup, c->bpoint, c->nflag, c->coords, c->centroids,
                                                          c->dim, lo, hi, wbase, c->wbuf, c->rowcnt, c->neumann);
    else
        k_ls<<<npb_blocks(hi - lo, T), T, 0, c->stream>>>(c->esup_ptr, c->esup, c->bpoint, c->nflag, c->coords, c->centroids,
                                                         lo, hi, wbase, c->wbuf, c->rowcnt, c->neumann);
    NPB_LAUNCH(c);
    NPB_CUDA(cudaGetLastError());
    return NPB_OK;
}
