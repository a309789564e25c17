// k2_gls.cu — per-node GLS weight construction (kernel group K2, tolerance method).
//
// Replaces GLSInterpolation.GLS and its helpers (ninpol/_methods/gls.pyx:75-474): build_ks_sv_arrays
// (:234-249), build_ls_matrices (:252-356), set_neumann_rows (:374-416), solve_ls (:420-474, LAPACK
// DGELS on an (E+3F+B) x (3E+1) system with E right-hand sides, keeping only the last row of X).
//
// Algebra: with M = [A | c] (A = the 3E gradient columns, c = last column = 1 on the E
// element rows, 0 elsewhere) and r = c - A argmin_g |A g - c| the residual of ONE least-squares problem,
//     weights_i = X[3E, i] = r_i / |r|^2 = r_i / sum_j r_j          (i < E),
// so one QR of A with the single right-hand side c replaces the reference's E solves.  The all-zero
// rows the reference leaves for boundary faces (gls.pyx:340-344) do not change a least-squares
// solution and are not materialised.  Reproduced reference behaviour: Q3 (neumann[p] is the weight of
// the node's LAST element; the Neumann right-hand side is dead), Q4 (that value is added to every
// weight of the row), Q8 (all faces boundary -> zero row).
//
// Structure exploited: A is block sparse — 3 columns per element; an element row touches one block, the
// 3 rows of an interior face touch the two blocks of its elements, a Neumann row touches one block.
// The kernel runs a MULTIFRONTAL Householder QR per node: element blocks are eliminated in greedy
// minimum-degree order (adjacency kept as per-lane bitmasks); eliminating block i gathers the row
// groups that contain i into a small dense front (rows x (3*|blocks|+1)), applies 3 Householder
// reflections, keeps the 3 pivot rows as rows of R and leaves the remaining rows as one new group.
// An interior node of the 50M-tet mesh (E=24, F=36) costs ~0.11 MFLOP this way instead of 1.19 MFLOP
// for a dense one-RHS QR (1.94 MFLOP in the reference).  Back substitution through the stored R rows
// gives g, then r_i = 1 - d_i . g_i on the element rows.
//
// Mapping: one warp per node (one-warp CTAs, persistent, atomic work counter).  The dense front being
// factored lives in shared memory (lanes over front columns, loops over front rows; the 3-column panel is
// factored with lanes over rows in registers + shuffles); the row groups — original rows and the
// contribution blocks left by earlier fronts — and the finished rows of R live in an append-only,
// L2-resident global slab per CTA and are fetched into the front with cp.async (LDGSTS), so shared
// memory holds only the hot data and ~10-27 nodes are resident per SM instead of 6.  Fronts taller than
// the shared buffer are processed in row chunks (the 3 pivot rows of a chunk are carried into the next).
// Nodes are bucketed into 7 size classes by star size; stars that do not fit (E > 64, a capacity
// overflow detected at run time) go to the dense global-memory kernel of k2_gls_dense.cu.  The kernel is
// latency / issue bound on this bookkeeping, not FP64 or HBM bound (profiles/).
#include <stdlib.h>
#include "gls_common.cuh"

typedef unsigned long long u64;

#define MF_NCLASS 9          // 0 = skipped node, 1..7 = shared-memory classes, 8 = dense fallback
#define MF_SCAP 64           // max row groups merged into one front

struct MfClass {
    int fcap;   // front buffer in shared memory (doubles)
    int ecap;   // max elements around the node
    int fcap_f; // max faces around the node
    int acap;   // capacity (doubles) of the CTA's global row-group arena; the R slab has the same size
};
#define MF_CLASS_TABLE {{0, 0, 0, 0}, {384, 8, 14, 3072}, {640, 12, 22, 6144}, {768, 16, 30, 8192}, {1408, 24, 40, 12288}, \
                        {1536, 32, 56, 18432}, {2560, 48, 80, 32768}, {4096, 64, 112, 49152}, {0, 0, 0, 0}}
__constant__ MfClass c_mf[MF_NCLASS] = MF_CLASS_TABLE;
static const MfClass h_mf[MF_NCLASS] = MF_CLASS_TABLE;

__host__ __device__ __forceinline__ int mf_ngcap(const MfClass &k) { return ((2 * k.ecap + k.fcap_f + 7) / 8) * 8; }
__host__ __device__ __forceinline__ int mf_mcap(const MfClass &k) { int m = k.ecap + 4 * k.fcap_f; return m < 96 ? m : 96; }
__host__ __device__ __forceinline__ size_t mf_smem_bytes(const MfClass &k)
{
    size_t d = (size_t)k.fcap + 4 * (size_t)mf_mcap(k) + 6 * (size_t)k.ecap;  // front, vbuf[.][4], gvec, dvec
    size_t b = d * 8 + (size_t)mf_ngcap(k) * (8 + 4 + 2 + 1) + (size_t)k.ecap * (8 + 4 + 4);   // group table, R table
    b += (size_t)k.ecap * 4 + MF_SCAP * 4 + 64;                             // es, S list, colblk
    return (b + 15) & ~(size_t)15;
}

// per node: Dirichlet / Q8 nodes are finished here (zero row); the others get a size class
__global__ void k_gls_classify(GlsArgs a, i64 lo, i64 hi, uint8_t *__restrict__ cls, int force_dense)
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
    int k = MF_NCLASS - 1;
    for (int q = 1; q < MF_NCLASS - 1 && !force_dense; q++)
        if (E <= c_mf[q].ecap && F <= c_mf[q].fcap_f) {
            k = q;
            break;
        }
    cls[p] = (uint8_t)k;
}

__device__ __forceinline__ u64 warp_or64(u64 v)
{
    unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
    unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
    return ((u64)hi << 32) | lo;
}
// 8-byte asynchronous global -> shared copy (LDGSTS): no register staging, no stall until the wait
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ int nth_set_bit(u64 m, int n)  // index of the n-th (0-based) set bit
{
    for (int q = 0; q < n; q++) m &= m - 1;
    return __ffsll((long long)m) - 1;
}

struct MfWs {
    double *arena;   // global: row groups (original rows, contribution blocks), append-only
    double *front;   // shared: the dense front being factored
    double *vbuf, *gvec, *dvec;
    u64 *g_mask, *r_mask;
    int *g_off, *r_off, *r_meta;   // r_meta = piv | npiv << 8 | c << 16
    unsigned short *g_nr;
    unsigned char *g_ld;
    int *es;
    int *s_list;            // packed (gid | rowbase << 16)
    unsigned char *colblk;  // block id of every 3-column slot of the current front
};

__device__ __forceinline__ MfWs mf_carve(unsigned char *base, const MfClass &k, double *garena)
{
    MfWs w;
    int ng = mf_ngcap(k);
    w.arena = garena;
    w.front = (double *)base;
    w.vbuf = w.front + k.fcap;                  // [mcap][4]: the three Householder vectors of a front
    w.gvec = w.vbuf + 4 * mf_mcap(k);
    w.dvec = w.gvec + 3 * k.ecap;
    w.g_mask = (u64 *)(w.dvec + 3 * k.ecap);
    w.r_mask = w.g_mask + ng;
    w.g_off = (int *)(w.r_mask + k.ecap);
    w.r_off = w.g_off + ng;
    w.r_meta = w.r_off + k.ecap;
    w.es = w.r_meta + k.ecap;
    w.s_list = w.es + k.ecap;
    w.g_nr = (unsigned short *)(w.s_list + MF_SCAP);
    w.g_ld = (unsigned char *)(w.g_nr + ng);
    w.colblk = w.g_ld + ng;                     // [64]
    return w;
}

// Closes the holes left by consumed groups in the (global) group arena: live table entries and their
// blocks slide down in table order, which equals arena order because entries are only ever appended.
// Keeps the slab region a CTA touches small enough that the resident CTAs' slabs stay in L2 (without it
// every contribution block ends up written to HBM once: 584 GB per pass over the 50M-tet mesh).
__device__ void mf_compact(MfWs &w, int &ng, int &top, int lane)
{
    int newng = 0, newtop = 0;
    for (int g0 = 0; g0 < ng; g0 += 32) {
        int g = g0 + lane;
        unsigned bal = __ballot_sync(0xffffffffu, g < ng && w.g_nr[g] > 0);
        while (bal) {
            int gi = g0 + __ffs(bal) - 1;
            bal &= bal - 1;
            u64 mk = w.g_mask[gi];
            int src = w.g_off[gi], nr = w.g_nr[gi], ld = w.g_ld[gi];
            int sz = nr * ld;
            if (src != newtop) {
                for (int i0 = 0; i0 < sz; i0 += 32) {
                    int i = i0 + lane;
                    double v = 0.0;
                    if (i < sz) v = w.arena[src + i];
                    __syncwarp();
                    if (i < sz) w.arena[newtop + i] = v;
                    __syncwarp();
                }
            }
            if (lane == 0) {
                w.g_mask[newng] = mk;
                w.g_off[newng] = newtop;
                w.g_nr[newng] = (unsigned short)nr;
                w.g_ld[newng] = (unsigned char)ld;
            }
            newtop += sz;
            newng++;
        }
    }
    __syncwarp();
    ng = newng;
    top = newtop;
}

#define MF_RPL 3   // front rows per lane in the panel factorisation: fronts of up to 96 rows

// Householder scalars of one column: alpha = -sign(x0) |x|, beta = 2 / |v|^2 with v = x - alpha e1, and
// rinv = 1 / alpha (kept on the diagonal of R for the back substitution).  sqrt and 1/alpha come from one
// rsqrt plus a Newton step instead of the IEEE sqrt and divide sequences.
__device__ __forceinline__ void hh_scalars(double sigma, double x0, double &alpha, double &beta, double &rinv)
{
    if (sigma == 0.0) {   // zero column: identity reflector, zero pivot (NaN takes the normal path and propagates)
        alpha = 0.0;
        beta = 0.0;
        rinv = 0.0;
        return;
    }
    double rs = rsqrt(sigma);
    double sq = sigma * rs;
    sq = fma(fma(-sq, sq, sigma), 0.5 * rs, sq);   // sq -> sqrt(sigma) to the last bit or so
    rs = fma(fma(-sq, rs, 1.0), rs, rs);           // rs -> 1 / sq
    alpha = (x0 >= 0.0) ? -sq : sq;
    rinv = (x0 >= 0.0) ? -rs : rs;
    beta = __drcp_rn(sigma - x0 * alpha);
}

// returns 0 on success, 1 when the star does not fit this class (caller reroutes it to the dense kernel)
__device__ int mf_node(const GlsArgs &a, int p, unsigned char *smem, const MfClass &kc, double *rslab, double *garena)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    MfWs w = mf_carve(smem, kc, garena);
    const int eb = a.esup_ptr[p], E = a.esup_ptr[p + 1] - eb;
    const int fb = a.fsup_ptr[p], F = a.fsup_ptr[p + 1] - fb;
    const bool neu = a.nflag[p] != 0;
    const double xv0 = a.coords[(i64)p * 3 + 0], xv1 = a.coords[(i64)p * 3 + 1], xv2 = a.coords[(i64)p * 3 + 2];
    const int ngcap = mf_ngcap(kc);

    // ---- setup: esup row, element groups ----
    for (int i = lane; i < E; i += 32) w.es[i] = a.esup[eb + i];
    __syncwarp();
    for (int i = lane; i < E; i += 32) {
        const double *cc = a.cent + (i64)w.es[i] * 3;
        double d0 = cc[0] - xv0, d1 = cc[1] - xv1, d2 = cc[2] - xv2;
        w.dvec[3 * i] = d0; w.dvec[3 * i + 1] = d1; w.dvec[3 * i + 2] = d2;
        double *row = w.arena + 4 * i;                    // [ d^T | 1 ]   gls.pyx:268-281
        row[0] = d0; row[1] = d1; row[2] = d2; row[3] = 1.0;
        w.g_mask[i] = 1ull << i;
        w.g_off[i] = 4 * i;
        w.g_nr[i] = 1;
        w.g_ld[i] = 4;
    }
    // ---- face groups (gls.pyx:291-356) and Neumann groups (:394-416) ----
    int n_if = 0;
    for (int f0 = 0; f0 < F; f0 += 32) {
        int fi = f0 + lane;
        bool interior = false;
        if (fi < F) interior = a.esuf2[a.fsup[fb + fi]].y >= 0;
        n_if += __popc(__ballot_sync(FULL, interior));
    }
    const int n_bf = F - n_if;
    const int off_if = 4 * E, off_bf = 4 * E + 21 * n_if;
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
        unsigned mi = __ballot_sync(FULL, interior);
        unsigned mb = __ballot_sync(FULL, boundary);
        unsigned below = (1u << lane) - 1u;
        if (interior) {
            int j = if_seen + __popc(mi & below);
            int I1 = 0, I2 = 0;
            for (int k = 0; k < E; k++) {
                int ek = w.es[k];
                if (ek == e2.x) I1 = k;
                if (ek == e2.y) I2 = k;
            }
            const double *Nn = a.fnormal + (i64)face * 3;
            const double *xs = a.fcent + (i64)face * 3;
            double N0 = Nn[0], N1 = Nn[1], N2 = Nn[2];
            double t0 = xv0 - xs[0], t1 = xv1 - xs[1], t2 = xv2 - xs[2];                     // T1 = x_v - x_S
            double c0 = N1 * t2 - N2 * t1, c1 = N2 * t0 - N0 * t2, c2 = N0 * t1 - N1 * t0;    // T2 = N x T1
            double eta = fmax(fmax(0.0, a.diff_mag[e2.x]), a.diff_mag[e2.y]);
            double tau = pow(sqrt(c0 * c0 + c1 * c1 + c2 * c2), -eta);
            const double *K1 = a.perm + (i64)e2.x * 9;
            const double *K2 = a.perm + (i64)e2.y * 9;
            // columns in ascending block order: the owner (smaller element id) has the smaller local index
            double s1 = -1.0, s2 = 1.0;
            int lo_i = I1, hi_i = I2;
            if (I2 < I1) { lo_i = I2; hi_i = I1; s1 = 1.0; s2 = -1.0; const double *t = K1; K1 = K2; K2 = t; }
            double *r1 = w.arena + off_if + 21 * j;   // 3 rows x (3 + 3 + rhs)
            double *r2 = r1 + 7, *r3 = r2 + 7;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                r1[q] = s1 * (K1[3 * q] * N0 + K1[3 * q + 1] * N1 + K1[3 * q + 2] * N2);
                r1[3 + q] = s2 * (K2[3 * q] * N0 + K2[3 * q + 1] * N1 + K2[3 * q + 2] * N2);
            }
            r2[0] = s1 * t0; r2[1] = s1 * t1; r2[2] = s1 * t2; r2[3] = s2 * t0; r2[4] = s2 * t1; r2[5] = s2 * t2;
            r3[0] = s1 * (tau * c0); r3[1] = s1 * (tau * c1); r3[2] = s1 * (tau * c2);
            r3[3] = s2 * (tau * c0); r3[4] = s2 * (tau * c1); r3[5] = s2 * (tau * c2);
            r1[6] = 0.0; r2[6] = 0.0; r3[6] = 0.0;
            int g = E + j;
            w.g_mask[g] = (1ull << lo_i) | (1ull << hi_i);
            w.g_off[g] = off_if + 21 * j;
            w.g_nr[g] = 3;
            w.g_ld[g] = 7;
        }
        if (boundary && neu) {
            int j = bf_seen + __popc(mb & below);
            int Ik = 0;
            for (int k = 0; k < E; k++)
                if (w.es[k] == e2.x) Ik = k;
            const double *Nn = a.fnormal + (i64)face * 3;
            const double *K1 = a.perm + (i64)e2.x * 9;
            double N0 = Nn[0], N1 = Nn[1], N2 = Nn[2];
            double *rr = w.arena + off_bf + 4 * j;
#pragma unroll
            for (int q = 0; q < 3; q++) rr[q] = -(K1[3 * q] * N0 + K1[3 * q + 1] * N1 + K1[3 * q + 2] * N2);
            rr[3] = 0.0;
            int g = E + n_if + j;
            w.g_mask[g] = 1ull << Ik;
            w.g_off[g] = off_bf + 4 * j;
            w.g_nr[g] = 1;
            w.g_ld[g] = 4;
        }
        if_seen += __popc(mi);
        bf_seen += __popc(mb);
    }
    int ng = E + n_if + (neu ? n_bf : 0);
    int top = off_bf + (neu ? 4 * n_bf : 0);
    __syncwarp();

    // ---- adjacency bitmasks: lane b owns blocks b and b + 32 ----
    u64 adjA = 0, adjB = 0;
    {
        const u64 bitA = 1ull << lane, bitB = 1ull << (lane + 32);
        for (int g = 0; g < ng; g++) {
            u64 mk = w.g_mask[g];
            if (mk & bitA) adjA |= mk;
            if (mk & bitB) adjB |= mk;
        }
    }
    u64 alive = (E >= 64) ? ~0ull : ((1ull << E) - 1ull);
    int nR = 0, rtop = 0;   // R rows written so far (entries / doubles in the global slab)

    // ---- elimination ----
    while (alive) {
        // (a) minimum-degree pivot block
        unsigned keyA = ((alive >> lane) & 1ull) ? (unsigned)((__popcll(adjA) << 8) | lane) : 0xffffffffu;
        unsigned keyB = ((alive >> (lane + 32)) & 1ull) ? (unsigned)((__popcll(adjB) << 8) | (lane + 32)) : 0xffffffffu;
        unsigned key = __reduce_min_sync(FULL, keyA < keyB ? keyA : keyB);
        const int piv = (int)(key & 0xffu);
        const u64 pbit = 1ull << piv;
        // rescue for stars whose contribution blocks outgrow the slab: close the holes left by consumed groups
        // (regular stars never get here; compacting eagerly would cut HBM write-back traffic but costs ~9 % time)
        if (top > (kc.acap / 4) * 3) {
            mf_compact(w, ng, top, lane);
        }
        // (b) row groups containing the pivot block
        int nS = 0, rho = 0;
        u64 U = 0;
        for (int g0 = 0; g0 < ng; g0 += 32) {
            int g = g0 + lane;
            bool in = false;
            int nr = 0;
            u64 mk = 0;
            if (g < ng) {
                nr = w.g_nr[g];
                mk = w.g_mask[g];
                in = nr > 0 && (mk & pbit);
            }
            unsigned bal = __ballot_sync(FULL, in);
            if (bal == 0) continue;
            int v = in ? nr : 0;   // exclusive prefix of the row counts over the selected lanes
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            int pos = nS + __popc(bal & ((1u << lane) - 1u));
            if (in && pos < MF_SCAP) w.s_list[pos] = g | ((rho + incl - v) << 16);
            nS += __popc(bal);
            rho += __shfl_sync(FULL, incl, 31);
            U |= warp_or64(in ? mk : 0ull);
        }
        const u64 Up = U & ~pbit;
        const int c = 3 * __popcll(U) + 1;
        // (c) capacity checks: group table, R slab; the front is processed in chunks of at most `rcap` rows
        const int rcap = min(32 * MF_RPL, min(mf_mcap(kc), kc.fcap / c));
        if (nS > MF_SCAP || rcap < 8 || rtop + 3 * c > kc.acap) return 1;
        // block id of every column slot: slot 0 = pivot, then the other blocks ascending
        if ((Up >> lane) & 1ull) w.colblk[1 + __popcll(Up & ((1ull << lane) - 1ull))] = (unsigned char)lane;
        if ((Up >> (lane + 32)) & 1ull) w.colblk[1 + __popcll(Up & ((1ull << (lane + 32)) - 1ull))] = (unsigned char)(lane + 32);
        if (lane == 0) w.colblk[0] = (unsigned char)piv;
        __syncwarp();
        double *Fm = w.front;
        const unsigned front_s = (unsigned)__cvta_generic_to_shared(Fm);
        const int ng0 = ng;         // groups appended below are not candidates of this elimination
        int t_next = 0, r_done = 0; // next group of the S list / rows of it already taken
        int carry = 0;              // pivot rows of the previous chunk, kept in front rows [0, carry)
        int npiv = 0;
        double fri0 = 0.0, fri1 = 0.0, fri2 = 0.0;   // 1 / alpha of the last chunk's pivots
        (void)rho;
        for (bool last = false; !last;) {
            // (d) assemble one chunk: lanes over front columns [pivot block | other blocks ascending | rhs]
            int rho_c = carry;
            while (t_next < nS && rho_c < rcap) {
                const int g = w.s_list[t_next] & 0xffff;
                const u64 mk = w.g_mask[g];
                const int nr = w.g_nr[g], ld = w.g_ld[g];
                const int take = min(nr - r_done, rcap - rho_c);
                const double *gsrc = w.arena + w.g_off[g] + r_done * ld;
                for (int j0 = 0; j0 < c; j0 += 32) {
                    int j = j0 + lane;
                    if (j >= c) continue;
                    const bool is_rhs = (j == c - 1);
                    const int blk = is_rhs ? 0 : w.colblk[j / 3];
                    const bool has = is_rhs || ((mk >> blk) & 1ull);
                    const int sc = is_rhs ? 3 * __popcll(mk) : 3 * __popcll(mk & ((1ull << blk) - 1ull)) + j % 3;
                    const double *src = gsrc + sc;
                    double *dst = Fm + rho_c * c + j;
                    if (has) {
                        unsigned ds = front_s + (unsigned)((rho_c * c + j) * 8);
                        for (int r = 0; r < take; r++) {
                            // row groups live in global memory: all copies of a chunk are in flight at once
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ds), "l"(src) : "memory");
                            src += ld;
                            ds += (unsigned)(c * 8);
                        }
                    } else {
                        for (int r = 0; r < take; r++) {
                            *dst = 0.0;
                            dst += c;
                        }
                    }
                }
                rho_c += take;
                r_done += take;
                if (r_done == nr) {
                    t_next++;
                    r_done = 0;
                }
            }
            last = (t_next >= nS);
            cp_async_wait_all();
            __syncwarp();
            // (e) panel: Householder on the three pivot columns, lanes over rows, entries in registers
            double a0[MF_RPL], a1[MF_RPL], a2[MF_RPL];
#pragma unroll
            for (int q = 0; q < MF_RPL; q++) {
                int r = lane + 32 * q;
                bool ok = r < rho_c;
                a0[q] = ok ? Fm[r * c + 0] : 0.0;
                a1[q] = ok ? Fm[r * c + 1] : 0.0;
                a2[q] = ok ? Fm[r * c + 2] : 0.0;
            }
            double alpha0, beta0, alpha1, beta1, alpha2, beta2, d10, d20, d21, ri0, ri1, ri2;
            {
                double sg = 0.0;
#pragma unroll
                for (int q = 0; q < MF_RPL; q++) sg += a0[q] * a0[q];
                sg = warp_sum(sg);
                double x00 = __shfl_sync(FULL, a0[0], 0);
                hh_scalars(sg, x00, alpha0, beta0, ri0);
                if (lane == 0) a0[0] = x00 - alpha0;                  // a0 now holds v0
                double t1 = 0.0, t2 = 0.0;
#pragma unroll
                for (int q = 0; q < MF_RPL; q++) { t1 += a0[q] * a1[q]; t2 += a0[q] * a2[q]; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { t1 += __shfl_xor_sync(FULL, t1, o); t2 += __shfl_xor_sync(FULL, t2, o); }
                t1 *= beta0; t2 *= beta0;
#pragma unroll
                for (int q = 0; q < MF_RPL; q++) { a1[q] -= t1 * a0[q]; a2[q] -= t2 * a0[q]; }
            }
            double r01 = __shfl_sync(FULL, a1[0], 0), r02 = __shfl_sync(FULL, a2[0], 0);
            {
                if (lane == 0) a1[0] = 0.0;                           // rows above the pivot do not take part
                double sg = 0.0;
#pragma unroll
                for (int q = 0; q < MF_RPL; q++) sg += a1[q] * a1[q];
                sg = warp_sum(sg);
                double x11 = __shfl_sync(FULL, a1[0], 1);
                hh_scalars(sg, x11, alpha1, beta1, ri1);
                if (lane == 1) a1[0] = x11 - alpha1;                  // a1 now holds v1
                if (lane == 0) a2[0] = 0.0;
                double t2 = 0.0;
#pragma unroll
                for (int q = 0; q < MF_RPL; q++) t2 += a1[q] * a2[q];
                t2 = warp_sum(t2) * beta1;
#pragma unroll
                for (int q = 0; q < MF_RPL; q++) a2[q] -= t2 * a1[q];
            }
            double r12 = __shfl_sync(FULL, a2[0], 1);
            {
                if (lane == 1) a2[0] = 0.0;
                double sg = 0.0;
#pragma unroll
                for (int q = 0; q < MF_RPL; q++) sg += a2[q] * a2[q];
                sg = warp_sum(sg);
                double x22 = __shfl_sync(FULL, a2[0], 2);
                hh_scalars(sg, x22, alpha2, beta2, ri2);
                if (lane == 2) a2[0] = x22 - alpha2;                  // a2 now holds v2
                d10 = 0.0; d20 = 0.0; d21 = 0.0;
#pragma unroll
                for (int q = 0; q < MF_RPL; q++) { d10 += a1[q] * a0[q]; d20 += a2[q] * a0[q]; d21 += a2[q] * a1[q]; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    d10 += __shfl_xor_sync(FULL, d10, o);
                    d20 += __shfl_xor_sync(FULL, d20, o);
                    d21 += __shfl_xor_sync(FULL, d21, o);
                }
            }
            // v vectors to shared memory as [row][4]; the pivot rows' panel entries of R
#pragma unroll
            for (int q = 0; q < MF_RPL; q++) {
                int r = lane + 32 * q;
                if (r < rho_c) {
                    double2 *vp = reinterpret_cast<double2 *>(w.vbuf + 4 * r);
                    vp[0] = make_double2(a0[q], a1[q]);
                    vp[1] = make_double2(a2[q], 0.0);
                }
            }
            if (lane == 0) {
                Fm[0] = alpha0; Fm[1] = r01; Fm[2] = r02;
                if (rho_c > 1) { Fm[c] = 0.0; Fm[c + 1] = alpha1; Fm[c + 2] = r12; }
                if (rho_c > 2) { Fm[2 * c] = 0.0; Fm[2 * c + 1] = 0.0; Fm[2 * c + 2] = alpha2; }
            }
            __syncwarp();
            // (f) apply the three reflections to the other columns: two passes over the rows
            for (int j0 = 3; j0 < c; j0 += 32) {
                int j = j0 + lane;
                if (j >= c) continue;
                double w0 = 0.0, w1 = 0.0, w2 = 0.0;
                {
                    const double *fp = Fm + j;
                    const double2 *vp = reinterpret_cast<const double2 *>(w.vbuf);
#pragma unroll 4
                    for (int r = 0; r < rho_c; r++) {
                        double f = *fp;
                        double2 va = vp[0], vb = vp[1];
                        w0 += va.x * f; w1 += va.y * f; w2 += vb.x * f;
                        fp += c; vp += 2;
                    }
                }
                double s0 = beta0 * w0;
                double s1 = beta1 * (w1 - d10 * s0);
                double s2 = beta2 * (w2 - d20 * s0 - d21 * s1);
                {
                    double *fp = Fm + j;
                    const double2 *vp = reinterpret_cast<const double2 *>(w.vbuf);
#pragma unroll 4
                    for (int r = 0; r < rho_c; r++) {
                        double2 va = vp[0], vb = vp[1];
                        *fp = *fp - (va.x * s0 + va.y * s1 + vb.x * s2);
                        fp += c; vp += 2;
                    }
                }
            }
            __syncwarp();
            fri0 = ri0; fri1 = ri1; fri2 = ri2;
            // (g) rows below the pivot rows (columns of the other blocks + rhs) are appended to the group arena
            //     as a new contribution block; the pivot rows stay in the front for the next chunk, or go to
            //     the R slab after the last one
            npiv = rho_c < 3 ? rho_c : 3;
            const int left = rho_c - npiv;
            const bool keep = left > 0 && Up != 0;
            if (keep) {
                const int cw = c - 3;
                if (ng + 1 > ngcap || top + left * cw > kc.acap) return 1;
                for (int j0 = 0; j0 < cw; j0 += 32) {
                    int j = j0 + lane;
                    if (j < cw) {
                        const double *src = Fm + npiv * c + 3 + j;
                        double *dst = w.arena + top + j;
                        for (int r = 0; r < left; r++) {
                            *dst = *src;
                            src += c;
                            dst += cw;
                        }
                    }
                }
                if (lane == 0) {
                    w.g_mask[ng] = Up;
                    w.g_off[ng] = top;
                    w.g_nr[ng] = (unsigned short)left;
                    w.g_ld[ng] = (unsigned char)cw;
                }
                top += left * cw;
                ng++;
            }
            carry = npiv;
            __syncwarp();
        }
        for (int t = lane; t < nS; t += 32) w.g_nr[w.s_list[t] & 0xffff] = 0;   // consumed
        (void)ng0;
        // R rows to the slab; the diagonal carries 1 / alpha so that the back substitution has no divisions
        for (int j = lane; j < npiv * c; j += 32) {
            double v = Fm[j];
            if (j == 0) v = fri0;
            if (j == c + 1) v = fri1;
            if (j == 2 * c + 2) v = fri2;
            rslab[rtop + j] = v;
        }
        if (lane == 0) {
            w.r_mask[nR] = U;
            w.r_off[nR] = rtop;
            w.r_meta[nR] = piv | (npiv << 8) | (c << 16);
        }
        nR++;
        rtop += npiv * c;
        // (h) adjacency update: the neighbours of the pivot become a clique
        if ((Up >> lane) & 1ull) adjA = (adjA | U) & ~pbit;
        if ((Up >> (lane + 32)) & 1ull) adjB = (adjB | U) & ~pbit;
        alive &= ~pbit;
        __syncwarp();
    }

    // ---- back substitution through the R rows, newest first (slab copied back into the free arena) ----
    __syncwarp();
    const bool r_in_smem = rtop <= kc.fcap;
    if (r_in_smem) {
        for (int i = lane; i < rtop; i += 32) cp_async8(w.front + i, rslab + i);
        cp_async_wait_all();
    }
    const double *Rbase = r_in_smem ? w.front : rslab;
    for (int i = lane; i < 3 * E; i += 32) w.gvec[i] = 0.0;
    __syncwarp();
    for (int g = nR - 1; g >= 0; g--) {
        const int meta = w.r_meta[g];
        const int piv = meta & 0xff, npiv = (meta >> 8) & 0xff, c = meta >> 16;
        const u64 Up = w.r_mask[g] & ~(1ull << piv);
        const double *R = Rbase + w.r_off[g];
        double p0 = 0.0, p1 = 0.0, p2 = 0.0;
        for (int j0 = 3; j0 < c - 1; j0 += 32) {
            int j = j0 + lane;
            if (j < c - 1) {
                int blk = nth_set_bit(Up, (j - 3) / 3);
                double gj = w.gvec[3 * blk + (j - 3) % 3];
                p0 += R[j] * gj;
                if (npiv > 1) p1 += R[c + j] * gj;
                if (npiv > 2) p2 += R[2 * c + j] * gj;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            p0 += __shfl_xor_sync(FULL, p0, o);
            p1 += __shfl_xor_sync(FULL, p1, o);
            p2 += __shfl_xor_sync(FULL, p2, o);
        }
        if (lane == 0) {
            double g0 = 0.0, g1 = 0.0, g2 = 0.0;   // the diagonal of R holds 1 / alpha (0 for a zero pivot)
            if (npiv > 2) g2 = (R[2 * c + c - 1] - p2) * R[2 * c + 2];
            if (npiv > 1) g1 = (R[c + c - 1] - p1 - R[c + 2] * g2) * R[c + 1];
            g0 = (R[c - 1] - p0 - R[1] * g1 - R[2] * g2) * R[0];
            w.gvec[3 * piv] = g0;
            w.gvec[3 * piv + 1] = g1;
            w.gvec[3 * piv + 2] = g2;
        }
        __syncwarp();
    }
    // ---- residual on the element rows, weights, CSR values ----
    double *wo = a.wbuf + ((i64)eb - a.wbase);
    double part = 0.0;
    for (int i = lane; i < E; i += 32) {
        double ri = 1.0 - (w.dvec[3 * i] * w.gvec[3 * i] + w.dvec[3 * i + 1] * w.gvec[3 * i + 1] + w.dvec[3 * i + 2] * w.gvec[3 * i + 2]);
        w.vbuf[i] = ri;
        part += ri;
    }
    double tot = warp_sum(part);
    __syncwarp();
    double nv = neu ? w.vbuf[E - 1] / tot : 0.0;   // gls.pyx:470-472 (Q3)
    int cnt = 0;
    for (int i = lane; i < E; i += 32) {
        double v = w.vbuf[i] / tot + nv;           // interpolator.pyx:618 (Q4)
        wo[i] = v;
        cnt += (v != 0.0) ? 1 : 0;
    }
    cnt = __reduce_add_sync(FULL, cnt);
    if (lane == 0) {
        a.rowcnt[p] = cnt;
        a.neumann[p] = nv;
    }
    __syncwarp();
    return 0;
}

// persistent: one warp per CTA; nodes handed out through an atomic counter; stars that do not fit are
// appended to the overflow list for the dense kernel
template <int MINBLOCKS>
__global__ void __launch_bounds__(32, MINBLOCKS)
k_gls_mf(GlsArgs a, const int32_t *__restrict__ list, int count, int *__restrict__ counter, int klass,
         int32_t *__restrict__ overflow, int *__restrict__ n_overflow, double *__restrict__ slabs)
{
    extern __shared__ __align__(16) unsigned char smem_mf[];
    const MfClass kc = c_mf[klass];
    double *rslab = slabs + (size_t)blockIdx.x * kc.acap * 2;   // per CTA: R slab, then the group arena
    double *garena = rslab + kc.acap;
    while (true) {
        int i = 0;
        if (threadIdx.x == 0) i = atomicAdd(counter, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= count) break;
        int p = list[i];
        int rc = mf_node(a, p, smem_mf, kc, rslab, garena);
        __syncwarp();
        if (rc != 0 && threadIdx.x == 0) overflow[atomicAdd(n_overflow, 1)] = p;
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
    if (!c->node_list) {
        // [0, n): work list of the current class; [n, 2n): overflow list; then n bytes of classes
        NPB_TRY(npb_alloc(c, (void **)&c->node_list, sizeof(int32_t) * 2 * (size_t)c->n_points + (size_t)c->n_points));
    }
    int32_t *list = c->node_list, *overflow = c->node_list + c->n_points;
    uint8_t *cls = (uint8_t *)(c->node_list + 2 * c->n_points);
    int *n_overflow = c->counters + 40;
    NPB_CUDA(cudaMemsetAsync(n_overflow, 0, sizeof(int), s));
    const char *force = getenv("NPB_FORCE_GLS_DENSE");   // tests: exercise the dense fallback kernel
    k_gls_classify<<<npb_blocks(nloc, 256), 256, 0, s>>>(a, lo, hi, cls, (force && force[0] == '1') ? 1 : 0);
    NPB_LAUNCH(c);
    float main_ms = 0.f;
    static const char *cls_names[MF_NCLASS] = {"", "k2_gls_c1", "k2_gls_c2", "k2_gls_c3", "k2_gls_c4", "k2_gls_c5", "k2_gls_c6", "k2_gls_c7", "k2_gls_dense"};
    for (int k = 1; k < MF_NCLASS; k++) c->timings.erase(cls_names[k]);
    int n_dense_direct = 0;
    for (int k = 1; k < MF_NCLASS - 1; k++) {
        int count = 0;
        NPB_TRY(npb_select_class(c, cls, lo, hi, k, list, &count));
        if (count == 0) continue;
        int *counter = c->counters + 20 + k;
        NPB_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), s));
        NpbTimer tk(c, cls_names[k]);
        int smem = (int)mf_smem_bytes(h_mf[k]);
        const bool small = smem <= 9 * 1024;   // small stars: trade registers for resident warps
        if (small)
            NPB_CUDA(cudaFuncSetAttribute(k_gls_mf<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        else
            NPB_CUDA(cudaFuncSetAttribute(k_gls_mf<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        const char *cap = getenv("NPB_GLS_CTAS_PER_SM");
        if (cap && atoi(cap) > 0 && per_sm > atoi(cap)) per_sm = atoi(cap);
        if (per_sm > 32) per_sm = 32;
        int grid = c->sm_count * per_sm;
        if (grid > count) grid = count;
        NPB_TRY(npb_ensure(&c->gls_ws, &c->gls_ws_cap, sizeof(double) * (size_t)grid * h_mf[k].acap * 2));
        if (small)
            k_gls_mf<24><<<grid, 32, smem, s>>>(a, list, count, counter, k, overflow, n_overflow, (double *)c->gls_ws);
        else
            k_gls_mf<12><<<grid, 32, smem, s>>>(a, list, count, counter, k, overflow, n_overflow, (double *)c->gls_ws);
        NPB_LAUNCH(c);
        NPB_CUDA(cudaGetLastError());
        tk.stop();
        if (c->timings[cls_names[k]] > main_ms) main_ms = c->timings[cls_names[k]];
    }
    // dense fallback: stars classified as too large, then whatever overflowed at run time
    {
        NpbTimer tk(c, cls_names[MF_NCLASS - 1]);
        NPB_TRY(npb_select_class(c, cls, lo, hi, MF_NCLASS - 1, list, &n_dense_direct));
        NPB_TRY(npb_gls_dense(c, a, list, n_dense_direct));
        int h_over = 0;
        NPB_CUDA(cudaMemcpyAsync(&h_over, n_overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
        NPB_CUDA(cudaStreamSynchronize(s));
        NPB_TRY(npb_gls_dense(c, a, overflow, h_over));
        tk.stop();
        c->timings["gls_dense_nodes"] = (float)(n_dense_direct + h_over);
        if (c->timings[cls_names[MF_NCLASS - 1]] > main_ms && (n_dense_direct + h_over) > 0) main_ms = c->timings[cls_names[MF_NCLASS - 1]];
    }
    c->timings["k2_main"] = main_ms;
    return NPB_OK;
}
