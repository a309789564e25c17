// k1_connectivity.cu — load_mesh connectivity on the device (kernel group K1, integer part).
//
// Replaces the serial / OpenMP loops of the reference's Grid.build (ninpol/_interpolator/grid.pyx):
//   build_esup  :233-267   node -> element CSR      = histogram + exclusive scan + fill + per-row sort
//   build_esuel :449-525   element -> element       = one thread per (element, local face)
//   build_infael:304-345   global face numbering    = exclusive scan of ownership flags
//   build_fsup  :347-379   node -> face CSR         = histogram + scan + fill + per-row sort
//   build_esuf  :381-444   face -> elements + boundary face / node tags
// All outputs are bit-identical to the reference's (rows ascending, first-encounter face numbering);
// the equivalences are spelled out in SURVEY.md App. A and checked in tests/test_gpu_parity.py.
// Everything here is HBM-bound integer work: ids are int32 on the device, rows are compact.
#include <stdlib.h>
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// int64 [n_elems, cstride] (-1 padded) host layout  ->  int32 [n_elems, spe] + uint8 element types
// ------------------------------------------------------------------------------------------------
// Node ids are validated here, before anything scatters through them: the first npoel[type] ids of an element
// must lie in [0, n_points) (a 1-based or mis-indexed mesh would otherwise write out of bounds on the device;
// the reference segfaults on the host).  bad[0] = lowest offending element id + 1 (0: none).
__global__ void k_convert_conn(const i64 *__restrict__ conn, int cstride, const i64 *__restrict__ types, i64 n_elems, int spe,
                               ElemTables tab, i64 n_points, int32_t *__restrict__ inpoel, uint8_t *__restrict__ etype,
                               unsigned long long *__restrict__ bad)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_elems * spe) return;
    i64 e = idx / spe;
    int j = (int)(idx - e * spe);
    int t = (int)types[e];
    i64 v = j < cstride ? conn[e * cstride + j] : -1;
    if (j < tab.npoel[t]) {
        if (v < 0 || v >= n_points) {
            atomicMin(bad, (unsigned long long)e);
            v = 0;   // keep the table in range; the build is abandoned below
        }
    } else
        v = -1;      // slots past the element's node count are padding whatever the caller left there
    inpoel[idx] = (int32_t)v;
    if (j == 0) etype[e] = (uint8_t)t;
}

// histogram of node incidences over the padded (element, local node) table
__global__ void k_count_nodes(const int32_t *__restrict__ table, i64 n, int32_t *__restrict__ cnt)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int p = table[idx];
    if (p >= 0) atomicAdd(&cnt[p], 1);
}

// scatter owner ids into the rows (arbitrary order inside a row; k_sort_rows restores ascending order)
__global__ void k_fill_rows(const int32_t *__restrict__ table, i64 n, int stride, const int32_t *__restrict__ ptr,
                            int32_t *__restrict__ cursor, int32_t *__restrict__ out)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int p = table[idx];
    if (p < 0) return;
    int pos = atomicAdd(&cursor[p], 1);
    out[(i64)ptr[p] + pos] = (int32_t)(idx / stride);
}

// ascending insertion sort of every CSR row (rows are short: <= MX_*_PER_POINT) + row-length maximum
__global__ void k_sort_rows(const int32_t *__restrict__ ptr, i64 n_rows, int32_t *__restrict__ vals, int *__restrict__ mx)
{
    i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    if (r < n_rows) {
        i64 b = ptr[r];
        len = ptr[r + 1] - ptr[r];
        int32_t *row = vals + b;
        for (int i = 1; i < len; i++) {
            int32_t v = row[i];
            int j = i - 1;
            while (j >= 0 && row[j] > v) {
                row[j + 1] = row[j];
                j--;
            }
            row[j + 1] = v;
        }
    }
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(mx, len);
}

// The same sort with the rows of a CTA staged in shared memory: SORT_ROWS consecutive rows are one contiguous slice of
// `vals`, so it is read and written back fully coalesced (TMA-friendly layout) and the per-row insertion sort — rows
// arrive almost sorted, because element / face ids grow with the launch order of the atomic fill — runs on shared
// memory instead of on scattered global words.  Slices larger than the buffer fall back to the in-place sort.
#define SORT_ROWS 256
#define SORT_CAP 12032   // ints of shared memory (47 KB): 256 rows of up to 47 entries
__global__ void __launch_bounds__(SORT_ROWS)
k_sort_rows_smem(const int32_t *__restrict__ ptr, i64 n_rows, int32_t *__restrict__ vals, int *__restrict__ mx)
{
    __shared__ int32_t buf[SORT_CAP];
    const i64 r0 = (i64)blockIdx.x * SORT_ROWS;
    const i64 r1 = min(r0 + SORT_ROWS, n_rows);
    const i64 r = r0 + threadIdx.x;
    const int b0 = ptr[r0], b1 = ptr[r1];
    const int ne = b1 - b0;
    int len = 0;
    if (ne <= SORT_CAP) {
        for (int i = threadIdx.x; i < ne; i += SORT_ROWS) buf[i] = vals[(i64)b0 + i];
        __syncthreads();
        if (r < r1) {
            const int b = ptr[r] - b0;
            len = ptr[r + 1] - ptr[r];
            int32_t *row = buf + b;
            for (int i = 1; i < len; i++) {
                int32_t v = row[i];
                int j = i - 1;
                while (j >= 0 && row[j] > v) {
                    row[j + 1] = row[j];
                    j--;
                }
                row[j + 1] = v;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < ne; i += SORT_ROWS) vals[(i64)b0 + i] = buf[i];
    } else if (r < r1) {
        i64 b = ptr[r];
        len = ptr[r + 1] - ptr[r];
        int32_t *row = vals + b;
        for (int i = 1; i < len; i++) {
            int32_t v = row[i];
            int j = i - 1;
            while (j >= 0 && row[j] > v) {
                row[j + 1] = row[j];
                j--;
            }
            row[j + 1] = v;
        }
    }
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(mx, len);
}

// ------------------------------------------------------------------------------------------------
// esuel: one thread per (element, local face).  Candidates are the elements around the face's
// lowest-degree node (first minimum in local order, grid.pyx:479-488); a candidate matches when one of
// its faces contains every node of this face (grid.pyx:502-512); first match in (esup order, local
// face order) wins.  On conforming meshes the match is unique, so the reference's reverse write
// esuel[jelem,l] = ielem (:517) is what the thread of (jelem,l) finds by itself.
// ------------------------------------------------------------------------------------------------
// The candidate test is evaluated with bitmasks: `hit` has bit k set when local node k of the candidate
// belongs to this face; face l of the candidate matches iff popc(hit & nodes_of_face[l]) equals the
// number of nodes of this face — literally the count `is_equal` of grid.pyx:503-512.
struct FaceMasks {
    unsigned char m[NPB_N_TYPES][NPB_MX_FE];
};

template <int SPE>
__global__ void __launch_bounds__(256)
k_esuel(ElemTables tab, FaceMasks fm, const int32_t *__restrict__ inpoel, const uint8_t *__restrict__ etype,
        const int32_t *__restrict__ esup_ptr, const int32_t *__restrict__ esup, i64 n_elems, int sfe,
        int32_t *__restrict__ esuel)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_elems * sfe) return;
    i64 e = idx / sfe;
    int j = (int)(idx - e * sfe);
    int t = etype[e];
    if (j >= tab.nfael[t]) {
        esuel[idx] = -1;
        return;
    }
    const int nj = tab.lnofa[t][j];
    int mine[NPB_MX_PF];
#pragma unroll
    for (int o = 0; o < NPB_MX_PF; o++) mine[o] = (o < nj) ? inpoel[e * SPE + tab.lpofa[t][j][o]] : -2;
    int point = mine[0];
    int nmin = esup_ptr[point + 1] - esup_ptr[point];
#pragma unroll
    for (int k = 1; k < NPB_MX_PF; k++) {
        if (k < nj) {
            int ne = esup_ptr[mine[k] + 1] - esup_ptr[mine[k]];
            if (ne < nmin) {
                point = mine[k];
                nmin = ne;
            }
        }
    }
    int res = -1;
    const int qb = esup_ptr[point], qe = esup_ptr[point + 1];
    for (int q = qb; q < qe; q++) {
        int je = esup[q];
        if (je == (int)e) continue;
        int row[SPE];
        const int4 *rp = reinterpret_cast<const int4 *>(inpoel + (i64)je * SPE);
#pragma unroll
        for (int v = 0; v < SPE / 4; v++) {
            int4 x = rp[v];
            row[4 * v] = x.x; row[4 * v + 1] = x.y; row[4 * v + 2] = x.z; row[4 * v + 3] = x.w;
        }
        unsigned hit = 0;
#pragma unroll
        for (int k = 0; k < SPE; k++) {
            int qn = row[k];
            bool h = (qn == mine[0]) | (qn == mine[1]) | (qn == mine[2]) | (qn == mine[3]);
            hit |= h ? (1u << k) : 0u;
        }
        if (__popc(hit) < nj) continue;       // cannot match any face
        int jt = etype[je];
        int nf = tab.nfael[jt];
        bool match = false;
        for (int l = 0; l < nf; l++) match = match || (__popc(hit & fm.m[jt][l]) == nj);
        if (match) {
            res = je;
            break;
        }
    }
    esuel[idx] = res;
}

// ------------------------------------------------------------------------------------------------
// esuel by node stars: one warp per node p matches, inside the star of p held in shared memory, the two sides of
// every face whose SMALLEST node id is p.  Both elements of such a face contain p, so both are in the star, and every
// face has exactly one smallest node: each interior face is paired exactly once, with no candidate connectivity
// fetched from global memory more than once per star (the per-face kernel above reads ~12 candidate rows per face).
// A side whose face has no equal node set in the star is a boundary face and keeps the -1 esuel was filled with.
// Equality of node sets is the reference's criterion on conforming meshes (grid.pyx:502-512).
// ------------------------------------------------------------------------------------------------
#define STAR_CAPE 64          // elements per star held in shared memory; larger stars set *too_big
#define STAR_PAIRS 128        // sides of faces whose smallest node is the star's node (2 per interior face)
#define STAR_WARPS 4
template <int SPE>
__global__ void __launch_bounds__(32 * STAR_WARPS)
k_esuel_star(ElemTables tab, const int32_t *__restrict__ inpoel, const uint8_t *__restrict__ etype,
             const int32_t *__restrict__ esup_ptr, const int32_t *__restrict__ esup, i64 n_points, int sfe,
             int32_t *__restrict__ esuel, int *__restrict__ too_big)
{
    __shared__ int s_conn[STAR_WARPS][STAR_CAPE * SPE];
    __shared__ int s_es[STAR_WARPS][STAR_CAPE];
    __shared__ unsigned char s_type[STAR_WARPS][STAR_CAPE];
    __shared__ int s_key[STAR_WARPS][STAR_PAIRS][3];     // the face's other nodes, ascending, -1 padded
    __shared__ unsigned short s_ij[STAR_WARPS][STAR_PAIRS];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    // The three dependent loads of a star (esup_ptr -> esup ids -> connectivity rows) are software-pipelined over the
    // nodes a warp visits: the row pointer is fetched two nodes ahead, the element ids one node ahead (two per lane:
    // STAR_CAPE = 64), so only the connectivity rows are waited for.
    const i64 step = (i64)gridDim.x * STAR_WARPS;
    const i64 p_first = (i64)blockIdx.x * STAR_WARPS + wid;
    int ebC = 0, EC = 0, eC0 = -1, eC1 = -1, ebB = 0, EB = 0;
    if (p_first < n_points) {
        ebC = esup_ptr[p_first];
        EC = esup_ptr[p_first + 1] - ebC;
        if (EC <= STAR_CAPE) {
            eC0 = lane < EC ? esup[ebC + lane] : -1;
            eC1 = lane + 32 < EC ? esup[ebC + 32 + lane] : -1;
        }
    }
    if (p_first + step < n_points) {
        ebB = esup_ptr[p_first + step];
        EB = esup_ptr[p_first + step + 1] - ebB;
    }
    for (i64 p = p_first; p < n_points; p += step) {
        const int eb = ebC, E = EC, e0 = eC0, e1 = eC1;
        (void)eb;
        // next node's element ids, the node after's row pointer
        int eN0 = -1, eN1 = -1, ebA = 0, EA = 0;
        if (p + step < n_points && EB <= STAR_CAPE) {
            eN0 = lane < EB ? esup[ebB + lane] : -1;
            eN1 = lane + 32 < EB ? esup[ebB + 32 + lane] : -1;
        }
        if (p + 2 * step < n_points) {
            ebA = esup_ptr[p + 2 * step];
            EA = esup_ptr[p + 2 * step + 1] - ebA;
        }
        ebC = ebB; EC = EB; eC0 = eN0; eC1 = eN1; ebB = ebA; EB = EA;
        if (E > STAR_CAPE) {
            if (lane == 0) atomicExch(too_big, 1);
            continue;
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int e = h == 0 ? e0 : e1, i = lane + 32 * h;
            if (e >= 0) {
                s_es[wid][i] = e;
                s_type[wid][i] = etype[e];
                const int4 *rp = reinterpret_cast<const int4 *>(inpoel + (i64)e * SPE);
#pragma unroll
                for (int v = 0; v < SPE / 4; v++) reinterpret_cast<int4 *>(&s_conn[wid][i * SPE])[v] = rp[v];
            }
        }
        __syncwarp();
        // sides (element i, local face j) whose smallest node is p, compacted
        int P = 0;
        const int total = E * sfe;      // sfe = 4 or 6: the most local faces an element of this mesh has
        for (int idx0 = 0; idx0 < total; idx0 += 32) {
            const int idx = idx0 + lane;
            bool ok = false;
            int k0 = -1, k1 = -1, k2 = -1, i = 0, j = 0;
            if (idx < total) {
                i = idx / sfe;
                j = idx - i * sfe;
                const int t = s_type[wid][i];
                if (j < tab.nfael[t]) {
                    const int nj = tab.lnofa[t][j];
                    int nd[NPB_MX_PF];
                    bool has = false;
                    int mn = 0x7fffffff;
#pragma unroll
                    for (int k = 0; k < NPB_MX_PF; k++) {
                        nd[k] = (k < nj) ? s_conn[wid][i * SPE + tab.lpofa[t][j][k]] : 0x7fffffff;
                        has = has || nd[k] == (int)p;
                        mn = min(mn, nd[k]);
                    }
                    ok = has && mn == (int)p;
                    if (ok) {
                        // the other nodes ascending (p itself sorts first and is dropped; padding sorts last)
#define NPB_CSWAP(a, b) { int lo_ = min(a, b), hi_ = max(a, b); a = lo_; b = hi_; }
                        NPB_CSWAP(nd[0], nd[1]) NPB_CSWAP(nd[2], nd[3]) NPB_CSWAP(nd[0], nd[2]) NPB_CSWAP(nd[1], nd[3]) NPB_CSWAP(nd[1], nd[2])
#undef NPB_CSWAP
                        k0 = nd[1];
                        k1 = nd[2];
                        k2 = nd[3] == 0x7fffffff ? -1 : nd[3];
                        if (k1 == 0x7fffffff) k1 = -1;
                    }
                }
            }
            const unsigned bal = __ballot_sync(FULL, ok);
            if (P + __popc(bal) > STAR_PAIRS) {      // warp-uniform: more sides than the table holds
                if (lane == 0) atomicExch(too_big, 1);
                P = 0;
                break;
            }
            if (ok) {
                const int at = P + __popc(bal & ((1u << lane) - 1u));
                s_key[wid][at][0] = k0;
                s_key[wid][at][1] = k1;
                s_key[wid][at][2] = k2;
                s_ij[wid][at] = (unsigned short)(i * 8 + j);
            }
            P += __popc(bal);
        }
        __syncwarp();
        // pair the sides.  Up to 32 of them (every star of the BASELINE meshes): one side per lane, the lanes holding the
        // same node set find each other with three __match_any_sync, each side writes its own entry.  More: side a looks
        // for the first later side b with the same node set and writes both entries.
        if (P <= 32) {
            const bool live = lane < P;
            const int k0 = live ? s_key[wid][lane][0] : -2 - lane, k1 = live ? s_key[wid][lane][1] : -2 - lane,
                      k2 = live ? s_key[wid][lane][2] : -2 - lane;
            const unsigned peers = __match_any_sync(FULL, k0) & __match_any_sync(FULL, k1) & __match_any_sync(FULL, k2);
            const unsigned others = peers & ~(1u << lane);
            if (live && others) {
                const int ija = s_ij[wid][lane], ijb = s_ij[wid][__ffs(others) - 1];
                const int ea = s_es[wid][ija >> 3], eb2 = s_es[wid][ijb >> 3];
                if (ea != eb2) esuel[(i64)ea * sfe + (ija & 7)] = eb2;
            }
        } else
        for (int a0 = 0; a0 < P; a0 += 32) {
            const int a = a0 + lane;
            if (a < P) {
                const int k0 = s_key[wid][a][0], k1 = s_key[wid][a][1], k2 = s_key[wid][a][2];
                for (int b = a + 1; b < P; b++) {
                    if (s_key[wid][b][0] == k0 && s_key[wid][b][1] == k1 && s_key[wid][b][2] == k2) {
                        const int ija = s_ij[wid][a], ijb = s_ij[wid][b];
                        const int ea = s_es[wid][ija >> 3], eb2 = s_es[wid][ijb >> 3];
                        if (ea != eb2) {
                            esuel[(i64)ea * sfe + (ija & 7)] = eb2;
                            esuel[(i64)eb2 * sfe + (ijb & 7)] = ea;
                            break;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Face numbering.  (e, j) owns its face iff it has no neighbour or e < neighbour: that is exactly the
// face the reference's serial loop numbers when it first meets it (grid.pyx:315-334).
// ------------------------------------------------------------------------------------------------
__global__ void k_owner_count(ElemTables tab, const uint8_t *__restrict__ etype, const int32_t *__restrict__ esuel,
                              i64 n_elems, int sfe, int32_t *__restrict__ ownc)
{
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e > n_elems) return;
    if (e == n_elems) {
        ownc[e] = 0;
        return;
    }
    int nf = tab.nfael[etype[e]];
    int c = 0;
    for (int j = 0; j < nf; j++) {
        int nb = esuel[e * sfe + j];
        c += (nb < 0 || (int)e < nb) ? 1 : 0;
    }
    ownc[e] = c;
}

// infael for every (e, j); owners also emit inpofa (owner's local ordering, grid.pyx:340-345),
// esuf as the pair (owner, other), boundary tags (grid.pyx:434-444) and the node->face histogram.
__global__ void k_faces(ElemTables tab, const int32_t *__restrict__ inpoel, const uint8_t *__restrict__ etype,
                        const int32_t *__restrict__ esuel, const int32_t *__restrict__ fbase, i64 n_elems, int spe,
                        int sfe, int32_t *__restrict__ infael, int32_t *__restrict__ inpofa, int2 *__restrict__ esuf2,
                        uint8_t *__restrict__ bface, uint8_t *__restrict__ bpoint, int32_t *__restrict__ fcnt,
                        int *__restrict__ n_bfaces)
{
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    int t = etype[e];
    int nf = tab.nfael[t];
    int rank = 0;
    int base = fbase[e];
    for (int j = 0; j < sfe; j++) {
        if (j >= nf) {
            infael[e * sfe + j] = -1;
            continue;
        }
        int nb = esuel[e * sfe + j];
        if (nb < 0 || (int)e < nb) {
            int f = base + rank;
            rank++;
            infael[e * sfe + j] = f;
            int nj = tab.lnofa[t][j];
            int4 fn = make_int4(-1, -1, -1, -1);
            int *fp = &fn.x;
#pragma unroll
            for (int k = 0; k < NPB_MX_PF; k++)
                if (k < nj) {
                    int p = inpoel[e * spe + tab.lpofa[t][j][k]];
                    fp[k] = p;
                    atomicAdd(&fcnt[p], 1);
                    if (nb < 0) bpoint[p] = 1;
                }
            reinterpret_cast<int4 *>(inpofa)[f] = fn;
            esuf2[f] = make_int2((int)e, nb);
            bface[f] = nb < 0 ? 1 : 0;
            if (nb < 0) atomicAdd(n_bfaces, 1);
        } else {
            // the owner is nb (< e): its id for this face = its base + number of owned faces before the
            // first local face l with esuel[nb, l] == e (grid.pyx:331-334)
            int kt = etype[nb];
            int nfk = tab.nfael[kt];
            int r = 0;
            for (int l = 0; l < nfk; l++) {
                int nbl = esuel[(i64)nb * sfe + l];
                if (nbl == (int)e) break;
                r += (nbl < 0 || nb < nbl) ? 1 : 0;
            }
            infael[e * sfe + j] = fbase[nb] + r;
        }
    }
}

// node -> face fill (rows sorted afterwards by k_sort_rows)
__global__ void k_fill_fsup(const int32_t *__restrict__ inpofa, i64 n_faces, const int32_t *__restrict__ ptr,
                            int32_t *__restrict__ cursor, int32_t *__restrict__ fsup)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_faces * NPB_MX_PF) return;
    int p = inpofa[idx];
    if (p < 0) return;
    int pos = atomicAdd(&cursor[p], 1);
    fsup[(i64)ptr[p] + pos] = (int32_t)(idx / NPB_MX_PF);
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
int npb_k1_build(npb_ctx *c, const i64 *h_conn, int cstride, const i64 *h_types, const double *h_coords)
{
    const int T = 256;
    cudaStream_t s = c->stream;
    i64 ne = c->n_elems, np = c->n_points;
    int spe = c->spe, sfe = c->sfe;

    // ---- H2D of the reference-layout inputs, then compaction to int32 on the device ----
    {
        NpbTimer tm(c, "h2d_mesh");
        NpbTmp t_conn, t_types;
        NPB_CUDA(t_conn.alloc(sizeof(i64) * ne * cstride));
        NPB_CUDA(t_types.alloc(sizeof(i64) * ne));
        i64 *d_conn = t_conn.as<i64>(), *d_types = t_types.as<i64>();
        NPB_TRY(npb_h2d(c, d_conn, h_conn, sizeof(i64) * ne * cstride));
        NPB_TRY(npb_h2d(c, d_types, h_types, sizeof(i64) * ne));
        NPB_TRY(npb_alloc(c, (void **)&c->inpoel, sizeof(int32_t) * ne * spe));
        NPB_TRY(npb_alloc(c, (void **)&c->etype, ne));
        NPB_TRY(npb_alloc(c, (void **)&c->coords, sizeof(double) * np * 3));
        NPB_TRY(npb_h2d(c, c->coords, h_coords, sizeof(double) * np * 3));
        unsigned long long *d_bad = (unsigned long long *)(c->counters + 64);   // 8-byte aligned slot of the 128-int counter block
        NPB_CUDA(cudaMemsetAsync(d_bad, 0xff, sizeof(unsigned long long), s));
        k_convert_conn<<<npb_blocks(ne * spe, T), T, 0, s>>>(d_conn, cstride, d_types, ne, spe, c->tab, np, c->inpoel, c->etype, d_bad);
        NPB_LAUNCH(c);
        tm.stop();
        unsigned long long h_bad = 0;
        NPB_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost, s));
        NPB_CUDA(cudaStreamSynchronize(s));
        if (h_bad != ~0ull) {
            npb_set_error("element %llu has a node id outside [0, %lld) among its first npoel[type] entries", h_bad, (long long)np);
            return NPB_ERR_ARG;
        }
    }
    // Device time of K1 = sum of the kernel segments below; cudaMalloc / host round trips for the sizes
    // (n_faces, row totals) sit between the segments and are not counted.
    for (const char *k : {"k1", "k1_esup", "k1_esuel", "k1_faces", "k1_fsup", "k1_geom"}) c->timings[k] = 0.f;
    int *d_mx = c->counters;  // [0] = mx_epp, [1] = mx_fpp, [2] = boundary faces
    NPB_CUDA(cudaMemsetAsync(d_mx, 0, sizeof(int) * 8, s));

    // ---- esup ----
    NpbTmp t_cursor, t_fbase;
    NPB_TRY(npb_alloc(c, (void **)&c->esup_ptr, sizeof(int32_t) * (np + 1)));
    NPB_TRY(npb_alloc(c, (void **)&c->esuel, sizeof(int32_t) * ne * sfe));
    NPB_TRY(npb_alloc(c, (void **)&c->infael, sizeof(int32_t) * ne * sfe));
    NPB_TRY(npb_alloc(c, (void **)&c->bpoint, (size_t)np));
    NPB_TRY(npb_alloc(c, (void **)&c->fsup_ptr, sizeof(int32_t) * (np + 1)));
    NPB_CUDA(t_cursor.alloc(sizeof(int32_t) * (np + 1)));
    NPB_CUDA(t_fbase.alloc(sizeof(int32_t) * (ne + 1)));
    int32_t *cursor = t_cursor.as<int32_t>(), *fbase = t_fbase.as<int32_t>();
    {
        NpbTimer tm(c, "k1_esup", true);
        NPB_CUDA(cudaMemsetAsync(c->esup_ptr, 0, sizeof(int32_t) * (np + 1), s));
        k_count_nodes<<<npb_blocks(ne * spe, T), T, 0, s>>>(c->inpoel, ne * spe, c->esup_ptr);
        NPB_LAUNCH(c);
        NPB_TRY(npb_exclusive_scan_i32(c, c->esup_ptr, c->esup_ptr, np + 1));
        tm.stop();
    }
    {
        int32_t total = 0;
        NPB_CUDA(cudaMemcpyAsync(&total, c->esup_ptr + np, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        NPB_CUDA(cudaStreamSynchronize(s));
        c->len_esup = total;
        NPB_TRY(npb_alloc(c, (void **)&c->esup, sizeof(int32_t) * (size_t)(total > 0 ? total : 1)));
    }
    {
        NpbTimer tm(c, "k1_esup", true);
        NPB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (np + 1), s));
        k_fill_rows<<<npb_blocks(ne * spe, T), T, 0, s>>>(c->inpoel, ne * spe, spe, c->esup_ptr, cursor, c->esup);
        NPB_LAUNCH(c);
        k_sort_rows_smem<<<npb_blocks(np, SORT_ROWS), SORT_ROWS, 0, s>>>(c->esup_ptr, np, c->esup, d_mx + 0);
        NPB_LAUNCH(c);
        tm.stop();
    }
    // ---- esuel ----
    {
        NpbTimer tm(c, "k1_esuel", true);
        FaceMasks fm;
        for (int t = 0; t < NPB_N_TYPES; t++)
            for (int l = 0; l < NPB_MX_FE; l++) {
                unsigned char m = 0;
                for (int k = 0; k < c->tab.lnofa[t][l]; k++) m |= (unsigned char)(1u << c->tab.lpofa[t][l][k]);
                fm.m[t][l] = l < c->tab.nfael[t] ? m : 0;
            }
        bool done = false;
        const char *plain = getenv("NPB_K1_ESUEL_PLAIN");   // A/B timing and tests: the per-face candidate search
        if (!(plain && plain[0] == '1') && ne * (i64)sfe < (1ll << 31)) {
            int *too_big = d_mx + 3;
            NPB_CUDA(cudaMemsetAsync(c->esuel, 0xff, sizeof(int32_t) * ne * sfe, s));   // -1: boundary faces and unused slots
            const int grid = c->sm_count * 16;
            if (spe == 4)
                k_esuel_star<4><<<grid, 32 * STAR_WARPS, 0, s>>>(c->tab, c->inpoel, c->etype, c->esup_ptr, c->esup, np, sfe, c->esuel, too_big);
            else
                k_esuel_star<8><<<grid, 32 * STAR_WARPS, 0, s>>>(c->tab, c->inpoel, c->etype, c->esup_ptr, c->esup, np, sfe, c->esuel, too_big);
            NPB_LAUNCH(c);
            int h_big = 0;
            NPB_TRY(npb_read_int(c, too_big, &h_big));
            done = h_big == 0;      // a star of more than STAR_CAPE elements: the per-face kernel rebuilds everything
        }
        if (done) {
        } else if (spe == 4)
            k_esuel<4><<<npb_blocks(ne * sfe, 256), 256, 0, s>>>(c->tab, fm, c->inpoel, c->etype, c->esup_ptr, c->esup, ne, sfe, c->esuel);
        else
            k_esuel<8><<<npb_blocks(ne * sfe, 256), 256, 0, s>>>(c->tab, fm, c->inpoel, c->etype, c->esup_ptr, c->esup, ne, sfe, c->esuel);
        NPB_LAUNCH(c);
        tm.stop();
    }
    // ---- faces: numbering, inpofa, esuf, tags, fsup histogram ----
    {
        NpbTimer tm(c, "k1_faces", true);
        k_owner_count<<<npb_blocks(ne + 1, T), T, 0, s>>>(c->tab, c->etype, c->esuel, ne, sfe, fbase);
        NPB_LAUNCH(c);
        NPB_TRY(npb_exclusive_scan_i32(c, fbase, fbase, ne + 1));
        tm.stop();
    }
    {
        int32_t nfaces = 0;
        NPB_CUDA(cudaMemcpyAsync(&nfaces, fbase + ne, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        NPB_CUDA(cudaStreamSynchronize(s));
        c->n_faces = nfaces;
        size_t nf1 = (size_t)(nfaces > 0 ? nfaces : 1);
        NPB_TRY(npb_alloc(c, (void **)&c->inpofa, sizeof(int32_t) * nf1 * NPB_MX_PF));
        NPB_TRY(npb_alloc(c, (void **)&c->esuf2, sizeof(int2) * nf1));
        NPB_TRY(npb_alloc(c, (void **)&c->bface, nf1));
        NPB_TRY(npb_alloc(c, (void **)&c->centroids, sizeof(double) * ne * NPB_CSTRIDE));
        NPB_TRY(npb_alloc(c, (void **)&c->fcent, sizeof(double) * nf1 * 3));
        NPB_TRY(npb_alloc(c, (void **)&c->fnormal, sizeof(double) * nf1 * 3));
        NPB_TRY(npb_alloc(c, (void **)&c->farea, sizeof(double) * nf1));
    }
    {
        NpbTimer tm(c, "k1_faces", true);
        NPB_CUDA(cudaMemsetAsync(c->bpoint, 0, (size_t)np, s));
        NPB_CUDA(cudaMemsetAsync(c->fsup_ptr, 0, sizeof(int32_t) * (np + 1), s));
        k_faces<<<npb_blocks(ne, 128), 128, 0, s>>>(c->tab, c->inpoel, c->etype, c->esuel, fbase, ne, spe, sfe, c->infael,
                                                   c->inpofa, c->esuf2, c->bface, c->bpoint, c->fsup_ptr, d_mx + 2);
        NPB_LAUNCH(c);
        tm.stop();
    }
    // ---- fsup ----
    {
        NpbTimer tm(c, "k1_fsup", true);
        NPB_TRY(npb_exclusive_scan_i32(c, c->fsup_ptr, c->fsup_ptr, np + 1));
        tm.stop();
    }
    {
        int32_t total = 0;
        NPB_CUDA(cudaMemcpyAsync(&total, c->fsup_ptr + np, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        NPB_CUDA(cudaStreamSynchronize(s));
        c->len_fsup = total;
        NPB_TRY(npb_alloc(c, (void **)&c->fsup, sizeof(int32_t) * (size_t)(total > 0 ? total : 1)));
    }
    {
        NpbTimer tm(c, "k1_fsup", true);
        NPB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (np + 1), s));
        k_fill_fsup<<<npb_blocks(c->n_faces * NPB_MX_PF, T), T, 0, s>>>(c->inpofa, c->n_faces, c->fsup_ptr, cursor, c->fsup);
        NPB_LAUNCH(c);
        k_sort_rows_smem<<<npb_blocks(np, SORT_ROWS), SORT_ROWS, 0, s>>>(c->fsup_ptr, np, c->fsup, d_mx + 1);
        NPB_LAUNCH(c);
        tm.stop();
    }
    int h_mx[3] = {0, 0, 0};
    NPB_CUDA(cudaMemcpyAsync(h_mx, d_mx, sizeof(int) * 3, cudaMemcpyDeviceToHost, s));
    NPB_CUDA(cudaStreamSynchronize(s));
    c->mx_epp = h_mx[0];
    c->mx_fpp = h_mx[1];
    // esuf rows are [owner] or [owner, other] (grid.pyx:390-416)
    c->len_esuf = 2 * c->n_faces - h_mx[2];
    c->mx_epf = c->n_faces == 0 ? 0 : (h_mx[2] < c->n_faces ? 2 : 1);
    // ---- geometry (k1_geometry.cu, compiled without FMA contraction) ----
    NPB_TRY(npb_k1_geometry(c));
    for (const char *k : {"k1_esup", "k1_esuel", "k1_faces", "k1_fsup", "k1_geom"}) c->timings["k1"] += c->timings[k];
    return NPB_OK;
}
