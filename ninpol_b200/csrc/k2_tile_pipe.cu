// k2_tile_pipe.cu — software-pipelined IDW / LS tile kernels (kernel groups K2 + K3), the HBM-roofline path.
//
// Same arithmetic and the same tile decomposition as k2_idw_ls_tile.cu (reference ninpol/_methods/idw.pyx:35-84,
// ls.pyx:33-135; evaluation order of SURVEY.md App. C, no FMA: -fmad=false plus _rn intrinsics) — a CTA owns tiles
// of NB consecutive nodes = one contiguous slice of esup — but the three dependent global round trips of a tile
// (esup_ptr slice -> esup slice -> centroid gather) no longer sit in front of its arithmetic.  They run one and
// two tiles AHEAD of it:
//   iteration t of a CTA:
//     * one elected thread issues TMA bulk copies (cp.async.bulk, completion on an mbarrier) of the CONTIGUOUS
//       inputs of tile t+2: esup_ptr / boundary flag / neumann flag / coordinate / indptr slices of its nodes and
//       its esup slice (whose bounds that thread fetched one iteration earlier);
//     * all threads wait for the mbarrier of tile t+1 and issue the GATHER of its centroids as 8-byte cp.async
//       (LDGSTS: global -> shared without register staging), 3 per esup entry;
//     * the arithmetic of tile t runs on data that landed in shared memory during iteration t-1:
//       per-entry phase (IDW: reciprocal distances), per-node phase (the order-sensitive sequential sums of the
//       reference), per-entry phase (final divide, coalesced stores of (index, value) into the CSR).
// Buffers: 3-deep ring for the bulk-copied slices, 2-deep for the gathered centroids; two block barriers per tile
// (three for IDW) instead of five.
#include <stdlib.h>
#include "common.cuh"
#include "tile_args.cuh"

#define IDW_EPS ((double)1.0000000036274937e-15f) /* float32(1e-15), idw.pyx:53 */
#define M2(a, b) __dmul_rn(a, b)
#define S2(a, b) __dsub_rn(a, b)
#define A2(a, b) __dadd_rn(a, b)

// ---- PTX wrappers: mbarrier, TMA bulk copy, cp.async ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared; 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_8(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// A slice [src, src + bytes) of a global array whose start is only element-aligned: the copy starts at the 16-byte
// boundary below it (`lead` bytes early) and is rounded up to 16 bytes; every array read this way is allocated with
// slack at its end (npb_alloc).  Returns the rounded size; the wanted data begins at dst + lead.
__device__ __forceinline__ unsigned bulk_slice(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned &lead)
{
    const unsigned long long s = (unsigned long long)src;
    lead = (unsigned)(s & 15ull);
    const unsigned size = (lead + bytes + 15u) & ~15u;
    tma_bulk_g2s(dst, (const void *)(s - lead), size, bar);
    return size;
}

template <int NB, int ECAP>
struct PipeSmem {
    // ring of 3: contiguous slices (16-byte aligned starts; up to 15 lead bytes + 15 tail bytes of slack each)
    // (every ring row is a multiple of 16 bytes: each slot is the 16-byte aligned destination of a bulk copy)
    alignas(16) int ptr[3][(NB + 1 + 8 + 3) & ~3];
    alignas(16) int out[3][(NB + 1 + 8 + 3) & ~3];
    alignas(16) double xyz[3][3 * NB + 4];
    alignas(16) unsigned char bp[3][NB + 32];
    alignas(16) unsigned char nf[3][NB + 32];
    alignas(16) int es[3][ECAP + 8];
    unsigned lead[3][6];          // byte offset of the wanted data inside each slice buffer (ptr, out, xyz, bp, nf, es)
    alignas(8) unsigned long long bar[3];
    // ring of 2: gathered centroids, one padded SoA plane per coordinate (entry i of node k at i + k: the per-node
    // sequential reads of phase B then fall into different banks)
    alignas(16) double cx[2][ECAP + NB], cy[2][ECAP + NB], cz[2][ECAP + NB];
    unsigned char rid[2][ECAP];   // node (within the tile) of every entry
    // per-tile scratch
    double lam[NB * 4];
    double tot[NB];
    int fz[NB], cnt[NB];
    unsigned char mode[NB], proc[NB];
};

static_assert((3 * 64 + 4) % 2 == 0 && (64 + 32) % 16 == 0 && (32 + 32) % 16 == 0 && (16 + 32) % 16 == 0 && (1408 + 8) % 4 == 0 &&
                  (768 + 8) % 4 == 0 && (384 + 8) % 4 == 0,
              "ring rows of PipeSmem must be multiples of 16 bytes");

template <int METHOD, int NB, int ECAP, int T, int MINB>
__global__ void __launch_bounds__(T, MINB) k_tile_pipe(TileArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef PipeSmem<NB, ECAP> S;
    S &s = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x;
    const i64 ntiles = (a.p_hi - a.p_lo + a.nb - 1) / a.nb;
    const i64 my_tiles = ((i64)blockIdx.x < ntiles) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (my_tiles == 0) return;
    if (tid == 0) {
        for (int q = 0; q < 3; q++) mbar_init(&s.bar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto tile_p0 = [&](i64 k) { return a.p_lo + ((i64)blockIdx.x + k * gridDim.x) * a.nb; };
    auto tile_nb = [&](i64 k) { return (int)min((i64)a.nb, a.p_hi - tile_p0(k)); };

    // ---- stage 1 (thread 0): TMA bulk copies of the contiguous inputs of local tile k into ring slot k % 3 ----
    int nxt_eb = 0, nxt_ee = 0;   // esup bounds of the NEXT tile to be bulk-loaded (fetched one iteration ahead)
    auto fetch_bounds = [&](i64 k) {
        if (k < my_tiles) {
            nxt_eb = a.esup_ptr[tile_p0(k)];
            nxt_ee = a.esup_ptr[tile_p0(k) + tile_nb(k)];
        }
    };
    auto issue_bulk = [&](i64 k) {   // uses nxt_eb / nxt_ee fetched for tile k
        if (k >= my_tiles) return;
        const int q = (int)(k % 3);
        const i64 p0 = tile_p0(k);
        const int nb = tile_nb(k);
        unsigned total = 0, lead;
        unsigned sizes[6];
        const void *srcs[6] = {a.esup_ptr + p0, a.indptr + p0, a.coords + p0 * 3, a.bpoint + p0, a.nflag + p0, a.esup + nxt_eb};
        void *dsts[6] = {s.ptr[q], s.out[q], s.xyz[q], s.bp[q], s.nf[q], s.es[q]};
        const unsigned want[6] = {(unsigned)(nb + 1) * 4u, (unsigned)(nb + 1) * 4u, (unsigned)nb * 24u, (unsigned)nb, (unsigned)nb,
                                  (unsigned)(nxt_ee - nxt_eb) * 4u};
        // expected bytes first (the barrier must know them before a copy can complete the phase)
#pragma unroll
        for (int j = 0; j < 6; j++) {
            const unsigned l = (unsigned)((unsigned long long)srcs[j] & 15ull);
            sizes[j] = (j == 1 && !a.direct) || want[j] == 0 ? 0u : ((l + want[j] + 15u) & ~15u);
            total += sizes[j];
        }
        mbar_expect_tx(&s.bar[q], total);
#pragma unroll
        for (int j = 0; j < 6; j++) {
            lead = 0;
            if (sizes[j]) bulk_slice(dsts[j], srcs[j], want[j], &s.bar[q], lead);
            s.lead[q][j] = lead;
        }
    };
    // ---- stage 2 (all threads): centroid gather of local tile k into ring slot k & 1 (needs stage 1 of tile k) ----
    auto issue_gather = [&](i64 k) {
        if (k >= my_tiles) return;
        const int q = (int)(k % 3), h = (int)(k & 1);
        mbar_wait(&s.bar[q], (unsigned)((k / 3) & 1));
        const int nb = tile_nb(k);
        const int *ptr = reinterpret_cast<const int *>(reinterpret_cast<const unsigned char *>(s.ptr[q]) + s.lead[q][0]);
        const int *es = reinterpret_cast<const int *>(reinterpret_cast<const unsigned char *>(s.es[q]) + s.lead[q][5]);
        const int eb = ptr[0], ne = ptr[nb] - eb;
        for (int i = tid; i < ne; i += T) {
            // node of entry i: last k with ptr[k] - eb <= i
            int lo = 0, hi = nb;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (ptr[mid] - eb <= i) lo = mid; else hi = mid;
            }
            s.rid[h][i] = (unsigned char)lo;
            const double *cc = a.cent + (i64)es[i] * NPB_CSTRIDE;
            cp_async_8(&s.cx[h][i + lo], cc);
            cp_async_8(&s.cy[h][i + lo], cc + 1);
            cp_async_8(&s.cz[h][i + lo], cc + 2);
        }
        cp_async_commit();
    };

    // ---- prologue: tiles 0 and 1 ----
    if (tid == 0) {
        fetch_bounds(0);
        issue_bulk(0);
        fetch_bounds(1);
        issue_bulk(1);
        fetch_bounds(2);
    }
    __syncthreads();      // the slice offsets thread 0 recorded (s.lead) are visible to everybody
    issue_gather(0);

    for (i64 k = 0; k < my_tiles; k++) {
        const int q = (int)(k % 3), h = (int)(k & 1);
        cp_async_wait0();     // my share of the gather of tile k has landed
        __syncthreads();      // ... and everybody's; everybody is done with tile k-1 (its ring slots are free)
        if (tid == 0) {
            issue_bulk(k + 2);
            fetch_bounds(k + 3);
        }
        issue_gather(k + 1);

        const i64 p0 = tile_p0(k);
        const int nb = tile_nb(k);
        const int *ptr = reinterpret_cast<const int *>(reinterpret_cast<const unsigned char *>(s.ptr[q]) + s.lead[q][0]);
        const int *out = reinterpret_cast<const int *>(reinterpret_cast<const unsigned char *>(s.out[q]) + s.lead[q][1]);
        const double *xyz = reinterpret_cast<const double *>(reinterpret_cast<const unsigned char *>(s.xyz[q]) + s.lead[q][2]);
        const unsigned char *bp = s.bp[q] + s.lead[q][3], *nf = s.nf[q] + s.lead[q][4];
        const int *es = reinterpret_cast<const int *>(reinterpret_cast<const unsigned char *>(s.es[q]) + s.lead[q][5]);
        const int eb = ptr[0], ne = ptr[nb] - eb;
        double *cx = s.cx[h], *cy = s.cy[h], *cz = s.cz[h];
        const unsigned char *rid = s.rid[h];

        if (METHOD == NPB_METHOD_IDW) {
            // per entry: reciprocal distance (idw.pyx:64-78), overwriting the x plane
            for (int i = tid; i < ne; i += T) {
                const int node = rid[i];
                if (bp[node] && !nf[node]) continue;
                double d0 = S2(xyz[3 * node], cx[i + node]);
                double dist = A2(0.0, M2(d0, d0));
                if (a.dim > 1) {
                    double d1 = S2(xyz[3 * node + 1], cy[i + node]);
                    dist = A2(dist, M2(d1, d1));
                }
                if (a.dim > 2) {
                    double d2 = S2(xyz[3 * node + 2], cz[i + node]);
                    dist = A2(dist, M2(d2, d2));
                }
                // coincident centroid (idw.pyx:69): marked with -1 (a reciprocal distance is never negative)
                cx[i + node] = (dist <= IDW_EPS) ? -1.0 : __ddiv_rn(1.0, __dsqrt_rn(dist));
            }
            __syncthreads();
        }
        // ---- per node: the reference's sequential sums, in esup order ----
        if (tid < nb) {
            const bool proc = !(bp[tid] && !nf[tid]);
            s.proc[tid] = proc;
            s.cnt[tid] = 0;
            if (proc) {
                const int b = ptr[tid] - eb + tid, E = ptr[tid + 1] - ptr[tid];
                if (METHOD == NPB_METHOD_IDW) {
                    double total = 0.0;
                    int fz = -1;
                    for (int j = 0; j < E; j++) {   // idw.pyx:79
                        double v = cx[b + j];
                        if (v == -1.0) {
                            fz = j;
                            break;
                        }
                        total = A2(total, v);
                    }
                    s.tot[tid] = total;
                    s.fz[tid] = fz;
                } else {
                    const double x0 = xyz[3 * tid], x1 = xyz[3 * tid + 1], x2 = xyz[3 * tid + 2];
                    double Ix = 0.0, Iy = 0.0, Iz = 0.0, Ixx = 0.0, Ixy = 0.0, Ixz = 0.0, Iyy = 0.0, Iyz = 0.0, Izz = 0.0;
                    for (int j = 0; j < E; j++) {   // ls.pyx:64-77
                        double vx = S2(cx[b + j], x0), vy = S2(cy[b + j], x1), vz = S2(cz[b + j], x2);
                        Ix = A2(Ix, vx); Iy = A2(Iy, vy); Iz = A2(Iz, vz);
                        Ixx = A2(Ixx, M2(vx, vx)); Ixy = A2(Ixy, M2(vx, vy)); Ixz = A2(Ixz, M2(vx, vz));
                        Iyy = A2(Iyy, M2(vy, vy)); Iyz = A2(Iyz, M2(vy, vz)); Izz = A2(Izz, M2(vz, vz));
                    }
                    bool flat = (Iz == 0.0 && Izz == 0.0 && Ixz == 0.0 && Iyz == 0.0);
                    if (flat) Izz = 1.0;   // ls.pyx:79-80
                    double D = A2(A2(M2(Ixx, S2(M2(Iyy, Izz), M2(Iyz, Iyz))), M2(Ixy, S2(M2(Iyz, Ixz), M2(Ixy, Izz)))),
                                  M2(Ixz, S2(M2(Ixy, Iyz), M2(Iyy, Ixz))));
                    if (D == 0.0) {  // inverse-distance fallback (ls.pyx:88-102): total of 1/|v| in esup order
                        double total = 0.0;
                        for (int j = 0; j < E; j++) {
                            double vx = S2(cx[b + j], x0), vy = S2(cy[b + j], x1), vz = S2(cz[b + j], x2);
                            total = A2(total, __ddiv_rn(1.0, __dsqrt_rn(A2(A2(M2(vx, vx), M2(vy, vy)), M2(vz, vz)))));
                        }
                        s.lam[4 * tid + 3] = total;
                        s.mode[tid] = 1;
                    } else {
                        // ls.pyx:105-106 repeats the test, but Izz is already 1.0 whenever it held: restated literally
                        if (Iz == 0.0 && Izz == 0.0 && Ixz == 0.0 && Iyz == 0.0) Izz = -1.0;
                        double lx = __ddiv_rn(A2(A2(M2(Ix, S2(M2(Iyz, Iyz), M2(Iyy, Izz))), M2(Iy, S2(M2(Ixy, Izz), M2(Iyz, Ixz)))),
                                                 M2(Iz, S2(M2(Iyy, Ixz), M2(Ixy, Iyz)))), D);
                        double ly = __ddiv_rn(A2(A2(M2(Ix, S2(M2(Ixy, Izz), M2(Iyz, Ixz))), M2(Iy, S2(M2(Ixz, Ixz), M2(Ixx, Izz)))),
                                                 M2(Iz, S2(M2(Ixx, Iyz), M2(Ixy, Ixz)))), D);
                        double lz = __ddiv_rn(A2(A2(M2(Ix, S2(M2(Iyy, Ixz), M2(Ixy, Iyz))), M2(Iy, S2(M2(Ixx, Iyz), M2(Ixy, Ixz)))),
                                                 M2(Iz, S2(M2(Ixy, Ixy), M2(Ixx, Iyy)))), D);
                        s.lam[4 * tid] = lx;
                        s.lam[4 * tid + 1] = ly;
                        s.lam[4 * tid + 2] = lz;
                        s.lam[4 * tid + 3] = A2(A2(A2((double)E, M2(lx, Ix)), M2(ly, Iy)), M2(lz, Iz));   // denom, ls.pyx:126
                        s.mode[tid] = 0;
                    }
                }
            }
        }
        __syncthreads();
        // ---- per entry: normalise and emit ----
        for (int i = tid; i < ne; i += T) {
            const int node = rid[i];
            if (!s.proc[node]) {
                if (!a.direct) a.wbuf[(i64)eb + i - a.wbase] = 0.0;
                continue;
            }
            const int j = i - (ptr[node] - eb);
            double w;
            if (METHOD == NPB_METHOD_IDW) {
                const int fz = s.fz[node];
                if (fz >= 0)
                    w = (j == fz) ? 1.0 : 0.0;
                else
                    w = A2(__ddiv_rn(cx[i + node], s.tot[node]), 0.0);
            } else {
                double vx = S2(cx[i + node], xyz[3 * node]), vy = S2(cy[i + node], xyz[3 * node + 1]),
                       vz = S2(cz[i + node], xyz[3 * node + 2]);
                const double den = s.lam[4 * node + 3];
                if (s.mode[node] == 0)
                    w = A2(A2(A2(1.0, M2(s.lam[4 * node], vx)), M2(s.lam[4 * node + 1], vy)), M2(s.lam[4 * node + 2], vz));
                else
                    w = __ddiv_rn(1.0, __dsqrt_rn(A2(A2(M2(vx, vx), M2(vy, vy)), M2(vz, vz))));
                w = A2(__ddiv_rn(w, den), 0.0);
            }
            if (a.direct) {
                const i64 pos = (i64)out[node] + j;
                a.data[pos] = w;
                a.indices[pos] = es[i];
                if (w == 0.0) atomicAdd(a.zero_counter, 1);
            } else {
                a.wbuf[(i64)eb + i - a.wbase] = w;
                if (w != 0.0) atomicAdd(&s.cnt[node], 1);
            }
        }
        if (!a.direct) {
            __syncthreads();
            if (tid < nb) {
                a.rowcnt[p0 + tid] = s.cnt[tid];
                a.neumann[p0 + tid] = 0.0;
            }
        }
    }
}

// Tile shapes (nodes x entries, threads, resident CTAs the launch bound asks for).  The prefetch rings cost 60 bytes of
// shared memory per entry, so big tiles mean few resident warps while the block barriers between the phases want many:
//   A  64 x 1408, 256 threads, 2 CTAs / SM (~100 KB each)      B  32 x 768, 128 threads, 4 CTAs / SM (~54 KB)
//   C  16 x 384,  128 threads, 5 CTAs / SM (~28 KB)            D  16 x 384,  64 threads, 8 CTAs / SM
template <int METHOD, int NB, int ECAP, int T, int MINB>
static int launch_variant(npb_ctx *c, TileArgs a)
{
    typedef PipeSmem<NB, ECAP> S;
    const int smem = (int)sizeof(S);
    auto kern = k_tile_pipe<METHOD, NB, ECAP, T, MINB>;
    static bool configured = false;
    if (!configured) {
        NPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    a.nb = ECAP / c->mx_epp > NB ? NB : ECAP / c->mx_epp;
    if (a.nb < 1) return NPB_ERR_ARG;
    const i64 ntiles = (a.p_hi - a.p_lo + a.nb - 1) / a.nb;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > MINB) per_sm = MINB;
    const i64 cap = (i64)c->sm_count * per_sm;
    int grid = (int)(ntiles < cap ? ntiles : cap);
    if (grid < 1) return NPB_OK;
    kern<<<grid, T, smem, c->stream>>>(a);
    return NPB_OK;
}

template <int METHOD>
static int launch_shape(npb_ctx *c, const TileArgs &a, char shape)
{
    switch (shape) {
    case 'A': return launch_variant<METHOD, 64, 1408, 256, 2>(c, a);
    case 'B': return launch_variant<METHOD, 32, 768, 128, 4>(c, a);
    case 'C': return launch_variant<METHOD, 16, 384, 128, 5>(c, a);
    case 'D': return launch_variant<METHOD, 16, 384, 64, 8>(c, a);
    }
    return NPB_ERR_ARG;
}

// *used = 1 when the pipelined kernel took the launch, 0 when the caller should use the plain tile kernel.
// Selected with NPB_TILE_PIPE=A|B|C|D (measured on B200: see DESIGN.md for which shape, if any, beats the plain
// kernels); any tile size gives the same results.
int npb_tile_pipe_launch(npb_ctx *c, const TileArgs &a, int method, int *used)
{
    *used = 0;
    const char *sel = getenv("NPB_TILE_PIPE");
    if (!sel || !sel[0] || sel[0] == '0') return NPB_OK;
    char shape = sel[0];
    if (shape < 'A' || shape > 'D' || c->mx_epp < 1) return NPB_OK;
    const int ecap = shape == 'A' ? 1408 : shape == 'B' ? 768 : 384;
    if (c->mx_epp > ecap) return NPB_OK;
    if (method == NPB_METHOD_IDW) NPB_TRY(launch_shape<NPB_METHOD_IDW>(c, a, shape));
    else NPB_TRY(launch_shape<NPB_METHOD_LS>(c, a, shape));
    *used = 1;
    return NPB_OK;
}
