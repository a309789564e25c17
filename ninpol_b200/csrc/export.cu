// export.cu — Grid attribute export in the reference's dtypes and shapes (grid.pxd:128-187), plus the
// structures no interpolation method consumes and that are therefore built lazily: psup
// (Grid.build_psup, ninpol/_interpolator/grid.pyx:269-302).  Device arrays are int32 / compact; the
// widening to int64, the -1 padding to [n,8] / [n,6] / [n,4] and the (ptr, flat) form of esuf are done
// by kernels here, then copied to caller memory.
#include <string.h>
#include "common.cuh"

__global__ void k_widen_rows(const int32_t *__restrict__ src, i64 rows, int src_stride, int dst_stride, i64 *__restrict__ dst)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * dst_stride) return;
    i64 r = idx / dst_stride;
    int j = (int)(idx - r * dst_stride);
    dst[idx] = j < src_stride ? (i64)src[r * src_stride + j] : -1;
}
__global__ void k_widen_u8(const uint8_t *__restrict__ src, i64 n, i64 *__restrict__ dst)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void k_esuf_count(const int2 *__restrict__ esuf2, i64 n_faces, int32_t *__restrict__ cnt)
{
    i64 f = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (f > n_faces) return;
    cnt[f] = f == n_faces ? 0 : (esuf2[f].y >= 0 ? 2 : 1);
}
__global__ void k_esuf_flat(const int2 *__restrict__ esuf2, const int32_t *__restrict__ ptr, i64 n_faces, i64 *__restrict__ out)
{
    i64 f = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    int2 e = esuf2[f];
    out[ptr[f]] = e.x;
    if (e.y >= 0) out[ptr[f] + 1] = e.y;
}

// ---- psup: first-occurrence order over (esup row, local node), grid.pyx:285-299 ----
#define PSUP_CAP 256
template <bool FILL>
__global__ void k_psup(ElemTables tab, const int32_t *__restrict__ inpoel, const uint8_t *__restrict__ etype,
                       const int32_t *__restrict__ esup_ptr, const int32_t *__restrict__ esup, i64 n_points, int spe,
                       int32_t *__restrict__ cnt_or_ptr, int32_t *__restrict__ psup, int *__restrict__ flags)
{
    i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_points) return;
    if (p == n_points) {
        if (!FILL) cnt_or_ptr[p] = 0;
        return;
    }
    int seen[PSUP_CAP];
    int n = 0;
    for (int q = esup_ptr[p]; q < esup_ptr[p + 1]; q++) {
        int e = esup[q];
        int npe = tab.npoel[etype[e]];
        for (int k = 0; k < npe; k++) {
            int v = inpoel[(i64)e * spe + k];
            if (v == (int)p) continue;
            bool dup = false;
            for (int s = 0; s < n; s++) dup = dup || (seen[s] == v);
            if (dup) continue;
            if (n < PSUP_CAP) seen[n] = v;
            else atomicExch(&flags[0], 1);
            n++;
        }
    }
    if (FILL) {
        int b = cnt_or_ptr[p];
        for (int s = 0; s < n && s < PSUP_CAP; s++) psup[b + s] = seen[s];
    } else {
        cnt_or_ptr[p] = n;
        atomicMax(&flags[1], n);
    }
}

static int out_i64(npb_ctx *c, const i64 *dev, i64 n, void *out, i64 cap);

static int build_psup(npb_ctx *c)
{
    if (c->psup_ptr) return NPB_OK;
    cudaStream_t s = c->stream;
    i64 np = c->n_points;
    int *flags = c->counters + 32;
    NPB_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * 2, s));
    int32_t *ptr = nullptr;
    NPB_TRY(npb_alloc(c, (void **)&ptr, sizeof(int32_t) * (np + 1)));
    k_psup<false><<<npb_blocks(np + 1, 128), 128, 0, s>>>(c->tab, c->inpoel, c->etype, c->esup_ptr, c->esup, np, c->spe, ptr,
                                                         nullptr, flags);
    NPB_LAUNCH(c);
    int h[2] = {0, 0};
    NPB_CUDA(cudaMemcpyAsync(h, flags, sizeof(int) * 2, cudaMemcpyDeviceToHost, s));
    NPB_CUDA(cudaStreamSynchronize(s));
    if (h[0]) {
        npb_set_error("psup: a node has more than %d neighbours", PSUP_CAP);
        return NPB_ERR_RANGE;
    }
    c->mx_ppp = h[1];
    NPB_TRY(npb_exclusive_scan_i32(c, ptr, ptr, np + 1));
    int32_t total = 0;
    NPB_CUDA(cudaMemcpyAsync(&total, ptr + np, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    NPB_CUDA(cudaStreamSynchronize(s));
    c->len_psup = total;
    NPB_TRY(npb_alloc(c, (void **)&c->psup, sizeof(int32_t) * (size_t)(total > 0 ? total : 1)));
    k_psup<true><<<npb_blocks(np + 1, 128), 128, 0, s>>>(c->tab, c->inpoel, c->etype, c->esup_ptr, c->esup, np, c->spe, ptr,
                                                        c->psup, flags);
    NPB_LAUNCH(c);
    c->psup_ptr = ptr;
    return NPB_OK;
}


// ---- edges: Grid.build_inedel, grid.pyx:527-580 -------------------------------------------------
// The reference numbers edges in first-encounter order over (element, local edge) and identifies an
// edge by the 32-bit truncation of myhash(sorted end points) (grid.pyx:29-43, unordered_map[int,int]
// at :539), so two edges whose hashes collide share one id.  Device restatement: hash every (e, j),
// stable radix sort of (hash, slot) pairs -> the head of each run is the first encounter; an exclusive
// scan of the head flags IN SLOT ORDER gives the first-encounter numbering; the head's id is copied
// down its run.  Bit-identical to the serial loop, collisions included.
__device__ __forceinline__ unsigned ref_myhash2_lo32(int a, int b)
{
    unsigned long long seed = 2ull;
    int v[2] = {a, b};
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int x = v[i];
        x = (int)((unsigned)((x >> 16) ^ x) * 0x45d9f3bu);
        x = (int)((unsigned)((x >> 16) ^ x) * 0x45d9f3bu);
        x = (x >> 16) ^ x;
        seed ^= (unsigned long long)((unsigned)x + 0x9e3779b9u) + (seed << 6) + (seed >> 2);
    }
    return (unsigned)seed;
}

__global__ void k_edge_count(EdgeTables et, const uint8_t *__restrict__ etype, i64 n_elems, int32_t *__restrict__ cnt)
{
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e > n_elems) return;
    cnt[e] = e == n_elems ? 0 : et.nedel[etype[e]];
}

__global__ void k_edge_keys(EdgeTables et, const int32_t *__restrict__ inpoel, const uint8_t *__restrict__ etype,
                            const int32_t *__restrict__ ebase, i64 n_elems, int spe, uint32_t *__restrict__ keys,
                            uint32_t *__restrict__ slots)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_elems * NPB_MX_EE) return;
    i64 e = idx / NPB_MX_EE;
    int j = (int)(idx - e * NPB_MX_EE);
    int t = etype[e];
    if (j >= et.nedel[t]) return;
    int a = inpoel[e * spe + et.lpoed[t][j][0]], b = inpoel[e * spe + et.lpoed[t][j][1]];
    int s0 = a < b ? a : b, s1 = a < b ? b : a;
    i64 q = (i64)ebase[e] + j;
    keys[q] = ref_myhash2_lo32(s0, s1);
    slots[q] = (uint32_t)idx;
}

// sorted order: head flags, and the same flags scattered to slot order (as 0/1 for the numbering scan)
__global__ void k_edge_heads(const uint32_t *__restrict__ skeys, const uint32_t *__restrict__ sslots, i64 n,
                             uint2 *__restrict__ pairs, int32_t *__restrict__ first_at_slot)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool head = (i == 0) || (skeys[i] != skeys[i - 1]);
    pairs[i] = make_uint2(head ? 1u : 0u, sslots[i]);
    if (head) first_at_slot[sslots[i]] = 1;
}

__global__ void k_edge_emit(const uint2 *__restrict__ pairs, const uint32_t *__restrict__ sslots, i64 n,
                            const int32_t *__restrict__ number_at_slot, EdgeTables et, const int32_t *__restrict__ inpoel,
                            const uint8_t *__restrict__ etype, int spe, i64 *__restrict__ inedel, i64 *__restrict__ inpoed)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t head_slot = pairs[i].y, slot = sslots[i];
    int id = number_at_slot[head_slot];
    inedel[slot] = id;
    if (slot == head_slot) {
        i64 e = slot / NPB_MX_EE;
        int j = (int)(slot - e * NPB_MX_EE);
        int t = etype[e];
        inpoed[(i64)id * 2 + 0] = inpoel[e * spe + et.lpoed[t][j][0]];
        inpoed[(i64)id * 2 + 1] = inpoel[e * spe + et.lpoed[t][j][1]];
    }
}

__global__ void k_fill_i64(i64 *p, i64 n, i64 v)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// builds c->inedel_d [n_elems,12] and c->inpoed_d [n_edges,2] (int64, reference layout) once
static int build_edges(npb_ctx *c)
{
    if (c->inedel_d) return NPB_OK;
    cudaStream_t s = c->stream;
    i64 ne = c->n_elems, nslots = ne * NPB_MX_EE;
    if (nslots >= (1ll << 31)) {
        npb_set_error("edge table too large for 32-bit slots");
        return NPB_ERR_RANGE;
    }
    NpbTmp t_ebase, t_flag, t_keys, t_slots, t_skeys, t_sslots, t_pairs;   // freed on every return path
    NPB_CUDA(t_ebase.alloc(sizeof(int32_t) * (ne + 1)));
    int32_t *ebase = t_ebase.as<int32_t>();
    k_edge_count<<<npb_blocks(ne + 1, 256), 256, 0, s>>>(c->etab, c->etype, ne, ebase);
    NPB_LAUNCH(c);
    NPB_TRY(npb_exclusive_scan_i32(c, ebase, ebase, ne + 1));
    int32_t total = 0;
    NPB_CUDA(cudaMemcpyAsync(&total, ebase + ne, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    NPB_CUDA(cudaStreamSynchronize(s));
    size_t n1 = (size_t)(total > 0 ? total : 1);
    NPB_CUDA(t_keys.alloc(4 * n1));
    NPB_CUDA(t_slots.alloc(4 * n1));
    NPB_CUDA(t_skeys.alloc(4 * n1));
    NPB_CUDA(t_sslots.alloc(4 * n1));
    NPB_CUDA(t_pairs.alloc(8 * n1));
    NPB_CUDA(t_flag.alloc(sizeof(int32_t) * (size_t)(nslots + 1)));
    uint32_t *keys = t_keys.as<uint32_t>(), *slots = t_slots.as<uint32_t>();
    uint32_t *skeys = t_skeys.as<uint32_t>(), *sslots = t_sslots.as<uint32_t>();
    uint2 *pairs = t_pairs.as<uint2>();
    int32_t *flag = t_flag.as<int32_t>();
    NPB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t) * (size_t)(nslots + 1), s));
    NPB_TRY(npb_alloc(c, (void **)&c->inedel_d, sizeof(i64) * (size_t)nslots));
    k_fill_i64<<<npb_blocks(nslots, 256), 256, 0, s>>>(c->inedel_d, nslots, -1);
    NPB_LAUNCH(c);
    if (total > 0) {
        k_edge_keys<<<npb_blocks(nslots, 256), 256, 0, s>>>(c->etab, c->inpoel, c->etype, ebase, ne, c->spe, keys, slots);
        NPB_LAUNCH(c);
        NPB_TRY(npb_sort_pairs_u32(c, keys, skeys, slots, sslots, total));
        k_edge_heads<<<npb_blocks(total, 256), 256, 0, s>>>(skeys, sslots, total, pairs, flag);
        NPB_LAUNCH(c);
        NPB_TRY(npb_propagate_heads(c, pairs, total));
        NPB_TRY(npb_exclusive_scan_i32(c, flag, flag, nslots + 1));
    }
    int32_t n_edges = 0;
    NPB_CUDA(cudaMemcpyAsync(&n_edges, flag + nslots, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    NPB_CUDA(cudaStreamSynchronize(s));
    c->n_edges = n_edges;
    NPB_TRY(npb_alloc(c, (void **)&c->inpoed_d, sizeof(i64) * 2 * (size_t)(n_edges > 0 ? n_edges : 1)));
    if (total > 0) {
        k_edge_emit<<<npb_blocks(total, 256), 256, 0, s>>>(pairs, sslots, total, flag, c->etab, c->inpoel, c->etype, c->spe,
                                                          c->inedel_d, c->inpoed_d);
        NPB_LAUNCH(c);
    }
    NPB_CUDA(cudaStreamSynchronize(s));
    return NPB_OK;
}

int npb_edges_stats(npb_ctx *c) { return c->build_edges ? build_edges(c) : NPB_OK; }

int npb_k1_extras(npb_ctx *c)
{
    // psup is built on first use (npb_export_array / npb_psup_stats); edges at load time when asked for,
    // like the reference (grid.pyx:219-227)
    return c->build_edges ? build_edges(c) : NPB_OK;
}

int npb_psup_stats(npb_ctx *c)
{
    return build_psup(c);
}

static int out_i64(npb_ctx *c, const i64 *dev, i64 n, void *out, i64 cap)
{
    if (cap < (i64)sizeof(i64) * n) {
        npb_set_error("output buffer too small: need %lld bytes, have %lld", (long long)(sizeof(i64) * n), (long long)cap);
        return NPB_ERR_ARG;
    }
    if (n > 0) NPB_CUDA(cudaMemcpyAsync(out, dev, sizeof(i64) * n, cudaMemcpyDeviceToHost, c->stream));
    NPB_CUDA(cudaStreamSynchronize(c->stream));
    return NPB_OK;
}

static int export_rows(npb_ctx *c, const int32_t *src, i64 rows, int ss, int ds, void *out, i64 cap)
{
    i64 n = rows * ds;
    NpbTmp t;
    NPB_CUDA(t.alloc(sizeof(i64) * (size_t)(n > 0 ? n : 1)));
    i64 *tmp = t.as<i64>();
    if (n > 0) {
        k_widen_rows<<<npb_blocks(n, 256), 256, 0, c->stream>>>(src, rows, ss, ds, tmp);
        NPB_LAUNCH(c);
    }
    return out_i64(c, tmp, n, out, cap);
}

__global__ void k_strip_pad(const double *__restrict__ in, i64 n, double *__restrict__ out)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * 3) return;
    i64 e = idx / 3;
    out[idx] = in[e * NPB_CSTRIDE + (idx - e * 3)];
}

static int export_f64(npb_ctx *c, const double *src, i64 n, void *out, i64 cap)
{
    if (cap < (i64)sizeof(double) * n) {
        npb_set_error("output buffer too small: need %lld bytes, have %lld", (long long)(sizeof(double) * n), (long long)cap);
        return NPB_ERR_ARG;
    }
    if (n > 0) NPB_CUDA(cudaMemcpyAsync(out, src, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    NPB_CUDA(cudaStreamSynchronize(c->stream));
    return NPB_OK;
}

int npb_export_array(npb_ctx *c, const char *name, void *out, i64 cap)
{
    i64 ne = c->n_elems, np = c->n_points, nf = c->n_faces;
    if (!strcmp(name, "point_coords")) return export_f64(c, c->coords, np * 3, out, cap);
    if (!strcmp(name, "centroids") && NPB_CSTRIDE == 3) return export_f64(c, c->centroids, ne * 3, out, cap);
    if (!strcmp(name, "centroids")) {   // device records are [n_elems, NPB_CSTRIDE]: strip the pad
        NpbTmp packed;
        NPB_CUDA(packed.alloc(sizeof(double) * (size_t)ne * 3));
        if (ne > 0) {
            k_strip_pad<<<npb_blocks(ne * 3, 256), 256, 0, c->stream>>>(c->centroids, ne, packed.as<double>());
            NPB_LAUNCH(c);
        }
        return export_f64(c, packed.as<double>(), ne * 3, out, cap);
    }
    if (!strcmp(name, "faces_centers")) return export_f64(c, c->fcent, nf * 3, out, cap);
    if (!strcmp(name, "normal_faces")) return export_f64(c, c->fnormal, nf * 3, out, cap);
    if (!strcmp(name, "faces_areas")) return export_f64(c, c->farea, nf, out, cap);
    if (!strcmp(name, "inpoel")) return export_rows(c, c->inpoel, ne, c->spe, NPB_MX_PE, out, cap);
    if (!strcmp(name, "esuel")) return export_rows(c, c->esuel, ne, c->sfe, NPB_MX_FE, out, cap);
    if (!strcmp(name, "infael")) return export_rows(c, c->infael, ne, c->sfe, NPB_MX_FE, out, cap);
    if (!strcmp(name, "inpofa")) return export_rows(c, c->inpofa, nf, NPB_MX_PF, NPB_MX_PF, out, cap);
    if (!strcmp(name, "esup")) return export_rows(c, c->esup, c->len_esup, 1, 1, out, cap);
    if (!strcmp(name, "esup_ptr")) return export_rows(c, c->esup_ptr, np + 1, 1, 1, out, cap);
    if (!strcmp(name, "fsup")) return export_rows(c, c->fsup, c->len_fsup, 1, 1, out, cap);
    if (!strcmp(name, "fsup_ptr")) return export_rows(c, c->fsup_ptr, np + 1, 1, 1, out, cap);
    if (!strcmp(name, "inedel") || !strcmp(name, "inpoed")) {
        if (!c->build_edges) {
            npb_set_error("edge structures were not requested (build_edges = 0)");
            return NPB_ERR_STATE;
        }
        NPB_TRY(build_edges(c));
        i64 n = !strcmp(name, "inedel") ? c->n_elems * NPB_MX_EE : c->n_edges * 2;
        return out_i64(c, !strcmp(name, "inedel") ? c->inedel_d : c->inpoed_d, n, out, cap);
    }
    if (!strcmp(name, "psup") || !strcmp(name, "psup_ptr")) {
        NPB_TRY(build_psup(c));
        if (!strcmp(name, "psup")) return export_rows(c, c->psup, c->len_psup, 1, 1, out, cap);
        return export_rows(c, c->psup_ptr, np + 1, 1, 1, out, cap);
    }
    if (!strcmp(name, "element_types") || !strcmp(name, "boundary_faces") || !strcmp(name, "boundary_points")) {
        const uint8_t *src = !strcmp(name, "element_types") ? c->etype : (!strcmp(name, "boundary_faces") ? c->bface : c->bpoint);
        i64 n = !strcmp(name, "element_types") ? ne : (!strcmp(name, "boundary_faces") ? nf : np);
        NpbTmp t;
        NPB_CUDA(t.alloc(sizeof(i64) * (size_t)(n > 0 ? n : 1)));
        i64 *tmp = t.as<i64>();
        if (n > 0) {
            k_widen_u8<<<npb_blocks(n, 256), 256, 0, c->stream>>>(src, n, tmp);
            NPB_LAUNCH(c);
        }
        return out_i64(c, tmp, n, out, cap);
    }
    if (!strcmp(name, "esuf") || !strcmp(name, "esuf_ptr")) {
        NpbTmp t_ptr, t_flat;
        NPB_CUDA(t_ptr.alloc(sizeof(int32_t) * (size_t)(nf + 1)));
        int32_t *ptr = t_ptr.as<int32_t>();
        k_esuf_count<<<npb_blocks(nf + 1, 256), 256, 0, c->stream>>>(c->esuf2, nf, ptr);
        NPB_LAUNCH(c);
        int rc = npb_exclusive_scan_i32(c, ptr, ptr, nf + 1);
        if (rc == NPB_OK) {
            if (!strcmp(name, "esuf_ptr")) {
                rc = export_rows(c, ptr, nf + 1, 1, 1, out, cap);
            } else {
                NPB_CUDA(t_flat.alloc(sizeof(i64) * (size_t)(c->len_esuf > 0 ? c->len_esuf : 1)));
                i64 *tmp = t_flat.as<i64>();
                if (nf > 0) {
                    k_esuf_flat<<<npb_blocks(nf, 256), 256, 0, c->stream>>>(c->esuf2, ptr, nf, tmp);
                    NPB_LAUNCH(c);
                }
                rc = out_i64(c, tmp, c->len_esuf, out, cap);
            }
        }
        return rc;
    }
    npb_set_error("npb_grid_array: unknown or unavailable array '%s'", name);
    return NPB_ERR_ARG;
}
