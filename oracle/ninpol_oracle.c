/*
 * ninpol_oracle.c — CPU restatement of the reference's nodal-interpolation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path; it is never linked
 * into, imported by or called from the product (`ninpol_b200`).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Every function is a serial, literal restatement of one reference routine (daviyan5/ninpol 1.0.2,
 * release build: cdivision=True, -O3 -fopenmp, x86-64 without FMA); the reference file:line each one
 * follows is cited at the function.  Arrays use the reference's own dtypes and layouts (int64 ids,
 * -1 padding, float64 geometry) so results compare with np.array_equal.
 *
 * Parity status: PINNED.  tests/test_oracle.py compares every output of this file with
 * the compiled reference (oracle/_ref, built by oracle/build_ref.py) on tet / hex / mixed meshes, and
 * tests/golden/ holds fixtures generated from that compiled reference (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/__init__.py does it).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef long long i64;

#define MX_PE 8   /* NINPOL_MAX_POINTS_PER_ELEMENT  (ninpol_defines.pxd:2) */
#define MX_FE 6   /* NINPOL_MAX_FACES_PER_ELEMENT   (ninpol_defines.pxd:3) */
#define MX_PF 4   /* NINPOL_MAX_POINTS_PER_FACE     (ninpol_defines.pxd:4) */
#define N_TYPES 8 /* NINPOL_NUM_ELEMENT_TYPES       (ninpol_defines.pxd:5) */

/* ------------------------------------------------------------------------------------------------
 * Grid.build_esup — grid.pyx:233-267.  esup_ptr has n_points+1 entries (zeroed here), esup has
 * sum(npoel[type]) entries.  Returns MX_ELEMENTS_PER_POINT.
 * ---------------------------------------------------------------------------------------------- */
i64 orc_build_esup(i64 n_elems, i64 n_points, const i64 *inpoel, const i64 *etype, const i64 *npoel,
                   i64 *esup_ptr, i64 *esup)
{
    i64 mx = 0;
    memset(esup_ptr, 0, sizeof(i64) * (size_t)(n_points + 1));
    for (i64 i = 0; i < n_elems; i++) {                       /* :244-250 */
        i64 t = etype[i];
        for (i64 j = 0; j < npoel[t]; j++) {
            i64 p = inpoel[i * MX_PE + j];
            esup_ptr[p + 1] += 1;
            if (esup_ptr[p + 1] > mx) mx = esup_ptr[p + 1];
        }
    }
    for (i64 i = 0; i < n_points; i++) esup_ptr[i + 1] += esup_ptr[i];   /* :253-254 */
    for (i64 i = 0; i < n_elems; i++) {                       /* :258-263 */
        i64 t = etype[i];
        for (i64 j = 0; j < npoel[t]; j++) {
            i64 p = inpoel[i * MX_PE + j];
            esup[esup_ptr[p]] = i;
            esup_ptr[p] += 1;
        }
    }
    for (i64 i = n_points; i > 0; i--) esup_ptr[i] = esup_ptr[i - 1];     /* :265-267 */
    esup_ptr[0] = 0;
    return mx;
}

/* ------------------------------------------------------------------------------------------------
 * Grid.build_psup — grid.pyx:269-302.  psup must hold esup_ptr[n_points]*(MX_PE-1) entries;
 * returns the used length; *mx_out = MX_POINTS_PER_POINT.
 * ---------------------------------------------------------------------------------------------- */
i64 orc_build_psup(i64 n_points, const i64 *inpoel, const i64 *etype, const i64 *npoel,
                   const i64 *esup_ptr, const i64 *esup, i64 *psup_ptr, i64 *psup, i64 *mx_out)
{
    i64 *tmp = (i64 *)malloc(sizeof(i64) * (size_t)n_points);
    i64 stor = 0, mx = 0;
    for (i64 i = 0; i < n_points; i++) tmp[i] = -1;
    psup_ptr[0] = 0;
    for (i64 i = 0; i < n_points; i++) {
        for (i64 j = esup_ptr[i]; j < esup_ptr[i + 1]; j++) {
            i64 e = esup[j];
            i64 t = etype[e];
            for (i64 k = 0; k < npoel[t]; k++) {
                i64 q = inpoel[e * MX_PE + k];
                if (q != i && tmp[q] != i) {
                    psup[stor] = q;
                    tmp[q] = i;
                    stor++;
                }
            }
        }
        psup_ptr[i + 1] = stor;
        if (psup_ptr[i + 1] - psup_ptr[i] > mx) mx = psup_ptr[i + 1] - psup_ptr[i];
    }
    free(tmp);
    *mx_out = mx;
    return stor;
}

/* ------------------------------------------------------------------------------------------------
 * Grid.build_esuel — grid.pyx:449-525 (the prange over ielem executed serially in ascending order;
 * on conforming meshes the result does not depend on the iteration order).  esuel is
 * [n_elems, MX_FE], filled with -1 here.
 * ---------------------------------------------------------------------------------------------- */
void orc_build_esuel(i64 n_elems, const i64 *inpoel, const i64 *etype, const i64 *nfael,
                     const i64 *lnofa, const i64 *lpofa, const i64 *esup_ptr, const i64 *esup, i64 *esuel)
{
    for (i64 i = 0; i < n_elems * MX_FE; i++) esuel[i] = -1;
    for (i64 ie = 0; ie < n_elems; ie++) {
        i64 it = etype[ie];
        for (i64 j = 0; j < nfael[it]; j++) {
            if (esuel[ie * MX_FE + j] != -1) continue;                               /* :476 */
            i64 nj = lnofa[it * MX_FE + j];
            i64 point = inpoel[ie * MX_PE + lpofa[(it * MX_FE + j) * MX_PF + 0]];    /* :479 */
            i64 nmin = esup_ptr[point + 1] - esup_ptr[point];
            for (i64 k = 0; k < nj; k++) {                                           /* :483-488 */
                i64 kp = inpoel[ie * MX_PE + lpofa[(it * MX_FE + j) * MX_PF + k]];
                i64 ne = esup_ptr[kp + 1] - esup_ptr[kp];
                if (ne < nmin) { point = kp; nmin = ne; }
            }
            int found = 0;
            for (i64 k = esup_ptr[point]; k < esup_ptr[point + 1]; k++) {            /* :494 */
                i64 je = esup[k];
                i64 jt = etype[je];
                if (je != ie) {
                    for (i64 l = 0; l < nfael[jt]; l++) {                            /* :502 */
                        i64 eq = 0;
                        for (i64 m = 0; m < lnofa[jt * MX_FE + l]; m++) {
                            i64 q = inpoel[je * MX_PE + lpofa[(jt * MX_FE + l) * MX_PF + m]];
                            for (i64 o = 0; o < nj; o++) {
                                if (q == inpoel[ie * MX_PE + lpofa[(it * MX_FE + j) * MX_PF + o]]) { eq++; break; }
                            }
                        }
                        if (eq == nj) {                                              /* :512-519 */
                            esuel[ie * MX_FE + j] = je;
                            esuel[je * MX_FE + l] = ie;
                            found = 1;
                        }
                        if (found) break;
                    }
                }
                if (found) break;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Grid.build_infael — grid.pyx:304-345.  infael [n_elems, MX_FE] (set to -1 here); face_to_elem is
 * caller scratch [n_elems*MX_FE, 2].  Returns n_faces.  orc_fill_inpofa is :337-345.
 * ---------------------------------------------------------------------------------------------- */
i64 orc_build_infael(i64 n_elems, const i64 *etype, const i64 *nfael, const i64 *esuel, i64 *infael,
                     i64 *face_to_elem)
{
    i64 face_index = 0;
    for (i64 i = 0; i < n_elems * MX_FE; i++) infael[i] = -1;
    for (i64 i = 0; i < n_elems; i++) {
        i64 it = etype[i];
        for (i64 j = 0; j < nfael[it]; j++) {
            if (infael[i * MX_FE + j] != -1) continue;
            infael[i * MX_FE + j] = face_index;
            face_index++;
            face_to_elem[infael[i * MX_FE + j] * 2 + 0] = i;
            face_to_elem[infael[i * MX_FE + j] * 2 + 1] = j;
            i64 k = esuel[i * MX_FE + j];
            if (k == -1) continue;
            i64 kt = etype[k];
            for (i64 l = 0; l < nfael[kt]; l++) {
                if (esuel[k * MX_FE + l] == i) { infael[k * MX_FE + l] = infael[i * MX_FE + j]; break; }
            }
        }
    }
    return face_index;
}

void orc_fill_inpofa(i64 n_faces, const i64 *inpoel, const i64 *etype, const i64 *lnofa, const i64 *lpofa,
                     const i64 *face_to_elem, i64 *inpofa)
{
    for (i64 i = 0; i < n_faces * MX_PF; i++) inpofa[i] = -1;
    for (i64 f = 0; f < n_faces; f++) {
        i64 i = face_to_elem[f * 2], j = face_to_elem[f * 2 + 1];
        i64 it = etype[i];
        for (i64 k = 0; k < lnofa[it * MX_FE + j]; k++)
            inpofa[f * MX_PF + k] = inpoel[i * MX_PE + lpofa[(it * MX_FE + j) * MX_PF + k]];
    }
}

/* ------------------------------------------------------------------------------------------------
 * Grid.build_fsup — grid.pyx:347-379.  Two calls: count (fsup == NULL) returns total length after
 * filling fsup_ptr with the prefix sums; fill writes fsup.  Returns MX_FACES_PER_POINT via *mx_out.
 * ---------------------------------------------------------------------------------------------- */
i64 orc_build_fsup(i64 n_faces, i64 n_points, const i64 *inpofa, i64 *fsup_ptr, i64 *fsup, i64 *mx_out)
{
    if (fsup == NULL) {
        i64 mx = 0;
        memset(fsup_ptr, 0, sizeof(i64) * (size_t)(n_points + 1));
        for (i64 i = 0; i < n_faces; i++)
            for (i64 j = 0; j < MX_PF; j++) {
                if (inpofa[i * MX_PF + j] == -1) break;
                i64 p = inpofa[i * MX_PF + j];
                fsup_ptr[p + 1] += 1;
                if (fsup_ptr[p + 1] > mx) mx = fsup_ptr[p + 1];
            }
        for (i64 i = 0; i < n_points; i++) fsup_ptr[i + 1] += fsup_ptr[i];
        *mx_out = mx;
        return fsup_ptr[n_points];
    }
    for (i64 i = 0; i < n_faces; i++)
        for (i64 j = 0; j < MX_PF; j++) {
            if (inpofa[i * MX_PF + j] == -1) break;
            i64 p = inpofa[i * MX_PF + j];
            fsup[fsup_ptr[p]] = i;
            fsup_ptr[p] += 1;
        }
    for (i64 i = n_points; i > 0; i--) fsup_ptr[i] = fsup_ptr[i - 1];
    fsup_ptr[0] = 0;
    return fsup_ptr[n_points];
}

/* ------------------------------------------------------------------------------------------------
 * Grid.build_esuf + boundary tagging — grid.pyx:381-444.  esuf has sum(nfael[type]) entries.
 * inpofa is re-derived from the first element of each face (:424-432), boundary_faces/points are
 * int64 0/1.  Returns MX_ELEMENTS_PER_FACE.
 * ---------------------------------------------------------------------------------------------- */
i64 orc_build_esuf(i64 n_elems, i64 n_faces, i64 n_points, const i64 *inpoel, const i64 *etype,
                   const i64 *nfael, const i64 *lnofa, const i64 *lpofa, const i64 *infael,
                   i64 *esuf_ptr, i64 *esuf, i64 *inpofa, i64 *boundary_faces, i64 *boundary_points)
{
    i64 mx = 0;
    memset(esuf_ptr, 0, sizeof(i64) * (size_t)(n_faces + 1));
    for (i64 i = 0; i < n_elems; i++) {
        i64 t = etype[i];
        for (i64 j = 0; j < nfael[t]; j++) {
            i64 f = infael[i * MX_FE + j];
            esuf_ptr[f + 1] += 1;
            if (esuf_ptr[f + 1] > mx) mx = esuf_ptr[f + 1];
        }
    }
    for (i64 i = 0; i < n_faces; i++) esuf_ptr[i + 1] += esuf_ptr[i];
    for (i64 i = 0; i < n_elems; i++) {
        i64 t = etype[i];
        for (i64 j = 0; j < nfael[t]; j++) {
            i64 f = infael[i * MX_FE + j];
            esuf[esuf_ptr[f]] = i;
            esuf_ptr[f] += 1;
        }
    }
    for (i64 i = n_faces; i > 0; i--) esuf_ptr[i] = esuf_ptr[i - 1];
    esuf_ptr[0] = 0;
    for (i64 f = 0; f < n_faces; f++) {                                   /* :424-432 */
        i64 e = esuf[esuf_ptr[f]];
        if (e != -1) {
            i64 t = etype[e];
            i64 j;
            for (j = 0; j < nfael[t]; j++)
                if (infael[e * MX_FE + j] == f) break;
            for (i64 k = 0; k < lnofa[t * MX_FE + j]; k++)
                inpofa[f * MX_PF + k] = inpoel[e * MX_PE + lpofa[(t * MX_FE + j) * MX_PF + k]];
        }
    }
    memset(boundary_faces, 0, sizeof(i64) * (size_t)n_faces);
    memset(boundary_points, 0, sizeof(i64) * (size_t)n_points);
    for (i64 f = 0; f < n_faces; f++) {                                   /* :438-444 */
        if (esuf_ptr[f + 1] - esuf_ptr[f] == 1) {
            boundary_faces[f] = 1;
            for (i64 j = 0; j < MX_PF; j++) {
                i64 p = inpofa[f * MX_PF + j];
                if (p == -1) break;
                boundary_points[p] = 1;
            }
        }
    }
    return mx;
}

/* ------------------------------------------------------------------------------------------------
 * Grid.calculate_centroids — grid.pyx:669-719.  centroids [n_elems,3] and faces_centers
 * [n_faces,3] must be zero-initialised by the caller (np.zeros in the reference).
 * ---------------------------------------------------------------------------------------------- */
void orc_centroids(i64 dim, i64 n_elems, i64 n_faces, const i64 *inpoel, const i64 *etype, const i64 *npoel,
                   const i64 *inpofa, const double *coords, double *centroids, double *faces_centers)
{
    for (i64 i = 0; i < n_elems; i++) {
        int npe = (int)npoel[etype[i]];
        for (int j = 0; j < npe; j++)
            for (i64 k = 0; k < dim; k++)
                centroids[i * 3 + k] += coords[inpoel[i * MX_PE + j] * 3 + k] / npe;      /* :704 */
    }
    for (i64 i = 0; i < n_faces; i++) {
        int npofa = 0;
        for (int j = 0; j < MX_PF; j++) {
            if (inpofa[i * MX_PF + j] == -1) break;
            npofa = npofa + 1;
            for (i64 k = 0; k < dim; k++) faces_centers[i * 3 + k] += coords[inpofa[i * MX_PF + j] * 3 + k];
        }
        for (i64 k = 0; k < dim; k++) faces_centers[i * 3 + k] /= npofa;                   /* :717 */
    }
}

/* ------------------------------------------------------------------------------------------------
 * Grid.calculate_normal_faces — grid.pyx:721-809.  The locals are C `float` in the reference
 * (:732-736): differences are taken in double and rounded to float, the cross product, norm and the
 * division are float.  normal_faces [n_faces,3], faces_areas [n_faces].
 * ---------------------------------------------------------------------------------------------- */
void orc_normals(i64 dim, i64 n_faces, const i64 *inpofa, const double *coords, double *normal_faces,
                 double *faces_areas)
{
    float v1x, v1y, v1z, v2x, v2y, v2z, normalx, normaly, normalz, norm;
    if (dim == 3) {
        for (i64 f = 0; f < n_faces; f++) {
            int npofa = inpofa[f * MX_PF + 3] == -1 ? 3 : 4;
            i64 p1 = inpofa[f * MX_PF + 0], p2 = inpofa[f * MX_PF + 1], p3 = inpofa[f * MX_PF + 2];
            v1x = (float)(coords[p1 * 3 + 0] - coords[p2 * 3 + 0]);
            v1y = (float)(coords[p1 * 3 + 1] - coords[p2 * 3 + 1]);
            v1z = (float)(coords[p1 * 3 + 2] - coords[p2 * 3 + 2]);
            v2x = (float)(coords[p3 * 3 + 0] - coords[p2 * 3 + 0]);
            v2y = (float)(coords[p3 * 3 + 1] - coords[p2 * 3 + 1]);
            v2z = (float)(coords[p3 * 3 + 2] - coords[p2 * 3 + 2]);
            normalx = v1y * v2z - v1z * v2y;
            normaly = v1z * v2x - v1x * v2z;
            normalz = v1x * v2y - v1y * v2x;
            norm = (float)sqrt((double)(normalx * normalx + normaly * normaly + normalz * normalz));
            norm = fabsf(norm);
            normal_faces[f * 3 + 0] = normalx / norm;
            normal_faces[f * 3 + 1] = normaly / norm;
            normal_faces[f * 3 + 2] = normalz / norm;
            if (npofa == 3) {
                faces_areas[f] = norm / 2.0;
            } else {
                i64 p4 = inpofa[f * MX_PF + 3];
                v1x = (float)(coords[p1 * 3 + 0] - coords[p4 * 3 + 0]);
                v1y = (float)(coords[p1 * 3 + 1] - coords[p4 * 3 + 1]);
                v1z = (float)(coords[p1 * 3 + 2] - coords[p4 * 3 + 2]);
                v2x = (float)(coords[p3 * 3 + 0] - coords[p4 * 3 + 0]);
                v2y = (float)(coords[p3 * 3 + 1] - coords[p4 * 3 + 1]);
                v2z = (float)(coords[p3 * 3 + 2] - coords[p4 * 3 + 2]);
                normalx = v1y * v2z - v1z * v2y;
                normaly = v1z * v2x - v1x * v2z;
                normalz = v1x * v2y - v1y * v2x;
                /* grid.pyx is compiled as C++ (setup.py:33-36): sqrt(float) resolves to the float
                 * overload, so the second norm and the sum are float; only the /2.0 is double. */
                float both = norm + sqrtf(normalx * normalx + normaly * normaly + normalz * normalz);
                faces_areas[f] = both / 2.0;
            }
        }
    } else {
        for (i64 f = 0; f < n_faces; f++) {                               /* :788-806 */
            i64 p1 = inpofa[f * MX_PF + 0], p2 = inpofa[f * MX_PF + 1];
            v1x = (float)(coords[p1 * 3 + 0] - coords[p2 * 3 + 0]);
            v1y = (float)(coords[p1 * 3 + 1] - coords[p2 * 3 + 1]);
            normalx = -v1y;
            normaly = v1x;
            norm = (float)sqrt((double)(normalx * normalx + normaly * normaly));
            norm = fabsf(norm);
            normal_faces[f * 3 + 0] = normalx / norm;
            normal_faces[f * 3 + 1] = normaly / norm;
            normal_faces[f * 3 + 2] = 0.0;
            faces_areas[f] = norm;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * IDWInterpolation.inverse_distance — idw.pyx:35-84, all nodes as targets.  weights is
 * [n_points, ncol] zero-initialised (interpolator.pyx:650).
 * ---------------------------------------------------------------------------------------------- */
/* `list` (n_list node ids) restricts the loop to a sample of target nodes — the reference's
 * `target_points` loop variable (idw.pyx:57-60) — and row i of `weights` then belongs to list[i];
 * list == NULL means every node, row = node id. */
void orc_idw_nodes(i64 dim, i64 n_list, const i64 *list, i64 ncol, const i64 *esup_ptr, const i64 *esup,
                   const i64 *boundary_points, const i64 *neumann_point, const double *coords, const double *centroids,
                   double *weights_out)
{
    const float machine_epsilon = (float)1e-15;                            /* :53 */
    for (i64 li = 0; li < n_list; li++) {
        const i64 point = list ? list[li] : li;
        double *weights = weights_out + (li - point) * ncol;               /* weights[point * ncol + k] is row li */
        int zero_found = 0;
        double total = 0.0;
        i64 n_source = 0;
        if (boundary_points[point] && !neumann_point[point]) continue;    /* :62 */
        i64 j = 0;
        for (i64 q = esup_ptr[point]; q < esup_ptr[point + 1]; q++, j++) {
            i64 src = esup[q];
            double distance = 0.0;
            for (i64 k = 0; k < dim; k++) {
                double d = coords[point * 3 + k] - centroids[src * 3 + k];
                distance = distance + d * d;                               /* (..)**2, :67 */
            }
            if (distance <= machine_epsilon) {                             /* :69-74 */
                zero_found = 1;
                for (i64 k = 0; k < n_source; k++) weights[point * ncol + k] = 0.;
                weights[point * ncol + j] = 1.;
                break;
            }
            distance = sqrt(distance);
            weights[point * ncol + j] += 1 / distance;
            total += 1 / distance;
            n_source = n_source + 1;
        }
        if (!zero_found)
            for (i64 k = 0; k < n_source; k++) weights[point * ncol + k] /= total;
    }
}

void orc_idw(i64 dim, i64 n_points, i64 ncol, const i64 *esup_ptr, const i64 *esup, const i64 *boundary_points,
             const i64 *neumann_point, const double *coords, const double *centroids, double *weights)
{
    orc_idw_nodes(dim, n_points, NULL, ncol, esup_ptr, esup, boundary_points, neumann_point, coords, centroids, weights);
}

/* ------------------------------------------------------------------------------------------------
 * LSInterpolation.LS — ls.pyx:33-135, all nodes as targets.
 * ---------------------------------------------------------------------------------------------- */
void orc_ls_nodes(i64 n_list, const i64 *list, i64 ncol, const i64 *esup_ptr, const i64 *esup,
                  const i64 *boundary_points, const i64 *neumann_point, const double *coords, const double *centroids,
                  double *weights_out)
{
    for (i64 li = 0; li < n_list; li++) {
        const i64 point = list ? list[li] : li;
        double *weights = weights_out + (li - point) * ncol;               /* see orc_idw_nodes */
        double Ix, Iy, Iz, Ixx, Ixy, Ixz, Iyy, Iyz, Izz, D, lx, ly, lz, denom, vx, vy, vz, total;
        if (boundary_points[point] && !neumann_point[point]) continue;
        Ix = Iy = Iz = 0.0;
        Ixx = Ixy = Ixz = Iyy = Iyz = Izz = 0.0;
        int n_vols = (int)(esup_ptr[point + 1] - esup_ptr[point]);
        for (i64 q = esup_ptr[point]; q < esup_ptr[point + 1]; q++) {      /* :64-77 */
            i64 vol = esup[q];
            vx = centroids[vol * 3 + 0] - coords[point * 3 + 0];
            vy = centroids[vol * 3 + 1] - coords[point * 3 + 1];
            vz = centroids[vol * 3 + 2] - coords[point * 3 + 2];
            Ix = Ix + vx;  Iy = Iy + vy;  Iz = Iz + vz;
            Ixx = Ixx + vx * vx;  Ixy = Ixy + vx * vy;  Ixz = Ixz + vx * vz;
            Iyy = Iyy + vy * vy;  Iyz = Iyz + vy * vz;  Izz = Izz + vz * vz;
        }
        if (Iz == 0.0 && Izz == 0.0 && Ixz == 0.0 && Iyz == 0.0) Izz = 1.0;   /* :79-80 */
        D = (Ixx * (Iyy * Izz - Iyz * Iyz) + Ixy * (Iyz * Ixz - Ixy * Izz) + Ixz * (Ixy * Iyz - Iyy * Ixz));
        if (D == 0.0) {                                                    /* :88-102 */
            total = 0.0;
            i64 i = 0;
            for (i64 q = esup_ptr[point]; q < esup_ptr[point + 1]; q++, i++) {
                i64 vol = esup[q];
                vx = centroids[vol * 3 + 0] - coords[point * 3 + 0];
                vy = centroids[vol * 3 + 1] - coords[point * 3 + 1];
                vz = centroids[vol * 3 + 2] - coords[point * 3 + 2];
                weights[point * ncol + i] = 1.0 / sqrt(vx * vx + vy * vy + vz * vz);
                total = total + 1.0 / sqrt(vx * vx + vy * vy + vz * vz);
            }
            for (i = 0; i < n_vols; i++) weights[point * ncol + i] = weights[point * ncol + i] / total;
            continue;
        }
        if (Iz == 0.0 && Izz == 0.0 && Ixz == 0.0 && Iyz == 0.0) Izz = -1.0;  /* :105-106 */
        lx = (Ix * (Iyz * Iyz - Iyy * Izz) + Iy * (Ixy * Izz - Iyz * Ixz) + Iz * (Iyy * Ixz - Ixy * Iyz)) / D;
        ly = (Ix * (Ixy * Izz - Iyz * Ixz) + Iy * (Ixz * Ixz - Ixx * Izz) + Iz * (Ixx * Iyz - Ixy * Ixz)) / D;
        lz = (Ix * (Iyy * Ixz - Ixy * Iyz) + Iy * (Ixx * Iyz - Ixy * Ixz) + Iz * (Ixy * Ixy - Ixx * Iyy)) / D;
        denom = n_vols + lx * Ix + ly * Iy + lz * Iz;                      /* :126 */
        i64 i = 0;
        for (i64 q = esup_ptr[point]; q < esup_ptr[point + 1]; q++, i++) {
            i64 vol = esup[q];
            vx = centroids[vol * 3 + 0] - coords[point * 3 + 0];
            vy = centroids[vol * 3 + 1] - coords[point * 3 + 1];
            vz = centroids[vol * 3 + 2] - coords[point * 3 + 2];
            weights[point * ncol + i] = (1. + lx * vx + ly * vy + lz * vz);
            weights[point * ncol + i] /= denom;
        }
    }
}

void orc_ls(i64 n_points, i64 ncol, const i64 *esup_ptr, const i64 *esup, const i64 *boundary_points,
            const i64 *neumann_point, const double *coords, const double *centroids, double *weights)
{
    orc_ls_nodes(n_points, NULL, ncol, esup_ptr, esup, boundary_points, neumann_point, coords, centroids, weights);
}

/* ------------------------------------------------------------------------------------------------
 * GLSInterpolation.GLS + build_ks_sv_arrays + build_ls_matrices + set_neumann_rows + solve_ls —
 * gls.pyx:75-474, all nodes as targets, one scratch slab (the reference has one per OpenMP thread).
 * dgels / dgemv are the Fortran-ABI entry points of scipy.linalg.cython_lapack / cython_blas (the very
 * functions the reference calls at gls.pyx:156,320,321,397,457); the Python wrapper extracts them
 * from scipy's capsules.  MXE / MXF are MX_ELEMENTS_PER_POINT / MX_FACES_PER_POINT.
 * Optionally dumps, for one node (`dump_point` >= 0), the dense system Mi (m x n, row-major with
 * row stride n) into dump_M and its sizes into dump_mn — used by tests to check conditioning.
 * Returns 0, or -1 on allocation failure.
 * ---------------------------------------------------------------------------------------------- */
typedef void (*dgels_t)(char *, int *, int *, int *, double *, int *, double *, int *, double *, int *, int *);
typedef void (*dgemv_t)(char *, int *, int *, double *, double *, int *, double *, int *, double *, double *, int *);

static void cross3(const double *a, const double *b, double *c)          /* gls.pyx:365-369 */
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

int orc_gls_nodes(i64 n_list, const i64 *list, i64 MXE, i64 MXF, const i64 *esup_ptr, const i64 *esup, const i64 *fsup_ptr,
            const i64 *fsup, const i64 *esuf_ptr, const i64 *esuf, const i64 *inpofa, const i64 *boundary_faces,
            const i64 *boundary_points, const i64 *neumann_point, const double *neumann_val,
            const double *coords, const double *centroids, const double *faces_centers,
            const double *normal_faces, const double *permeability /*[n_elems,3,3]*/, const double *diff_mag,
            dgels_t dgels, dgemv_t dgemv, double *weights_out /*[n_list,MXE]*/, double *neumann_out,
            i64 dump_point, double *dump_M, i64 *dump_mn)
{
    int M_MAX = (int)(MXE + 3 * MXF + MXF), N_MAX = (int)(3 * MXE + 1), NRHS_MAX = (int)(MXE + 1);
    double *Mi = (double *)malloc(sizeof(double) * (size_t)M_MAX * N_MAX);
    double *Ni = (double *)malloc(sizeof(double) * (size_t)M_MAX * NRHS_MAX);
    i64 *KSetv = (i64 *)malloc(sizeof(i64) * (size_t)MXE);
    i64 *Sv = (i64 *)malloc(sizeof(i64) * (size_t)MXF), *Svb = (i64 *)malloc(sizeof(i64) * (size_t)MXF);
    double *dKv = (double *)malloc(sizeof(double) * 3 * (size_t)MXE);
    double *T1 = (double *)malloc(sizeof(double) * 3 * (size_t)MXF), *tT2 = (double *)malloc(sizeof(double) * 3 * (size_t)MXF);
    double *nL1 = (double *)malloc(sizeof(double) * 3 * (size_t)MXF), *nL2 = (double *)malloc(sizeof(double) * 3 * (size_t)MXF);
    double *nL = (double *)malloc(sizeof(double) * 3 * (size_t)MXF);
    i64 *KsSv = (i64 *)malloc(sizeof(i64) * 2 * (size_t)MXF), *KsSvb = (i64 *)malloc(sizeof(i64) * (size_t)MXF);
    double *A = NULL, *B = NULL, *work = NULL;
    if (!Mi || !Ni || !KSetv || !Sv || !Svb || !dKv || !T1 || !tT2 || !nL1 || !nL2 || !nL || !KsSv || !KsSvb) return -1;
    /* workspace query, gls.pyx:156-158 */
    int lwork = -1, info = 0;
    {
        int m = M_MAX, n = N_MAX, nrhs = NRHS_MAX, lda = m > 1 ? m : 1, ldb = lda;
        double wq = 0.0;
        char tr = 'N';
        dgels(&tr, &m, &n, &nrhs, Mi, &lda, Ni, &ldb, &wq, &lwork, &info);
        lwork = (int)wq;
        if (lwork < 1) lwork = 1;
        work = (double *)malloc(sizeof(double) * (size_t)lwork);
        A = (double *)malloc(sizeof(double) * (size_t)M_MAX * N_MAX);
        B = (double *)malloc(sizeof(double) * (size_t)M_MAX * NRHS_MAX);
        if (!work || !A || !B) return -1;
    }
    for (i64 li = 0; li < n_list; li++) {
        const i64 point = list ? list[li] : li;                               /* row li of the outputs, see orc_idw_nodes */
        double *weights = weights_out + (li - point) * MXE;
        double *neumann_ws = neumann_out + (li - point);
        if (boundary_points[point] && !neumann_point[point]) continue;        /* :165 */
        int n_elem = (int)(esup_ptr[point + 1] - esup_ptr[point]);
        int n_face = (int)(fsup_ptr[point + 1] - fsup_ptr[point]);
        int n_bface = 0;
        for (i64 i = fsup_ptr[point]; i < fsup_ptr[point + 1]; i++)
            if (boundary_faces[fsup[i]] == 1) n_bface++;
        int m = n_elem + 3 * n_face + n_bface;                                 /* :177-181 */
        int n = 3 * n_elem + 1;
        int is_neu = (int)neumann_point[point];
        int nrhs = n_elem + is_neu;
        int lda = m > 1 ? m : 1, ldb = lda;
        memset(Mi, 0, sizeof(double) * (size_t)M_MAX * N_MAX);                /* :184-190 */
        memset(Ni, 0, sizeof(double) * (size_t)M_MAX * NRHS_MAX);
        /* build_ks_sv_arrays, :234-249 */
        for (i64 i = esup_ptr[point]; i < esup_ptr[point + 1]; i++) KSetv[i - esup_ptr[point]] = esup[i];
        {
            int j = 0;
            for (i64 i = fsup_ptr[point]; i < fsup_ptr[point + 1]; i++) {
                i64 face = fsup[i];
                Sv[i - fsup_ptr[point]] = face;
                if (boundary_faces[face] == 1) { Svb[j] = face; j++; }
            }
        }
        /* build_ls_matrices, :252-356 */
        if (!(n_bface >= n_face)) {                                            /* :266-267 */
            const double *xv = &coords[point * 3];
            for (int i = 0; i < n_elem; i++) {
                const double *xK = &centroids[KSetv[i] * 3];
                dKv[i * 3 + 0] = xK[0] - xv[0];
                dKv[i * 3 + 1] = xK[1] - xv[1];
                dKv[i * 3 + 2] = xK[2] - xv[2];
            }
            for (int i = 0; i < n_elem; i++) {                                 /* :275-281 */
                Mi[i * N_MAX + 3 * i] = dKv[i * 3 + 0];
                Mi[i * N_MAX + 3 * i + 1] = dKv[i * 3 + 1];
                Mi[i * N_MAX + 3 * i + 2] = dKv[i * 3 + 2];
                Mi[i * N_MAX + 3 * n_elem] = 1.0;
                Ni[i * NRHS_MAX + i] = 1.0;
            }
            int n_iface = n_face - n_bface, j = 0;
            int three = 3, one = 1;
            double alpha = 1.0, beta = 0.0;
            char tr = 'T';
            for (int i = 0; i < n_face; i++) {                                 /* :291-323 */
                i64 S = Sv[i];
                int n_esuf = (int)(esuf_ptr[S + 1] - esuf_ptr[S]);
                if (n_esuf < 2) continue;
                const double *xS = &faces_centers[S * 3];
                double Nsj[3] = {normal_faces[S * 3], normal_faces[S * 3 + 1], normal_faces[S * 3 + 2]};
                double eta = 0.0;
                for (int k = 0; k < n_esuf; k++) {
                    KsSv[j * 2 + k] = esuf[esuf_ptr[S] + k];
                    double dm = diff_mag[KsSv[j * 2 + k]];
                    eta = (dm > eta) ? dm : eta;        /* max(eta_j[j], diff_mag[..]), :304 */
                }
                T1[j * 3 + 0] = xv[0] - xS[0];
                T1[j * 3 + 1] = xv[1] - xS[1];
                T1[j * 3 + 2] = xv[2] - xS[2];
                double T2[3];
                cross3(Nsj, &T1[j * 3], T2);
                double tau = pow(sqrt(T2[0] * T2[0] + T2[1] * T2[1] + T2[2] * T2[2]), -eta);   /* :314,372 */
                tT2[j * 3 + 0] = tau * T2[0];
                tT2[j * 3 + 1] = tau * T2[1];
                tT2[j * 3 + 2] = tau * T2[2];
                dgemv(&tr, &three, &three, &alpha, (double *)&permeability[KsSv[j * 2 + 0] * 9], &three, Nsj, &one, &beta, &nL1[j * 3], &one);
                dgemv(&tr, &three, &three, &alpha, (double *)&permeability[KsSv[j * 2 + 1] * 9], &three, Nsj, &one, &beta, &nL2[j * 3], &one);
                j++;
            }
            int start = n_elem;
            for (int i = 0; i < n_iface; i++) {                                /* :328-356 */
                int I1 = -1, I2 = -1;
                for (int e = 0; e < n_elem; e++) {
                    if (KSetv[e] == KsSv[i * 2 + 0]) I1 = e;
                    if (KSetv[e] == KsSv[i * 2 + 1]) I2 = e;
                }
                int r1 = start, r2 = start + 1, r3 = start + 2;
                start += 3;
                for (int c = 0; c < 3; c++) {
                    Mi[r1 * N_MAX + 3 * I1 + c] = nL1[i * 3 + c] * -1;
                    Mi[r1 * N_MAX + 3 * I2 + c] = nL2[i * 3 + c] * 1;
                    Mi[r2 * N_MAX + 3 * I1 + c] = T1[i * 3 + c] * -1;
                    Mi[r2 * N_MAX + 3 * I2 + c] = T1[i * 3 + c] * 1;
                    Mi[r3 * N_MAX + 3 * I1 + c] = tT2[i * 3 + c] * -1;
                    Mi[r3 * N_MAX + 3 * I2 + c] = tT2[i * 3 + c] * 1;
                }
            }
        }
        if (is_neu) {                                                          /* set_neumann_rows :374-416 */
            int start = n_elem + 3 * n_face;
            int three = 3, one = 1;
            double alpha = 1.0, beta = 0.0;
            char tr = 'T';
            for (int i = 0; i < n_bface; i++) {
                int row = start + i;
                KsSvb[i] = esuf[esuf_ptr[Svb[i]]];
                dgemv(&tr, &three, &three, &alpha, (double *)&permeability[KsSvb[i] * 9], &three,
                      (double *)&normal_faces[Svb[i] * 3], &one, &beta, &nL[i * 3], &one);
                int total_b = 0;
                for (int q = 0; q < MX_PF; q++) {
                    i64 bp = inpofa[Svb[i] * MX_PF + q];
                    if (bp == -1) break;
                    total_b += 1;
                    Ni[row * NRHS_MAX + n_elem] += neumann_val[bp];
                }
                Ni[row * NRHS_MAX + n_elem] /= total_b;
            }
            for (int i = 0; i < n_bface; i++) {
                int row = start + i, Ik = -1;
                for (int e = 0; e < n_elem; e++)
                    if (KSetv[e] == KsSvb[i]) Ik = e;
                Mi[row * N_MAX + 3 * Ik] = -nL[i * 3 + 0];
                Mi[row * N_MAX + 3 * Ik + 1] = -nL[i * 3 + 1];
                Mi[row * N_MAX + 3 * Ik + 2] = -nL[i * 3 + 2];
            }
        }
        if (point == dump_point && dump_M) {
            for (int r = 0; r < m; r++)
                for (int c = 0; c < n; c++) dump_M[(size_t)r * n + c] = Mi[r * N_MAX + c];
            dump_mn[0] = m;
            dump_mn[1] = n;
        }
        /* solve_ls, :420-474 */
        for (int col = 0; col < n; col++)
            for (int row = 0; row < m; row++) A[row + (size_t)col * lda] = Mi[row * N_MAX + col];
        for (int col = 0; col < nrhs; col++)
            for (int row = 0; row < m; row++) B[row + (size_t)col * ldb] = Ni[row * NRHS_MAX + col];
        {
            char trn = 'N';
            info = 0;
            dgels(&trn, &m, &n, &nrhs, A, &lda, B, &ldb, work, &lwork, &info);
        }
        int w_total = nrhs - is_neu;
        for (int i = 0; i < w_total; i++) {
            weights[point * MXE + i] = 0;
            weights[point * MXE + i] += B[(n - 1) + (size_t)i * ldb];
        }
        if (is_neu) {
            neumann_ws[point] = 0;
            neumann_ws[point] += B[(n - 1) + (size_t)(w_total - 1) * ldb];     /* :470-472, off-by-one kept */
        }
    }
    free(Mi); free(Ni); free(KSetv); free(Sv); free(Svb); free(dKv); free(T1); free(tT2);
    free(nL1); free(nL2); free(nL); free(KsSv); free(KsSvb); free(A); free(B); free(work);
    return 0;
}

int orc_gls(i64 n_points, i64 MXE, i64 MXF, const i64 *esup_ptr, const i64 *esup, const i64 *fsup_ptr,
            const i64 *fsup, const i64 *esuf_ptr, const i64 *esuf, const i64 *inpofa, const i64 *boundary_faces,
            const i64 *boundary_points, const i64 *neumann_point, const double *neumann_val,
            const double *coords, const double *centroids, const double *faces_centers,
            const double *normal_faces, const double *permeability, const double *diff_mag,
            dgels_t dgels, dgemv_t dgemv, double *weights /*[n_points,MXE]*/, double *neumann_ws,
            i64 dump_point, double *dump_M, i64 *dump_mn)
{
    return orc_gls_nodes(n_points, NULL, MXE, MXF, esup_ptr, esup, fsup_ptr, fsup, esuf_ptr, esuf, inpofa, boundary_faces,
                         boundary_points, neumann_point, neumann_val, coords, centroids, faces_centers, normal_faces,
                         permeability, diff_mag, dgels, dgemv, weights, neumann_ws, dump_point, dump_M, dump_mn);
}

/* ------------------------------------------------------------------------------------------------
 * Grid.build_inedel — grid.pyx:527-580 with myhash (:29-43).  The reference keys its hash map by the
 * hash VALUE of the sorted end points, truncated to a C int (unordered_map[int, int], :539): two
 * different edges whose 32-bit hashes collide are merged into one edge id.  Reproduced as is.
 * inedel [n_elems, 12] and inpoed [n_elems*12, 2] are set to -1 here; returns n_edges.
 * ---------------------------------------------------------------------------------------------- */
static size_t ref_myhash2(i64 a, i64 b)
{
    size_t seed = 2;                                    /* len(vec) */
    i64 v[2] = {a, b};
    for (int i = 0; i < 2; i++) {
        int x = (int)v[i];
        x = (int)((unsigned int)((x >> 16) ^ x) * 0x45d9f3bu);
        x = (int)((unsigned int)((x >> 16) ^ x) * 0x45d9f3bu);
        x = (x >> 16) ^ x;
        seed ^= (unsigned int)((unsigned int)x + 0x9e3779b9u) + (seed << 6) + (seed >> 2);
    }
    return seed;
}

i64 orc_build_inedel(i64 n_elems, const i64 *inpoel, const i64 *etype, const i64 *nedel, const i64 *lpoed,
                     i64 *inedel, i64 *inpoed)
{
    const int MXE = 12;
    size_t cap = 64;
    while (cap < (size_t)n_elems * MXE * 2) cap <<= 1;
    int *keys = (int *)malloc(sizeof(int) * cap);
    int *vals = (int *)malloc(sizeof(int) * cap);
    unsigned char *used = (unsigned char *)calloc(cap, 1);
    i64 n_edges = 0;
    for (i64 i = 0; i < n_elems * MXE; i++) inedel[i] = -1;
    for (i64 i = 0; i < n_elems * MXE * 2; i++) inpoed[i] = -1;
    for (i64 i = 0; i < n_elems; i++) {
        i64 t = etype[i];
        for (i64 j = 0; j < nedel[t]; j++) {
            i64 e0 = inpoel[i * MX_PE + lpoed[(t * MXE + j) * 2 + 0]];
            i64 e1 = inpoel[i * MX_PE + lpoed[(t * MXE + j) * 2 + 1]];
            i64 s0 = e0, s1 = e1;
            if (e0 > e1) { s0 = e1; s1 = e0; }
            int key = (int)ref_myhash2(s0, s1);        /* size_t -> int key of the map */
            size_t h = ((size_t)(unsigned int)key * 0x9E3779B97F4A7C15ull) & (cap - 1);
            while (used[h] && keys[h] != key) h = (h + 1) & (cap - 1);
            i64 idx;
            if (!used[h]) {
                used[h] = 1;
                keys[h] = key;
                vals[h] = (int)n_edges;
                idx = n_edges;
                inpoed[idx * 2 + 0] = e0;
                inpoed[idx * 2 + 1] = e1;
                n_edges++;
            } else {
                idx = vals[h];
            }
            inedel[i * MXE + j] = idx;
        }
    }
    free(keys); free(vals); free(used);
    return n_edges;
}
