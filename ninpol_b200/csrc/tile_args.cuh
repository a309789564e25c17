// tile_args.cuh — argument block shared by the IDW / LS tile kernels (k2_idw_ls_tile.cu, k2_tile_pipe.cu).
#pragma once
#include "common.cuh"

struct TileArgs {
    const int32_t *esup_ptr, *esup;
    const uint8_t *bpoint, *nflag;
    const double *coords, *cent;
    const int32_t *indptr;
    int32_t *indices;
    double *data;
    int *zero_counter;
    double *neumann;
    double *wbuf;        // two-pass mode: values go to wbuf (esup-indexed from wbase) and rowcnt[p] = surviving entries
    int32_t *rowcnt;
    i64 wbase;
    i64 p_lo, p_hi;      // node range of this launch
    int nb, dim;
    int direct;          // 1: write the CSR at indptr[] positions and count exact zeros; 0: two-pass mode
};

// k2_tile_pipe.cu: the software-pipelined (TMA bulk + cp.async) variant; *used = 0 -> take the plain tile kernel
int npb_tile_pipe_launch(npb_ctx *c, const TileArgs &a, int method, int *used);
