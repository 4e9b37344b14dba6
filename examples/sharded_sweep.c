/* sharded_sweep.c -- the multi-GPU side of the C ABI from plain C: the hot loop of full_trace (src/PupilSampling.jl:115-146)
 * with the y-rows of the pupil grid block-sharded over several B200s and the per-field statistics combined INSIDE
 * libort_b200.so (ncclAllGather + rank-order merge kernel, ort_opts.gather_stats).  Two forms:
 *
 *   one process per GPU:   sharded_sweep rank <r> <world> <idfile> y1 y2 y_EP u h_prime focus
 *        rank 0 creates the communicator id and writes it to <idfile>; the other ranks wait for the file.
 *        Every rank traces its block of rows and prints the MERGED statistics (identical on all ranks).
 *   one process, n GPUs:   sharded_sweep multi <n> y1 y2 y_EP u h_prime focus
 *        ort_comm_init_all + ort_trace3d_grid_multi: one call, whole grid, outputs in the reference's order.
 *
 *   gcc -std=c99 -I include examples/sharded_sweep.c -o sharded_sweep -L opticalraytracing.jl_b200/lib -lort_b200 -lm
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ort_b200.h"

enum { NY = 64, NX = 32, NN = NY * NX };

static void die(const char *what, ort_ctx *ctx) { fprintf(stderr, "%s: %s\n", what, ort_last_error(ctx)); exit(1); }

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: see the header comment\n"); return 2; }
    const int multi = strcmp(argv[1], "multi") == 0;
    const int base = multi ? 3 : 5;
    if (argc < base + 6) { fprintf(stderr, "usage: see the header comment\n"); return 2; }
    const double y1 = atof(argv[base]), y2 = atof(argv[base + 1]), y_EP = atof(argv[base + 2]), u = atof(argv[base + 3]);
    const double h_prime = atof(argv[base + 4]), focus = atof(argv[base + 5]);
    /* Cooke triplet (test/runtests.jl:19-35), image plane appended, t[end-1] = focus (src/PupilSampling.jl:111-114) */
    double R[9] = {INFINITY, 37.40, -341.48, -42.65, 36.40, INFINITY, 204.52, -37.05, INFINITY};
    double t[9] = {0.0, 5.90, 12.93, 2.50, 2.00, 9.85, 5.90, 0.0, 0.0};
    double n[9] = {1.0, 1.61272, 1.0, 1.64769, 1.0, 1.0, 1.61272, 1.0, 1.0};
    t[7] = focus;
    static double ys[NY], xs[NX], ex[NN], ey[NN];
    static uint8_t mask[NN];
    const double sy = (y2 - y1) / (NY - 1), sx = y_EP / (NX - 1);
    for (int i = 0; i < NY; i++) ys[i] = y1 + i * sy;
    for (int i = 0; i < NX; i++) xs[i] = 0.0 + i * sx;
    ys[NY - 1] = y2; xs[NX - 1] = y_EP;
    ort_field field; memset(&field, 0, sizeof field);
    field.mode = 0; field.u = u; field.v = 0.0; field.h_prime = h_prime;
    ort_opts opts; memset(&opts, 0, sizeof opts);
    opts.arith = ORT_ARITH_FAST; opts.compact = 1;
    ort_grid_out out; memset(&out, 0, sizeof out);
    out.ex = ex; out.ey = ey; out.mask = mask;

    if (multi) {
        const int ng = atoi(argv[2]);
        ort_ctx *ctxs[ORT_MAX_GPUS];
        if (ng < 1 || ng > ORT_MAX_GPUS) return 2;
        for (int d = 0; d < ng; d++)
            if (ort_init(&ctxs[d], d) != ORT_OK) die("ort_init", NULL);
        if (ng > 1 && ort_comm_init_all(ctxs, ng) != ORT_OK) die("ort_comm_init_all", ctxs[0]);
        if (ort_set_layout(ctxs[0], 9, R, t, n, NULL) != ORT_OK) die("ort_set_layout", ctxs[0]);
        ort_stats merged, local[ORT_MAX_GPUS];
        out.stats = &merged; out.stats_local = local;
        if (ort_trace3d_grid_multi(ctxs, ng, &field, 1, ys, NY, xs, NX, 5, 10.3, &opts, &out) != ORT_OK)
            die("ort_trace3d_grid_multi", ctxs[0]);
        long long sum = 0;
        for (int d = 0; d < ng; d++) sum += (long long)local[d].n_kept;
        printf("multi n=%d n_kept=%lld sum_local=%lld rms=%.17g mean_y=%.17g ex1=%.17g ey1=%.17g exlast=%.17g\n", ng,
               (long long)merged.n_kept, sum, ort_rms_from_stats(&merged), merged.mean_y, ex[1], ey[1], ex[merged.n_kept - 1]);
        for (int d = 0; d < ng; d++) ort_free(ctxs[d]);
        return 0;
    }

    const int rank = atoi(argv[2]), world = atoi(argv[3]);
    const char *idfile = argv[4];
    unsigned char id[ORT_COMM_ID_BYTES];
    if (rank == 0) {
        if (ort_comm_unique_id(id) != ORT_OK) die("ort_comm_unique_id", NULL);
        char tmp[4096];
        snprintf(tmp, sizeof tmp, "%s.tmp", idfile);
        FILE *f = fopen(tmp, "wb");
        if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) { perror("idfile"); return 1; }
        fclose(f);
        rename(tmp, idfile);
    } else {
        FILE *f = NULL;
        for (int i = 0; i < 600 && !(f = fopen(idfile, "rb")); i++) { struct timespec ts = {0, 100000000}; nanosleep(&ts, NULL); }
        if (!f || fread(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "rank %d: no communicator id in %s\n", rank, idfile); return 1; }
        fclose(f);
    }
    ort_ctx *ctx = NULL;
    if (ort_init(&ctx, rank) != ORT_OK) die("ort_init", NULL);
    if (ort_comm_init_rank(ctx, id, rank, world) != ORT_OK) die("ort_comm_init_rank", ctx);
    if (ort_set_layout(ctx, 9, R, t, n, NULL) != ORT_OK) die("ort_set_layout", ctx);
    int64_t lo, hi;
    ort_comm_range(NY, rank, world, &lo, &hi);
    ort_stats merged, mine;
    out.stats = &merged; out.stats_local = &mine;
    opts.gather_stats = 1;
    /* this rank's rows: ys[lo .. hi); its compacted segment has mine.n_kept entries */
    if (ort_trace3d_grid(ctx, &field, 1, ys + lo, (int)(hi - lo), xs, NX, 5, 10.3, &opts, &out) != ORT_OK) die("ort_trace3d_grid", ctx);
    int nccl = 0, r = 0, w = 0;
    ort_comm_info(ctx, &r, &w, &nccl);
    printf("rank=%d/%d nccl=%d rows=[%lld,%lld) n_local=%lld n_kept=%lld rms=%.17g mean_y=%.17g\n", r, w, nccl, (long long)lo,
           (long long)hi, (long long)mine.n_kept, (long long)merged.n_kept, ort_rms_from_stats(&merged), merged.mean_y);
    ort_free(ctx);
    return 0;
}
