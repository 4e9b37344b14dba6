/* full_trace.c -- the C ABI of libort_b200.so used from plain C (what a ccall/cgo/JNI binding does).
 * Cooke triplet (reference test fixture, test/runtests.jl:19-35), one field, 64 x 32 half-pupil grid:
 * the hot loop of full_trace (src/PupilSampling.jl:115-146) in one call.
 *   gcc -std=c99 -I include examples/full_trace.c -o full_trace -L opticalraytracing.jl_b200/lib -lort_b200 -lm
 * The aimed inputs (y1, y2, y_EP, u, h', focus) come from the host prelude; they are passed on the command line. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ort_b200.h"

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: %s y1 y2 y_EP u h_prime focus\n", argv[0]); return 2; }
    const double y1 = atof(argv[1]), y2 = atof(argv[2]), y_EP = atof(argv[3]), u = atof(argv[4]);
    const double h_prime = atof(argv[5]), focus = atof(argv[6]);
    /* extended surfaces: image plane appended, t[end-1] = focus (src/PupilSampling.jl:111-114) */
    double R[9] = {INFINITY, 37.40, -341.48, -42.65, 36.40, INFINITY, 204.52, -37.05, INFINITY};
    double t[9] = {0.0, 5.90, 12.93, 2.50, 2.00, 9.85, 5.90, 0.0, 0.0};
    double n[9] = {1.0, 1.61272, 1.0, 1.64769, 1.0, 1.0, 1.61272, 1.0, 1.0};
    t[7] = focus;
    enum { NY = 64, NX = 32, NN = NY * NX };
    double ys[NY], xs[NX];
    /* collect(range(y1, y2, k)) / collect(range(0, y_EP, k/2)), evaluated like numpy.linspace: start + i*step, end point exact */
    const double sy = (y2 - y1) / (NY - 1), sx = y_EP / (NX - 1);
    for (int i = 0; i < NY; i++) ys[i] = y1 + i * sy;
    for (int i = 0; i < NX; i++) xs[i] = 0.0 + i * sx;
    ys[NY - 1] = y2; xs[NX - 1] = y_EP;

    ort_ctx *ctx = NULL;
    if (ort_init(&ctx, 0) != ORT_OK) { fprintf(stderr, "ort_init: %s\n", ort_last_error(NULL)); return 1; }
    if (ort_set_layout(ctx, 9, R, t, n, NULL) != ORT_OK) { fprintf(stderr, "%s\n", ort_last_error(ctx)); return 1; }
    static double ex[NN], ey[NN];
    static uint8_t mask[NN];
    ort_stats stats;
    ort_field field = {0};
    field.mode = 0; field.u = u; field.v = 0.0; field.h_prime = h_prime;
    ort_opts opts = {0};
    opts.arith = ORT_ARITH_FAST; opts.compact = 1;
    ort_grid_out out = {0};
    out.ex = ex; out.ey = ey; out.mask = mask; out.stats = &stats;
    if (ort_trace3d_grid(ctx, &field, 1, ys, NY, xs, NX, 5, 10.3, &opts, &out) != ORT_OK) {
        fprintf(stderr, "ort_trace3d_grid: %s\n", ort_last_error(ctx));
        return 1;
    }
    printf("n_kept=%lld rms=%.15g mean_y=%.15g ex0=%.17g ey0=%.17g launches=%lld\n", (long long)stats.n_kept,
           ort_rms_from_stats(&stats), stats.mean_y, ex[1], ey[1], (long long)ort_launch_count(ctx));
    ort_free(ctx);
    return 0;
}
