/* sphere_root_forms.c -- error study for the three algebraically equal forms of the ray-sphere path parameter in the
 * FAST (K-form) tracer of csrc/ort_internal.cuh:
 *
 *     stable     s = F / (G + sqrt(disc))            one division per sphere (4 DFMA/DMUL + 1 MUFU after the seed)
 *     cheap      s = (G - sqrt(disc)) * (1 / (c n1^2))   one subtraction + one multiplication by a per-surface constant
 *     centre     the same root written about the sphere's centre C = (0, 0, R):  b = (P - C).K,  q = |P - C|^2 - R^2,
 *                s = (-b - sgn sqrt(b^2 - n1^2 q)) / n1^2   -- 3 FP64 operations fewer again, q cancels like eps R^2
 *
 * The cheap form cancels: its absolute error is ~ eps * |R| per surface, independent of the gap.  This program traces
 * the same rays through the double-Gauss of BASELINE config 2 in double (both forms, fma as the GPU contracts) and in
 * 80-bit long double (stable form = truth), and prints the worst image-plane error of each form relative to the
 * position scale, for the nominal radii and for the same lens with one weak surface of growing radius.
 *
 *   gcc -O2 -ffp-contract=off -o srf tools/sphere_root_forms.c -lm && ./srf [rays per study, default 20000]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define ROWS 13
static double Rr[ROWS] = {INFINITY, 54.153, 152.522, 35.951, INFINITY, 22.270, INFINITY, -25.685, INFINITY, -36.980, 196.417, -67.148, INFINITY};
static double Tt[ROWS] = {0.0, 8.747, 0.5, 14.0, 3.777, 14.253, 12.428, 3.777, 10.834, 0.5, 6.858, 57.5, 0.0};
static double Nn[ROWS] = {1.0, 1.60738, 1.0, 1.62041, 1.60342, 1.0, 1.0, 1.60342, 1.62041, 1.0, 1.62041, 1.0, 1.0};

#define TRACE(NAME, T, FMA, SQRT, CHEAP)                                                                              \
    static void NAME(double y0, double x0, double u, double v, T* xo, T* yo)                                           \
    {                                                                                                                  \
        T x = x0, y = y0, z = 0, inv = (T)Nn[0] / SQRT((T)v * v + (T)u * u + 1);                                       \
        T Kx = v * inv, Ky = u * inv, Kz = inv;                                                                        \
        for (int i = 0; i < ROWS - 1; i++) {                                                                           \
            const T t = Tt[i], n1 = Nn[i], n2 = Nn[i + 1], dn2 = (n2 - n1) * (n2 + n1);                                \
            if (isinf(Rr[i + 1])) {                                                                                    \
                const T s = (t - z) / Kz;                                                                              \
                x = FMA(s, Kx, x); y = FMA(s, Ky, y); z = 0;                                                           \
                if (n1 != n2) Kz = SQRT(FMA(Kz, Kz, dn2));                                                             \
                continue;                                                                                              \
            }                                                                                                          \
            const T c = (T)1 / (T)Rr[i + 1], cn1sq = c * n1 * n1;                                                      \
            const T zr = z - t;                                                                                        \
            const T PD = FMA(x, Kx, FMA(y, Ky, zr * Kz)), P2 = FMA(x, x, FMA(y, y, zr * zr));                          \
            const T F = FMA(c, P2, -2 * zr), G = FMA(-c, PD, Kz), disc = FMA(G, G, -(cn1sq * F));                      \
            const T ssq = SQRT(disc);                                                                                  \
            const T s = (CHEAP) ? (G - ssq) * ((T)1 / cn1sq) : F / (G + ssq);                                          \
            x = FMA(s, Kx, x); y = FMA(s, Ky, y); z = FMA(s, Kz, zr);                                                  \
            const T g = ssq - SQRT(disc + dn2), gc = g * c;                                                            \
            Kx = FMA(gc, x, Kx); Ky = FMA(gc, y, Ky); Kz = FMA(gc, z, Kz - g);                                         \
        }                                                                                                              \
        *xo = x; *yo = y;                                                                                              \
    }

TRACE(trace_stable, double, fma, sqrt, 0)
TRACE(trace_cheap, double, fma, sqrt, 1)
TRACE(trace_truth, long double, fmal, sqrtl, 0)


/* centre form as the SIMPLE kernels run it: z is carried relative to the reference point of the last surface (a sphere's
 * centre, a plane's vertex), every surface subtracts its constant zoff = t + ref - ref_prev; all square roots are of
 * R^2-scaled quantities, so the incidence cosine needs no extra multiplication: 30 FP64 operations per sphere */
static void trace_centre(double y0, double x0, double u, double v, double* xo, double* yo)
{
    double x = x0, y = y0, z = 0, inv = Nn[0] / sqrt(v * v + u * u + 1);
    double Kx = v * inv, Ky = u * inv, Kz = inv, ref_prev = 0;
    for (int i = 0; i < ROWS - 1; i++) {
        const double t = Tt[i], n1 = Nn[i], n2 = Nn[i + 1], dn2 = (n2 - n1) * (n2 + n1);
        if (isinf(Rr[i + 1])) {
            const double zoff = t - ref_prev;
            const double s = (zoff - z) / Kz;
            x = fma(s, Kx, x); y = fma(s, Ky, y); z = 0; ref_prev = 0;
            if (n1 != n2) Kz = sqrt(fma(Kz, Kz, dn2));
            continue;
        }
        const double R = Rr[i + 1], c = 1 / R, n1sq = n1 * n1, minvn = -1 / n1sq, sg = R < 0 ? -1.0 : 1.0;
        const double zoff = t + R - ref_prev, R2 = R * R, dn2R2 = dn2 * R2, ccabs = c * fabs(c);
        const double w = z - zoff;
        const double b = fma(x, Kx, fma(y, Ky, w * Kz));
        const double q = fma(x, x, fma(y, y, fma(w, w, -R2)));
        const double disc = fma(b, b, -(n1sq * q));
        const double root = sqrt(disc);                      /* |R| n1 cos I */
        const double s = fma(sg, root, b) * minvn;
        x = fma(s, Kx, x); y = fma(s, Ky, y); z = fma(s, Kz, w);
        const double gR = root - sqrt(disc + dn2R2), gc = gR * ccabs;
        Kx = fma(gc, x, Kx); Ky = fma(gc, y, Ky); Kz = fma(gc, z, Kz);
        ref_prev = R;
    }
    *xo = x; *yo = y;
}

static double urand(void) { return (double)rand() / RAND_MAX; }

static int n_rays = 20000;

static void study(const char* label)
{
    double worst_s = 0, worst_c = 0, worst_z = 0;
    const double scale = 25.0;
    srand(1);
    for (int k = 0; k < n_rays; k++) {
        const double y0 = -16 + 32 * urand(), x0 = 16 * urand(), u = tan(0.2374 * urand());
        double xs, ys, xc, yc, xz, yz; long double xt, yt;
        trace_truth(y0, x0, u, 0.0, &xt, &yt);
        if (!isfinite((double)xt) || !isfinite((double)yt)) continue;
        trace_stable(y0, x0, u, 0.0, &xs, &ys);
        trace_cheap(y0, x0, u, 0.0, &xc, &yc);
        trace_centre(y0, x0, u, 0.0, &xz, &yz);
        const double es = fmax(fabs((double)(xs - xt)), fabs((double)(ys - yt))) / scale;
        const double ec = fmax(fabs((double)(xc - xt)), fabs((double)(yc - yt))) / scale;
        if (es > worst_s) worst_s = es;
        if (ec > worst_c) worst_c = ec;
        const double ez = fmax(fabs((double)(xz - xt)), fabs((double)(yz - yt))) / scale;
        if (ez > worst_z) worst_z = ez;
    }
    printf("%-34s stable %.2e   cheap %.2e   centre %.2e   (relative to %.0f mm)\n", label, worst_s, worst_c, worst_z, scale);
}

int main(int argc, char** argv)
{
    if (argc > 1) n_rays = atoi(argv[1]);
    study("double-Gauss, nominal radii");
    const double weak[] = {1e3, 1e4, 1e5, 1e6, 1e8};
    for (int w = 0; w < 5; w++) {
        char lab[64];
        Rr[2] = weak[w];                       /* second surface made progressively weaker */
        snprintf(lab, sizeof lab, "surface 2 with R = %.0e", weak[w]);
        study(lab);
    }
    return 0;
}
