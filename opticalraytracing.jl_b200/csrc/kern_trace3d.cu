// kern_trace3d.cu -- the 3-D skew real-ray trace kernels (sm_100a, FP64 CUDA cores).
//
//   k_grid<ARITH>      K1: pupil-grid sweep, replaces the hot loop of full_trace
//                      (src/PupilSampling.jl:115-138): one thread per ray, whole surface loop in
//                      registers, prescription in the constant bank (__grid_constant__ kernel
//                      parameter), fused mask / eps / r / theta / wavegrad epilogue, per-thread
//                      shifted-moment accumulation + warp-shuffle Chan merge for the spot statistics.
//                      Persistent: gridDim.x * gridDim.y CTAs = SMs * resident CTAs/SM, tile-strided.
//   k_grid_finalize    K6: deterministic fold of the per-CTA partials into ort_stats per field.
//   k_tile_scan / k_compact   ordered compaction in the reference's push! order (:134-137).
//   k_rays<ARITH>      arbitrary rays, every surface recorded (raytrace(...,Vector{RealRay}) :34-65).
//   k_candidates<ARITH> K5: one CTA per candidate prescription staged in shared memory.
#include "kern.cuh"

// ------------------------------------------------------------------------------------------
// per-ray trace drivers shared by the kernels
// ------------------------------------------------------------------------------------------
struct Hit {               // what the grid epilogue needs from one traced ray
    double xs, ys;         // position at the stop surface   (xv[stop], yv[stop])
    double xf, yf;         // position at the last surface   (xv[end],  yv[end])
    unsigned flags;
};

template <class SurfArray>
__device__ __forceinline__ Hit trace_strict(const SurfArray& S, int nsurf, int stop,
                                            double y, double x, double u, double v)
{
    RayS r;
    strict_init(r, y, x, u, v);
    Hit h; h.xs = h.ys = CUDART_NAN;
    for (int i = 0; i < nsurf; i++) {
        strict_step(S[i], r);
        if (i == stop - 1) { h.xs = r.x; h.ys = r.y; }
    }
    h.xf = r.x; h.yf = r.y; h.flags = r.flags;
    return h;
}

template <class SurfArray>
__device__ __forceinline__ Hit trace_fast(const SurfArray& S, int nsurf, int stop,
                                          double y, double x, double u, double v, bool& amb)
{
    RayF r;
    fast_init(r, y, x, u, v);
    Hit h; h.xs = h.ys = CUDART_NAN;
    for (int i = 0; i < nsurf; i++) {
        fast_step(S[i], r);
        if (i == stop - 1) { h.xs = r.x; h.ys = r.y; }
    }
    h.xf = r.x; h.yf = r.y; h.flags = r.flags; amb = r.amb;
    return h;
}

// the strict re-trace of guard-band rays lives out of line so it does not bloat the hot loop
template <class SurfArray>
__device__ __noinline__ Hit trace_strict_cold(const SurfArray& S, int nsurf, int stop,
                                              double y, double x, double u, double v)
{
    return trace_strict(S, nsurf, stop, y, x, u, v);
}

__device__ __forceinline__ bool is_nan_bits(double a)
{
    return (hi32(a) & 0x7FFFFFFF) > 0x7FF00000 ||
           ((hi32(a) & 0x7FFFFFFF) == 0x7FF00000 && __double2loint(a) != 0);
}

// per-thread running moments about the first kept sample (cheap: 2 DADD + 2 DFMA per ray and axis)
struct Acc {
    int n;
    double cx, cy, s1x, s2x, s1y, s2y, rmax;   // rmax holds r (strict) or r^2 (fast)
    int nmiss, ntir, ndom, nclip;
};
__device__ __forceinline__ void acc_zero(Acc& a)
{
    a.n = 0; a.cx = a.cy = a.s1x = a.s2x = a.s1y = a.s2y = 0.0; a.rmax = -CUDART_INF;
    a.nmiss = a.ntir = a.ndom = a.nclip = 0;
}
__device__ __forceinline__ void acc_add(Acc& a, double ex, double ey, double rr)
{
    if (a.n == 0) { a.cx = ex; a.cy = ey; }
    double dx = ex - a.cx, dy = ey - a.cy;
    a.s1x += dx; a.s2x = fma(dx, dx, a.s2x);
    a.s1y += dy; a.s2y = fma(dy, dy, a.s2y);
    a.rmax = fmax(a.rmax, rr);
    a.n++;
}
__device__ __forceinline__ void acc_flags(Acc& a, unsigned f)
{
    a.nmiss += (f & ORT_FLAG_MISS) ? 1 : 0; a.ntir += (f & ORT_FLAG_TIR) ? 1 : 0;
    a.ndom += (f & ORT_FLAG_DOMAIN) ? 1 : 0; a.nclip += (f & ORT_FLAG_CLIP) ? 1 : 0;
}
__device__ __forceinline__ Part acc_to_part(const Acc& a, bool rmax_is_squared)
{
    Part p; part_zero(p);
    p.nmiss = a.nmiss; p.ntir = a.ntir; p.ndom = a.ndom; p.nclip = a.nclip;
    if (a.n > 0) {
        double n = (double)a.n;
        p.n = a.n;
        p.mx = a.cx + a.s1x / n;
        p.my = a.cy + a.s1y / n;
        p.m2x = fmax(a.s2x - a.s1x * a.s1x / n, 0.0);
        p.m2y = fmax(a.s2y - a.s1y * a.s1y / n, 0.0);
        p.rmax = rmax_is_squared ? sqrt(a.rmax) : a.rmax;
    }
    return p;
}

// ------------------------------------------------------------------------------------------
// K1: pupil-grid sweep
// ------------------------------------------------------------------------------------------
template <int ARITH>
__global__ void __launch_bounds__(ORT_TILE)
k_grid(const __grid_constant__ Presc P, const __grid_constant__ GridArgs A)
{
    __shared__ Part s_part[ORT_TILE / 32];
    const int f = blockIdx.y;
    const ort_field fld = A.fields[f];
    const int nsurf = P.nsurf;
    const unsigned NN = A.NN;
    const unsigned ntiles = (NN + ORT_TILE - 1) / ORT_TILE;
    const size_t fbase = (size_t)f * NN;
    Acc acc; acc_zero(acc);

    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned i = tile * ORT_TILE + threadIdx.x;
        int kept = 0;
        if (i < NN) {
            const unsigned iy = i / (unsigned)A.nx, ix = i - iy * (unsigned)A.nx;   // y outer, x inner (:123)
            const double y0 = __ldg(A.ys + (size_t)f * A.ys_stride + iy), x0 = __ldg(A.xs + ix);
            double u = fld.u, v = fld.v;
            if (fld.mode == 1) {                         // RayBasis: :124-127 then tan at :38-39
                u = tan(SD(SS(fld.ybar, y0), fld.z0));
                v = tan(SD(-x0, fld.z0));
            }
            Hit h;
            bool amb = false;
            double ri, r2 = 0.0;
            bool clip;
            if (ARITH == ORT_ARITH_FAST) {
                h = trace_fast(P.s, nsurf, A.stop, y0, x0, u, v, amb);
                r2 = fma(h.xs, h.xs, h.ys * h.ys);
                if (tiny_vs(r2 - A.a_stop2, A.a_stop2)) amb = true;     // within 2^-30 of the stop edge
                clip = r2 > A.a_stop2;
                ri = 0.0;
            }
            if (ARITH == ORT_ARITH_STRICT || amb) {
                h = (ARITH == ORT_ARITH_STRICT) ? trace_strict(P.s, nsurf, A.stop, y0, x0, u, v)
                                                : trace_strict_cold(P.s, nsurf, A.stop, y0, x0, u, v);
                ri = jl_hypot(h.xs, h.ys);                              // :131
                clip = ri > A.a_stop;
                r2 = ri * ri;
            } else if (A.r || A.theta) {
                ri = (r2 > 0.0) ? fast_sqrt(r2) : r2;
            }
            const bool drop = clip || is_nan_bits(h.xf) || is_nan_bits(h.yf);   // :132
            unsigned flags = h.flags | (clip ? ORT_FLAG_CLIP : 0u);
            kept = !drop;
            const double ex = h.xf;                                             // :135
            const double ey = (ARITH == ORT_ARITH_STRICT || amb) ? SS(h.yf, fld.h_prime)
                                                                 : h.yf - fld.h_prime;   // :134
            const size_t o = fbase + i;
            if (A.ex) A.ex[o] = ex;
            if (A.ey) A.ey[o] = ey;
            if (A.r) A.r[o] = ri;                                               // :136
            if (A.theta) A.theta[o] = atan2(h.ys, h.xs);                        // :133
            if (A.wx) A.wx[o] = SD(SM(ex, A.wg_nu), A.wg_lambda);               // :166
            if (A.wy) A.wy[o] = SD(SM(ey, A.wg_nu), A.wg_lambda);
            if (A.mask) A.mask[o] = (uint8_t)kept;
            if (A.flags) A.flags[o] = (uint8_t)flags;
            acc_flags(acc, flags);
            if (kept) acc_add(acc, ex, ey, (ARITH == ORT_ARITH_STRICT) ? ri : r2);
        }
        if (A.tile_counts) {
            int c = __syncthreads_count(kept);
            if (threadIdx.x == 0) A.tile_counts[(size_t)f * ntiles + tile] = c;
        }
    }
    Part p = acc_to_part(acc, ARITH != ORT_ARITH_STRICT);
    part_block_reduce<ORT_TILE / 32>(p, s_part);
    if (threadIdx.x == 0) A.partials[(size_t)f * gridDim.x + blockIdx.x] = p;
}

// K6: fold the per-CTA partials of one field in a fixed order (bit-reproducible run to run).
__global__ void __launch_bounds__(256) k_grid_finalize(const Part* partials, int nparts, ort_stats* stats)
{
    __shared__ Part s_part[8];
    const int f = blockIdx.x;
    Part p; part_zero(p);
    for (int j = threadIdx.x; j < nparts; j += 256) part_merge(p, partials[(size_t)f * nparts + j]);
    part_block_reduce<8>(p, s_part);
    if (threadIdx.x == 0) {
        ort_stats s;
        s.n_kept = p.n; s.mean_x = p.mx; s.mean_y = p.my; s.m2_x = p.m2x; s.m2_y = p.m2y;
        s.r_max = p.rmax;
        s.n_miss = p.nmiss; s.n_tir = p.ntir; s.n_domain = p.ndom; s.n_clip = p.nclip;
        stats[f] = s;
    }
}

// exclusive scan of the per-tile kept counts of one field (one CTA per field)
__global__ void __launch_bounds__(1024) k_tile_scan(int* counts, unsigned ntiles)
{
    __shared__ unsigned s_sum[1024];
    int* c = counts + (size_t)blockIdx.x * ntiles;
    const unsigned per = (ntiles + 1023) / 1024;
    const unsigned lo = threadIdx.x * per, hi = min(lo + per, ntiles);
    unsigned sum = 0;
    for (unsigned j = lo; j < hi; j++) sum += (unsigned)c[j];
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {             // Hillis-Steele inclusive scan
        unsigned v = (threadIdx.x >= d) ? s_sum[threadIdx.x - d] : 0;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = s_sum[threadIdx.x] - sum;
    for (unsigned j = lo; j < hi; j++) { unsigned v = (unsigned)c[j]; c[j] = (int)run; run += v; }
}

// order-preserving scatter of up to 6 arrays: one CTA per (tile, field)
__global__ void __launch_bounds__(ORT_TILE) k_compact(CompactArgs C)
{
    __shared__ int s_warp[ORT_TILE / 32];
    const unsigned tile = blockIdx.x, f = blockIdx.y;
    const unsigned ntiles = (C.NN + ORT_TILE - 1) / ORT_TILE;
    const unsigned i = tile * ORT_TILE + threadIdx.x;
    const size_t fbase = (size_t)f * C.NN;
    const int m = (i < C.NN) ? C.mask[fbase + i] : 0;
    const unsigned ball = __ballot_sync(0xffffffffu, m);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(ball);
    __syncthreads();
    int off = C.tile_offsets[(size_t)f * ntiles + tile];
    for (int w = 0; w < warp; w++) off += s_warp[w];
    off += __popc(ball & ((1u << lane) - 1u));
    if (m) {
#pragma unroll
        for (int a = 0; a < 6; a++)
            if (C.src[a]) C.dst[a][fbase + off] = C.src[a][fbase + i];
    }
}

// ------------------------------------------------------------------------------------------
// arbitrary rays, all surfaces recorded
// ------------------------------------------------------------------------------------------
template <int ARITH>
__global__ void __launch_bounds__(256)
k_rays(const __grid_constant__ Presc P, RaysArgs A)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= A.N) return;
    const double y0 = A.y0[i], x0 = A.x0[i], u0 = A.u0[i], v0 = A.v0[i];
    const int nsurf = P.nsurf;
    bool strict = (ARITH == ORT_ARITH_STRICT);
    if (!strict) {
        RayF r;
        fast_init(r, y0, x0, u0, v0);
        for (int s = 0; s < nsurf; s++) {
            fast_step(P.s[s], r);
            if (A.xv) A.xv[(size_t)s * A.N + i] = r.x;
            if (A.yv) A.yv[(size_t)s * A.N + i] = r.y;
        }
        if (!r.amb) {
            if (A.kout) { A.kout[i] = r.L; A.kout[A.N + i] = r.M; A.kout[2 * A.N + i] = r.N; }
            if (A.flags) A.flags[i] = (uint8_t)r.flags;
        } else strict = true;
    }
    if (strict) {
        RayS r;
        strict_init(r, y0, x0, u0, v0);
        for (int s = 0; s < nsurf; s++) {
            strict_step(P.s[s], r);
            if (A.xv) A.xv[(size_t)s * A.N + i] = r.x;
            if (A.yv) A.yv[(size_t)s * A.N + i] = r.y;
        }
        if (A.kout) { A.kout[i] = r.k1; A.kout[A.N + i] = r.k2; A.kout[2 * A.N + i] = r.k3; }
        if (A.flags) A.flags[i] = (uint8_t)r.flags;
    }
}

// ------------------------------------------------------------------------------------------
// K5: candidate prescriptions, one CTA each, prescription staged in shared memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void derive_surface(SurfK& S, double R, double K, double t, double n1, double n2)
{
    S.R = R; S.K = K; S.t = t; S.n1 = n1; S.n2 = n2;
    S.sgnR = (R < 0.0) ? -1.0 : ((R > 0.0) ? 1.0 : R);
    S.c = isfinite(R) ? 1.0 / R : 0.0;
    S.eta = n1 / n2;
    S.eta2 = S.eta * S.eta;
    S.ome2 = 1.0 - S.eta2;
    S.onepK = 1.0 + K;
    S.kind = !isfinite(R) ? SURF_PLANE : (K == 0.0 ? SURF_SPHERE : SURF_CONIC);
    S.refr = (n1 != n2);
}

template <int ARITH>
__global__ void __launch_bounds__(ORT_TILE)
k_candidates(CandArgs A)
{
    __shared__ SurfK s_surf[ORT_MAX_ROWS - 1];
    __shared__ Part s_part[ORT_TILE / 32];
    const long long c = blockIdx.x;
    const int rows = A.rows, nsurf = rows - 1;
    const double* Rc = A.RtnK + (size_t)c * 4 * rows;
    if (threadIdx.x < nsurf) {
        const int i = threadIdx.x;
        derive_surface(s_surf[i], Rc[i + 1], Rc[3 * rows + i + 1], Rc[rows + i], Rc[2 * rows + i],
                       Rc[2 * rows + i + 1]);
    }
    __syncthreads();
    Acc acc; acc_zero(acc);
    const unsigned NN = (unsigned)A.ny * (unsigned)A.nx;
    for (unsigned i = threadIdx.x; i < NN; i += ORT_TILE) {
        const unsigned iy = i / (unsigned)A.nx, ix = i - iy * (unsigned)A.nx;
        const double y0 = __ldg(A.ys + iy), x0 = __ldg(A.xs + ix);
        Hit h; bool amb = false; double ri = 0.0, r2 = 0.0; bool clip;
        if (ARITH == ORT_ARITH_FAST) {
            h = trace_fast(s_surf, nsurf, A.stop, y0, x0, A.u, A.v, amb);
            r2 = fma(h.xs, h.xs, h.ys * h.ys);
            if (tiny_vs(r2 - A.a_stop2, A.a_stop2)) amb = true;
            clip = r2 > A.a_stop2;
        }
        if (ARITH == ORT_ARITH_STRICT || amb) {
            h = (ARITH == ORT_ARITH_STRICT) ? trace_strict(s_surf, nsurf, A.stop, y0, x0, A.u, A.v)
                                            : trace_strict_cold(s_surf, nsurf, A.stop, y0, x0, A.u, A.v);
            ri = jl_hypot(h.xs, h.ys);
            clip = ri > A.a_stop;
            r2 = ri * ri;
        }
        const bool drop = clip || is_nan_bits(h.xf) || is_nan_bits(h.yf);
        if (!drop) acc_add(acc, h.xf, h.yf - A.h_prime, r2);
    }
    Part p = acc_to_part(acc, true);
    part_block_reduce<ORT_TILE / 32>(p, s_part);
    if (threadIdx.x == 0) {
        double* o = A.out + 4 * c;
        o[0] = (double)p.n;
        if (p.n > 0) { o[1] = p.mx; o[2] = p.my; o[3] = sqrt((p.m2x + p.m2y) / (double)p.n); }
        else { o[1] = o[2] = o[3] = CUDART_NAN; }
    }
}

// ------------------------------------------------------------------------------------------
// launch wrappers (called from ort_api.cu)
// ------------------------------------------------------------------------------------------
int grid_blocks_per_sm(int arith)
{
    int nb = 0;
    cudaError_t e = (arith == ORT_ARITH_FAST)
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_grid<ORT_ARITH_FAST>, ORT_TILE, 0)
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_grid<ORT_ARITH_STRICT>, ORT_TILE, 0);
    if (e != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

cudaError_t launch_grid(const Presc& P, const GridArgs& A, int arith, dim3 grid, cudaStream_t st)
{
    if (arith == ORT_ARITH_FAST) k_grid<ORT_ARITH_FAST><<<grid, ORT_TILE, 0, st>>>(P, A);
    else k_grid<ORT_ARITH_STRICT><<<grid, ORT_TILE, 0, st>>>(P, A);
    return cudaGetLastError();
}

cudaError_t launch_grid_finalize(const Part* partials, int nparts, int n_fields, ort_stats* stats,
                                 cudaStream_t st)
{
    k_grid_finalize<<<n_fields, 256, 0, st>>>(partials, nparts, stats);
    return cudaGetLastError();
}

cudaError_t launch_compact(int* tile_counts, const CompactArgs& C, int n_fields, cudaStream_t st)
{
    const unsigned ntiles = (C.NN + ORT_TILE - 1) / ORT_TILE;
    k_tile_scan<<<n_fields, 1024, 0, st>>>(tile_counts, ntiles);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_compact<<<dim3(ntiles, n_fields), ORT_TILE, 0, st>>>(C);
    return cudaGetLastError();
}

cudaError_t launch_rays(const Presc& P, const RaysArgs& A, int arith, cudaStream_t st)
{
    const unsigned nb = (unsigned)((A.N + 255) / 256);
    if (nb == 0) return cudaSuccess;
    if (arith == ORT_ARITH_FAST) k_rays<ORT_ARITH_FAST><<<nb, 256, 0, st>>>(P, A);
    else k_rays<ORT_ARITH_STRICT><<<nb, 256, 0, st>>>(P, A);
    return cudaGetLastError();
}

cudaError_t launch_candidates(const CandArgs& A, int arith, cudaStream_t st)
{
    if (A.C == 0) return cudaSuccess;
    if (arith == ORT_ARITH_FAST) k_candidates<ORT_ARITH_FAST><<<(unsigned)A.C, ORT_TILE, 0, st>>>(A);
    else k_candidates<ORT_ARITH_STRICT><<<(unsigned)A.C, ORT_TILE, 0, st>>>(A);
    return cudaGetLastError();
}
