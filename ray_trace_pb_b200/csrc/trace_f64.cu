// trace_f64.cu -- the fused fp64 sequential trace: one thread carries one ray in registers through every surface.
//
// Replaces the Python loop of System.ray_trace (reference raytrace.py:641-661) and the whole-array NumPy bodies
// of RefractingSurface.propagate (1160-1234), ReflectingSurface.propagate (1238-1303), PerfectLens.propagate
// (1601-1801), propagate_ray2plane (241-306), SphericalSurface.get_intersect / get_normal / is_pt_on_surface
// (1467-1535), FlatSurface / PlaneMirror geometry (1323-1347, 1392-1412) and Material.n (materials.py:39-51).
//
// THIS TRANSLATION UNIT MUST BE COMPILED WITH -fmad=false.
// The fp64 mode promises results bit-identical to the NumPy path (values and NaN masks), because the reference's
// validity tests use an absolute 1e-12 tolerance that flips on last-bit differences.  That
// needs every + - * rounded on its own (no FMA contraction), 3-term sums associated left to right, IEEE / and
// sqrt, and NaN-aware selects written in the reference's polarity.  Only value-preserving rewrites are used:
//   * x*2*pi == x*(2*pi) (scaling by 2 is exact); norm*(+-1) done as a sign flip; sqrt(disc) evaluated once;
//   * n(lambda) looked up / evaluated once per medium per ray, n1/n2 taken from a host-divided table;
//   * divisions by one denominator share the Newton-refined reciprocal (exact_math.cuh: same bits as IEEE `/`);
//   * `norm <= aperture` and `|norm - |R|| < 1e-12` are tested on the squared sum against host-computed exact
//     thresholds (monotonicity of the correctly rounded sqrt), so those square roots are never taken;
//   * the NaN fills of culled rays are applied once, where a slab is stored or feeds the next surface.
#include <cmath>
#include <math_constants.h>

#include "exact_math.cuh"
#include "rtb_device.cuh"
#include "surface_steps.cuh"

namespace rtb {

namespace {

// per-block shared prescription-derived constants
struct SharedConsts {
    double rcp_radius[kMaxSurfaces]; // refined 1/R per spherical surface (1/f for perfect lenses)
    unsigned long long rcp_ok;       // bit k: den_ok of that denominator
};

// Material.n (materials.py:39-51); Constant.n (materials.py:72-79) ignores the wavelength, NaN included
__device__ __forceinline__ double eval_index(const DevMaterial &m, double wl)
{
    if (m.kind == RTB_MAT_CONSTANT) return m.n_const;
    if (m.kind == RTB_MAT_TABLE_ONLY) return nan64(); // only the host knows this medium
    const double w2 = wl * wl;
    const double acc = (xm::div(m.b0 * w2, w2 - m.c0) + xm::div(m.b1 * w2, w2 - m.c1)) + xm::div(m.b2 * w2, w2 - m.c2);
    return xm::sqrt(acc + 1.0);
}

// a ray whose wavelength is not in the host table (the table may be built from a sample of the batch)
static __device__ __noinline__ double index_for_unlisted(const DevMaterial *m, double wl) { return eval_index(*m, wl); }

// ---- ray I/O ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_ray(const double *base, long long i, Ray &r)
{
    const double *p = base + 8 * i;
    ld256_stream(p, r.ox, r.oy, r.oz, r.dx);
    ld256_stream(p + 4, r.dy, r.dz, r.ph, r.wl);
}

__device__ __forceinline__ void store_ray(double *base, long long i, const Ray &r)
{
    double *p = base + 8 * i;
    st256(p, r.ox, r.oy, r.oz, r.dx);
    st256(p + 4, r.dy, r.dz, r.ph, r.wl);
}

// on-device ray sources (raytrace.py:45-161 and this library's Cartesian grid); see rtb_source in rtb.h
__device__ __forceinline__ double linspace_at(long long i, long long n, double start, double step, double stop)
{
    const double v = (double)i * step + start;
    return (n > 1 && i == n - 1) ? stop : v;
}

__device__ __noinline__ Ray make_ray(const DevSource &g, long long idx)
{
    Ray r;
    r.ph = 0.0;
    r.wl = g.wavelength;
    if (g.kind == RTB_SRC_GRID) {
        const long long iu = idx % g.n_a, iv = idx / g.n_a;
        const double u = linspace_at(iu, g.n_a, g.a_start, g.a_step, g.a_stop);
        const double v = linspace_at(iv, g.n_b, g.b_start, g.b_step, g.b_stop);
        r.ox = (g.px + g.e1x * u) + g.e2x * v;
        r.oy = (g.py + g.e1y * u) + g.e2y * v;
        r.oz = (g.pz + g.e1z * u) + g.e2z * v;
        r.dx = g.axx; r.dy = g.axy; r.dz = g.axz;
    } else if (g.kind == RTB_SRC_COLLIMATED) {
        const long long ip = idx % g.n_b, id = idx / g.n_b;
        const double off = linspace_at(id, g.n_a, g.a_start, g.a_step, g.a_stop);
        const double phi = ((double)ip * kTwoPi) / (double)g.n_b + g.b_start;
        double sp, cp;
        sincos(phi, &sp, &cp);
        const double a = off * cp, b = off * sp;
        r.ox = (g.px + g.e1x * a) + g.e2x * b;
        r.oy = (g.py + g.e1y * a) + g.e2y * b;
        r.oz = (g.pz + g.e1z * a) + g.e2z * b;
        r.dx = g.axx; r.dy = g.axy; r.dz = g.axz;
    } else {
        const long long it = idx % g.n_a, ip = idx / g.n_a;
        const double theta = linspace_at(it, g.n_a, g.a_start, g.a_step, g.a_stop);
        const double phi = ((double)ip * kTwoPi) / (double)g.n_b;
        double st, ct, sp, cp;
        sincos(theta, &st, &ct);
        sincos(phi, &sp, &cp);
        r.ox = g.px; r.oy = g.py; r.oz = g.pz;
        r.dx = (g.axx * ct + (g.e1x * cp) * st) + (g.e2x * sp) * st;
        r.dy = (g.axy * ct + (g.e1y * cp) * st) + (g.e2y * sp) * st;
        r.dz = (g.axz * ct + (g.e1z * cp) * st) + (g.e2z * sp) * st;
    }
    return r;
}

// ---- fused reductions (rtb_reduce in rtb.h) -----------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ void atomic_min_f64(double *addr, double v)
{
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a;
    while (v < __longlong_as_double((long long)old)) {
        const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}
__device__ __forceinline__ void atomic_max_f64(double *addr, double v)
{
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a;
    while (v > __longlong_as_double((long long)old)) {
        const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}

// Running sums of one thread, flushed once per thread block at the end of the kernel.
struct Tally {
    double cnt, su, sv, suu, svv, suv, sp, spp, umin, umax, vmin, vmax;
};

__device__ __forceinline__ void tally_init(Tally &t)
{
    t.cnt = t.su = t.sv = t.suu = t.svv = t.suv = t.sp = t.spp = 0.0;
    t.umin = t.vmin = CUDART_INF;
    t.umax = t.vmax = -CUDART_INF;
}

__device__ __noinline__ void reduce_sample(const DevReduce &R, const Ray r, Tally &t)
{
    const double px = r.ox - R.ox, py = r.oy - R.oy, pz = r.oz - R.oz;
    const double u = dot3(px, py, pz, R.e1x, R.e1y, R.e1z);
    const double v = dot3(px, py, pz, R.e2x, R.e2y, R.e2z);
    const double ph = r.ph - R.phase_ref;
    if (!(isfinite(u) && isfinite(v) && isfinite(ph))) return;
    if (R.stats) {
        t.cnt += 1.0;
        t.su += u; t.sv += v;
        t.suu += u * u; t.svv += v * v; t.suv += u * v;
        t.sp += ph; t.spp += ph * ph;
        t.umin = fmin(t.umin, u); t.umax = fmax(t.umax, u);
        t.vmin = fmin(t.vmin, v); t.vmax = fmax(t.vmax, v);
    }
    if (R.grid) {
        const double fu = floor((u + R.half_width) * R.inv_cell);
        const double fv = floor((v + R.half_width) * R.inv_cell);
        const double g = (double)R.grid_n;
        if (fu >= 0.0 && fu < g && fv >= 0.0 && fv < g) {
            const long long cell = (long long)fv * R.grid_n + (long long)fu;
            const long long plane = (long long)R.grid_n * R.grid_n;
            double s, c;
            sincos(ph, &s, &c);
            atomicAdd(R.grid + cell, c);
            atomicAdd(R.grid + plane + cell, s);
            atomicAdd(R.grid + 2 * plane + cell, 1.0);
        }
    }
}

__device__ __noinline__ void tally_flush(const DevReduce &R, Tally &t)
{
    if (!R.stats) return;
    __shared__ double part[12][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    double v[12] = {warp_sum(t.cnt), warp_sum(t.su),  warp_sum(t.sv),  warp_sum(t.suu),
                    warp_sum(t.svv), warp_sum(t.suv), warp_sum(t.sp),  warp_sum(t.spp),
                    warp_min(t.umin), warp_max(t.umax), warp_min(t.vmin), warp_max(t.vmax)};
    if (lane == 0)
        for (int k = 0; k < 12; k++) part[k][warp] = v[k];
    __syncthreads();
    if (warp == 0) {
        for (int k = 0; k < 12; k++) {
            double x;
            if (k < 8) {
                x = (lane < nwarp) ? part[k][lane] : 0.0;
                x = warp_sum(x);
                if (lane == 0 && x != 0.0) atomicAdd(R.stats + k, x);
            } else if (k == 8 || k == 10) {
                x = (lane < nwarp) ? part[k][lane] : CUDART_INF;
                x = warp_min(x);
                if (lane == 0 && x < CUDART_INF) atomic_min_f64(R.stats + k, x);
            } else {
                x = (lane < nwarp) ? part[k][lane] : -CUDART_INF;
                x = warp_max(x);
                if (lane == 0 && x > -CUDART_INF) atomic_max_f64(R.stats + k, x);
            }
        }
    }
}

// ---- the kernel --------------------------------------------------------------------------------------------
// USE_TABLE   refractive indices (and n1/n2) from the host table (any material), else in-register Sellmeier.
// FROM_SOURCE rays are produced by the on-device source instead of being read from memory.
// MODE        0: only the final slab is stored; 1: general (any slab selection, fused reductions).
template <bool USE_TABLE, bool FROM_SOURCE, int MODE>
__global__ void __launch_bounds__(kTraceThreads, kTraceMinBlocks) trace_f64_kernel(const __grid_constant__ TraceParams P)
{
    constexpr bool GENERAL = MODE == 1;
    __shared__ double s_ntab[USE_TABLE ? (kMaxWavelengths + 1) * kMaxMedia : 1];
    __shared__ double s_ratio[USE_TABLE ? (kMaxWavelengths + 1) * kMaxMedia : 1];
    __shared__ SharedConsts s_c;
    const int n_med = P.n_surf + 1;
    if (USE_TABLE) {
        const int count = (P.n_wl + 1) * n_med;
        for (int k = threadIdx.x; k < count; k += blockDim.x) {
            s_ntab[k] = P.n_tab[k];
            s_ratio[k] = P.ratio_tab[k];
        }
    }
    if (threadIdx.x == 0) s_c.rcp_ok = 0ull;
    __syncthreads();
    for (int k = threadIdx.x; k < P.n_surf; k += blockDim.x) {
        const DevSurface &s = P.surf[k];
        const double den = (s.kind == RTB_SURF_PERFECT_LENS) ? s.focal_len : s.radius;
        s_c.rcp_radius[k] = xm::refine_rcp(den);
        if (xm::den_ok(den)) atomicOr(&s_c.rcp_ok, 1ull << k);
    }
    __syncthreads();

    const bool reducing = GENERAL && P.red.slab >= 0;
    const bool intersect_only = GENERAL && (P.flags & RTB_FLAG_INTERSECT_ONLY) != 0;
    Tally tally;
    if (GENERAL) tally_init(tally);

    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n_rays; i += stride) {
        Ray cur;
        if (FROM_SOURCE)
            cur = make_ray(P.src, P.src.first + i);
        else
            load_ray(P.rays_in, i, cur);

        // refractive indices are functions of the launch wavelength; see DESIGN.md ("n is taken at launch")
        const double wl0 = cur.wl;
        const xm::Rcp rcp_wl = xm::make_rcp(wl0);
        int row = 0;
        bool unlisted = false;
        if (USE_TABLE) {
            row = P.n_wl; // the NaN-wavelength row
            const long long bits = __double_as_longlong(wl0);
#pragma unroll 1
            for (int k = 0; k < P.n_wl; k++)
                if (__double_as_longlong(P.wl[k]) == bits) row = k;
            // a valid wavelength the table does not list: evaluate Sellmeier / constant media for this ray
            unlisted = (row == P.n_wl) && (wl0 == wl0);
            row *= n_med;
        }
        if (GENERAL) {
            if (P.slab_pos[0] >= 0) store_ray(P.out + P.slab_pos[0] * P.out_stride, i, cur);
            if (reducing && P.red.slab == 0) reduce_sample(P.red, cur, tally);
        }

        double n1 = !USE_TABLE ? eval_index(P.mat[0], wl0)
                               : (unlisted ? index_for_unlisted(&P.mat[0], wl0) : s_ntab[row]);
        bool dead = false;
#pragma unroll 1
        for (int k = 0; k < P.n_surf; k++) {
            const DevSurface &s = P.surf[k];
            const double n2 = !USE_TABLE ? eval_index(P.mat[k + 1], wl0)
                                         : (unlisted ? index_for_unlisted(&P.mat[k + 1], wl0) : s_ntab[row + k + 1]);
            // what this surface's two slabs are needed for (uniform): bit 0 store "at", 1 store "after",
            // 2 reduce "at", 3 reduce "after"
            const int act = GENERAL ? P.slab_act[k] : 0;
            const bool need_at = (act & 5) != 0;
            Ray after;
            auto emit_at = [&](const Ray &at) {
                if (act & 1) store_ray(P.out + P.slab_pos[2 * k + 1] * P.out_stride, i, at);
                if (act & 4) reduce_sample(P.red, at, tally);
            };
            if (dead) {
                // an all-NaN ray stays all-NaN through every kind of surface: skip the arithmetic
                set_nan(after);
                if (need_at) emit_at(after);
            } else {
                xm::Rcp rcp_k;
                rcp_k.b = (s.kind == RTB_SURF_PERFECT_LENS) ? s.focal_len : s.radius;
                rcp_k.y = s_c.rcp_radius[k];
                rcp_k.ok = (s_c.rcp_ok >> k) & 1ull;
                // run the surface optimistically; one flag says whether every intermediate stayed in the fast
                // paths' domain, otherwise redo this surface with the Careful arithmetic (surface_steps.cuh)
                Optimistic m;
                AtRaw raw;
                StepResult redo;
                if (s.kind == RTB_SURF_FLAT || s.kind == RTB_SURF_SPHERE) {
                    const double ratio = (USE_TABLE && !unlisted) ? s_ratio[row + k] : xm::div(n1, n2);
                    dead = refracting_step<Optimistic>(m, s, cur, n1, ratio, rcp_wl, rcp_k, !intersect_only, raw, after);
                    if (!m.ok) redo = careful_refracting(&s, cur, n1, ratio, !intersect_only);
                } else if (s.kind == RTB_SURF_MIRROR) {
                    dead = mirror_step<Optimistic>(m, s, cur, n1, rcp_wl, raw, after);
                    if (!m.ok) redo = careful_mirror(&s, cur, n1);
                } else {
                    dead = perfect_lens_step<Optimistic>(m, s, cur, n1, n2, rcp_wl, rcp_k, intersect_only, need_at, raw,
                                                         after);
                    if (!m.ok) redo = careful_lens(&s, cur, n1, n2, intersect_only);
                }
                if (m.ok) {
                    if (need_at) {
                        Ray at;
                        fill_at(raw, cur, at);
                        emit_at(at);
                    }
                } else {
                    if (need_at) emit_at(redo.at);
                    after = redo.after;
                    dead = redo.dead;
                }
            }
            if (GENERAL) {
                if (act & 2) store_ray(P.out + P.slab_pos[2 * k + 2] * P.out_stride, i, after);
                if (act & 8) reduce_sample(P.red, after, tally);
            }
            cur = after;
            n1 = n2;
        }
        if (!GENERAL) store_ray(P.out, i, cur);
    }
    if (GENERAL && reducing) tally_flush(P.red, tally);
}

template <bool T, bool S, int M>
cudaError_t launch_one(const TraceParams &P, unsigned blocks, cudaStream_t stream)
{
    trace_f64_kernel<T, S, M><<<blocks, kTraceThreads, 0, stream>>>(P);
    return cudaGetLastError();
}

} // namespace

// launcher used by rtb_api.cu
cudaError_t launch_trace_f64(const TraceParams &P, int sm_count, cudaStream_t stream)
{
    if (P.n_rays <= 0) return cudaSuccess;
    long long blocks = (P.n_rays + kTraceThreads - 1) / kTraceThreads;
    // a whole number of waves over the SMs, grid-stride inside
    const long long max_blocks = (long long)sm_count * kTraceMinBlocks * 4;
    if (blocks > max_blocks) blocks = max_blocks;
    const bool table = P.n_wl > 0;
    const bool source = P.src.kind >= 0;
    const bool fast = P.store_last_only && P.red.slab < 0 && P.flags == 0;
    const unsigned b = (unsigned)blocks;
    if (fast) {
        if (table) return source ? launch_one<true, true, 0>(P, b, stream) : launch_one<true, false, 0>(P, b, stream);
        return source ? launch_one<false, true, 0>(P, b, stream) : launch_one<false, false, 0>(P, b, stream);
    }
    if (table) return source ? launch_one<true, true, 1>(P, b, stream) : launch_one<true, false, 1>(P, b, stream);
    return source ? launch_one<false, true, 1>(P, b, stream) : launch_one<false, false, 1>(P, b, stream);
}

} // namespace rtb
