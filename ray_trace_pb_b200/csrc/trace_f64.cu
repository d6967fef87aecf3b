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
//   x*2*pi == x*(2*pi) (scaling by 2 is exact), norm*(+-1) done as a sign flip, sqrt(disc) evaluated once,
//   n(lambda) looked up / evaluated once per medium per ray, -(a)/b == (-a)/b.
#include <cmath>
#include <math_constants.h>

#include "rtb_device.cuh"

namespace rtb {

namespace {

constexpr double kPi = 3.141592653589793;       // np.pi
constexpr double kTwoPi = 6.283185307179586;    // 2 * np.pi (exact doubling)
constexpr double kOnSurfaceTol = 1e-12;         // raytrace.py:1343, 1408, 1528
constexpr double kPerpTol = 1e-12;              // raytrace.py:1714

struct Ray {
    double ox, oy, oz;
    double dx, dy, dz;
    double ph;
    double wl;
};

__device__ __forceinline__ double nan64() { return CUDART_NAN; }

__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz)
{
    return (ax * bx + ay * by) + az * bz;
}

__device__ __forceinline__ double norm3(double x, double y, double z) { return sqrt((x * x + y * y) + z * z); }

__device__ __forceinline__ void set_nan(Ray &r)
{
    const double q = nan64();
    r.ox = q; r.oy = q; r.oz = q;
    r.dx = q; r.dy = q; r.dz = q;
    r.ph = q;
    r.wl = q;
}

// np.sign: 0 for 0, NaN for NaN
__device__ __forceinline__ double np_sign(double x)
{
    double s = (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0);
    return (x != x) ? x : s;
}

// Material.n (materials.py:39-51); Constant.n (materials.py:72-79) ignores the wavelength, NaN included
__device__ __forceinline__ double eval_index(const DevMaterial &m, double wl)
{
    if (m.kind == RTB_MAT_CONSTANT) return m.n_const;
    const double w2 = wl * wl;
    const double acc = ((m.b0 * w2) / (w2 - m.c0) + (m.b1 * w2) / (w2 - m.c1)) + (m.b2 * w2) / (w2 - m.c2);
    return sqrt(acc + 1.0);
}

// propagate_ray2plane (raytrace.py:241-306) without the optional back-propagation cull; returns t
__device__ __forceinline__ double to_plane(const Ray &in, double nx, double ny, double nz, double cx, double cy,
                                           double cz, double n_medium, Ray &out)
{
    const double num = ((in.ox - cx) * nx + (in.oy - cy) * ny) + (in.oz - cz) * nz;
    const double den = (in.dx * nx + in.dy * ny) + in.dz * nz;
    const double t = (-num) / den;
    const double vx = in.dx * t, vy = in.dy * t, vz = in.dz * t;
    out.ox = in.ox + vx;
    out.oy = in.oy + vy;
    out.oz = in.oz + vz;
    double len = norm3(vx, vy, vz);
    len = (t < 0.0) ? -len : len;                       // * prop_direction (+-1), raytrace.py:291-297
    out.ph = in.ph + ((len * kTwoPi) / in.wl) * n_medium;
    out.dx = in.dx; out.dy = in.dy; out.dz = in.dz;
    out.wl = in.wl;
    return t;
}

// FlatSurface / PlaneMirror .is_pt_on_surface (raytrace.py:1339-1347, 1405-1412)
__device__ __forceinline__ bool on_plane_surface(const Ray &at, const DevSurface &s)
{
    const double px = at.ox - s.cx, py = at.oy - s.cy, pz = at.oz - s.cz;
    const bool on_plane = fabs(dot3(px, py, pz, s.nx, s.ny, s.nz)) < kOnSurfaceTol;
    const bool in_aperture = norm3(px, py, pz) <= s.aperture;
    return on_plane && in_aperture;
}

// SphericalSurface.is_pt_on_surface (raytrace.py:1518-1535): aperture measured from the axis through the origin
__device__ __forceinline__ bool on_sphere_surface(const Ray &at, const DevSurface &s)
{
    const double px = at.ox - s.cx, py = at.oy - s.cy, pz = at.oz - s.cz;
    const bool on_surface = fabs(norm3(px, py, pz) - s.abs_radius) < kOnSurfaceTol;
    const double along = dot3(at.ox, at.oy, at.oz, s.ax, s.ay, s.az);
    const double qx = at.ox - along * s.ax, qy = at.oy - along * s.ay, qz = at.oz - along * s.az;
    const bool in_aperture = norm3(qx, qy, qz) <= s.aperture;
    return on_surface && in_aperture;
}

// the (normal, nb, nc) construction of raytrace.py:1203-1209 / 1271-1277: returns nc
__device__ __forceinline__ void tangent_basis(double dx, double dy, double dz, double nx, double ny, double nz,
                                              double &cx, double &cy, double &cz)
{
    double bx = dy * nz - dz * ny;
    double by = dz * nx - dx * nz;
    double bz = dx * ny - dy * nx;
    double l = norm3(bx, by, bz);
    bx = bx / l; by = by / l; bz = bz / l;
    bx = (bx != bx) ? 0.0 : bx;
    by = (by != by) ? 0.0 : by;
    bz = (bz != bz) ? 0.0 : bz;
    cx = ny * bz - nz * by;
    cy = nz * bx - nx * bz;
    cz = nx * by - ny * bx;
    l = norm3(cx, cy, cz);
    cx = cx / l; cy = cy / l; cz = cz / l;
    cx = (cx != cx) ? 0.0 : cx;
    cy = (cy != cy) ? 0.0 : cy;
    cz = (cz != cz) ? 0.0 : cz;
}

// the tail every RefractingSurface/ReflectingSurface shares: outgoing ray from the at-surface ray (raytrace.py:1218-1226)
__device__ __forceinline__ void finish_after(const Ray &at, double ex, double ey, double ez, bool on_surface,
                                             Ray &after)
{
    const bool dead_dir = (ex != ex);                     // only the x component is inspected, raytrace.py:1221
    const double q = nan64();
    after.ox = dead_dir ? q : at.ox;
    after.oy = dead_dir ? q : at.oy;
    after.oz = dead_dir ? q : at.oz;
    after.dx = ex; after.dy = ey; after.dz = ez;
    after.ph = at.ph;
    after.wl = at.wl;
    if (!on_surface) set_nan(after);
}

// FlatSurface + SphericalSurface through RefractingSurface.propagate (raytrace.py:1160-1234)
__device__ __forceinline__ void refracting_step(const DevSurface &s, const Ray &in, double n1, double n2,
                                                bool front_cull, Ray &at, Ray &after)
{
    double nx, ny, nz;
    if (s.kind == RTB_SURF_FLAT) {
        // get_intersect with exclude_backward_propagation=True (raytrace.py:1331-1337, 303-304)
        const double t = to_plane(in, s.nx, s.ny, s.nz, s.cx, s.cy, s.cz, n1, at);
        if (t < 0.0) set_nan(at);
        nx = s.nx; ny = s.ny; nz = s.nz;
    } else {
        // SphericalSurface.get_intersect (raytrace.py:1479-1516)
        const double qx = in.ox - s.cx, qy = in.oy - s.cy, qz = in.oz - s.cz;
        const double B = 2.0 * dot3(in.dx, in.dy, in.dz, qx, qy, qz);
        const double C = ((qx * qx + qy * qy) + qz * qz) - s.radius_sq;
        const double root = sqrt(B * B - 4.0 * C);
        double t1 = 0.5 * (-B + root);
        double t2 = 0.5 * (-B - root);
        t1 = (t1 < 0.0) ? CUDART_INF : t1;
        t2 = (t2 < 0.0) ? CUDART_INF : t2;
        double t = (t1 < t2) ? t1 : t2;                    // np.min over the two roots ...
        t = (t1 != t1 || t2 != t2) ? nan64() : t;          // ... which propagates NaN
        t = (t == CUDART_INF) ? nan64() : t;
        at.ox = in.ox + in.dx * t;
        at.oy = in.oy + in.dy * t;
        at.oz = in.oz + in.dz * t;
        const double len = norm3(at.ox - in.ox, at.oy - in.oy, at.oz - in.oz);
        at.ph = in.ph + ((len * kTwoPi) / in.wl) * n1;
        at.dx = in.dx; at.dy = in.dy; at.dz = in.dz;
        at.wl = in.wl;
        // get_normal (raytrace.py:1476): (p - c) / R, sign follows R, not re-normalised
        nx = (at.ox - s.cx) / s.radius;
        ny = (at.oy - s.cy) / s.radius;
        nz = (at.oz - s.cz) / s.radius;
    }
    // front-side cull with the *incoming* direction and input_axis (raytrace.py:1187-1192)
    if (front_cull && dot3(in.dx, in.dy, in.dz, s.ax, s.ay, s.az) < 0.0) set_nan(at);

    // Snell (raytrace.py:1197-1216)
    double cx, cy, cz;
    tangent_basis(at.dx, at.dy, at.dz, nx, ny, nz, cx, cy, cz);
    const double mag_nc = (n1 / n2) * dot3(cx, cy, cz, at.dx, at.dy, at.dz);
    const double w = np_sign(dot3(nx, ny, nz, at.dx, at.dy, at.dz)) * sqrt(1.0 - mag_nc * mag_nc);
    const double ex = mag_nc * cx + w * nx;
    const double ey = mag_nc * cy + w * ny;
    const double ez = mag_nc * cz + w * nz;

    const bool on = (s.kind == RTB_SURF_FLAT) ? on_plane_surface(at, s) : on_sphere_surface(at, s);
    finish_after(at, ex, ey, ez, on, after);
}

// PlaneMirror through ReflectingSurface.propagate (raytrace.py:1238-1303, get_intersect 1398-1403)
__device__ __forceinline__ void mirror_step(const DevSurface &s, const Ray &in, double n1, Ray &at, Ray &after)
{
    const double t = to_plane(in, s.nx, s.ny, s.nz, s.cx, s.cy, s.cz, n1, at);
    if (t < 0.0) set_nan(at);
    double cx, cy, cz;
    tangent_basis(at.dx, at.dy, at.dz, s.nx, s.ny, s.nz, cx, cy, cz);
    const double mag_na = -dot3(s.nx, s.ny, s.nz, at.dx, at.dy, at.dz);
    const double mag_nc = dot3(cx, cy, cz, at.dx, at.dy, at.dz);
    const double ex = mag_na * s.nx + mag_nc * cx;
    const double ey = mag_na * s.ny + mag_nc * cy;
    const double ez = mag_na * s.nz + mag_nc * cz;
    finish_after(at, ex, ey, ez, on_plane_surface(at, s), after);
}

// PerfectLens.propagate (raytrace.py:1601-1801)
__device__ __forceinline__ void perfect_lens_step(const DevSurface &s, const Ray &in, double n1, double n2,
                                                  bool as_get_intersect, Ray &before, Ray &after)
{
    // front / back focal points, per ray because they scale with n(lambda) (raytrace.py:1682-1687)
    const double fx = s.cx - s.nfx * n1, fy = s.cy - s.nfy * n1, fz = s.cz - s.nfz * n1;
    const double gx = s.cx + s.nfx * n2, gy = s.cy + s.nfy * n2, gz = s.cz + s.nfz * n2;

    Ray rf;
    to_plane(in, s.nx, s.ny, s.nz, fx, fy, fz, n1, rf);             // raytrace.py:1693-1697

    // transverse unit vector of the ray direction (raytrace.py:1704-1715)
    const double rnd = dot3(rf.dx, rf.dy, rf.dz, s.nx, s.ny, s.nz);
    double px = rf.dx - rnd * s.nx, py = rf.dy - rnd * s.ny, pz = rf.dz - rnd * s.nz;
    const double pn = norm3(px, py, pz);
    if (pn > kPerpTol) {
        px = px / pn; py = py / pn; pz = pz / pn;
    }
    // height vector in the front focal plane (raytrace.py:1720-1728)
    const double hx = rf.ox - fx, hy = rf.oy - fy, hz = rf.oz - fz;
    const double hn = norm3(hx, hy, hz);
    double ux = hx, uy = hy, uz = hz;
    if (hn != 0.0) {
        ux = ux / hn; uy = uy / hn; uz = uz / hn;
    }
    const double sin_t1 = dot3(px, py, pz, rf.dx, rf.dy, rf.dz);     // raytrace.py:1731

    // ray in the back focal plane (raytrace.py:1736-1752)
    Ray rb;
    const double scale = (n1 * s.focal_len) * sin_t1;
    rb.ox = scale * px + gx;
    rb.oy = scale * py + gy;
    rb.oz = scale * pz + gz;
    const double sin_t2 = ((-hn) / s.focal_len) / n2;
    const double cos_t2 = sqrt(1.0 - sin_t2 * sin_t2);
    rb.dx = sin_t2 * ux + cos_t2 * s.nx;
    rb.dy = sin_t2 * uy + cos_t2 * s.ny;
    rb.dz = sin_t2 * uz + cos_t2 * s.nz;
    rb.wl = in.wl;
    // NA cull blanks the row (raytrace.py:1757-1760) *before* the phase column is written (1775)
    if (fabs(sin_t1) > s.sin_alpha || fabs(sin_t2) > s.sin_alpha) set_nan(rb);
    const double k = kTwoPi / in.wl;
    const double plane_wave = dot3(hx, hy, hz, rf.dx, rf.dy, rf.dz);
    rb.ph = (rf.ph - (k * n1) * plane_wave) + k * ((n1 * n1) * s.focal_len + (n2 * n2) * s.focal_len);

    to_plane(rb, s.nx, s.ny, s.nz, s.cx, s.cy, s.cz, n2, after);     // raytrace.py:1783-1787
    const double tb = to_plane(in, s.nx, s.ny, s.nz, s.cx, s.cy, s.cz, n1, before);  // raytrace.py:1790-1793
    if (as_get_intersect && tb < 0.0) set_nan(before);               // PerfectLens.get_intersect, raytrace.py:1580-1584
}

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

__device__ __forceinline__ void make_ray(const DevSource &g, long long idx, Ray &r)
{
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
    // monotone CAS loop; values are never NaN here
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

// Per-thread running sums, flushed once per thread block at the end of the kernel.
struct Tally {
    double cnt, su, sv, suu, svv, suv, sp, spp, umin, umax, vmin, vmax;
};

__device__ __forceinline__ void tally_init(Tally &t)
{
    t.cnt = t.su = t.sv = t.suu = t.svv = t.suv = t.sp = t.spp = 0.0;
    t.umin = t.vmin = CUDART_INF;
    t.umax = t.vmax = -CUDART_INF;
}

__device__ __forceinline__ void reduce_sample(const DevReduce &R, const Ray &r, bool live, Tally &t)
{
    const double px = r.ox - R.ox, py = r.oy - R.oy, pz = r.oz - R.oz;
    const double u = dot3(px, py, pz, R.e1x, R.e1y, R.e1z);
    const double v = dot3(px, py, pz, R.e2x, R.e2y, R.e2z);
    const double ph = r.ph - R.phase_ref;
    const bool ok = live && isfinite(u) && isfinite(v) && isfinite(ph);
    if (!ok) return;
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

__device__ __forceinline__ void tally_flush(const DevReduce &R, Tally &t)
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
// USE_TABLE: refractive indices from the host table (any material), else in-register Sellmeier / constant.
// FROM_SOURCE: rays are produced by the on-device source instead of being read from memory.
template <bool USE_TABLE, bool FROM_SOURCE>
__global__ void __launch_bounds__(128) trace_f64_kernel(const __grid_constant__ TraceParams P)
{
    __shared__ double s_ntab[USE_TABLE ? (kMaxWavelengths + 1) * kMaxMedia : 1];
    const int n_med = P.n_surf + 1;
    if (USE_TABLE) {
        const int count = (P.n_wl + 1) * n_med;
        for (int k = threadIdx.x; k < count; k += blockDim.x) s_ntab[k] = P.n_tab[k];
        __syncthreads();
    }
    const bool reducing = P.red.slab >= 0;
    const bool intersect_only = (P.flags & RTB_FLAG_INTERSECT_ONLY) != 0;
    Tally tally;
    tally_init(tally);

    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n_rays; i += stride) {
        Ray cur;
        if (FROM_SOURCE)
            make_ray(P.src, P.src.first + i, cur);
        else
            load_ray(P.rays_in, i, cur);

        // refractive indices are functions of the launch wavelength; see DESIGN.md ("n is taken at launch")
        const double wl0 = cur.wl;
        int row = 0;
        if (USE_TABLE) {
            row = P.n_wl; // NaN / unlisted wavelength row
            const long long bits = __double_as_longlong(wl0);
#pragma unroll 1
            for (int k = 0; k < P.n_wl; k++)
                if (__double_as_longlong(P.wl[k]) == bits) row = k;
            row *= n_med;
        }
        if (!P.store_last_only && P.slab_pos[0] >= 0) store_ray(P.out + P.slab_pos[0] * P.out_stride, i, cur);
        if (reducing && P.red.slab == 0) reduce_sample(P.red, cur, true, tally);

        double n1 = USE_TABLE ? s_ntab[row] : eval_index(P.mat[0], wl0);
#pragma unroll 1
        for (int k = 0; k < P.n_surf; k++) {
            const DevSurface &s = P.surf[k];
            const double n2 = USE_TABLE ? s_ntab[row + k + 1] : eval_index(P.mat[k + 1], wl0);
            Ray at, after;
            if (s.kind == RTB_SURF_FLAT || s.kind == RTB_SURF_SPHERE)
                refracting_step(s, cur, n1, n2, !intersect_only, at, after);
            else if (s.kind == RTB_SURF_MIRROR)
                mirror_step(s, cur, n1, at, after);
            else
                perfect_lens_step(s, cur, n1, n2, intersect_only, at, after);

            if (!P.store_last_only) {
                const int pa = P.slab_pos[2 * k + 1], pb = P.slab_pos[2 * k + 2];
                if (pa >= 0) store_ray(P.out + pa * P.out_stride, i, at);
                if (pb >= 0) store_ray(P.out + pb * P.out_stride, i, after);
            }
            if (reducing) {
                if (P.red.slab == 2 * k + 1) reduce_sample(P.red, at, true, tally);
                if (P.red.slab == 2 * k + 2) reduce_sample(P.red, after, true, tally);
            }
            cur = after;
            n1 = n2;
        }
        if (P.store_last_only) store_ray(P.out, i, cur);
    }
    if (reducing) tally_flush(P.red, tally);
}

} // namespace

// launcher used by rtb_api.cu
cudaError_t launch_trace_f64(const TraceParams &P, int sm_count, cudaStream_t stream)
{
    if (P.n_rays <= 0) return cudaSuccess;
    const int threads = 128;
    long long blocks = (P.n_rays + threads - 1) / threads;
    // persistent-style grid: a whole number of waves over the SMs, grid-stride inside
    const long long max_blocks = (long long)sm_count * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    const bool table = P.n_wl > 0;
    const bool source = P.src.kind >= 0;
    if (table && source)
        trace_f64_kernel<true, true><<<(unsigned)blocks, threads, 0, stream>>>(P);
    else if (table)
        trace_f64_kernel<true, false><<<(unsigned)blocks, threads, 0, stream>>>(P);
    else if (source)
        trace_f64_kernel<false, true><<<(unsigned)blocks, threads, 0, stream>>>(P);
    else
        trace_f64_kernel<false, false><<<(unsigned)blocks, threads, 0, stream>>>(P);
    return cudaGetLastError();
}

} // namespace rtb
