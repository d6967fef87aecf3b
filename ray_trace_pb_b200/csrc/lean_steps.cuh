// lean_steps.cuh -- the refracting steps of the "lean" trace kernel (trace_lean.cu): the same roundings as
// surface_steps.cuh (reference raytrace.py:1160-1234, 1323-1347, 1467-1535, 241-306), arranged for the B200's
// operand-read port.
//
// What binds the exact fp64 trace on this chip (tools/ubench/fp64_switch.cu, rf_bandwidth.cu):
// a sub-partition reads ONE 64-bit register operand (an even/odd register pair) per cycle, for every pipe.  A DADD /
// DMUL of two registers occupies the port for 2 cycles (= the FP64 pipe's own 2 cycles), a DFMA of three distinct
// registers for 3, and every integer / select / move instruction in between for 1 or 2 more -- cycles per ray are the
// SUM of those, not the maximum.  So these steps
//   * read the prescription as warp-uniform operands (the surface loop's index is uniform, constants come from the
//     constant bank / uniform registers and cost no port cycle),
//   * update the ray in place (no second copy of the state, no register moves at the loop's back edge),
//   * turn +-1 multiplications and np.sign selects into one integer operation on the sign bit,
//   * pick the sphere's root before it is computed (one subtraction instead of two roots and two selects),
//   * collect every domain test into one predicate with chained compares, and carry NO in-line handling of exact
//     zeros but one (the z component of d x n at a sphere, see unit3_zero_z): a ray whose flag comes back false is re-traced from its launch state by the careful per-surface
//     machinery of surface_steps.cuh (trace_lean.cu: redo_ray), and surfaces where whole bundles produce zeros are
//     found by a probe launch and run the zero-tolerant steps of surface_steps.cuh instead (policy mask).
// A true flag means every intermediate was an ordinary normal number, so every rounding below is the one the
// reference performs and the results are the same bits.  Steps exist for surfaces whose normal / input axis is
// exactly (0, 0, +-1); any other surface runs the general steps of surface_steps.cuh.
#pragma once

#include "exact_math.cuh"
#include "rtb_device.cuh"
#include "surface_steps.cuh"

namespace rtb {
namespace lean {

struct State {          // the loop-carried ray: wavelength and its reciprocal live outside
    double ox, oy, oz;
    double dx, dy, dz;
    double ph;
};

// mag (>= +0, not NaN) with the sign bit of s XOR-ed in: mag * (+-1) and np.sign(s) * mag for s != 0
__device__ __forceinline__ double with_sign_of(double mag, double s)
{
    return __hiloint2double(__double2hiint(mag) ^ (__double2hiint(s) & (int)0x80000000), __double2loint(mag));
}

__device__ __forceinline__ double sqrt_chk(bool &ok, double x)
{
    const int probe = __double2hiint(x) + (int)0xfcb00000;
    ok &= (unsigned)probe < 0x7ca00000u;
    return xm::sqrt_core(x, probe);
}

// |v|^2 -> |v| for a vector that is then divided by |v| (see Optimistic::sqrt_unit: requires 2^-969 <= |v|^2 < 2^104)
__device__ __forceinline__ double sqrt_unit_chk(bool &ok, double x)
{
    const int probe = __double2hiint(x) + (int)0xfcb00000;
    ok &= (unsigned)probe < 0x43200000u;
    return xm::sqrt_core(x, probe);
}

// v / |v| with l = sqrt_unit_chk(|v|^2): quotients of numerators >= 2^-969 are normal (Optimistic::rcp_unit / divz_unit
// without the zero forms: an exactly zero component fails the flag)
__device__ __forceinline__ void unit3(bool &ok, double &x, double &y, double &z, double l)
{
    const double r = xm::refine_rcp(l);
    ok &= xm::num_ok(x) & xm::num_ok(y) & xm::num_ok(z);
    x = xm::div_core(x, l, r);
    y = xm::div_core(y, l, r);
    z = xm::div_core(z, l, r);
}

// The same for d x n at a sphere on the z axis, whose z component may be an exact zero.  For a ray in a plane through the
// lens axis -- every ray of a beam parallel to the axis, at every surface of a coaxial train -- dx ny - dy nx vanishes
// mathematically, and in floating point it is rounding noise that comes out exactly zero for a good part of the rays;
// for a beam along z itself it is (+-0) ny - (+-0) nx.  (+-0) / l = +-0: the fast quotient gives +0 for both, so the sign
// bit of the numerator is copied onto it (a no-op for a non-zero numerator: l > 0).
__device__ __forceinline__ void unit3_zero_z(bool &ok, double &x, double &y, double &z, double l)
{
    const double r = xm::refine_rcp(l);
    const bool z_is_zero = ((__double2hiint(z) & 0x7fffffff) | __double2loint(z)) == 0;
    ok &= xm::num_ok(x) & xm::num_ok(y) & (xm::num_ok(z) | z_is_zero);
    x = xm::div_core(x, l, r);
    y = xm::div_core(y, l, r);
    const double q = xm::div_core(z, l, r);
    z = __hiloint2double(__double2hiint(q) | (__double2hiint(z) & (int)0x80000000), __double2loint(q));
}

// (x, y, z) / l for a general positive l (Optimistic::rcp + div: every quotient range-tested), the z component allowed to
// be an exact zero: a perfect lens on the z axis removes the axial part of a vector exactly (dz - (d . n) nz = +0) and
// measures heights in a plane z = const
__device__ __forceinline__ void div3_zero_z(bool &ok, double &x, double &y, double &z, double l)
{
    const double r = xm::refine_rcp(l);
    const bool z_is_zero = ((__double2hiint(z) & 0x7fffffff) | __double2loint(z)) == 0;
    const double qx = xm::div_core(x, l, r), qy = xm::div_core(y, l, r), qz = xm::div_core(z, l, r);
    ok &= xm::quo_ok(r) & xm::num_ok(x) & xm::num_ok(y) & xm::quo_ok(qx) & xm::quo_ok(qy) &
          (z_is_zero | (xm::num_ok(z) & xm::quo_ok(qz)));
    x = qx;
    y = qy;
    z = __hiloint2double(__double2hiint(qz) | (__double2hiint(z) & (int)0x80000000), __double2loint(qz));
}

__device__ __forceinline__ void unit2(bool &ok, double &x, double &y, double l)
{
    const double r = xm::refine_rcp(l);
    ok &= xm::num_ok(x) & xm::num_ok(y);
    x = xm::div_core(x, l, r);
    y = xm::div_core(y, l, r);
}

// SphericalSurface through RefractingSurface.propagate, input axis (0, 0, +-1).  `wl`, `wl_rcp`: the launch wavelength
// (2^-100 <= |wl| <= 2^100, checked once per ray) and its refined reciprocal; `r_rcp`: refined 1/R, usable (checked per
// block).  Returns "alive": on the sphere inside the aperture and not culled; `kill`: the at-surface slab's row is blank
// (raytrace.py:1187-1192; position and phase of that slab are the ones left in `r`).  Updates the ray in place.
__device__ __forceinline__ bool sphere_axial(bool &ok, const DevSurface &s, double r_rcp, State &r, double n1, double ratio,
                                             double wl, double wl_rcp, bool &kill)
{
    // get_intersect (raytrace.py:1479-1516)
    const double qx = r.ox - s.cx, qy = r.oy - s.cy, qz = r.oz - s.cz;
    const double B = 2.0 * dot3(r.dx, r.dy, r.dz, qx, qy, qz);
    const double C = sumsq3(qx, qy, qz) - s.radius_sq;
    const double root = sqrt_chk(ok, B * B - 4.0 * C);
    // The smaller root 0.5 * (-B - root) is negative exactly when root > -B (halving is exact, the difference of two
    // doubles of this size cannot underflow): then the larger one is taken.  0.5 * ((+-root) - B) is the same addition
    // the reference performs, so one of them is computed instead of two plus selects.  Both negative: no hit, flag.
    const double sroot = __hiloint2double(__double2hiint(root) ^ ((root > -B) ? 0 : (int)0x80000000), __double2loint(root));
    const double t = 0.5 * (sroot - B);
    ok &= t >= 0.0;
    const double px = r.ox + r.dx * t, py = r.oy + r.dy * t, pz = r.oz + r.dz * t;
    const double len = sqrt_chk(ok, sumsq3(px - r.ox, py - r.oy, pz - r.oz));
    // (len * 2 pi) / wl: numerator in [2^-482, 2^515], quotient normal for the wavelengths admitted -- no range tests
    r.ph = r.ph + xm::div_core(len * kTwoPi, wl, wl_rcp) * n1;
    // get_normal (raytrace.py:1476) and is_pt_on_surface (1518-1535)
    const double rx = px - s.cx, ry = py - s.cy, rz = pz - s.cz;
    ok &= xm::num_ok(rx) & xm::num_ok(ry) & xm::num_ok(rz);
    const double nx = xm::div_core(rx, s.radius, r_rcp), ny = xm::div_core(ry, s.radius, r_rcp),
                 nz = xm::div_core(rz, s.radius, r_rcp);
    const double s_on = sumsq3(rx, ry, rz);
    const double s_ap = px * px + py * py;
    // (bitwise, not short-circuit: four chained compares instead of branches around them)
    kill = r.dz * s.az < 0.0;                              // front-side cull, raytrace.py:1187-1192
    const bool on = (s_on >= s.on_sq_lo) & (s_on <= s.on_sq_hi) & (s_ap <= s.ap_sq_max) & !kill;
    // Snell (raytrace.py:1197-1216)
    double bx = r.dy * nz - r.dz * ny;
    double by = r.dz * nx - r.dx * nz;
    double bz = r.dx * ny - r.dy * nx;
    unit3_zero_z(ok, bx, by, bz, sqrt_unit_chk(ok, sumsq3(bx, by, bz)));
    double cx = ny * bz - nz * by;
    double cy = nz * bx - nx * bz;
    double cz = nx * by - ny * bx;
    unit3(ok, cx, cy, cz, sqrt_unit_chk(ok, sumsq3(cx, cy, cz)));
    const double mag_nc = ratio * dot3_np(cx, cy, cz, r.dx, r.dy, r.dz);
    const double cosn = dot3(nx, ny, nz, r.dx, r.dy, r.dz);
    ok &= (cosn != 0.0);                                   // np.sign(0) = 0: left to the careful path
    const double w = with_sign_of(sqrt_chk(ok, 1.0 - mag_nc * mag_nc), cosn);
    r.dx = mag_nc * cx + w * nx;
    r.dy = mag_nc * cy + w * ny;
    r.dz = mag_nc * cz + w * nz;
    r.ox = px; r.oy = py; r.oz = pz;
    return on;
}

// (x, y) / l, l = sqrt_unit_chk(x^2 + y^2), each component allowed to be an exact zero (the zero-tolerant flat step)
__device__ __forceinline__ void unit2_zero(bool &ok, double &x, double &y, double l)
{
    const double r = xm::refine_rcp(l);
    const bool x_zero = ((__double2hiint(x) & 0x7fffffff) | __double2loint(x)) == 0;
    const bool y_zero = ((__double2hiint(y) & 0x7fffffff) | __double2loint(y)) == 0;
    ok &= (xm::num_ok(x) | x_zero) & (xm::num_ok(y) | y_zero);
    const double qx = xm::div_core(x, l, r), qy = xm::div_core(y, l, r);
    x = __hiloint2double(__double2hiint(qx) | (__double2hiint(x) & (int)0x80000000), __double2loint(qx));
    y = __hiloint2double(__double2hiint(qy) | (__double2hiint(y) & (int)0x80000000), __double2loint(qy));
}

// FlatSurface through RefractingSurface.propagate, normal (+-0, +-0, +-1) and input axis (0, 0, +-1).
// With that normal the reference's arithmetic collapses exactly (x - (+-0) = x, (+-0) - x = -x, x + (+-0) = x for
// x != 0):
//   t        = -((oz - cz) nz) / (dz nz)                                   (to_plane's axis shortcut)
//   d x n    = (dy nz, -(dx nz), +-0)            n x nb = (-(nz nb_y), nz nb_x, +-0)
//   nc . d   = (0 + nc_x dx) + nc_y dy                                      n . d = nz dz
//   d'       = (m nc_x + w nx, m nc_y + w ny, w nz)
// ZF = false: the hot form.  Every quantity above is checked non-zero through the flag, the +-0 terms are dropped.
// ZF = true: the form for the surfaces where the probe found whole bundles failing that -- what the reference's scripts
// put in front of almost every system: rays that START ON the plane (t = +-0, whose sign is that of the full three-term
// sum), rays ALONG the normal (d x n = 0: the reference's 0/0 -> NaN -> 0 fix-ups leave nb = nc = (+0, +0, +0), and the
// general formulas evaluated with that nc give d' = +-n with the reference's zero signs), rays in the planes x = 0 or
// y = 0 (one exact-zero component of d x n).
template <bool ZF>
__device__ __forceinline__ bool flat_axial(bool &ok, const DevSurface &s, State &r, double n1, double ratio, double wl,
                                           double wl_rcp, bool &kill)
{
    // propagate_ray2plane (raytrace.py:241-306) with exclude_backward_propagation (303-304)
    double num = (r.oz - s.cz) * s.nz;
    const double den = r.dz * s.nz;
    const bool t_zero = ZF && ((__double2hiint(num) & 0x7fffffff) | __double2loint(num)) == 0;
    if (ZF && t_zero) num = ((r.ox - s.cx) * s.nx + (r.oy - s.cy) * s.ny) + num;      // the sign of the zero
    const double den_rcp = xm::refine_rcp(den);
    const double t_fast = xm::div_core(-num, den, den_rcp);
    const double t = t_zero ? __dmul_rn(-num, den_rcp) : t_fast;
    ok &= xm::den_ok(den) & xm::quo_ok(den_rcp) & (t_zero | (xm::num_ok(num) & xm::quo_ok(t_fast)));
    const double vx = r.dx * t, vy = r.dy * t, vz = r.dz * t;
    const double px = r.ox + vx, py = r.oy + vy, pz = r.oz + vz;
    const double vv = sumsq3(vx, vy, vz);
    const int vv_probe = __double2hiint(vv) + (int)0xfcb00000;
    ok &= t_zero | ((unsigned)vv_probe < 0x7ca00000u);
    const double root = with_sign_of(xm::sqrt_core(vv, vv_probe), t);                 // * prop_direction
    const double len = t_zero ? 0.0 : root;                                           // (+1 for t = -0)
    const double turns = len * kTwoPi;
    r.ph = r.ph + (t_zero ? __dmul_rn(turns, wl_rcp) : xm::div_core(turns, wl, wl_rcp)) * n1;
    // is_pt_on_surface (raytrace.py:1339-1347), front-side cull (1187-1192)
    const double rx = px - s.cx, ry = py - s.cy, rz = pz - s.cz;
    kill = (t < 0.0) | (r.dz * s.az < 0.0);
    const bool on = !kill & (fabs(rz * s.nz) < kOnSurfaceTol) & (sumsq3(rx, ry, rz) <= s.ap_sq_max);
    // Snell in the plane
    double bx = r.dy * s.nz, by = -(r.dx * s.nz);
    double ex, ey;
    if (!ZF) {
        unit2(ok, bx, by, sqrt_unit_chk(ok, bx * bx + by * by));
        double cx = -(s.nz * by), cy = s.nz * bx;
        unit2(ok, cx, cy, sqrt_unit_chk(ok, cx * cx + cy * cy));
        const double mag_nc = ratio * (cx * r.dx + cy * r.dy);
        const double w = with_sign_of(sqrt_chk(ok, 1.0 - mag_nc * mag_nc), den);      // sign(n . d) = sign(nz dz), non-zero
        ex = mag_nc * cx;
        ey = mag_nc * cy;
        // m nc_x + w (+-0) keeps m nc_x only if it is not itself a zero: products this small leave the lean path
        ok &= xm::num_ok(ex) & xm::num_ok(ey);
        r.dz = w * s.nz;
    } else {
        const bool along = (((__double2hiint(bx) | __double2hiint(by)) & 0x7fffffff) | __double2loint(bx) | __double2loint(by)) == 0;
        bool basis_ok = true;
        unit2_zero(basis_ok, bx, by, sqrt_unit_chk(basis_ok, bx * bx + by * by));
        double cx = -(s.nz * by), cy = s.nz * bx;
        unit2_zero(basis_ok, cx, cy, sqrt_unit_chk(basis_ok, cx * cx + cy * cy));
        ok &= along | basis_ok;
        cx = along ? 0.0 : cx;
        cy = along ? 0.0 : cy;
        const double mag_nc = ratio * ((0.0 + cx * r.dx) + cy * r.dy);
        const double w = with_sign_of(sqrt_chk(ok, 1.0 - mag_nc * mag_nc), den);
        ex = mag_nc * cx + w * s.nx;
        ey = mag_nc * cy + w * s.ny;
        r.dz = w * s.nz;
    }
    r.dx = ex;
    r.dy = ey;
    r.ox = px; r.oy = py; r.oz = pz;
    return on;
}

// FlatSurface through RefractingSurface.propagate, any normal and input axis (as stored): refracting_step's flat branch
// with the plain three-term forms, no zero forms.  n . d is the plane propagation's own denominator (the same three
// products, multiplication commutes), so it is not computed twice.
__device__ __forceinline__ bool flat_any(bool &ok, const DevSurface &s, State &r, double n1, double ratio, double wl,
                                         double wl_rcp, bool &kill)
{
    const double num = dot3(r.ox - s.cx, r.oy - s.cy, r.oz - s.cz, s.nx, s.ny, s.nz);
    const double den = dot3(r.dx, r.dy, r.dz, s.nx, s.ny, s.nz);
    const double den_rcp = xm::refine_rcp(den);
    const double t = xm::div_core(-num, den, den_rcp);
    ok &= xm::den_ok(den) & xm::quo_ok(den_rcp) & xm::num_ok(num) & xm::quo_ok(t);
    const double vx = r.dx * t, vy = r.dy * t, vz = r.dz * t;
    const double px = r.ox + vx, py = r.oy + vy, pz = r.oz + vz;
    const double len = with_sign_of(sqrt_chk(ok, sumsq3(vx, vy, vz)), t);      // * prop_direction; t != 0 here
    r.ph = r.ph + xm::div_core(len * kTwoPi, wl, wl_rcp) * n1;
    const double rx = px - s.cx, ry = py - s.cy, rz = pz - s.cz;
    kill = (t < 0.0) | (dot3(r.dx, r.dy, r.dz, s.ax, s.ay, s.az) < 0.0);
    const bool on = !kill & (fabs(dot3(rx, ry, rz, s.nx, s.ny, s.nz)) < kOnSurfaceTol) & (sumsq3(rx, ry, rz) <= s.ap_sq_max);
    double bx = r.dy * s.nz - r.dz * s.ny;
    double by = r.dz * s.nx - r.dx * s.nz;
    double bz = r.dx * s.ny - r.dy * s.nx;
    unit3(ok, bx, by, bz, sqrt_unit_chk(ok, sumsq3(bx, by, bz)));
    double cx = s.ny * bz - s.nz * by;
    double cy = s.nz * bx - s.nx * bz;
    double cz = s.nx * by - s.ny * bx;
    unit3(ok, cx, cy, cz, sqrt_unit_chk(ok, sumsq3(cx, cy, cz)));
    const double mag_nc = ratio * dot3_np(cx, cy, cz, r.dx, r.dy, r.dz);
    const double w = with_sign_of(sqrt_chk(ok, 1.0 - mag_nc * mag_nc), den);   // sign(n . d); den != 0 here
    r.dx = mag_nc * cx + w * s.nx;
    r.dy = mag_nc * cy + w * s.ny;
    r.dz = mag_nc * cz + w * s.nz;
    r.ox = px; r.oy = py; r.oz = pz;
    return on;
}

// PerfectLens.propagate (raytrace.py:1601-1801), any normal: perfect_lens_step of surface_steps.cuh without its general
// zero forms and its "before" slab.  One zero IS handled in line, because whole bundles produce it: a ray that starts
// exactly in the lens's front focal plane (t = +-0) -- in a 4f train the previous surface sits there.  A beam along the
// axis (no transverse direction) or through the front focal point (no height) fails the flag; so does a ray that is
// not culled by the numerical aperture and still has no real cos(theta_2).  `f_rcp`: refined 1 / focal_len (usable,
// checked per block).  Returns "alive" (not culled by the NA test); the lens has no "at" slab here (kill = false).
__device__ __forceinline__ bool lens_any(bool &ok, const DevSurface &s, double f_rcp, State &r, double n1, double n2, double wl,
                                         double wl_rcp)
{
    // front / back focal points (raytrace.py:1682-1687)
    const double fx = s.cx - s.nfx * n1, fy = s.cy - s.nfy * n1, fz = s.cz - s.nfz * n1;
    const double gx = s.cx + s.nfx * n2, gy = s.cy + s.nfy * n2, gz = s.cz + s.nfz * n2;
    // the ray in the front focal plane (raytrace.py:1693-1697): to_plane<ZF = true>
    const double num = dot3(r.ox - fx, r.oy - fy, r.oz - fz, s.nx, s.ny, s.nz);
    const double den = dot3(r.dx, r.dy, r.dz, s.nx, s.ny, s.nz);
    const double den_rcp = xm::refine_rcp(den);
    const bool t_zero = ((__double2hiint(num) & 0x7fffffff) | __double2loint(num)) == 0;
    const double t_fast = xm::div_core(-num, den, den_rcp);
    const double t = t_zero ? __dmul_rn(-num, den_rcp) : t_fast;                   // (+-0) / den = the signed zero
    ok &= xm::den_ok(den) & xm::quo_ok(den_rcp) & (t_zero | (xm::num_ok(num) & xm::quo_ok(t_fast)));
    const double vx = r.dx * t, vy = r.dy * t, vz = r.dz * t;
    const double ax = r.ox + vx, ay = r.oy + vy, az = r.oz + vz;
    const double vv = sumsq3(vx, vy, vz);
    const int vv_probe = __double2hiint(vv) + (int)0xfcb00000;
    ok &= t_zero | ((unsigned)vv_probe < 0x7ca00000u);
    const double len0 = t_zero ? 0.0 : xm::sqrt_core(vv, vv_probe);
    const double len = (t < 0.0) ? -len0 : len0;                                    // * prop_direction (+1 for t = -0)
    const double turns = len * kTwoPi;
    const double ph_ffp = r.ph + (t_zero ? __dmul_rn(turns, wl_rcp) : xm::div_core(turns, wl, wl_rcp)) * n1;
    // transverse unit vector of the direction (raytrace.py:1704-1715)
    const double rnd = dot3_np(r.dx, r.dy, r.dz, s.nx, s.ny, s.nz);
    double px = r.dx - rnd * s.nx, py = r.dy - rnd * s.ny, pz = r.dz - rnd * s.nz;
    // (a vector over its own norm, like d x n at a sphere: with 2^-969 <= |v|^2 < 2^104 and numerators that pass num_ok
    // the quotients are normal numbers -- no range test per quotient)
    const double pn = sqrt_unit_chk(ok, sumsq3(px, py, pz));
    ok &= pn > kPerpTol;                                                            // else: left un-normalised (careful path)
    unit3_zero_z(ok, px, py, pz, pn);
    // height vector in the front focal plane (raytrace.py:1720-1728)
    const double hx = ax - fx, hy = ay - fy, hz = az - fz;
    const double hn = sqrt_unit_chk(ok, sumsq3(hx, hy, hz));
    double ux = hx, uy = hy, uz = hz;
    unit3_zero_z(ok, ux, uy, uz, hn);
    const double sin_t1 = dot3_np(px, py, pz, r.dx, r.dy, r.dz);                   // raytrace.py:1731
    // the ray in the back focal plane (raytrace.py:1736-1752)
    const double scale = (n1 * s.focal_len) * sin_t1;
    const double bx = scale * px + gx, by = scale * py + gy, bz = scale * pz + gz;
    const double n2_rcp = xm::refine_rcp(n2);
    const double s2a = xm::div_core(-hn, s.focal_len, f_rcp);
    const double sin_t2 = xm::div_core(s2a, n2, n2_rcp);
    ok &= xm::quo_ok(s2a) & xm::den_ok(n2) & xm::quo_ok(n2_rcp) & xm::num_ok(s2a) & xm::quo_ok(sin_t2);   // (hn is normal)
    // NA cull (raytrace.py:1757-1760): a culled ray's row is blank whatever else is computed from here on
    const bool culled = (fabs(sin_t1) > s.sin_alpha) | (fabs(sin_t2) > s.sin_alpha);
    bool ok2 = true;                                                                // tests that only matter for a surviving ray
    const double cos_t2 = sqrt_chk(ok2, 1.0 - sin_t2 * sin_t2);
    const double ex = sin_t2 * ux + cos_t2 * s.nx, ey = sin_t2 * uy + cos_t2 * s.ny, ez = sin_t2 * uz + cos_t2 * s.nz;
    const double k = xm::div_core(kTwoPi, wl, wl_rcp);
    const double plane_wave = dot3_np(hx, hy, hz, r.dx, r.dy, r.dz);
    const double ph_b = (ph_ffp - (k * n1) * plane_wave) + k * ((n1 * n1) * s.focal_len + (n2 * n2) * s.focal_len);
    // back to the lens plane in the second medium (raytrace.py:1783-1787)
    const double num2 = dot3(bx - s.cx, by - s.cy, bz - s.cz, s.nx, s.ny, s.nz);
    const double den2 = dot3(ex, ey, ez, s.nx, s.ny, s.nz);
    const double den2_rcp = xm::refine_rcp(den2);
    const double t2 = xm::div_core(-num2, den2, den2_rcp);
    ok2 &= xm::den_ok(den2) & xm::quo_ok(den2_rcp) & xm::num_ok(num2) & xm::quo_ok(t2);
    const double wx = ex * t2, wy = ey * t2, wz = ez * t2;
    const double len2 = with_sign_of(sqrt_chk(ok2, sumsq3(wx, wy, wz)), t2);
    ok &= culled | ok2;
    r.ox = bx + wx; r.oy = by + wy; r.oz = bz + wz;
    r.dx = ex; r.dy = ey; r.dz = ez;
    r.ph = ph_b + xm::div_core(len2 * kTwoPi, wl, wl_rcp) * n2;
    return !culled;
}

} // namespace lean
} // namespace rtb
