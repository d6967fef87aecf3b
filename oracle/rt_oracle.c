/*
 * rt_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar, one-ray-at-a-time CPU restatement of the hot path of QI2lab/ray_trace_pb
 * (System.ray_trace and everything it calls).  It exists only so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg can check / time something that is known to equal the reference; nothing under
 * ray_trace_pb_b200/ may call it.
 *
 * Pinning: the reference's own test-suite does not touch this path (tests/rt_unittest.py covers Seidel sums
 * only), so the oracle is pinned against outputs of the reference itself, run in the build container:
 * tests/golden/make_golden.py imports /root/reference/src/raytrace and writes tests/golden/*.npz, and
 * tests/test_oracle_golden.py requires this file to reproduce them bit for bit (values and NaN masks).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference/src/raytrace/).
 * Arithmetic rules that make the result bit-identical to NumPy (verified against the reference):
 *   - every + - * is rounded individually (build with -ffp-contract=off: no FMA),
 *   - 3-term sums are left associated, exactly like np.sum(axis=1) over 3 columns,
 *   - / and sqrt are the IEEE correctly rounded ones,
 *   - x**2 is x*x,
 *   - comparisons with NaN are false, written in the polarity the reference uses.
 *
 * Surface records arrive as rows of ORC_SURF_STRIDE doubles (see oracle.py):
 *   [0] kind (0 flat, 1 sphere, 2 plane mirror, 3 perfect lens)
 *   [1..3] center  [4..6] normal  [7..9] input_axis  [10] radius  [11] radius**2  [12] abs(radius)
 *   [13] aperture_rad  [14] focal_len  [15..17] normal*focal_len  [18] sin(alpha)
 * Refractive indices arrive as a table ntab[row][medium] with one row per distinct wavelength of the batch
 * (evaluated in Python by the materials' own n()), and wl_row[i] says which row ray i uses.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_SURF_STRIDE 20

typedef struct {
    double o[3];
    double d[3];
    double phase;
    double wl;
} oray;

static const double ORC_PI = 3.141592653589793; /* np.pi */

/* xp.sum(a * b, axis=1) of (N,3) rows.  numpy's add.reduce starts from the identity +0.0, so three products that are
 * all -0 sum to +0 (measured on numpy 2.3.5 for every N: np.sum([[-0., -0., -0.]], axis=1) == [+0.]); for any other
 * operands this is (a0 b0 + a1 b1) + a2 b2.  Sums the reference writes out as a + b + c (propagate_ray2plane,
 * SphericalSurface.get_intersect) have no such start value and are spelled inline below. */
static double dot3(const double *a, const double *b) { return ((0.0 + a[0] * b[0]) + a[1] * b[1]) + a[2] * b[2]; }

/* np.linalg.norm(x, axis=1): sqrt(add.reduce(x*x)) */
static double norm3(const double *a) { return sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]); }

/* np.cross on (N,3) rows */
static void cross3(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

static void ray_set_nan(oray *r)
{
    r->o[0] = r->o[1] = r->o[2] = NAN;
    r->d[0] = r->d[1] = r->d[2] = NAN;
    r->phase = NAN;
    r->wl = NAN;
}

static double np_sign(double x)
{
    if (x > 0) return 1.0;
    if (x < 0) return -1.0;
    if (x == 0) return 0.0;
    return NAN;
}

/*
 * propagate_ray2plane, raytrace.py:241-306.  `c` is the (possibly per-ray) point on the plane.
 * Returns t; writes the ray at the plane into *out.
 */
static double ray2plane(const oray *in, const double *nrm, const double *c, double n_medium, int exclude_backward,
                        oray *out)
{
    /* raytrace.py:287 */
    double num = ((in->o[0] - c[0]) * nrm[0] + (in->o[1] - c[1]) * nrm[1]) + (in->o[2] - c[2]) * nrm[2];
    double den = (in->d[0] * nrm[0] + in->d[1] * nrm[1]) + in->d[2] * nrm[2];
    double t = -num / den;
    /* raytrace.py:291-292: integer +-1, -1 only where t < 0 (False for NaN) */
    double dir = (t < 0) ? -1.0 : 1.0;
    /* raytrace.py:295-297 */
    double v[3] = {in->d[0] * t, in->d[1] * t, in->d[2] * t};
    out->o[0] = in->o[0] + v[0];
    out->o[1] = in->o[1] + v[1];
    out->o[2] = in->o[2] + v[2];
    double shift = norm3(v) * dir * 2 * ORC_PI / in->wl * n_medium;
    out->d[0] = in->d[0];
    out->d[1] = in->d[1];
    out->d[2] = in->d[2];
    out->phase = in->phase + shift;
    out->wl = in->wl;
    /* raytrace.py:303-304 */
    if (exclude_backward && dir == -1.0) ray_set_nan(out);
    return t;
}

/* SphericalSurface.get_intersect, raytrace.py:1479-1516 */
static void sphere_intersect(const oray *in, const double *s, double n_medium, oray *out)
{
    const double *c = s + 1;
    double oc[3] = {in->o[0] - c[0], in->o[1] - c[1], in->o[2] - c[2]};
    double B = 2 * ((in->d[0] * oc[0] + in->d[1] * oc[1]) + in->d[2] * oc[2]);
    double C = ((oc[0] * oc[0] + oc[1] * oc[1]) + oc[2] * oc[2]) - s[11];
    double disc = B * B - 4 * C; /* A = 1 */
    double t1 = 0.5 * (-B + sqrt(disc));
    double t2 = 0.5 * (-B - sqrt(disc));
    if (t1 < 0) t1 = INFINITY;
    if (t2 < 0) t2 = INFINITY;
    /* np.min propagates NaN */
    double t;
    if (isnan(t1) || isnan(t2))
        t = NAN;
    else
        t = (t1 < t2) ? t1 : t2;
    if (t == INFINITY) t = NAN;

    out->o[0] = in->o[0] + in->d[0] * t;
    out->o[1] = in->o[1] + in->d[1] * t;
    out->o[2] = in->o[2] + in->d[2] * t;
    double dp[3] = {out->o[0] - in->o[0], out->o[1] - in->o[1], out->o[2] - in->o[2]};
    double shift = norm3(dp) * 2 * ORC_PI / in->wl * n_medium;
    out->d[0] = in->d[0];
    out->d[1] = in->d[1];
    out->d[2] = in->d[2];
    out->phase = in->phase + shift;
    out->wl = in->wl;
}

/* FlatSurface / PlaneMirror .is_pt_on_surface, raytrace.py:1339-1347, 1405-1412 */
static int plane_pt_on_surface(const double *p, const double *s)
{
    const double *c = s + 1, *nrm = s + 4;
    double pc[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    int on_plane = fabs((pc[0] * nrm[0] + pc[1] * nrm[1]) + pc[2] * nrm[2]) < 1e-12;
    int in_aperture = norm3(pc) <= s[13];
    return on_plane && in_aperture;
}

/* SphericalSurface.is_pt_on_surface, raytrace.py:1518-1535 */
static int sphere_pt_on_surface(const double *p, const double *s)
{
    const double *c = s + 1, *ax = s + 7;
    double pc[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    double dist = norm3(pc);
    int on_surface = fabs(dist - s[12]) < 1e-12;
    /* the aperture is measured from the axis through the ORIGIN: p itself, not p - center */
    double along = (p[0] * ax[0] + p[1] * ax[1]) + p[2] * ax[2];
    double ortho[3] = {p[0] - along * ax[0], p[1] - along * ax[1], p[2] - along * ax[2]};
    int in_aperture = norm3(ortho) <= s[13];
    return on_surface && in_aperture;
}

/* (normal, nb, nc) basis of raytrace.py:1203-1209 / 1271-1277; NaN components are zeroed one by one */
static void snell_basis(const double *ds, const double *nrm, double *nc)
{
    double nb[3];
    cross3(ds, nrm, nb);
    double l = norm3(nb);
    for (int k = 0; k < 3; k++) {
        nb[k] = nb[k] / l;
        if (isnan(nb[k])) nb[k] = 0;
    }
    cross3(nrm, nb, nc);
    l = norm3(nc);
    for (int k = 0; k < 3; k++) {
        nc[k] = nc[k] / l;
        if (isnan(nc[k])) nc[k] = 0;
    }
}

/* RefractingSurface.propagate for FlatSurface and SphericalSurface, raytrace.py:1160-1234 */
static void refracting_propagate(const oray *in, const double *s, double n1, double n2, oray *at, oray *after)
{
    int kind = (int)s[0];
    double nrm[3];
    if (kind == 0) {
        /* FlatSurface.get_intersect / get_normal, raytrace.py:1323-1337 */
        ray2plane(in, s + 4, s + 1, n1, 1, at);
        nrm[0] = s[4];
        nrm[1] = s[5];
        nrm[2] = s[6];
    } else {
        sphere_intersect(in, s, n1, at);
        /* SphericalSurface.get_normal, raytrace.py:1476 */
        nrm[0] = (at->o[0] - s[1]) / s[10];
        nrm[1] = (at->o[1] - s[2]) / s[10];
        nrm[2] = (at->o[2] - s[3]) / s[10];
    }
    /* raytrace.py:1187-1192: uses the direction of the incoming rays and input_axis */
    double cos_in = (in->d[0] * s[7] + in->d[1] * s[8]) + in->d[2] * s[9];
    if (cos_in < 0) ray_set_nan(at);

    /* raytrace.py:1197-1216 */
    const double *ds = at->d;
    double nc[3];
    snell_basis(ds, nrm, nc);
    double mag_nc = n1 / n2 * dot3(nc, ds);
    double sign_na = np_sign(dot3(nrm, ds));
    double w = sign_na * sqrt(1 - mag_nc * mag_nc);
    double dout[3];
    for (int k = 0; k < 3; k++) dout[k] = mag_nc * nc[k] + w * nrm[k];

    /* raytrace.py:1218-1221 */
    after->o[0] = at->o[0];
    after->o[1] = at->o[1];
    after->o[2] = at->o[2];
    after->d[0] = dout[0];
    after->d[1] = dout[1];
    after->d[2] = dout[2];
    after->phase = at->phase;
    after->wl = at->wl;
    if (isnan(dout[0])) after->o[0] = after->o[1] = after->o[2] = NAN;

    /* raytrace.py:1225-1226 */
    int on = (kind == 0) ? plane_pt_on_surface(at->o, s) : sphere_pt_on_surface(at->o, s);
    if (!on) ray_set_nan(after);
}

/* ReflectingSurface.propagate for PlaneMirror, raytrace.py:1238-1303 with get_intersect at 1398-1403 */
static void mirror_propagate(const oray *in, const double *s, double n1, oray *at, oray *after)
{
    const double *nrm = s + 4;
    double t = ray2plane(in, nrm, s + 1, n1, 0, at);
    if (t < 0) ray_set_nan(at);

    const double *ds = at->d;
    double nc[3];
    snell_basis(ds, nrm, nc);
    double mag_na = -dot3(nrm, ds);
    double mag_nc = dot3(nc, ds);
    double dout[3];
    for (int k = 0; k < 3; k++) dout[k] = mag_na * nrm[k] + mag_nc * nc[k];

    after->o[0] = at->o[0];
    after->o[1] = at->o[1];
    after->o[2] = at->o[2];
    after->d[0] = dout[0];
    after->d[1] = dout[1];
    after->d[2] = dout[2];
    after->phase = at->phase;
    after->wl = at->wl;
    if (isnan(dout[0])) after->o[0] = after->o[1] = after->o[2] = NAN;
    if (!plane_pt_on_surface(at->o, s)) ray_set_nan(after);
}

/* PerfectLens.propagate, raytrace.py:1601-1801 */
static void perfect_lens_propagate(const oray *in, const double *s, double n1, double n2, oray *before, oray *after)
{
    const double *c = s + 1, *nrm = s + 4, *nf = s + 15;
    double f = s[14], sin_alpha = s[18];
    double wl = in->wl;

    /* raytrace.py:1682-1687 */
    double ffp[3], bfp[3];
    for (int k = 0; k < 3; k++) {
        ffp[k] = c[k] - nf[k] * n1;
        bfp[k] = c[k] + nf[k] * n2;
    }
    /* raytrace.py:1693-1697 */
    oray rf;
    ray2plane(in, nrm, ffp, n1, 0, &rf);

    /* raytrace.py:1704-1715 */
    const double *s1 = rf.d;
    double rnd = dot3(s1, nrm);
    double sp[3] = {s1[0] - rnd * nrm[0], s1[1] - rnd * nrm[1], s1[2] - rnd * nrm[2]};
    double sp_norm = norm3(sp);
    if (sp_norm > 1e-12)
        for (int k = 0; k < 3; k++) sp[k] = sp[k] / sp_norm;

    /* raytrace.py:1720-1728 */
    double r1[3] = {rf.o[0] - ffp[0], rf.o[1] - ffp[1], rf.o[2] - ffp[2]};
    double r1_norm = norm3(r1);
    double r1u[3] = {r1[0], r1[1], r1[2]};
    if (r1_norm != 0) {
        double l = norm3(r1u);
        for (int k = 0; k < 3; k++) r1u[k] = r1u[k] / l;
    }
    /* raytrace.py:1731 */
    double sin_t1 = dot3(sp, s1);

    /* raytrace.py:1736-1752 */
    oray rb;
    rb.wl = wl;
    for (int k = 0; k < 3; k++) rb.o[k] = n1 * f * sin_t1 * sp[k] + bfp[k];
    double sin_t2 = -r1_norm / f / n2;
    double cos_t2 = sqrt(1 - sin_t2 * sin_t2);
    for (int k = 0; k < 3; k++) rb.d[k] = sin_t2 * r1u[k] + cos_t2 * nrm[k];
    rb.phase = 0;

    /* raytrace.py:1757-1760: NaN the whole row; the phase column is overwritten just below */
    if (fabs(sin_t1) > sin_alpha || fabs(sin_t2) > sin_alpha) ray_set_nan(&rb);

    /* raytrace.py:1773-1777 */
    double pwp = dot3(r1, s1);
    rb.phase = rf.phase - 2 * ORC_PI / wl * n1 * pwp + 2 * ORC_PI / wl * (n1 * n1 * f + n2 * n2 * f);

    /* raytrace.py:1783-1793 */
    ray2plane(&rb, nrm, c, n2, 0, after);
    ray2plane(in, nrm, c, n1, 0, before);
}

/*
 * System.ray_trace, raytrace.py:641-661, for rays [0, n_rays).
 * out: keep_all != 0 -> (2S+1, n_rays, 8) history (slab 0 = copy of the input); else (n_rays, 8) last slab only.
 * n_threads > 1 uses OpenMP when compiled with it (the CPU baseline leg of bench.py); results do not depend on it.
 */
int oracle_trace(const double *surf, int n_surf, const double *ntab, const int32_t *wl_row, const double *rays_in,
                 int64_t n_rays, double *out, int keep_all, int n_threads)
{
    const int n_med = n_surf + 1;
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int64_t i = 0; i < n_rays; i++) {
        oray cur, at, after;
        memcpy(&cur, rays_in + 8 * i, sizeof(oray));
        const double *nrow = ntab + (int64_t)wl_row[i] * n_med;
        if (keep_all) memcpy(out + 8 * i, &cur, sizeof(oray));
        for (int k = 0; k < n_surf; k++) {
            const double *s = surf + (int64_t)k * ORC_SURF_STRIDE;
            double n1 = nrow[k], n2 = nrow[k + 1];
            switch ((int)s[0]) {
            case 0:
            case 1:
                refracting_propagate(&cur, s, n1, n2, &at, &after);
                break;
            case 2:
                mirror_propagate(&cur, s, n1, &at, &after);
                break;
            default:
                perfect_lens_propagate(&cur, s, n1, n2, &at, &after);
                break;
            }
            if (keep_all) {
                memcpy(out + ((int64_t)(2 * k + 1) * n_rays + i) * 8, &at, sizeof(oray));
                memcpy(out + ((int64_t)(2 * k + 2) * n_rays + i) * 8, &after, sizeof(oray));
            }
            cur = after;
        }
        if (!keep_all) memcpy(out + 8 * i, &cur, sizeof(oray));
    }
    return 0;
}

/* Material.n, materials.py:39-51, for checking the in-kernel Sellmeier evaluation */
void oracle_sellmeier(const double *b, const double *c, const double *wl, int64_t n, double *out)
{
    for (int64_t i = 0; i < n; i++) {
        double w2 = wl[i] * wl[i];
        double val = b[0] * w2 / (w2 - c[0]) + b[1] * w2 / (w2 - c[1]) + b[2] * w2 / (w2 - c[2]);
        out[i] = sqrt(val + 1);
    }
}

/*
 * intersect_rays, raytrace.py:164-238 (both inputs already broadcast to n rows).
 * Note raytrace.py:208-212: use_xy / use_yz are logical_and with the *value* of a determinant, i.e. "non-zero"
 * (NaN counts as true).
 */
void oracle_intersect_rays(const double *ray1, const double *ray2, int64_t n, double *pts)
{
    for (int64_t i = 0; i < n; i++) {
        const double *a = ray1 + 8 * i, *b = ray2 + 8 * i;
        double x1 = a[0], y1 = a[1], z1 = a[2], dx1 = a[3], dy1 = a[4], dz1 = a[5];
        double x2 = b[0], y2 = b[1], z2 = b[2], dx2 = b[3], dy2 = b[4], dz2 = b[5];

        double s = NAN;
        double det_xz = dx2 * dz1 - dz2 * dx1;
        double det_xy = dx2 * dy1 - dy2 * dx1;
        double det_yz = dz2 * dy1 - dy2 * dz1;
        int use_xz = det_xz != 0;
        int use_xy = !use_xz && (det_xy != 0);
        int use_yz = !use_xz && !use_xy && (det_yz != 0);
        if (use_xz) s = ((z2 - z1) * dx1 - (x2 - x1) * dz1) / det_xz;
        if (use_xy) s = ((y2 - y1) * dx1 - (x2 - x1) * dy1) / det_xy;
        if (use_yz) s = ((y2 - y1) * dz1 - (z2 - z1) * dy1) / det_yz;

        double t = NAN;
        int use_z = dz1 != 0;
        int use_y = !use_z && (dy1 != 0);
        int use_x = !(use_z || use_y);
        if (use_z) t = (z2 + s * dz2 - z1) / dz1;
        if (use_y) t = (y2 + s * dy2 - y1) / dy1;
        if (use_x) t = (x2 + s * dx2 - x1) / dx1;

        double p1[3] = {x1 + t * dx1, y1 + t * dy1, z1 + t * dz1};
        double p2[3] = {x2 + s * dx2, y2 + s * dy2, z2 + s * dz2};
        /* np.max propagates NaN; NaN > 1e-12 is False, so NaN points pass through as NaN */
        double m = fabs(p1[0] - p2[0]);
        double e1 = fabs(p1[1] - p2[1]), e2 = fabs(p1[2] - p2[2]);
        if (isnan(m) || isnan(e1) || isnan(e2))
            m = NAN;
        else {
            if (e1 > m) m = e1;
            if (e2 > m) m = e2;
        }
        if (m > 1e-12) p1[0] = p1[1] = p1[2] = NAN;
        pts[3 * i] = p1[0];
        pts[3 * i + 1] = p1[1];
        pts[3 * i + 2] = p1[2];
    }
}
