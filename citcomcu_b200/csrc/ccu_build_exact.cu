// ccu_build_exact.cu -- operator construction on the device:
// element geometry (get_global_shape_fn), mass matrix and transfer weights (mass_matrix), pressure
// gradient rows (get_elt_g), viscosity from temperature (get_system_viscosity / visc_from_T),
// viscosity coarsening (project_viscosity), element stiffness + augmented Lagrangian (get_elt_k,
// get_aug_k), node-stored half matrix and inverse diagonal (construct_node_ks, build_diagonal_of_K),
// pressure preconditioner (build_diagonal_of_Ahat) and the body-force vector (assemble_forces).
//
// This translation unit is compiled with -fmad=false: every expression follows the reference's
// operand types and evaluation order (float products where both operands are float, double
// otherwise, float accumulators where the reference accumulates into float arrays), so that the
// coefficients written to HBM are bit-identical to the ones the reference's host code builds.
// Scatter loops of the reference are restated as gathers in ascending element order, which is
// the order the reference's element loops add contributions in.
#include "ccu_ctx.cuh"
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cub/device/device_partition.cuh>
#include <thrust/iterator/counting_iterator.h>

#define CCU_F_VBX 1
#define CCU_F_VBY 2
#define CCU_F_VBZ 4

// master-element tables (construct_shape_functions, Shape_functions.c:53; indices GNVINDEX / GNVXINDEX,
// element_definitions.h:59-63): Nv[8*(n-1)+(v-1)], Nxv[64*d + 8*(n-1) + (v-1)], Np[n-1], Nxp[8*d + (n-1)]
// N is double, Nx float in the reference (global_defs.h:258-269)
struct ShapeTables { double Nv[64], Np[8]; float Nxv[192], Nxp[24]; };
__constant__ ShapeTables c_sh;

static double h_lpoly(int p, double y) { return p == 1 ? 0.5 * (1 - y) : (p == 2 ? 0.5 * (1 + y) : 0.0); }
static double h_lpolydash(int p) { return p == 1 ? -0.5 : (p == 2 ? 0.5 : 0.0); }

static void make_shape_tables(ShapeTables &t)
{
    const float B = (float)0.57735026918962576451;
    const float gx[9][3] = { {0,0,0}, {-B,-B,-B}, {B,-B,-B}, {B,B,-B}, {-B,B,-B}, {-B,-B,B}, {B,-B,B}, {B,B,B}, {-B,B,B} };   // g_point
    const int bb[3][9] = { {0,1,2,2,1,1,2,2,1}, {0,1,1,2,2,1,1,2,2}, {0,1,1,1,1,2,2,2,2} };
    for(int i = 1; i <= 8; i++)
    {
        for(int j = 1; j <= 8; j++)
        {
            double n = 1.0;
            for(int d = 0; d < 3; d++) n *= h_lpoly(bb[d][i], (double)gx[j][d]);
            t.Nv[8 * (i - 1) + (j - 1)] = n;
            for(int dd = 0; dd < 3; dd++)
            {
                float v = (float)h_lpolydash(bb[dd][i]);
                for(int d = 0; d < 3; d++) if(d != dd) v = (float)((double)v * h_lpoly(bb[d][i], (double)gx[j][d]));
                t.Nxv[64 * dd + 8 * (i - 1) + (j - 1)] = v;
            }
        }
        {
            double n = 1.0;
            for(int d = 0; d < 3; d++) n *= h_lpoly(bb[d][i], 0.0);
            t.Np[i - 1] = n;
            for(int dd = 0; dd < 3; dd++)
            {
                float v = (float)h_lpolydash(bb[dd][i]);
                for(int d = 0; d < 3; d++) if(d != dd) v = (float)((double)v * h_lpoly(bb[d][i], 0.0));
                t.Nxp[8 * dd + (i - 1)] = v;
            }
        }
    }
}

__device__ __constant__ int c_OFFS[9][3] = CCU_OFFS_INIT;      // local node -> (dz, dx, dy)

// natural node id of local node a of element (ey, ex, ez)
__device__ __forceinline__ int elt_node(const CcuGeom &g, int ey, int ex, int ez, int a)
{
    return (ez + c_OFFS[a][0]) + g.noz * ((ex + c_OFFS[a][1]) + g.nox * (ey + c_OFFS[a][2]));
}

// get_global_shape_fn (Size_does_matter.c:60-179), Cartesian branch, one integration point:
// NX = master-element derivatives at that point, [d][node]; out: gnx[d][node] (float), return Jacobian (double)
__device__ __forceinline__ double gp_geom(const float X[3][8], const float *NX, int stride_d, int stride_n, float gnx[3][8])
{
    double dxda[3][3];
    for(int d = 0; d < 3; d++) for(int e = 0; e < 3; e++) dxda[d][e] = 0.0;
    for(int i = 0; i < 8; i++)
        for(int d = 0; d < 3; d++)
            for(int e = 0; e < 3; e++)
                dxda[d][e] += X[e][i] * NX[d * stride_d + i * stride_n];         // float * float
    const double (*A)[3] = dxda;
    const double jac = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                       A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    double cof[3][3];
    for(int d = 0; d < 3; d++)
        for(int e = 0; e < 3; e++)
        {
            const int r0 = (d == 0) ? 1 : 0, r1 = (d == 2) ? 1 : 2, c0 = (e == 0) ? 1 : 0, c1 = (e == 2) ? 1 : 2;
            const double det2 = A[r0][c0] * A[r1][c1] - A[r0][c1] * A[r1][c0];
            cof[d][e] = (((d + e) & 1) ? -1 : 1) * det2;
        }
    for(int j = 0; j < 8; j++)
        for(int d = 0; d < 3; d++)
        {
            float v = 0.0f;
            for(int e = 0; e < 3; e++) v = (float)((double)v + NX[e * stride_d + j * stride_n] * cof[e][d]);
            v = (float)((double)v / jac);
            gnx[d][j] = v;
        }
    return jac;
}

__device__ __forceinline__ void load_elt_coords(const CcuGeom &g, const float *__restrict__ XX, int ey, int ex, int ez, float X[3][8])
{
    for(int a = 1; a <= 8; a++)
    {
        const int n = elt_node(g, ey, ex, ez, a);
        for(int e = 0; e < 3; e++) X[e][a - 1] = XX[(size_t)e * g.nno + n];
    }
}

// ---------------------------------------------------------------- geometry-derived arrays
// mass_matrix (Size_does_matter.c:618-757): TWW, ECO.size; get_elt_g (Element_calculations.c:831, CART3D :921-927)
__global__ void __launch_bounds__(128) bk_elt_geometry(const CcuGeom g, const float *__restrict__ XX, float *TWW, double *TWWd, float *eco, float *elt_del)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float X[3][8], gnx[3][8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    double temp[8];
    for(int a = 0; a < 8; a++) temp[a] = 0.0;
    for(int k = 0; k < 8; k++)
    {
        const float gda = (float)gp_geom(X, c_sh.Nxv + k, 64, 8, gnx);
        for(int a = 0; a < 8; a++) temp[a] += gda * 1.0f * c_sh.Nv[8 * a + k];          // (float*float)*double
    }
    for(int a = 0; a < 8; a++) { TWW[(size_t)e * 8 + a] = (float)temp[a]; TWWd[(size_t)e * 8 + a] = temp[a]; }
    {   // element sizes: n[] are 1-based local nodes in the reference expressions (:688-695)
        const float *x1 = X[0], *x2 = X[1], *x3 = X[2];
        float d;
        d = (float)(0.25 * (double)(x1[1] + x1[2] + x1[5] + x1[6] - x1[0] - x1[3] - x1[4] - x1[7]));   // float sum, double scale
        eco[(size_t)e * 3 + 0] = (float)sqrt((double)(d * d));
        d = (float)(0.25 * (double)(x2[2] + x2[3] + x2[6] + x2[7] - x2[0] - x2[1] - x2[4] - x2[5]));
        eco[(size_t)e * 3 + 1] = (float)sqrt((double)(d * d));
        d = (float)(0.25 * (double)(x3[4] + x3[5] + x3[6] + x3[7] - x3[0] - x3[1] - x3[2] - x3[3]));
        eco[(size_t)e * 3 + 2] = (float)sqrt((double)(d * d));
    }
    {   // pressure point: elt_del[p+d] = -GNX.ppt(d,a) * (p_point weight * GDA.ppt)
        const float gdap = (float)gp_geom(X, c_sh.Nxp, 8, 1, gnx);
        const double temp_p = (double)(8.0f * gdap);
        for(int a = 0; a < 8; a++)
            for(int d = 0; d < 3; d++) elt_del[(size_t)e * 24 + 3 * a + d] = (float)(-gnx[d][a] * temp_p);
    }
}


// ---------------------------------------------------------------- regional-spherical geometry (E->control.Rsphere)
// XX holds the CARTESIAN coordinates of the nodes, SXX their (theta, phi, r).  Velocity components at a node are (u_theta, u_phi,
// u_r) in that node's own basis; the operators project them on the basis of the integration point (the Cc / Ccx matrices of
// construct_c3x3matrix_el, Size_does_matter.c:253-440).  The reference evaluates Cc once per radial column of elements
// ((el-1) % ELZ == 0) and reuses it up the column -- the colatitude / longitude of an integration point do not depend on the
// radius in exact arithmetic, but the float node coordinates make it differ by ~1e-8 -- and the kernels do the same (sph_column_point_trig).
struct SphTrig { double ct, st, cf, sf; };
__device__ __forceinline__ SphTrig sph_trig(double tt, double ff) { SphTrig t; t.ct = cos(tt); t.cf = cos(ff); t.st = sin(tt); t.sf = sin(ff); return t; }
__device__ __forceinline__ double sph_myatan(double y, double x) { double fi = atan2(y, x); if(fi < 0.0) fi += 2 * 3.14159265358979323846; return fi; }
// point = sum_a X_a N_a   (N: master-element shape functions at the integration point, stride between nodes `sn`)
__device__ __forceinline__ void sph_point(const float X[3][8], const double *N, int sn, double x[3])
{
    for(int d = 0; d < 3; d++)
    {
        double v = 0.0;
        for(int i = 0; i < 8; i++) v += X[d][i] * N[i * sn];
        x[d] = v;
    }
}
// get_rtf / form_rtf_bc (Size_does_matter.c:182-250): rtf = (theta, phi, 1/r)
__device__ __forceinline__ void sph_rtf(const double x[3], double &th, double &ph, double &ri)
{
    ri = 1.0 / sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    th = acos(x[2] * ri);
    ph = sph_myatan(x[1], x[0]);
}
// basis of the integration point as construct_c3x3matrix_el forms it (tt = acos(x3 / rr))
__device__ __forceinline__ SphTrig sph_point_trig(const double x[3])
{
    const double rr = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    return sph_trig(acos(x[2] / rr), sph_myatan(x[1], x[0]));
}
// derivatives with respect to (theta, phi, r) from the Cartesian ones (get_global_shape_fn sphere branch, :111-126)
__device__ __forceinline__ void sph_rotate_gnx(const double x[3], float gnx[3][8])
{
    double th, ph, ri;
    sph_rtf(x, th, ph, ri);
    double bc[3][3];
    bc[0][0] = x[2] * cos(ph); bc[0][1] = x[2] * sin(ph); bc[0][2] = -sin(th) / ri;
    bc[1][0] = -x[1];          bc[1][1] = x[0];           bc[1][2] = 0.0;
    bc[2][0] = x[0] * ri;      bc[2][1] = x[1] * ri;      bc[2][2] = x[2] * ri;
    for(int j = 0; j < 8; j++)
    {
        const float g0 = gnx[0][j], g1 = gnx[1][j], g2 = gnx[2][j];
        for(int d = 0; d < 3; d++) gnx[d][j] = (float)(bc[d][0] * g0 + bc[d][1] * g1 + bc[d][2] * g2);
    }
}
// rows of the point basis u (theta-hat, phi-hat, r-hat) and their theta / phi derivatives, all dotted with row i of the node basis ua:
//   c[m]  = Cc(m+1, i)        = ua[i] . u[m]
//   c1[m] = Ccx(m+1, i, 1)    = ua[i] . d u[m] / d theta
//   c2[m] = Ccx(m+1, i, 2)    = ua[i] . d u[m] / d phi
__device__ __forceinline__ void sph_cc(const SphTrig &P, const SphTrig &A, int i, double c[3], double c1[3], double c2[3])
{
    double ua[3];
    if(i == 0) { ua[0] = A.ct * A.cf; ua[1] = A.ct * A.sf; ua[2] = -A.st; }
    else if(i == 1) { ua[0] = -A.sf; ua[1] = A.cf; ua[2] = 0.0; }
    else { ua[0] = A.st * A.cf; ua[1] = A.st * A.sf; ua[2] = A.ct; }
    const double u[3][3] = { { P.ct * P.cf, P.ct * P.sf, -P.st }, { -P.sf, P.cf, 0.0 }, { P.st * P.cf, P.st * P.sf, P.ct } };
    const double ux1[3][3] = { { -P.st * P.cf, -P.st * P.sf, -P.ct }, { 0.0, 0.0, 0.0 }, { P.ct * P.cf, P.ct * P.sf, -P.st } };
    const double ux2[3][3] = { { -P.ct * P.sf, P.ct * P.cf, 0.0 }, { -P.cf, -P.sf, 0.0 }, { -P.st * P.sf, P.st * P.cf, 0.0 } };
    for(int m = 0; m < 3; m++)
    {
        c[m] = ua[0] * u[m][0] + ua[1] * u[m][1] + ua[2] * u[m][2];
        c1[m] = ua[0] * ux1[m][0] + ua[1] * ux1[m][1] + ua[2] * ux1[m][2];
        c2[m] = ua[0] * ux2[m][0] + ua[1] * ux2[m][1] + ua[2] * ux2[m][2];
    }
}
// get_ba (Element_calculations.c:300-348), one (node, component, point): the six strain-rate rows
__device__ __forceinline__ void sph_ba(const SphTrig &P, const SphTrig &A, int i, double gnx0, double gnx1, double gnx2, double shp,
                                       double ra, double si, double ct, double ba[6])
{
    double c[3], c1[3], c2[3];
    sph_cc(P, A, i, c, c1, c2);
    const double cc1 = c[0], cc2 = c[1], cc3 = c[2];
    ba[0] = ((gnx0 * cc1 + shp * c1[0]) + shp * cc3) * ra;
    ba[1] = (shp * cc1 * ct + shp * cc3 + (gnx1 * cc2 + shp * c2[1]) * si) * ra;
    ba[2] = gnx2 * cc3;
    ba[3] = ((gnx0 * cc2 + shp * c1[1]) - shp * cc2 * ct + (gnx1 * cc1 + shp * c2[0]) * si) * ra;
    ba[4] = gnx2 * cc1 + (gnx0 * cc3 + shp * (c1[2] - cc1)) * ra;
    ba[5] = gnx2 * cc2 + ((gnx1 * cc3 + shp * c2[2]) * si - shp * cc2) * ra;
}
// strain_rate_2_inv, Rsphere branch (Viscosity_structures.c:996-1040), before the SQRT / halving step: the six strain-rate components
// at the pressure point from the nodal (u_theta, u_phi, u_r) taken as scalars, accumulated in float as the reference's Vxyz
__device__ __forceinline__ float sph_strain2(const float X[3][8], const float VV[3][8], float gnx[3][8])
{
    double x[3], th, ph, ri;
    sph_point(X, c_sh.Np, 1, x);
    sph_rotate_gnx(x, gnx);                          // gnx comes in as the Cartesian derivatives at the pressure point
    sph_rtf(x, th, ph, ri);
    const double ct = cos(th), st = sin(th);
    float v1 = 0.0f, v2 = 0.0f, v3 = 0.0f, v4 = 0.0f, v5 = 0.0f, v6 = 0.0f;
    for(int i = 0; i < 8; i++)
    {
        const double N = c_sh.Np[i];
        v1 = (float)((double)v1 + ((double)(VV[0][i] * gnx[0][i]) + VV[2][i] * N) * ri);
        v2 = (float)((double)v2 + (((double)(VV[1][i] * gnx[1][i]) + VV[0][i] * N * ct) / st + VV[2][i] * N) * ri);
        v3 = v3 + VV[2][i] * gnx[2][i];
        v4 = (float)((double)v4 + (((double)(VV[0][i] * gnx[1][i]) - VV[1][i] * N * ct) / st + (double)(VV[1][i] * gnx[0][i])) * ri);
        v5 = (float)((double)v5 + ((double)(VV[0][i] * gnx[2][i]) + ri * ((double)(VV[2][i] * gnx[0][i]) - VV[0][i] * N)));
        v6 = (float)((double)v6 + ((double)(VV[1][i] * gnx[2][i]) + ri * ((double)(VV[2][i] * gnx[1][i]) / st - VV[1][i] * N)));
    }
    const double e11 = 2.0 * v1, e22 = 2.0 * v2, e33 = 2.0 * v3, e12 = v4, e13 = v5, e23 = v6;
    return (float)(e11 * e11 + e12 * e12 * 2.0 + e22 * e22 + e23 * e23 * 2.0 + e33 * e33 + e13 * e13 * 2.0);
}
// the reference evaluates Cc / Ccx on the FIRST element of a radial column ((el-1) % ELZ == 0) and keeps them for the elements above
// (static struct CC in get_elt_k / get_elt_g / get_elt_f): the point basis comes from that element's (float) coordinates
__device__ __forceinline__ SphTrig sph_column_point_trig(const CcuGeom &g, const float *__restrict__ XX, int ey, int ex, int ez, const float X[3][8],
                                                         const double *N, int sn)
{
    double x[3];
    if(ez == 0) { sph_point(X, N, sn, x); return sph_point_trig(x); }
    float X0[3][8];
    for(int a = 1; a <= 8; a++)
    {
        const int n = (0 + c_OFFS[a][0]) + g.noz * ((ex + c_OFFS[a][1]) + g.nox * (ey + c_OFFS[a][2]));
        for(int e = 0; e < 3; e++) X0[e][a - 1] = XX[(size_t)e * g.nno + n];
    }
    sph_point(X0, N, sn, x);
    return sph_point_trig(x);
}
__device__ __forceinline__ void load_elt_sph(const CcuGeom &g, const float *__restrict__ SXX, int ey, int ex, int ez, SphTrig A[8])
{
    for(int a = 1; a <= 8; a++)
    {
        const int n = elt_node(g, ey, ex, ez, a);
        A[a - 1] = sph_trig((double)SXX[n], (double)SXX[(size_t)g.nno + n]);
    }
}
// mass_matrix ECO.size (Size_does_matter.c:661-680) and get_elt_g (Element_calculations.c:896-920) for Rsphere: overwrite what
// bk_elt_geometry wrote for these two arrays (TWW / MASS use the Jacobian of the Cartesian coordinates in both geometries)
__global__ void __launch_bounds__(128) bk_elt_geometry_sph(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ SXX, float *eco, float *elt_del)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float X[3][8], gnx[3][8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    float S[3][8];
    for(int a = 1; a <= 8; a++)
    {
        const int n = elt_node(g, ey, ex, ez, a);
        for(int d = 0; d < 3; d++) S[d][a - 1] = SXX[(size_t)d * g.nno + n];
    }
    {
        double centre[3];
        for(int i = 0; i < 3; i++)
        {
            double v = 0.0;
            for(int a = 0; a < 8; a++) v += X[i][a];
            centre[i] = v / 8;
        }
        const float c3 = (float)sqrt(centre[0] * centre[0] + centre[1] * centre[1] + centre[2] * centre[2]);
        const float c1 = (float)acos(centre[2] / c3);
        float d1 = (float)fmax(fabs((double)(S[0][2] - S[0][0])), fabs((double)(S[0][1] - S[0][3])));
        eco[(size_t)e * 3 + 0] = d1 * c3;
        float d2 = (float)fmax(fabs((double)(S[1][2] - S[1][0])), fabs((double)(S[1][1] - S[1][3])));
        eco[(size_t)e * 3 + 1] = (float)((double)(d2 * c3) * sin((double)c1));
        const float d3 = (float)(0.25 * (double)(S[2][4] + S[2][5] + S[2][6] + S[2][7] - S[2][0] - S[2][1] - S[2][2] - S[2][3]));
        eco[(size_t)e * 3 + 2] = (float)sqrt((double)(d3 * d3));
    }
    {
        const float gdap = (float)gp_geom(X, c_sh.Nxp, 8, 1, gnx);
        double x[3];
        sph_point(X, c_sh.Np, 1, x);
        sph_rotate_gnx(x, gnx);
        const double temp_p = (double)(8.0f * gdap);
        double th, ph, ra;
        sph_rtf(x, th, ph, ra);
        const double si = 1.0 / sin(th), ct = cos(th) * si;
        const SphTrig P = sph_column_point_trig(g, XX, ey, ex, ez, X, c_sh.Np, 1);
        for(int a = 0; a < 8; a++)
        {
            const SphTrig A = sph_trig((double)S[0][a], (double)S[1][a]);
            const double shp = c_sh.Np[a];
            for(int i = 0; i < 3; i++)
            {
                double c[3], c1[3], c2[3];
                sph_cc(P, A, i, c, c1, c2);
                const double xi = gnx[2][a] * c[2] + 2.0 * ra * shp * c[2] +
                                  ra * (gnx[0][a] * c[0] + shp * c1[0] + ct * shp * c[0] + si * (gnx[1][a] * c[1] + shp * c2[1]));
                elt_del[(size_t)e * 24 + 3 * a + i] = (float)(-xi * temp_p);
            }
        }
    }
}
// get_elt_k, Rsphere branch (Element_calculations.c:249-260): bdbmu[i][j] = sum_k W[k] (2 (ba1 ba1 + ba2 ba2 + ba3 ba3) + ba4 ba4 + ba5 ba5 + ba6 ba6)
struct SphElt
{
    float gn[8][3][8];        // [gauss point][d][node], derivatives with respect to (theta, phi, r)
    double W[8], ra[8], si[8], ct[8];
    SphTrig A[8], P[8];       // node bases, integration-point bases
};
// column_cache: the point bases of the first element of the radial column (get_elt_k called with iconv = 0 from construct_node_ks) or
// of the element itself (iconv = 1, the imposed-velocity term of get_elt_f)
__device__ __forceinline__ void sph_elt_setup(const CcuGeom &g, const float *__restrict__ XX, const float *__restrict__ SXX, const float *__restrict__ EVI,
                                              const int e, const bool column_cache, SphElt &S)
{
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float X[3][8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    load_elt_sph(g, SXX, ey, ex, ez, S.A);
    for(int k = 0; k < 8; k++)
    {
        const float gda = (float)gp_geom(X, c_sh.Nxv + k, 64, 8, S.gn[k]);
        double x[3], th, ph;
        sph_point(X, c_sh.Nv + k, 8, x);
        sph_rotate_gnx(x, S.gn[k]);
        sph_rtf(x, th, ph, S.ra[k]);
        S.si[k] = 1.0 / sin(th);
        S.ct[k] = cos(th) * S.si[k];
        S.P[k] = column_cache ? sph_column_point_trig(g, XX, ey, ex, ez, X, c_sh.Nv + k, 8) : sph_point_trig(x);
        S.W[k] = (double)(1.0f * gda * EVI[(size_t)e * 8 + k]);
    }
}
__device__ __forceinline__ void sph_elt_pair(const SphElt &S, const int a, const int b, double bd[3][3])
{
    for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++) bd[i][j] = 0.0;
    for(int k = 0; k < 8; k++)
    {
        double bb[3][6];
        for(int j = 0; j < 3; j++) sph_ba(S.P[k], S.A[b], j, S.gn[k][0][b], S.gn[k][1][b], S.gn[k][2][b], c_sh.Nv[8 * b + k], S.ra[k], S.si[k], S.ct[k], bb[j]);
        for(int i = 0; i < 3; i++)
        {
            double aa[6];
            sph_ba(S.P[k], S.A[a], i, S.gn[k][0][a], S.gn[k][1][a], S.gn[k][2][a], c_sh.Nv[8 * a + k], S.ra[k], S.si[k], S.ct[k], aa);
            for(int j = 0; j < 3; j++)
                bd[i][j] += S.W[k] * (2.0 * (aa[0] * bb[j][0] + aa[1] * bb[j][1] + aa[2] * bb[j][2]) + aa[3] * bb[j][3] + aa[4] * bb[j][4] + aa[5] * bb[j][5]);
        }
    }
}
__global__ void __launch_bounds__(64) bk_elt_k_sph(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ SXX, const float *__restrict__ EVI,
                                                  const int e_begin, const int e_count, double *blocks)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= e_count) return;
    SphElt S;
    sph_elt_setup(g, XX, SXX, EVI, e_begin + t, true, S);
    int pair = 0;
    for(int a = 0; a < 8; a++)
        for(int b = a; b < 8; b++, pair++)
        {
            double bd[3][3];
            sph_elt_pair(S, a, b, bd);
            for(int i = 0; i < 3; i++)
                for(int j = 0; j < 3; j++) blocks[(size_t)(pair * 9 + 3 * i + j) * e_count + t] = bd[i][j];
        }
}
// get_elt_f, Rsphere branch (Element_calculations.c:1072-1086): the radial body force projected on each node's own basis
__global__ void __launch_bounds__(128) bk_forces_elt_sph(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ SXX,
                                                        const float *__restrict__ buoy, double *EF)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float X[3][8], gnx[3][8], force[8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    SphTrig A[8];
    load_elt_sph(g, SXX, ey, ex, ez, A);
    for(int q = 1; q <= 8; q++) force[q - 1] = buoy[elt_node(g, ey, ex, ez, q)];
    double ef[24];
    for(int p = 0; p < 24; p++) ef[p] = 0.0;
    for(int q = 0; q < 8; q++)
    {
        double fg = 0.0;
        for(int kk = 0; kk < 8; kk++) fg += (double)force[kk] * c_sh.Nv[8 * kk + q];
        const float gda = (float)gp_geom(X, c_sh.Nxv + q, 64, 8, gnx);
        const SphTrig P = sph_column_point_trig(g, XX, ey, ex, ez, X, c_sh.Nv + q, 8);
        for(int a = 0; a < 8; a++)
            for(int i = 0; i < 3; i++)
            {
                double c[3], c1[3], c2[3];
                sph_cc(P, A[a], i, c, c1, c2);
                ef[3 * a + i] += fg * c_sh.Nv[8 * a + q] * gda * 1.0f * c[2];
            }
    }
    for(int p = 0; p < 24; p++) EF[(size_t)p * g.nel + e] = ef[p];
}
__global__ void __launch_bounds__(128) bk_forces_gather_sph(const CcuGeom g, const double *__restrict__ EF, const unsigned char *__restrict__ flags, double *F)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    double f[3] = { 0.0, 0.0, 0.0 };
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int a = LUT[k - ez][j - ex][i - ey] - 1;
                const size_t el = (size_t)(ez + g.elz * (ex + g.elx * ey));
                for(int d = 0; d < 3; d++) f[d] += EF[(size_t)(3 * a + d) * g.nel + el];
            }
        }
    }
    const int s = ccu_sidx(g, i, j, k);
    const unsigned char fl = flags[s];
    F[s] = (fl & CCU_F_VBX) ? 0.0 : f[0];
    F[(size_t)g.NS + s] = (fl & CCU_F_VBY) ? 0.0 : f[1];
    F[2 * (size_t)g.NS + s] = (fl & CCU_F_VBZ) ? 0.0 : f[2];
}

// MASS[node] = 1 / sum over its elements (ascending) of TWW   (float accumulation, Size_does_matter.c:718,739)
__global__ void __launch_bounds__(128) bk_mass(const CcuGeom g, const double *__restrict__ TWWd, float *MASS)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    float m = 0.0f;
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int e = ez + g.elz * (ex + g.elx * ey);
                m = (float)((double)m + TWWd[(size_t)e * 8 + LUT[k - ez][j - ex][i - ey] - 1]);     // float += double temp[node]
            }
        }
    }
    MASS[n] = m;                 // the lumped mass; inverted by bk_invert after the halo sum (Size_does_matter.c:733-739)
}
__global__ void __launch_bounds__(128) bk_invert(const int n, float *v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) v[i] = (float)(1.0 / (double)v[i]);
}
__global__ void __launch_bounds__(128) bk_mul(const int n, float *v, const float *__restrict__ w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) v[i] = v[i] * w[i];
}
__global__ void __launch_bounds__(128) bk_invert_BI(const size_t n, double *v)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if(i < n) v[i] = (v[i] != 0.0) ? 1.0 / v[i] : 0.0;       // halo / padding slots stay zero
}

// ---------------------------------------------------------------- viscosity
// visc_from_T (Viscosity_structures.c:491-742) rheol 0, 1, 2, 3, 4, 10, 11 + visc_from_mat (:477) + min/max clip (:411-425).
// X3 = E->X[3] (Cartesian depth coordinate of the nodes) for the depth-dependent laws 2 and 4.  The operand types follow
// the reference expression by expression (float temperature sums, double depth sums, the double constant 0.5 of law 11).
__global__ void __launch_bounds__(128) bk_visc(const CcuGeom g, const CcuViscParams vp, const int *__restrict__ mat,
                                               const float *__restrict__ T, const float *__restrict__ X3, float *EVI, const int clip)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    const int l = mat[e] - 1;
    const float tempa = vp.N0[l];
    float TT[8], ZZ[8];
    const bool depth = vp.tdepv && (vp.rheol == 2 || vp.rheol == 4);
    for(int a = 1; a <= 8; a++)
    {
        const int n = elt_node(g, ey, ex, ez, a);
        TT[a - 1] = T[n];
        ZZ[a - 1] = depth ? X3[n] : 0.0f;
    }
    for(int jj = 0; jj < 8; jj++)
    {
        float v;
        if(!vp.tdepv) v = tempa;
        else
        {
            float temp = 1.0e-32f;
            double zz = 0.0;
            for(int kk = 0; kk < 8; kk++)
            {
                temp = (float)((double)temp + fmaxf(0.0f, TT[kk]) * c_sh.Nv[8 * kk + jj]);   // float*double
                if(depth) zz += ZZ[kk] * c_sh.Nv[8 * kk + jj];
            }
            const float El = vp.E[l], Tl = vp.T[l], Zl = vp.Z[l];
            switch(vp.rheol)
            {
            case 0: v = (float)((double)tempa * exp((double)(El * (1.0f - temp)))); break;                    // eta0 exp(E (1 - T))
            case 1: v = (float)((double)tempa * exp((double)(El / (temp + Tl)))); break;                      // eta0 exp(E / (T + T0))
            case 2: v = (float)((double)tempa * exp(((double)El + (1 - zz) * (double)Zl) / (double)(temp + Tl))); break;   // eta0 exp((E + (1-z) Z0) / (T + T0))
            case 3: v = (float)((double)tempa * exp((double)(El * (Tl - temp)))); break;                      // eta0 exp(E (T0 - T))
            case 4: v = (float)((double)tempa * exp((double)(El * (Tl - temp)) + (1 - zz) * (double)Zl)); break;            // eta0 exp(E (Tc - T) + (1-z) Z)
            case 10: v = (float)((double)tempa * exp((double)(El / (temp + Tl) + Zl))); break;                // eta0 exp(E / (T + T0) + Z)
            default: v = (float)((double)tempa * exp((double)(El / (temp + Tl)) - (double)El / (0.5 + (double)Tl))); break; // 11: eta0 exp(E/(T+T0) - E/(0.5+T0))
            }
        }
        if(clip && vp.vmax && v > vp.max_value) v = vp.max_value;
        if(clip && vp.vmin && v < vp.min_value) v = vp.min_value;
        EVI[(size_t)e * 8 + jj] = v;
    }
}
// visc_from_S (Viscosity_structures.c:744-975), sdepv_rheology 1 (power law) and 2 (the same without the leading factor two,
// only below a transition temperature and composition): per element the second invariant of the strain rate at the pressure
// point (strain_rate_2_inv with SQRT, :979-1090; 1 on the very first call of a run), then per Gauss point
//     eta <- [2] eta / (1 + scale eta^(1 - 1/n)),   scale = (2 edot / sigma_trans)^(1 - 1/n)
__global__ void __launch_bounds__(64) bk_visc_sdepv(const CcuGeom g, const CcuViscParams vp, const int first, const int *__restrict__ mat,
                                                    const float *__restrict__ XX, const float *__restrict__ V, const float *__restrict__ T,
                                                    const float *__restrict__ Cn, float *EVI, const int sph)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float eedot = 1.0f;
    if(!first)
    {
        float X[3][8], gnx[3][8], VV[3][8];
        load_elt_coords(g, XX, ey, ex, ez, X);
        gp_geom(X, c_sh.Nxp, 8, 1, gnx);
        for(int a = 1; a <= 8; a++)
        {
            const int n = elt_node(g, ey, ex, ez, a);
            for(int d = 0; d < 3; d++) VV[d][a - 1] = V[(size_t)d * g.nno + n];
        }
        double dudx[3][3];
        for(int p = 0; p < 3; p++) for(int q = 0; q < 3; q++) dudx[p][q] = 0.0;
        for(int i = 0; i < 8; i++)
            for(int p = 0; p < 3; p++)
                for(int q = 0; q < 3; q++) dudx[p][q] += VV[p][i] * gnx[q][i];
        double ed[3][3];
        for(int p = 0; p < 3; p++) for(int q = 0; q < 3; q++) ed[p][q] = 0.5 * (dudx[p][q] + dudx[q][p]);
        float ee = (float)(ed[0][0] * ed[0][0] + ed[0][1] * ed[0][1] * 2.0 + ed[1][1] * ed[1][1] + ed[1][2] * ed[1][2] * 2.0 +
                           ed[2][2] * ed[2][2] + ed[0][2] * ed[0][2] * 2.0);
        if(sph) ee = sph_strain2(X, VV, gnx);
        eedot = (float)sqrt(0.5 * (double)ee);
    }
    const int l = mat[e] - 1;
    const float exponent1 = (float)(1.0 - 1.0 / (double)vp.sdepv_expt[l]);
    const float scale = (float)pow(2.0 * (double)eedot / (double)vp.sdepv_trns[l], (double)exponent1);
    float TT[8], CC[8];
    if(vp.sdepv_rheology == 2)
        for(int a = 1; a <= 8; a++)
        {
            const int n = elt_node(g, ey, ex, ez, a);
            TT[a - 1] = T[n]; CC[a - 1] = Cn ? Cn[n] : 0.0f;
        }
    for(int jj = 0; jj < 8; jj++)
    {
        const float eta = EVI[(size_t)e * 8 + jj];
        if(vp.sdepv_rheology == 1)
            EVI[(size_t)e * 8 + jj] = (float)(2.0 * (double)eta / (1.0 + (double)scale * pow((double)eta, (double)exponent1)));
        else
        {
            float temp = 0.0f, comp = 0.0f;
            for(int kk = 0; kk < 8; kk++)
            {
                temp = (float)((double)temp + fmax(0.0, (double)TT[kk]) * c_sh.Nv[8 * kk + jj]);
                comp = (float)((double)comp + fmax(0.0, (double)CC[kk]) * c_sh.Nv[8 * kk + jj]);
            }
            if(temp < vp.sdepv_trns_T && comp < vp.sdepv_trns_c)
                EVI[(size_t)e * 8 + jj] = (float)((double)eta / (1.0 + (double)scale * pow((double)eta, (double)exponent1)));
        }
    }
}
__global__ void __launch_bounds__(256) bk_visc_clip(const size_t n, const CcuViscParams vp, float *EVI)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if(i >= n) return;
    float v = EVI[i];
    if(vp.vmax && v > vp.max_value) v = vp.max_value;
    if(vp.vmin && v < vp.min_value) v = vp.min_value;
    EVI[i] = v;
}

// visc_from_C (Viscosity_structures.c:1784-1935), the modes without flavours / lithosphere overrides: the viscosity of every
// integration point is multiplied with exp(c log(pre_comp[1]) + (1 - c) log(pre_comp[0])), c the composition there (or, with
// cdepv_absolute, replaced by exp(c log(pre_comp[1]) + (1 - c) log(eta)))
__global__ void __launch_bounds__(128) bk_visc_cdepv(const CcuGeom g, const CcuViscParams vp, const int *__restrict__ mat, const float *__restrict__ C, float *EVI)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    const int l2 = 2 * (vp.cdepv_layer ? mat[e] - 1 : 0);
    float CC[8];
    for(int a = 1; a <= 8; a++)
    {
        float v = C[elt_node(g, ey, ex, ez, a)];
        if(vp.cdepv_check_range) { if(v < 0) v = 0.0f; if(v > 1) v = 1.0f; }
        CC[a - 1] = v;
    }
    for(int jj = 0; jj < 8; jj++)
    {
        double cc = 0.0;
        for(int kk = 0; kk < 8; kk++) cc += CC[kk] * c_sh.Nv[8 * kk + jj];
        const size_t q = (size_t)e * 8 + jj;
        if(vp.cdepv_absolute) EVI[q] = (float)exp(cc * vp.cdepv_logv[l2 + 1] + (1.0 - cc) * log((double)EVI[q]));
        else EVI[q] = (float)((double)EVI[q] * exp(cc * vp.cdepv_logv[l2 + 1] + (1.0 - cc) * vp.cdepv_logv[l2]));
    }
}

// visc_from_B (Viscosity_structures.c:1470-1755), regular plasticity: yield stress tau = a depth + b (capped at lambda; or (a z[m] + b) lambda /
// tau_scale in the dimensional form), "Byerlee viscosity" tau / (2 (eII + 1e-7)) + offset, combined with the viscosity so far as a
// harmonic sum (plasticity_trans) or a minimum.  eII: second invariant at the pressure point, 1 on the very first call of a run.
__global__ void __launch_bounds__(64) bk_visc_bdepv(const CcuGeom g, const CcuViscParams vp, const int first, const int *__restrict__ mat,
                                                    const float *__restrict__ XX, const float *__restrict__ depthco, const float *__restrict__ V,
                                                    float *EVI, const int sph)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float eedot = 1.0f;
    if(!first)
    {
        float X[3][8], gnx[3][8], VV[3][8];
        load_elt_coords(g, XX, ey, ex, ez, X);
        gp_geom(X, c_sh.Nxp, 8, 1, gnx);
        for(int a = 1; a <= 8; a++)
        {
            const int n = elt_node(g, ey, ex, ez, a);
            for(int d = 0; d < 3; d++) VV[d][a - 1] = V[(size_t)d * g.nno + n];
        }
        float ee;
        if(sph) ee = sph_strain2(X, VV, gnx);
        else
        {
            double dudx[3][3];
            for(int p = 0; p < 3; p++) for(int q = 0; q < 3; q++) dudx[p][q] = 0.0;
            for(int i = 0; i < 8; i++)
                for(int p = 0; p < 3; p++)
                    for(int q = 0; q < 3; q++) dudx[p][q] += VV[p][i] * gnx[q][i];
            double ed[3][3];
            for(int p = 0; p < 3; p++) for(int q = 0; q < 3; q++) ed[p][q] = 0.5 * (dudx[p][q] + dudx[q][p]);
            ee = (float)(ed[0][0] * ed[0][0] + ed[0][1] * ed[0][1] * 2.0 + ed[1][1] * ed[1][1] + ed[1][2] * ed[1][2] * 2.0 +
                         ed[2][2] * ed[2][2] + ed[0][2] * ed[0][2] * 2.0);
        }
        eedot = (float)sqrt(0.5 * (double)ee);
    }
    const int l = mat[e] - 1;
    float zz[8];
    for(int a = 1; a <= 8; a++)
    {
        zz[a - 1] = (float)(1.0 - (double)depthco[elt_node(g, ey, ex, ez, a)]);
        if(vp.bdepv_dimensional) zz[a - 1] *= vp.bdepv_ndz_to_m;
    }
    for(int jj = 0; jj < 8; jj++)
    {
        float zzz = 0.0f;
        for(int kk = 0; kk < 8; kk++) zzz += zz[kk] * (float)c_sh.Nv[8 * kk + jj];
        float tau;
        if(vp.bdepv_dimensional)
        {
            tau = (vp.abyerlee[l] * zzz + vp.bbyerlee[l]) * vp.lbyerlee[l];
            tau /= vp.bdepv_tau_scale;
        }
        else
        {
            tau = vp.abyerlee[l] * zzz + vp.bbyerlee[l];
            tau = tau < vp.lbyerlee[l] ? tau : vp.lbyerlee[l];
        }
        const float ettby = (float)((double)tau / (2.0 * ((double)eedot + 1e-7)) + (double)vp.bdepv_offset);
        const size_t q = (size_t)e * 8 + jj;
        const float eta = EVI[q];
        EVI[q] = vp.bdepv_trans ? (float)(1.0 / (1.0 / (double)eta + 1.0 / (double)ettby)) : (eta < ettby ? eta : ettby);
    }
}

// visc_from_gint_to_ele (Nodal_mesh.c:559-581): element mean of the eight Gauss-point values (double sum)
__global__ void __launch_bounds__(128) bk_gint_to_ele(const int nel, const float *__restrict__ EVI, float *VN)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= nel) return;
    double t = 0.0;
    for(int i = 0; i < 8; i++) t += EVI[(size_t)e * 8 + i];
    t = t / 8;
    VN[e] = (float)t;
}
// inject_scalar (Solver_multigrid.c:641-673): the coarse node takes the value of the coincident fine node
__global__ void __launch_bounds__(128) bk_inject_scalar(const CcuGeom gc, const CcuGeom gf, const float *__restrict__ AU, float *AD)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= gc.nno) return;
    const int K = n % gc.noz, J = (n / gc.noz) % gc.nox, I = n / (gc.noz * gc.nox);
    AD[n] = AU[2 * K + gf.noz * (2 * J + gf.nox * 2 * I)];
}
// inject_scalar_e (Solver_multigrid.c:675-698): Gauss point i of a coarse element = the mean of its i-th fine sub-element;
// project_scalar_e (:306-342) + visc_from_ele_to_gint (Nodal_mesh.c:541-557): every Gauss point = the mean over the eight
// sub-elements (float sum in the order of EL.sub, times the double weight 1/8)
template <int AVERAGE>
__global__ void __launch_bounds__(128) bk_scalar_e_to_gint(const CcuGeom gc, const CcuGeom gf, const float *__restrict__ VNf, float *EVIc)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= gc.nel) return;
    const int ez = e % gc.elz, ex = (e / gc.elz) % gc.elx, ey = e / (gc.elz * gc.elx);
    float sub[8];
    for(int i = 1; i <= 8; i++)      // EL[lev][e].sub[i]: fine element at twice the coarse indices plus the offset of local node i (Construct_arrays.c:659)
        sub[i - 1] = VNf[(2 * ez + c_OFFS[i][0]) + gf.elz * ((2 * ex + c_OFFS[i][1]) + gf.elx * (2 * ey + c_OFFS[i][2]))];
    if(AVERAGE)
    {
        const double weight = (double)1.0 / 8;
        float average = 0.0f;
        for(int i = 0; i < 8; i++) average += sub[i];
        const float ad = (float)(average * weight);
        for(int i = 0; i < 8; i++) EVIc[(size_t)e * 8 + i] = ad;
    }
    else
        for(int i = 0; i < 8; i++) EVIc[(size_t)e * 8 + i] = sub[i];
}

// visc_from_gint_to_nodes (Nodal_mesh.c:583-615)
__global__ void __launch_bounds__(128) bk_gint_to_nodes(const CcuGeom g, const float *__restrict__ EVI, const float *__restrict__ TWW,
                                                         const float *__restrict__ MASS, float *VN)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    float v = 0.0f;
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int e = ez + g.elz * (ex + g.elx * ey);
                double tv = 0.0;
                for(int q = 0; q < 8; q++) tv += EVI[(size_t)e * 8 + q];
                tv = tv / 8;
                v = (float)((double)v + TWW[(size_t)e * 8 + LUT[k - ez][j - ex][i - ey] - 1] * tv);
            }
        }
    }
    VN[n] = MASS ? v * MASS[n] : v;
}
// project_scalar (Solver_multigrid.c:344-388): fine nodal field -> coarse nodal field
__global__ void __launch_bounds__(128) bk_project_scalar(const CcuGeom gc, const CcuGeom gf, const float *__restrict__ TWWc,
                                                          const float *__restrict__ MASSc, const float *__restrict__ AU, float *AD)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= gc.nno) return;
    const int K = n % gc.noz, J = (n / gc.noz) % gc.nox, I = n / (gc.noz * gc.nox);
    const double weight = (double)1.0 / 8;
    float ad = 0.0f;
    for(int ey = I - 1; ey <= I; ey++)
    {
        if(ey < 0 || ey >= gc.ely) continue;
        for(int ex = J - 1; ex <= J; ex++)
        {
            if(ex < 0 || ex >= gc.elx) continue;
            for(int ez = K - 1; ez <= K; ez++)
            {
                if(ez < 0 || ez >= gc.elz) continue;
                const int oy = I - ey, ox = J - ex, oz = K - ez;
                const int el = ez + gc.elz * (ex + gc.elx * ey);
                float average = 0.0f;
                for(int q = 1; q <= 8; q++) average += AU[elt_node(gf, 2 * ey + oy, 2 * ex + ox, 2 * ez + oz, q)];
                const float w = (float)(weight * average);
                ad += w * TWWc[(size_t)el * 8 + LUT[oz][ox][oy] - 1];
            }
        }
    }
    AD[n] = MASSc ? ad * MASSc[n] : ad;
}
// visc_from_nodes_to_gint (Nodal_mesh.c:617-640)
__global__ void __launch_bounds__(128) bk_nodes_to_gint(const CcuGeom g, const float *__restrict__ VN, float *EVI)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float vn[8];
    for(int a = 1; a <= 8; a++) vn[a - 1] = VN[elt_node(g, ey, ex, ez, a)];
    for(int i = 0; i < 8; i++)
    {
        double tv = 0.0;
        for(int j = 0; j < 8; j++) tv += c_sh.Nv[8 * j + i] * vn[j];     // double * float
        EVI[(size_t)e * 8 + i] = (float)tv;
    }
}

// ---------------------------------------------------------------- element stiffness
// get_elt_k (Element_calculations.c:133-293, CART3D :223-248): the 36 node-pair blocks a<=b of one
// element, bdbmu[i][j], written SoA (pair*9 + 3*i + j) over the chunk's elements.
__global__ void __launch_bounds__(64) bk_elt_k(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ EVI,
                                              const int e_begin, const int e_count, double *blocks)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= e_count) return;
    const int e = e_begin + t;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float X[3][8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    float gn[8][3][8];        // [gauss point][d][node]
    double W[8];
    for(int k = 0; k < 8; k++)
    {
        const float gda = (float)gp_geom(X, c_sh.Nxv + k, 64, 8, gn[k]);
        W[k] = (double)(1.0f * gda * EVI[(size_t)e * 8 + k]);             // float*float*float
    }
    int pair = 0;
    for(int a = 0; a < 8; a++)
        for(int b = a; b < 8; b++, pair++)
        {
            double bd[3][3];
            for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++) bd[i][j] = 0.0;
            for(int k = 0; k < 8; k++)
                for(int j = 0; j < 3; j++)
                    for(int i = 0; i < 3; i++)
                        bd[i][j] += W[k] * gn[k][j][a] * gn[k][i][b];
            double temp = 0.0;
            for(int k = 0; k < 8; k++)
                temp += W[k] * (gn[k][0][a] * gn[k][0][b] + gn[k][1][a] * gn[k][1][b] + gn[k][2][a] * gn[k][2][b]);   // float sum of float products
            bd[0][0] += temp; bd[1][1] += temp; bd[2][2] += temp;
            for(int i = 0; i < 3; i++)
                for(int j = 0; j < 3; j++) blocks[(size_t)(pair * 9 + 3 * i + j) * e_count + t] = bd[i][j];
        }
}

__host__ __device__ constexpr int pair_index(int a, int b) { return a * 8 - (a * (a - 1)) / 2 + (b - a); }   // a<=b, 0-based
// fixed slot of the neighbour at (dy,dx,dz): 0 self, 1..13 = CCU_LO order, -1 upper neighbour
__host__ __device__ constexpr int lo_slot(int dy, int dx, int dz)
{
    return (dy == 0 && dx == 0 && dz == 0) ? 0
         : (dy == -1) ? (1 + (dx + 1) * 3 + (dz + 1))
         : (dy == 0 && dx == -1) ? (10 + (dz + 1))
         : (dy == 0 && dx == 0 && dz == -1) ? 13 : -1;
}

// construct_node_ks + build_diagonal_of_K + get_aug_k (Construct_arrays.c:359-535, Element_calculations.c:624,1127):
// one thread per node gathers its <= 8 elements in ascending order; every half-matrix entry is a float
// accumulator exactly as Eqn_k1-3 are; BI accumulates the (augmented) element diagonals in double.
template <int OY, int OX, int OZ>
__device__ __forceinline__ void node_gather_element(const CcuGeom &g, const int e_rel, const int e_count, const double *__restrict__ blocks,
                                                    const float *__restrict__ gdel, const double visc_aug, const int use_aug,
                                                    const unsigned char fn, const unsigned char *__restrict__ flags, const int i, const int j,
                                                    const int k, float (&acc)[14][9], double (&diag)[3])
{
    constexpr int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    constexpr int OFFS[9][3] = CCU_OFFS_INIT;
    constexpr int a = LUT[OZ][OX][OY] - 1;            // local index of this node in the element
    const double w[3] = { (fn & CCU_F_VBX) ? 0.0 : 1.0, (fn & CCU_F_VBY) ? 0.0 : 1.0, (fn & CCU_F_VBZ) ? 0.0 : 1.0 };
#pragma unroll
    for(int b = 0; b < 8; b++)
    {
        const int dz = OFFS[b + 1][0] - OZ, dx = OFFS[b + 1][1] - OX, dy = OFFS[b + 1][2] - OY;
        const int slot = lo_slot(dy, dx, dz);
        if(slot < 0) continue;
        const unsigned char fb = flags[ccu_sidx(g, i + dy, j + dx, k + dz)];
        const double ww[3] = { (fb & CCU_F_VBX) ? 0.0 : 1.0, (fb & CCU_F_VBY) ? 0.0 : 1.0, (fb & CCU_F_VBZ) ? 0.0 : 1.0 };
        const int pr = (a <= b) ? pair_index(a, b) : pair_index(b, a);
#pragma unroll
        for(int r = 0; r < 3; r++)
#pragma unroll
            for(int cc = 0; cc < 3; cc++)
            {
                // elt_k[(a,r),(b,cc)]: stored block is bdbmu_{min,max}; transposed when a > b (:270-288)
                double v = (a <= b) ? blocks[(size_t)(pr * 9 + 3 * r + cc) * e_count + e_rel] : blocks[(size_t)(pr * 9 + 3 * cc + r) * e_count + e_rel];
                if(use_aug) v += visc_aug * gdel[3 * a + r] * gdel[3 * b + cc];
                acc[slot][3 * r + cc] = (float)((double)acc[slot][3 * r + cc] + w[r] * ww[cc] * v);
                if(slot == 0 && r == cc) diag[r] += v;
            }
    }
}

__global__ void __launch_bounds__(64) bk_node_ks(const CcuGeom g, const int i_begin, const int i_end, const int ey_begin, const int e_count,
                                                const double *__restrict__ blocks, const float *__restrict__ elt_del,
                                                const float *__restrict__ EVI, const unsigned char *__restrict__ flags,
                                                const int use_aug, const double augmented, float *K, double *BI, const int invert)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int plane = g.nox * g.noz;
    if(t >= (i_end - i_begin) * plane) return;
    const int i = i_begin + t / plane, rem = t % plane, j = rem / g.noz, k = rem % g.noz;
    const int s = ccu_sidx(g, i, j, k);
    const unsigned char fn = flags[s];
    float acc[14][9];
    double diag[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
    for(int q = 0; q < 14; q++)
#pragma unroll
        for(int r = 0; r < 9; r++) acc[q][r] = 0.0f;
#define GATHER(OY, OX, OZ) { const int ey = i - OY, ex = j - OX, ez = k - OZ; \
        if(ey >= 0 && ey < g.ely && ex >= 0 && ex < g.elx && ez >= 0 && ez < g.elz) { \
            const int e = ez + g.elz * (ex + g.elx * ey); \
            double va = 0.0; \
            if(use_aug) { for(int q = 0; q < 8; q++) va += EVI[(size_t)e * 8 + q]; va = va / 8; va = va * augmented; } \
            node_gather_element<OY, OX, OZ>(g, e - ey_begin * g.elx * g.elz, e_count, blocks, elt_del + (size_t)e * 24, va, use_aug, fn, flags, i, j, k, acc, diag); } }
    // ascending element number: ey = i-1 first (OY = 1), then ex = j-1 (OX = 1), then ez = k-1 (OZ = 1)
    GATHER(1, 1, 1) GATHER(1, 1, 0) GATHER(1, 0, 1) GATHER(1, 0, 0) GATHER(0, 1, 1) GATHER(0, 1, 0) GATHER(0, 0, 1) GATHER(0, 0, 0)
#undef GATHER
    const size_t NS = (size_t)g.NS;
#pragma unroll
    for(int q = 0; q < 14; q++)
#pragma unroll
        for(int r = 0; r < 9; r++) K[(size_t)(q * 9 + r) * NS + s] = acc[q][r];
    for(int d = 0; d < 3; d++) BI[(size_t)d * NS + s] = invert ? 1.0 / diag[d] : diag[d];
}

// build_diagonal_of_Ahat / assemble_dAhatp_entry (Element_calculations.c:654-685, 771-824)
__global__ void __launch_bounds__(128) bk_BPI(const CcuGeom g, const float *__restrict__ elt_del, const double *__restrict__ BI,
                                              const int precondition, double *BPI)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    if(!precondition) { BPI[e] = 1.0; return; }
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    const float *gd = elt_del + (size_t)e * 24;
    double gradP[24];
    for(int a = 1; a <= 8; a++)
    {
        const int s = ccu_sidx(g, ey + c_OFFS[a][2], ex + c_OFFS[a][1], ez + c_OFFS[a][0]);
        for(int d = 0; d < 3; d++) gradP[3 * (a - 1) + d] = 0.0 + BI[(size_t)d * g.NS + s] * gd[3 * (a - 1) + d];
    }
    double divU = 0.0;
    for(int p = 0; p < 24; p++) divU += gd[p] * gradP[p];
    BPI[e] = (divU != 0.0) ? 1.0 / divU : 1.0;
}

// assemble_forces + get_elt_f (Element_calculations.c:74-125, 989-1070), CART3D without imposed non-zero
// velocities: F(3n+2) = sum over elements (ascending) of sum_j force_at_gs[j] N(a,j) gDA[j] w[j]; stripped.
// Two passes: the element pass evaluates each element's eight nodal contributions once (a node-centred single pass
// recomputes every element's Jacobian determinants eight times: 13.5 ms at 256x256x128), the node pass adds them in
// ascending element order (ey, ex, ez) as the reference's element loop does.
__global__ void __launch_bounds__(128) bk_forces_elt(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ buoy, double *EF)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float X[3][8], gnx[3][8], force[8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    for(int q = 1; q <= 8; q++) force[q - 1] = buoy[elt_node(g, ey, ex, ez, q)];
    double ef[8];
    for(int a = 0; a < 8; a++) ef[a] = 0.0;
    for(int q = 0; q < 8; q++)
    {
        double fg = 0.0;
        for(int kk = 0; kk < 8; kk++) fg += (double)force[kk] * c_sh.Nv[8 * kk + q];
        const float gda = (float)gp_geom(X, c_sh.Nxv + q, 64, 8, gnx);
        for(int a = 0; a < 8; a++) ef[a] += fg * c_sh.Nv[8 * a + q] * gda * 1.0f;
    }
    for(int a = 0; a < 8; a++) EF[(size_t)a * g.nel + e] = ef[a];
}
__global__ void __launch_bounds__(128) bk_forces_gather(const CcuGeom g, const double *__restrict__ EF, const unsigned char *__restrict__ flags, double *F)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    double f = 0.0;
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int a = LUT[k - ez][j - ex][i - ey] - 1;
                f += EF[(size_t)a * g.nel + (ez + g.elz * (ex + g.elx * ey))];
            }
        }
    }
    const int s = ccu_sidx(g, i, j, k);
    const unsigned char fl = flags[s];
    F[s] = 0.0;
    F[(size_t)g.NS + s] = 0.0;
    F[2 * (size_t)g.NS + s] = (fl & CCU_F_VBZ) ? 0.0 : f;
}

// ---- imposed non-zero boundary velocities: the K.VB term of get_elt_f (Element_calculations.c:1038-1063)
// raw node bits of the reference (global_defs.h): VBX 0x2, VBZ 0x4, VBY 0x8
__device__ __forceinline__ unsigned vb_type(int d) { return d == 0 ? 0x2u : (d == 1 ? 0x8u : 0x4u); }
// marks the elements with a flagged node carrying a non-zero VB and hands each a slot (slot order is irrelevant: one slot, one element)
__global__ void __launch_bounds__(128) bk_vb_mark(const CcuGeom g, const unsigned *__restrict__ node, const float *__restrict__ VB1,
                                                   const float *__restrict__ VB2, const float *__restrict__ VB3, int *slot, int *elems, int *count)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    bool any = false;
    for(int a = 1; a <= 8; a++)
    {
        const int n = elt_node(g, ey, ex, ez, a);
        const unsigned f = node[n];
        any = any || ((f & 0x2u) && VB1[n] != 0.0f) || ((f & 0x8u) && VB2[n] != 0.0f) || ((f & 0x4u) && VB3[n] != 0.0f);
    }
    int sl = -1;
    if(any) { sl = atomicAdd(count, 1); if(elems) elems[sl] = e; }
    if(slot) slot[e] = sl;
}
__global__ void __launch_bounds__(128) bk_vb_slots(const int n_vb, const int *__restrict__ elems, int *slot)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t < n_vb) slot[elems[t]] = t;
}
// elt_f[p] -= elt_k[p][q] VB[q] over the flagged dofs q != p with VB != 0; elt_k as get_elt_k builds it (bk_elt_k above)
__global__ void __launch_bounds__(64) bk_forces_vb_elt(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ EVI,
                                                      const unsigned *__restrict__ node, const float *__restrict__ VB1,
                                                      const float *__restrict__ VB2, const float *__restrict__ VB3,
                                                      const int n_vb, const int *__restrict__ elems, double *EF, const float *__restrict__ SXX)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if(t >= n_vb) return;
    const int e = elems[t];
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    double vb[24], ef[24];
    if(SXX)
    {   // regional sphere: the element matrix of get_elt_k's Rsphere branch, point bases of the element itself (iconv = 1, :1102)
        SphElt S;
        sph_elt_setup(g, XX, SXX, EVI, e, false, S);
        for(int a = 0; a < 8; a++)
        {
            const int n = elt_node(g, ey, ex, ez, a + 1);
            const unsigned f = node[n];
            vb[3 * a + 0] = (f & 0x2u) ? (double)VB1[n] : 0.0;
            vb[3 * a + 1] = (f & 0x8u) ? (double)VB2[n] : 0.0;
            vb[3 * a + 2] = (f & 0x4u) ? (double)VB3[n] : 0.0;
        }
        for(int p = 0; p < 24; p++) ef[p] = 0.0;
        for(int a = 0; a < 8; a++)
            for(int b = a; b < 8; b++)
            {
                double bd[3][3];
                sph_elt_pair(S, a, b, bd);
                for(int i = 0; i < 3; i++)
                    for(int j = 0; j < 3; j++)
                    {
                        if(!(a == b && i == j)) ef[3 * a + i] -= bd[i][j] * vb[3 * b + j];
                        if(a != b) ef[3 * b + i] -= bd[j][i] * vb[3 * a + j];
                    }
            }
        for(int p = 0; p < 24; p++) EF[(size_t)p * n_vb + t] = ef[p];
        return;
    }
    float X[3][8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    float gn[8][3][8];
    double W[8];
    for(int k = 0; k < 8; k++)
    {
        const float gda = (float)gp_geom(X, c_sh.Nxv + k, 64, 8, gn[k]);
        W[k] = (double)(1.0f * gda * EVI[(size_t)e * 8 + k]);
    }
    for(int a = 0; a < 8; a++)
    {
        const int n = elt_node(g, ey, ex, ez, a + 1);
        const unsigned f = node[n];
        vb[3 * a + 0] = (f & 0x2u) ? (double)VB1[n] : 0.0;
        vb[3 * a + 1] = (f & 0x8u) ? (double)VB2[n] : 0.0;
        vb[3 * a + 2] = (f & 0x4u) ? (double)VB3[n] : 0.0;
    }
    for(int p = 0; p < 24; p++) ef[p] = 0.0;
    for(int a = 0; a < 8; a++)
        for(int b = a; b < 8; b++)
        {
            double bd[3][3];
            for(int i = 0; i < 3; i++) for(int j = 0; j < 3; j++) bd[i][j] = 0.0;
            for(int k = 0; k < 8; k++)
                for(int j = 0; j < 3; j++)
                    for(int i = 0; i < 3; i++)
                        bd[i][j] += W[k] * gn[k][j][a] * gn[k][i][b];
            double temp = 0.0;
            for(int k = 0; k < 8; k++)
                temp += W[k] * (gn[k][0][a] * gn[k][0][b] + gn[k][1][a] * gn[k][1][b] + gn[k][2][a] * gn[k][2][b]);
            bd[0][0] += temp; bd[1][1] += temp; bd[2][2] += temp;
            // elt_k[3a+i][3b+j] = bd[i][j], elt_k[3b+i][3a+j] = bd[j][i] (Element_calculations.c:268-288)
            for(int i = 0; i < 3; i++)
                for(int j = 0; j < 3; j++)
                {
                    if(!(a == b && i == j)) ef[3 * a + i] -= bd[i][j] * vb[3 * b + j];
                    if(a != b) ef[3 * b + i] -= bd[j][i] * vb[3 * a + j];
                }
        }
    for(int p = 0; p < 24; p++) EF[(size_t)p * n_vb + t] = ef[p];
}
// adds the elements' -K.VB contributions into the free dofs of F (ascending element order), after bk_forces_gather and before the
// interface sum
__global__ void __launch_bounds__(128) bk_forces_vb_gather(const CcuGeom g, const int *__restrict__ slot, const double *__restrict__ EF, const int n_vb,
                                                           const unsigned char *__restrict__ flags, double *F)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    double f[3] = { 0.0, 0.0, 0.0 };
    bool any = false;
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int sl = slot[ez + g.elz * (ex + g.elx * ey)];
                if(sl < 0) continue;
                const int a = LUT[k - ez][j - ex][i - ey] - 1;
                for(int d = 0; d < 3; d++) f[d] += EF[(size_t)(3 * a + d) * n_vb + sl];
                any = true;
            }
        }
    }
    if(!any) return;
    const int s = ccu_sidx(g, i, j, k);
    const unsigned char fl = flags[s];
    if(!(fl & CCU_F_VBX)) F[s] += f[0];
    if(!(fl & CCU_F_VBY)) F[(size_t)g.NS + s] += f[1];
    if(!(fl & CCU_F_VBZ)) F[2 * (size_t)g.NS + s] += f[2];
}
__global__ void __launch_bounds__(256) bk_strip_u(const CcuGeom g, const unsigned char *__restrict__ flags, double *U)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if(s >= g.NS) return;
    const unsigned char fl = flags[s];
    if(fl & CCU_F_VBX) U[s] = 0.0;
    if(fl & CCU_F_VBY) U[(size_t)g.NS + s] = 0.0;
    if(fl & CCU_F_VBZ) U[2 * (size_t)g.NS + s] = 0.0;
}
// velocities_conform_bcs (Boundary_conditions.c:993-1021) on the resident U
__global__ void __launch_bounds__(256) bk_conform_vbcs(const CcuGeom g, const unsigned *__restrict__ node, const float *__restrict__ VB1,
                                                       const float *__restrict__ VB2, const float *__restrict__ VB3, double *U)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const unsigned f = node[n];
    if(!(f & 0xeu)) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    const int s = ccu_sidx(g, i, j, k);
    if(f & 0x2u) U[s] = (double)VB1[n];
    if(f & 0x8u) U[(size_t)g.NS + s] = (double)VB2[n];
    if(f & 0x4u) U[2 * (size_t)g.NS + s] = (double)VB3[n];
}

// device K -> reference layout Eqn_k1-3 (test / drop-in read-back)
__global__ void bk_stiffness_to_ref(const CcuGeom g, const float *__restrict__ K, float *k1, float *k2, float *k3)
{
    const int LO[13][3] = CCU_LO_INIT;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    const int s = ccu_sidx(g, i, j, k);
    const size_t NS = (size_t)g.NS, base = (size_t)n * 42;
    float *kk[3] = { k1 + base, k2 + base, k3 + base };
    for(int q = 0; q < 42; q++) { kk[0][q] = 0.0f; kk[1][q] = 0.0f; kk[2][q] = 0.0f; }
    for(int a = 0; a < 3; a++) for(int b = 0; b < 3; b++) kk[a][b] = K[(size_t)(a * 3 + b) * NS + s];
    int rs = 0;
    for(int q = 0; q < 13; q++)
    {
        const int ii = i + LO[q][0], jj = j + LO[q][1], kz = k + LO[q][2];
        if(!(ii >= 0 && jj >= 0 && jj < g.nox && kz >= 0 && kz < g.noz)) continue;
        rs++;
        for(int a = 0; a < 3; a++) for(int b = 0; b < 3; b++) kk[a][3 * rs + b] = K[(size_t)((q + 1) * 9 + a * 3 + b) * NS + s];
    }
}
__global__ void bk_vec_to_nat(const CcuGeom g, const double *__restrict__ dev, double *nat)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    const int s = ccu_sidx(g, i, j, k);
    for(int d = 0; d < 3; d++) nat[3 * (size_t)n + d] = dev[(size_t)d * g.NS + s];
}

// ================================================================= energy step
#define CCU_TB_ANY (0x10u | 0x20u | 0x40u)     // TBX | TBZ | TBY (global_defs.h:65-89)
#define CCU_FBZ 0x100000u

// v_from_vector (Stokes_flow_Incomp.c:530-552): V[d][node] = (float) U[eq]; imposed velocities are zero here
__global__ void __launch_bounds__(128) ek_v_from_vector(const CcuGeom g, const double *__restrict__ U, float *V)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    const int s = ccu_sidx(g, i, j, k);
    for(int d = 0; d < 3; d++) V[(size_t)d * g.nno + n] = (float)U[(size_t)d * g.NS + s];
}

__device__ __forceinline__ void atomic_min_pos_float(float *addr, float v) { atomicMin((int *)addr, __float_as_int(v)); }   // v > 0
__device__ __forceinline__ void atomic_max_float(float *addr, float v)
{   // any sign: ints order like floats for >= 0, reversed for < 0
    if(v >= 0.0f) atomicMax((int *)addr, __float_as_int(v)); else atomicMin((unsigned *)addr, __float_as_uint(v));
}

// std_timestep (Advection_diffusion.c:737-810): mode 0 = min over elements of size^2 (diffusive limit, once),
// mode 1 = min over elements of 0.5 / (|uc1|/dx + |uc2|/dy + |uc3|/dz); float/double operand types as the reference
__global__ void __launch_bounds__(128) ek_timestep(const CcuGeom g, const float *__restrict__ eco, const float *__restrict__ V,
                                                   const int mode, float *red)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    float best = 1.0e8f;
    if(e < g.nel)
    {
        if(mode == 0)
        {
            for(int d = 0; d < 3; d++) { const float ts = eco[(size_t)e * 3 + d] * eco[(size_t)e * 3 + d]; if(best > ts) best = ts; }
        }
        else
        {
            const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
            float uc1 = 0.0f, uc2 = 0.0f, uc3 = 0.0f;
            for(int a = 1; a <= 8; a++)
            {
                const int n = elt_node(g, ey, ex, ez, a);
                uc1 = (float)((double)uc1 + c_sh.Np[a - 1] * (double)V[n]);
                uc2 = (float)((double)uc2 + c_sh.Np[a - 1] * (double)V[(size_t)g.nno + n]);
                uc3 = (float)((double)uc3 + c_sh.Np[a - 1] * (double)V[2 * (size_t)g.nno + n]);
            }
            const float uc = (float)(fabs((double)uc1) / (double)eco[(size_t)e * 3] + fabs((double)uc2) / (double)eco[(size_t)e * 3 + 1] +
                                     fabs((double)uc3) / (double)eco[(size_t)e * 3 + 2]);
            const float step = (float)(0.5 / (double)uc);
            best = fminf(best, step);
        }
    }
    for(int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if((threadIdx.x & 31) == 0) atomic_min_pos_float(red, best);
}
__global__ void ek_set_red(float *red, float vmin, float vmax) { red[0] = vmin; red[1] = vmax; }
__global__ void __launch_bounds__(256) ek_max(const int n, const float *__restrict__ T, float *red)
{
    float best = -10.0f;                      // Tmax (Global_operations.c:394)
    for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) best = fmaxf(T[i], best);
    for(int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if((threadIdx.x & 31) == 0) atomic_max_float(red + 1, best);
}

// predictor / corrector (Advection_diffusion.c:352-392)
__global__ void __launch_bounds__(256) ek_predictor(const int n, const unsigned *__restrict__ node, const float multiplier, float *field, float *fielddot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(!(node[i] & CCU_TB_ANY)) field[i] = field[i] + multiplier * fielddot[i];
    fielddot[i] = 0.0f;
}
__global__ void __launch_bounds__(256) ek_corrector(const int n, const unsigned *__restrict__ node, const float multiplier, float *field, float *fielddot,
                                                    const float *__restrict__ Dfielddot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(!(node[i] & CCU_TB_ANY)) field[i] = field[i] + multiplier * Dfielddot[i];
    fielddot[i] = fielddot[i] + Dfielddot[i];
}

// pg_shape_fn + element_residual (Advection_diffusion.c:448-557, 564-735), CART3D, no FBZ flux term: Eres[8] per element.
// The Petrov-Galerkin weights PG(j,i) = N(j,i) + adiff * (u_i . grad N_j) are formed on the fly per Gauss point
// (u_i is the same sum the reference forms twice, in pg_shape_fn and as v1..v3 in element_residual).
__global__ void __launch_bounds__(64) ek_element_residual(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ eco,
                                                          const unsigned *__restrict__ node, const float *__restrict__ T,
                                                          const float *__restrict__ Tdot, const float *__restrict__ V,
                                                          const float *__restrict__ diffusivity, const float Q0,
                                                          const float *__restrict__ heat_adi, const float *__restrict__ heat_visc,
                                                          const float *__restrict__ heat_latent, double *Eres, const int sph)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float X[3][8], gnx[3][8], vel[3][8];
    double Tn[8], DTn[8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    for(int a = 1; a <= 8; a++)
    {
        const int n = elt_node(g, ey, ex, ez, a);
        Tn[a - 1] = (double)T[n];
        DTn[a - 1] = (node[n] & CCU_TB_ANY) ? 0.0 : (double)Tdot[n];
        for(int d = 0; d < 3; d++) vel[d][a - 1] = V[(size_t)d * g.nno + n];
    }
    const float diff = (float)((double)(diffusivity[ez] + diffusivity[ez + 1]) * 0.5);
    // upwind parameter (pg_shape_fn :476-504)
    const double twodiff = 2.0 * (double)diff;
    double uc1 = 0.0, uc2 = 0.0, uc3 = 0.0;
    for(int a = 0; a < 8; a++)
    {
        uc1 += c_sh.Np[a] * (double)vel[0][a];
        uc2 += c_sh.Np[a] * (double)vel[1][a];
        uc3 += c_sh.Np[a] * (double)vel[2][a];
    }
    const double uxse = fabs(uc1 * (double)eco[(size_t)e * 3]), ueta = fabs(uc2 * (double)eco[(size_t)e * 3 + 1]),
                 ufai = fabs(uc3 * (double)eco[(size_t)e * 3 + 2]);
    const double xse = (uxse > twodiff) ? (1.0 - twodiff / uxse) : 0.0;
    const double eta = (ueta > twodiff) ? (1.0 - twodiff / ueta) : 0.0;
    const double fai = (ufai > twodiff) ? (1.0 - twodiff / ufai) : 0.0;
    const double unorm = uc1 * uc1 + uc2 * uc2 + uc3 * uc3;
    const double adiff = (unorm > 0.000001) ? ((uxse * xse + ueta * eta + ufai * fai) / (2.0 * unorm)) : 0.0;
    // Q = (rad_heat.total - heating_adi[el] + heating_visc[el]) * heating_latent[el]  (:643-647; the arrays hold 0, 0, 1
    // unless process_heating filled them)
    const float h_adi = heat_adi ? heat_adi[e] : 0.0f, h_visc = heat_visc ? heat_visc[e] : 0.0f, h_lat = heat_latent ? heat_latent[e] : 1.0f;
    const double Q = ((double)Q0 - (double)h_adi + (double)h_visc) * (double)h_lat;
    const bool diffusion = (diff != 0.0f);
    const float dl = diff * h_lat;                                                     // diff * heating_latent[el]  (float * float)
    double res[8];
    for(int j = 0; j < 8; j++) res[j] = 0.0;
    for(int i = 0; i < 8; i++)
    {
        const float gda = (float)gp_geom(X, c_sh.Nxv + i, 64, 8, gnx);
        // Rsphere (:506-528, 620-640, 668-674): gNX holds d/dtheta, d/dphi, d/dr; rtf3 = 1/r, rtf2 = 1/(r sin theta) scale the first two
        double rtf3 = 1.0, rtf2 = 1.0;
        if(sph)
        {
            double x[3], th, ph;
            sph_point(X, c_sh.Nv + i, 8, x);
            sph_rotate_gnx(x, gnx);
            sph_rtf(x, th, ph, rtf3);
            rtf2 = rtf3 / sin(th);
        }
        double dT = 0.0, tx1 = 0.0, tx2 = 0.0, tx3 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        for(int j = 0; j < 8; j++)
        {
            const double sfn = c_sh.Nv[8 * j + i];
            dT += DTn[j] * sfn;
            if(sph)
            {
                tx1 += (double)gnx[0][j] * Tn[j] * rtf3;
                tx2 += (double)gnx[1][j] * Tn[j] * rtf2;
            }
            else
            {
                tx1 += (double)gnx[0][j] * Tn[j];
                tx2 += (double)gnx[1][j] * Tn[j];
            }
            tx3 += (double)gnx[2][j] * Tn[j];
            v1 += (double)vel[0][j] * sfn;
            v2 += (double)vel[1][j] * sfn;
            v3 += (double)vel[2][j] * sfn;
        }
        const double adv = dT - Q + v1 * tx1 + v2 * tx2 + v3 * tx3;
        for(int j = 0; j < 8; j++)
        {
            const double prod1 = sph ? (v1 * (double)gnx[0][j] * rtf3 + v2 * (double)gnx[1][j] * rtf2 + v3 * (double)gnx[2][j])
                                     : (v1 * (double)gnx[0][j] + v2 * (double)gnx[1][j] + v3 * (double)gnx[2][j]);
            const double pg = c_sh.Nv[8 * j + i] + adiff * prod1;
            double term = pg * (double)gda * adv;
            if(diffusion)
                term = term + (double)(dl * gda) * (sph ? ((double)gnx[0][j] * tx1 * rtf3 + (double)gnx[1][j] * tx2 * rtf2 + (double)gnx[2][j] * tx3)
                                                        : ((double)gnx[0][j] * tx1 + (double)gnx[1][j] * tx2 + (double)gnx[2][j] * tx3));
            res[j] -= term;
        }
    }
    for(int j = 0; j < 8; j++) Eres[(size_t)e * 8 + j] = res[j];
}
// process_heating (Advection_diffusion.c:813-957), CART3D, without phase changes: per element
//   heating_visc = (Di/Atemp) * mean_gp(EVI) * 0.5 * (second invariant of the strain rate at the element centre)^2-sum
//                  (strain_rate_2_inv, Viscosity_structures.c:1043-1088, SQRT = 0),
//   heating_adi  = mean_nodes( Vz (T + Ts) Di ) * (expansivity[ez] + expansivity[ez+1]) / 2,   heating_latent = 1.
// Operand types as the reference: products of float operands are float, sums run in double.
__global__ void __launch_bounds__(64) ek_process_heating(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ EVI,
                                                         const float *__restrict__ T, const float *__restrict__ V,
                                                         const float *__restrict__ expansivity, const int adi_on, const int visc_on,
                                                         const float disptn, const float surf_temp, const float Atemp,
                                                         float *heat_adi, float *heat_visc, float *heat_latent, const int sph)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    const double temp1 = (double)(disptn / Atemp);                       // float division, kept in a double
    if(visc_on)
    {
        float X[3][8], gnx[3][8], VV[3][8];
        load_elt_coords(g, XX, ey, ex, ez, X);
        gp_geom(X, c_sh.Nxp, 8, 1, gnx);                                 // gNX[e].ppt: derivatives at the pressure point
        for(int a = 1; a <= 8; a++)
        {
            const int n = elt_node(g, ey, ex, ez, a);
            for(int d = 0; d < 3; d++) VV[d][a - 1] = V[(size_t)d * g.nno + n];
        }
        double dudx[3][3];
        for(int p = 0; p < 3; p++) for(int q = 0; q < 3; q++) dudx[p][q] = 0.0;
        for(int i = 0; i < 8; i++)
            for(int p = 0; p < 3; p++)
                for(int q = 0; q < 3; q++) dudx[p][q] += VV[p][i] * gnx[q][i];          // float * float
        double ed[3][3];
        for(int p = 0; p < 3; p++) for(int q = 0; q < 3; q++) ed[p][q] = 0.5 * (dudx[p][q] + dudx[q][p]);
        float eedot = (float)(ed[0][0] * ed[0][0] + ed[0][1] * ed[0][1] * 2.0 + ed[1][1] * ed[1][1] + ed[1][2] * ed[1][2] * 2.0 +
                              ed[2][2] * ed[2][2] + ed[0][2] * ed[0][2] * 2.0);
        if(sph) eedot = sph_strain2(X, VV, gnx);
        eedot = (float)((double)eedot * 0.5);
        double temp2 = 0.0;
        for(int i = 0; i < 8; i++) temp2 += EVI[(size_t)e * 8 + i];
        temp2 = temp2 / 8;
        heat_visc[e] = (float)(temp1 * temp2 * (double)eedot);
    }
    if(adi_on)
    {
        double temp2 = 0.0;
        for(int a = 1; a <= 8; a++)
        {
            const int n = elt_node(g, ey, ex, ez, a);
            temp2 = temp2 + (double)(V[2 * (size_t)g.nno + n] * (T[n] + surf_temp) * disptn);   // float expression
        }
        temp2 = temp2 / 8;
        heat_adi[e] = (float)(temp2 * (double)(expansivity[ez] + expansivity[ez + 1]) * 0.5);
    }
    (void)heat_latent;
}

__global__ void bk_fill_f32(const size_t n, const float v, float *x)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if(i < n) x[i] = v;
}
// phase_change (Phase_change.c:43-165), CART3D: Fas = 0.5 (1 + tanh(width * (z_phase - z - clapeyron (T - transT))))
struct CcuPhase { float zlm, z410, Ra670, clap670, width670, transT670, Ra410, clap410, width410, transT410; };
// transition temperatures: the layer-average temperature interpolated to the phase depth (:84-110); layer = sums | weights
__global__ void ek_phase_transT(const CcuGeom g, const float *__restrict__ XX, const double *__restrict__ layer, const float zlm, const float z410, float *transT)
{
    if(threadIdx.x > 1 || blockIdx.x) return;
    const float zp = threadIdx.x == 0 ? zlm : z410;
    const float *x3 = XX + 2 * (size_t)g.nno;                         // z of nodes 0 .. noz-1 (the first column)
    double temp1 = 0.0;
    for(int i = 0; i < g.noz - 1; i++)
        if(zp <= x3[i + 1] && zp >= x3[i])
        {
            const float H0 = (float)(layer[i] / layer[g.noz + i]), H1 = (float)(layer[i + 1] / layer[g.noz + i + 1]);
            temp1 = (double)(H0 + (H1 - H0) * (zp - x3[i]) / (x3[i + 1] - x3[i]));      // float expression
            break;
        }
    transT[threadIdx.x] = (float)temp1;
}
__global__ void ek_transT_to_tab(const float *__restrict__ transT, double *slot) { if(threadIdx.x < 2) slot[threadIdx.x] = (double)transT[threadIdx.x]; }
__global__ void ek_transT_from_tab(const double *__restrict__ slot, float *transT) { if(threadIdx.x < 2) transT[threadIdx.x] = (float)slot[threadIdx.x]; }
__global__ void __launch_bounds__(256) ek_phase_functions(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ T,
                                                          const CcuPhase ph, const float *__restrict__ transT, float *Fas670, float *Fas410)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const float z = XX[2 * (size_t)g.nno + n];
    double ep = (double)(ph.zlm - z - ph.clap670 * (T[n] - transT[0]));                 // float expression
    Fas670[n] = (float)(0.5 * (1.0 + tanh((double)ph.width670 * ep)));
    ep = (double)(ph.z410 - z - ph.clap410 * (T[n] - transT[1]));
    Fas410[n] = (float)(0.5 * (1.0 + tanh((double)ph.width410 * ep)));
}
__global__ void __launch_bounds__(256) ek_phase_buoyancy(const int nno, const float Ra670, const float Ra410, const float *__restrict__ Fas670,
                                                         const float *__restrict__ Fas410, float *buoy)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= nno) return;
    buoy[n] -= Ra670 * Fas670[n] + Ra410 * Fas410[n];
}
// latent-heating part of process_heating (Advection_diffusion.c:889-946): 670 first, then 410, then latent = 1 / latent
__global__ void __launch_bounds__(64) ek_latent_heating(const CcuGeom g, const float *__restrict__ T, const float *__restrict__ V,
                                                        const float *__restrict__ Fas670, const float *__restrict__ Fas410, const CcuPhase ph,
                                                        const float disptn, const float surf_temp, const float Atemp,
                                                        float *heat_adi, float *heat_latent)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float adi = heat_adi[e], lat = 1.0f;
    for(int which = 0; which < 2; which++)
    {
        const float Ra = which ? ph.Ra410 : ph.Ra670, clap = which ? ph.clap410 : ph.clap670, width = which ? ph.width410 : ph.width670;
        const float *Fas = which ? Fas410 : Fas670;
        if(Ra == 0.0f) continue;
        const double temp1 = 2.0 * width * clap * Ra / Atemp;
        double temp2 = 0.0, temp3 = 0.0;
        for(int a = 1; a <= 8; a++)
        {
            const int j = elt_node(g, ey, ex, ez, a);
            const float ts = T[j] + surf_temp;
            temp2 = temp2 + temp1 * (1.0 - Fas[j]) * Fas[j] * V[2 * (size_t)g.nno + j] * ts * disptn;
            temp3 = temp3 + temp1 * clap * (1.0 - Fas[j]) * Fas[j] * ts * disptn;
        }
        temp2 = temp2 / 8;
        temp3 = temp3 / 8;
        adi = (float)((double)adi + temp2);
        lat = (float)((double)lat + temp3);
    }
    heat_adi[e] = adi;
    heat_latent[e] = (float)(1.0 / lat);
}

// the scatter DTdot[node] += Eres[a] of pg_solver (:425-429) as a gather in ascending element order, float accumulator
__global__ void __launch_bounds__(128) ek_gather_residual(const CcuGeom g, const double *__restrict__ Eres, const float *__restrict__ MASS, float *DTdot)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    float acc = 0.0f;
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int e = ez + g.elz * (ex + g.elx * ey);
                acc = (float)((double)acc + Eres[(size_t)e * 8 + LUT[k - ez][j - ex][i - ey] - 1]);
            }
        }
    }
    DTdot[n] = MASS ? acc * MASS[n] : acc;
}

// thermal_buoyancy (thermal only) and the layer sums of return_horiz_ave (Global_operations.c:133-252).  For the
// rectangular faces of a Cartesian mesh the 2x2 Gauss quadrature of the bilinear interpolant is the trapezoid rule:
// node (j, k) of a layer carries the weight wx[j] * wy[k], w = half the sum of the adjacent spacings.
__global__ void __launch_bounds__(256) ek_buoyancy(const CcuGeom g, const float Atemp, const float *__restrict__ T,
                                                   const float *__restrict__ expansivity, float *buoy)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    buoy[n] = Atemp * T[n] * expansivity[n % g.noz];
}
__global__ void __launch_bounds__(256) ek_layer_sums(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ X, double *layer)
{
    const int kz = blockIdx.x;                       // one block per z layer
    double s = 0.0, w = 0.0;
    for(int t = threadIdx.x; t < g.nox * g.noy; t += blockDim.x)
    {
        const int j = t % g.nox, i = t / g.nox;
        const float *x1 = XX, *x2 = XX + g.nno;
        const int n0 = g.noz * (j + g.nox * i);      // node (i, j, 0): coordinates do not depend on z in a box
        double wx = 0.0, wy = 0.0;
        if(j > 0) wx += 0.5 * ((double)x1[n0] - (double)x1[n0 - g.noz]);
        if(j < g.nox - 1) wx += 0.5 * ((double)x1[n0 + g.noz] - (double)x1[n0]);
        if(i > 0) wy += 0.5 * ((double)x2[n0] - (double)x2[n0 - g.noz * g.nox]);
        if(i < g.noy - 1) wy += 0.5 * ((double)x2[n0 + g.noz * g.nox] - (double)x2[n0]);
        s += (double)X[n0 + kz] * wx * wy;
        w += wx * wy;
    }
    __shared__ double sh[2][8];
    for(int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); w += __shfl_down_sync(0xffffffffu, w, o); }
    if((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = w; }
    __syncthreads();
    if(threadIdx.x == 0)
    {
        double a = 0.0, b = 0.0;
        for(int q = 0; q < 8; q++) { a += sh[0][q]; b += sh[1][q]; }
        layer[kz] = a; layer[g.noz + kz] = b;
    }
}
// return_horiz_ave (Global_operations.c:133-215) as the reference integrates it, for regional-spherical meshes: layer kz sums, over the
// element columns, the 2 x 2 Gauss quadrature of the bilinear field on the BOTTOM face of element (min(kz, elz-1), ex, ey) -- the top
// layer borrows the last element's bottom-face areas, :185-199 -- with the surface Jacobian of get_global_1d_shape_fn's Rsphere branch
// (Size_does_matter.c:492-527: face nodes rotated into the frame of the element centre)
__global__ void __launch_bounds__(256) ek_layer_sums_sph(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ X, double *layer)
{
    const int kz = blockIdx.x;
    const int ez = kz < g.elz ? kz : g.elz - 1;
    const double B = (double)(float)0.57735026918962576451;
    const double gp[4][2] = { { -B, -B }, { B, -B }, { B, B }, { -B, B } };
    const int sx[4] = { -1, 1, 1, -1 }, sy[4] = { -1, -1, 1, 1 };          // master-element corners of face nodes 1..4
    double s = 0.0, w = 0.0;
    for(int t = threadIdx.x; t < g.elx * g.ely; t += blockDim.x)
    {
        const int ex = t % g.elx, ey = t / g.elx;
        float Xe[3][8];
        load_elt_coords(g, XX, ey, ex, ez, Xe);
        double centre[3];
        for(int i = 0; i < 3; i++)
        {
            double v = 0.0;
            for(int a = 0; a < 8; a++) v += Xe[i][a];
            centre[i] = v / 8;
        }
        const float c3 = (float)sqrt(centre[0] * centre[0] + centre[1] * centre[1] + centre[2] * centre[2]);
        const float to = (float)acos(centre[2] / c3), fo = (float)sph_myatan(centre[1], centre[0]);     // ECO.centre[1], [2] (float)
        const double ct = cos((double)to), st = sin((double)to), cf = cos((double)fo), sf = sin((double)fo);
        double xx[2][4];                                                  // the two in-surface coordinates of the four face nodes
        for(int i = 0; i < 4; i++)
        {
            xx[0][i] = Xe[0][i] * (ct * cf) + Xe[1][i] * (ct * sf) + Xe[2][i] * (-st);
            xx[1][i] = Xe[0][i] * (-sf) + Xe[1][i] * cf + Xe[2][i] * 0.0;
        }
        double val[4];
        {
            const int base = kz + g.noz * (ex + g.nox * ey);
            val[0] = (double)X[base]; val[1] = (double)X[base + g.noz]; val[2] = (double)X[base + g.noz + g.noz * g.nox]; val[3] = (double)X[base + g.noz * g.nox];
        }
        for(int k = 0; k < 4; k++)
        {
            double dxda[2][2] = { { 0.0, 0.0 }, { 0.0, 0.0 } }, M[4];
            for(int i = 0; i < 4; i++)
            {
                const double lx = 0.5 * (1.0 + sx[i] * gp[k][0]), ly = 0.5 * (1.0 + sy[i] * gp[k][1]);
                M[i] = lx * ly;
                const double mx0 = 0.5 * sx[i] * ly, mx1 = 0.5 * sy[i] * lx;      // Mx.vpt(0, i, k), Mx.vpt(1, i, k)
                dxda[0][0] += xx[0][i] * mx0; dxda[0][1] += xx[1][i] * mx0;
                dxda[1][0] += xx[0][i] * mx1; dxda[1][1] += xx[1][i] * mx1;
            }
            const double jac = dxda[0][0] * dxda[1][1] - dxda[0][1] * dxda[1][0];
            for(int d = 0; d < 4; d++) { s += val[d] * M[d] * jac; w += M[d] * jac; }
        }
    }
    __shared__ double sh[2][8];
    for(int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); w += __shfl_down_sync(0xffffffffu, w, o); }
    if((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = w; }
    __syncthreads();
    if(threadIdx.x == 0)
    {
        double a = 0.0, b = 0.0;
        for(int q = 0; q < 8; q++) { a += sh[0][q]; b += sh[1][q]; }
        layer[kz] = a; layer[g.noz + kz] = b;
    }
}
__global__ void __launch_bounds__(256) ek_remove_layer_ave(const CcuGeom g, const double *__restrict__ layer, float *X)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int kz = n % g.noz;
    if(layer[g.noz + kz] != 0.0) X[n] = X[n] - (float)(layer[kz] / layer[g.noz + kz]);
}

// ================================================================= diagnostics: heat_flux (Process_buoyancy.c:63-203), Nusselt numbers
// per element: uT = sum_gp (u_z T - kappa dT/dz) gDA / area  (:105-131), area = ECO.area (Size_does_matter.c:712)
__global__ void __launch_bounds__(64) hf_element(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ T, const float *__restrict__ V,
                                                 const float *__restrict__ diffusivity, double *uT_out, float *area_out, const int sph)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float X[3][8], gnx[3][8];
    double VZ[8], Tn[8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    for(int a = 1; a <= 8; a++)
    {
        const int n = elt_node(g, ey, ex, ez, a);
        VZ[a - 1] = (double)V[2 * (size_t)g.nno + n];
        Tn[a - 1] = (double)T[n];
    }
    const double diff = (double)((diffusivity[ez] + diffusivity[ez + 1]) * 0.5);
    double uT = 0.0, area = 0.0;
    for(int i = 0; i < 8; i++)
    {
        const float gda = (float)gp_geom(X, c_sh.Nxv + i, 64, 8, gnx);
        if(sph)
        {   // regional sphere: gNX holds d/dtheta, d/dphi, d/dr (get_global_shape_fn's sphere branch), V[3] is the radial velocity
            double x[3];
            sph_point(X, c_sh.Nv + i, 8, x);
            sph_rotate_gnx(x, gnx);
        }
        double u = 0.0, Tg = 0.0, dTdz = 0.0;
        for(int j = 0; j < 8; j++)
        {
            u += VZ[j] * c_sh.Nv[8 * j + i];
            Tg += Tn[j] * c_sh.Nv[8 * j + i];
            dTdz = dTdz + Tn[j] * (double)gnx[2][j];
        }
        uT = uT + (u * Tg - diff * dTdz) * (double)gda;
        area += 1.0 * (double)gda;
    }
    const float areaf = (float)area;
    uT /= (double)areaf;
    uT_out[e] = uT;
    area_out[e] = areaf;
}
// heatflux[node] += TWW * uT in ascending element order (float accumulator, :133-138); * Mass unless halo sums come first
__global__ void __launch_bounds__(128) hf_nodal(const CcuGeom g, const float *__restrict__ TWW, const float *__restrict__ MASS,
                                                const double *__restrict__ uT, float *hf)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    float acc = 0.0f;
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int e = ez + g.elz * (ex + g.elx * ey);
                acc = (float)((double)acc + (double)TWW[(size_t)e * 8 + LUT[k - ez][j - ex][i - ey] - 1] * uT[e]);
            }
        }
    }
    hf[n] = MASS ? acc * MASS[n] : acc;
}
// surface / bottom extrapolation (:151-158) and the area-weighted sums over the surface elements (:162-176): out[0..3] =
// hfb, areab, hft, areat of this subdomain (zero where it does not touch the bottom / top of the box)
__global__ void __launch_bounds__(256) hf_surface_sums(const CcuGeom g, const float *__restrict__ hf, const float *__restrict__ area,
                                                       const int at_bottom, const int at_top, double *out)
{
    double s[4] = { 0.0, 0.0, 0.0, 0.0 };
    for(int t = blockIdx.x * blockDim.x + threadIdx.x; t < g.elx * g.ely; t += gridDim.x * blockDim.x)
    {
        const int ex = t % g.elx, ey = t / g.elx;
        double tempb = 0.0, tempt = 0.0;
        for(int q = 0; q < 4; q++)
        {
            const int j = ex + (q & 1), i = ey + (q >> 1);
            const size_t col = (size_t)g.noz * (j + (size_t)g.nox * i);
            const float sh = 2 * hf[col + g.noz - 1] - hf[col + g.noz - 2];
            const float bh = 2 * hf[col] - hf[col + 1];
            tempt += (double)sh; tempb += (double)bh;
        }
        const size_t eb = (size_t)g.elz * (ex + (size_t)g.elx * ey), et = eb + g.elz - 1;
        if(at_bottom) { s[0] += tempb * (double)area[eb]; s[1] += (double)area[eb]; }
        if(at_top) { s[2] += tempt * (double)area[et]; s[3] += (double)area[et]; }
    }
    __shared__ double sh[4][8];
    for(int q = 0; q < 4; q++)
    {
        double v = s[q];
        for(int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if(threadIdx.x < 4)
    {
        double v = 0.0;
        for(int w = 0; w < 8; w++) v += sh[threadIdx.x][w];
        out[threadIdx.x] = v;
    }
}

// ================================================================= markers
#define CCU_SIDEE 0x800000u
struct MkGrid
{
    const double *XP1, *XP2, *XP3;   // 1-D node coordinates (0-based here: XP[d][i] of the reference is XPd[i-1])
    const int *RG3;                   // RG[3][0..rnoz]
    double dx, dy, dzz;
    int rnoz;
};
// get_element (Composition_adv.c:1086-1180): uniform spacing in x and y, table lookup in z; returns the 1-based element
// number (0 if the marker fell out of the z table); dX = offsets from the element's first node
__device__ __forceinline__ int mk_get_element(const CcuGeom &g, const MkGrid &m, const double x1, const double x2, const double x3, double dX[3])
{
    double t = (x1 - m.XP1[0]) / m.dx + 1;
    const int IX = (int)(t < (double)g.elx ? t : (double)g.elx);
    dX[0] = x1 - m.XP1[IX - 1];
    t = (x2 - m.XP2[0]) / m.dy + 1;
    const int IY = (int)(t < (double)g.ely ? t : (double)g.ely);
    dX[1] = x2 - m.XP2[IY - 1];
    t = (x3 - m.XP3[0]) / m.dzz + 1;
    const int i1 = (int)(t < (double)(m.rnoz - 1) ? t : (double)(m.rnoz - 1));
    int IZ = m.RG3[i1];
    dX[2] = 0.0;
    if(IZ) dX[2] = x3 - m.XP3[IZ - 1];
    else
        for(int i = 0; i <= 2; i += 2)
        {
            const int j1 = m.RG3[i1 - 1 + i];
            if(j1 >= 1 && j1 <= g.elz && x3 >= m.XP3[j1 - 1] && x3 <= m.XP3[j1]) { IZ = j1; dX[2] = x3 - m.XP3[IZ - 1]; }
        }
    if(!IZ) return 0;
    return IZ + (IX - 1) * g.elz + (IY - 1) * g.elz * g.elx;
}
// velocity_markers (Composition_adv.c:990-1075): trilinear interpolation of the nodal velocity at XMC (con 0 -> VO) or
// XMCpred (con 1 -> Vpred); also records the element
__global__ void __launch_bounds__(128) mk_velocity(const CcuGeom g, const MkGrid m, const int n, const int cap, const double *__restrict__ X,
                                                   const float *__restrict__ eco, const float *__restrict__ V, float *Vout, int *CElement, int *err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double dX[3];
    const int el = mk_get_element(g, m, X[i], X[(size_t)cap + i], X[2 * (size_t)cap + i], dX);
    if(!el) { atomicAdd(err, 1); return; }
    const int e = el - 1;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    const double sz = (double)eco[(size_t)e * 3 + 2];
    const double w[8] = { (m.dx - dX[0]) * (m.dy - dX[1]) * (sz - dX[2]), dX[0] * (m.dy - dX[1]) * (sz - dX[2]), dX[0] * dX[1] * (sz - dX[2]),
                          (m.dx - dX[0]) * dX[1] * (sz - dX[2]), (m.dx - dX[0]) * (m.dy - dX[1]) * dX[2], dX[0] * (m.dy - dX[1]) * dX[2],
                          dX[0] * dX[1] * dX[2], (m.dx - dX[0]) * dX[1] * dX[2] };
    const double area = m.dx * m.dy * sz;
    int nd[8];
    for(int a = 1; a <= 8; a++) nd[a - 1] = elt_node(g, ey, ex, ez, a);
    for(int d = 0; d < 3; d++)
    {
        const float *Vd = V + (size_t)d * g.nno;
        double s = w[0] * (double)Vd[nd[0]];
        for(int a = 1; a < 8; a++) s = s + w[a] * (double)Vd[nd[a]];
        Vout[(size_t)d * cap + i] = (float)(s / area);
    }
    CElement[i] = el;
}
// Euler predictor / modified-Euler corrector of the positions (Composition_adv.c:115-121, 71-77): dt*VO is a FLOAT product
__global__ void __launch_bounds__(256) mk_advance(const int n, const int cap, const float dt, const int corrector, const float *__restrict__ VO,
                                                  const float *__restrict__ Vpred, double *X, double *Xpred, const int sph)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(sph)
    {   // Rsphere (Composition_adv.c:79-92, 124-135): positions are (theta, phi, r); u_theta / r, u_phi / (r sin theta), all from the OLD position
        const size_t q0 = i, q1 = (size_t)cap + i, q2 = 2 * (size_t)cap + i;
        const double t = X[q0], f = X[q1], r = X[q2];
        if(!corrector)
        {
            Xpred[q0] = t + (double)(dt * VO[q0]) / r;
            Xpred[q1] = f + (double)(dt * VO[q1]) / (r * sin(t));
            Xpred[q2] = r + (double)(dt * VO[q2]);
        }
        else
        {
            X[q0] = t + 0.5 * (double)dt * (double)(VO[q0] + Vpred[q0]) / r;
            X[q1] = f + 0.5 * (double)dt * (double)(VO[q1] + Vpred[q1]) / (r * sin(t));
            X[q2] = r + 0.5 * (double)dt * (double)(VO[q2] + Vpred[q2]);
        }
        return;
    }
    for(int d = 0; d < 3; d++)
    {
        const size_t q = (size_t)d * cap + i;
        if(!corrector) Xpred[q] = X[q] + (double)(dt * VO[q]);
        else X[q] = X[q] + 0.5 * (double)dt * (double)(VO[q] + Vpred[q]);
    }
}
// move_tracers_to_neighbors (Composition_adv.c:218-400), one subdomain: markers of boundary (SIDEE) elements are clamped
// into [XG1, XG2]
__global__ void __launch_bounds__(256) mk_clamp(const int n, const int cap, const unsigned *__restrict__ Element, const int *__restrict__ CElement,
                                                const double g1x, const double g1y, const double g1z, const double g2x, const double g2y,
                                                const double g2z, double *X)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(!(Element[CElement[i] - 1] & CCU_SIDEE)) return;
    const double lo[3] = { g1x, g1y, g1z }, hi[3] = { g2x, g2y, g2z };
    for(int d = 0; d < 3; d++)
    {
        double v = X[(size_t)d * cap + i];
        v = v < hi[d] ? v : hi[d];
        v = v > lo[d] ? v : lo[d];
        X[(size_t)d * cap + i] = v;
    }
}
// ---- markers changing subdomain (transfer_markers_processors, Composition_adv.c:148-226)
// locate_processor (:628-669) for the markers of side elements (move_tracers_to_neighbors :382-420): the subdomain one step
// up / down per axis when the (clamped) position lies beyond the local mesh.  code = (ox+1) + 3 (oy+1) + 9 (oz+1); 13 = stays.
#define CCU_MK_REC 8             // doubles per migrating marker: XMC[3], XMCpred[3], {VO0, VO1}, {VO2, C12}
__global__ void __launch_bounds__(256) mk_dest(const int n, const int cap, const unsigned *__restrict__ Element, const int *__restrict__ CElement,
                                               const double *__restrict__ X, const double lo0, const double lo1, const double lo2,
                                               const double hi0, const double hi1, const double hi2, const int me0, const int me1, const int me2,
                                               const int np0, const int np1, const int np2, unsigned char *code, unsigned char *stay)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    int c = 13;
    if(Element[CElement[i] - 1] & CCU_SIDEE)
    {
        const double lo[3] = { lo0, lo1, lo2 }, hi[3] = { hi0, hi1, hi2 };
        const int me[3] = { me0, me1, me2 }, np[3] = { np0, np1, np2 };
        int o[3];
        for(int d = 0; d < 3; d++)
        {
            const double v = X[(size_t)d * cap + i];
            int m = me[d];
            if(v > hi[d]) m = min(np[d] - 1, me[d] + 1);
            else if(v < lo[d]) m = max(0, me[d] - 1);
            o[d] = m - me[d];
        }
        c = (o[0] + 1) + 3 * (o[1] + 1) + 9 * (o[2] + 1);
    }
    code[i] = (unsigned char)c;
    stay[i] = (unsigned char)(c == 13);
}
// leavers (perm[n_stay .. n) in reverse order of their original index) -> (index, code) pairs for the host's stable sort
__global__ void mk_leaver_list(const int n, const int n_stay, const int *__restrict__ perm, const unsigned char *__restrict__ code, int *lv_idx, int *lv_code)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= n - n_stay) return;
    const int i = perm[n - 1 - q];
    lv_idx[q] = i; lv_code[q] = code[i];
}
__global__ void mk_pack(const int nl, const int cap, const int *__restrict__ idx, const double *__restrict__ X, const double *__restrict__ Xpred,
                        const float *__restrict__ VO, const int *__restrict__ C12, double *buf)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nl) return;
    const int i = idx[q];
    double *r = buf + (size_t)q * CCU_MK_REC;
    for(int d = 0; d < 3; d++) { r[d] = X[(size_t)d * cap + i]; r[3 + d] = Xpred[(size_t)d * cap + i]; }
    float *f = (float *)(r + 6);
    f[0] = VO[i]; f[1] = VO[(size_t)cap + i]; f[2] = VO[2 * (size_t)cap + i];
    ((int *)f)[3] = C12[i];
}
__global__ void mk_unpack(const int nr, const int cap, const int first, const double *__restrict__ buf, double *X, double *Xpred, float *VO, int *C12, int *CElement)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= nr) return;
    const int i = first + q;
    const double *r = buf + (size_t)q * CCU_MK_REC;
    for(int d = 0; d < 3; d++) { X[(size_t)d * cap + i] = r[d]; Xpred[(size_t)d * cap + i] = r[3 + d]; }
    const float *f = (const float *)(r + 6);
    VO[i] = f[0]; VO[(size_t)cap + i] = f[1]; VO[2 * (size_t)cap + i] = f[2];
    C12[i] = ((const int *)f)[3];
    CElement[i] = 1;                 // element_markers assigns it from the position next
}
// stayers keep their relative order: arrays gathered through the partition's permutation
template <class T, int ND>
__global__ void mk_gather(const int n_stay, const int cap, const int *__restrict__ perm, const T *__restrict__ src, T *dst)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if(q >= n_stay) return;
    const int i = perm[q];
#pragma unroll
    for(int d = 0; d < ND; d++) dst[(size_t)d * cap + q] = src[(size_t)d * cap + i];
}

// element_markers (Composition_adv.c:960-985) + the per-element counts of get_C_from_markers (:757-760)
__global__ void __launch_bounds__(128) mk_assign_count(const CcuGeom g, const MkGrid m, const int n, const int cap, const double *__restrict__ X,
                                                       const int *__restrict__ C12, int *CElement, int *count, int *err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double dX[3];
    const int el = mk_get_element(g, m, X[i], X[(size_t)cap + i], X[2 * (size_t)cap + i], dX);
    if(!el) { atomicAdd(err, 1); return; }
    CElement[i] = el;
    atomicAdd(count + (size_t)C12[i] * g.nel + (el - 1), 1);
}
// ratio method (Composition_adv.c:786-803): CE = dense / (regular + dense), elements without markers keep their CE
__global__ void __launch_bounds__(256) mk_element_C(const int nel, const int *__restrict__ count, float *CE)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= nel) return;
    const int c0 = count[e], c1 = count[(size_t)nel + e];
    if(c0 || c1) { const float t0 = (float)c0, t1 = (float)c1; CE[e] = t1 / (t0 + t1); }
}
// C[node] = Mass * sum over its elements (ascending) of TWW * CE   (:797-811), float accumulator
__global__ void __launch_bounds__(128) mk_nodal_C(const CcuGeom g, const float *__restrict__ TWW, const float *__restrict__ MASS,
                                                  const float *__restrict__ CE, float *C)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    float acc = 0.0f;
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int e = ez + g.elz * (ex + g.elx * ey);
                acc = acc + TWW[(size_t)e * 8 + LUT[k - ez][j - ex][i - ey] - 1] * CE[e];
            }
        }
    }
    C[n] = MASS ? acc * MASS[n] : acc;
}
// thermo-chemical buoyancy (Pan_problem_misc_functions.c:125-128): Atemp*T*expansivity - Acomp*C
__global__ void __launch_bounds__(256) ek_buoyancy_comp(const CcuGeom g, const float Atemp, const float Acomp, const float *__restrict__ T,
                                                        const float *__restrict__ C, const float *__restrict__ expansivity, float *buoy)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    buoy[n] = Atemp * T[n] * expansivity[n % g.noz] - Acomp * C[n];
}

// ================================================================= host side
static bool g_tables_ready = false;
static int ensure_tables(ccu_ctx *c)
{
    (void)c;
    if(g_tables_ready) return 0;
    ShapeTables t;
    make_shape_tables(t);
    CK(cudaMemcpyToSymbol(c_sh, &t, sizeof(t)));
    g_tables_ready = true;
    return 0;
}

int ccu_set_coordinates(ccu_ctx *c, int lev, const float *X1, const float *X2, const float *X3)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    const size_t n = (size_t)L.g.nno;
    if(!L.XX) CK(cudaMalloc(&L.XX, sizeof(float) * 3 * n));
    CK(cudaMemcpyAsync(L.XX, X1, sizeof(float) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(L.XX + n, X2, sizeof(float) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(L.XX + 2 * n, X3, sizeof(float) * n, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    L.have_xx = true;
    return 0;
}

// E->SXX[lev][1..3] (theta, phi, r): switches the context to regional-spherical geometry (E->control.Rsphere); ccu_set_coordinates
// then carries the CARTESIAN coordinates of the nodes (E->XX), as in the reference.  Every level needs both before ccu_build_geometry.
int ccu_set_spherical_coordinates(ccu_ctx *c, int lev, const float *S1, const float *S2, const float *S3)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    const size_t n = (size_t)L.g.nno;
    if(!L.SXX) CK(cudaMalloc(&L.SXX, sizeof(float) * 3 * n));
    CK(cudaMemcpyAsync(L.SXX, S1, sizeof(float) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(L.SXX + n, S2, sizeof(float) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(L.SXX + 2 * n, S3, sizeof(float) * n, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    L.have_sxx = true;
    c->rsphere = true;
    return 0;
}
int ccu_build_geometry(ccu_ctx *c)
{
    if(!c) FAIL("null context");
    if(ensure_tables(c)) return 1;
    for(int lev = c->cfg.levmin; lev <= c->cfg.levmax; lev++)
    {
        Level &L = c->L[lev];
        if(!L.have_xx) FAIL("build_geometry: coordinates missing");
        if(ccu_ensure_stage(c, sizeof(double) * 8 * (size_t)L.g.nel)) return 1;
        LAUNCH(c, bk_elt_geometry, cdiv(L.g.nel, 128), 128, L.g, L.XX, L.TWW, (double *)c->stage, L.eco, L.elt_del);
        if(c->rsphere)
        {
            if(!L.have_sxx) FAIL("build_geometry: spherical coordinates missing on a level (ccu_set_spherical_coordinates)");
            LAUNCH(c, bk_elt_geometry_sph, cdiv(L.g.nel, 128), 128, L.g, L.XX, L.SXX, L.eco, L.elt_del);
        }
        ccu_elt_del_changed(c, lev);
        ccu_eco_changed(c, lev);
        LAUNCH(c, bk_mass, cdiv(L.g.nno, 128), 128, L.g, (const double *)c->stage, L.MASS);
        if(ccu_halo_sum_nodal(c, lev, L.MASS)) return 1;                  // exchange_node_f20 (Size_does_matter.c:733)
        LAUNCH(c, bk_invert, cdiv(L.g.nno, 128), 128, L.g.nno, L.MASS);
        SYNC(c);
        L.have_tw = true;
    }
    SYNC(c);
    return 0;
}

int ccu_set_viscosity_law(ccu_ctx *c, int tdepv, int rheol, int num_mat, const float *N0, const float *E, const float *T, const float *Z,
                          int vmin, float min_value, int vmax, float max_value, int smooth_cycles)
{
    if(!c) FAIL("null context");
    if(num_mat < 1 || num_mat > 40) FAIL("bad num_mat");
    if(tdepv && !(rheol == 0 || rheol == 1 || rheol == 2 || rheol == 3 || rheol == 4 || rheol == 10 || rheol == 11))
        FAIL("viscosity law: RHEOL option undefined in TDEPV (the reference knows 0, 1, 2, 3, 4, 10, 11)");
    if(smooth_cycles < 0 || smooth_cycles > 3) FAIL("project_viscosity: visc_smooth_cycles must be 0 .. 3");
    CcuViscParams &v = c->visc;
    v.tdepv = tdepv; v.rheol = rheol; v.num_mat = num_mat; v.vmin = vmin; v.vmax = vmax; v.min_value = min_value; v.max_value = max_value;
    v.smooth_cycles = smooth_cycles;
    for(int i = 0; i < num_mat; i++) { v.N0[i] = N0[i]; v.E[i] = E[i]; v.T[i] = T[i]; v.Z[i] = Z[i]; }
    return 0;
}

// E->viscosity.{BDEPV, abyerlee, bbyerlee, lbyerlee, plasticity_dimensional, plasticity_trans, plasticity_viscosity_offset},
// E->monitor.{length_scale, tau_scale} (Viscosity_structures.c:69-99, 186-200): the regular branch of visc_from_B.  Iterated with the
// velocity like SDEPV (need_to_iterate), with the same misfit / damping / iteration cap (ccu_set_sdepv carries those; on = 0 there is fine).
int ccu_set_bdepv(ccu_ctx *c, int on, const float *abyerlee, const float *bbyerlee, const float *lbyerlee, int dimensional, float length_scale,
                  float tau_scale, int plasticity_trans, float viscosity_offset)
{
    if(!c) FAIL("null context");
    CcuViscParams &v = c->visc;
    v.bdepv = on != 0; v.bdepv_visits = 0;
    if(!on) return 0;
    if(!abyerlee || !bbyerlee || !lbyerlee) FAIL("set_bdepv: yield-stress parameters missing");
    for(int i = 0; i < v.num_mat && i < 40; i++) { v.abyerlee[i] = abyerlee[i]; v.bbyerlee[i] = bbyerlee[i]; v.lbyerlee[i] = lbyerlee[i]; }
    v.bdepv_dimensional = dimensional != 0; v.bdepv_ndz_to_m = length_scale; v.bdepv_tau_scale = tau_scale;
    v.bdepv_trans = plasticity_trans != 0; v.bdepv_offset = viscosity_offset;
    return 0;
}
// E->viscosity.{CDEPV, layer_pre_comp, pre_comp, cdepv_absolute}, E->control.check_c_irange (Viscosity_structures.c:178-281): pre_comp holds
// 2 * num_mat values with layer_pre_comp, else 2.  The flavour / lithosphere / crust variants of visc_from_C are not implemented.
int ccu_set_cdepv(ccu_ctx *c, int on, int layer_pre_comp, const float *pre_comp, int absolute, int check_c_irange)
{
    if(!c) FAIL("null context");
    CcuViscParams &v = c->visc;
    v.cdepv = on != 0; v.cdepv_layer = layer_pre_comp != 0; v.cdepv_absolute = absolute != 0; v.cdepv_check_range = check_c_irange != 0;
    if(!on) return 0;
    if(!pre_comp) FAIL("set_cdepv: pre_comp missing");
    const int n = layer_pre_comp ? 2 * v.num_mat : 2;
    if(n > 80) FAIL("set_cdepv: too many material layers");
    for(int i = 0; i < n; i++)
    {
        if(!(pre_comp[i] > 0.0f)) FAIL("set_cdepv: pre_comp must be positive");
        v.cdepv_logv[i] = log((double)pre_comp[i]);
    }
    return 0;
}
// E->C (nodal composition, [nno] in the reference's node order) for hosts that keep the markers themselves
int ccu_set_composition(ccu_ctx *c, const float *C)
{
    if(!c || !C) FAIL("set_composition: null argument");
    Level &L = c->L[c->cfg.levmax];
    if(!c->Cnode) CK(cudaMalloc(&c->Cnode, sizeof(float) * (size_t)L.g.nno));
    CK(cudaMemcpyAsync(c->Cnode, C, sizeof(float) * (size_t)L.g.nno, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    return 0;
}

// E->viscosity.{SDEPV, sdepv_rheology, sdepv_expt, sdepv_trns, sdepv_misfit, sdepv_iter_damp, sdepv_start_from_newtonian, sdepv_trns_T,
// sdepv_trns_c}, E->monitor.max_sdep_visc_iter (Viscosity_structures.c:150-310)
int ccu_set_sdepv(ccu_ctx *c, int on, int rheology, const float *expt, const float *trns, float misfit, float iter_damp, int max_iter,
                  int start_from_newtonian, float trns_T, float trns_c)
{
    if(!c) FAIL("null context");
    CcuViscParams &v = c->visc;
    if(on && rheology != 1 && rheology != 2) FAIL("stress-dependent viscosity: sdepv_rheology 1 and 2 are implemented on the device (3, the dimensional Arrhenius law, is not)");
    v.sdepv = on != 0; v.sdepv_rheology = rheology; v.sdepv_misfit = misfit; v.sdepv_iter_damp = iter_damp; v.sdepv_max_iter = max_iter;
    v.sdepv_start_from_newtonian = start_from_newtonian; v.sdepv_trns_T = trns_T; v.sdepv_trns_c = trns_c; v.sdepv_visits = 0;
    for(int i = 0; i < v.num_mat && i < 40; i++) { v.sdepv_expt[i] = expt ? expt[i] : 1.0f; v.sdepv_trns[i] = trns ? trns[i] : 1.0f; }
    return 0;
}

int ccu_set_material(ccu_ctx *c, const int *mat)
{
    if(!c) FAIL("null context");
    Level &L = c->L[c->cfg.levmax];
    if(!c->mat) CK(cudaMalloc(&c->mat, sizeof(int) * (size_t)L.g.nel));
    CK(cudaMemcpyAsync(c->mat, mat, sizeof(int) * (size_t)L.g.nel, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    return 0;
}

static int ensure_nodal(ccu_ctx *c)
{
    Level &L = c->L[c->cfg.levmax];
    if(!c->T) CK(cudaMalloc(&c->T, sizeof(float) * (size_t)L.g.nno));
    if(!c->buoy) CK(cudaMalloc(&c->buoy, sizeof(float) * (size_t)L.g.nno));
    if(!c->nodal_tmp) CK(cudaMalloc(&c->nodal_tmp, sizeof(float) * (size_t)L.g.nno));
    if(!c->nodal_tmp2) CK(cudaMalloc(&c->nodal_tmp2, sizeof(float) * (size_t)L.g.nno));
    for(int lev = c->cfg.levmin; lev <= c->cfg.levmax; lev++)
        if(!c->L[lev].EVI) CK(cudaMalloc(&c->L[lev].EVI, sizeof(float) * 8 * (size_t)c->L[lev].g.nel));
    return 0;
}

int ccu_set_temperature(ccu_ctx *c, const float *T)
{
    if(!c) FAIL("null context");
    if(ensure_nodal(c)) return 1;
    CK(cudaMemcpyAsync(c->T, T, sizeof(float) * (size_t)c->L[c->cfg.levmax].g.nno, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    return 0;
}

int ccu_set_element_viscosity(ccu_ctx *c, int lev, const float *EVI)
{
    if(ccu_check_lev(c, lev)) return 2;
    if(ensure_nodal(c)) return 1;
    Level &L = c->L[lev];
    CK(cudaMemcpyAsync(L.EVI, EVI, sizeof(float) * 8 * (size_t)L.g.nel, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    L.have_evi = true;
    return 0;
}

// get_system_viscosity (Viscosity_structures.c:369) at the finest level from the resident temperature
int ccu_get_system_viscosity(ccu_ctx *c)
{
    if(!c) FAIL("null context");
    if(ensure_tables(c) || ensure_nodal(c)) return 1;
    if(!c->mat) FAIL("get_system_viscosity: material groups missing");
    Level &L = c->L[c->cfg.levmax];
    if(c->visc.tdepv && (c->visc.rheol == 2 || c->visc.rheol == 4) && !L.have_xx) FAIL("get_system_viscosity: depth-dependent law needs the node coordinates");
    const bool sd = c->visc.sdepv != 0, cd = c->visc.cdepv != 0, bd = c->visc.bdepv != 0;
    const bool post = sd || cd || bd;                  // anything between the temperature law and the min / max clip
    const float *Ccomp = c->mk.ready ? (const float *)c->mk.C : (const float *)c->Cnode;
    if(cd && !Ccomp) FAIL("get_system_viscosity: composition-dependent viscosity needs the nodal composition (device markers or ccu_set_composition)");
    // the depth coordinate of laws 2 and 4 and of the yield stress: z of the box, r of the regional sphere (Xtmp = E->SX, Viscosity_structures.c:525-530)
    const float *zco = c->rsphere ? (L.SXX ? L.SXX + 2 * (size_t)L.g.nno : (const float *)nullptr) : (L.XX ? L.XX + 2 * (size_t)L.g.nno : (const float *)nullptr);
    if(((c->visc.tdepv && (c->visc.rheol == 2 || c->visc.rheol == 4)) || bd) && !zco) FAIL("get_system_viscosity: depth-dependent law needs the node coordinates");
    // get_system_viscosity's order (Viscosity_structures.c:386-425): temperature law, stress, composition, plastic yielding, then the clip
    LAUNCH(c, bk_visc, cdiv(L.g.nel, 128), 128, L.g, c->visc, c->mat, c->T, zco, L.EVI, post ? 0 : 1);
    if(sd)
    {
        CcuViscParams &v = c->visc;
        const int first = v.sdepv_visits == 0;                     // a run that did not restart: unit strain rate on the first call
        if(!first && !c->en.have_v) FAIL("get_system_viscosity: stress-dependent viscosity needs the velocity of the last solve (ccu_v_from_vector)");
        if(!L.have_xx) FAIL("get_system_viscosity: coordinates missing");
        if(!v.sdepv_start_from_newtonian || v.sdepv_visits)
            LAUNCH(c, bk_visc_sdepv, cdiv(L.g.nel, 64), 64, L.g, c->visc, first, (const int *)c->mat, (const float *)L.XX, (const float *)c->en.V, (const float *)c->T,
                   Ccomp, L.EVI, c->rsphere ? 1 : 0);
        v.sdepv_visits++;
    }
    if(cd) LAUNCH(c, bk_visc_cdepv, cdiv(L.g.nel, 128), 128, L.g, c->visc, (const int *)c->mat, Ccomp, L.EVI);
    if(bd)
    {
        CcuViscParams &v = c->visc;
        const int first = v.bdepv_visits == 0;
        if(!first && !c->en.have_v) FAIL("get_system_viscosity: plastic yielding needs the velocity of the last solve (ccu_v_from_vector)");
        if(!L.have_xx) FAIL("get_system_viscosity: coordinates missing");
        LAUNCH(c, bk_visc_bdepv, cdiv(L.g.nel, 64), 64, L.g, c->visc, first, (const int *)c->mat, (const float *)L.XX, zco, (const float *)c->en.V, L.EVI,
               c->rsphere ? 1 : 0);
        v.bdepv_visits++;
    }
    if(post) LAUNCH(c, bk_visc_clip, cdiv(8 * (size_t)L.g.nel, 256), 256, 8 * (size_t)L.g.nel, c->visc, L.EVI);
    CK(cudaGetLastError());
    L.have_evi = true;
    return 0;
}

// construct_stiffness_B_matrix (Construct_arrays.c:834-889): project_viscosity, node_ks, BI, BPI on every level
int ccu_construct_stiffness_B_matrix(ccu_ctx *c, int augmented_Lagr, double augmented, int precondition)
{
    if(!c) FAIL("null context");
    if(ensure_tables(c) || ensure_nodal(c)) return 1;
    const int levmax = c->cfg.levmax, levmin = c->cfg.levmin;
    if(!c->L[levmax].have_evi) FAIL("construct_stiffness_B_matrix: finest-level viscosity missing");
    // project_viscosity (Solver_multigrid.c:398-474), the four modes of visc_smooth_cycles
    const int mode = c->visc.smooth_cycles;
    for(int lv = levmax; lv > levmin; lv--)
    {
        Level &Lf = c->L[lv], &Lc = c->L[lv - 1];
        if(!Lf.have_tw || !Lc.have_tw) FAIL("construct_stiffness_B_matrix: geometry not built");
        if(mode == 2 || mode == 3)
        {   // element means, injected per sub-element (2) or averaged over the sub-elements (3): no node sums, no exchange
            LAUNCH(c, bk_gint_to_ele, cdiv(Lf.g.nel, 128), 128, Lf.g.nel, (const float *)Lf.EVI, c->nodal_tmp);
            if(mode == 2) LAUNCH(c, bk_scalar_e_to_gint<0>, cdiv(Lc.g.nel, 128), 128, Lc.g, Lf.g, (const float *)c->nodal_tmp, Lc.EVI);
            else LAUNCH(c, bk_scalar_e_to_gint<1>, cdiv(Lc.g.nel, 128), 128, Lc.g, Lf.g, (const float *)c->nodal_tmp, Lc.EVI);
            Lc.have_evi = true;
            continue;
        }
        if(mode == 0)
        {   // nodal values of the fine level injected at the coincident nodes
            if(!c->multi()) LAUNCH(c, bk_gint_to_nodes, cdiv(Lf.g.nno, 128), 128, Lf.g, Lf.EVI, Lf.TWW, Lf.MASS, c->nodal_tmp);
            else
            {
                LAUNCH(c, bk_gint_to_nodes, cdiv(Lf.g.nno, 128), 128, Lf.g, Lf.EVI, Lf.TWW, (const float *)nullptr, c->nodal_tmp);
                if(ccu_halo_sum_nodal(c, lv, c->nodal_tmp)) return 1;
                LAUNCH(c, bk_mul, cdiv(Lf.g.nno, 128), 128, Lf.g.nno, c->nodal_tmp, Lf.MASS);
            }
            LAUNCH(c, bk_inject_scalar, cdiv(Lc.g.nno, 128), 128, Lc.g, Lf.g, (const float *)c->nodal_tmp, c->nodal_tmp2);
            LAUNCH(c, bk_nodes_to_gint, cdiv(Lc.g.nel, 128), 128, Lc.g, c->nodal_tmp2, Lc.EVI);
            Lc.have_evi = true;
            continue;
        }
        if(!c->multi())
        {
            LAUNCH(c, bk_gint_to_nodes, cdiv(Lf.g.nno, 128), 128, Lf.g, Lf.EVI, Lf.TWW, Lf.MASS, c->nodal_tmp);
            LAUNCH(c, bk_project_scalar, cdiv(Lc.g.nno, 128), 128, Lc.g, Lf.g, Lc.TWW, Lc.MASS, c->nodal_tmp, c->nodal_tmp2);
        }
        else
        {   // exchange_node_f20 between the element gather and the mass factor (Nodal_mesh.c:607-612, Solver_multigrid.c:380-386)
            LAUNCH(c, bk_gint_to_nodes, cdiv(Lf.g.nno, 128), 128, Lf.g, Lf.EVI, Lf.TWW, (const float *)nullptr, c->nodal_tmp);
            if(ccu_halo_sum_nodal(c, lv, c->nodal_tmp)) return 1;
            LAUNCH(c, bk_mul, cdiv(Lf.g.nno, 128), 128, Lf.g.nno, c->nodal_tmp, Lf.MASS);
            LAUNCH(c, bk_project_scalar, cdiv(Lc.g.nno, 128), 128, Lc.g, Lf.g, Lc.TWW, (const float *)nullptr, c->nodal_tmp, c->nodal_tmp2);
            if(ccu_halo_sum_nodal(c, lv - 1, c->nodal_tmp2)) return 1;
            LAUNCH(c, bk_mul, cdiv(Lc.g.nno, 128), 128, Lc.g.nno, c->nodal_tmp2, Lc.MASS);
        }
        LAUNCH(c, bk_nodes_to_gint, cdiv(Lc.g.nel, 128), 128, Lc.g, c->nodal_tmp2, Lc.EVI);
        Lc.have_evi = true;
    }
    // scratch for element blocks: chunks of node planes, budget ~3 GB
    const size_t budget_elems = ((size_t)3 << 30) / (324 * sizeof(double));
    for(int lev = levmax; lev >= levmin; lev--)
    {
        Level &L = c->L[lev];
        if(!L.have_flags || !L.have_xx) FAIL("construct_stiffness_B_matrix: flags/coordinates missing");
        const CcuGeom &g = L.g;
        const size_t per_plane = (size_t)g.elx * g.elz;
        int planes = (int)(budget_elems / per_plane);
        if(planes < 2) planes = 2;
        if(planes > g.ely) planes = g.ely;
        const size_t need = per_plane * planes;
        if(need > c->eltK_elems)
        {
            if(c->eltK) cudaFree(c->eltK);
            c->eltK = nullptr; c->eltK_elems = 0;
            CK(cudaMalloc(&c->eltK, sizeof(double) * 324 * need));
            c->eltK_elems = need;
        }
        // node planes [i0, i1) need element planes [max(i0-1,0), min(i1, ely))
        int i0 = 0;
        while(i0 < g.noy)
        {
            const int ey0 = i0 > 0 ? i0 - 1 : 0;
            int ey1 = ey0 + planes; if(ey1 > g.ely) ey1 = g.ely;
            int i1 = (ey1 == g.ely) ? g.noy : ey1;          // nodes of plane ey1 would need element plane ey1
            const int e_count = (int)(per_plane * (ey1 - ey0));
            if(c->rsphere)
            {
                if(!L.have_sxx) FAIL("construct_stiffness_B_matrix: spherical coordinates missing");
                LAUNCH(c, bk_elt_k_sph, cdiv(e_count, 64), 64, g, L.XX, L.SXX, L.EVI, (int)(per_plane * ey0), e_count, c->eltK);
            }
            else LAUNCH(c, bk_elt_k, cdiv(e_count, 64), 64, g, L.XX, L.EVI, (int)(per_plane * ey0), e_count, c->eltK);
            const size_t nn = (size_t)(i1 - i0) * g.nox * g.noz;
            LAUNCH(c, bk_node_ks, cdiv(nn, 64), 64, g, i0, i1, ey0, e_count, c->eltK, L.elt_del, L.EVI, L.flags, augmented_Lagr, augmented, L.K, L.BI, c->multi() ? 0 : 1);
            i0 = i1;
        }
        if(c->multi())
        {   // the diagonal of a duplicated node is the sum over its owners (exchange_id_d20, Construct_arrays.c:497)
            if(ccu_halo_sum_vec(c, lev, L.BI)) return 1;
            LAUNCH(c, bk_invert_BI, cdiv(L.vlen(), 128), 128, L.vlen(), L.BI);
        }
        LAUNCH(c, bk_BPI, cdiv(g.nel, 128), 128, g, L.elt_del, L.BI, precondition, L.BPI);
        if(c->multi() && ccu_damp_face_BI(c, lev)) return 1;              // rebuild_BI_on_boundary (Construct_arrays.c:892)
        L.have_K = true; L.have_p = true;
        if(ccu_col_refresh(c, lev)) return 1;                            // column-major copy for the column kernels (ccu_col.cuh)
    }
    CK(cudaGetLastError());
    SYNC(c);
    if(c->coarse)
    {   // replicated coarse levels: their operators are built from the gathered viscosity of level agg_lev, exactly as a
        // single subdomain owning the whole mesh would build them
        if(ensure_nodal(c->coarse)) return 1;
        c->coarse->visc = c->visc;
        if(ccu_agg_gather_evi(c)) return 1;
        return ccu_construct_stiffness_B_matrix(c->coarse, augmented_Lagr, augmented, precondition);
    }
    return 0;
}

// assemble_forces (Element_calculations.c:74): buoyancy[nno] -> resident F (CCU_VEC_F); optional host copy
int ccu_assemble_forces(ccu_ctx *c, const float *buoyancy, double *F_out)
{
    if(!c) FAIL("null context");
    if(ensure_tables(c) || ensure_nodal(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    if(!L.have_xx || !L.have_flags) FAIL("assemble_forces: coordinates/flags missing");
    if(buoyancy) CK(cudaMemcpyAsync(c->buoy, buoyancy, sizeof(float) * (size_t)L.g.nno, cudaMemcpyHostToDevice, c->st));
    if(!c->forceEF) CK(cudaMalloc(&c->forceEF, sizeof(double) * (c->rsphere ? 24 : 8) * (size_t)L.g.nel));
    if(c->rsphere)
    {
        if(!L.have_sxx) FAIL("assemble_forces: spherical coordinates missing");
        LAUNCH(c, bk_forces_elt_sph, cdiv(L.g.nel, 128), 128, L.g, L.XX, L.SXX, c->buoy, c->forceEF);
        LAUNCH(c, bk_forces_gather_sph, cdiv(L.g.nno, 128), 128, L.g, c->forceEF, L.flags, L.vec[CCU_VEC_F]);
    }
    else
    {
        LAUNCH(c, bk_forces_elt, cdiv(L.g.nel, 128), 128, L.g, L.XX, c->buoy, c->forceEF);
        LAUNCH(c, bk_forces_gather, cdiv(L.g.nno, 128), 128, L.g, c->forceEF, L.flags, L.vec[CCU_VEC_F]);
    }
    if(c->have_vb)
    {
        if(!L.have_evi || !L.node) FAIL("assemble_forces: the imposed-velocity term needs the viscosity as it stands (ccu_get_system_viscosity) and the node flags");
        if(c->vb_dirty)
        {   // which elements carry the term: counted, then listed (flags and VB only change through their setters)
            int *cnt = nullptr;
            CK(cudaMalloc(&cnt, sizeof(int)));
            CK(cudaMemsetAsync(cnt, 0, sizeof(int), c->st));
            LAUNCH(c, bk_vb_mark, cdiv(L.g.nel, 128), 128, L.g, L.node, c->VB[0], c->VB[1], c->VB[2], (int *)nullptr, (int *)nullptr, cnt);
            CK(cudaMemcpyAsync(&c->n_vb, cnt, sizeof(int), cudaMemcpyDeviceToHost, c->st));
            SYNC(c);
            cudaFree(c->vb_elems); cudaFree(c->vbEF); c->vb_elems = nullptr; c->vbEF = nullptr;
            if(!c->vb_slot) CK(cudaMalloc(&c->vb_slot, sizeof(int) * (size_t)L.g.nel));
            if(c->n_vb > 0)
            {
                CK(cudaMalloc(&c->vb_elems, sizeof(int) * (size_t)c->n_vb));
                CK(cudaMalloc(&c->vbEF, sizeof(double) * 24 * (size_t)c->n_vb));
                CK(cudaMemsetAsync(cnt, 0, sizeof(int), c->st));
                LAUNCH(c, bk_vb_mark, cdiv(L.g.nel, 128), 128, L.g, L.node, c->VB[0], c->VB[1], c->VB[2], c->vb_slot, c->vb_elems, cnt);
            }
            SYNC(c);
            cudaFree(cnt);
            c->vb_dirty = false;
        }
        if(c->n_vb > 0)
        {
            LAUNCH(c, bk_forces_vb_elt, cdiv(c->n_vb, 64), 64, L.g, L.XX, L.EVI, L.node, c->VB[0], c->VB[1], c->VB[2], c->n_vb, c->vb_elems, c->vbEF,
                   c->rsphere ? (const float *)L.SXX : (const float *)nullptr);
            LAUNCH(c, bk_forces_vb_gather, cdiv(L.g.nno, 128), 128, L.g, c->vb_slot, c->vbEF, c->n_vb, L.flags, L.vec[CCU_VEC_F]);
        }
    }
    if(ccu_halo_sum_vec(c, c->cfg.levmax, L.vec[CCU_VEC_F])) return 1;   // exchange_id_d20 (Element_calculations.c:119)
    if(F_out)
    {
        if(ccu_ensure_stage(c, sizeof(double) * L.g.neq)) return 1;
        LAUNCH(c, bk_vec_to_nat, cdiv(L.g.nno, 256), 256, L.g, L.vec[CCU_VEC_F], (double *)c->stage);
        CK(cudaMemcpyAsync(F_out, c->stage, sizeof(double) * L.g.neq, cudaMemcpyDeviceToHost, c->st));
    }
    SYNC(c);
    return 0;
}


// E->VB (global_defs.h:1038): imposed boundary velocities, [nno] per direction in the reference's node order; all NULL clears them.
// Non-zero values enter the force vector (the K.VB term of get_elt_f) and U (velocities_conform_bcs).
int ccu_set_velocity_bcs(ccu_ctx *c, const float *VB1, const float *VB2, const float *VB3)
{
    if(!c) FAIL("null context");
    Level &L = c->L[c->cfg.levmax];
    if(!VB1 && !VB2 && !VB3) { c->have_vb = false; return 0; }
    if(!VB1 || !VB2 || !VB3) FAIL("set_velocity_bcs: give all three directions or none");
    const float *src[3] = { VB1, VB2, VB3 };
    for(int d = 0; d < 3; d++)
    {
        if(!c->VB[d]) CK(cudaMalloc(&c->VB[d], sizeof(float) * (size_t)L.g.nno));
        CK(cudaMemcpyAsync(c->VB[d], src[d], sizeof(float) * (size_t)L.g.nno, cudaMemcpyHostToDevice, c->st));
    }
    SYNC(c);
    c->have_vb = true;      // also when every value is zero here: the other subdomains of a run may carry the moving boundary
    c->vb_dirty = true;
    return 0;
}
int ccu_conform_velocity_bcs(ccu_ctx *c)
{
    if(!c) FAIL("null context");
    Level &L = c->L[c->cfg.levmax];
    if(!L.have_flags || !L.node) FAIL("conform_velocity_bcs: node flags missing");
    if(!c->have_vb)
    {
        LAUNCH(c, bk_strip_u, cdiv(L.g.NS, 256), 256, L.g, L.flags, L.vec[CCU_VEC_U]);
        return 0;
    }
    LAUNCH(c, bk_conform_vbcs, cdiv(L.g.nno, 256), 256, L.g, L.node, c->VB[0], c->VB[1], c->VB[2], L.vec[CCU_VEC_U]);
    return 0;
}

// ---- read-back in the reference's layouts (tests, drop-in diagnostics)
int ccu_get_stiffness(ccu_ctx *c, int lev, float *k1, float *k2, float *k3, double *BI)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    const size_t n42 = (size_t)L.g.nno * 42;
    if(ccu_ensure_stage(c, sizeof(float) * 3 * n42 + sizeof(double) * L.g.neq)) return 1;
    float *s = (float *)c->stage;
    LAUNCH(c, bk_stiffness_to_ref, cdiv(L.g.nno, 128), 128, L.g, L.K, s, s + n42, s + 2 * n42);
    CK(cudaMemcpyAsync(k1, s, sizeof(float) * n42, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(k2, s + n42, sizeof(float) * n42, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(k3, s + 2 * n42, sizeof(float) * n42, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    if(BI)
    {
        double *sb = (double *)c->stage;
        LAUNCH(c, bk_vec_to_nat, cdiv(L.g.nno, 256), 256, L.g, L.BI, sb);
        CK(cudaMemcpyAsync(BI, sb, sizeof(double) * L.g.neq, cudaMemcpyDeviceToHost, c->st));
        SYNC(c);
    }
    return 0;
}

int ccu_get_level_array(ccu_ctx *c, int lev, int which, void *out)
{
    if(ccu_check_lev(c, lev)) return 2;
    Level &L = c->L[lev];
    const void *src = nullptr; size_t bytes = 0;
    switch(which)
    {
    case CCU_ARR_TWW: src = L.TWW; bytes = sizeof(float) * 8 * (size_t)L.g.nel; break;
    case CCU_ARR_MASS: src = L.MASS; bytes = sizeof(float) * (size_t)L.g.nno; break;
    case CCU_ARR_ECO_SIZE: src = L.eco; bytes = sizeof(float) * 3 * (size_t)L.g.nel; break;
    case CCU_ARR_ELT_DEL: src = L.elt_del; bytes = sizeof(float) * 24 * (size_t)L.g.nel; break;
    case CCU_ARR_BPI: src = L.BPI; bytes = sizeof(double) * (size_t)L.g.npno; break;
    case CCU_ARR_EVI: src = L.EVI; bytes = sizeof(float) * 8 * (size_t)L.g.nel; break;
    default: FAIL("get_level_array: unknown array id");
    }
    if(!src) FAIL("get_level_array: array not allocated");
    CK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}

// ================================================================= energy step, host side
static int ensure_energy(ccu_ctx *c)
{
    if(ensure_tables(c) || ensure_nodal(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    auto &E = c->en;
    const size_t nno = (size_t)L.g.nno;
    if(!E.Tdot)
    {
        CK(cudaMalloc(&E.Tdot, sizeof(float) * nno)); CK(cudaMemsetAsync(E.Tdot, 0, sizeof(float) * nno, c->st));
        CK(cudaMalloc(&E.DTdot, sizeof(float) * nno));
        CK(cudaMalloc(&E.V, sizeof(float) * 3 * nno)); CK(cudaMemsetAsync(E.V, 0, sizeof(float) * 3 * nno, c->st));
        CK(cudaMalloc(&E.T1, sizeof(float) * nno)); CK(cudaMalloc(&E.Tdot1, sizeof(float) * nno));
        CK(cudaMalloc(&E.diffusivity, sizeof(float) * L.g.noz)); CK(cudaMalloc(&E.expansivity, sizeof(float) * L.g.noz));
        CK(cudaMalloc(&E.Eres, sizeof(double) * 8 * (size_t)L.g.nel));
        CK(cudaMalloc(&E.layer, sizeof(double) * 2 * L.g.noz));
        CK(cudaMalloc(&E.red, sizeof(float) * 4));
    }
    return 0;
}
int ccu_set_energy_params(ccu_ctx *c, float fine_tune_dt, float fixed_timestep, float gamma, int temp_iterations,
                          const float *diffusivity, const float *expansivity, float Q0)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    auto &E = c->en;
    E.fine_tune_dt = fine_tune_dt; E.fixed_timestep = fixed_timestep; E.gamma = gamma; E.temp_iterations = temp_iterations; E.Q0 = Q0;
    CK(cudaMemcpyAsync(E.diffusivity, diffusivity, sizeof(float) * L.g.noz, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(E.expansivity, expansivity, sizeof(float) * L.g.noz, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    E.have_params = true; E.diff_timestep = -1.0f;
    return 0;
}
// E->control.{adi_heating, visc_heating, Atemp}, E->data.{disptn_number, surf_temp}: extended-Boussinesq heating terms
int ccu_set_heating_params(ccu_ctx *c, int adi_heating, int visc_heating, float disptn_number, float surf_temp, float Atemp)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    auto &E = c->en;
    E.adi_heating = adi_heating; E.visc_heating = visc_heating; E.disptn = disptn_number; E.surf_temp = surf_temp; E.Atemp_heat = Atemp;
    const size_t nel = (size_t)c->L[c->cfg.levmax].g.nel;
    if((adi_heating || visc_heating) && !E.heat_adi)
    {
        CK(cudaMalloc(&E.heat_adi, sizeof(float) * nel)); CK(cudaMemsetAsync(E.heat_adi, 0, sizeof(float) * nel, c->st));
        CK(cudaMalloc(&E.heat_visc, sizeof(float) * nel)); CK(cudaMemsetAsync(E.heat_visc, 0, sizeof(float) * nel, c->st));
    }
    return 0;
}
// the reference's own E->heating_adi / heating_visc / heating_latent (float[nel] = ptr + 1), e.g. when process_heating stays
// on the host (phase-change latent heating); NULL leaves the resident array (or the neutral value) in place
int ccu_set_heating_arrays(ccu_ctx *c, const float *heating_adi, const float *heating_visc, const float *heating_latent)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    auto &E = c->en;
    const size_t nel = (size_t)c->L[c->cfg.levmax].g.nel;
    const float *src[3] = { heating_adi, heating_visc, heating_latent };
    float **dst[3] = { &E.heat_adi, &E.heat_visc, &E.heat_latent };
    for(int q = 0; q < 3; q++)
    {
        if(!src[q]) continue;
        if(!*dst[q]) CK(cudaMalloc(dst[q], sizeof(float) * nel));
        CK(cudaMemcpyAsync(*dst[q], src[q], sizeof(float) * nel, cudaMemcpyHostToDevice, c->st));
    }
    SYNC(c);
    return 0;
}
// return_horiz_ave across subdomains (Global_operations.c:205-238): the layer sums of the ranks that share a z position
// (the reference's horizontal sub-communicator) are added.  One allreduce over ALL ranks of a table with one slot per z
// position does the same: every rank adds its sums into slot me_z, and reads that slot back.
// layer sums of a nodal field into E.layer (sums | weights), in the geometry of the context
static void launch_layer_sums(ccu_ctx *c, const float *field)
{
    Level &L = c->L[c->cfg.levmax];
    if(c->rsphere) LAUNCH(c, ek_layer_sums_sph, L.g.noz, 256, L.g, (const float *)L.XX, field, c->en.layer);
    else LAUNCH(c, ek_layer_sums, L.g.noz, 256, L.g, (const float *)L.XX, field, c->en.layer);
}
static int layer_allreduce(ccu_ctx *c)
{
    if(!c->multi()) return 0;
    auto &E = c->en;
    const CcuComm *m = c->comm;
    const int noz = c->L[c->cfg.levmax].g.noz, npz = m->nproc[2], mez = m->me[2];
    const size_t slot = 2 * (size_t)noz;
    if(!E.layer_tab) CK(cudaMalloc(&E.layer_tab, sizeof(double) * slot * npz));
    CK(cudaMemsetAsync(E.layer_tab, 0, sizeof(double) * slot * npz, c->st));
    CK(cudaMemcpyAsync(E.layer_tab + slot * mez, E.layer, sizeof(double) * slot, cudaMemcpyDeviceToDevice, c->st));
    if(ccu_allreduce_buffer(c, E.layer_tab, (int)(slot * npz), 0)) return 1;
    CK(cudaMemcpyAsync(E.layer, E.layer_tab + slot * mez, sizeof(double) * slot, cudaMemcpyDeviceToDevice, c->st));
    return 0;
}
// phase changes: E->viscosity.{zlm, z410} and E->control.{Ra_670, clapeyron670, width670, Ra_410, clapeyron410, width410} AS THE
// REFERENCE HOLDS THEM AFTER ITS FIRST phase_change CALL (Phase_change.c:51-67 rescales them once, in place)
int ccu_set_phase_params(ccu_ctx *c, float zlm, float z410, float Ra_670, float clapeyron670, float width670,
                         float Ra_410, float clapeyron410, float width410)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    auto &E = c->en;
    E.ph.zlm = zlm; E.ph.z410 = z410; E.ph.Ra670 = Ra_670; E.ph.clap670 = clapeyron670; E.ph.width670 = width670;
    E.ph.Ra410 = Ra_410; E.ph.clap410 = clapeyron410; E.ph.width410 = width410;
    E.phase_on = (Ra_670 != 0.0f || Ra_410 != 0.0f);
    const size_t nno = (size_t)c->L[c->cfg.levmax].g.nno, nel = (size_t)c->L[c->cfg.levmax].g.nel;
    if(E.phase_on && !E.Fas670)
    {
        CK(cudaMalloc(&E.Fas670, sizeof(float) * nno)); CK(cudaMalloc(&E.Fas410, sizeof(float) * nno));
        CK(cudaMalloc(&E.transT, sizeof(float) * 2)); CK(cudaMemsetAsync(E.transT, 0, sizeof(float) * 2, c->st));
        if(!E.heat_adi) { CK(cudaMalloc(&E.heat_adi, sizeof(float) * nel)); CK(cudaMemsetAsync(E.heat_adi, 0, sizeof(float) * nel, c->st)); }
        if(!E.heat_latent) CK(cudaMalloc(&E.heat_latent, sizeof(float) * nel));
        LAUNCH(c, bk_fill_f32, cdiv(nel, 256), 256, nel, 1.0f, E.heat_latent);
    }
    return 0;
}
// phase_change (Phase_change.c:43) on the resident T: the transition temperatures are re-read from the layer-average T
// when update_transT (the reference: first call and every 10th step, :82), then the nodal phase functions
int ccu_phase_change(ccu_ctx *c, int update_transT, float *Fas670_out, float *Fas410_out, float *transT_out /*[2]: 670, 410*/)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    auto &E = c->en;
    Level &L = c->L[c->cfg.levmax];
    if(!E.phase_on) FAIL("phase_change: ccu_set_phase_params first");
    // the depth coordinate: z of the box, r of the regional sphere (Xtmp = E->SX there, Phase_change.c:78-87); both [3][nno]
    if(c->rsphere && !L.have_sxx) FAIL("phase_change: spherical coordinates missing");
    const float *zco = c->rsphere ? (const float *)L.SXX : (const float *)L.XX;
    if(update_transT)
    {
        launch_layer_sums(c, (const float *)c->T);
        if(layer_allreduce(c)) return 1;
        LAUNCH(c, ek_phase_transT, 1, 32, L.g, zco, (const double *)E.layer, E.ph.zlm, E.ph.z410, E.transT);
        if(c->multi() && c->comm->nproc[2] > 1)
        {   // sum_across_depth (Global_operations.c:763, Phase_change.c:103,116): the phase depth lies in ONE z subdomain of a
            // vertical column of ranks (the others found 0): sum over the ranks with this rank's (x, y) through one allreduce table
            const CcuComm *m = c->comm;
            const int slots = 2 * m->nproc[0] * m->nproc[1], mine = 2 * (m->me[0] + m->nproc[0] * m->me[1]);
            if(!E.transT_tab) CK(cudaMalloc(&E.transT_tab, sizeof(double) * slots));
            CK(cudaMemsetAsync(E.transT_tab, 0, sizeof(double) * slots, c->st));
            LAUNCH(c, ek_transT_to_tab, 1, 32, (const float *)E.transT, E.transT_tab + mine);
            if(ccu_allreduce_buffer(c, E.transT_tab, slots, 0)) return 1;
            LAUNCH(c, ek_transT_from_tab, 1, 32, (const double *)(E.transT_tab + mine), E.transT);
        }
    }
    CcuPhase ph; ph.zlm = E.ph.zlm; ph.z410 = E.ph.z410; ph.Ra670 = E.ph.Ra670; ph.clap670 = E.ph.clap670; ph.width670 = E.ph.width670;
    ph.Ra410 = E.ph.Ra410; ph.clap410 = E.ph.clap410; ph.width410 = E.ph.width410; ph.transT670 = 0; ph.transT410 = 0;
    LAUNCH(c, ek_phase_functions, cdiv(L.g.nno, 256), 256, L.g, zco, (const float *)c->T, ph, (const float *)E.transT, E.Fas670, E.Fas410);
    const size_t nno = (size_t)L.g.nno;
    if(Fas670_out) CK(cudaMemcpyAsync(Fas670_out, E.Fas670, sizeof(float) * nno, cudaMemcpyDeviceToHost, c->st));
    if(Fas410_out) CK(cudaMemcpyAsync(Fas410_out, E.Fas410, sizeof(float) * nno, cudaMemcpyDeviceToHost, c->st));
    if(transT_out) CK(cudaMemcpyAsync(transT_out, E.transT, sizeof(float) * 2, cudaMemcpyDeviceToHost, c->st));
    CK(cudaGetLastError());
    SYNC(c);
    return 0;
}
// process_heating (Advection_diffusion.c:813) from the resident T, V and EVI[levmax]; outputs optional (float[nel])
int ccu_process_heating(ccu_ctx *c, float *heating_adi_out, float *heating_visc_out)
{   // (heating_latent: ccu_get_heating_latent)
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    auto &E = c->en;
    Level &L = c->L[c->cfg.levmax];
    if(!(E.adi_heating || E.visc_heating || E.phase_on)) return 0;
    if(!E.have_v) FAIL("process_heating: velocity missing (ccu_set_velocity / ccu_v_from_vector)");
    if(E.visc_heating && !L.have_evi) FAIL("process_heating: viscosity missing");
    if(!E.have_params) FAIL("process_heating: ccu_set_energy_params first (expansivity)");
    if(E.adi_heating || E.visc_heating)
        LAUNCH(c, ek_process_heating, cdiv(L.g.nel, 64), 64, L.g, L.XX, L.EVI, c->T, E.V, E.expansivity, E.adi_heating, E.visc_heating,
               E.disptn, E.surf_temp, E.Atemp_heat, E.heat_adi, E.heat_visc, (float *)nullptr, c->rsphere ? 1 : 0);
    if(E.phase_on)
    {   // latent heating from the phase functions of the last phase_change call (the reference reads E->Fas670 / Fas410 as they stand)
        CcuPhase ph; ph.zlm = E.ph.zlm; ph.z410 = E.ph.z410; ph.Ra670 = E.ph.Ra670; ph.clap670 = E.ph.clap670; ph.width670 = E.ph.width670;
        ph.Ra410 = E.ph.Ra410; ph.clap410 = E.ph.clap410; ph.width410 = E.ph.width410; ph.transT670 = 0; ph.transT410 = 0;
        if(!E.adi_heating) CK(cudaMemsetAsync(E.heat_adi, 0, sizeof(float) * (size_t)L.g.nel, c->st));
        LAUNCH(c, ek_latent_heating, cdiv(L.g.nel, 64), 64, L.g, (const float *)c->T, (const float *)E.V, (const float *)E.Fas670,
               (const float *)E.Fas410, ph, E.disptn, E.surf_temp, E.Atemp_heat, E.heat_adi, E.heat_latent);
    }
    const size_t nel = (size_t)L.g.nel;
    if(heating_adi_out) CK(cudaMemcpyAsync(heating_adi_out, E.heat_adi, sizeof(float) * nel, cudaMemcpyDeviceToHost, c->st));
    if(heating_visc_out && E.heat_visc) CK(cudaMemcpyAsync(heating_visc_out, E.heat_visc, sizeof(float) * nel, cudaMemcpyDeviceToHost, c->st));
    CK(cudaGetLastError());
    SYNC(c);
    return 0;
}
int ccu_set_tdot(ccu_ctx *c, const float *Tdot)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    const size_t nno = (size_t)c->L[c->cfg.levmax].g.nno;
    if(Tdot) CK(cudaMemcpyAsync(c->en.Tdot, Tdot, sizeof(float) * nno, cudaMemcpyHostToDevice, c->st));
    else CK(cudaMemsetAsync(c->en.Tdot, 0, sizeof(float) * nno, c->st));
    SYNC(c);
    return 0;
}
int ccu_set_velocity(ccu_ctx *c, const float *V1, const float *V2, const float *V3)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    const size_t nno = (size_t)c->L[c->cfg.levmax].g.nno;
    CK(cudaMemcpyAsync(c->en.V, V1, sizeof(float) * nno, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->en.V + nno, V2, sizeof(float) * nno, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->en.V + 2 * nno, V3, sizeof(float) * nno, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    c->en.have_v = true;
    return 0;
}
int ccu_v_from_vector(ccu_ctx *c, float *V_out)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    LAUNCH(c, ek_v_from_vector, cdiv(L.g.nno, 128), 128, L.g, (const double *)L.vec[CCU_VEC_U], c->en.V);
    c->en.have_v = true;
    if(V_out) CK(cudaMemcpyAsync(V_out, c->en.V, sizeof(float) * 3 * (size_t)L.g.nno, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}
// global_fmin / global_fmax (Global_operations.c:409,416) of the device scalar red[which]
static int reduce_scalar(ccu_ctx *c, int which, float *out)
{
    float v;
    CK(cudaMemcpyAsync(&v, c->en.red + which, sizeof(float), cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    if(c->multi())
    {
        double *buf = c->comm->dotstage;
        double d = (which == 0) ? -(double)v : (double)v;        // min through max
        CK(cudaMemcpyAsync(buf, &d, sizeof(double), cudaMemcpyHostToDevice, c->st));
        if(ccu_allreduce_buffer(c, buf, 1, 1)) return 1;
        CK(cudaMemcpyAsync(&d, buf, sizeof(double), cudaMemcpyDeviceToHost, c->st));
        SYNC(c);
        v = (float)((which == 0) ? -d : d);
    }
    *out = v;
    return 0;
}
static int std_timestep(ccu_ctx *c, float *dt)
{
    Level &L = c->L[c->cfg.levmax];
    auto &E = c->en;
    if(!E.have_params || !E.have_v) FAIL("std_timestep: energy parameters / velocity missing");
    if(E.fixed_timestep != 0.0f) { *dt = E.fixed_timestep; return 0; }
    if(E.diff_timestep < 0.0f)
    {
        LAUNCH(c, ek_set_red, 1, 1, E.red, 1.0e8f, -10.0f);
        LAUNCH(c, ek_timestep, cdiv(L.g.nel, 128), 128, L.g, L.eco, E.V, 0, E.red);
        float m;
        if(reduce_scalar(c, 0, &m)) return 1;
        E.diff_timestep = (float)(0.5 * (double)m);
    }
    LAUNCH(c, ek_set_red, 1, 1, E.red, 1.0e8f, -10.0f);
    LAUNCH(c, ek_timestep, cdiv(L.g.nel, 128), 128, L.g, L.eco, E.V, 1, E.red);
    float adv;
    CK(cudaMemcpyAsync(&adv, E.red, sizeof(float), cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    const float prod = E.fine_tune_dt * adv;
    adv = (float)(1.0e-32 + (double)(prod < E.diff_timestep ? prod : E.diff_timestep));
    // global_fmin of the per-rank candidates
    CK(cudaMemcpyAsync(E.red, &adv, sizeof(float), cudaMemcpyHostToDevice, c->st));
    return reduce_scalar(c, 0, dt);
}
int ccu_std_timestep(ccu_ctx *c, float *dt_out)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    return std_timestep(c, dt_out);
}
static int pg_solver(ccu_ctx *c)
{
    Level &L = c->L[c->cfg.levmax];
    auto &E = c->en;
    LAUNCH(c, ek_element_residual, cdiv(L.g.nel, 64), 64, L.g, L.XX, L.eco, L.node, c->T, E.Tdot, E.V, E.diffusivity, E.Q0,
           (const float *)E.heat_adi, (const float *)E.heat_visc, (const float *)E.heat_latent, E.Eres, c->rsphere ? 1 : 0);
    if(!c->multi()) { LAUNCH(c, ek_gather_residual, cdiv(L.g.nno, 128), 128, L.g, E.Eres, L.MASS, E.DTdot); return 0; }
    LAUNCH(c, ek_gather_residual, cdiv(L.g.nno, 128), 128, L.g, E.Eres, (const float *)nullptr, E.DTdot);
    if(ccu_halo_sum_nodal(c, c->cfg.levmax, E.DTdot)) return 1;      // exchange_node_f20 (Advection_diffusion.c:432)
    LAUNCH(c, bk_mul, cdiv(L.g.nno, 128), 128, L.g.nno, E.DTdot, L.MASS);
    return 0;
}
static int energy_ready(ccu_ctx *c)
{
    Level &L = c->L[c->cfg.levmax];
    if(!c->en.have_params || !c->en.have_v) FAIL("energy step: ccu_set_energy_params and a velocity (ccu_v_from_vector / ccu_set_velocity) are needed first");
    if(!L.have_xx || !L.have_tw || !L.node) FAIL("energy step: coordinates / geometry / node flags missing");
    return 0;
}
int ccu_pg_solver(ccu_ctx *c, float *DTdot_out)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c) || energy_ready(c)) return 1;
    if(pg_solver(c)) return 1;
    if(DTdot_out) CK(cudaMemcpyAsync(DTdot_out, c->en.DTdot, sizeof(float) * (size_t)c->L[c->cfg.levmax].g.nno, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}
static int tmax(ccu_ctx *c, float *out)
{
    Level &L = c->L[c->cfg.levmax];
    LAUNCH(c, ek_set_red, 1, 1, c->en.red, 1.0e8f, -10.0f);
    LAUNCH(c, ek_max, min(cdiv(L.g.nno, 256), 148u * 8u), 256, L.g.nno, (const float *)c->T, c->en.red);
    return reduce_scalar(c, 1, out);
}
int ccu_PG_timestep(ccu_ctx *c, float *T, float *Tdot, float *dt_out, float *T_interior_out)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c) || energy_ready(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    auto &E = c->en;
    const size_t nno = (size_t)L.g.nno;
    if(T) CK(cudaMemcpyAsync(c->T, T, sizeof(float) * nno, cudaMemcpyHostToDevice, c->st));
    if(Tdot) CK(cudaMemcpyAsync(E.Tdot, Tdot, sizeof(float) * nno, cudaMemcpyHostToDevice, c->st));
    float timestep;
    if(std_timestep(c, &timestep)) return 1;
    CK(cudaMemcpyAsync(E.T1, c->T, sizeof(float) * nno, cudaMemcpyDeviceToDevice, c->st));
    CK(cudaMemcpyAsync(E.Tdot1, E.Tdot, sizeof(float) * nno, cudaMemcpyDeviceToDevice, c->st));
    const float T_maxvaried = (float)1.01;
    float T_interior1, T_interior = 0.0f;
    if(tmax(c, &T_interior1)) return 1;
    float dt_reduced = 1.0f;
    int last_sub_iterations = 1, iredo;
    do
    {
        timestep = timestep * dt_reduced;
        iredo = 0;
        const float mult_p = (float)((1.0 - (double)E.gamma) * (double)timestep);
        LAUNCH(c, ek_predictor, cdiv(nno, 256), 256, (int)nno, L.node, mult_p, c->T, E.Tdot);
        for(int pass = 0; pass < E.temp_iterations; pass++)
        {
            if(pg_solver(c)) return 1;
            LAUNCH(c, ek_corrector, cdiv(nno, 256), 256, (int)nno, L.node, E.gamma * timestep, c->T, E.Tdot, (const float *)E.DTdot);
        }
        if(tmax(c, &T_interior)) return 1;
        if(T_interior / T_interior1 > T_maxvaried)
        {
            CK(cudaMemcpyAsync(c->T, E.T1, sizeof(float) * nno, cudaMemcpyDeviceToDevice, c->st));
            CK(cudaMemcpyAsync(E.Tdot, E.Tdot1, sizeof(float) * nno, cudaMemcpyDeviceToDevice, c->st));
            iredo = 1;
            dt_reduced = (float)((double)dt_reduced * 0.5);
            last_sub_iterations++;
        }
    } while(iredo == 1 && last_sub_iterations <= 5);
    if(dt_out) *dt_out = timestep;
    if(T_interior_out) *T_interior_out = T_interior;
    if(T) CK(cudaMemcpyAsync(T, c->T, sizeof(float) * nno, cudaMemcpyDeviceToHost, c->st));
    if(Tdot) CK(cudaMemcpyAsync(Tdot, E.Tdot, sizeof(float) * nno, cudaMemcpyDeviceToHost, c->st));
    CK(cudaGetLastError());
    SYNC(c);
    return 0;
}
int ccu_thermal_buoyancy(ccu_ctx *c, float Atemp, float *buoyancy_out)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    auto &E = c->en;
    if(!E.have_params || !L.have_xx) FAIL("thermal_buoyancy: energy parameters / coordinates missing");
    if(c->mk.ready) LAUNCH(c, ek_buoyancy_comp, cdiv(L.g.nno, 256), 256, L.g, Atemp, c->mk.Acomp, (const float *)c->T, (const float *)c->mk.C, (const float *)E.expansivity, c->buoy);
    else LAUNCH(c, ek_buoyancy, cdiv(L.g.nno, 256), 256, L.g, Atemp, (const float *)c->T, (const float *)E.expansivity, c->buoy);
    if(E.phase_on)
    {   // phase_change + buoyancy -= Ra_670 Fas670 + Ra_410 Fas410 (Pan_problem_misc_functions.c:129-134)
        if(ccu_phase_change(c, (E.step % 10 == 0) ? 1 : 0, nullptr, nullptr, nullptr)) return 1;
        LAUNCH(c, ek_phase_buoyancy, cdiv(L.g.nno, 256), 256, L.g.nno, E.ph.Ra670, E.ph.Ra410, (const float *)E.Fas670, (const float *)E.Fas410, c->buoy);
    }
    launch_layer_sums(c, (const float *)c->buoy);
    if(layer_allreduce(c)) return 1;                                   // the ranks of one horizontal plane share their sums
    LAUNCH(c, ek_remove_layer_ave, cdiv(L.g.nno, 256), 256, L.g, (const double *)E.layer, c->buoy);
    if(buoyancy_out) CK(cudaMemcpyAsync(buoyancy_out, c->buoy, sizeof(float) * (size_t)L.g.nno, cudaMemcpyDeviceToHost, c->st));
    CK(cudaGetLastError());
    SYNC(c);
    return 0;
}
int ccu_set_step(ccu_ctx *c, int solution_cycles)
{
    if(!c) FAIL("null context");
    c->en.step = solution_cycles;
    return 0;
}
int ccu_get_heating_latent(ccu_ctx *c, float *heating_latent_out)
{
    if(!c) FAIL("null context");
    if(!c->en.heat_latent) FAIL("get_heating_latent: no phase changes configured");
    CK(cudaMemcpyAsync(heating_latent_out, c->en.heat_latent, sizeof(float) * (size_t)c->L[c->cfg.levmax].g.nel, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}
// ================================================================= get_stress / get_STD_topo (Topo_gravity.c:352-562, 307-335)
// per element: S = sum_gp pre * strain rate / area - P (diagonal), pre = EVI * gDA, float arithmetic as the reference
__global__ void __launch_bounds__(64) st_element(const CcuGeom g, const float *__restrict__ XX, const float *__restrict__ EVI, const float *__restrict__ V,
                                                 const double *__restrict__ P, float *S6, const int sph)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= g.nel) return;
    const int ez = e % g.elz, ex = (e / g.elz) % g.elx, ey = e / (g.elz * g.elx);
    float X[3][8], gnx[3][8], VX[8], VY[8], VZ[8];
    load_elt_coords(g, XX, ey, ex, ez, X);
    for(int a = 1; a <= 8; a++)
    {
        const int n = elt_node(g, ey, ex, ez, a);
        VX[a - 1] = V[n]; VY[a - 1] = V[(size_t)g.nno + n]; VZ[a - 1] = V[2 * (size_t)g.nno + n];
    }
    float Sxx = 0.f, Syy = 0.f, Szz = 0.f, Sxy = 0.f, Sxz = 0.f, Szy = 0.f;
    double area = 0.0;
    for(int i = 0; i < 8; i++)
    {
        const float gda = (float)gp_geom(X, c_sh.Nxv + i, 64, 8, gnx);
        const float pre = EVI[(size_t)e * 8 + i] * gda;
        float Vzz = 0.f, Vxx = 0.f, Vyy = 0.f, Vxy = 0.f, Vxz = 0.f, Vzy = 0.f;
        if(sph)
        {   // Rsphere branch (Topo_gravity.c:429-439): x, y, z stand for theta, phi, r; gNX holds d/dtheta, d/dphi, d/dr
            double x[3], th, ph, ri;
            sph_point(X, c_sh.Nv + i, 8, x);
            sph_rotate_gnx(x, gnx);
            sph_rtf(x, th, ph, ri);
            const double ct = cos(th), sn = sin(th);
            for(int j = 0; j < 8; j++)
            {
                const double N = c_sh.Nv[8 * j + i];
                Vzz += VZ[j] * gnx[2][j];
                Vxx = (float)((double)Vxx + ((double)(VX[j] * gnx[0][j]) + VZ[j] * N) * ri);
                Vxz = (float)((double)Vxz + ((double)(VX[j] * gnx[2][j]) + ri * ((double)(VZ[j] * gnx[0][j]) - VX[j] * N)));
                Vyy = (float)((double)Vyy + (((double)(VY[j] * gnx[1][j]) + VX[j] * N * ct) / sn + VZ[j] * N) * ri);
                Vxy = (float)((double)Vxy + (((double)(VX[j] * gnx[1][j]) - VY[j] * N * ct) / sn + (double)(VY[j] * gnx[0][j])) * ri);
                Vzy = (float)((double)Vzy + ((double)(VY[j] * gnx[2][j]) + ri * ((double)(VZ[j] * gnx[1][j]) / sn - VY[j] * N)));
            }
        }
        else
        for(int j = 0; j < 8; j++)
        {
            Vzz += VZ[j] * gnx[2][j];
            Vxx += VX[j] * gnx[0][j];
            Vxz += (VX[j] * gnx[2][j] + VZ[j] * gnx[0][j]);
            Vyy += VY[j] * gnx[1][j];
            Vxy += (VX[j] * gnx[1][j] + VY[j] * gnx[0][j]);
            Vzy += (VY[j] * gnx[2][j] + VZ[j] * gnx[1][j]);
        }
        Sxx = (float)((double)Sxx + 2.0 * (double)pre * (double)Vxx);
        Syy = (float)((double)Syy + 2.0 * (double)pre * (double)Vyy);
        Szz = (float)((double)Szz + 2.0 * (double)pre * (double)Vzz);
        Sxy += pre * Vxy; Sxz += pre * Vxz; Szy += pre * Vzy;
        area += 1.0 * (double)gda;
    }
    const float ar = (float)area;                       // E->eco[e].area (Size_does_matter.c:712)
    Sxx /= ar; Syy /= ar; Szz /= ar; Sxz /= ar; Sxy /= ar; Szy /= ar;
    const double p = P[e];
    Szz = (float)((double)Szz - p); Sxx = (float)((double)Sxx - p); Syy = (float)((double)Syy - p);
    const size_t nel = (size_t)g.nel;                   // order of the output planes: SXX, SXY, SXZ, SYY, SZY, SZZ (get_stress's arguments)
    S6[e] = Sxx; S6[nel + e] = Sxy; S6[2 * nel + e] = Sxz; S6[3 * nel + e] = Syy; S6[4 * nel + e] = Szy; S6[5 * nel + e] = Szz;
}
// nodal value += TWW * element value in ascending element order, for the six components; * Mass unless a halo sum comes first
__global__ void __launch_bounds__(128) st_nodal(const CcuGeom g, const float *__restrict__ TWW, const float *__restrict__ MASS,
                                                const float *__restrict__ S6e, float *S6n)
{
    const int LUT[2][2][2] = { { {1, 4}, {2, 3} }, { {5, 8}, {6, 7} } };
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= g.nno) return;
    const int k = n % g.noz, j = (n / g.noz) % g.nox, i = n / (g.noz * g.nox);
    float acc[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
    for(int ey = i - 1; ey <= i; ey++)
    {
        if(ey < 0 || ey >= g.ely) continue;
        for(int ex = j - 1; ex <= j; ex++)
        {
            if(ex < 0 || ex >= g.elx) continue;
            for(int ez = k - 1; ez <= k; ez++)
            {
                if(ez < 0 || ez >= g.elz) continue;
                const int e = ez + g.elz * (ex + g.elx * ey);
                const float w = TWW[(size_t)e * 8 + LUT[k - ez][j - ex][i - ey] - 1];
                for(int q = 0; q < 6; q++) acc[q] += w * S6e[(size_t)q * g.nel + e];
            }
        }
    }
    const float mss = MASS ? MASS[n] : 1.0f;
    for(int q = 0; q < 6; q++) S6n[(size_t)q * g.nno + n] = acc[q] * mss;
}
__global__ void __launch_bounds__(128) st_topo(const CcuGeom g, const float *__restrict__ SZZ, float *tpg, float *tpgb)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;       // surface node snode - 1; surf_node[snode] = snode * noz (Construct_arrays.c:116)
    if(s >= g.nox * g.noy) return;
    const size_t top = (size_t)s * g.noz + (g.noz - 1), bot = (size_t)s * g.noz;
    tpg[s] = -2 * SZZ[top] + SZZ[top - 1];
    tpgb[s] = 2 * SZZ[bot] - SZZ[bot + 1];
}
// S_out: float[6][nno] in the order SXX, SXY, SXZ, SYY, SZY, SZZ (may be NULL); tpg / tpgb: float[nox*noy] (may be NULL).
// Needs the resident velocity, viscosity and pressure of the last solve (or ccu_pvec_upload).
int ccu_get_stress_topo(ccu_ctx *c, float *S_out, float *tpg_out, float *tpgb_out)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    auto &E = c->en;
    if(!E.have_v) FAIL("get_stress: velocity missing (ccu_set_velocity / ccu_v_from_vector)");
    if(!L.have_evi || !L.have_tw || !L.have_xx) FAIL("get_stress: viscosity / geometry missing");
    if(!c->P) FAIL("get_stress: pressure missing (no Stokes solve yet; ccu_pvec_upload)");
    const size_t nel = (size_t)L.g.nel, nno = (size_t)L.g.nno, nsf = (size_t)L.g.nox * L.g.noy;
    float *Se = nullptr, *Sn = nullptr, *tp = nullptr;
    CK(cudaMalloc(&Se, sizeof(float) * 6 * nel)); CK(cudaMalloc(&Sn, sizeof(float) * 6 * nno)); CK(cudaMalloc(&tp, sizeof(float) * 2 * nsf));
    LAUNCH(c, st_element, cdiv(nel, 64), 64, L.g, (const float *)L.XX, (const float *)L.EVI, (const float *)E.V, (const double *)c->P, Se, c->rsphere ? 1 : 0);
    int rc = 0;
    if(!c->multi()) LAUNCH(c, st_nodal, cdiv(nno, 128), 128, L.g, (const float *)L.TWW, (const float *)L.MASS, (const float *)Se, Sn);
    else
    {   // six exchange_node_f20 between the element sums and the mass factor (Topo_gravity.c:538-553)
        LAUNCH(c, st_nodal, cdiv(nno, 128), 128, L.g, (const float *)L.TWW, (const float *)nullptr, (const float *)Se, Sn);
        for(int q = 0; q < 6 && !rc; q++)
        {
            rc = ccu_halo_sum_nodal(c, c->cfg.levmax, Sn + q * nno);
            LAUNCH(c, bk_mul, cdiv(nno, 128), 128, L.g.nno, Sn + q * nno, L.MASS);
        }
    }
    LAUNCH(c, st_topo, cdiv(nsf, 128), 128, L.g, (const float *)(Sn + 5 * nno), tp, tp + nsf);
    if(!rc && S_out) rc = cudaMemcpyAsync(S_out, Sn, sizeof(float) * 6 * nno, cudaMemcpyDeviceToHost, c->st) != cudaSuccess;
    if(!rc && tpg_out) rc = cudaMemcpyAsync(tpg_out, tp, sizeof(float) * nsf, cudaMemcpyDeviceToHost, c->st) != cudaSuccess;
    if(!rc && tpgb_out) rc = cudaMemcpyAsync(tpgb_out, tp + nsf, sizeof(float) * nsf, cudaMemcpyDeviceToHost, c->st) != cudaSuccess;
    if(cudaStreamSynchronize(c->st) != cudaSuccess) rc = 1;
    cudaFree(Se); cudaFree(Sn); cudaFree(tp);
    if(rc) FAIL("get_stress: device error");
    CK_LAUNCHES(c);
    return 0;
}

// averages (Process_velocity.c:179-224): horizontal averages per z layer of the nodal viscosity, of the composition and
// of |V|^2 (-> layer vrms), from the resident velocity / viscosity / markers; return_horiz_ave (Global_operations.c:133) over
// the ranks of a horizontal plane as in thermal_buoyancy.  Outputs float[noz] each, any may be NULL.
__global__ void __launch_bounds__(256) ek_speed2(const int nno, const float *__restrict__ V, float *out)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if(n >= nno) return;
    const float a = V[n], b = V[(size_t)nno + n], cc = V[2 * (size_t)nno + n];
    out[n] = a * a + b * b + cc * cc;                  // float arithmetic, as the reference's temp[i]
}
__global__ void __launch_bounds__(128) ek_layer_finish(const int noz, const double *__restrict__ layer, const int root, float *out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k >= noz) return;
    float v = layer[noz + k] != 0.0 ? (float)(layer[k] / layer[noz + k]) : 0.0f;
    if(root) v = (float)sqrt((double)v);
    out[k] = v;
}
int ccu_averages(ccu_ctx *c, float *vrms_out, float *visc_out, float *C_out)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    auto &E = c->en;
    if(!L.have_xx) FAIL("averages: coordinates missing");
    const int noz = L.g.noz;
    float *lay = nullptr;
    CK(cudaMalloc(&lay, sizeof(float) * 3 * (size_t)noz));
    auto one = [&](const float *field, int root, float *dev_out, float *host_out) -> int
    {
        launch_layer_sums(c, field);
        if(layer_allreduce(c)) return 1;
        LAUNCH(c, ek_layer_finish, cdiv(noz, 128), 128, noz, (const double *)E.layer, root, dev_out);
        CK(cudaMemcpyAsync(host_out, dev_out, sizeof(float) * noz, cudaMemcpyDeviceToHost, c->st));
        return 0;
    };
    int rc = 0;
    if(vrms_out && !rc)
    {
        if(!E.have_v) { cudaFree(lay); FAIL("averages: velocity missing (ccu_set_velocity / ccu_v_from_vector)"); }
        LAUNCH(c, ek_speed2, cdiv(L.g.nno, 256), 256, L.g.nno, (const float *)E.V, c->nodal_tmp);
        rc = one(c->nodal_tmp, 1, lay, vrms_out);
    }
    if(visc_out && !rc)
    {
        if(!L.have_evi || !L.have_tw) { cudaFree(lay); FAIL("averages: viscosity missing"); }
        if(!c->multi()) LAUNCH(c, bk_gint_to_nodes, cdiv(L.g.nno, 128), 128, L.g, L.EVI, L.TWW, L.MASS, c->nodal_tmp);
        else
        {
            LAUNCH(c, bk_gint_to_nodes, cdiv(L.g.nno, 128), 128, L.g, L.EVI, L.TWW, (const float *)nullptr, c->nodal_tmp);
            rc = ccu_halo_sum_nodal(c, c->cfg.levmax, c->nodal_tmp);
            LAUNCH(c, bk_mul, cdiv(L.g.nno, 128), 128, L.g.nno, c->nodal_tmp, L.MASS);
        }
        if(!rc) rc = one(c->nodal_tmp, 0, lay + noz, visc_out);
    }
    if(C_out && !rc)
    {
        if(!c->mk.ready) { cudaFree(lay); FAIL("averages: no compositional field (markers) resident"); }
        rc = one((const float *)c->mk.C, 0, lay + 2 * noz, C_out);
    }
    if(!rc) { const cudaError_t e = cudaStreamSynchronize(c->st); if(e != cudaSuccess) rc = 1; }
    cudaFree(lay);
    if(rc) return rc;
    CK_LAUNCHES(c);
    return 0;
}
// heat_flux (Process_buoyancy.c:63-203): Nusselt numbers at the top and bottom of the box from the resident T and velocity
int ccu_heat_flux(ccu_ctx *c, float *Nut_out, float *Nub_out)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c) || energy_ready(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    auto &E = c->en;
    const size_t nel = (size_t)L.g.nel, nno = (size_t)L.g.nno;
    if(!E.hf) { CK(cudaMalloc(&E.hf, sizeof(float) * nno)); CK(cudaMalloc(&E.hf_area, sizeof(float) * nel)); CK(cudaMalloc(&E.hf_sums, sizeof(double) * 4)); }
    LAUNCH(c, hf_element, cdiv(nel, 64), 64, L.g, (const float *)L.XX, (const float *)c->T, (const float *)E.V, (const float *)E.diffusivity, E.Eres, E.hf_area, c->rsphere ? 1 : 0);
    int at_bottom = 1, at_top = 1;
    if(!c->multi()) LAUNCH(c, hf_nodal, cdiv(nno, 128), 128, L.g, (const float *)L.TWW, (const float *)L.MASS, (const double *)E.Eres, E.hf);
    else
    {
        LAUNCH(c, hf_nodal, cdiv(nno, 128), 128, L.g, (const float *)L.TWW, (const float *)nullptr, (const double *)E.Eres, E.hf);
        if(ccu_halo_sum_nodal(c, c->cfg.levmax, E.hf)) return 1;                   // exchange_node_f20 (:141)
        LAUNCH(c, bk_mul, cdiv(nno, 128), 128, L.g.nno, E.hf, L.MASS);
        at_bottom = c->comm->me[2] == 0; at_top = c->comm->me[2] == c->comm->nproc[2] - 1;
    }
    LAUNCH(c, hf_surface_sums, 1, 256, L.g, (const float *)E.hf, (const float *)E.hf_area, at_bottom, at_top, E.hf_sums);
    if(ccu_allreduce_buffer(c, E.hf_sums, 4, 0)) return 1;                          // return_horiz_sum over every plane = global sum
    double h[4];
    CK(cudaMemcpyAsync(h, E.hf_sums, sizeof(double) * 4, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    if(Nub_out) *Nub_out = (float)((float)h[0] / ((float)h[1] * 4));               // inp[] / outp[] are float (:62,178-186)
    if(Nut_out) *Nut_out = (float)((float)h[2] / ((float)h[3] * 4));
    return 0;
}
int ccu_get_temperature(ccu_ctx *c, float *T, float *Tdot)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    const size_t nno = (size_t)c->L[c->cfg.levmax].g.nno;
    if(T) CK(cudaMemcpyAsync(T, c->T, sizeof(float) * nno, cudaMemcpyDeviceToHost, c->st));
    if(Tdot) CK(cudaMemcpyAsync(Tdot, c->en.Tdot, sizeof(float) * nno, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}

// ================================================================= markers, host side
int ccu_markers_setup(ccu_ctx *c, int capacity, int markers_per_ele, int rnoz, const double *XP1, const double *XP2, const double *XP3,
                      const int *RG3, const double *XG1, const double *XG2, const unsigned *Element, float Acomp)
{
    if(!c) FAIL("null context");
    if(ensure_energy(c)) return 1;
    Level &L = c->L[c->cfg.levmax];
    auto &M = c->mk;
    if(M.X) FAIL("markers: already set up");
    if(capacity < 1) FAIL("markers: bad capacity");
    const size_t cap = (size_t)capacity, nel = (size_t)L.g.nel, nno = (size_t)L.g.nno;
    M.cap = capacity; M.markers_per_ele = markers_per_ele; M.rnoz = rnoz; M.Acomp = Acomp;
    CK(cudaMalloc(&M.X, sizeof(double) * 3 * cap)); CK(cudaMalloc(&M.Xpred, sizeof(double) * 3 * cap));
    CK(cudaMalloc(&M.VO, sizeof(float) * 3 * cap)); CK(cudaMalloc(&M.Vpred, sizeof(float) * 3 * cap));
    CK(cudaMemsetAsync(M.Xpred, 0, sizeof(double) * 3 * cap, c->st)); CK(cudaMemsetAsync(M.VO, 0, sizeof(float) * 3 * cap, c->st));
    CK(cudaMemsetAsync(M.Vpred, 0, sizeof(float) * 3 * cap, c->st));
    CK(cudaMalloc(&M.C12, sizeof(int) * cap)); CK(cudaMalloc(&M.CElement, sizeof(int) * cap));
    CK(cudaMalloc(&M.count, sizeof(int) * 2 * nel));
    CK(cudaMalloc(&M.CE, sizeof(float) * nel)); CK(cudaMemsetAsync(M.CE, 0, sizeof(float) * nel, c->st));
    CK(cudaMalloc(&M.C, sizeof(float) * nno)); CK(cudaMemsetAsync(M.C, 0, sizeof(float) * nno, c->st));
    const int nx = L.g.nox, ny = L.g.noy, nz = L.g.noz;
    CK(cudaMalloc(&M.XP, sizeof(double) * (nx + ny + nz)));
    CK(cudaMemcpyAsync(M.XP, XP1, sizeof(double) * nx, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(M.XP + nx, XP2, sizeof(double) * ny, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(M.XP + nx + ny, XP3, sizeof(double) * nz, cudaMemcpyHostToDevice, c->st));
    CK(cudaMalloc(&M.RG3, sizeof(int) * (rnoz + 1)));
    CK(cudaMemcpyAsync(M.RG3, RG3, sizeof(int) * (rnoz + 1), cudaMemcpyHostToDevice, c->st));
    CK(cudaMalloc(&M.Element, sizeof(unsigned) * nel));
    CK(cudaMemcpyAsync(M.Element, Element, sizeof(unsigned) * nel, cudaMemcpyHostToDevice, c->st));
    CK(cudaMalloc(&M.err, sizeof(int))); CK(cudaMemsetAsync(M.err, 0, sizeof(int), c->st));
    for(int d = 0; d < 3; d++) { M.XG1[d] = XG1[d]; M.XG2[d] = XG2[d]; }
    // dx, dy, dzz exactly as get_element's statics (Composition_adv.c:1113-1116)
    SYNC(c);
    M.ready = true;
    return 0;
}
static MkGrid mk_grid(ccu_ctx *c, const double *h_ends)
{
    Level &L = c->L[c->cfg.levmax];
    auto &M = c->mk;
    MkGrid m;
    m.XP1 = M.XP; m.XP2 = M.XP + L.g.nox; m.XP3 = M.XP + L.g.nox + L.g.noy; m.RG3 = M.RG3; m.rnoz = M.rnoz;
    m.dx = (h_ends[1] - h_ends[0]) / L.g.elx;
    m.dy = (h_ends[3] - h_ends[2]) / L.g.ely;
    m.dzz = (h_ends[5] - h_ends[4]) / (M.rnoz - 1);
    return m;
}
static int mk_ends(ccu_ctx *c, double e[6])
{
    Level &L = c->L[c->cfg.levmax];
    const int nx = L.g.nox, ny = L.g.noy, nz = L.g.noz;
    const int idx[6] = { 0, nx - 1, nx, nx + ny - 1, nx + ny, nx + ny + nz - 1 };
    for(int q = 0; q < 6; q++) CK(cudaMemcpyAsync(e + q, c->mk.XP + idx[q], sizeof(double), cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}
int ccu_markers_upload(ccu_ctx *c, int n, const double *X1, const double *X2, const double *X3, const int *C12, const int *CElement, const float *CE)
{
    if(!c) FAIL("null context");
    auto &M = c->mk;
    if(!M.ready) FAIL("markers: ccu_markers_setup first");
    if(n < 0 || n > M.cap) FAIL("markers: more markers than the capacity (markers_uplimit)");
    const size_t cap = (size_t)M.cap;
    M.n = n;
    CK(cudaMemcpyAsync(M.X, X1, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(M.X + cap, X2, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(M.X + 2 * cap, X3, sizeof(double) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(M.C12, C12, sizeof(int) * n, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(M.CElement, CElement, sizeof(int) * n, cudaMemcpyHostToDevice, c->st));
    if(CE) CK(cudaMemcpyAsync(M.CE, CE, sizeof(float) * (size_t)c->L[c->cfg.levmax].g.nel, cudaMemcpyHostToDevice, c->st));
    SYNC(c);
    return 0;
}
int ccu_markers_download(ccu_ctx *c, double *X, double *Xpred, float *VO, float *Vpred, int *CElement, float *C, float *CE)
{
    if(!c) FAIL("null context");
    auto &M = c->mk;
    if(!M.ready) FAIL("markers: ccu_markers_setup first");
    Level &L = c->L[c->cfg.levmax];
    const size_t cap = (size_t)M.cap, n = (size_t)M.n;
    for(int d = 0; d < 3; d++)
    {
        if(X) CK(cudaMemcpyAsync(X + d * n, M.X + d * cap, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
        if(Xpred) CK(cudaMemcpyAsync(Xpred + d * n, M.Xpred + d * cap, sizeof(double) * n, cudaMemcpyDeviceToHost, c->st));
        if(VO) CK(cudaMemcpyAsync(VO + d * n, M.VO + d * cap, sizeof(float) * n, cudaMemcpyDeviceToHost, c->st));
        if(Vpred) CK(cudaMemcpyAsync(Vpred + d * n, M.Vpred + d * cap, sizeof(float) * n, cudaMemcpyDeviceToHost, c->st));
    }
    if(CElement) CK(cudaMemcpyAsync(CElement, M.CElement, sizeof(int) * n, cudaMemcpyDeviceToHost, c->st));
    if(C) CK(cudaMemcpyAsync(C, M.C, sizeof(float) * (size_t)L.g.nno, cudaMemcpyDeviceToHost, c->st));
    if(CE) CK(cudaMemcpyAsync(CE, M.CE, sizeof(float) * (size_t)L.g.nel, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}
// ---- transfer_marker_properties (Composition_adv.c:682-704): clamp, [markers to their new subdomain], element_markers,
// get_C_from_markers
static int mk_scratch(ccu_ctx *c)
{
    auto &M = c->mk;
    if(M.code) return 0;
    const size_t cap = (size_t)M.cap;
    CK(cudaMalloc(&M.code, cap)); CK(cudaMalloc(&M.stayf, cap));
    CK(cudaMalloc(&M.perm, sizeof(int) * cap)); CK(cudaMalloc(&M.lv_idx, sizeof(int) * cap)); CK(cudaMalloc(&M.lv_code, sizeof(int) * cap));
    CK(cudaMalloc(&M.nsel, sizeof(int)));
    CK(cudaMalloc(&M.sX, sizeof(double) * 3 * cap)); CK(cudaMalloc(&M.sXpred, sizeof(double) * 3 * cap));
    CK(cudaMalloc(&M.sVO, sizeof(float) * 3 * cap)); CK(cudaMalloc(&M.sVpred, sizeof(float) * 3 * cap));
    CK(cudaMalloc(&M.sC12, sizeof(int) * cap)); CK(cudaMalloc(&M.sCElement, sizeof(int) * cap));
    CK(cudaMalloc(&M.sendbuf, sizeof(double) * CCU_MK_REC * cap)); CK(cudaMalloc(&M.recvbuf, sizeof(double) * CCU_MK_REC * cap));
    M.cub_bytes = 0;
    cub::DevicePartition::Flagged(nullptr, M.cub_bytes, thrust::counting_iterator<int>(0), M.stayf, M.perm, M.nsel, M.cap, c->st);
    CK(cudaMalloc(&M.cub_tmp, M.cub_bytes));
    return 0;
}
static bool mk_decomposed(const ccu_ctx *c) { return c->mk.np[0] * c->mk.np[1] * c->mk.np[2] > 1; }
// clamp + split: markers that now belong to a neighbouring subdomain are packed into M.sendbuf grouped by destination
// code (ascending, original order inside a group); the others are compacted in place, order kept.  sendcnt[27] on the host.
static int mk_clamp_split(ccu_ctx *c, double *&Xuse, const double ends[6], int sendcnt[27])
{
    auto &M = c->mk;
    const int n = M.n;
    const bool usePred = (Xuse == M.Xpred);
    LAUNCH(c, mk_clamp, cdiv(n, 256), 256, n, M.cap, (const unsigned *)M.Element, (const int *)M.CElement, M.XG1[0], M.XG1[1], M.XG1[2], M.XG2[0],
           M.XG2[1], M.XG2[2], Xuse);
    for(int q = 0; q < 27; q++) sendcnt[q] = 0;
    if(!mk_decomposed(c)) return 0;
    if(mk_scratch(c)) return 1;
    LAUNCH(c, mk_dest, cdiv(n, 256), 256, n, M.cap, (const unsigned *)M.Element, (const int *)M.CElement, (const double *)Xuse, ends[0], ends[2], ends[4],
           ends[1], ends[3], ends[5], M.me[0], M.me[1], M.me[2], M.np[0], M.np[1], M.np[2], M.code, M.stayf);
    size_t bytes = M.cub_bytes;
    CK(cub::DevicePartition::Flagged(M.cub_tmp, bytes, thrust::counting_iterator<int>(0), M.stayf, M.perm, M.nsel, n, c->st));
    c->launches++;
    int n_stay = 0;
    CK(cudaMemcpyAsync(&n_stay, M.nsel, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    const int nl = n - n_stay;
    if(nl == 0) return 0;
    // leavers: stable sort by destination on the host (they are a sliver of the markers)
    LAUNCH(c, mk_leaver_list, cdiv(nl, 256), 256, n, n_stay, (const int *)M.perm, (const unsigned char *)M.code, M.lv_idx, M.lv_code);
    std::vector<int> idx(nl), code(nl), order(nl);
    CK(cudaMemcpyAsync(idx.data(), M.lv_idx, sizeof(int) * nl, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(code.data(), M.lv_code, sizeof(int) * nl, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    for(int q = 0; q < nl; q++) { order[q] = q; sendcnt[code[q]]++; }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return code[a] < code[b]; });
    std::vector<int> sorted(nl);
    for(int q = 0; q < nl; q++) sorted[q] = idx[order[q]];
    CK(cudaMemcpyAsync(M.lv_idx, sorted.data(), sizeof(int) * nl, cudaMemcpyHostToDevice, c->st));
    LAUNCH(c, mk_pack, cdiv(nl, 256), 256, nl, M.cap, (const int *)M.lv_idx, (const double *)M.X, (const double *)M.Xpred, (const float *)M.VO,
           (const int *)M.C12, M.sendbuf);
    // stayers, order kept
    LAUNCH(c, (mk_gather<double, 3>), cdiv(n_stay, 256), 256, n_stay, M.cap, (const int *)M.perm, (const double *)M.X, M.sX);
    LAUNCH(c, (mk_gather<double, 3>), cdiv(n_stay, 256), 256, n_stay, M.cap, (const int *)M.perm, (const double *)M.Xpred, M.sXpred);
    LAUNCH(c, (mk_gather<float, 3>), cdiv(n_stay, 256), 256, n_stay, M.cap, (const int *)M.perm, (const float *)M.VO, M.sVO);
    LAUNCH(c, (mk_gather<float, 3>), cdiv(n_stay, 256), 256, n_stay, M.cap, (const int *)M.perm, (const float *)M.Vpred, M.sVpred);
    LAUNCH(c, (mk_gather<int, 1>), cdiv(n_stay, 256), 256, n_stay, M.cap, (const int *)M.perm, (const int *)M.C12, M.sC12);
    LAUNCH(c, (mk_gather<int, 1>), cdiv(n_stay, 256), 256, n_stay, M.cap, (const int *)M.perm, (const int *)M.CElement, M.sCElement);
    SYNC(c);                       // `sorted` must outlive the upload
    std::swap(M.X, M.sX); std::swap(M.Xpred, M.sXpred); std::swap(M.VO, M.sVO); std::swap(M.Vpred, M.sVpred);
    std::swap(M.C12, M.sC12); std::swap(M.CElement, M.sCElement);
    Xuse = usePred ? M.Xpred : M.X;
    M.n = n_stay;
    return 0;
}
// arrivals (records on the device) appended behind the resident markers
static int mk_append(ccu_ctx *c, int nrecv, const double *records_dev)
{
    auto &M = c->mk;
    if(nrecv == 0) return 0;
    if(M.n + nrecv > M.cap) FAIL("markers: number of markers over the limit (markers_uplimit), as the reference's Composition_adv.c:207");
    LAUNCH(c, mk_unpack, cdiv(nrecv, 256), 256, nrecv, M.cap, M.n, records_dev, M.X, M.Xpred, M.VO, M.C12, M.CElement);
    M.n += nrecv;
    return 0;
}
static int mk_finish(ccu_ctx *c, const MkGrid &m, double *Xuse)
{
    Level &L = c->L[c->cfg.levmax];
    auto &M = c->mk;
    const int n = M.n;
    CK(cudaMemsetAsync(M.count, 0, sizeof(int) * 2 * (size_t)L.g.nel, c->st));
    LAUNCH(c, mk_assign_count, cdiv(n, 128), 128, L.g, m, n, M.cap, (const double *)Xuse, (const int *)M.C12, M.CElement, M.count, M.err);
    LAUNCH(c, mk_element_C, cdiv(L.g.nel, 256), 256, L.g.nel, (const int *)M.count, M.CE);
    if(!c->multi()) LAUNCH(c, mk_nodal_C, cdiv(L.g.nno, 128), 128, L.g, (const float *)L.TWW, (const float *)L.MASS, (const float *)M.CE, M.C);
    else
    {   // exchange_node_f20 between the element sums and the mass factor (Composition_adv.c:797-800)
        LAUNCH(c, mk_nodal_C, cdiv(L.g.nno, 128), 128, L.g, (const float *)L.TWW, (const float *)nullptr, (const float *)M.CE, M.C);
        if(ccu_halo_sum_nodal(c, c->cfg.levmax, M.C)) return 1;
        LAUNCH(c, bk_mul, cdiv(L.g.nno, 128), 128, L.g.nno, M.C, L.MASS);
    }
    int err = 0;
    CK(cudaMemcpyAsync(&err, M.err, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    if(err) FAIL("markers: " + std::to_string(err) + " marker(s) left the z lookup table (the reference terminates here: '!!!overflow', Composition_adv.c:1176)");
    return 0;
}
// velocity at the markers and the position update of Euler (corrector = 0) / Runge_Kutta (1); returns the array that holds
// the positions the transfer works on (XMCpred / XMC)
static int mk_advect(ccu_ctx *c, float timestep, int corrector, MkGrid &m, double ends[6], double *&Xuse)
{
    if(!c) FAIL("null context");
    auto &M = c->mk;
    if(!M.ready) FAIL("markers: ccu_markers_setup / ccu_markers_upload first");      // a subdomain may hold no marker: it still takes part in every collective
    if(!c->en.have_v) FAIL("markers: a velocity (ccu_v_from_vector / ccu_set_velocity) is needed first");
    Level &L = c->L[c->cfg.levmax];
    if(mk_ends(c, ends)) return 1;
    m = mk_grid(c, ends);
    const int n = M.n;
    if(!corrector)
    {
        LAUNCH(c, mk_velocity, cdiv(n, 128), 128, L.g, m, n, M.cap, (const double *)M.X, (const float *)L.eco, (const float *)c->en.V, M.VO, M.CElement, M.err);
        LAUNCH(c, mk_advance, cdiv(n, 256), 256, n, M.cap, timestep, 0, (const float *)M.VO, (const float *)M.Vpred, M.X, M.Xpred, c->rsphere ? 1 : 0);
        Xuse = M.Xpred;
        return 0;
    }
    LAUNCH(c, mk_velocity, cdiv(n, 128), 128, L.g, m, n, M.cap, (const double *)M.Xpred, (const float *)L.eco, (const float *)c->en.V, M.Vpred, M.CElement, M.err);
    LAUNCH(c, mk_advance, cdiv(n, 256), 256, n, M.cap, timestep, 1, (const float *)M.VO, (const float *)M.Vpred, M.X, M.Xpred, c->rsphere ? 1 : 0);
    Xuse = M.X;
    return 0;
}
static int mk_step(ccu_ctx *c, float timestep, int corrector)
{
    MkGrid m; double ends[6]; double *Xuse = nullptr;
    if(mk_advect(c, timestep, corrector, m, ends, Xuse)) return 1;
    auto &M = c->mk;
    if(c->multi() && !mk_decomposed(c))
    {   // subdomain-per-GPU run: the decomposition is the communicator's
        for(int d = 0; d < 3; d++) { M.np[d] = c->comm->nproc[d]; M.me[d] = c->comm->me[d]; }
    }
    int sendcnt[27], recvcnt[27], nrecv = 0;
    if(mk_clamp_split(c, Xuse, ends, sendcnt)) return 1;
    if(mk_decomposed(c))
    {
        if(!c->multi()) FAIL("markers: a decomposition without a communicator: use ccu_markers_step_export / ccu_markers_import_finish");
        if(ccu_marker_exchange(c, sendcnt, M.sendbuf, CCU_MK_REC, recvcnt, M.recvbuf, (size_t)M.cap, &nrecv, M.n)) return 1;
        if(mk_append(c, nrecv, M.recvbuf)) return 1;
    }
    return mk_finish(c, m, Xuse);
}
// ---- the same step in two halves with the migrating markers handed over on the host (tests: two subdomains in one process)
int ccu_markers_set_decomp(ccu_ctx *c, const int nproc[3], const int me[3])
{
    if(!c) FAIL("null context");
    for(int d = 0; d < 3; d++) { c->mk.np[d] = nproc[d]; c->mk.me[d] = me[d]; }
    return 0;
}
int ccu_markers_step_export(ccu_ctx *c, float timestep, int corrector, int sendcnt[27], double *records_out, int max_records)
{
    MkGrid m; double ends[6]; double *Xuse = nullptr;
    if(mk_advect(c, timestep, corrector, m, ends, Xuse)) return 1;
    if(mk_clamp_split(c, Xuse, ends, sendcnt)) return 1;
    int nl = 0;
    for(int q = 0; q < 27; q++) nl += sendcnt[q];
    if(nl > max_records) FAIL("markers_step_export: record buffer too small");
    if(nl) CK(cudaMemcpyAsync(records_out, c->mk.sendbuf, sizeof(double) * CCU_MK_REC * nl, cudaMemcpyDeviceToHost, c->st));
    SYNC(c);
    return 0;
}
int ccu_markers_import_finish(ccu_ctx *c, int corrector, int nrecv, const double *records)
{
    if(!c) FAIL("null context");
    auto &M = c->mk;
    if(!M.ready) FAIL("markers: ccu_markers_setup first");
    if(nrecv)
    {
        if(mk_scratch(c)) return 1;
        if(nrecv > M.cap) FAIL("markers_import_finish: more records than the capacity");
        CK(cudaMemcpyAsync(M.recvbuf, records, sizeof(double) * CCU_MK_REC * nrecv, cudaMemcpyHostToDevice, c->st));
        if(mk_append(c, nrecv, M.recvbuf)) return 1;
    }
    double ends[6];
    if(mk_ends(c, ends)) return 1;
    const MkGrid m = mk_grid(c, ends);
    return mk_finish(c, m, corrector ? M.X : M.Xpred);
}
int ccu_markers_count(ccu_ctx *c) { return c ? c->mk.n : -1; }
int ccu_Euler(ccu_ctx *c, float timestep) { return mk_step(c, timestep, 0); }
int ccu_Runge_Kutta(ccu_ctx *c, float timestep) { return mk_step(c, timestep, 1); }
