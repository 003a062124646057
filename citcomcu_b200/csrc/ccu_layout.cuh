// ccu_layout.cuh -- HBM layout of one multigrid level (see DESIGN.md "Data layout").
//
// Nodes are stored 8-colour blocked: colour c = 4*(i&1) + 2*(j&1) + (k&1) with (i,j,k) the
// (y,x,z) indices of the reference numbering n = k + noz*(j + nox*i) (Construct_arrays.c:158-165).
// Each colour is a dense box of cells (ic,jc,kc) = (i>>1, j>>1, k>>1) with a one-cell zero halo,
// kc fastest, so that (a) one colour pass of the smoother streams its coefficients with unit
// stride, and (b) every stencil neighbour of a node is at a compile-time-constant offset from it
// (no Node_map, no bounds tests: out-of-grid neighbours are zero-valued halo entries multiplied by
// zero coefficients, the same "dummy equation" idea as the reference's neq+1 slot,
// Construct_arrays.c:305-309).
//
// Vectors are SoA: v[d*NS + s], d = dof, s = storage index.  Stiffness: K[(b*9 + a*3 + bb)*NS + s]
// = K(row dof a of node s, column dof bb of its neighbour b), b = 0 self, b = 1..13 the thirteen
// lexicographically lower neighbours in the fixed order CCU_LO below: exactly the 42x3 half-matrix
// entries the reference keeps per node (Eqn_k1-3, 504 B/node), each stored once.
#pragma once
#include <cstdint>

#ifndef CCU_KD_ALIGN
#define CCU_KD_ALIGN 1
#endif

struct CcuGeom
{
    int nox, noy, noz;       // nodes in x, y, z
    int elx, ely, elz;
    int nno, nel, neq, npno;
    int Kd, Jd, Id;          // cells per colour box incl. halo: z, x, y
    int JK;                  // Jd*Kd
    int NC;                  // cells per colour box
    int NS;                  // storage slots per dof (8*NC rounded up to 64)
};

__host__ __device__ inline CcuGeom ccu_make_geom(int nox, int noy, int noz)
{
    CcuGeom g;
    g.nox = nox; g.noy = noy; g.noz = noz;
    g.elx = nox - 1; g.ely = noy - 1; g.elz = noz - 1;
    g.nno = nox * noy * noz; g.nel = g.elx * g.ely * g.elz; g.neq = 3 * g.nno; g.npno = g.nel;
    g.Kd = (noz + 1) / 2 + 2; g.Jd = (nox + 1) / 2 + 2; g.Id = (noy + 1) / 2 + 2;
    if(g.Kd > 16) g.Kd = ((g.Kd + CCU_KD_ALIGN - 1) / CCU_KD_ALIGN) * CCU_KD_ALIGN;   // z rows start on a 32-byte sector (fp32) boundary
    g.JK = g.Jd * g.Kd;
    g.NC = g.Id * g.JK;
    g.NS = ((8 * g.NC + 63) / 64) * 64;
    return g;
}

__host__ __device__ inline int ccu_sidx(const CcuGeom &g, int i, int j, int k)
{
    const int c = ((i & 1) << 2) | ((j & 1) << 1) | (k & 1);
    return c * g.NC + ((i >> 1) + 1) * g.JK + ((j >> 1) + 1) * g.Kd + ((k >> 1) + 1);
}

// decode (colour, cell) -> node indices; returns false for halo / padding cells
__host__ __device__ inline bool ccu_decode(const CcuGeom &g, int c, int cell, int &i, int &j, int &k)
{
    const int ic = cell / g.JK, rem = cell - ic * g.JK;
    const int jc = rem / g.Kd, kc = rem - jc * g.Kd;
    i = 2 * (ic - 1) + ((c >> 2) & 1);
    j = 2 * (jc - 1) + ((c >> 1) & 1);
    k = 2 * (kc - 1) + (c & 1);
    return ic >= 1 && jc >= 1 && kc >= 1 && i < g.noy && j < g.nox && k < g.noz;
}

// thirteen lower neighbours (dy, dx, dz) in the reference's enumeration order (Construct_arrays.c:328-341)
#define CCU_LO_INIT { {-1,-1,-1},{-1,-1,0},{-1,-1,1},{-1,0,-1},{-1,0,0},{-1,0,1},{-1,1,-1},{-1,1,0},{-1,1,1}, \
                      {0,-1,-1},{0,-1,0},{0,-1,1},{0,0,-1} }

// cell shift along one axis when stepping d in {-1,0,1} from a node of parity p
__host__ __device__ constexpr int ccu_shift(int p, int d) { return d == 0 ? 0 : (d < 0 ? (p ? 0 : -1) : (p ? 1 : 0)); }

// element-local node a (1..8) -> (dz, dx, dy)   (element_definitions.h:211-222 used by Construct_arrays.c:74-82)
#define CCU_OFFS_INIT { {0,0,0},{0,0,0},{0,1,0},{0,1,1},{0,0,1},{1,0,0},{1,1,0},{1,1,1},{1,0,1} }
