/* pa_oracle.c -- CPU restatement of the reference's grad / curvature stencil path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's library; the product path never does.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this restatement bit-for-bit against
 * outputs of the compiled, unmodified reference (oracle/_ref, built by oracle/build_ref.py) on the
 * committed fixtures under tests/golden/ (the reference has no golden vectors of its own).
 *
 * The code follows the reference's own structure step by step (same objects, same loop order,
 * same floating-point expression order; compiled without FMA contraction) rather than the fused
 * formulation the CUDA product uses, so the two are independent derivations.  Citations use
 *   AX/ = Submodules/PelePhysics/Submodules/amrex/Src/      R/ = the PeleAnalysis root.
 *
 * Data convention of the C interface: a "field" is the concatenation, level-major then box order,
 * of every box's VALID region stored [k][j][i] (i fastest) -- the FArrayBox order without ghosts.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int lo[3], hi[3]; } box_t;

typedef struct {
    box_t dom;
    double dx[3], dxinv[3];
    int nb;
    box_t *bx;
    int64_t *off;        /* offset of each box's valid data inside a flat field */
    int ratio;           /* refinement ratio to the next coarser level (1 on level 0) */
} lev_t;

typedef struct pao_hier {
    int nlev;
    lev_t *lev;
    int is_per[3];
    int bc_kind[3];      /* per direction, used only when !is_per: 0 = Neumann, 1 = reflect_odd (R/Src/grad.cpp:182-191) */
    int64_t total;
} pao_hier;

/* ---------------------------------------------------------------- small box algebra */
static int b_ok(const box_t *b) { return b->lo[0] <= b->hi[0] && b->lo[1] <= b->hi[1] && b->lo[2] <= b->hi[2]; }
static int64_t b_npts(const box_t *b) { return (int64_t)(b->hi[0]-b->lo[0]+1)*(b->hi[1]-b->lo[1]+1)*(b->hi[2]-b->lo[2]+1); }
static box_t b_grow(box_t b, int n) { for (int d=0; d<3; ++d) { b.lo[d]-=n; b.hi[d]+=n; } return b; }
static box_t b_shift(box_t b, const int *s) { for (int d=0; d<3; ++d) { b.lo[d]+=s[d]; b.hi[d]+=s[d]; } return b; }
static box_t b_isect(box_t a, const box_t *b) {
    for (int d=0; d<3; ++d) { if (b->lo[d]>a.lo[d]) a.lo[d]=b->lo[d]; if (b->hi[d]<a.hi[d]) a.hi[d]=b->hi[d]; }
    return a;
}
static int b_contains(const box_t *b, int i, int j, int k) {
    return i>=b->lo[0] && i<=b->hi[0] && j>=b->lo[1] && j<=b->hi[1] && k>=b->lo[2] && k<=b->hi[2];
}
/* amrex::coarsen for possibly negative indices (floor division) */
static int crs(int i, int r) { return (i < 0) ? -((-i + r - 1) / r) : i / r; }
static box_t b_coarsen(box_t b, int r) { for (int d=0; d<3; ++d) { b.lo[d]=crs(b.lo[d],r); b.hi[d]=crs(b.hi[d],r); } return b; }

/* a FAB: data over a (possibly grown) box, i fastest */
typedef struct { box_t b; int n[3]; double *p; } fab_t;
static fab_t fab_make(box_t b, double init) {
    fab_t f; f.b = b;
    for (int d=0; d<3; ++d) f.n[d] = b.hi[d]-b.lo[d]+1;
    int64_t n = b_npts(&b);
    f.p = (double*)malloc(sizeof(double)*(size_t)n);
    for (int64_t q=0; q<n; ++q) f.p[q] = init;
    return f;
}
static inline double *fab_at(const fab_t *f, int i, int j, int k) {
    return f->p + ((int64_t)(k-f->b.lo[2])*f->n[1] + (j-f->b.lo[1]))*f->n[0] + (i-f->b.lo[0]);
}
typedef struct { box_t b; int n[3]; int *p; } ifab_t;
static ifab_t ifab_make(box_t b, int init) {
    ifab_t f; f.b = b;
    for (int d=0; d<3; ++d) f.n[d] = b.hi[d]-b.lo[d]+1;
    int64_t n = b_npts(&b);
    f.p = (int*)malloc(sizeof(int)*(size_t)n);
    for (int64_t q=0; q<n; ++q) f.p[q] = init;
    return f;
}
static inline int *ifab_at(const ifab_t *f, int i, int j, int k) {
    return f->p + ((int64_t)(k-f->b.lo[2])*f->n[1] + (j-f->b.lo[1]))*f->n[0] + (i-f->b.lo[0]);
}

/* ---------------------------------------------------------------- hierarchy */
pao_hier *pao_hier_create(int nlev, const int *domains, const double *dx, const int *ratios,
                          const int *nboxes, const int *boxes, const int *is_per, const int *bc_kind)
{
    pao_hier *h = (pao_hier*)calloc(1, sizeof(pao_hier));
    h->nlev = nlev;
    h->lev = (lev_t*)calloc((size_t)nlev, sizeof(lev_t));
    const int *bp = boxes;
    int64_t off = 0;
    for (int l=0; l<nlev; ++l) {
        lev_t *L = &h->lev[l];
        for (int d=0; d<3; ++d) {
            L->dom.lo[d] = domains[6*l+d]; L->dom.hi[d] = domains[6*l+3+d];
            L->dx[d] = dx[3*l+d];
            L->dxinv[d] = 1.0 / L->dx[d];            /* AX/Base/AMReX_Geometry.cpp:520-521 */
        }
        L->ratio = (l == 0) ? 1 : ratios[l-1];
        L->nb = nboxes[l];
        L->bx = (box_t*)malloc(sizeof(box_t)*(size_t)L->nb);
        L->off = (int64_t*)malloc(sizeof(int64_t)*(size_t)L->nb);
        for (int b=0; b<L->nb; ++b, bp+=6) {
            for (int d=0; d<3; ++d) { L->bx[b].lo[d]=bp[d]; L->bx[b].hi[d]=bp[3+d]; }
            L->off[b] = off; off += b_npts(&L->bx[b]);
        }
    }
    h->total = off;
    for (int d=0; d<3; ++d) { h->is_per[d]=is_per[d]; h->bc_kind[d]=bc_kind[d]; }
    return h;
}
void pao_hier_destroy(pao_hier *h) {
    if (!h) return;
    for (int l=0; l<h->nlev; ++l) { free(h->lev[l].bx); free(h->lev[l].off); }
    free(h->lev); free(h);
}
int64_t pao_total_cells(const pao_hier *h) { return h->total; }
int64_t pao_level_offset(const pao_hier *h, int lev) { return h->lev[lev].off[0]; }

/* Periodicity::shiftIntVect (AX/Base/AMReX_Periodicity.cpp:8-33): all combinations of
 * {-per..per step period} in periodic directions, per = smallest multiple of the period >= nghost. */
static int pshifts(const box_t *dom, const int *is_per, int ng, int (*out)[3], int maxn) {
    int per[3]={0,0,0}, jmp[3]={1,1,1};
    for (int d=0; d<3; ++d) if (is_per[d]) {
        int period = dom->hi[d]-dom->lo[d]+1;
        per[d] = jmp[d] = period;
        while (per[d] < ng) per[d] += period;
    }
    int n=0;
    for (int i=-per[0]; i<=per[0]; i+=jmp[0])
    for (int j=-per[1]; j<=per[1]; j+=jmp[1])
    for (int k=-per[2]; k<=per[2]; k+=jmp[2]) {
        if (n<maxn) { out[n][0]=i; out[n][1]=j; out[n][2]=k; }
        ++n;
    }
    return n;
}

/* FabArray::FillBoundary local-copy rule (AX/Base/AMReX_FabArrayBase.cpp:739-794, copy loop
 * AX/Base/AMReX_FBI.H:211-267): for every periodic shift p and every box k meeting grow(vbx,ng)+p,
 * ghost cells (isect-p) \ vbx of the receiving fab take src_k(isect).
 * If srcmap != NULL it records, per grown-fab cell, (src_box<<40 | linear index in the src VALID box), else -1. */
static void fill_boundary(const lev_t *L, const int *is_per, fab_t *fabs, int ng, int64_t **srcmap)
{
    int sh[729][3];
    int ns = pshifts(&L->dom, is_per, ng, sh, 729);
    for (int r=0; r<L->nb; ++r) {
        box_t vbx = L->bx[r];
        box_t rcv = b_grow(vbx, ng);
        for (int s=0; s<ns; ++s) {
            box_t rs = b_shift(rcv, sh[s]);
            for (int k=0; k<L->nb; ++k) {
                box_t is = b_isect(rs, &L->bx[k]);
                if (!b_ok(&is)) continue;
                const box_t *sb = &L->bx[k];
                int sn0 = sb->hi[0]-sb->lo[0]+1, sn1 = sb->hi[1]-sb->lo[1]+1;
                for (int kk=is.lo[2]; kk<=is.hi[2]; ++kk)
                for (int jj=is.lo[1]; jj<=is.hi[1]; ++jj)
                for (int ii=is.lo[0]; ii<=is.hi[0]; ++ii) {
                    int di=ii-sh[s][0], dj=jj-sh[s][1], dk=kk-sh[s][2];
                    if (b_contains(&vbx, di, dj, dk)) continue;           /* boxDiff(dst, vbx) */
                    if (fabs) *fab_at(&fabs[r], di, dj, dk) = *fab_at(&fabs[k], ii, jj, kk);
                    if (srcmap) {
                        int64_t lin = ((int64_t)(kk-sb->lo[2])*sn1 + (jj-sb->lo[1]))*sn0 + (ii-sb->lo[0]);
                        int n0 = rcv.hi[0]-rcv.lo[0]+1, n1 = rcv.hi[1]-rcv.lo[1]+1;
                        srcmap[r][((int64_t)(dk-rcv.lo[2])*n1 + (dj-rcv.lo[1]))*n0 + (di-rcv.lo[0])] = ((int64_t)k<<40) | lin;
                    }
                }
            }
        }
    }
}

/* BATbndryReg box for a cell-centred register (AX/Base/AMReX_BoxArray.H:174-230) */
static box_t bndry_box(box_t b, int face /*0..2 lo x,y,z; 3..5 hi*/, int in_rad, int out_rad, int extent)
{
    int d = face % 3, islo = face < 3;
    box_t r = b;
    if (islo) r.hi[d] = r.lo[d]; else r.lo[d] = r.hi[d];
    for (int t=0; t<3; ++t) { r.lo[t] -= extent; r.hi[t] += extent; }
    if (islo) { r.lo[d] = b.lo[d] - out_rad; r.hi[d] = b.lo[d] + in_rad - 1; }
    else      { r.lo[d] = b.hi[d] + 1 - in_rad; r.hi[d] = b.hi[d] + out_rad; }
    return r;
}

enum { M_COVERED=0, M_NOT_COVERED=1, M_OUTSIDE=2 };   /* AX/Boundary/AMReX_BndryData.H:44 */

/* MultiMask::define (AX/Boundary/AMReX_MultiMask.cpp:25-71) for one box / face */
static ifab_t make_mask(const lev_t *L, const int *is_per, int b, int face, int in_rad, int out_rad, int extent)
{
    box_t mb = bndry_box(L->bx[b], face, in_rad, out_rad, extent);
    ifab_t m = ifab_make(mb, M_NOT_COVERED);
    int ngrow = out_rad > extent ? out_rad : extent;
    box_t dom = L->dom;
    for (int d=0; d<3; ++d) if (is_per[d]) { dom.lo[d]-=ngrow; dom.hi[d]+=ngrow; }
    for (int k=mb.lo[2]; k<=mb.hi[2]; ++k)
    for (int j=mb.lo[1]; j<=mb.hi[1]; ++j)
    for (int i=mb.lo[0]; i<=mb.hi[0]; ++i)
        *ifab_at(&m,i,j,k) = b_contains(&dom,i,j,k) ? M_NOT_COVERED : M_OUTSIDE;
    /* setVal(covered, CPC(mask fabs <- valid boxes, periodic)) */
    int sh[27][3];
    int ns = pshifts(&L->dom, is_per, 0, sh, 27);
    for (int s=0; s<ns; ++s) {
        box_t ms = b_shift(mb, sh[s]);
        for (int q=0; q<L->nb; ++q) {
            box_t is = b_isect(ms, &L->bx[q]);
            if (!b_ok(&is)) continue;
            for (int k=is.lo[2]; k<=is.hi[2]; ++k)
            for (int j=is.lo[1]; j<=is.hi[1]; ++j)
            for (int i=is.lo[0]; i<=is.hi[0]; ++i)
                *ifab_at(&m, i-sh[s][0], j-sh[s][1], k-sh[s][2]) = M_COVERED;
        }
    }
    return m;
}

/* BndryRegister on the coarsened fine box (in 0, out 1, extent 2; NaN-initialised,
 * AX/Boundary/AMReX_BndryRegister.H:147-190) + copyFrom(coarse VALID cells, nghost 0, coarse periodicity)
 * (AX/Boundary/AMReX_BndryRegister.H:266-276, AX/LinearSolvers/MLMG/AMReX_MLCellLinOp.H:649-661). */
static fab_t make_crse_register(const lev_t *LC, const int *is_per, const fab_t *cfabs, box_t fine_bx, int ratio, int face)
{
    box_t cb = b_coarsen(fine_bx, ratio);
    box_t rb = bndry_box(cb, face, 0, 1, 2);
    fab_t reg = fab_make(rb, NAN);
    int sh[27][3];
    int ns = pshifts(&LC->dom, is_per, 0, sh, 27);
    for (int s=0; s<ns; ++s) {
        box_t rs = b_shift(rb, sh[s]);
        for (int q=0; q<LC->nb; ++q) {
            box_t is = b_isect(rs, &LC->bx[q]);
            if (!b_ok(&is)) continue;
            for (int k=is.lo[2]; k<=is.hi[2]; ++k)
            for (int j=is.lo[1]; j<=is.hi[1]; ++j)
            for (int i=is.lo[0]; i<=is.hi[0]; ++i)
                *fab_at(&reg, i-sh[s][0], j-sh[s][1], k-sh[s][2]) = *fab_at(&cfabs[q], i, j, k);
        }
    }
    return reg;
}

/* interpbndrydata_{x,y,z}_o3 (AX/Boundary/AMReX_InterpBndryData_3D_K.H:23-119), written once with the
 * two tangential directions (t1 < t2) as parameters; the expression order is that of the reference. */
static double interp_o3(const fab_t *crse, const ifab_t *mask, int dir, int i, int j, int k, int r)
{
    int idx[3] = {i,j,k};
    int c[3] = {crs(i,r), crs(j,r), crs(k,r)};
    int t1 = (dir==0) ? 1 : 0;
    int t2 = (dir==2) ? 1 : 2;
    int e1[3]={0,0,0}, e2[3]={0,0,0}; e1[t1]=1; e2[t2]=1;
#define MASK(a1,a2) (*ifab_at(mask, i+(a1)*r*e1[0]+(a2)*r*e2[0], j+(a1)*r*e1[1]+(a2)*r*e2[1], k+(a1)*r*e1[2]+(a2)*r*e2[2]))
#define CR(a1,a2)   (*fab_at(crse, c[0]+(a1)*e1[0]+(a2)*e2[0], c[1]+(a1)*e1[1]+(a2)*e2[1], c[2]+(a1)*e1[2]+(a2)*e2[2]))
    int lo = (MASK(-1,0) == M_NOT_COVERED) ? -1 : 0;
    int hi = (MASK( 1,0) == M_NOT_COVERED) ?  1 : 0;
    double fac = (hi == lo+1) ? 1.0 : 0.5;
    double d1 = fac*(CR(hi,0)-CR(lo,0));
    double d11 = (hi==lo+2) ? 0.5*(CR(1,0) - 2.*CR(0,0) + CR(-1,0)) : 0.;

    lo = (MASK(0,-1) == M_NOT_COVERED) ? -1 : 0;
    hi = (MASK(0, 1) == M_NOT_COVERED) ?  1 : 0;
    fac = (hi == lo+1) ? 1.0 : 0.5;
    double d2 = fac*(CR(0,hi)-CR(0,lo));
    double d22 = (hi==lo+2) ? 0.5*(CR(0,1) - 2.*CR(0,0) + CR(0,-1)) : 0.;

    double d12 = (MASK(-1,-1) == M_NOT_COVERED && MASK(1,-1) == M_NOT_COVERED &&
                  MASK(-1, 1) == M_NOT_COVERED && MASK(1, 1) == M_NOT_COVERED)
        ? 0.25*(CR(1,1)-CR(-1,1)+CR(-1,-1)-CR(1,-1)) : 0.0;

    double x1 = -0.5 + (idx[t1]-c[t1]*r+0.5)/r;
    double x2 = -0.5 + (idx[t2]-c[t2]*r+0.5)/r;
    return CR(0,0) + x1*d1 + (x1*x1)*d11 + x2*d2 + (x2*x2)*d22 + x1*x2*d12;
#undef MASK
#undef CR
}

/* poly_interp_coeff (AX/Boundary/AMReX_LOUtil_K.H:24-37) */
static void poly_interp_coeff(double xInt, const double *x, int N, double *c)
{
    for (int j=0; j<N; ++j) {
        double num=1.0, den=1.0;
        for (int i=0; i<N; ++i) if (i!=j) { num *= xInt-x[i]; den *= x[j]-x[i]; }
        c[j] = num/den;
    }
}

/* One scalar on one level: FillBoundary, coarse register + o3 interpolation, applyBC.
 * fabs: fine fabs grown by ng with valid data set.  cfabs: coarse fabs (valid only used) or NULL.
 * (setLevelBC / updateSolBC / applyBC: AX/LinearSolvers/MLMG/AMReX_MLCellLinOp.H:513-590,649-661,680-889;
 *  setBoxBC: AX/LinearSolvers/MLMG/AMReX_MLMGBndry.H:107-155; kernels AX/LinearSolvers/MLMG/AMReX_MLLinOp_K.H:14-325) */
static void apply_bc(const pao_hier *h, int l, fab_t *fabs, int ng, const fab_t *cfabs, int ratio)
{
    const lev_t *L = &h->lev[l];
    const int maxorder = 4;
    fill_boundary(L, h->is_per, fabs, ng, NULL);
    for (int b=0; b<L->nb; ++b) {
        box_t vbx = L->bx[b];
        for (int face=0; face<6; ++face) {
            int d = face%3, islo = face<3, s = islo ? 1 : -1;
            int physical = ((islo ? vbx.lo[d]==L->dom.lo[d] : vbx.hi[d]==L->dom.hi[d]) && !h->is_per[d]);
            ifab_t mv = make_mask(L, h->is_per, b, face, 0, 1, 0);          /* m_maskvals */
            box_t fb = mv.b;                                                /* = adjCell(vbx, face) */
            if (physical) {
                for (int k=fb.lo[2]; k<=fb.hi[2]; ++k)
                for (int j=fb.lo[1]; j<=fb.hi[1]; ++j)
                for (int i=fb.lo[0]; i<=fb.hi[0]; ++i) {
                    if (*ifab_at(&mv,i,j,k) > 0) {
                        int in[3]={i,j,k}; in[d]+=s;
                        double v = *fab_at(&fabs[b], in[0],in[1],in[2]);
                        *fab_at(&fabs[b],i,j,k) = (h->bc_kind[d]==1) ? -v : v;
                    }
                }
            } else {
                /* coarse-fine (Dirichlet tag, bcl = 0.5*ratio*dx).  The face may equally be next to another
                 * same-level box or a periodic image: then every mask value is 'covered' and nothing is written. */
                int any = 0;
                for (int64_t q=0; q<b_npts(&fb); ++q) if (mv.p[q] > 0) { any = 1; break; }
                if (any) {
                    fab_t bv = fab_make(bndry_box(vbx, face, 0, 1, 1), NAN);    /* BndryData::bndry, extent 1 */
                    if (cfabs) {
                        ifab_t bm = make_mask(L, h->is_per, b, face, 0, 2, 5);  /* BndryData::masks (AMReX_BndryData.H:265) */
                        fab_t reg = make_crse_register(&h->lev[l-1], h->is_per, cfabs, vbx, ratio, face);
                        for (int k=bv.b.lo[2]; k<=bv.b.hi[2]; ++k)
                        for (int j=bv.b.lo[1]; j<=bv.b.hi[1]; ++j)
                        for (int i=bv.b.lo[0]; i<=bv.b.hi[0]; ++i)
                            *fab_at(&bv,i,j,k) = interp_o3(&reg, &bm, d, i, j, k, ratio);
                        free(reg.p); free(bm.p);
                    }
                    int blen = vbx.hi[d]-vbx.lo[d]+1;
                    int NX = (blen+1 < maxorder) ? blen+1 : maxorder;
                    double bcl = 0.5*(double)ratio*L->dx[d];
                    double x[4] = {-bcl*L->dxinv[d], 0.5, 1.5, 2.5};
                    double coef[4] = {0,0,0,0};
                    poly_interp_coeff(-0.5, x, NX, coef);
                    for (int k=fb.lo[2]; k<=fb.hi[2]; ++k)
                    for (int j=fb.lo[1]; j<=fb.hi[1]; ++j)
                    for (int i=fb.lo[0]; i<=fb.hi[0]; ++i) {
                        if (*ifab_at(&mv,i,j,k) > 0) {
                            double tmp = 0.0;
                            for (int m=1; m<NX; ++m) {
                                int in[3]={i,j,k}; in[d]+=m*s;
                                tmp += *fab_at(&fabs[b], in[0],in[1],in[2]) * coef[m];
                            }
                            double *ph = fab_at(&fabs[b],i,j,k);
                            *ph = tmp;
                            *ph += *fab_at(&bv,i,j,k)*coef[0];
                        }
                    }
                    free(bv.p);
                }
            }
            free(mv.p);
        }
    }
}

/* getFluxes + average_face_to_cellcenter + mult(-1): writes 3 cell-centred gradient components.
 * (mlpoisson_flux_{x,y,z}: AX/LinearSolvers/MLMG/AMReX_MLPoisson_3D_K.H:36-126; betainv = 1/(-1):
 *  AX/LinearSolvers/MLMG/AMReX_MLCellABecLap.H:272-291; amrex_avg_fc_to_cc: AX/Base/AMReX_MultiFabUtil_3D_C.H:39-49) */
static void cell_gradient(const lev_t *L, const fab_t *fab, box_t vbx, double *g[3])
{
    int n0 = vbx.hi[0]-vbx.lo[0]+1, n1 = vbx.hi[1]-vbx.lo[1]+1;
    for (int d=0; d<3; ++d) {
        box_t fbx = vbx; fbx.hi[d] += 1;                       /* surroundingNodes */
        fab_t fl = fab_make(fbx, 0.0);
        double dxinv = L->dxinv[d];
        for (int k=fbx.lo[2]; k<=fbx.hi[2]; ++k)
        for (int j=fbx.lo[1]; j<=fbx.hi[1]; ++j)
        for (int i=fbx.lo[0]; i<=fbx.hi[0]; ++i) {
            int m[3]={i,j,k}; m[d]-=1;
            double f = dxinv*(*fab_at(fab,i,j,k) - *fab_at(fab,m[0],m[1],m[2]));
            *fab_at(&fl,i,j,k) = f * (1.0/(-1.0));              /* flux.mult(betainv) */
        }
        for (int k=vbx.lo[2]; k<=vbx.hi[2]; ++k)
        for (int j=vbx.lo[1]; j<=vbx.hi[1]; ++j)
        for (int i=vbx.lo[0]; i<=vbx.hi[0]; ++i) {
            int p[3]={i,j,k}; p[d]+=1;
            double cc = 0.5 * (*fab_at(&fl,i,j,k) + *fab_at(&fl,p[0],p[1],p[2]));
            g[d][((int64_t)(k-vbx.lo[2])*n1 + (j-vbx.lo[1]))*n0 + (i-vbx.lo[0])] = cc * (-1.0);
        }
        free(fl.p);
    }
}

static fab_t *level_fabs(const lev_t *L, const double *field, int ng, double ghost_init)
{
    fab_t *f = (fab_t*)malloc(sizeof(fab_t)*(size_t)L->nb);
    for (int b=0; b<L->nb; ++b) {
        box_t v = L->bx[b];
        f[b] = fab_make(b_grow(v, ng), ghost_init);
        const double *src = field + L->off[b];
        int n0 = v.hi[0]-v.lo[0]+1, n1 = v.hi[1]-v.lo[1]+1;
        for (int k=v.lo[2]; k<=v.hi[2]; ++k)
        for (int j=v.lo[1]; j<=v.hi[1]; ++j)
            memcpy(fab_at(&f[b], v.lo[0], j, k), src + ((int64_t)(k-v.lo[2])*n1 + (j-v.lo[1]))*n0, sizeof(double)*(size_t)n0);
    }
    return f;
}
static void free_fabs(fab_t *f, int nb) { for (int b=0; b<nb; ++b) free(f[b].p); free(f); }

/* Gradient of one scalar on every level, ghost cells by the rules above.
 * crse_ratio_override > 0 forces that ratio for the coarse register (curvature hard-codes 2,
 * R/Src/curvature.cpp:445,518); 0 = ratio of the level domains (composite grad, AX/LinearSolvers/MLMG/AMReX_MLLinOp.H:856-885). */
static void grad_all_levels(const pao_hier *h, const double *s, const double *crse_src,
                            double *gx, double *gy, double *gz, int lev_only, int ratio_override)
{
    for (int l=0; l<h->nlev; ++l) {
        if (lev_only >= 0 && l != lev_only) continue;
        const lev_t *L = &h->lev[l];
        fab_t *fabs = level_fabs(L, s, 1, 0.0);
        fab_t *cf = NULL;
        int needs_crse = (l > 0);
        if (needs_crse) cf = level_fabs(&h->lev[l-1], crse_src, 0, 0.0);
        int ratio = (ratio_override > 0) ? ratio_override : L->ratio;
        apply_bc(h, l, fabs, 1, cf, ratio);
        for (int b=0; b<L->nb; ++b) {
            double *g[3] = { gx + L->off[b], gy + L->off[b], gz + L->off[b] };
            cell_gradient(L, &fabs[b], L->bx[b], g);
        }
        free_fabs(fabs, L->nb);
        if (cf) free_fabs(cf, h->lev[l-1].nb);
    }
}

/* R/Src/grad.cpp:151-236.  out = 4 fields (gx, gy, gz, ||grad||), each `total` long. */
int pao_grad(const pao_hier *h, const double *s, double *out)
{
    int64_t T = h->total;
    double *gx = out, *gy = out+T, *gz = out+2*T, *mag = out+3*T;
    grad_all_levels(h, s, s, gx, gy, gz, -1, 0);
    for (int64_t q=0; q<T; ++q)
        mag[q] = sqrt(gx[q]*gx[q] + gy[q]*gy[q] + gz[q]*gz[q]);
    return 0;
}

/* R/Src/curvature.cpp:283-326 (progress variable) and :418-572 (mean curvature), default options plus
 * threshold_prog.  out = 5 fields: Progress, MeanCurvature, FlameNormalX, FlameNormalY, FlameNormalZ.
 * Optional outputs (may be NULL): gauss (GaussianCurvature, :575-677). */
int pao_curvature_ex(const pao_hier *h, const double *S, double progMin, double progMax,
                     int do_threshold, double threshold, int crse_ratio, double *out, double *gauss,
                     const double *U, double *sr, double *rost, double *veln);

int pao_curvature(const pao_hier *h, const double *S, double progMin, double progMax,
                  int do_threshold, double threshold, int crse_ratio, double *out, double *gauss)
{
    return pao_curvature_ex(h, S, progMin, progMax, do_threshold, threshold, crse_ratio, out, gauss, NULL, NULL, NULL, NULL);
}

/* Same with the velocity-based optional branches: U = 3 fields (x,y,z velocity) of `total` cells each;
 * sr = StrainRate (curvature.cpp:679-759; the -nn:grad u term is overwritten at :745-747, leaving div u, not clipped),
 * rost = 9 fields dU_m/dx_n at index 3m+n (:755-757), veln = u.n with the clipped normal (:761-789). */
int pao_curvature_ex(const pao_hier *h, const double *S, double progMin, double progMax,
                     int do_threshold, double threshold, int crse_ratio, double *out, double *gauss,
                     const double *U, double *sr, double *rost, double *veln)
{
    int64_t T = h->total;
    double *c = out, *K = out+T, *n[3] = { out+2*T, out+3*T, out+4*T };
    double invdenom = 1.0 / (progMax - progMin);
    for (int64_t q=0; q<T; ++q) c[q] = (S[q] - progMin) * invdenom;
    double *G[3], *tmp[3], *nrm = (double*)malloc(sizeof(double)*(size_t)T);
    for (int d=0; d<3; ++d) { G[d] = (double*)malloc(sizeof(double)*(size_t)T); tmp[d] = (double*)malloc(sizeof(double)*(size_t)T); }
    for (int l=0; l<h->nlev; ++l) {
        const lev_t *L = &h->lev[l];
        int64_t o0 = L->off[0];
        int64_t o1 = (l+1 < h->nlev) ? h->lev[l+1].off[0] : T;
        grad_all_levels(h, c, c, G[0], G[1], G[2], l, crse_ratio);
        for (int64_t q=o0; q<o1; ++q) {
            double v = fmax(1e-14, sqrt(pow(G[0][q],2.0) + pow(G[1][q],2.0) + pow(G[2][q],2.0)));
            nrm[q] = -v;
            for (int d=0; d<3; ++d) n[d][q] = G[d][q] / nrm[q];
        }
        for (int64_t q=o0; q<o1; ++q) K[q] = 0.0;
        for (int d=0; d<3; ++d) {
            /* coarse data for level l = flame_normal[l-1][d] AFTER its threshold clip (:514-518) */
            grad_all_levels(h, n[d], n[d], tmp[0], tmp[1], tmp[2], l, crse_ratio);
            for (int64_t q=o0; q<o1; ++q) K[q] += tmp[d][q];
        }
        for (int64_t q=o0; q<o1; ++q) K[q] *= 0.5;
        if (gauss) {
            /* Hessian rows from the un-normalised gradient ("cell_normal"), coarse = same on l-1 (:582-613) */
            double *H = (double*)malloc(sizeof(double)*9*(size_t)(o1-o0));
            for (int d=0; d<3; ++d) {
                grad_all_levels(h, G[d], G[d], tmp[0], tmp[1], tmp[2], l, crse_ratio);
                for (int e=0; e<3; ++e) memcpy(H + (size_t)(3*d+e)*(size_t)(o1-o0), tmp[e]+o0, sizeof(double)*(size_t)(o1-o0));
            }
            int64_t N = o1-o0;
            for (int64_t q=0; q<N; ++q) {
#define Hx(e) H[(0+(e))*N+q]
#define Hy(e) H[(3+(e))*N+q]
#define Hz(e) H[(6+(e))*N+q]
                double Ax0 = Hy(1)*Hz(2) - Hz(1)*Hy(2);
                double Ay0 = Hy(2)*Hz(0) - Hz(2)*Hy(0);
                double Az0 = Hy(0)*Hz(1) - Hz(0)*Hy(1);
                double Ax1 = Hx(2)*Hz(1) - Hz(2)*Hx(1);
                double Ay1 = Hx(0)*Hz(2) - Hz(0)*Hx(2);
                double Az1 = Hx(1)*Hz(0) - Hz(1)*Hx(0);
                double Ax2 = Hx(1)*Hy(2) - Hy(1)*Hx(2);
                double Ay2 = Hx(2)*Hy(0) - Hy(2)*Hx(0);
                double Az2 = Hx(0)*Hy(1) - Hy(0)*Hx(1);
#undef Hx
#undef Hy
#undef Hz
                double Cx = G[0][o0+q], Cy = G[1][o0+q], Cz = G[2][o0+q];
                double v = ( Cx * ( Ax0*Cx + Ax1*Cy + Ax2*Cz ) +
                             Cy * ( Ay0*Cx + Ay1*Cy + Ay2*Cz ) +
                             Cz * ( Az0*Cx + Az1*Cy + Az2*Cz ) ) / pow(nrm[o0+q], 4.0);
                if (do_threshold && (c[o0+q] < threshold || c[o0+q] > 1.0-threshold)) v = 0.0;
                gauss[o0+q] = v;
            }
            free(H);
        }
        if (do_threshold)
            for (int64_t q=o0; q<o1; ++q)
                if (c[q] < threshold || c[q] > 1.0-threshold) { K[q] = 0.0; n[0][q] = n[1][q] = n[2][q] = 0.0; }
        if (U && (sr || rost)) {
            double *dU = (double*)malloc(sizeof(double)*9*(size_t)(o1-o0));
            int64_t N = o1-o0;
            for (int m=0; m<3; ++m) {
                grad_all_levels(h, U + (int64_t)m*T, U + (int64_t)m*T, tmp[0], tmp[1], tmp[2], l, crse_ratio);
                for (int e=0; e<3; ++e) memcpy(dU + (size_t)(3*m+e)*(size_t)N, tmp[e]+o0, sizeof(double)*(size_t)N);
            }
            if (sr) for (int64_t q=0; q<N; ++q) sr[o0+q] = dU[0*N+q] + dU[4*N+q] + dU[8*N+q];
            if (rost) for (int m=0; m<9; ++m) memcpy(rost + (int64_t)m*T + o0, dU + (size_t)m*(size_t)N, sizeof(double)*(size_t)N);
            free(dU);
        }
        if (U && veln)
            for (int64_t q=o0; q<o1; ++q) {
                double v = U[q]*n[0][q] + U[T+q]*n[1][q] + U[2*T+q]*n[2][q];
                if (do_threshold && (c[q] < threshold || c[q] > 1.0-threshold)) v = 0.0;
                veln[q] = v;
            }
    }
    for (int d=0; d<3; ++d) { free(G[d]); free(tmp[d]); }
    free(nrm);
    return 0;
}

/* ---------------------------------------------------------------- inspection helpers for integer parity */

/* Ghost-filled fabs of one level for one scalar (same-level copies + BC fill), grown by ng,
 * concatenated in box order.  Ghost cells nothing writes keep `ghost_init`. */
int pao_filled_fabs(const pao_hier *h, int l, const double *s, int ng, double ghost_init, int crse_ratio, double *out)
{
    const lev_t *L = &h->lev[l];
    fab_t *fabs = level_fabs(L, s, ng, ghost_init);
    fab_t *cf = (l>0) ? level_fabs(&h->lev[l-1], s, 0, 0.0) : NULL;
    apply_bc(h, l, fabs, ng, cf, crse_ratio > 0 ? crse_ratio : L->ratio);
    int64_t o = 0;
    for (int b=0; b<L->nb; ++b) { int64_t n = b_npts(&fabs[b].b); memcpy(out+o, fabs[b].p, sizeof(double)*(size_t)n); o += n; }
    free_fabs(fabs, L->nb);
    if (cf) free_fabs(cf, h->lev[l-1].nb);
    return 0;
}

/* FillBoundary source map of one level: per grown-fab cell (src_box<<40 | src linear index) or -1. */
int pao_fb_source_map(const pao_hier *h, int l, int ng, int64_t *out)
{
    const lev_t *L = &h->lev[l];
    int64_t **maps = (int64_t**)malloc(sizeof(int64_t*)*(size_t)L->nb);
    int64_t o = 0;
    for (int b=0; b<L->nb; ++b) {
        box_t g = b_grow(L->bx[b], ng);
        int64_t n = b_npts(&g);
        maps[b] = out + o;
        for (int64_t q=0; q<n; ++q) maps[b][q] = -1;
        o += n;
    }
    fill_boundary(L, h->is_per, NULL, ng, maps);
    free(maps);
    return 0;
}

/* Mask plane of (box, face): kind 0 = m_maskvals (out 1, extent 0), kind 1 = BndryData mask
 * (out 2, extent 5).  Returns the number of ints written; box of the plane in pbox[6]. */
int64_t pao_mask(const pao_hier *h, int l, int b, int face, int kind, int *pbox, int *out)
{
    ifab_t m = kind ? make_mask(&h->lev[l], h->is_per, b, face, 0, 2, 5)
                    : make_mask(&h->lev[l], h->is_per, b, face, 0, 1, 0);
    int64_t n = b_npts(&m.b);
    for (int d=0; d<3; ++d) { pbox[d]=m.b.lo[d]; pbox[3+d]=m.b.hi[d]; }
    if (out) memcpy(out, m.p, sizeof(int)*(size_t)n);
    free(m.p);
    return n;
}
