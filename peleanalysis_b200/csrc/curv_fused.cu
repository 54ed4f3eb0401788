// curv_fused.cu -- the curvature tool's two stencil passes as ONE persistent, TMA-staged kernel for sm_100a.
//
// Reference data flow (curvature.cpp:310-567): Progress c = (S - pmin) * inv; G = grad c; n = G / -max(1e-14, |G|);
// FillBoundary(n); K = 0.5 * (d n_x/dx + d n_y/dy + d n_z/dz).  Round 1 ran this as two stencil kernels with the flame
// normal making a round trip through HBM (S -> c, n ; n -> K : 8 + 32 + 24 + 8 = 72 bytes per cell against 48 algorithmic).
// Here a work item is a block of K cells (rows x planes of one box, full x) swept along z:
//   * the raw scalar S is staged with a TWO-cell halo in y and z (one in x) by the same cp.async.bulk / mbarrier ring and
//     neighbour links as stencil_tma.cu; each consumer thread normalises its own four cells of a freshly landed plane IN
//     PLACE (one normalisation per staged value) and keeps them in registers for the z differences;
//   * n is computed for the block plus a ONE-cell rim in y and z -- always from valid cells of the box itself, whose
//     width-1 ghost cells come from the links / the materialised ghost cells exactly as in MODE_NORMAL_S, so every n value
//     is bit-identical to the one the owning tile writes;
//   * n_x neighbours travel by warp shuffle, n_y rows through a double-buffered shared-memory plane (one named barrier per
//     z step), n_z stays in a register queue; K leaves in the same sweep.  n is written (it is an output) but never re-read.
// Cells whose K stencil leaves the box (the outermost cell layer: they need n of a neighbour box, of the coarse level or of
// a wall) are NOT computed here: k_div_shell (kernels.cu) does those few per cent afterwards from the ghost-filled n, by the
// same rules as MODE_DIV.  Traffic: 8 (S) + 8 (c) + 24 (n) + 8 (K) per cell plus halo re-reads that mostly hit L2.
// Arithmetic is the reference's expression order with separate IEEE mul/add (-fmad=false): bit-exact.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "kernels.cuh"
#include "stencil_dev.cuh"

namespace pa {

namespace {

constexpr int CF_MAX_STAGES = 8;
constexpr int CF_XG_LANES = 30;                       // producer lanes 1 .. 30 fetch the x ghosts: (side, n-row) cells lane-1 and lane-1+30
constexpr int CF_MAX_ROWS = 32;                       // staged rows per item (K rows + 4); 2 * (rows - 2) x-ghost cells <= 60
constexpr int CF_STATIC_SMEM = 8 * 1024;              // upper bound of the static shared memory below

struct CfRec {
    PaTile t;                // y0, ny, z0, nz = the K rows / planes of the item: subsets of [1, n-2]
    int links;               // bit f set: face f has a neighbour link
    int pad;
    PaBoxDev bx;
    PaLayDev li, lo;
};

#ifndef PA_HOST_EMULATION
__device__ __forceinline__ void consumer_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
// generic-proxy writes to a stage (the in-place normalisation) are ordered before the async-proxy refill of that stage
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#else
__device__ __forceinline__ void consumer_bar(int nthreads) { cuemu::named_bar(1, nthreads); }
__device__ __forceinline__ void fence_proxy_async() {}
#endif

__device__ __forceinline__ void lds4(const double* p, double v[4]) {
    const double2 a = lds2(p), b = lds2(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
// first nv (1..4) values of a quad whose first element is 16-byte aligned
__device__ __forceinline__ void st4(double* p, const double v[4], int nv) {
    if (nv >= 2) stg2(p, v[0], v[1]); else p[0] = v[0];
    if (nv == 4) stg2(p + 2, v[2], v[3]); else if (nv == 3) p[2] = v[2];
}
// the reference's face differences of four consecutive cells: v[0..5] = x-1, the quad, x+4
__device__ __forceinline__ void cdiff4(double dxi, const double v[6], double d[4]) {
    double f[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) f[j] = dxi * (v[j + 1] - v[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = -(0.5 * ((-f[j]) + (-f[j + 1])));
}

template <int CW, bool PLAIN>
__global__ void __launch_bounds__((CW + 1) * 32, 1) k_curv_fused(const PaTile* __restrict__ tiles, int ntiles, GridArgs ga, StencilExtra ex,
                                                                  int stage_doubles /* multiple of 16 */, int S /* ring depth in planes */,
                                                                  unsigned long long* __restrict__ ticket, unsigned long long ticket_base,
                                                                  int abl /* bits 0-7: timing experiments only (PA_CF_ABLATE: results are wrong); bits 8-23: producer poll interval, ns */) {
    constexpr int CONSUMER_THREADS = CW * 32;
    PA_DYN_SMEM(smem_raw);
    double* sm = reinterpret_cast<double*>(smem_raw);                 // [S][stage_doubles] ring of S planes
    double* nybuf = sm + (long long)S * stage_doubles;                // [2][CONSUMER_THREADS * 4]: n_y of the two latest n planes
    double* stg_out = nybuf + 2 * CONSUMER_THREADS * 4;               // [4][CONSUMER_THREADS * 4]: n_x, n_z, K, Progress rows on their way out (TMA stores)
    __shared__ __align__(8) uint64_t full_bar[CF_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[CF_MAX_STAGES];
    __shared__ __align__(16) double xg_s[CF_MAX_STAGES][2][CF_MAX_ROWS];   // x ghosts of linked x faces: [stage][lo/hi][staged row]
    __shared__ __align__(16) CfRec rec_s[CF_MAX_STAGES];
    __shared__ __align__(8) uint64_t row_bar[CW][2];                  // per consumer warp: "step g done", g even / odd

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1u + CF_XG_LANES); mbar_init(&empty_bar[s], CW); }
        for (int w = 0; w < CW; ++w) { mbar_init(&row_bar[w][0], 1u); mbar_init(&row_bar[w][1], 1u); }
#ifndef PA_HOST_EMULATION
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    }
    __syncthreads();

    if (warp == CW) {
        // ===================================== producer warp =====================================
        // as in k_stencil_tma, with the staged region grown to K rows / planes +- 2 (always inside the box plus its ghost layer)
        if (lane > CF_XG_LANES) return;
        const unsigned psleep = (unsigned)(abl >> 8) & 0xffffu;          // nanoseconds between polls of the empty barrier (0: spin)
        int stage = 0;
        uint32_t ephase = 1;
        for (;;) {
            unsigned long long tk = 0;
            if (lane == 0) tk = atomicAdd(ticket, 1ULL) - ticket_base;
            tk = __shfl_sync(0x7fffffffu, tk, 0);
            if (tk >= (unsigned long long)ntiles) {
                if (psleep) mbar_wait_sleep(&empty_bar[stage], ephase, psleep); else mbar_wait(&empty_bar[stage], ephase);
                if (lane == 0) { rec_s[stage].t.lev = -1; mbar_arrive(&full_bar[stage]); }
                else cp_async_arrive_noinc(&full_bar[stage]);
                break;
            }
            const PaTile t = tiles[(int)tk];
            const LevArgs& L = ga.L[t.lev];
            const PaBoxDev bx = L.boxes[t.box];
            const PaLayDev li = L.lay_in[t.box];
            const PaNbr nb = L.nbr[t.box];
            const int c0 = L.in_comp;
            const int rows = t.ny + 4, nplanes = t.nz + 4, nrn = t.ny + 2;
            const int yf = t.y0 - 2, zf = t.z0 - 2;                // first staged row / plane (box-relative, >= -1)
            const uint32_t row_bytes = (uint32_t)li.P * 8u, plane_bytes = (uint32_t)rows * row_bytes;
            auto link_src = [&](int face, long long& cs) -> const double* {
                const PaNbrFace F = nb.f[face];
                cs = 0;
                if (F.nb < 0) return nullptr;
                const PaPeerSlab ps = L.peers[F.rank];
                const PaLayDev ln = L.lay_in[F.nb];
                cs = ps.cs;
                return ps.base + (long long)c0 * ps.cs + ln.off + (long long)(F.rel[2] + ln.ng) * ln.PS + (long long)(F.rel[1] + ln.ng) * ln.P + F.rel[0];
            };
            // row 0 / plane 0 of the box (first valid row of the first valid plane, x pad 0)
            const double* own = L.in + li.off + (long long)li.ng * li.PS + (long long)li.ng * li.P;
            long long cs_unused = 0;
            const double *s_ylo = nullptr, *s_zlo = nullptr, *s_yhi = nullptr, *s_zhi = nullptr;
            const double* xsrc[2] = {nullptr, nullptr};
            double* xdst[2] = {nullptr, nullptr};
            if (lane == 0) {
                s_ylo = link_src(1, cs_unused);
                s_zlo = link_src(2, cs_unused);
                s_yhi = link_src(4, cs_unused);
                s_zhi = link_src(5, cs_unused);
            } else {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int idx = lane - 1 + CF_XG_LANES * k;
                    const int side = idx / nrn, xr = idx - side * nrn;     // side 0 = x-lo, 1 = x-hi; n-row xr = staged row xr + 1
                    if (side < 2) {
                        const double* s0 = link_src(side ? 3 : 0, cs_unused);
                        if (s0) {
                            xsrc[k] = s0 + (long long)zf * li.PS + (long long)(t.y0 - 1 + xr) * li.P + ((side ? bx.n[0] : -1) + li.ng + li.xoff);
                            xdst[k] = &xg_s[0][side][xr + 1];
                        }
                    }
                }
            }
            const int nyb = bx.n[1], nzb = bx.n[2];
            for (int p = 0; p < nplanes; ++p) {
                if (psleep) mbar_wait_sleep(&empty_bar[stage], ephase, psleep); else mbar_wait(&empty_bar[stage], ephase);
                if (lane == 0) {
                    if (p == 0) {
                        CfRec& R = rec_s[stage];
                        R.t = t; R.bx = bx; R.li = li; R.lo = L.lay_out[t.box];
                        int lk = 0;
#pragma unroll
                        for (int f = 0; f < 6; ++f) lk |= (nb.f[f].nb >= 0) ? (1 << f) : 0;
                        R.links = lk;
                    }
                    mbar_expect_tx(&full_bar[stage], plane_bytes);
                    const int z = zf + p;
                    const double* zs = (z < 0) ? s_zlo : (z >= nzb ? s_zhi : nullptr);
                    double* dst = sm + (long long)stage * stage_doubles;
                    if (zs) {
                        tma_load_1d(dst, zs + (long long)z * li.PS + (long long)yf * li.P, plane_bytes, &full_bar[stage]);
                    } else {
                        int r0 = yf, r1 = yf + rows - 1;
                        if (r0 < 0 && s_ylo) {
                            tma_load_1d(dst, s_ylo + (long long)z * li.PS - li.P, row_bytes, &full_bar[stage]);
                            r0 = 0;
                        }
                        if (r1 >= nyb && s_yhi) {
                            tma_load_1d(dst + (long long)(rows - 1) * li.P, s_yhi + (long long)z * li.PS + (long long)nyb * li.P, row_bytes, &full_bar[stage]);
                            r1 = nyb - 1;
                        }
                        tma_load_1d(dst + (long long)(r0 - yf) * li.P, own + (long long)z * li.PS + (long long)r0 * li.P,
                                    (uint32_t)(r1 - r0 + 1) * row_bytes, &full_bar[stage]);
                    }
                } else {
                    if (xsrc[0]) cp_async_8(xdst[0] + stage * (2 * CF_MAX_ROWS), xsrc[0] + (long long)p * li.PS);
                    if (xsrc[1]) cp_async_8(xdst[1] + stage * (2 * CF_MAX_ROWS), xsrc[1] + (long long)p * li.PS);
                    cp_async_arrive_noinc(&full_bar[stage]);
                }
                if (++stage == S) { stage = 0; ephase ^= 1u; }
            }
        }
        return;
    }

    // ===================================== consumer warps =====================================
    // Warp w owns n-row(s) 1 + w * (32 / LPR) ... of the staged block (staged rows 0 and rows-1 are halo rows: read only).
    // The ring holds the RAW scalar; every value is normalised by its reader (the thread's own quad once per plane, kept in
    // registers; the rows above / below when they are read), so a step needs nothing from other warps until its K part:
    // n_y of the rows above / below of the plane before.  Those travel through a double-buffered shared-memory plane and are
    // synchronised warp to warp -- each warp arrives on its own mbarrier once per step and waits for its two neighbour warps
    // only, late in the step -- so the warps of a CTA drift apart instead of marching in lock-step through load / FP64 /
    // store phases.  Within a step the in-plane differences of plane s-1 (whose data landed a step ago) are computed BEFORE
    // the wait for plane s, so the TMA latency hides behind them.
    const int tid = threadIdx.x;
    int sc = 0;                    // stage of the plane being received
    uint32_t fphase = 0;
    unsigned gstep = 0;            // steps done by this warp since the kernel started (same sequence in every consumer warp)
    const double pmin = ex.pmin, pinv = ex.inv;
    constexpr int NYB = CONSUMER_THREADS * 4;                          // doubles per n_y buffer
    double* const nyme = nybuf + 4 * tid;                               // this thread's n_y quad in buffer 0
    const bool abl_nosync = (abl & 1) != 0, abl_nostore = (abl & 2) != 0, abl_nok = (abl & 4) != 0, abl_nochain = (abl & 8) != 0;

    for (;;) {
        mbar_wait(&full_bar[sc], fphase);
        const CfRec& R = rec_s[sc];
        const PaTile t = R.t;
        if (t.lev < 0) break;                          // end marker
        const int links = R.links;
        const int nx = R.bx.n[0], nyb = R.bx.n[1], nzb = R.bx.n[2];
        const int P = R.li.P;
        const long long lo_PS = R.lo.PS;
        const LevArgs& L = ga.L[t.lev];
        const double dxi = L.dxi[0], dyi = L.dxi[1], dzi = L.dxi[2];
        const long long cs_out = L.cs_out;
        const int rows = t.ny + 4, nplanes = t.nz + 4;
        const int xbase = R.li.ng + R.li.xoff;         // even
        // lanes per row: the power of two >= number of quads of a row; a warp holds 32 / LPR consecutive n-rows
        const int nq4 = (nx + 3) >> 2;
        int lsh = 0;
        while ((1 << lsh) < nq4) ++lsh;
        const int q = lane & ((1 << lsh) - 1);
        const int r = 1 + (warp << (5 - lsh)) + (lane >> lsh);             // staged row of this thread: box row y0 - 2 + r
        const int nv = nx - 4 * q <= 0 ? 0 : (nx - 4 * q >= 4 ? 4 : nx - 4 * q);   // valid cells of the quad
        const bool fullq = (nx & 3) == 0;                                  // every quad that has cells has four (warp-uniform)
        const bool isN = (r <= rows - 2) & (nv > 0);
        const bool isK = isN & (r >= 2) & (r <= rows - 3);
        const int yb = t.y0 - 2 + r;
        // c and n are written by the item that holds the cell as a K row / plane; the box's outermost rows / planes (no K
        // there) go with the first / last item
        const int wy0 = (t.y0 == 1) ? 0 : t.y0, wy1 = (t.y0 + t.ny == nyb - 1) ? nyb - 1 : t.y0 + t.ny - 1;
        const int wz0 = (t.z0 == 1) ? 0 : t.z0, wz1 = (t.z0 + t.nz == nzb - 1) ? nzb - 1 : t.z0 + t.nz - 1;
        const bool wrow = isN & (yb >= wy0) & (yb <= wy1);
        const int rs = isN ? r : 1, qs = isN ? q : 0;                      // idle threads read row 1: stay inside the stage
        const int soff = rs * P + xbase + 4 * qs;
        // staged values are the raw scalar S wherever they come from valid cells (own or a linked neighbour's) and already
        // progress values where they are this box's materialised ghost cells (unlinked faces, GhostXform fill): the row
        // below / above of this thread is such a row only at the box's first / last row.  (x - 0) * 1 leaves a value as it is.
        const bool ym_mat = (yb - 1 < 0) && !(links & 2), yp_mat = (yb + 1 >= nyb) && !(links & 16);
        const double pm_m = ym_mat ? 0.0 : pmin, pv_m = ym_mat ? 1.0 : pinv;
        const double pm_p = yp_mat ? 0.0 : pmin, pv_p = yp_mat ? 1.0 : pinv;
        const bool xlo_link = (links & 1) != 0, xhi_link = (links & 8) != 0;
        const bool first_q = isN & (q == 0), last_q = isN & (q == ((nx - 1) >> 2));
        const int jlast = (nx - 1) & 3;
        const int xg_row = rs < CF_MAX_ROWS ? rs : CF_MAX_ROWS - 1;
        // output element of the quad in staged plane 0; advanced by planes
        const long long oo = R.lo.off + (long long)(t.z0 - 2 + R.lo.ng) * lo_PS + (long long)(yb + R.lo.ng) * R.lo.P + (4 * q + R.lo.ng + R.lo.xoff);
        double* const out_n = L.out + oo;                                  // n_x; n_y, n_z follow at cs_out
        double* const out_c = ex.cout[t.lev] + oo;
        double* const out_k = ex.kout[t.lev] + oo;
        double* const aux_base = ex.aux[t.lev] ? ex.aux[t.lev] + oo : nullptr;
        const long long cg = ex.cs_aux[t.lev];
        const bool do_thr = ex.do_threshold != 0;
        const double thr_lo = ex.threshold, thr_hi = 1.0 - ex.threshold;
        const int lpr4 = 4 << lsh;
        const int nym = isK ? -lpr4 : 0, nyp = isK ? lpr4 : 0;             // n_y quads of the rows below / above in an n_y buffer
        // PA_CF_BULK=1: results leave as TMA bulk stores of whole rows (shared -> global, asynchronous) from per-thread staging
        // slots, issued by the first lane of a row, instead of vector stores.  Rows whose length is not a multiple of 4 cells
        // (16-byte granularity) and the optional un-normalised gradient output always take vector stores.
        const bool bulk = fullq && !aux_base && (abl & 16);
        const bool lead = isN & (q == 0);
        const uint32_t row_out_bytes = (uint32_t)nx * 8u;
        double* const snx = stg_out + 4 * tid;
        double* const snz = snx + NYB;
        double* const sk = snz + NYB;
        double* const scc = sk + NYB;

        double cB[4];                        // progress of plane s-1 (this thread's quad)
        double fzl[4];                       // z face difference dzi * (c(s-1) - c(s-2))
        double nzB[4], gzl[4];               // n_z of plane s-2 and its lower z face difference dzi * (n_z(s-2) - n_z(s-3))
        double dxh[4];                       // d n_x / dx of plane s-2
        unsigned clip = 0;                   // threshold clip of K: bit j = cell j of plane s-2 is outside [threshold, 1 - threshold]
#pragma unroll
        for (int j = 0; j < 4; ++j) cB[j] = fzl[j] = nzB[j] = gzl[j] = dxh[j] = 0.0;
        int sp = sc;                         // stage of plane s-1

        for (int s = 0; s < nplanes; ++s) {
            // ---- in-plane differences of plane s-1 (landed a step ago) ----
            double gx[4], gy[4];
            if (s >= 2) {
                const double* Sp = sm + (long long)sp * stage_doubles;
                double v[6], ym[4], yp[4];
                lds4(Sp + soff - P, ym);
                lds4(Sp + soff + P, yp);
                double xm = __shfl_up_sync(0xffffffffu, cB[3], 1);
                double xp = __shfl_down_sync(0xffffffffu, cB[0], 1);
                if (first_q) xm = xlo_link ? (xg_s[sp][0][xg_row] - pmin) * pinv : Sp[soff - 1];
                v[0] = xm; v[1] = cB[0]; v[2] = cB[1]; v[3] = cB[2]; v[4] = cB[3]; v[5] = xp;
                if (last_q) {
                    const double eh = xhi_link ? (xg_s[sp][1][xg_row] - pmin) * pinv : Sp[soff + jlast + 1];
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j == jlast) v[j + 2] = eh;
                }
                cdiff4(dxi, v, gx);
#pragma unroll
                for (int j = 0; j < 4; ++j) gy[j] = cdiff(dyi, (ym[j] - pm_m) * pv_m, cB[j], (yp[j] - pm_p) * pv_p);
            }
            // ---- plane s: this thread's quad, normalised once (curvature.cpp:316-320) ----
            if (s > 0) mbar_wait(&full_bar[sc], fphase);
            const int zb = t.z0 - 2 + s;
            const bool plane_mat = (zb < 0 && !(links & 4)) || (zb >= nzb && !(links & 32));
            double cN[4];
            lds4(sm + (long long)sc * stage_doubles + soff, cN);
            if (!plane_mat) {
#pragma unroll
                for (int j = 0; j < 4; ++j) cN[j] = (cN[j] - pmin) * pinv;
            }
            if (s >= 2) {
                // ---- flame normal of plane s-1 ----
                double gz[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double fzh = dzi * (cN[j] - cB[j]);
                    gz[j] = -(0.5 * ((-fzl[j]) + (-fzh)));
                    fzl[j] = fzh;
                }
                double n0[4], n1[4], n2[4];
                if (abl_nochain) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { n0[j] = gx[j]; n1[j] = gy[j]; n2[j] = gz[j]; }
                } else if (PLAIN) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double nn = -fmax(1e-14, sqrt(gx[j] * gx[j] + gy[j] * gy[j] + gz[j] * gz[j]));
                        n0[j] = gx[j] / nn; n1[j] = gy[j] / nn; n2[j] = gz[j] / nn;
                    }
                } else {
                    normal_quad(gx, gy, gz, n0, n1, n2);
                }
                const int zn = zb - 1;                                       // box plane of this n plane
                const bool wplane = wrow & (zn >= wz0) & (zn <= wz1) & !abl_nostore;
                if (!bulk) {
                    if (wplane) {
                        const long long o = (long long)(s - 1) * lo_PS;
                        if (fullq) {
                            stg2(out_c + o, cB[0], cB[1]); stg2(out_c + o + 2, cB[2], cB[3]);          // Progress (curvature.cpp:310-321)
                            stg2(out_n + o, n0[0], n0[1]); stg2(out_n + o + 2, n0[2], n0[3]);
                            stg2(out_n + o + cs_out, n1[0], n1[1]); stg2(out_n + o + cs_out + 2, n1[2], n1[3]);
                            stg2(out_n + o + 2 * cs_out, n2[0], n2[1]); stg2(out_n + o + 2 * cs_out + 2, n2[2], n2[3]);
                        } else {
                            st4(out_c + o, cB, nv);
                            st4(out_n + o, n0, nv);
                            st4(out_n + o + cs_out, n1, nv);
                            st4(out_n + o + 2 * cs_out, n2, nv);
                        }
                        if (aux_base) { st4(aux_base + o, gx, nv); st4(aux_base + o + cg, gy, nv); st4(aux_base + o + 2 * cg, gz, nv); }
                    }
                } else {
                    // the previous step's row stores have finished reading the staging slots (all groups but the latest, K's)
                    if (lead) bulk_wait_read<1>();
                    __syncwarp();
                    *reinterpret_cast<double2*>(scc) = make_double2(cB[0], cB[1]);
                    *reinterpret_cast<double2*>(scc + 2) = make_double2(cB[2], cB[3]);
                    *reinterpret_cast<double2*>(snx) = make_double2(n0[0], n0[1]);
                    *reinterpret_cast<double2*>(snx + 2) = make_double2(n0[2], n0[3]);
                    *reinterpret_cast<double2*>(snz) = make_double2(n2[0], n2[1]);
                    *reinterpret_cast<double2*>(snz + 2) = make_double2(n2[2], n2[3]);
                }
                // d n_x / dx of this plane from the x neighbours' n_x; the first / last cell of a row get a meaningless value
                // (they are K cells of k_div_shell)
                double w6[6], dxn[4];
                w6[0] = __shfl_up_sync(0xffffffffu, n0[3], 1);
                w6[5] = __shfl_down_sync(0xffffffffu, n0[0], 1);
                w6[1] = n0[0]; w6[2] = n0[1]; w6[3] = n0[2]; w6[4] = n0[3];
                cdiff4(dxi, w6, dxn);
                // ---- the neighbour warps have finished step gstep-1: their n_y of plane s-2 is written and they no longer read
                //      the n_y buffer this step overwrites ----
                if (gstep > 0 && !abl_nosync) {
                    const unsigned b = (gstep - 1) & 1u, ph = ((gstep - 1) >> 1) & 1u;
                    if (warp > 0) mbar_wait(&row_bar[warp - 1][b], ph);
                    if (warp + 1 < CW) mbar_wait(&row_bar[warp + 1][b], ph);
                }
                // this plane's n_y for the rows above / below (read by them one step later) -- and, with bulk stores, its way out
                double* const nyw = nyme + ((s - 1) & 1) * NYB;
                *reinterpret_cast<double2*>(nyw) = make_double2(n1[0], n1[1]);
                *reinterpret_cast<double2*>(nyw + 2) = make_double2(n1[2], n1[3]);
                if (bulk) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lead & wplane) {
                        const long long o = (long long)(s - 1) * lo_PS;
                        tma_store_1d(out_c + o, scc, row_out_bytes);                             // Progress (curvature.cpp:310-321)
                        tma_store_1d(out_n + o, snx, row_out_bytes);
                        tma_store_1d(out_n + o + cs_out, nyw, row_out_bytes);
                        tma_store_1d(out_n + o + 2 * cs_out, snz, row_out_bytes);
                    }
                    if (lead) bulk_commit();
                }
                if (s >= 4 && !abl_nok) {
                    // ---- K of plane s-2: n_y of rows r-1, r, r+1 written one step ago, n_z faces, d n_x/dx held ----
                    const double* nyr = nyme + (s & 1) * NYB;
                    double a4[4], b4[4], c4[4];
                    lds4(nyr, b4);
                    lds4(nyr + nym, a4);
                    lds4(nyr + nyp, c4);
                    double kk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double dy = cdiff(dyi, a4[j], b4[j], c4[j]);
                        const double gzh = dzi * (n2[j] - nzB[j]);
                        const double dz = -(0.5 * ((-gzl[j]) + (-gzh)));
                        kk[j] = 0.5 * (((0.0 + dxh[j]) + dy) + dz);                 // curvature.cpp:505-547
                    }
                    if (do_thr) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (clip & (1u << j)) kk[j] = 0.0;        // :549-567 (K only; n is clipped afterwards)
                    }
                    if (bulk) {
                        if (lead) bulk_wait_read<1>();                   // the previous step's K row has left its slot
                        __syncwarp();
                        *reinterpret_cast<double2*>(sk) = make_double2(kk[0], kk[1]);
                        *reinterpret_cast<double2*>(sk + 2) = make_double2(kk[2], kk[3]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lead & isK & !abl_nostore) tma_store_1d(out_k + (long long)(s - 2) * lo_PS, sk, row_out_bytes);
                        if (lead) bulk_commit();
                    } else if (isK & !abl_nostore) {
                        if (fullq) { double* pk = out_k + (long long)(s - 2) * lo_PS; stg2(pk, kk[0], kk[1]); stg2(pk + 2, kk[2], kk[3]); }
                        else st4(out_k + (long long)(s - 2) * lo_PS, kk, nv);
                    }
                } else if (bulk && lead) {
                    bulk_commit();                                       // keeps "all but the latest group" = "up to this step's rows"
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    gzl[j] = dzi * (n2[j] - nzB[j]);
                    nzB[j] = n2[j]; dxh[j] = dxn[j];
                }
                if (do_thr) {
                    clip = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (cB[j] < thr_lo || cB[j] > thr_hi) clip |= 1u << j;
                }
            } else {
                // no neighbour data is read in the first two steps of an item, but the wait is made all the same: a warp must
                // never get two steps ahead of its neighbours, or the phase bits of their two alternating barriers alias
                if (gstep > 0 && !abl_nosync) {
                    const unsigned b = (gstep - 1) & 1u, ph = ((gstep - 1) >> 1) & 1u;
                    if (warp > 0) mbar_wait(&row_bar[warp - 1][b], ph);
                    if (warp + 1 < CW) mbar_wait(&row_bar[warp + 1][b], ph);
                }
                if (s == 1) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) fzl[j] = dzi * (cN[j] - cB[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) cB[j] = cN[j];
            // this warp's n_y of the step is visible before it arrives; plane s-1 is no longer read by this warp
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&row_bar[warp][gstep & 1u]);
                if (s >= 1) mbar_arrive(&empty_bar[sp]);
            }
            ++gstep;
            sp = sc;
            if (++sc == S) { sc = 0; fphase ^= 1u; }
        }
        if (bulk) {                                                      // the staging slots are free for the next item
            if (lead) bulk_wait_read<0>();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[sp]);
    }
}

}  // namespace

namespace {

template <int CW, bool PLAIN>
cudaError_t launch_cf(const PaTile* tiles, int ntiles, int stage_doubles, const GridArgs& ga, const StencilExtra& ex, cudaStream_t st) {
    constexpr int THREADS = (CW + 1) * 32;
    const size_t stage_bytes = (size_t)stage_doubles * sizeof(double);
    const char* eb = getenv("PA_CF_BULK");
    const bool bulk = eb && eb[0] == '1';
    const size_t ny_bytes = (size_t)(2 + (bulk ? 4 : 0)) * CW * 32 * 4 * sizeof(double);   // n_y exchange (2) + staging of bulk-stored rows (4)
    const size_t budget = (size_t)227 * 1024 - CF_STATIC_SMEM - 1024 - ny_bytes;
    // two planes are being read (s-1 and s); the rest of the ring is data in flight
    int S = 2 + 2;
    const char* es = getenv("PA_CF_STAGES");
    if (es) S = std::max(3, atoi(es));
    if ((size_t)S * stage_bytes > budget) S = (int)(budget / stage_bytes);
    if (S > CF_MAX_STAGES) S = CF_MAX_STAGES;
    if (S < 3) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)S * stage_bytes + ny_bytes;
    static std::map<int, size_t> configured;
    static std::mutex mu;
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(k_curv_fused<CW, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev] = smem;
        }
    }
    cudaError_t e = cudaSuccess;
    const int nsm = stencil_num_sms(&e);
    if (e != cudaSuccess) return e;
    const int grid = std::min(ntiles, nsm);
    unsigned long long* tdev = nullptr;
    unsigned long long tbase = 0;
    e = stencil_ticket(st, (unsigned long long)ntiles, grid, &tdev, &tbase);
    if (e != cudaSuccess) return e;
    const char* ea = getenv("PA_CF_ABLATE");
    const char* ep = getenv("PA_CF_PSLEEP");
    const int flags = (((ea ? atoi(ea) : 0) & 0xef) | (bulk ? 16 : 0)) | ((ep ? std::min(std::max(atoi(ep), 0), 0xffff) : 200) << 8);
    PA_LAUNCH(grid, THREADS, smem, st, k_curv_fused<CW, PLAIN>)(tiles, ntiles, ga, ex, stage_doubles, S, tdev, tbase, flags);
    return cudaGetLastError();
}

}  // namespace

// Consumer warps per CTA (one CTA per SM).  Warps are allocated in fours, so the register budget per thread is set by
// CW + 1 rounded up to a multiple of four: 15 + 1 -> 128 registers, 19 + 1 -> 96.  PA_CF_CW picks (read once).
int curv_fused_consumer_warps() {
    static int cw = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (cw == 0) {
        const char* e = getenv("PA_CF_CW");
        cw = e ? atoi(e) : 15;                      // measured: 15 (128 registers) 5.7 ms, 19 (96 registers, spills) 7.2 ms
        if (cw != 15 && cw != 19) cw = 15;
    }
    return cw;
}
int curv_fused_max_rows() { return CF_MAX_ROWS; }
// largest staged plane (doubles) the kernel accepts: three stages plus the exchange / staging buffers must fit
int curv_fused_max_plane_doubles() { return (int)(((size_t)227 * 1024 - CF_STATIC_SMEM - 1024 - (size_t)6 * 19 * 32 * 4 * 8) / (3 * 8)) & ~15; }

cudaError_t launch_curv_fused(const PaTile* tiles, int ntiles, int max_plane_doubles, const GridArgs& ga, const StencilExtra& ex,
                              bool plain_math, cudaStream_t st) {
    if (ntiles <= 0) return cudaSuccess;
    const int stage_doubles = (max_plane_doubles + 15) & ~15;
    cudaError_t e;
    if (curv_fused_consumer_warps() == 15)
        e = plain_math ? launch_cf<15, true>(tiles, ntiles, stage_doubles, ga, ex, st) : launch_cf<15, false>(tiles, ntiles, stage_doubles, ga, ex, st);
    else
        e = plain_math ? launch_cf<19, true>(tiles, ntiles, stage_doubles, ga, ex, st) : launch_cf<19, false>(tiles, ntiles, stage_doubles, ga, ex, st);
    ++g_launches;
    return e;
}

}  // namespace pa
