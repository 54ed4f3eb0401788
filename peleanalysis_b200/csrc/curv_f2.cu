// curv_f2.cu -- second fused curvature kernel (PA_CURV_FUSED=2, opt-in): S -> Progress, flame normal and K in one sweep,
// built the other way round from curv_fused.cu.
//
// Reference data flow (curvature.cpp:310-567): c = (S - pmin) * inv; G = grad c; n = G / -max(1e-14, |G|); FillBoundary(n);
// K = 0.5 * div n.  curv_fused.cu keeps the z history of c and n in registers (four cells per thread, 128 registers, spills) and
// hands rows between warps through mbarriers; measured on a B200 it issues 823 instructions per warp and plane of which a
// third are FP64 and is bound by issue and latency.  Here nothing is carried in registers:
//   * one CTA of 512 threads per work item (K rows x K planes of one box, full x), swept along z;
//   * a plane of the scalar (item rows +- 2, x +- 1) is loaded with plain 128-bit loads one plane AHEAD into registers and,
//     one step later, normalised ONCE and stored into a four-plane shared-memory ring -- what the source of a cell is
//     (own valid cell, linked neighbour, materialised ghost cell already in progress space) is decided there, per row,
//     so the arithmetic loops below read uniform progress values and carry no flags;
//   * the flame normal of the middle plane goes to shared memory (n_x, n_y: two planes each, n_z: a ring of three) and,
//     where this item owns the cell, to HBM; K of the plane before follows from shared memory.  Two block barriers per plane;
//   * every thread handles two independent x-pairs per phase (four sqrt -> reciprocal -> quotient chains in flight).
// Cells whose K stencil leaves the box are left to k_div_shell, exactly as with curv_fused.cu.  Arithmetic is the reference's
// expression order with separate IEEE multiplies and adds (-fmad=false): bit-identical to the separate kernels.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>

#include "kernels.cuh"
#include "stencil_dev.cuh"

namespace pa {

namespace {

constexpr int F2_THREADS = 512;
constexpr int F2_KR = 14;                 // K rows per item
constexpr int F2_NR = F2_KR + 2;          // rows of n
constexpr int F2_SR = F2_KR + 4;          // rows of the scalar
constexpr int F2_NXMAX = 128;
constexpr int F2_PW = F2_NXMAX + 4;       // row pitch of a progress plane: x = -1 at 1, x = 0 at 2 (pairs stay 16-byte aligned)
constexpr int F2_CPL = F2_SR * F2_PW;     // doubles per progress plane
constexpr int F2_NPL = F2_NR * F2_NXMAX;  // doubles per normal-component plane
constexpr int F2_LOADS = (F2_SR * (F2_NXMAX / 2) + F2_THREADS - 1) / F2_THREADS;   // x-pairs a thread loads per plane (3)
constexpr int F2_RING = 4;               // progress planes: three are read by a step, the fourth is being staged for the next
constexpr size_t F2_SMEM = (size_t)(F2_RING * F2_CPL + 7 * F2_NPL) * sizeof(double);

#ifdef PA_HOST_EMULATION
__device__ __forceinline__ double2 ldg2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double ldg1(const double* p) { return *p; }
#else
__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ double ldg1(const double* p) { return __ldg(p); }
#endif

// LNXP >= 0: every box of the launch is 2 << LNXP cells wide, so (row, pair) of a flattened index are a shift and a mask.  The
// first version divided by the runtime pair count in every loop: ncu showed 4.0 G instructions with IMAD / ISETP / IABS / MUFU.RCP /
// I2F at the top of the mix (integer division), FP64 pipe 27 %.  LNXP < 0 keeps the division (any even width).
template <int LNXP>
__global__ void __launch_bounds__(F2_THREADS, 1) k_curv_f2(const PaTile* __restrict__ tiles, GridArgs ga, StencilExtra ex) {
    PA_DYN_SMEM(smem_raw);
    double* const C = reinterpret_cast<double*>(smem_raw);       // [F2_RING][F2_SR][F2_PW] progress ring
    double* const NXs = C + F2_RING * F2_CPL;                          // [2][F2_NR][128]
    double* const NYs = NXs + 2 * F2_NPL;                        // [2][F2_NR][128]
    double* const NZs = NYs + 2 * F2_NPL;                        // [3][F2_NR][128]

    const PaTile t = tiles[blockIdx.x];
    const LevArgs& L = ga.L[t.lev];
    const PaBoxDev bx = L.boxes[t.box];
    const PaLayDev li = L.lay_in[t.box], lo = L.lay_out[t.box];
    const PaNbr nb = L.nbr[t.box];
    const int nx = bx.n[0], nyb = bx.n[1], nzb = bx.n[2];
    const int nxp = LNXP >= 0 ? (1 << (LNXP >= 0 ? LNXP : 0)) : (nx >> 1);
    auto split = [&](int p, int& r, int& q) {
        if (LNXP >= 0) { r = p >> (LNXP >= 0 ? LNXP : 0); q = p & ((1 << (LNXP >= 0 ? LNXP : 0)) - 1); }
        else { r = p / nxp; q = p - r * nxp; }
    };
    const int KR = t.ny, NR = t.ny + 2, SR = t.ny + 4;
    const int nplanes = t.nz + 4;                                // scalar planes z0-2 .. z0+nz+1
    const int yS0 = t.y0 - 2, zS0 = t.z0 - 2;                    // first scalar row / plane (box-relative, >= -1)
    const double pmin = ex.pmin, pinv = ex.inv;
    const double dxi = L.dxi[0], dyi = L.dxi[1], dzi = L.dxi[2];
    const int tid = threadIdx.x;

    // ---- sources of the scalar: own slab, or a linked neighbour's slab read in place ----
    const int c0 = L.in_comp;
    const int xb = li.ng + li.xoff;
    const double* const own = L.in + li.off + (long long)li.ng * li.PS + (long long)li.ng * li.P + xb;      // cell (0, 0, 0)
    const double* lnk[6];
#pragma unroll
    for (int f = 0; f < 6; ++f) {
        const PaNbrFace F = nb.f[f];
        lnk[f] = nullptr;
        if (F.nb >= 0) {
            const PaPeerSlab ps = L.peers[F.rank];
            const PaLayDev ln = L.lay_in[F.nb];
            lnk[f] = ps.base + (long long)c0 * ps.cs + ln.off + (long long)(F.rel[2] + ln.ng) * ln.PS + (long long)(F.rel[1] + ln.ng) * ln.P +
                     (F.rel[0] + ln.ng + ln.xoff);                                                         // my cell (0, 0, 0) in the neighbour's slab
        }
    }
    // row (y, z) of the scalar: pointer to its x = 0 element and whether it holds the RAW scalar (valid cells of this box or of a
    // linked neighbour) or progress values already (this box's materialised ghost cells, written in progress space by the fill)
    auto row_src = [&](int y, int z, bool& raw) -> const double* {
        const double* base = own;
        raw = true;
        if (z < 0) { if (lnk[2]) base = lnk[2]; else raw = false; }
        else if (z >= nzb) { if (lnk[5]) base = lnk[5]; else raw = false; }
        else if (y < 0) { if (lnk[1]) base = lnk[1]; else raw = false; }
        else if (y >= nyb) { if (lnk[4]) base = lnk[4]; else raw = false; }
        return base + (long long)z * li.PS + (long long)y * li.P;
    };

    // ---- the loads of one plane: F2_LOADS x-pairs per thread and, for threads 0 .. 2 SR - 1, one x-ghost cell ----
    double2 pre[F2_LOADS];
    double preg = 0.0;
    unsigned praw = 0;                                           // bit i: pre[i] is raw; bit 8: preg is raw
    auto issue_loads = [&](int ps) {
        const int z = zS0 + ps;
        praw = 0;
#pragma unroll
        for (int i = 0; i < F2_LOADS; ++i) {
            const int p = tid + i * F2_THREADS;
            pre[i] = make_double2(0.0, 0.0);
            if (p < SR * nxp) {
                int rs, q;
                split(p, rs, q);
                bool raw;
                const double* src = row_src(yS0 + rs, z, raw);
                pre[i] = ldg2(src + 2 * q);
                if (raw) praw |= 1u << i;
            }
        }
        if (tid < 2 * SR) {
            const int rs = tid >> 1, hi = tid & 1;
            const int y = yS0 + rs;
            bool raw;
            const double* src = row_src(y, z, raw);
            const bool inbox = (y >= 0) & (y < nyb) & (z >= 0) & (z < nzb);
            if (inbox) {                                         // the x faces of a valid row: linked neighbour, or own ghost cell
                const double* l = lnk[hi ? 3 : 0];
                if (l) src = l + (long long)z * li.PS + (long long)y * li.P; else raw = false;
            }
            preg = ldg1(src + (hi ? nx : -1));
            if (raw) praw |= 1u << 8;
        }
    };
    auto store_loads = [&](int ps) {                             // normalise once (curvature.cpp:316-320) and stage
        double* Cs = C + (ps % F2_RING) * F2_CPL;
#pragma unroll
        for (int i = 0; i < F2_LOADS; ++i) {
            const int p = tid + i * F2_THREADS;
            if (p < SR * nxp) {
                int rs, q;
                split(p, rs, q);
                double2 v = pre[i];
                if (praw & (1u << i)) { v.x = (v.x - pmin) * pinv; v.y = (v.y - pmin) * pinv; }
                *reinterpret_cast<double2*>(Cs + rs * F2_PW + 2 + 2 * q) = v;
            }
        }
        if (tid < 2 * SR) {
            const int rs = tid >> 1, hi = tid & 1;
            double v = preg;
            if (praw & (1u << 8)) v = (v - pmin) * pinv;
            Cs[rs * F2_PW + (hi ? nx + 2 : 1)] = v;
        }
    };

    // ---- ownership of c and n: the item that holds the cell as a K row / plane; the box's outermost rows / planes go with the
    //      first / last item ----
    const int wy0 = (t.y0 == 1) ? 0 : t.y0, wy1 = (t.y0 + t.ny == nyb - 1) ? nyb - 1 : t.y0 + t.ny - 1;
    const int wz0 = (t.z0 == 1) ? 0 : t.z0, wz1 = (t.z0 + t.nz == nzb - 1) ? nzb - 1 : t.z0 + t.nz - 1;
    const long long cs_out = L.cs_out;
    double* const out_n = L.out;
    double* const out_c = ex.cout[t.lev];
    double* const out_k = ex.kout[t.lev];
    double* const aux = ex.aux[t.lev];
    const long long cg = ex.cs_aux[t.lev];
    const bool do_thr = ex.do_threshold != 0;
    const double thr_lo = ex.threshold, thr_hi = 1.0 - ex.threshold;
    const long long obase = lo.off + (long long)lo.ng * lo.PS + (long long)lo.ng * lo.P + (lo.ng + lo.xoff);   // output cell (0, 0, 0)

    issue_loads(0);
    for (int ps = 0; ps < nplanes; ++ps) {
        store_loads(ps);
        if (ps + 1 < nplanes) issue_loads(ps + 1);               // in flight across the two phases below
        __syncthreads();                                         // plane ps staged; the K phase of the step before is done
        if (ps >= 2) {
            // ---- flame normal of plane zn = zS0 + ps - 1 from the progress planes ps-2, ps-1, ps ----
            const double* C0 = C + ((ps - 2) % F2_RING) * F2_CPL;
            const double* C1 = C + ((ps - 1) % F2_RING) * F2_CPL;
            const double* C2 = C + (ps % F2_RING) * F2_CPL;
            double* nxw = NXs + (ps & 1) * F2_NPL;
            double* nyw = NYs + (ps & 1) * F2_NPL;
            double* nzw = NZs + ((ps - 1) % 3) * F2_NPL;
            const int zn = zS0 + ps - 1;
            const bool wplane = (zn >= wz0) & (zn <= wz1);
#pragma unroll 2
            for (int p = tid; p < NR * nxp; p += F2_THREADS) {
                int rn, q;
                split(p, rn, q);
                const int o = (rn + 1) * F2_PW + 2 + 2 * q;
                const double2 c = lds2(C1 + o);
                const double xm = C1[o - 1], xp = C1[o + 2];
                const double2 ym = lds2(C1 + o - F2_PW), yp = lds2(C1 + o + F2_PW);
                const double2 zm = lds2(C0 + o), zp = lds2(C2 + o);
                const double ax = cdiff(dxi, xm, c.x, c.y), ay = cdiff(dxi, c.x, c.y, xp);
                const double bx0 = cdiff(dyi, ym.x, c.x, yp.x), by0 = cdiff(dyi, ym.y, c.y, yp.y);
                const double g0 = cdiff(dzi, zm.x, c.x, zp.x), g1 = cdiff(dzi, zm.y, c.y, zp.y);
                double r0[3], r1[3];
                normal_pair(ax, bx0, g0, ay, by0, g1, r0, r1);    // curvature.cpp:467-502
                const int no = rn * F2_NXMAX + 2 * q;
                *reinterpret_cast<double2*>(nxw + no) = make_double2(r0[0], r1[0]);
                *reinterpret_cast<double2*>(nyw + no) = make_double2(r0[1], r1[1]);
                *reinterpret_cast<double2*>(nzw + no) = make_double2(r0[2], r1[2]);
                const int y = t.y0 - 1 + rn;
                if (wplane & (y >= wy0) & (y <= wy1)) {
                    const long long oo = obase + (long long)zn * lo.PS + (long long)y * lo.P + 2 * q;
                    stg2(out_c + oo, c.x, c.y);                                        // Progress (curvature.cpp:310-321)
                    stg2(out_n + oo, r0[0], r1[0]);
                    stg2(out_n + oo + cs_out, r0[1], r1[1]);
                    stg2(out_n + oo + 2 * cs_out, r0[2], r1[2]);
                    if (aux) { stg2(aux + oo, ax, ay); stg2(aux + oo + cg, bx0, by0); stg2(aux + oo + 2 * cg, g0, g1); }
                }
            }
        }
        __syncthreads();                                         // n of plane zn staged
        if (ps >= 4) {
            // ---- K of plane zk = zS0 + ps - 2: n_x / n_y of that plane (written one step ago), n_z of zk-1, zk, zk+1 ----
            const double* nxr = NXs + ((ps - 1) & 1) * F2_NPL;
            const double* nyr = NYs + ((ps - 1) & 1) * F2_NPL;
            const double* z0p = NZs + ((ps - 3) % 3) * F2_NPL;
            const double* z1p = NZs + ((ps - 2) % 3) * F2_NPL;
            const double* z2p = NZs + ((ps - 1) % 3) * F2_NPL;
            const double* Ck = C + ((ps - 2) % F2_RING) * F2_CPL;      // not the slot the next step stages into: (ps + 1) % 4 = (ps - 3) % 4
            const int zk = zS0 + ps - 2;
#pragma unroll 2
            for (int p = tid; p < KR * nxp; p += F2_THREADS) {
                int rk, q;
                split(p, rk, q);
                const int no = (rk + 1) * F2_NXMAX + 2 * q;
                const double2 a = lds2(nxr + no);
                const double am = nxr[no - 1], ap = nxr[no + 2];       // out of the row for the first / last pair: those cells are not stored
                const double2 bm = lds2(nyr + no - F2_NXMAX), b = lds2(nyr + no), bp = lds2(nyr + no + F2_NXMAX);
                const double2 gm = lds2(z0p + no), g = lds2(z1p + no), gp = lds2(z2p + no);
                const double dx0 = cdiff(dxi, am, a.x, a.y), dx1 = cdiff(dxi, a.x, a.y, ap);
                const double dy0 = cdiff(dyi, bm.x, b.x, bp.x), dy1 = cdiff(dyi, bm.y, b.y, bp.y);
                const double dz0 = cdiff(dzi, gm.x, g.x, gp.x), dz1 = cdiff(dzi, gm.y, g.y, gp.y);
                double k0 = 0.5 * (((0.0 + dx0) + dy0) + dz0);                       // curvature.cpp:505-547
                double k1 = 0.5 * (((0.0 + dx1) + dy1) + dz1);
                if (do_thr) {                                                          // :549-567 (K only; n is clipped afterwards)
                    const double2 pc = lds2(Ck + (rk + 2) * F2_PW + 2 + 2 * q);
                    if (pc.x < thr_lo || pc.x > thr_hi) k0 = 0.0;
                    if (pc.y < thr_lo || pc.y > thr_hi) k1 = 0.0;
                }
                double* pk = out_k + obase + (long long)zk * lo.PS + (long long)(t.y0 + rk) * lo.P + 2 * q;
                const bool first = q == 0, last = q == nxp - 1;                     // x = 0 and x = nx - 1 belong to k_div_shell
                if (!first && !last) stg2(pk, k0, k1);
                else { if (!first) pk[0] = k0; if (!last) pk[1] = k1; }
            }
        }
    }
}

}  // namespace

int curv_f2_rows() { return F2_KR; }
int curv_f2_max_nx() { return F2_NXMAX; }

cudaError_t launch_curv_f2(const PaTile* tiles, int ntiles, int lnxp, const GridArgs& ga, const StencilExtra& ex, cudaStream_t st) {
    if (ntiles <= 0) return cudaSuccess;
    static std::map<std::pair<int, int>, bool> configured;
    static std::mutex mu;
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    auto go = [&](auto kern, int key) -> cudaError_t {
        {
            std::lock_guard<std::mutex> lock(mu);
            if (!configured[std::make_pair(dev, key)]) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F2_SMEM);
                if (e != cudaSuccess) return e;
                configured[std::make_pair(dev, key)] = true;
            }
        }
        PA_LAUNCH(ntiles, F2_THREADS, F2_SMEM, st, kern)(tiles, ga, ex);
        return cudaGetLastError();
    };
    cudaError_t e;
    switch (lnxp) {
    case 6: e = go(k_curv_f2<6>, 6); break;      // 128-wide boxes
    case 5: e = go(k_curv_f2<5>, 5); break;      // 64
    case 4: e = go(k_curv_f2<4>, 4); break;      // 32
    case 3: e = go(k_curv_f2<3>, 3); break;      // 16
    default: e = go(k_curv_f2<-1>, -1); break;
    }
    ++g_launches;
    return e;
}

}  // namespace pa
