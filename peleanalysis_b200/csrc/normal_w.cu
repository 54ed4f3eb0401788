// normal_w.cu -- flame normal without shared memory and without barriers (PA_NORMAL_W=1): S -> Progress, n = G / -max(1e-14, |G|),
// the work of MODE_NORMAL_S (curvature.cpp:310-321, 426-502).
//
// Why another kernel.  Three implementations of this pass -- the TMA ring with a register z history (stencil_tma.cu), progress
// planes in shared memory fed from registers and fed by cp.async (curv_f3.cu) -- all end at 3.7-3.9 ms on the north-star
// hierarchy with 28-33 % of the FP64 pipe and 30-37 % of the issue slots: they synchronise 8-17 warps once per plane and their
// ~110 registers per thread cap the SM at 16 warps.  The plain-load fallback kernel (k_stencil_simple, 76 registers, no
// synchronisation at all) issues 2.7 times as many instructions per cell and still reaches 61 % of the issue slots
// (profiles/r02_launches_simple_route.csv).  This kernel keeps that shape and removes its instructions (measured: no faster --
// 4.3 ms, see the end of this comment):
//   * a warp owns one row of one x strip (at most 32 pairs) and sweeps along z; nothing is shared between warps, so nothing waits;
//   * a thread owns one x pair: the centre values of the planes z-1, z, z+1 stay in registers (normalised once, on arrival),
//     x neighbours come from the adjacent lanes by shuffle, the y neighbours are loaded (they are the centre rows of the
//     neighbouring warps: L1 / L2 hits) and normalised on arrival;
//   * pointers are set up once per item and advanced by the plane stride; the sqrt -> reciprocal -> quotient chains are the
//     branch-free forms of stencil_dev.cuh (two per thread).
// Arithmetic is the reference's expression order with separate IEEE multiplies and adds (-fmad=false): bit-identical to the
// other routes.
// Result on a B200 (profiles/r02_ncu_normal_w_summary.txt): 275 instructions per pair and plane, 23 warps per SM, and 4.27 ms
// against 3.73 ms for MODE_NORMAL_S -- the warps wait for their loads (long scoreboard 9-13 cycles per issued instruction)
// whether the operands are requested in the same step, one step ahead, prefetched into L2 eight planes ahead or kept four
// steps in flight by cp.async.  Opt-in, parity-tested; not the default.  (The ablation of the plane-staged kernel at the end of
// round 2 explains it: the pass is paced by its write-heavy memory traffic, not by arithmetic or by how its loads are issued --
// profiles/r02_ablate_normal_f3.txt)
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"
#include "stencil_dev.cuh"

namespace pa {

namespace {

constexpr int NW_ROWS = 8;                // rows per item = warps per CTA
constexpr int NW_THREADS = 32 * NW_ROWS;
constexpr int NW_KQ = 32;                 // pairs per strip at most (one lane each)

#ifdef PA_HOST_EMULATION
__device__ __forceinline__ double2 ldg2_nw(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double ldg1_nw(const double* p) { return *p; }
__device__ __forceinline__ void prefetch_l2(const double*) {}
#else
__device__ __forceinline__ void prefetch_l2(const double* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ double2 ldg2_nw(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ double ldg1_nw(const double* p) { return __ldg(p); }
#endif

// Warp-private staging ring (RING): every lane copies ITS OWN operands of the steps up to NW_DEPTH ahead into shared memory with
// cp.async and reads them back itself -- no other thread ever touches them, so no barrier and no __syncwarp is needed.  Per warp
// and step: centre pairs of plane z+1 (512 bytes), the y-1 / y+1 pairs of plane z (2 x 512) and the strip's edge cells (256).
constexpr int NW_DEPTH = 4;
constexpr int NW_SLOT = 512 * 3 + 256;    // bytes per warp and step
constexpr int NW_SMEM = NW_DEPTH * NW_SLOT * NW_ROWS;
#ifdef PA_HOST_EMULATION
struct NwSm { unsigned char* p; };
__device__ __forceinline__ NwSm nw_sm(unsigned char* raw, int off) { return NwSm{raw + off}; }
__device__ __forceinline__ void nw_cpa16(NwSm d, int off, const double* src) { cuemu::cp_async_16(d.p + off, src); }
__device__ __forceinline__ void nw_cpa8(NwSm d, int off, const double* src) { cuemu::cp_async_n(d.p + off, src, 8); }
__device__ __forceinline__ void nw_commit() { cuemu::cp_async_commit(); }
__device__ __forceinline__ void nw_wait(int n) { cuemu::cp_async_wait(n); }
__device__ __forceinline__ double2 nw_ld2(NwSm d, int off) { return *reinterpret_cast<const double2*>(d.p + off); }
__device__ __forceinline__ double nw_ld1(NwSm d, int off) { return *reinterpret_cast<const double*>(d.p + off); }
#else
struct NwSm { uint32_t a; };
__device__ __forceinline__ NwSm nw_sm(unsigned char* raw, int off) { return NwSm{smem_u32(raw) + (uint32_t)off}; }
__device__ __forceinline__ void nw_cpa16(NwSm d, int off, const double* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d.a + (uint32_t)off), "l"(src) : "memory"); }
__device__ __forceinline__ void nw_cpa8(NwSm d, int off, const double* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d.a + (uint32_t)off), "l"(src) : "memory"); }
__device__ __forceinline__ void nw_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void nw_wait(int n) {
    switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}
__device__ __forceinline__ double2 nw_ld2(NwSm d, int off) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(d.a + (uint32_t)off));
    return v;
}
__device__ __forceinline__ double nw_ld1(NwSm d, int off) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(d.a + (uint32_t)off));
    return v;
}
#endif

// MINB: CTAs per SM the register allocation aims at
template <int MINB, bool RING>
__global__ void __launch_bounds__(NW_THREADS, MINB) k_normal_w(const PaTile* __restrict__ tiles, GridArgs ga, StencilExtra ex, int pfd) {
    const PaTile t = tiles[blockIdx.x];
    const int lev = t.lev & 0xff, xq0 = (t.lev >> 8) & 0xff, KQ = (t.lev >> 16) & 0xff;   // level, first pair and pairs of the strip
    const LevArgs& L = ga.L[lev];
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
    if (row >= t.ny) return;                                     // warps are independent: a short item simply uses fewer of them
    const PaBoxDev bx = L.boxes[t.box];
    const PaLayDev li = L.lay_in[t.box], lo = L.lay_out[t.box];
    const int nx = bx.n[0], nyb = bx.n[1], nzb = bx.n[2];
    const int y = t.y0 + row;
    const bool act = lane < KQ;
    const int q = xq0 + (act ? lane : 0);                        // idle lanes shadow lane 0 (they take part in the shuffles only)
    const int x = 2 * q;
    const long long PS = li.PS;

    // ---- sources: own slab, or a linked neighbour's slab read in place.  A value is RAW (the scalar: valid cells of this box or
    //      of a linked neighbour) or already in progress space (this box's materialised ghost cells) ----
    const double* const own = L.in + li.off + (long long)li.ng * li.PS + (long long)li.ng * li.P + (li.ng + li.xoff);   // cell (0, 0, 0)
    auto link = [&](int f) -> const double* {
        const PaNbrFace F = L.nbr[t.box].f[f];
        if (F.nb < 0) return nullptr;
        const PaPeerSlab ps = L.peers[F.rank];
        const PaLayDev ln = L.lay_in[F.nb];
        return ps.base + (long long)L.in_comp * ps.cs + ln.off + (long long)(F.rel[2] + ln.ng) * ln.PS + (long long)(F.rel[1] + ln.ng) * ln.P +
               (F.rel[0] + ln.ng + ln.xoff);                     // my cell (0, 0, 0) in the neighbour's slab
    };
    const long long rowoff = (long long)y * li.P + x;
    unsigned fl = 0;                                             // 1: y-1 row raw, 2: y+1 row raw, 4: edge cell raw, 8: left edge lane, 16: right edge lane,
                                                                 // 32 / 64: plane -1 / plane nz raw
    const double* pc = own + rowoff + (long long)(t.z0 - 1) * PS;      // centre pair, plane about to be loaded
    const double* pym = pc - li.P + PS;                                 // y-1 / y+1 pairs of the plane whose n is computed next (z0)
    const double* pyp = pc + li.P + PS;
    fl |= 3u;
    if (y == 0) {
        const double* l = link(1);
        if (l) pym = l + rowoff - li.P + (long long)t.z0 * PS; else fl &= ~1u;
    }
    if (y == nyb - 1) {
        const double* l = link(4);
        if (l) pyp = l + rowoff + li.P + (long long)t.z0 * PS; else fl &= ~2u;
    }
    // the cell left of the strip (lane 0) / right of it (the strip's last lane): the same row of this box -- a valid cell of the
    // next strip or a materialised ghost cell -- or, across a linked x face, the neighbour's valid cell
    const double* pe = pc + PS;                                  // plane z0
    if (lane == 0) {
        fl |= 8u | 4u;
        pe = pe - 1;
        if (q == 0) {
            const double* l = link(0);
            if (l) pe = l + rowoff - 1 + (long long)t.z0 * PS; else fl &= ~4u;
        }
    }
    if (act && lane == KQ - 1) {                                 // (a strip has at least two pairs: never lane 0 as well)
        fl |= 16u | 4u;
        pe = pe + 2;
        if (x + 2 >= nx) {
            const double* l = link(3);
            if (l) pe = l + rowoff + 2 + (long long)t.z0 * PS; else fl &= ~4u;
        }
    }
    // the box's z ghost planes: the z neighbour's valid plane, or own ghost cells
    const double* pzlo = pc;                                     // plane z0 - 1 if that is -1
    const double* pzhi = nullptr;                                // plane nz
    if (t.z0 == 0) {
        const double* l = link(2);
        if (l) { pzlo = l + rowoff - PS; fl |= 32u; }
    } else {
        fl |= 32u;                                               // plane z0 - 1 is a valid plane of this box
    }
    if (t.z0 + t.nz == nzb) {
        const double* l = link(5);
        pzhi = (l ? l : own) + rowoff + (long long)nzb * PS;
        if (l) fl |= 64u;
    }

    const double pmin = ex.pmin, pinv = ex.inv;
    const double dxi = L.dxi[0], dyi = L.dxi[1], dzi = L.dxi[2];
    auto norm2 = [&](double2 v, bool raw) -> double2 {           // curvature.cpp:316-320
        if (raw) { v.x = (v.x - pmin) * pinv; v.y = (v.y - pmin) * pinv; }
        return v;
    };
    const long long cs_out = L.cs_out;
    double* pon = L.out + lo.off + (long long)(lo.ng + t.z0) * lo.PS + (long long)(lo.ng + y) * lo.P + (lo.ng + lo.xoff + x);
    double* poc = ex.cout[lev] + (pon - L.out);
    double* poa = ex.aux[lev] ? ex.aux[lev] + (pon - L.out) : nullptr;
    const long long cg = ex.cs_aux[lev];
    const long long oPS = lo.PS;

    if constexpr (RING) {
        PA_DYN_SMEM(smem_raw);
        const NwSm ring = nw_sm(smem_raw, row * (NW_DEPTH * NW_SLOT));
        const int depth = pfd < 1 ? 1 : (pfd > NW_DEPTH ? NW_DEPTH : pfd);
        const bool edge = (fl & 24u) != 0;
        double2 zm = norm2(ldg2_nw(pzlo), (fl & 32u) != 0);
        pc += PS;
        double2 c = norm2(ldg2_nw(pc), true);
        pc += PS;                                                // plane z0 + 1: the centre operand of step 0
        auto issue = [&](int k) {                                // operands of step k -> slot k % depth; one group per step
            if (k < t.nz) {
                const int so = (k % depth) * NW_SLOT;
                nw_cpa16(ring, so + lane * 16, (t.z0 + k + 1 == nzb) ? pzhi : pc);
                nw_cpa16(ring, so + 512 + lane * 16, pym);
                nw_cpa16(ring, so + 1024 + lane * 16, pyp);
                if (edge) nw_cpa8(ring, so + 1536 + lane * 8, pe);
                pc += PS; pym += PS; pyp += PS; pe += PS;
            }
            nw_commit();
        };
        for (int d = 0; d < depth; ++d) issue(d);
        for (int k = 0; k < t.nz; ++k) {
            const bool hi = t.z0 + k + 1 == nzb;
            nw_wait(depth - 1);                                  // this lane's copies for step k have landed
            const int so = (k % depth) * NW_SLOT;
            double2 zp = nw_ld2(ring, so + lane * 16), ym = nw_ld2(ring, so + 512 + lane * 16), yp = nw_ld2(ring, so + 1024 + lane * 16);
            double e = edge ? nw_ld1(ring, so + 1536 + lane * 8) : 0.0;
            issue(k + depth);
            zp = norm2(zp, hi ? (fl & 64u) != 0 : true);
            ym = norm2(ym, (fl & 1u) != 0);
            yp = norm2(yp, (fl & 2u) != 0);
            double xm = __shfl_up_sync(0xffffffffu, c.y, 1);
            double xp = __shfl_down_sync(0xffffffffu, c.x, 1);
            if (edge) {
                if (fl & 4u) e = (e - pmin) * pinv;
                if (fl & 8u) xm = e; else xp = e;
            }
            const double ax = cdiff(dxi, xm, c.x, c.y), ay = cdiff(dxi, c.x, c.y, xp);
            const double bx0 = cdiff(dyi, ym.x, c.x, yp.x), by0 = cdiff(dyi, ym.y, c.y, yp.y);
            const double g0 = cdiff(dzi, zm.x, c.x, zp.x), g1 = cdiff(dzi, zm.y, c.y, zp.y);
            double r0[3], r1[3];
            normal_pair(ax, bx0, g0, ay, by0, g1, r0, r1);        // curvature.cpp:467-502
            if (act) {
                stg2(poc, c.x, c.y);                              // Progress (curvature.cpp:310-321)
                stg2(pon, r0[0], r1[0]);
                stg2(pon + cs_out, r0[1], r1[1]);
                stg2(pon + 2 * cs_out, r0[2], r1[2]);
                if (poa) { stg2(poa, ax, ay); stg2(poa + cg, bx0, by0); stg2(poa + 2 * cg, g0, g1); }
            }
            zm = c; c = zp;
            pon += oPS; poc += oPS;
            if (poa) poa += oPS;
        }
        return;
    }
    // planes z0 - 1 and z0; the loads of a step are issued one step ahead (measured without: 12.8 cycles of long-scoreboard
    // stall per issued instruction -- 24 independent warps per SM do not hide a DRAM round trip per plane on their own)
    double2 zm = norm2(ldg2_nw(pzlo), (fl & 32u) != 0);
    pc += PS;
    double2 c = norm2(ldg2_nw(pc), true);
    pc += PS;
    const bool edge = (fl & 24u) != 0;
    double2 zp_n = ldg2_nw((t.z0 + 1 == nzb) ? pzhi : pc), ym_n = ldg2_nw(pym), yp_n = ldg2_nw(pyp);
    double e_n = edge ? ldg1_nw(pe) : 0.0;
    for (int k = 0; k < t.nz; ++k) {
        const int z = t.z0 + k;
        const bool hi = z + 1 == nzb;
        double2 zp = zp_n, ym = ym_n, yp = yp_n;
        double e = e_n;
        pc += PS; pym += PS; pyp += PS; pe += PS;
        // The register prefetch covers one step (~2 us for a warp that shares its scheduler with five others): under load a DRAM
        // round trip takes as long, and 61 % of the stall samples sat on the first use of zp.  A prefetch into L2 a few planes
        // further ahead costs no registers and turns that load into an L2 hit.  Own slab only (planes <= nz exist there).
        if (pfd > 0 && z + 2 + pfd <= nzb) prefetch_l2(pc + (long long)pfd * PS);
        if (k + 1 < t.nz) {                                      // next step's planes
            zp_n = ldg2_nw((z + 2 == nzb) ? pzhi : pc);
            ym_n = ldg2_nw(pym);
            yp_n = ldg2_nw(pyp);
            if (edge) e_n = ldg1_nw(pe);
        }
        zp = norm2(zp, hi ? (fl & 64u) != 0 : true);
        ym = norm2(ym, (fl & 1u) != 0);
        yp = norm2(yp, (fl & 2u) != 0);
        double xm = __shfl_up_sync(0xffffffffu, c.y, 1);
        double xp = __shfl_down_sync(0xffffffffu, c.x, 1);
        if (edge) {
            if (fl & 4u) e = (e - pmin) * pinv;
            if (fl & 8u) xm = e; else xp = e;
        }
        const double ax = cdiff(dxi, xm, c.x, c.y), ay = cdiff(dxi, c.x, c.y, xp);
        const double bx0 = cdiff(dyi, ym.x, c.x, yp.x), by0 = cdiff(dyi, ym.y, c.y, yp.y);
        const double g0 = cdiff(dzi, zm.x, c.x, zp.x), g1 = cdiff(dzi, zm.y, c.y, zp.y);
        double r0[3], r1[3];
        normal_pair(ax, bx0, g0, ay, by0, g1, r0, r1);            // curvature.cpp:467-502
        if (act) {
            stg2(poc, c.x, c.y);                                  // Progress (curvature.cpp:310-321)
            stg2(pon, r0[0], r1[0]);
            stg2(pon + cs_out, r0[1], r1[1]);
            stg2(pon + 2 * cs_out, r0[2], r1[2]);
            if (poa) { stg2(poa, ax, ay); stg2(poa + cg, bx0, by0); stg2(poa + 2 * cg, g0, g1); }
        }
        zm = c; c = zp;
        pon += oPS; poc += oPS;
        if (poa) poa += oPS;
    }
}

}  // namespace

int normal_w_rows() { return NW_ROWS; }
int normal_w_strip_pairs() { return NW_KQ; }

template <int MINB, bool RING>
static cudaError_t launch_nw(const PaTile* tiles, int ntiles, const GridArgs& ga, const StencilExtra& ex, int pfd, cudaStream_t st) {
    const size_t smem = RING ? (size_t)NW_SMEM : 0;
#ifndef PA_HOST_EMULATION
    if (RING) {
        cudaError_t e = cudaFuncSetAttribute(k_normal_w<MINB, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_normal_w<MINB, RING>, cudaFuncAttributePreferredSharedMemoryCarveout, 86);   // MINB CTAs of 56 KB: 196 KB
        if (e != cudaSuccess) return e;
    }
#endif
    PA_LAUNCH(ntiles, NW_THREADS, smem, st, k_normal_w<MINB, RING>)(tiles, ga, ex, pfd);
    return cudaGetLastError();
}

cudaError_t launch_normal_w(const PaTile* tiles, int ntiles, const GridArgs& ga, const StencilExtra& ex, cudaStream_t st) {
    if (ntiles <= 0) return cudaSuccess;
    const char* e = getenv("PA_NW_CTAS");                        // CTAs per SM the build aims at: 2 (128 registers), 3 (80, default), 4 (64)
    const int ctas = e ? atoi(e) : 3;
    // Measured on a B200 (profiles/r02_ncu_normal_w_summary.txt): operands one step ahead in registers 7.04 ms per curvature step,
    // the cp.async ring 7.38, L2 prefetch no gain -- so the register form without L2 prefetch is what PA_NORMAL_W=1 runs.
    const char* er = getenv("PA_NW_RING");                       // 1: the warp-private cp.async ring instead of the register prefetch
    const bool ring = er && er[0] == '1';
    const char* ep = getenv("PA_NW_PF");                         // ring: steps in flight (1 .. 4); registers: planes prefetched into L2 ahead (0: off)
    const int pfd = ep ? std::max(0, atoi(ep)) : (ring ? 4 : 0);
    cudaError_t err;
    if (ring) err = ctas == 2 ? launch_nw<2, true>(tiles, ntiles, ga, ex, pfd, st) : ctas == 4 ? launch_nw<4, true>(tiles, ntiles, ga, ex, pfd, st)
                                                                                              : launch_nw<3, true>(tiles, ntiles, ga, ex, pfd, st);
    else err = ctas == 2 ? launch_nw<2, false>(tiles, ntiles, ga, ex, pfd, st) : ctas == 4 ? launch_nw<4, false>(tiles, ntiles, ga, ex, pfd, st)
                                                                                           : launch_nw<3, false>(tiles, ntiles, ga, ex, pfd, st);
    ++g_launches;
    return err;
}

}  // namespace pa
