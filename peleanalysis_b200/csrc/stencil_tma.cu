// stencil_tma.cu -- the hot stencil kernels as a TMA-staged shared-memory pipeline for sm_100a.
//
// One CTA owns a tile = rows [y0,y0+ny) x planes [z0,z0+nz) of one box, full x extent, and sweeps it along z
// (2.5-D blocking).  Because a tile spans whole rows, the (ny+2) rows of one z-plane -- y halo and x ghosts
// included -- are ONE contiguous, 16-byte aligned run in the box's slab, so each plane is staged with a single
// cp.async.bulk (TMA, SASS UBLKCP) that completes on an mbarrier.  A producer warp keeps STAGES planes in flight;
// eight consumer warps read a plane exactly once (128-bit LDS for the centre and y neighbours, warp shuffles for
// the x neighbours), keep the z-1 / z / z+1 centre values in a register queue, and release the stage back to the
// producer through a second mbarrier -- no __syncthreads in the steady state.
// Arithmetic is the reference's expression order with separate IEEE mul/add (-fmad=false): bit-exact.
#include <cstdint>

#include "kernels.cuh"

namespace pa {

namespace {

constexpr int STAGES = 4;
constexpr int CONSUMER_WARPS = 8;
constexpr int CONSUMER_THREADS = CONSUMER_WARPS * 32;
constexpr int THREADS = CONSUMER_THREADS + 32;       // + 1 producer warp
constexpr int MAX_ITEMS = 2;                          // x-pairs per consumer thread per plane
constexpr int TILE_ROWS = 8;                          // default TY
constexpr int MAX_TILE_ROWS = 15;                     // 2 * rows x-ghost lanes must fit in the producer warp next to lane 0

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 8-byte asynchronous global -> shared copy (SASS LDGSTS) and its completion hooked to an mbarrier arrival
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ double cdiff(double dxi, double m, double c, double p) {
    return 0.5 * (dxi * (c - m) + dxi * (p - c));
}
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void stg2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

template <int MODE> struct ModeTraits;
template <> struct ModeTraits<MODE_GRAD> { static constexpr int NIN = 1, NOUT = 4; };
template <> struct ModeTraits<MODE_GRAD3> { static constexpr int NIN = 1, NOUT = 3; };
template <> struct ModeTraits<MODE_NORMAL> { static constexpr int NIN = 1, NOUT = 3; };
template <> struct ModeTraits<MODE_DIV> { static constexpr int NIN = 3, NOUT = 1; };
template <> struct ModeTraits<MODE_NORMAL_S> { static constexpr int NIN = 1, NOUT = 3; };

// per-item register state carried from plane to plane
struct ItemState {
    double2 cm, c0;       // centre values of planes p-2 and p-1 (MODE_DIV: of the z component)
    double2 a0, b0;       // in-plane derivatives of plane p-1 (x and y)
};

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) k_stencil_tma(const PaTile* __restrict__ tiles, GridArgs ga, StencilExtra ex,
                                                         int stage_doubles /* per input component, multiple of 16 */) {
    constexpr int NIN = ModeTraits<MODE>::NIN;
    constexpr int NOUT = ModeTraits<MODE>::NOUT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);                 // [STAGES][NIN][stage_doubles]
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(16) double xg_s[STAGES][2][16];              // x ghosts of linked x faces: [stage][lo/hi][row]

    const PaTile t = tiles[blockIdx.x];
    const LevArgs& L = ga.L[t.lev];
    const PaBoxDev bx = L.boxes[t.box];
    const PaLayDev li = L.lay_in[t.box];
    const PaLayDev lo = L.lay_out[t.box];
    const int v = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int rows = t.ny + 2;                         // y0-1 .. y0+ny
    const int plane_elems = rows * li.P;               // contiguous in global memory
    const uint32_t plane_bytes = (uint32_t)plane_elems * 8u;
    const int nplanes = t.nz + 2;                      // z0-1 .. z0+nz
    const double* __restrict__ in0 = L.in + (MODE == MODE_DIV ? 0 : (long long)v * L.cs_in);

    // x faces with a neighbour link: their ghost column is fetched cell by cell from the neighbour by producer lanes
    const bool xlo_link = L.nbr[t.box].f[0].nb >= 0, xhi_link = L.nbr[t.box].f[3].nb >= 0;
    if (threadIdx.x == 0) {
        // a stage is full when the TMA bytes have landed (lane 0's arrive.expect_tx) and every x-ghost lane's copy has
        const uint32_t nfull = 1u + (xlo_link ? t.ny : 0) + (xhi_link ? t.ny : 0);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], nfull); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ===== producer warp =====
        // lane 0 drives the TMA ring.  Linked faces (PaNbrFace) are staged straight from the neighbour box -- this
        // rank's slab or a peer's over NVLink: a z-ghost plane is one bulk copy from the neighbour's plane, a y-ghost
        // row one row copy.  Unlinked faces come from this box's own (materialised) ghost cells inside the main run.
        // lanes 1 .. 2*ny each own one x-ghost cell per plane (side, row) of a linked x face and fetch it with an
        // 8-byte cp.async (LDGSTS) into xg_s; its completion is one more arrival on the stage's full barrier.
        const PaNbr nb = L.nbr[t.box];
        const int c0 = L.in_comp + (MODE == MODE_DIV ? 0 : v);
        // per linked face: address of the neighbour element matching (row 0 of the padded block, plane 0) of
        // component c0, and the component stride of the slab it lives in
        auto link_src = [&](int face, long long& cs) -> const double* {
            const PaNbrFace F = nb.f[face];
            cs = 0;
            if (F.nb < 0) return nullptr;
            const PaPeerSlab ps = L.peers[F.rank];
            const PaLayDev ln = L.lay_in[F.nb];
            cs = ps.cs;
            // neighbour-relative (x, y, z) = own-relative + rel
            return ps.base + (long long)c0 * ps.cs + ln.off + (long long)(F.rel[2] + ln.ng) * ln.PS + (long long)(F.rel[1] + ln.ng) * ln.P + F.rel[0];
        };
        const int nyb = bx.n[1], nzb = bx.n[2];
        if (lane == 0) {
            long long cs_ylo, cs_zlo, cs_yhi, cs_zhi;
            const double* s_ylo = link_src(1, cs_ylo);        // rel[0] == 0 for y and z faces: rows are x-aligned
            const double* s_zlo = link_src(2, cs_zlo);
            const double* s_yhi = link_src(4, cs_yhi);
            const double* s_zhi = link_src(5, cs_zhi);
            const uint32_t row_bytes = (uint32_t)li.P * 8u;
            for (int p = 0; p < nplanes; ++p) {
                const int s = p % STAGES;
                const int n = p / STAGES;
                if (n > 0) mbar_wait(&empty_bar[s], (uint32_t)((n - 1) & 1));
                mbar_expect_tx(&full_bar[s], plane_bytes * NIN);
                const int z = t.z0 - 1 + p;
                const double* zs = (z < 0) ? s_zlo : (z >= nzb ? s_zhi : nullptr);
                const long long zcs = (z < 0) ? cs_zlo : cs_zhi;
#pragma unroll
                for (int c = 0; c < NIN; ++c) {
                    double* dst = sm + ((long long)s * NIN + c) * stage_doubles;
                    if (zs) {
                        tma_load_1d(dst, zs + (long long)c * zcs + (long long)z * li.PS + (long long)(t.y0 - 1) * li.P, plane_bytes, &full_bar[s]);
                        continue;
                    }
                    int r0 = t.y0 - 1, r1 = t.y0 + t.ny;                     // first / last staged row (box-relative)
                    if (r0 < 0 && s_ylo) {
                        tma_load_1d(dst, s_ylo + (long long)c * cs_ylo + (long long)z * li.PS - li.P, row_bytes, &full_bar[s]);
                        r0 = 0;
                    }
                    if (r1 >= nyb && s_yhi) {
                        tma_load_1d(dst + (long long)(rows - 1) * li.P, s_yhi + (long long)c * cs_yhi + (long long)z * li.PS + (long long)nyb * li.P,
                                    row_bytes, &full_bar[s]);
                        r1 = nyb - 1;
                    }
                    tma_load_1d(dst + (long long)(r0 - (t.y0 - 1)) * li.P,
                                in0 + (long long)c * L.cs_in + li.off + (long long)(z + li.ng) * li.PS + (long long)(r0 + li.ng) * li.P,
                                (uint32_t)(r1 - r0 + 1) * row_bytes, &full_bar[s]);
                }
            }
        } else {
            const int idx = lane - 1;
            const int side = idx / t.ny, r = idx - side * t.ny;              // side 0 = x-lo, 1 = x-hi
            if (side < 2 && (side ? xhi_link : xlo_link)) {
                long long cs;
                // own-relative ghost cell (gi, y0 + r, z0 - 1 + p): gi = -1 (lo) or nx (hi)
                const int gi = side ? bx.n[0] : -1;
                const double* src = link_src(side ? 3 : 0, cs) + (long long)(t.z0 - 1) * li.PS + (long long)(t.y0 + r) * li.P + (gi + li.ng + li.xoff);
                for (int p = 0; p < nplanes; ++p) {
                    const int s = p % STAGES;
                    const int n = p / STAGES;
                    if (n > 0) mbar_wait(&empty_bar[s], (uint32_t)((n - 1) & 1));
                    cp_async_8(&xg_s[s][side][r], src + (long long)p * li.PS);
                    cp_async_arrive_noinc(&full_bar[s]);
                }
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int nx = bx.n[0];
    const int nq = (nx + 1) >> 1;
    const int items = nq * t.ny;                       // <= MAX_ITEMS * CONSUMER_THREADS by tile construction
    const double dxi = L.dxi[0], dyi = L.dxi[1], dzi = L.dxi[2];
    const int xbase = li.ng + li.xoff;                 // even
    double* __restrict__ out0 = L.out + (long long)v * NOUT * L.cs_out;

    // MODE_NORMAL_S: staged values are the raw scalar S wherever they come from valid cells (own or a linked neighbour's)
    // and already-normalised progress values c where they are this box's materialised ghost cells.  n?_raw = "the
    // ghost layer of that face is materialised" (face not linked).
    constexpr bool XS = (MODE == MODE_NORMAL_S);
    bool xlo_raw = false, xhi_raw = false, ylo_raw = false, yhi_raw = false, zlo_raw = false, zhi_raw = false;
    if (XS) {
        const PaNbr nb = L.nbr[t.box];
        xlo_raw = nb.f[0].nb < 0; ylo_raw = nb.f[1].nb < 0; zlo_raw = nb.f[2].nb < 0;
        xhi_raw = nb.f[3].nb < 0; yhi_raw = nb.f[4].nb < 0; zhi_raw = nb.f[5].nb < 0;
    }
    const double pmin = ex.pmin, pinv = ex.inv;
    auto prog = [&](double sv) { return (sv - pmin) * pinv; };          // curvature.cpp:316-320

    ItemState st[MAX_ITEMS];
#pragma unroll
    for (int it = 0; it < MAX_ITEMS; ++it) st[it].cm = st[it].c0 = st[it].a0 = st[it].b0 = make_double2(0., 0.);

    for (int p = 0; p < nplanes; ++p) {
        const int s = p % STAGES;
        mbar_wait(&full_bar[s], (uint32_t)((p / STAGES) & 1));
        const double* S0 = sm + (long long)s * NIN * stage_doubles;
#pragma unroll
        for (int it = 0; it < MAX_ITEMS; ++it) {
            const int w = threadIdx.x + it * CONSUMER_THREADS;
            const bool active = w < items;
            const int q = active ? (w % nq) : 0;
            const int r = active ? (w / nq) : 0;
            const int xi = xbase + 2 * q;
            const double* Sc = S0 + (r + 1) * li.P + xi;               // centre row of this item, component 0
            // centre pair of the component whose x derivative we take
            double2 c = lds2(Sc);
            const bool last_odd = (nx & 1) && (q == nq - 1);              // the pair's second cell is the x-hi ghost
            if (XS) {
                // a z-ghost plane of an unlinked face is materialised (already c); everything else staged here is S
                const int z = t.z0 - 1 + p;
                const bool plane_raw = (z < 0 && zlo_raw) || (z >= bx.n[2] && zhi_raw);
                if (!plane_raw) { c.x = prog(c.x); if (!(last_odd && xhi_raw)) c.y = prog(c.y); }
            }
            // x neighbours: from the adjacent lanes when they hold the same row, else from shared memory
            double xm = __shfl_up_sync(0xffffffffu, c.y, 1);
            double xp = __shfl_down_sync(0xffffffffu, c.x, 1);
            if (lane == 0 || q == 0) { xm = Sc[-1]; if (XS && !(q == 0 && xlo_raw)) xm = prog(xm); }
            if (lane == 31 || q == nq - 1 || !active || (w + 1 >= items)) { xp = Sc[2]; if (XS && !(q == nq - 1 && xhi_raw)) xp = prog(xp); }
            if (q == 0 && xlo_link) {                                     // ghost = the linked neighbour's last valid cell
                xm = xg_s[s][0][r]; if (XS) xm = prog(xm);
            }
            if (q == nq - 1 && xhi_link) {
                double xg = xg_s[s][1][r]; if (XS) xg = prog(xg);
                if (nx & 1) c.y = xg;                                     // odd row length: the pair's second cell IS the ghost
                else xp = xg;
            }
            double2 a1, b1, cp;
            a1.x = cdiff(dxi, xm, c.x, c.y);
            a1.y = cdiff(dxi, c.x, c.y, xp);
            if (MODE != MODE_DIV) {
                double2 ym = lds2(Sc - li.P), yp = lds2(Sc + li.P);
                if (XS) {
                    if (!(t.y0 + r == 0 && ylo_raw)) { ym.x = prog(ym.x); ym.y = prog(ym.y); }
                    if (!(t.y0 + r == bx.n[1] - 1 && yhi_raw)) { yp.x = prog(yp.x); yp.y = prog(yp.y); }
                }
                b1.x = cdiff(dyi, ym.x, c.x, yp.x);
                b1.y = cdiff(dyi, ym.y, c.y, yp.y);
                cp = c;
            } else {
                const double* Sy = Sc + stage_doubles;                  // component 1 (n_y)
                const double2 cy = lds2(Sy), ym = lds2(Sy - li.P), yp = lds2(Sy + li.P);
                b1.x = cdiff(dyi, ym.x, cy.x, yp.x);
                b1.y = cdiff(dyi, ym.y, cy.y, yp.y);
                cp = lds2(Sc + 2 * stage_doubles);                      // component 2 (n_z) centre
            }
            // finish plane p-1 (needs centre of p-2, p-1, p) once it is an interior plane of the tile
            if (p >= 2 && active) {
                const int jy = t.y0 + r, kz = t.z0 + p - 2;
                const double g0 = cdiff(dzi, st[it].cm.x, st[it].c0.x, cp.x);
                const double g1 = cdiff(dzi, st[it].cm.y, st[it].c0.y, cp.y);
                const long long o = lo.off + (long long)(kz + lo.ng) * lo.PS + (long long)(jy + lo.ng) * lo.P + (2 * q + lo.ng + lo.xoff);
                const bool two = (2 * q + 1 < nx);
                double r0[4], r1[4];
                const double ax = st[it].a0.x, ay = st[it].a0.y, bx0 = st[it].b0.x, by0 = st[it].b0.y;
                if (MODE == MODE_GRAD) {
                    r0[0] = ax; r0[1] = bx0; r0[2] = g0; r0[3] = sqrt(ax * ax + bx0 * bx0 + g0 * g0);
                    r1[0] = ay; r1[1] = by0; r1[2] = g1; r1[3] = sqrt(ay * ay + by0 * by0 + g1 * g1);
                } else if (MODE == MODE_GRAD3) {
                    r0[0] = ax; r0[1] = bx0; r0[2] = g0;
                    r1[0] = ay; r1[1] = by0; r1[2] = g1;
                } else if (MODE == MODE_NORMAL || MODE == MODE_NORMAL_S) {
                    if (XS) {                                                // Progress of plane p-1 (curvature.cpp:310-321)
                        double* pc = ex.cout[t.lev] + o;
                        if (two) stg2(pc, st[it].c0.x, st[it].c0.y); else pc[0] = st[it].c0.x;
                    }
                    const double n0 = -fmax(1e-14, sqrt(ax * ax + bx0 * bx0 + g0 * g0));
                    const double n1 = -fmax(1e-14, sqrt(ay * ay + by0 * by0 + g1 * g1));
                    r0[0] = ax / n0; r0[1] = bx0 / n0; r0[2] = g0 / n0;
                    r1[0] = ay / n1; r1[1] = by0 / n1; r1[2] = g1 / n1;
                    if (ex.aux[t.lev]) {
                        double* g = ex.aux[t.lev] + o;
                        const long long cg = ex.cs_aux[t.lev];
                        if (two) { stg2(g, ax, ay); stg2(g + cg, bx0, by0); stg2(g + 2 * cg, g0, g1); }
                        else { g[0] = ax; g[cg] = bx0; g[2 * cg] = g0; }
                    }
                } else {
                    r0[0] = 0.5 * (((0.0 + ax) + bx0) + g0);
                    r1[0] = 0.5 * (((0.0 + ay) + by0) + g1);
                    if (ex.do_threshold) {
                        const long long ai = li.off + (long long)(kz + li.ng) * li.PS + (long long)(jy + li.ng) * li.P + (2 * q + xbase);
                        const double2 pc = *reinterpret_cast<const double2*>(ex.prog[t.lev] + ai);
                        if (pc.x < ex.threshold || pc.x > 1.0 - ex.threshold) r0[0] = 0.0;
                        if (pc.y < ex.threshold || pc.y > 1.0 - ex.threshold) r1[0] = 0.0;
                    }
                }
                double* po = out0 + o;
#pragma unroll
                for (int m = 0; m < NOUT; ++m) {
                    if (two) stg2(po + m * L.cs_out, r0[m], r1[m]); else po[m * L.cs_out] = r0[m];
                }
            }
            st[it].cm = st[it].c0;
            st[it].c0 = cp;
            st[it].a0 = a1;
            st[it].b0 = b1;
        }
        // this warp is done reading stage s
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
}

template <int MODE>
cudaError_t launch_mode(const PaTile* tiles, int ntiles, int stage_doubles, const GridArgs& ga, const StencilExtra& ex,
                        int nvar, cudaStream_t st) {
    constexpr int NIN = ModeTraits<MODE>::NIN;
    size_t smem = (size_t)STAGES * NIN * stage_doubles * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_stencil_tma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    k_stencil_tma<MODE><<<dim3(ntiles, nvar), THREADS, smem, st>>>(tiles, ga, ex, stage_doubles);
    return cudaGetLastError();
}

}  // namespace

int stencil_tma_tile_rows() { return TILE_ROWS; }
int stencil_tma_max_tile_rows() { return MAX_TILE_ROWS; }
int stencil_tma_max_items() { return MAX_ITEMS * CONSUMER_THREADS; }
// largest staged plane ((TY+2) rows x pitch) the pipeline accepts per input component
int stencil_tma_max_plane_doubles() { return (200 * 1024) / (STAGES * 3 * 8); }

cudaError_t launch_stencil_tma(int mode, const PaTile* tiles, int ntiles, int max_plane_doubles, const GridArgs& ga,
                               const StencilExtra& ex, int nvar, cudaStream_t st) {
    if (ntiles <= 0) return cudaSuccess;
    int stage_doubles = (max_plane_doubles + 15) & ~15;
    cudaError_t e;
    switch (mode) {
        case MODE_GRAD: e = launch_mode<MODE_GRAD>(tiles, ntiles, stage_doubles, ga, ex, nvar, st); break;
        case MODE_GRAD3: e = launch_mode<MODE_GRAD3>(tiles, ntiles, stage_doubles, ga, ex, nvar, st); break;
        case MODE_NORMAL: e = launch_mode<MODE_NORMAL>(tiles, ntiles, stage_doubles, ga, ex, nvar, st); break;
        case MODE_DIV: e = launch_mode<MODE_DIV>(tiles, ntiles, stage_doubles, ga, ex, nvar, st); break;
        case MODE_NORMAL_S: e = launch_mode<MODE_NORMAL_S>(tiles, ntiles, stage_doubles, ga, ex, nvar, st); break;
        default: return cudaErrorInvalidValue;
    }
    ++g_launches;
    return e;
}

}  // namespace pa
