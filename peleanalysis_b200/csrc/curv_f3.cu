// curv_f3.cu -- third fused curvature kernel (PA_CURV_FUSED=3): S -> Progress, flame normal and K in one sweep.
//
// Reference data flow (curvature.cpp:310-567): c = (S - pmin) * inv; G = grad c; n = G / -max(1e-14, |G|); FillBoundary(n);
// K = 0.5 * div n.  Its predecessor curv_f2.cu (one 512-thread CTA per SM, two block barriers per plane; removed) was measured at 259
// instructions per cell, of which only 30 % are FP64, and 18 % of the warp time at the barriers.
// This kernel keeps its data flow -- progress planes and flame-normal planes in shared memory, nothing carried in registers
// along z -- and changes what that profile blamed:
//   * work item = K rows x K planes x an x-STRIP of at most 64 cells (+ one rim pair on the strip's inner sides); 256 threads and
//     94 KB of shared memory per CTA, so TWO CTAs share an SM and one computes while the other waits at its barrier;
//   * ONE block barrier per plane: a step stages scalar plane p, computes n of plane p-2 from the planes staged by EARLIER steps
//     and K of plane p-3, whose n_z(p-2) the same thread has just computed (K is mapped like n: a thread's K cell is its n cell);
//   * what a thread loads / computes is fixed for the whole item: source pointers, shared-memory offsets, ownership flags are
//     set up once per item, the plane loop only advances pointers (f2 re-derived them per pair and plane: 54 + 65 of its 259
//     instructions per cell);
//   * the two (row, pair) slots of a thread run through ONE basic block (four sqrt -> reciprocal -> quotient chains in flight).
// Measured on a B200: 5.87 ms (curvature step 7.31 ms against 6.52 ms for the separate kernels); its K-less form (FK = false,
// PA_NORMAL_F3) 3.85 ms against 3.73 ms for MODE_NORMAL_S, and 2.0 ms with its global stores removed (PA_NF3_ABLATE, DESIGN.md
// section 6): the pass is paced by its memory traffic.  Both are opt-in.
// Cells whose K stencil leaves the box are left to k_div_shell, exactly as with the other fused kernels.  Arithmetic is the
// reference's expression order with separate IEEE multiplies and adds (-fmad=false): bit-identical to the separate kernels.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>

#include "kernels.cuh"
#include "stencil_dev.cuh"

namespace pa {

namespace {

constexpr int F3_THREADS = 256;
constexpr int F3_NL = 2;                  // (row, pair) slots of n / K per thread
// Tile geometry.  FK = true: the fused kernel (n with a one-cell rim around the K cells, flame-normal planes in shared memory).
// FK = false: the same sweep without the K part -- S -> Progress and n only, no rim, whole rows of up to 128 cells, 45 KB of
// shared memory; it stands in for MODE_NORMAL_S of the TMA pipeline in front of the separate divergence kernel (PA_NORMAL_F3).
template <bool FK> struct F3Geo {
    static constexpr int KR = FK ? 13 : 8;                // K rows (FK) / rows of n (!FK) per item at most
    static constexpr int NR = FK ? KR + 2 : KR;           // rows of n
    static constexpr int SR = NR + 2;                     // rows of the scalar
    static constexpr int KQ = FK ? 32 : 64;               // pairs (x) per strip at most
    static constexpr int NQ = FK ? KQ + 2 : KQ;           // pairs of n per row at most (FK: one rim pair per inner side)
    static constexpr int SQ = NQ + 2;                     // pairs of the scalar per row at most
    static constexpr int PW = 2 * SQ;                     // row pitch of a progress plane (doubles)
    static constexpr int NW = 2 * NQ;                     // row pitch of a normal-component plane
    static constexpr int CPL = SR * PW;                   // doubles per progress plane
    static constexpr int NPL = FK ? NR * NW : 0;          // doubles per normal-component plane
    static constexpr int SL = (SR * SQ + F3_THREADS - 1) / F3_THREADS;   // scalar pairs a thread stages per plane (3)
    // progress planes in shared memory.  FK: three are read by a step, the fourth is being staged from registers (the plane
    // loaded one step ahead).  !FK: the planes arrive by cp.async up to four steps ahead and are normalised in place: three
    // read + one being normalised + four in flight.  (The register prefetch of one plane was not enough for the shorter
    // steps of the K-less kernel: ncu showed 3.5 cycles of long-scoreboard and 5.7 of barrier stall per issued instruction.)
    static constexpr int RING = FK ? 4 : 8;
    static_assert(NR * NQ <= F3_NL * F3_THREADS, "n slots");
};

// Shared memory is addressed through 32-bit shared-window addresses and ld/st.shared with immediate offsets: with generic
// pointers nvcc re-derives the window base (S2UR SR_CgaCtaId + ULEA) in every basic block and spends a LEA per access.
// SmA + d = d doubles further; the accessors' template argument is a compile-time offset in doubles.
#ifdef PA_HOST_EMULATION
__device__ __forceinline__ double2 ldg2_f3(const double* p) { return *reinterpret_cast<const double2*>(p); }
struct SmA {
    double* p;
    __device__ SmA operator+(int d) const { return SmA{p + d}; }
};
__device__ __forceinline__ SmA sm_base(void* raw) { return SmA{reinterpret_cast<double*>(raw)}; }
template <int D> __device__ __forceinline__ double2 sm_ld2(SmA a) { return *reinterpret_cast<const double2*>(a.p + D); }
template <int D> __device__ __forceinline__ double sm_ld1(SmA a) { return a.p[D]; }
template <int D> __device__ __forceinline__ void sm_st2(SmA a, double x, double y) { *reinterpret_cast<double2*>(a.p + D) = make_double2(x, y); }
#else
__device__ __forceinline__ double2 ldg2_f3(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }
struct SmA {
    uint32_t a;
    __device__ __forceinline__ SmA operator+(int d) const { return SmA{a + 8u * (uint32_t)d}; }
};
__device__ __forceinline__ SmA sm_base(void* raw) { return SmA{smem_u32(raw)}; }
template <int D> __device__ __forceinline__ double2 sm_ld2(SmA a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a.a), "n"(8 * D));
    return v;
}
template <int D> __device__ __forceinline__ double sm_ld1(SmA a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a.a), "n"(8 * D));
    return v;
}
template <int D> __device__ __forceinline__ void sm_st2(SmA a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a.a), "n"(8 * D), "d"(x), "d"(y) : "memory");
}
#endif

// 16-byte asynchronous copy global -> shared (LDGSTS), commit / wait as in PTX
#ifdef PA_HOST_EMULATION
__device__ __forceinline__ void cpa16(SmA dst, const double* src) { cuemu::cp_async_16(dst.p, src); }
__device__ __forceinline__ void cpa_commit() { cuemu::cp_async_commit(); }
__device__ __forceinline__ void cpa_wait(int n) { cuemu::cp_async_wait(n); }
#else
__device__ __forceinline__ void cpa16(SmA dst, const double* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst.a), "l"(src) : "memory"); }
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpa_wait(int n) {
    switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}
#endif

// tiles [0, end[0]) of the launch belong to level 0, [end[0], end[1]) to level 1, ...: the level of a CTA follows from
// blockIdx alone, so the compiler knows it is uniform and keeps the level's pointers and constants in uniform registers
// (a level read from the tile record made every use of them a constant-bank load with a computed index)
struct F3Levels {
    int end[PA_MAX_LEVELS];
};

// Per-item constants live in shared memory and are re-read (one LDS each) where they are used.  Kept in registers they did
// not fit next to the four flame-normal chains: the first version spilled 100 bytes, and because 188 KB of the SM's 228 KB
// L1 / shared memory are carved out as shared memory the spill reloads missed L1 -- 31 % of its stall samples sat on
// instructions waiting for an LDL (profiles/r02_ncu_curv_f3_summary.txt).
enum {
    IC_DXI, IC_DYI, IC_DZI,       // 1 / dx
    IC_PC, IC_PN, IC_PK, IC_PA,   // address of the box's cell (0, 0, 0) in Progress / n_x / K / aux (0: no aux output)
    IC_CSN, IC_CSA,               // component strides of n and aux in BYTES
    IC_GZ0, IC_GZ1,               // source of cell (0, 0, z) on the box's ghost planes z = -1 / z = nz, bit 0: raw scalar
    IC_N
};
template <bool FK> constexpr size_t f3_smem() {
    return (size_t)(F3Geo<FK>::RING * F3Geo<FK>::CPL + 7 * F3Geo<FK>::NPL + IC_N + F3Geo<FK>::SL * F3_THREADS / 2) * sizeof(double);   // + srow[SL][256] ints
}

#ifdef PA_HOST_EMULATION
template <int D> __device__ __forceinline__ long long sm_ldq(SmA a) { return reinterpret_cast<const long long*>(a.p)[D]; }
template <int D> __device__ __forceinline__ void sm_stq(SmA a, long long v) { reinterpret_cast<long long*>(a.p)[D] = v; }
__device__ __forceinline__ int sm_ldi(SmA a, int i) { return reinterpret_cast<const int*>(a.p)[i]; }
__device__ __forceinline__ void sm_sti(SmA a, int i, int v) { reinterpret_cast<int*>(a.p)[i] = v; }
#else
template <int D> __device__ __forceinline__ long long sm_ldq(SmA a) {
    long long v;
    asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(v) : "r"(a.a), "n"(8 * D));
    return v;
}
template <int D> __device__ __forceinline__ void sm_stq(SmA a, long long v) {
    asm volatile("st.shared.b64 [%0+%1], %2;" ::"r"(a.a), "n"(8 * D), "l"(v) : "memory");
}
__device__ __forceinline__ int sm_ldi(SmA a, int i) {
    int v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a.a + 4u * (uint32_t)i));
    return v;
}
__device__ __forceinline__ void sm_sti(SmA a, int i, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a.a + 4u * (uint32_t)i), "r"(v) : "memory"); }
#endif
template <int D> __device__ __forceinline__ char* sm_ldp(SmA a) { return reinterpret_cast<char*>(sm_ldq<D>(a)); }
// 16-byte store to global memory at a byte address (the output addresses come out of shared memory as integers: an ordinary
// store through them would be a generic ST)
#ifdef PA_HOST_EMULATION
__device__ __forceinline__ void stg2b(char* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }
__device__ __forceinline__ void stg1b(char* p, double x) { *reinterpret_cast<double*>(p) = x; }
__device__ __forceinline__ void stg2b_cs(char* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }
#else
__device__ __forceinline__ void stg2b(char* p, double x, double y) { asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory"); }
__device__ __forceinline__ void stg1b(char* p, double x) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(x) : "memory"); }
__device__ __forceinline__ void stg2b_cs(char* p, double x, double y) { asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory"); }
#endif

// ABL != 0: timing experiments, WRONG RESULTS (PA_NF3_ABLATE, K-less form only) -- bit 0: n = G instead of the sqrt -> reciprocal
// -> quotient chain, bit 1: no global stores (the values stay live behind a condition that is never true); bit 2 (results stay
// right): streaming stores (st.global.cs) for Progress and n
template <bool FK, int MINB, int ABL = 0>
__global__ void __launch_bounds__(F3_THREADS, MINB) k_curv_f3(const PaTile* __restrict__ tiles, F3Levels lv, GridArgs ga, StencilExtra ex, int depth) {
    using G = F3Geo<FK>;
    constexpr int F3_PW = G::PW, F3_NW = G::NW, F3_CPL = G::CPL, F3_NPL = G::NPL, F3_SL = G::SL;
    constexpr int RH = FK ? 1 : 0;                               // rim of n around the item's own rows / planes
    constexpr int F3_RING = G::RING;
    PA_DYN_SMEM(smem_raw);
    const SmA C = sm_base(smem_raw);                             // [F3_RING][SR][PW] progress ring
    const SmA NXY = C + F3_RING * F3_CPL;                        // [2]{n_x plane, n_y plane}
    const SmA NZs = NXY + 4 * F3_NPL;                            // [3] n_z ring
    const SmA IC = NZs + 3 * F3_NPL;                             // item constants
    const SmA SROW = IC + IC_N;                                  // [SL][F3_THREADS] ints: y * P + x of the pairs a thread stages
    const int tid = threadIdx.x;

    int nplanes, g0, g1, mw0, mw1, PSin, PSout;                  // the plane loop's scalars
    const double* sp[F3_SL];                                     // the pair on the NEXT plane to load, if that plane is inside the box
    int ssm[F3_SL];                                              // offset inside a staged plane, -1: slot not used
    unsigned sraw = 0;                                           // bit i: slot i is raw on planes inside the box
    int so[F3_NL], no[F3_NL];                                    // offsets of the pair inside a progress plane / a normal plane
    int oe[F3_NL];                                               // output element of the pair relative to the box's cell (0, 0, 0), plane of the next n
    unsigned fl = 0;                                             // per slot j: bit j = slot used, 4+j = writes c / n, 8+j = writes K, 12+j / 16+j = pair holds x = 0 / x = nx-1
    {
        const PaTile t = tiles[blockIdx.x];
        int lev = 0;
#pragma unroll
        for (int l = 0; l < PA_MAX_LEVELS - 1; ++l) lev += (int)blockIdx.x >= lv.end[l] ? 1 : 0;
        const int xq0 = (t.lev >> 8) & 0xff, KQ = (t.lev >> 16) & 0xff;   // first K pair and K pairs of the strip
        const LevArgs& L = ga.L[lev];
        const PaBoxDev bx = L.boxes[t.box];
        const PaLayDev li = L.lay_in[t.box], lo = L.lay_out[t.box];
        const PaNbr nb = L.nbr[t.box];
        const int nx = bx.n[0], nyb = bx.n[1], nzb = bx.n[2];
        const int nxp = nx >> 1;
        const int nq0 = (FK && xq0 > 0) ? xq0 - 1 : xq0, nq1 = (FK && xq0 + KQ + 1 <= nxp) ? xq0 + KQ + 1 : xq0 + KQ;   // pairs of n this item computes
        const int NQ = nq1 - nq0, SQ = NQ + 2;
        const int sq0 = nq0 - 1;                                 // first staged pair; pair -1 = cells (-2, -1), pair nxp = (nx, nx + 1)
        const int KR = t.ny, NR = KR + 2 * RH, SR = NR + 2;
        const int yS0 = t.y0 - RH - 1, zS0 = t.z0 - RH - 1;      // first scalar row / plane (box-relative, >= -1)
        nplanes = t.nz + 2 * RH + 2;                             // scalar planes
        g0 = zS0 < 0 ? 0 : -1;                                   // plane indices of the box's z ghost planes inside the item, -1: none
        g1 = zS0 + nplanes - 1 >= nzb ? nplanes - 1 : -1;
        PSin = li.PS; PSout = lo.PS;

        // ---- sources of the scalar: own slab, or a linked neighbour's slab read in place ----
        const int c0 = L.in_comp;
        const double* const own = L.in + li.off + (long long)li.ng * li.PS + (long long)li.ng * li.P + (li.ng + li.xoff);   // cell (0, 0, 0)
        const double* lnk[6];
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            const PaNbrFace F = nb.f[f];
            lnk[f] = nullptr;
            if (F.nb >= 0) {
                const PaPeerSlab ps = L.peers[F.rank];
                const PaLayDev ln = L.lay_in[F.nb];
                lnk[f] = ps.base + (long long)c0 * ps.cs + ln.off + (long long)(F.rel[2] + ln.ng) * ln.PS + (long long)(F.rel[1] + ln.ng) * ln.P +
                         (F.rel[0] + ln.ng + ln.xoff);                                                     // my cell (0, 0, 0) in the neighbour's slab
            }
        }

        // ---- what this thread stages: F3_SL pairs (row, pair) of every plane.  A pair holds the RAW scalar (valid cells of this
        //      box or of a linked neighbour) or progress values already (this box's materialised ghost cells, written in progress
        //      space by the ghost fill).  Edge cells (outside the box in two directions) are staged from any readable address: no
        //      stencil of this kernel touches them. ----
#pragma unroll
        for (int i = 0; i < F3_SL; ++i) {
            const int p = tid + i * F3_THREADS;
            ssm[i] = -1; sp[i] = own;
            int srow = 0;
            if (p < SR * SQ) {
                const int rs = p / SQ, qs = p - rs * SQ;
                const int y = yS0 + rs, xg = 2 * (sq0 + qs);
                const bool yin = (y >= 0) & (y < nyb);
                const double* base = own;
                bool raw = true;
                if (xg < 0) { if (yin && lnk[0]) base = lnk[0]; else raw = false; }
                else if (xg >= nx) { if (yin && lnk[3]) base = lnk[3]; else raw = false; }
                else if (y < 0) { if (lnk[1]) base = lnk[1]; else raw = false; }
                else if (y >= nyb) { if (lnk[4]) base = lnk[4]; else raw = false; }
                srow = y * li.P + xg;
                sp[i] = base + ((long long)zS0 * li.PS + srow);
                ssm[i] = rs * F3_PW + 2 * qs;
                if (raw) sraw |= 1u << i;
            }
            sm_sti(SROW, i * F3_THREADS + tid, srow);
        }

        // ---- what this thread computes: F3_NL (row, pair) slots of n, and K of the same cells one plane behind.  Progress and n
        //      are written by the item that holds the cell as a K row / plane / strip column; the box's outermost rows / planes
        //      go with the first / last item ----
        const int wy0 = (FK && t.y0 == 1) ? 0 : t.y0, wy1 = (FK && t.y0 + t.ny == nyb - 1) ? nyb - 1 : t.y0 + t.ny - 1;
        const int wz0 = (FK && t.z0 == 1) ? 0 : t.z0, wz1 = (FK && t.z0 + t.nz == nzb - 1) ? nzb - 1 : t.z0 + t.nz - 1;
        mw0 = wz0 - zS0; mw1 = wz1 - zS0;                        // planes (item-relative) whose c / n this item writes
#pragma unroll
        for (int j = 0; j < F3_NL; ++j) {
            int p = tid + j * F3_THREADS;
            const bool act = p < NR * NQ;
            if (!act) p = 0;
            const int rn = p / NQ, qn = p - rn * NQ;
            const int y = t.y0 - RH + rn, q = nq0 + qn;
            so[j] = (rn + 1) * F3_PW + 2 * (qn + 1);
            no[j] = rn * F3_NW + 2 * qn;
            oe[j] = (zS0 + 1) * lo.PS + y * lo.P + 2 * q;
            const bool incol = (q >= xq0) & (q < xq0 + KQ);
            if (act) {
                fl |= 1u << j;
                if (incol & (y >= wy0) & (y <= wy1)) fl |= 16u << j;
                if (FK && (incol & (rn >= 1) & (rn <= KR))) fl |= 256u << j;
                if (q == 0) fl |= 4096u << j;
                if (q == nxp - 1) fl |= 65536u << j;
            }
        }
        if (tid == 0) {
            const long long ob = lo.off + (long long)lo.ng * lo.PS + (long long)lo.ng * lo.P + (lo.ng + lo.xoff);   // output cell (0, 0, 0)
            sm_stq<IC_DXI>(IC, __double_as_longlong(L.dxi[0]));
            sm_stq<IC_DYI>(IC, __double_as_longlong(L.dxi[1]));
            sm_stq<IC_DZI>(IC, __double_as_longlong(L.dxi[2]));
            sm_stq<IC_PC>(IC, (long long)(ex.cout[lev] + ob));
            sm_stq<IC_PN>(IC, (long long)(L.out + ob));
            if (FK) sm_stq<IC_PK>(IC, (long long)(ex.kout[lev] + ob));
            sm_stq<IC_PA>(IC, ex.aux[lev] ? (long long)(ex.aux[lev] + ob) : 0ll);
            sm_stq<IC_CSN>(IC, L.cs_out * 8);
            sm_stq<IC_CSA>(IC, ex.cs_aux[lev] * 8);
            const double* z0s = (lnk[2] ? lnk[2] : own) - li.PS;
            const double* z1s = (lnk[5] ? lnk[5] : own) + (long long)nzb * li.PS;
            sm_stq<IC_GZ0>(IC, (long long)z0s | (lnk[2] ? 1 : 0));
            sm_stq<IC_GZ1>(IC, (long long)z1s | (lnk[5] ? 1 : 0));
        }
    }
    __syncthreads();

    const double pmin = ex.pmin, pinv = ex.inv;
    double2 pre[F3_SL];
    unsigned praw = 0;
    auto issue_loads = [&](int ps) {                             // FK: plane ps of the item into registers
        if ((ps != g0) & (ps != g1)) {
#pragma unroll
            for (int i = 0; i < F3_SL; ++i)
                if (ssm[i] >= 0) pre[i] = ldg2_f3(sp[i]);
            praw = sraw;
        } else {                                                 // the box's z ghost plane: the z neighbour's valid plane, or own ghost cells
            const long long gz = ps == g0 ? sm_ldq<IC_GZ0>(IC) : sm_ldq<IC_GZ1>(IC);
            praw = (gz & 1) ? 7u : 0u;
            const double* base = reinterpret_cast<const double*>(gz & ~1ll);
#pragma unroll
            for (int i = 0; i < F3_SL; ++i)
                if (ssm[i] >= 0) pre[i] = ldg2_f3(base + sm_ldi(SROW, i * F3_THREADS + tid));
        }
#pragma unroll
        for (int i = 0; i < F3_SL; ++i) sp[i] += PSin;
    };
    auto store_loads = [&](int ps) {                             // normalise once (curvature.cpp:316-320) and stage
        const SmA Cs = C + (ps & (F3_RING - 1)) * F3_CPL;
#pragma unroll
        for (int i = 0; i < F3_SL; ++i)
            if (ssm[i] >= 0) {
                double2 v = pre[i];
                if (praw & (1u << i)) { v.x = (v.x - pmin) * pinv; v.y = (v.y - pmin) * pinv; }
                sm_st2<0>(Cs + ssm[i], v.x, v.y);
            }
    };
    auto issue_async = [&](int ps) {                             // !FK: plane ps of the item straight into its ring slot; one group per plane
        if (ps < nplanes) {
            const SmA Cs = C + (ps & (F3_RING - 1)) * F3_CPL;
            if ((ps != g0) & (ps != g1)) {
#pragma unroll
                for (int i = 0; i < F3_SL; ++i)
                    if (ssm[i] >= 0) cpa16(Cs + ssm[i], sp[i]);
            } else {
                const long long gz = ps == g0 ? sm_ldq<IC_GZ0>(IC) : sm_ldq<IC_GZ1>(IC);
                const double* base = reinterpret_cast<const double*>(gz & ~1ll);
#pragma unroll
                for (int i = 0; i < F3_SL; ++i)
                    if (ssm[i] >= 0) cpa16(Cs + ssm[i], base + sm_ldi(SROW, i * F3_THREADS + tid));
            }
#pragma unroll
            for (int i = 0; i < F3_SL; ++i) sp[i] += PSin;
        }
        cpa_commit();
    };
    auto normalise_own = [&](int ps) {                           // !FK: the pairs this thread copied, in place (curvature.cpp:316-320)
        unsigned raw = sraw;
        if ((ps == g0) | (ps == g1)) raw = ((ps == g0 ? sm_ldq<IC_GZ0>(IC) : sm_ldq<IC_GZ1>(IC)) & 1) ? 7u : 0u;
        const SmA Cs = C + (ps & (F3_RING - 1)) * F3_CPL;
#pragma unroll
        for (int i = 0; i < F3_SL; ++i)
            if (ssm[i] >= 0 && (raw & (1u << i))) {
                const double2 v = sm_ld2<0>(Cs + ssm[i]);
                sm_st2<0>(Cs + ssm[i], (v.x - pmin) * pinv, (v.y - pmin) * pinv);
            }
    };

    const bool do_thr = ex.do_threshold != 0;
    const double thr_lo = ex.threshold, thr_hi = 1.0 - ex.threshold;
    int r3 = 1;                                                  // (plane of n) % 3 without the division
    if (FK) issue_loads(0);
    else for (int d = 0; d < depth; ++d) issue_async(d);
    for (int ps = 0; ps <= nplanes; ++ps) {
        if (FK) {
            if (ps < nplanes) {
                store_loads(ps);
                if (ps + 1 < nplanes) issue_loads(ps + 1);       // in flight across the arithmetic below
            }
        } else {
            cpa_wait(depth - 1);                                 // this thread's copies of plane ps have landed
            if (ps < nplanes) normalise_own(ps);
            issue_async(ps + depth);
        }
        if (ps >= 3) {
            // ---- flame normal of plane m = ps - 2 from the progress planes m-1, m, m+1 staged by earlier steps ----
            const int m = ps - 2;
            const SmA C0 = C + ((ps - 3) & (F3_RING - 1)) * F3_CPL;   // plane m - 1
            const SmA C1 = C + ((ps - 2) & (F3_RING - 1)) * F3_CPL;
            const SmA C2 = C + ((ps - 1) & (F3_RING - 1)) * F3_CPL;
            const SmA nw = NXY + (m & 1) * (2 * F3_NPL);
            const SmA nr = NXY + ((m & 1) ^ 1) * (2 * F3_NPL);
            const int r3m1 = r3 == 0 ? 2 : r3 - 1, r3m2 = r3m1 == 0 ? 2 : r3m1 - 1;
            const SmA zw = NZs + r3 * F3_NPL;
            const SmA z1 = NZs + r3m1 * F3_NPL;
            const SmA z0 = NZs + r3m2 * F3_NPL;
            const double dxi = sm_ld1<IC_DXI>(IC), dyi = sm_ld1<IC_DYI>(IC), dzi = sm_ld1<IC_DZI>(IC);
            double gx[4], gy[4], gz[4], n0[4], n1[4], n2[4];
            double2 cc[F3_NL];
#pragma unroll
            for (int j = 0; j < F3_NL; ++j) {
                const SmA c1 = C1 + so[j];
                const double2 c = sm_ld2<0>(c1);
                const double xm = sm_ld1<-1>(c1), xp = sm_ld1<2>(c1);
                const double2 ym = sm_ld2<-F3_PW>(c1), yp = sm_ld2<F3_PW>(c1);
                const double2 zm = sm_ld2<0>(C0 + so[j]), zp = sm_ld2<0>(C2 + so[j]);
                gx[2 * j] = cdiff(dxi, xm, c.x, c.y); gx[2 * j + 1] = cdiff(dxi, c.x, c.y, xp);
                gy[2 * j] = cdiff(dyi, ym.x, c.x, yp.x); gy[2 * j + 1] = cdiff(dyi, ym.y, c.y, yp.y);
                gz[2 * j] = cdiff(dzi, zm.x, c.x, zp.x); gz[2 * j + 1] = cdiff(dzi, zm.y, c.y, zp.y);
                cc[j] = c;
            }
            if (ABL & 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { n0[i] = gx[i]; n1[i] = gy[i]; n2[i] = gz[i]; }
            } else {
                normal_quad(gx, gy, gz, n0, n1, n2);             // curvature.cpp:467-502
            }
            const bool wplane = (ABL & 2) ? depth == 12345 : (m >= mw0) & (m <= mw1);
#pragma unroll
            for (int j = 0; j < F3_NL; ++j) {
                if (FK && (fl & (1u << j))) {
                    sm_st2<0>(nw + no[j], n0[2 * j], n0[2 * j + 1]);
                    sm_st2<F3_NPL>(nw + no[j], n1[2 * j], n1[2 * j + 1]);
                    sm_st2<0>(zw + no[j], n2[2 * j], n2[2 * j + 1]);
                }
                if (wplane && (fl & (16u << j))) {
                    const long long ob = 8ll * oe[j];
                    char* const pn = sm_ldp<IC_PN>(IC) + ob;
                    const long long csn = sm_ldq<IC_CSN>(IC);
                    if (ABL & 4) {
                        stg2b_cs(sm_ldp<IC_PC>(IC) + ob, cc[j].x, cc[j].y);
                        stg2b_cs(pn, n0[2 * j], n0[2 * j + 1]);
                        stg2b_cs(pn + csn, n1[2 * j], n1[2 * j + 1]);
                        stg2b_cs(pn + 2 * csn, n2[2 * j], n2[2 * j + 1]);
                    } else {
                        stg2b(sm_ldp<IC_PC>(IC) + ob, cc[j].x, cc[j].y);                 // Progress (curvature.cpp:310-321)
                        stg2b(pn, n0[2 * j], n0[2 * j + 1]);
                        stg2b(pn + csn, n1[2 * j], n1[2 * j + 1]);
                        stg2b(pn + 2 * csn, n2[2 * j], n2[2 * j + 1]);
                    }
                    char* const pa = sm_ldp<IC_PA>(IC);
                    if (pa) {
                        const long long csa = sm_ldq<IC_CSA>(IC);
                        stg2b(pa + ob, gx[2 * j], gx[2 * j + 1]);
                        stg2b(pa + ob + csa, gy[2 * j], gy[2 * j + 1]);
                        stg2b(pa + ob + 2 * csa, gz[2 * j], gz[2 * j + 1]);
                    }
                }
            }
            if (FK && m >= 3) {
                // ---- K of plane m - 1: n_x / n_y of that plane (staged one step ago), n_z of m-2, m-1 (shared memory) and m (registers).
                //      Every thread evaluates its cells; only K rows / columns are stored ----
#pragma unroll
                for (int j = 0; j < F3_NL; ++j) {
                    const SmA na = nr + no[j];
                    const double2 a = sm_ld2<0>(na);
                    const double am = sm_ld1<-1>(na), ap = sm_ld1<2>(na);   // outside the staged row for a box's first / last pair: those cells are not stored
                    const double2 bm = sm_ld2<F3_NPL - F3_NW>(na), b = sm_ld2<F3_NPL>(na), bp = sm_ld2<F3_NPL + F3_NW>(na);
                    const double2 gm = sm_ld2<0>(z0 + no[j]), g = sm_ld2<0>(z1 + no[j]);
                    const double dx0 = cdiff(dxi, am, a.x, a.y), dx1 = cdiff(dxi, a.x, a.y, ap);
                    const double dy0 = cdiff(dyi, bm.x, b.x, bp.x), dy1 = cdiff(dyi, bm.y, b.y, bp.y);
                    const double dz0 = cdiff(dzi, gm.x, g.x, n2[2 * j]), dz1 = cdiff(dzi, gm.y, g.y, n2[2 * j + 1]);
                    double k0 = 0.5 * (((0.0 + dx0) + dy0) + dz0);                       // curvature.cpp:505-547
                    double k1 = 0.5 * (((0.0 + dx1) + dy1) + dz1);
                    if (do_thr) {                                                          // :549-567 (K only; n is clipped afterwards)
                        const double2 pc = sm_ld2<0>(C0 + so[j]);
                        if (pc.x < thr_lo || pc.x > thr_hi) k0 = 0.0;
                        if (pc.y < thr_lo || pc.y > thr_hi) k1 = 0.0;
                    }
                    if (fl & (256u << j)) {
                        char* pk = sm_ldp<IC_PK>(IC) + 8ll * (oe[j] - PSout);
                        const bool first = (fl & (4096u << j)) != 0, last = (fl & (65536u << j)) != 0;   // x = 0 and x = nx - 1 belong to k_div_shell
                        if (!first && !last) stg2b(pk, k0, k1);
                        else { if (!first) stg1b(pk, k0); if (!last) stg1b(pk + 8, k1); }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < F3_NL; ++j) oe[j] += PSout;
            r3 = r3 == 2 ? 0 : r3 + 1;
        }
        __syncthreads();                                         // plane ps and n of plane ps - 2 staged for the next step
    }
}

}  // namespace

int curv_f3_rows() { return F3Geo<true>::KR; }
int curv_f3_strip_pairs() { return F3Geo<true>::KQ; }
int normal_f3_rows() { return F3Geo<false>::KR; }
int normal_f3_strip_pairs() { return F3Geo<false>::KQ; }

template <bool FK, int MINB, int ABL = 0>
static cudaError_t launch_f3(const PaTile* tiles, int ntiles, const int* level_end, int nlev, const GridArgs& ga, const StencilExtra& ex,
                             int depth, cudaStream_t st) {
    if (ntiles <= 0) return cudaSuccess;
    F3Levels lv;
    for (int l = 0; l < PA_MAX_LEVELS; ++l) lv.end[l] = l < nlev ? level_end[l] : 0x7fffffff;
    static std::map<int, bool> configured;
    static std::mutex mu;
    int dev = 0;
    { cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e; }
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(k_curv_f3<FK, MINB, ABL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f3_smem<FK>());
            if (e != cudaSuccess) return e;
#ifndef PA_HOST_EMULATION
            e = cudaFuncSetAttribute(k_curv_f3<FK, MINB, ABL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);   // FK: two CTAs per SM need 195 KB
            if (e != cudaSuccess) return e;
#endif
            configured[dev] = true;
        }
    }
    PA_LAUNCH(ntiles, F3_THREADS, f3_smem<FK>(), st, k_curv_f3<FK, MINB, ABL>)(tiles, lv, ga, ex, depth);
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_curv_f3(const PaTile* tiles, int ntiles, const int* level_end, int nlev, const GridArgs& ga, const StencilExtra& ex,
                           cudaStream_t st) {
    return launch_f3<true, 2>(tiles, ntiles, level_end, nlev, ga, ex, 1, st);
}
cudaError_t launch_normal_f3(const PaTile* tiles, int ntiles, const int* level_end, int nlev, const GridArgs& ga, const StencilExtra& ex,
                             cudaStream_t st) {
    const char* e = getenv("PA_NF3_CTAS");                      // 3: 80-register build, three CTAs per SM (spills 160 bytes)
    const char* ed = getenv("PA_NF3_DEPTH");                    // planes in flight per CTA (1 .. 4)
    const int depth = ed ? std::min(4, std::max(1, atoi(ed))) : 3;
    const char* ea = getenv("PA_NF3_ABLATE");                   // timing experiments (wrong results): 1 no chain, 2 no stores, 3 both
    if (ea && ea[0] == '1') return launch_f3<false, 2, 1>(tiles, ntiles, level_end, nlev, ga, ex, depth, st);
    if (ea && ea[0] == '2') return launch_f3<false, 2, 2>(tiles, ntiles, level_end, nlev, ga, ex, depth, st);
    if (ea && ea[0] == '3') return launch_f3<false, 2, 3>(tiles, ntiles, level_end, nlev, ga, ex, depth, st);
    if (ea && ea[0] == '4') return launch_f3<false, 2, 4>(tiles, ntiles, level_end, nlev, ga, ex, depth, st);
    if (ea && ea[0] == '5') return launch_f3<false, 2, 5>(tiles, ntiles, level_end, nlev, ga, ex, depth, st);
    if (e && e[0] == '3') return launch_f3<false, 3>(tiles, ntiles, level_end, nlev, ga, ex, depth, st);
    return launch_f3<false, 2>(tiles, ntiles, level_end, nlev, ga, ex, depth, st);
}

}  // namespace pa
