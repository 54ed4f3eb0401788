// tests/emu/cuemu.cpp -- TEST INFRASTRUCTURE: the CUDA execution-model emulator behind tests/emu/cuda_runtime.h.
//
// One OS thread.  A kernel launch runs its blocks one after the other; inside a block every CUDA thread is a fiber
// (ucontext) and the scheduler runs the live fibers round-robin (or, with CUEMU_SEED != 0, in a fresh pseudo-random order
// every round).  A fiber runs until it blocks in a synchronisation primitive:
//   __syncthreads            all live threads of the block
//   __syncwarp / __shfl_*    the lanes named by the mask (double-buffered value slots)
//   mbarrier                 pending-arrival count + transaction bytes + phase bit, try_wait on the phase parity
//   cp.async.bulk / cp.async queued; they complete (copy the bytes, then complete_tx / arrive) a pseudo-random number of
//                            scheduler rounds later (CUEMU_SEED == 0: within the same round), so a consumer that reads a
//                            stage before its barrier flipped, or a producer that refills a stage too early, shows up
// A scheduler round in which nothing happened (no arrival, no completion, no fiber finished, nothing pending) is a
// deadlock: the emulator prints what every live thread is waiting for and aborts.
#include <ucontext.h>

#include <cstdio>
#include <map>
#include <mutex>
#include <sys/mman.h>
#include <vector>

#include "cuda_runtime.h"

namespace cuemu {

uint3 g_threadIdx = {0, 0, 0}, g_blockIdx = {0, 0, 0};
dim3 g_blockDim, g_gridDim;

namespace {

constexpr size_t STACK_BYTES = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = true;
    const char* waiting = "";       // what the fiber is blocked on (deadlock report)
    const void* wait_obj = nullptr;
    unsigned long long async_due = 0;   // latest completion tick of this thread's outstanding cp.async copies
    // bulk-store groups (cp.async.bulk shared -> global): ops still reading shared memory per committed group, oldest first;
    // the last entry is the open (uncommitted) group
    std::vector<int> store_groups = std::vector<int>(1, 0);
    int groups_retired = 0;             // committed groups already dropped from the front of store_groups
    // cp.async (non-bulk) groups: completion tick of every committed group that may still be in flight, oldest first, and of
    // the open group
    std::vector<unsigned long long> cpa_groups;
    unsigned long long cpa_open = 0;
};

struct WarpState {
    unsigned long long vals[2][32];
    int gen = 0, arrived = 0;
};

struct MBar {
    uint32_t count = 0;
    long long pending = 0, tx = 0;
    uint32_t phase = 0;
};

struct Async {
    int kind;                       // 0: bulk copy + complete_tx, 1: plain copy, 2: arrive, 3: bulk store (shared -> global)
    void* dst; const void* src; size_t bytes; uint64_t* bar;
    unsigned long long due;
    int fiber = -1, group = -1;     // kind 3: issuing thread and its (absolute) bulk-store group
};

std::vector<Fiber> fibers;          // grows to the largest block seen; stacks are reused
ucontext_t sched_ctx;
Fiber* cur = nullptr;
int n_threads = 0, live = 0;
unsigned long long progress = 0, tick = 0;
int sync_gen = 0, sync_arrived = 0;
struct NamedBar { int gen = 0, arrived = 0; };
std::map<int, NamedBar> named_bars;
std::vector<WarpState> warps;
std::map<const void*, MBar> mbars;
std::vector<Async> asyncq;
const std::function<void()>* body = nullptr;
std::vector<unsigned char> dyn;
unsigned char* dyn_aligned = nullptr;
unsigned long long rng_state = 0;
long long seed = -1;
int starve = 0;

unsigned long long rnd() {          // splitmix64
    unsigned long long z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void trampoline() {
    (*body)();
    cur->done = true;
    --live;
    ++progress;
    swapcontext(&cur->ctx, &sched_ctx);
}

void yield(const char* why, const void* obj) {
    cur->waiting = why; cur->wait_obj = obj;
    swapcontext(&cur->ctx, &sched_ctx);
    cur->waiting = "";
}

void mbar_check(MBar& b) {
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; }
    ++progress;
}
MBar& mbar_of(uint64_t* bar) {
    auto it = mbars.find(bar);
    if (it == mbars.end()) { std::fprintf(stderr, "[cuemu] mbarrier %p used before mbarrier.init\n", (void*)bar); std::abort(); }
    return it->second;
}

void run_async(bool all) {
    // stable: entries that are due complete in issue order
    size_t w = 0;
    for (size_t i = 0; i < asyncq.size(); ++i) {
        Async& a = asyncq[i];
        if (all || a.due <= tick) {
            if (a.kind != 2) std::memcpy(a.dst, a.src, a.bytes);       // kind 3 reads shared memory only NOW: reuse before the wait shows
            if (a.kind == 3) {
                Fiber& f = fibers[a.fiber];
                const int rel = a.group - f.groups_retired;
                if (rel >= 0 && rel < (int)f.store_groups.size()) --f.store_groups[rel];
            }
            if (a.kind == 0) { MBar& b = mbar_of(a.bar); b.tx -= (long long)a.bytes; mbar_check(b); }
            if (a.kind == 2) { MBar& b = mbar_of(a.bar); --b.pending; mbar_check(b); }
            ++progress;
        } else {
            asyncq[w++] = a;
        }
    }
    asyncq.resize(w);
}

unsigned long long delay() { return seed == 0 ? 0ULL : rnd() % 4ULL; }

[[noreturn]] void deadlock() {
    std::fprintf(stderr, "[cuemu] DEADLOCK in block (%u,%u,%u) of grid (%u,%u,%u): %d live threads, none can proceed\n", g_blockIdx.x,
                 g_blockIdx.y, g_blockIdx.z, g_gridDim.x, g_gridDim.y, g_gridDim.z, live);
    int shown = 0;
    for (int t = 0; t < n_threads && shown < 48; ++t)
        if (!fibers[t].done) { std::fprintf(stderr, "   thread %4d (warp %d lane %d): %s %p\n", t, t / 32, t % 32, fibers[t].waiting, fibers[t].wait_obj); ++shown; }
    for (auto& kv : mbars)
        std::fprintf(stderr, "   mbarrier %p: count %u pending %lld tx %lld phase %u\n", kv.first, kv.second.count, kv.second.pending, kv.second.tx, kv.second.phase);
    std::abort();
}

void run_block() {
    std::vector<int> order(n_threads);
    for (int t = 0; t < n_threads; ++t) order[t] = t;
    while (live > 0) {
        const unsigned long long before = progress;
        if (seed > 0)
            for (int t = n_threads - 1; t > 0; --t) std::swap(order[t], order[rnd() % (unsigned long long)(t + 1)]);
        // CUEMU_STARVE=p (1..9): every scheduler round each WARP is left out with probability p/10 -- warps then drift apart by
        // many instructions, as they do on an SM whose schedulers favour other warps.  Protocols that only hold while all warps
        // advance at the same pace (a phase bit that aliases when one side runs two steps ahead) fail here, not on the GPU.
        std::vector<char> skip_warp;
        bool skipped_any = false;
        if (starve > 0) {
            skip_warp.assign((n_threads + 31) / 32, 0);
            for (auto& x : skip_warp) { x = (char)((int)(rnd() % 10ULL) < starve); }
        }
        for (int k = 0; k < n_threads; ++k) {
            Fiber& f = fibers[order[k]];
            if (f.done) continue;
            const int t = order[k];
            if (starve > 0 && skip_warp[t / 32]) { skipped_any = true; continue; }
            g_threadIdx.x = (unsigned)t % g_blockDim.x;
            g_threadIdx.y = ((unsigned)t / g_blockDim.x) % g_blockDim.y;
            g_threadIdx.z = (unsigned)t / (g_blockDim.x * g_blockDim.y);
            cur = &f;
            swapcontext(&sched_ctx, &f.ctx);
        }
        ++tick;
        run_async(false);
        if (progress == before && skipped_any) continue;
        if (progress == before) {
            if (!asyncq.empty()) { tick = asyncq.front().due; for (auto& a : asyncq) tick = std::max(tick, a.due); run_async(true); }
            else deadlock();
        }
    }
    run_async(true);
}

}  // namespace

unsigned char* dyn_smem() { return dyn_aligned; }

// ---- streams: created streams may defer their work to the latest point the events allow ---------------------------
// The emulator itself is single-threaded (one set of fibers); host threads that drive different "GPUs" -- the one-process,
// one-thread-per-GPU hosts -- serialise on this lock, one whole kernel launch at a time.
static std::recursive_mutex g_emu_mutex;
namespace {
std::map<cudaStream_t, std::vector<std::function<void()>>> created_streams;
std::map<cudaEvent_t, cudaStream_t> event_stream;
bool defer_side() { const char* e = getenv("CUEMU_DEFER_SIDE"); return e && e[0] == '1'; }
void flush_stream(cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> lock(g_emu_mutex);
    auto it = created_streams.find(st);
    if (it == created_streams.end()) return;
    std::vector<std::function<void()>> q;
    q.swap(it->second);
    for (auto& w : q) w();
}
}  // namespace
void submit(cudaStream_t st, std::function<void()> work) {
    std::lock_guard<std::recursive_mutex> lock(g_emu_mutex);
    auto it = created_streams.find(st);
    if (it != created_streams.end() && defer_side()) it->second.push_back(std::move(work));
    else work();
}

void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn) {
    {   // CUEMU_SEED is read at every launch (tests flip it); the random stream restarts whenever the seed changes
        const char* e = getenv("CUEMU_SEED");
        const long long sd = e ? atoll(e) : 0;
        if (sd != seed) { seed = sd; rng_state = (unsigned long long)seed * 0x2545F4914F6CDD1DULL + 1; }
        const char* es = getenv("CUEMU_STARVE");
        starve = es ? std::min(9, std::max(0, atoi(es))) : 0;
    }
    if (cur != nullptr) { std::fprintf(stderr, "[cuemu] nested kernel launch\n"); std::abort(); }
    n_threads = (int)(block.x * block.y * block.z);
    if (n_threads <= 0 || n_threads > 1024) { std::fprintf(stderr, "[cuemu] bad block size %d\n", n_threads); std::abort(); }
    while ((int)fibers.size() < n_threads) {
        fibers.emplace_back();
        void* s = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (s == MAP_FAILED) { std::perror("[cuemu] mmap"); std::abort(); }
        fibers.back().stack = (char*)s;
    }
    dyn.assign(smem + 256, 0xA5);                      // shared memory starts as garbage, like the real thing
    dyn_aligned = (unsigned char*)(((uintptr_t)dyn.data() + 127) & ~(uintptr_t)127);
    g_blockDim = block; g_gridDim = grid;
    body = &fn;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = uint3{bx, by, bz};
                live = n_threads;
                sync_gen = 0; sync_arrived = 0;
                warps.assign((n_threads + 31) / 32, WarpState());
                mbars.clear();
                named_bars.clear();
                asyncq.clear();
                for (int t = 0; t < n_threads; ++t) {
                    Fiber& f = fibers[t];
                    f.done = false; f.waiting = ""; f.async_due = 0;
                    f.store_groups.assign(1, 0); f.groups_retired = 0; f.cpa_groups.clear(); f.cpa_open = 0;
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = STACK_BYTES;
                    f.ctx.uc_link = nullptr;
                    makecontext(&f.ctx, trampoline, 0);
                }
                run_block();
            }
    cur = nullptr;
    body = nullptr;
}

void syncthreads() {
    const int gen = sync_gen;
    ++sync_arrived; ++progress;
    while (sync_gen == gen) {
        if (sync_arrived >= live) { ++sync_gen; sync_arrived = 0; ++progress; break; }
        yield("__syncthreads", nullptr);
    }
}

// bar.sync id, count: the first `count` arrivals of a generation release together (threads name the same count)
void named_bar(int id, int count) {
    NamedBar& B = named_bars[id];
    const int gen = B.gen;
    ++B.arrived; ++progress;
    if (B.arrived >= count) { B.arrived = 0; ++B.gen; ++progress; return; }
    while (B.gen == gen) yield("bar.sync (named barrier)", &B);
}

static int popc(unsigned m) { return __builtin_popcount(m); }

const unsigned long long* warp_exchange(unsigned mask, const void* v, size_t bytes) {
    const int t = (int)(g_threadIdx.x + g_blockDim.x * (g_threadIdx.y + g_blockDim.y * g_threadIdx.z));
    WarpState& W = warps[t / 32];
    const int lane = t % 32;
    if (!((mask >> lane) & 1u)) { std::fprintf(stderr, "[cuemu] lane %d calls a warp collective it is not named in (mask %08x)\n", lane, mask); std::abort(); }
    const int gen = W.gen;
    unsigned long long x = 0;
    if (v) std::memcpy(&x, v, bytes);
    W.vals[gen & 1][lane] = x;
    ++W.arrived; ++progress;
    if (W.arrived >= popc(mask)) { W.arrived = 0; ++W.gen; }
    while (W.gen == gen) yield("warp collective", &W);
    return W.vals[gen & 1];
}
void syncwarp(unsigned mask) { (void)warp_exchange(mask, nullptr, 0); }

void mbar_init(uint64_t* bar, uint32_t count) { MBar b; b.count = count; b.pending = count; mbars[bar] = b; ++progress; }
void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    MBar& b = mbar_of(bar);
    b.tx += bytes;
    if (b.tx > (1 << 20) - 1) { std::fprintf(stderr, "[cuemu] mbarrier tx-count %lld exceeds the hardware range (2^20 - 1)\n", b.tx); std::abort(); }
    --b.pending;
    mbar_check(b);
}
void mbar_arrive(uint64_t* bar) { MBar& b = mbar_of(bar); --b.pending; mbar_check(b); }
void mbar_wait(uint64_t* bar, uint32_t parity) {
    MBar& b = mbar_of(bar);
    while (b.phase == (parity & 1u)) yield("mbarrier wait", bar);
}
void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    if ((bytes & 15u) || ((uintptr_t)dst & 15u) || ((uintptr_t)src & 15u)) {
        std::fprintf(stderr, "[cuemu] cp.async.bulk needs 16-byte aligned addresses and size: dst %p src %p bytes %u\n", dst, src, bytes);
        std::abort();
    }
    asyncq.push_back(Async{0, dst, src, bytes, bar, tick + delay()});
    ++progress;
}
void cp_async_8(void* dst, const void* src) {
    if (((uintptr_t)dst & 7u) || ((uintptr_t)src & 7u)) { std::fprintf(stderr, "[cuemu] cp.async 8: misaligned\n"); std::abort(); }
    const unsigned long long due = tick + delay();
    cur->async_due = std::max(cur->async_due, due);
    asyncq.push_back(Async{1, dst, src, 8, nullptr, due});
    ++progress;
}
void cp_async_n(void* dst, const void* src, int bytes) {
    const uintptr_t m = (uintptr_t)bytes - 1;
    if ((bytes != 4 && bytes != 8 && bytes != 16) || ((uintptr_t)dst & m) || ((uintptr_t)src & m)) {
        std::fprintf(stderr, "[cuemu] cp.async %d: bad size or misaligned\n", bytes); std::abort();
    }
    const unsigned long long due = tick + delay();
    cur->cpa_open = std::max(cur->cpa_open, due + 1);             // lands in run_async once tick >= due, visible from tick due + 1
    asyncq.push_back(Async{1, dst, src, (uint32_t)bytes, nullptr, due});
    ++progress;
}
void cp_async_16(void* dst, const void* src) { cp_async_n(dst, src, 16); }
void cp_async_commit() { cur->cpa_groups.push_back(cur->cpa_open); cur->cpa_open = 0; ++progress; }
void cp_async_wait(int n) {
    // all committed groups of this thread except the n most recent are complete
    for (;;) {
        const int must = (int)cur->cpa_groups.size() - n;
        bool pending = false;
        for (int g = 0; g < must; ++g) pending |= cur->cpa_groups[g] > tick;
        if (!pending) break;
        yield("cp.async.wait_group", cur);
    }
    const int must = (int)cur->cpa_groups.size() - n;
    if (must > 0) cur->cpa_groups.erase(cur->cpa_groups.begin(), cur->cpa_groups.begin() + must);
}
// cp.async.bulk.global.shared::cta (TMA store) + bulk_group commit / wait_group.read
void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    if ((bytes & 15u) || ((uintptr_t)gdst & 15u) || ((uintptr_t)ssrc & 15u)) {
        std::fprintf(stderr, "[cuemu] cp.async.bulk store needs 16-byte aligned addresses and size: dst %p src %p bytes %u\n", gdst, ssrc, bytes);
        std::abort();
    }
    Async a{3, gdst, ssrc, bytes, nullptr, tick + 1 + delay()};
    a.fiber = (int)(cur - &fibers[0]);
    a.group = cur->groups_retired + (int)cur->store_groups.size() - 1;
    ++cur->store_groups.back();
    asyncq.push_back(a);
    ++progress;
}
void bulk_commit() { cur->store_groups.push_back(0); ++progress; }
void bulk_wait_read(int n) {
    // all committed groups except the n most recent have finished reading shared memory
    for (;;) {
        const int committed = (int)cur->store_groups.size() - 1;
        bool pending = false;
        for (int g = 0; g < committed - n; ++g) pending |= cur->store_groups[g] > 0;
        if (!pending) break;
        yield("cp.async.bulk.wait_group.read", cur);
    }
    while (cur->store_groups.size() > 1 && cur->store_groups.front() == 0 && (int)cur->store_groups.size() - 1 > n) {
        cur->store_groups.erase(cur->store_groups.begin());
        ++cur->groups_retired;
    }
}
void cp_async_arrive_noinc(uint64_t* bar) {
    asyncq.push_back(Async{2, nullptr, nullptr, 0, bar, std::max(cur->async_due, tick)});
    ++progress;
}

}  // namespace cuemu

// ---- runtime API ---------------------------------------------------------------------------------------------------
const char* cudaGetErrorString(cudaError_t e) {
    switch (e) {
        case cudaSuccess: return "no error";
        case cudaErrorInvalidValue: return "invalid argument";
        case cudaErrorMemoryAllocation: return "out of memory";
        case cudaErrorInvalidConfiguration: return "invalid configuration argument";
        default: return "operation not supported (emulator)";
    }
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { const char* e = getenv("CUEMU_DEVICES"); *n = e ? std::max(1, atoi(e)) : 1; return cudaSuccess; }
static thread_local int t_device = 0;                 // like CUDA: the current device is a property of the host thread
cudaError_t cudaSetDevice(int d) { t_device = d; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = t_device; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; std::strcpy(p->name, "cuemu"); return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { const char* e = getenv("CUEMU_SMS"); *v = e ? std::max(1, atoi(e)) : 2; return cudaSuccess; }
cudaError_t cudaMallocBytes(void** p, size_t n) {
    *p = nullptr;
    if (posix_memalign(p, 256, std::max<size_t>(n, 256)) != 0) return cudaErrorMemoryAllocation;
    std::memset(*p, 0xCD, n);                          // device memory starts as garbage
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMallocBytes(p, n); }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t st) { cuemu::flush_stream(st); std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st) { cuemu::flush_stream(st); std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t st) {
    cuemu::flush_stream(st);
    for (size_t r = 0; r < height; ++r) std::memcpy((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms* p, cudaStream_t st) {
    cuemu::flush_stream(st);
    const cudaPitchedPtr &S = p->srcPtr, &D = p->dstPtr;
    for (size_t z = 0; z < p->extent.depth; ++z)
        for (size_t y = 0; y < p->extent.height; ++y) {
            const char* s = (const char*)S.ptr + ((p->srcPos.z + z) * S.ysize + (p->srcPos.y + y)) * S.pitch + p->srcPos.x;
            char* d = (char*)D.ptr + ((p->dstPos.z + z) * D.ysize + (p->dstPos.y + y)) * D.pitch + p->dstPos.x;
            std::memcpy(d, s, p->extent.width);
        }
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t st) { cuemu::flush_stream(st); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned) {
    std::lock_guard<std::recursive_mutex> lock(cuemu::g_emu_mutex);
    *st = (cudaStream_t)std::malloc(8);
    cuemu::created_streams[*st];
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> lock(cuemu::g_emu_mutex);
    cuemu::flush_stream(st); cuemu::created_streams.erase(st); std::free(st); return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* ev, unsigned) { *ev = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t ev) { std::lock_guard<std::recursive_mutex> lock(cuemu::g_emu_mutex); cuemu::event_stream.erase(ev); std::free(ev); return cudaSuccess; }
// an event recorded in a deferring stream completes only when that stream's queue has run: whoever waits for it runs it
cudaError_t cudaEventRecord(cudaEvent_t ev, cudaStream_t st) { std::lock_guard<std::recursive_mutex> lock(cuemu::g_emu_mutex); cuemu::event_stream[ev] = st; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t ev, unsigned) {
    std::lock_guard<std::recursive_mutex> lock(cuemu::g_emu_mutex);
    auto it = cuemu::event_stream.find(ev);
    if (it != cuemu::event_stream.end()) cuemu::flush_stream(it->second);
    return cudaSuccess;
}
// "IPC" inside one process: the handle carries the pointer.  Lets a test stand up several ranks of a peer-linked hierarchy
// in one address space, so the kernels' in-place reads of another rank's slab run under the emulator too.
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof(*p)); return *p ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
